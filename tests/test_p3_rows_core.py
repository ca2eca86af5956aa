"""CPU check of the P3 row kernels' shared host/device core (lehrfempp_b200/csrc/rows_p3_core.h).

tests/cpp/p3_rows_emul.cc compiles the product's plan and row functions with g++ and runs them on the arrays the symbolic pass
would hold; here the resulting matrix rows are compared with the oracle (the restated reference) on structured, refined and
unstructured triangle meshes, for both storage orders and all constant-coefficient kinds.  What cannot be checked here --
the CUDA wrappers around these functions -- is covered by tests/test_gpu_p3_rows.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import lfo
from oracle.lfo_gmsh import GmshReader as OracleReader

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-12


@pytest.fixture(scope="module")
def emul():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "cpp"), "-s", "libp3emul.so"])
    L = C.CDLL(os.path.join(HERE, "cpp", "libp3emul.so"))
    L.p3_rows_emulate.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int, C.c_double] + [C.c_void_p] * 8
    L.p3_rows_emulate_cell_coords.argtypes = [C.c_void_p]
    return L


def reference_tensors():
    """Khat^{ij}[a][b] = sum_k w_k d_i phi_a d_j phi_b, Mhat[a][b] = sum_k w_k phi_a phi_b for FeLagrangeO3Tria and the default
    rule make_QuadRule(kTria, 6) (loc_comp_ellbvp.h:227-228)."""
    pts, w = lfo.quad_rule(3, 6)
    phi, grad, _ = lfo.eval_fe(3, 3, pts)
    gx, gy = grad[:, 0::2], grad[:, 1::2]
    k = {"k00": (gx * w) @ gx.T, "k01": (gx * w) @ gy.T, "k10": (gy * w) @ gx.T, "k11": (gy * w) @ gy.T, "km": (phi * w) @ phi.T}
    return {n: np.ascontiguousarray(v) for n, v in k.items()}


def meshes():
    yield "tp_tria 7x6", lfo.Mesh.tp_tria(7, 6, 0.25, -0.5, 1.75, 0.5)
    yield "tp_tria 3x3 refined twice", lfo.Mesh.tp_tria(3, 3).refine_regular().refine_regular()
    xy, en, cn, _ = OracleReader(os.path.join(HERE, "golden", "msh", "circle_first_order.msh")).arrays()
    yield "gmsh circle", lfo.Mesh.from_arrays(xy, cn, edge_nodes=en)
    # wheels with the hub as local vertex 0, 1 or 2 and every third cell listed clockwise: edge directions disagree with the
    # cells' local edge directions in every combination (the reversal of the two edge dofs, dofhandler.cc:245-260)
    for m in range(3, 9):
        ang = 2 * np.pi * (np.arange(m) + 0.1 * np.sin(np.arange(m))) / m
        wxy = np.vstack([[0.05, -0.03], np.stack([np.cos(ang), 0.8 * np.sin(ang)], axis=1)])
        rows = []
        for k in range(m):
            a, b = 1 + k, 1 + (k + 1) % m
            rows.append([[0, a, b], [b, 0, a], [b, a, 0]][k % 3] + [0xFFFFFFFF])
        yield "wheel %d" % m, lfo.Mesh.from_arrays(wxy, np.array(rows, dtype=np.uint32))
    from scipy.spatial import Delaunay
    pts = np.random.default_rng(11).random((150, 2))
    tri = Delaunay(pts).simplices.astype(np.uint32)
    flip = np.arange(len(tri)) % 4 == 1  # a quarter of the cells clockwise
    tri[flip] = tri[flip][:, [0, 2, 1]]
    yield "delaunay 150, mixed orientation", lfo.Mesh.from_arrays(pts, np.hstack([tri, np.full((len(tri), 1), 0xFFFFFFFF, np.uint32)]))


COEFFS = [
    ("laplace", 1.0, None, 0.0),
    ("reaction-diffusion", 2.5, None, 0.75),
    ("tensor", None, [[2.0, 0.5], [-0.25, 1.5]], 1.25),
]


@pytest.mark.parametrize("general", [False, True], ids=["ring6", "ring3to8"])
@pytest.mark.parametrize("csr", [True, False], ids=["csr", "csc"])
@pytest.mark.parametrize("coeff", COEFFS, ids=[c[0] for c in COEFFS])
def test_rows_match_oracle(emul, coeff, csr, general):
    """general = vertex rows through vertex_plan_general / vertex_row_general (closed rings of 3..8 cells, the opt-in kernel for
    unstructured meshes) instead of the valence-6 functions"""
    _, a_scalar, a_tensor, gamma = coeff
    K = reference_tensors()
    emul.p3_rows_emulate_general(1 if general else 0)
    seen = set()
    for name, om in meshes():
        ex = om.export()
        assert om.n_quad == 0
        dofs, nl = om.cell_dofs(3)
        assert np.all(nl == 10)
        oalpha = lfo.coeff.const(a_scalar) if a_tensor is None else lfo.coeff.const2x2(a_tensor)
        outer, inner, vals, _, _ = om.assemble_rd(3, oalpha, lfo.coeff.const(gamma), csr=csr)
        n_dofs = outer.size - 1
        if a_tensor is None:
            alpha4 = np.array([a_scalar, 0.0, 0.0, a_scalar])
        else:
            A = np.array(a_tensor)
            alpha4 = (A.T if csr else A).ravel().copy()  # transposed for row-major output, as assemble.cu passes it
        d32 = np.ascontiguousarray(dofs, dtype=np.int32)
        cn = np.ascontiguousarray(ex["cell_nodes"], dtype=np.uint32)
        xy = np.ascontiguousarray(ex["node_coords"], dtype=np.float64)
        out = np.zeros(vals.size)
        regular = np.zeros(n_dofs, np.uint8)
        counts = np.zeros(3, np.int64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = emul.p3_rows_emulate(om.n_nodes, om.n_cells, p(cn), p(xy), d32.shape[1], p(d32), n_dofs, p(outer), p(inner), p(alpha4),
                                  int(a_tensor is not None), gamma, p(K["k00"]), p(K["k01"]), p(K["k10"]), p(K["k11"]), p(K["km"]),
                                  p(out), p(regular), p(counts))
        assert rc == 0
        # every cell row, every edge with two cells and every closed ring of six cells must have been taken
        n_cells, n_edges = om.n_cells, om.n_edges
        bd = om.boundary_edges().astype(bool)
        assert counts[2] == n_cells, name
        assert counts[1] == 2 * int((~bd).sum()), name
        valence = np.bincount(ex["cell_nodes"][:, :3].ravel(), minlength=om.n_nodes)
        bd_nodes = np.zeros(om.n_nodes, bool)
        bd_nodes[ex["edge_nodes"][bd].ravel()] = True
        if general:
            assert counts[0] == int(((valence >= 3) & (valence <= 8) & ~bd_nodes).sum()), name
            seen |= set(valence[(valence >= 3) & (valence <= 8) & ~bd_nodes])
        else:
            assert counts[0] == int(((valence == 6) & ~bd_nodes).sum()), name
        if name.startswith("tp_tria"):
            assert counts[0] > 0
        # values of the rows taken: the oracle's, within the bar of the path
        row_of = np.repeat(np.arange(n_dofs), np.diff(outer))
        sel = regular[row_of].astype(bool)
        assert not np.isnan(out[sel]).any() and np.isnan(out[~sel]).all()
        err = np.abs(out[sel] - vals[sel]).max() / np.abs(vals).max()
        assert err <= TOL, (name, err)
    emul.p3_rows_emulate_general(0)
    if general:
        assert seen >= {3, 4, 5, 6, 7, 8}


@pytest.mark.parametrize("csr", [True, False], ids=["csr", "csc"])
@pytest.mark.parametrize("coeff", COEFFS, ids=[c[0] for c in COEFFS])
def test_rows_with_cell_corners_match_oracle(emul, coeff, csr):
    """Meshes whose cells carry their own corner coordinates (the reference's Geometry objects; after RefineRegular they differ
    from the node positions in the last bits): vertex_plan / edge_plan hand out (cell, corner) words, vertex_row_cv / edge_row2
    compute every cell from ITS corners.  The corners are moved by up to 5 % of the cell size here so that a kernel reading node
    positions (or the wrong cell's corners) would be off by orders of magnitude more than the bar."""
    _, a_scalar, a_tensor, gamma = coeff
    K = reference_tensors()
    rng = np.random.default_rng(3)
    for name, om0 in meshes():
        ex = om0.export()
        cn = np.ascontiguousarray(ex["cell_nodes"], dtype=np.uint32)
        xy = np.ascontiguousarray(ex["node_coords"], dtype=np.float64)
        corners = xy[cn[:, :3]]                                            # [n_cells][3][2]
        e1, e2 = corners[:, 1] - corners[:, 0], corners[:, 2] - corners[:, 0]
        size = np.sqrt(np.abs(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]))
        cc = np.zeros((len(cn), 4, 2))
        cc[:, :3] = corners + 0.05 * size[:, None, None] * (rng.random(corners.shape) - 0.5)
        om = lfo.Mesh.from_arrays(xy, cn, cell_coords=cc, cell_geo=np.ones(len(cn), np.uint8), edge_nodes=ex["edge_nodes"])
        dofs, nl = om.cell_dofs(3)
        oalpha = lfo.coeff.const(a_scalar) if a_tensor is None else lfo.coeff.const2x2(a_tensor)
        outer, inner, vals, _, _ = om.assemble_rd(3, oalpha, lfo.coeff.const(gamma), csr=csr)
        plain = om0.assemble_rd(3, oalpha, lfo.coeff.const(gamma), csr=csr)[2]
        assert np.abs(plain - vals).max() > 1e-4 * np.abs(vals).max()      # the perturbation is visible
        n_dofs = outer.size - 1
        if a_tensor is None:
            alpha4 = np.array([a_scalar, 0.0, 0.0, a_scalar])
        else:
            A = np.array(a_tensor)
            alpha4 = (A.T if csr else A).ravel().copy()
        d32 = np.ascontiguousarray(dofs, dtype=np.int32)
        out = np.zeros(vals.size)
        regular = np.zeros(n_dofs, np.uint8)
        counts = np.zeros(3, np.int64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        emul.p3_rows_emulate_general(0)
        emul.p3_rows_emulate_cell_coords(p(cc))
        try:
            rc = emul.p3_rows_emulate(om.n_nodes, om.n_cells, p(cn), p(xy), d32.shape[1], p(d32), n_dofs, p(outer), p(inner), p(alpha4),
                                      int(a_tensor is not None), gamma, p(K["k00"]), p(K["k01"]), p(K["k10"]), p(K["k11"]), p(K["km"]),
                                      p(out), p(regular), p(counts))
        finally:
            emul.p3_rows_emulate_cell_coords(None)
        assert rc == 0
        bd = om.boundary_edges().astype(bool)
        assert counts[2] == om.n_cells and counts[1] == 2 * int((~bd).sum()), name
        row_of = np.repeat(np.arange(n_dofs), np.diff(outer))
        sel = regular[row_of].astype(bool)
        err = np.abs(out[sel] - vals[sel]).max() / np.abs(vals).max()
        assert err <= TOL, (name, err)
