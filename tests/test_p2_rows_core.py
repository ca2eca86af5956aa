"""CPU check of the general-valence P2 vertex rows (lehrfempp_b200/csrc/rows_p2_core.h): the product's plan and row functions,
compiled with g++ (tests/cpp/p2_rows_emul.cc), against the oracle on Gmsh meshes, wheels with 3..8 spokes, a Delaunay
triangulation of random points (every ring length 3..8 occurs), a structured mesh (valence 6 through the general path) and a
refined mesh; both storage orders, all constant coefficients."""
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
    subprocess.check_call(["make", "-C", os.path.join(HERE, "cpp"), "-s", "libp2emul.so"])
    L = C.CDLL(os.path.join(HERE, "cpp", "libp2emul.so"))
    L.p2_vertex_rows_emulate.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int, C.c_double] + [C.c_void_p] * 8
    return L


def reference_tensors():
    pts, w = lfo.quad_rule(3, 4)  # default rule of FeLagrangeO2: degree 2 * 2
    phi, grad, _ = lfo.eval_fe(2, 3, pts)
    gx, gy = grad[:, 0::2], grad[:, 1::2]
    k = {"k00": (gx * w) @ gx.T, "k01": (gx * w) @ gy.T, "k10": (gy * w) @ gx.T, "k11": (gy * w) @ gy.T, "km": (phi * w) @ phi.T}
    return {n: np.ascontiguousarray(v) for n, v in k.items()}


def meshes():
    for name in ("circle_first_order.msh", "circle_first_order_v4.msh"):
        xy, en, cn, _ = OracleReader(os.path.join(HERE, "golden", "msh", name)).arrays()
        yield name, lfo.Mesh.from_arrays(xy, cn, edge_nodes=en)
    # wheels: one interior vertex with m = 3..8 cells (every other ring length than the Gmsh meshes' 5 and 6)
    for m in range(3, 9):
        ang = 2 * np.pi * (np.arange(m) + 0.1 * np.sin(np.arange(m))) / m
        xy = np.vstack([[0.05, -0.03], np.stack([np.cos(ang), 0.8 * np.sin(ang)], axis=1)])
        # the node is local vertex 0, 1 or 2, and every third cell is listed clockwise (both traversal directions of the plan)
        rows = []
        for k in range(m):
            a, b = 1 + k, 1 + (k + 1) % m
            rows.append([[0, a, b], [b, 0, a], [b, a, 0]][k % 3] + [0xFFFFFFFF])
        cn = np.array(rows, dtype=np.uint32)
        yield "wheel %d" % m, lfo.Mesh.from_arrays(xy, cn)
    # Delaunay triangulation of seeded random points: valences 3..9 mixed
    from scipy.spatial import Delaunay
    pts = np.random.default_rng(11).random((150, 2))
    tri = Delaunay(pts).simplices.astype(np.uint32)
    yield "delaunay 150", lfo.Mesh.from_arrays(pts, np.hstack([tri, np.full((len(tri), 1), 0xFFFFFFFF, np.uint32)]))
    yield "tp_tria 6x5", lfo.Mesh.tp_tria(6, 5, 0.25, -0.5, 1.75, 0.5)
    yield "tp_tria 2x2 refined twice", lfo.Mesh.tp_tria(2, 2).refine_regular().refine_regular()


COEFFS = [("laplace", 1.0, None, 0.0), ("reaction-diffusion", 2.5, None, 0.75), ("tensor", None, [[2.0, 0.5], [-0.25, 1.5]], 1.25)]


@pytest.mark.parametrize("csr", [True, False], ids=["csr", "csc"])
@pytest.mark.parametrize("coeff", COEFFS, ids=[c[0] for c in COEFFS])
def test_general_vertex_rows_match_oracle(emul, coeff, csr):
    _, a_scalar, a_tensor, gamma = coeff
    K = reference_tensors()
    seen_valences = set()
    for name, om in meshes():
        ex = om.export()
        dofs, nl = om.cell_dofs(2)
        assert np.all(nl == 6)
        oalpha = lfo.coeff.const(a_scalar) if a_tensor is None else lfo.coeff.const2x2(a_tensor)
        outer, inner, vals, _, _ = om.assemble_rd(2, oalpha, lfo.coeff.const(gamma), csr=csr)
        n_dofs = outer.size - 1
        alpha4 = np.array([a_scalar, 0.0, 0.0, a_scalar]) if a_tensor is None else (np.array(a_tensor).T if csr else np.array(a_tensor)).ravel().copy()
        d32 = np.ascontiguousarray(dofs, dtype=np.int32)
        cn = np.ascontiguousarray(ex["cell_nodes"], dtype=np.uint32)
        xy = np.ascontiguousarray(ex["node_coords"], dtype=np.float64)
        out = np.zeros(vals.size)
        regular = np.zeros(n_dofs, np.uint8)
        hist = np.zeros(9, np.int64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = emul.p2_vertex_rows_emulate(om.n_nodes, om.n_cells, p(cn), p(xy), d32.shape[1], p(d32), n_dofs, p(outer), p(inner), p(alpha4),
                                         int(a_tensor is not None), gamma, p(K["k00"]), p(K["k01"]), p(K["k10"]), p(K["k11"]), p(K["km"]),
                                         p(out), p(regular), p(hist))
        assert rc == 0
        # every interior vertex with 3..8 cells must have been taken, with the ring length of its valence
        bd = om.boundary_edges().astype(bool)
        bd_nodes = np.zeros(om.n_nodes, bool)
        bd_nodes[ex["edge_nodes"][bd].ravel()] = True
        valence = np.bincount(ex["cell_nodes"][:, :3].ravel(), minlength=om.n_nodes)
        want = np.bincount(valence[~bd_nodes & (valence >= 3) & (valence <= 8)], minlength=9)[:9]
        assert np.array_equal(hist, want), (name, hist, want)
        seen_valences |= set(np.flatnonzero(hist))
        row_of = np.repeat(np.arange(n_dofs), np.diff(outer))
        sel = regular[row_of].astype(bool)
        assert sel.any() and not np.isnan(out[sel]).any() and np.isnan(out[~sel]).all()
        err = np.abs(out[sel] - vals[sel]).max() / np.abs(vals).max()
        assert err <= TOL, (name, err)
    assert seen_valences >= {3, 4, 5, 6, 7, 8}
