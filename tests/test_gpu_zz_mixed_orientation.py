"""Row kernels (P1 fan, P2, P3) on triangle meshes whose cells are listed with MIXED orientation and with the row's vertex at
any local position -- legal input for the reference (its test meshes do it) and the case in which the two cells of an edge
run along it in the SAME direction.  Wheels and a Delaunay triangulation with a quarter of the cells flipped, against the
oracle.  (The plan logic of the P3 and general P2 kernels is checked for these meshes on the CPU; this covers the device-only
plan kernels of P1 / P2.)  Added after the round's GPU minutes were spent; sorts last on purpose."""
import numpy as np
import pytest

from oracle import lfo

pytestmark = pytest.mark.gpu
NIL = 0xFFFFFFFF


@pytest.fixture(scope="module")
def lf():
    import lehrfempp_b200 as lf
    return lf


@pytest.fixture(scope="module")
def ctx(lf):
    c = lf.Context(0)
    yield c
    c.close()


def meshes():
    for m in (5, 6, 7):
        ang = 2 * np.pi * (np.arange(m) + 0.1 * np.sin(np.arange(m))) / m
        xy = np.vstack([[0.05, -0.03], np.stack([np.cos(ang), 0.8 * np.sin(ang)], axis=1)])
        rows = []
        for k in range(m):
            a, b = 1 + k, 1 + (k + 1) % m
            rows.append([[0, a, b], [b, 0, a], [b, a, 0]][k % 3] + [NIL])
        yield "wheel %d" % m, xy, np.array(rows, dtype=np.uint32)
    from scipy.spatial import Delaunay
    pts = np.random.default_rng(11).random((150, 2))
    tri = Delaunay(pts).simplices.astype(np.uint32)
    flip = np.arange(len(tri)) % 4 == 1
    tri[flip] = tri[flip][:, [0, 2, 1]]
    yield "delaunay 150, mixed orientation", pts, np.hstack([tri, np.full((len(tri), 1), NIL, np.uint32)])
    # a structured mesh with every third cell flipped: valence-6 rings, i.e. the static P2 / P3 vertex kernels
    ex = lfo.Mesh.tp_tria(8, 7).export()
    cn = ex["cell_nodes"].copy()
    sel = np.arange(len(cn)) % 3 == 2
    cn[sel, 1], cn[sel, 2] = ex["cell_nodes"][sel, 2], ex["cell_nodes"][sel, 1]
    yield "tp_tria 8x7, every third cell clockwise", ex["node_coords"], cn


@pytest.mark.parametrize("degree", [1, 2, 3])
def test_row_kernels_on_mixed_orientation_meshes(ctx, lf, degree):
    for name, xy, cn in meshes():
        om = lfo.Mesh.from_arrays(xy, cn)
        gm = ctx.mesh_upload(xy, cn)
        dm = gm.dofmap_lagrange(degree)
        od, onl = om.cell_dofs(degree)
        gd, gnl = dm.download()
        assert np.array_equal(gd, od) and np.array_equal(gnl, onl), name
        for major, csr in ((lf.ROW_MAJOR, True), (lf.COL_MAJOR, False)):
            pat = dm.symbolic(major=major)
            o = om.assemble_rd(degree, lfo.coeff.const2x2([[2.0, 0.5], [-0.25, 1.5]]), lfo.coeff.const(1.25), csr=csr)
            outer, inner = pat.download()
            assert np.array_equal(outer, o[0]) and np.array_equal(inner, o[1]), name
            v = pat.assemble_reaction_diffusion(degree, lf.Coeff.const2x2([[2.0, 0.5], [-0.25, 1.5]]), lf.Coeff.const(1.25)).to_host()
            assert np.abs(v - o[2]).max() <= 1e-12 * np.abs(o[2]).max(), (name, major)
            g = pat.assemble_reaction_diffusion(degree, lf.Coeff.const2x2([[2.0, 0.5], [-0.25, 1.5]]), lf.Coeff.const(1.25),
                                                algo=lf.ALGO_GATHER).to_host()
            assert np.abs(v - g).max() <= 1e-13 * np.abs(g).max(), (name, major)
