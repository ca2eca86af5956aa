"""Run with LFGPU_P2_GENERAL=1 LFGPU_P3_GENERAL=1 (tests/test_gpu_zz_p2_general.py does): P2 and P3 assembly on unstructured
triangle meshes through the row kernels with the general-valence vertex plans, against the oracle.  Prints P2_GENERAL_OK."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lehrfempp_b200 as lf  # noqa: E402
from oracle import lfo  # noqa: E402
from oracle.lfo_gmsh import GmshReader as OracleReader  # noqa: E402

assert os.environ.get("LFGPU_P2_GENERAL") == "1" and os.environ.get("LFGPU_P3_GENERAL") == "1"
ctx = lf.Context(0)
NIL = 0xFFFFFFFF


def meshes():
    path = os.path.join(ROOT, "tests", "golden", "msh", "circle_first_order.msh")
    xy, en, cn, _ = OracleReader(path).arrays()
    yield "gmsh circle", xy, cn, en
    from scipy.spatial import Delaunay
    for seed, n in ((11, 150), (12, 4000)):
        pts = np.random.default_rng(seed).random((n, 2))
        tri = Delaunay(pts).simplices
        # drop the slivers of the convex hull (the device rejects degenerate cells like the reference's geometry classes)
        a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
        area2 = np.abs((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1]))
        longest2 = np.maximum.reduce([((b - a) ** 2).sum(1), ((c - b) ** 2).sum(1), ((a - c) ** 2).sum(1)])
        tri = tri[area2 > 0.05 * longest2].astype(np.uint32)
        used = np.unique(tri)
        remap = np.full(len(pts), -1)
        remap[used] = np.arange(len(used))
        yield "delaunay %d" % n, pts[used], np.hstack([remap[tri].astype(np.uint32), np.full((len(tri), 1), NIL, np.uint32)]), None
    om = lfo.Mesh.tp_tria(9, 7)
    ex = om.export()
    yield "tp_tria uploaded", ex["node_coords"], ex["cell_nodes"], ex["edge_nodes"]


for name, xy, cn, en in meshes():
    om = lfo.Mesh.from_arrays(xy, cn, edge_nodes=en)
    gm = ctx.mesh_upload(xy, cn)
    gm.build_topology(en)
    for degree in (2, 3):
        dm = gm.dofmap_lagrange(degree)
        for major, csr in ((lf.ROW_MAJOR, True), (lf.COL_MAJOR, False)):
            pat = dm.symbolic(major=major)
            for ga, gg, oa, og in ((lf.Coeff.const(1.0), lf.Coeff.const(0.0), lfo.coeff.const(1.0), lfo.coeff.const(0.0)),
                                   (lf.Coeff.const2x2([[2.0, 0.5], [-0.25, 1.5]]), lf.Coeff.const(1.25),
                                    lfo.coeff.const2x2([[2.0, 0.5], [-0.25, 1.5]]), lfo.coeff.const(1.25))):
                o = om.assemble_rd(degree, oa, og, csr=csr)
                outer, inner = pat.download()
                assert np.array_equal(outer, o[0]) and np.array_equal(inner, o[1]), name
                try:
                    v = pat.assemble_reaction_diffusion(degree, ga, gg, algo=lf.ALGO_FAN).to_host()
                except lf.LfgpuError as e:
                    assert e.code == -7, e  # too few regular rows on this mesh
                    print(name, "P%d: row kernels declined" % degree)
                    continue
                err = np.abs(v - o[2]).max() / np.abs(o[2]).max()
                ref = pat.assemble_reaction_diffusion(degree, ga, gg, algo=lf.ALGO_GATHER).to_host()
                assert err <= 1e-12 and np.abs(v - ref).max() <= 1e-13 * np.abs(ref).max(), (name, degree, err)
                print(name, "P%d major %d err %.2e" % (degree, major, err))
print("P2_GENERAL_OK")
