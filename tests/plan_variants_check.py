"""Run by tests/test_gpu_zz_plan_variants.py under different LFGPU_P2_COMPACT / LFGPU_L2_HINTS / LFGPU_P2_BULK settings
(the switches are read once per process): the P2 / P3 row kernels with every plan format and copy-out variant against the oracle
and the generic kernel.  Prints PLAN_VARIANTS_OK."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lehrfempp_b200 as lf  # noqa: E402
from oracle import lfo  # noqa: E402

ctx = lf.Context(0)
COEFFS = [
    (lf.Coeff.const(1.0), lf.Coeff.const(0.0), lfo.coeff.const(1.0), lfo.coeff.const(0.0)),
    (lf.Coeff.const2x2([[2.0, 0.5], [-0.25, 1.5]]), lf.Coeff.const(1.25), lfo.coeff.const2x2([[2.0, 0.5], [-0.25, 1.5]]), lfo.coeff.const(1.25)),
]


def check(name, gm, om, degree, oracle=True):
    for major, csr in ((lf.ROW_MAJOR, True), (lf.COL_MAJOR, False)):
        pat = gm.dofmap_lagrange(degree).symbolic(major=major)
        for ga, gg, oa, og in COEFFS:
            v = pat.assemble_reaction_diffusion(degree, ga, gg, algo=lf.ALGO_FAN).to_host()
            ref = pat.assemble_reaction_diffusion(degree, ga, gg, algo=lf.ALGO_GATHER).to_host()
            assert np.abs(v - ref).max() <= 1e-13 * np.abs(ref).max(), (name, degree, major)
            if oracle:
                o = om.assemble_rd(degree, oa, og, csr=csr)
                assert np.abs(v - o[2]).max() <= 1e-12 * np.abs(o[2]).max(), (name, degree, major)
            # accumulate on top of what is there (assembler.h:84-88)
            out = ctx.to_device(ref.copy())
            v2 = pat.assemble_reaction_diffusion(degree, ga, gg, algo=lf.ALGO_FAN, out=out, beta=1.0).to_host()
            assert np.abs(v2 - 2.0 * ref).max() <= 1e-13 * np.abs(ref).max(), (name, degree, major, "beta")
        # a row range (one GPU's share in the replicated modes)
        n = pat.rows  # square matrices: rows == cols == outer size
        part = pat.assemble_reaction_diffusion_range(degree, COEFFS[0][0], COEFFS[0][1], n // 3, 2 * n // 3 - n // 3, algo=lf.ALGO_FAN).to_host()
        full = pat.assemble_reaction_diffusion(degree, COEFFS[0][0], COEFFS[0][1], algo=lf.ALGO_GATHER).to_host()
        outer, _ = pat.download()
        lo, hi = outer[n // 3], outer[2 * n // 3]
        assert np.abs(part[lo:hi] - full[lo:hi]).max() <= 1e-13 * np.abs(full).max(), (name, degree, "rows")
    print(name, "P%d ok" % degree)


for degree in (2, 3):
    check("tp_tria 37x23", ctx.mesh_tp_tria(37, 23, 0.25, -0.5, 1.75, 0.5), lfo.Mesh.tp_tria(37, 23, 0.25, -0.5, 1.75, 0.5), degree)
    check("refined 5x4 x3", ctx.mesh_tp_tria(5, 4).refine_regular().refine_regular().refine_regular(),
          lfo.Mesh.tp_tria(5, 4).refine_regular().refine_regular().refine_regular(), degree)
    # above one wave of CTAs (prefetch branch), against the generic kernel only
    check("tp_tria 600x500", ctx.mesh_tp_tria(600, 500), None, degree, oracle=False)
# node numbers shuffled: neighbours are further than 32767 apart, the plan must stay in its full format
om = lfo.Mesh.tp_tria(210, 190)
ex = om.export()
rng = np.random.default_rng(5)
perm = rng.permutation(len(ex["node_coords"]))            # new number of old node i
inv = np.argsort(perm)
xy = ex["node_coords"][inv]
cn = ex["cell_nodes"].copy()
mask = cn != 0xFFFFFFFF
cn[mask] = perm[cn[mask]].astype(np.uint32)
gm = ctx.mesh_upload(xy, cn)
gm.build_topology(None)
for degree in (2, 3):
    check("shuffled 210x190", gm, None, degree, oracle=False)
print("PLAN_VARIANTS_OK")
