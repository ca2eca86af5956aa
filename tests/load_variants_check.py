"""Run by tests/test_gpu_zz_plan_variants.py under LFGPU_LOAD_TWOPASS / LFGPU_LOAD_FAN settings (read once per
process): the load vector of LFGPU_ALGO_AUTO against the oracle and the gather kernel.  Prints LOAD_VARIANTS_OK."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lehrfempp_b200 as lf  # noqa: E402
from oracle import lfo  # noqa: E402

ctx = lf.Context(0)
for om, gm in ((lfo.Mesh.tp_tria(29, 17, 0.25, -0.5, 1.75, 0.5), ctx.mesh_tp_tria(29, 17, 0.25, -0.5, 1.75, 0.5)),
               (lfo.Mesh.hybrid(9, 0.2, 3), ctx.mesh_hybrid(9, 0.2, 3))):
    for degree in (1, 2):
        dm = gm.dofmap_lagrange(degree)
        stride = max(lf.default_quad_rule(3, 2 * degree).weights.size, lf.default_quad_rule(4, 2 * degree).weights.size)
        xy = gm.qp_coords(degree, stride).to_host().reshape(om.n_cells, stride, 2)
        tab = np.ascontiguousarray(1.0 + xy[..., 0] * xy[..., 1] + np.sin(xy[..., 1]))
        for oc, gc in ((lfo.coeff.table(tab), lf.Coeff.per_qp(ctx.to_device(tab), stride)), (lfo.coeff.const(2.0), lf.Coeff.const(2.0)),
                       (lfo.coeff.table(np.ascontiguousarray(tab[:, 0])), lf.Coeff.per_cell(ctx.to_device(np.ascontiguousarray(tab[:, 0]))))):
            ov, _ = om.assemble_load(degree, oc)
            v = dm.assemble_load(degree, gc)
            assert np.abs(v.to_host() - ov).max() <= 1e-12 * np.abs(ov).max()
            g = dm.assemble_load(degree, gc, algo=lf.ALGO_GATHER).to_host()
            assert np.abs(v.to_host() - g).max() <= 1e-13 * np.abs(g).max()
            dm.assemble_load(degree, gc, beta=1.0, out=v)
            assert np.abs(v.to_host() - 2 * ov).max() <= 1e-12 * np.abs(ov).max()
big = ctx.mesh_tp_tria(700, 500)
dm = big.dofmap_lagrange(1)
xy = big.qp_coords(1, 4).to_host().reshape(big.n_cells, 4, 2)
gc = lf.Coeff.per_qp(ctx.to_device(np.ascontiguousarray(1.0 + xy[..., 0] * xy[..., 1])), 4)
v = dm.assemble_load(1, gc).to_host()
g = dm.assemble_load(1, gc, algo=lf.ALGO_GATHER).to_host()
assert np.abs(v - g).max() <= 1e-13 * np.abs(g).max() and abs(v.sum() - 1.25) <= 1e-9
print("LOAD_VARIANTS_OK")
