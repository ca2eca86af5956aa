"""Times the load vector at the headline size (P1, 1.0e8 triangles) with the atomic and the gather kernel (CUDA events, warm-up
3, 10 steps) and the one-time cost of the gather plan.  Prints one JSON line (kept under profiles/)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lehrfempp_b200 as lf  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 7071
degree = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = lf.Context(0)
mesh = ctx.mesh_tp_tria(n, n)
dm = mesh.dofmap_lagrange(degree)
vec = ctx.zeros(dm.num_dofs)
source = sys.argv[3] if len(sys.argv) > 3 else "const"   # const | per_qp (tabulated at the quadrature points, as MeshFunctionGlobal is)
if source == "per_qp":
    nq = max(lf.default_quad_rule(3, 2 * degree).weights.size, lf.default_quad_rule(4, 2 * degree).weights.size)  # table stride
    xy = mesh.qp_coords(degree, nq).to_host().reshape(mesh.n_cells, nq, 2)
    f = lf.Coeff.per_qp(ctx.to_device(np.ascontiguousarray(1.0 + xy[..., 0] * xy[..., 1])), nq)
else:
    f = lf.Coeff.const(1.0)
out = {"cells": mesh.n_cells, "dofs": dm.num_dofs, "degree": degree, "source": source}
t = time.time()
dm.assemble_load(degree, f, out=vec, algo=lf.ALGO_GATHER)
ctx.synchronize()
out["gather_first_call_s"] = time.time() - t
res = {}
# auto = the vertex-ring kernel (P1, constant source) or the two-pass kernels
for name, algo in (("gather", lf.ALGO_GATHER), ("atomic", lf.ALGO_ATOMIC), ("auto", lf.ALGO_AUTO)):
    for _ in range(3):
        dm.assemble_load(degree, f, out=vec, algo=algo)
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(10):
        dm.assemble_load(degree, f, out=vec, algo=algo)
    ctx.record(e1)
    out[name + "_ms"] = ctx.elapsed_ms(e0, e1) / 10
    res[name] = vec.to_host()
out["rel_diff"] = float(np.abs(res["gather"] - res["atomic"]).max() / np.abs(res["atomic"]).max())
out["rel_diff_auto"] = float(np.abs(res["gather"] - res["auto"]).max() / np.abs(res["gather"]).max())
print(json.dumps(out))
