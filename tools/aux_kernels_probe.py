"""Timing of the kernels around the hot path at the headline size (1.0e8 triangles, P1, CSR): boundary edge terms,
Dirichlet elimination (in place), SpMV and CG iterations.  CUDA events on the context stream; one JSON line.

  python tools/aux_kernels_probe.py [n]      (default n = 7071)
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lehrfempp_b200 as lf  # noqa: E402


def timed(ctx, fn, reps):
    fn()
    ctx.synchronize()
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(reps):
        fn()
    ctx.record(e1)
    return ctx.elapsed_ms(e0, e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 7071
    ctx = lf.Context(0)
    mesh = ctx.mesh_tp_tria(n, n)
    dm = mesh.dofmap_lagrange(1)
    pat = dm.symbolic(major=lf.ROW_MAJOR)
    one, zero = lf.Coeff.const(1.0), lf.Coeff.const(0.0)
    vals = pat.assemble_reaction_diffusion(1, one, one)
    rhs = dm.assemble_load(1, one)
    N, nnz = dm.num_dofs, pat.nnz
    out = {"workload": "P1 on TP-triangle mesh n=%d: %d cells, %d dofs, %d nnz" % (n, mesh.n_cells, N, nnz)}
    bd = mesh.boundary_edges()
    out["boundary_edges"] = int(bd.to_host().sum())
    out["edge_mass_boundary_ms"] = timed(ctx, lambda: pat.assemble_edge_mass(dm, 1, one, vals, active_edges=bd), 5)
    out["edge_mass_all_edges_ms"] = timed(ctx, lambda: pat.assemble_edge_mass(dm, 1, one, vals), 3)
    out["edge_load_all_edges_ms"] = timed(ctx, lambda: dm.assemble_edge_load(1, one, out=rhs), 3)
    # Dirichlet data on the boundary dofs
    mark = dm.assemble_edge_load(1, one, active_edges=bd).to_host()
    fixed = ctx.to_device((mark > 0).astype(np.uint8))
    xhat = ctx.zeros(N)
    vals = pat.assemble_reaction_diffusion(1, one, one, out=vals)
    t = timed(ctx, lambda: pat.fix_flagged_solution_components(vals, rhs, fixed, xhat), 5)
    out["fix_in_place_ms"] = t
    out["fix_in_place_GBps"] = (nnz * (8 + 4 + 8) + N * 30) / t / 1e6  # values r+w, inner, rhs r+w, flags, outer
    x = ctx.to_device(np.random.default_rng(0).standard_normal(N))
    y = ctx.empty(N)
    t = timed(ctx, lambda: pat.spmv(vals, x, out=y), 20)
    out["spmv_ms"] = t
    out["spmv_GBps"] = (nnz * 12 + N * 20) / t / 1e6  # values + inner, outer + x + y
    rhs = dm.assemble_load(1, one)
    pat.fix_flagged_solution_components(vals, rhs, fixed, xhat)
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    sol, iters, res = pat.cg_solve(vals, rhs, rel_tol=0.0, max_iter=25)
    ctx.record(e1)
    out["cg_ms_per_iteration"] = ctx.elapsed_ms(e0, e1) / max(iters, 1)
    out["cg_iterations"] = iters
    out["cg_rel_residual_after"] = res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
