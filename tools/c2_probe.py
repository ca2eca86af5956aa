"""Where does config C2 (P1 reaction-diffusion, per-point coefficients, hybrid mesh; 7.6 % of the HBM roofline in round 1) lose
its time?  Times the P1 numeric pass at C2's size on three meshes (triangles, quadrilaterals, C2's hybrid mesh) with three
coefficient kinds (constant, per cell, per quadrature point) through every algorithm that accepts the combination (auto = the
vertex-fan kernel on triangles with constant coefficients, else the item kernel; gather = the item kernel; atomic = FP64
atomics after a zero-fill): CUDA events on the ctx stream, warm-up 3, 10 steps each.  The differences along each axis separate the cost of the quadrature loop
(per-point vs. per-cell coefficients on triangles), of the non-affine geometry (quadrilaterals vs. triangles) and of the mixed
warps (hybrid vs. the weighted mean of the two pure meshes).
usage: c2_probe.py [n]   (default n = 1633: 4.0e6 cells on the hybrid mesh);  prints one JSON line (kept under profiles/)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lehrfempp_b200 as lf  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1633
STEPS, WARMUP, STRIDE = 10, 3, 4
ctx = lf.Context(0)
out = {"n": n, "env": {k: v for k, v in os.environ.items() if k.startswith("LFGPU_")}, "meshes": {}}
# the hybrid mesh has ~1.5 n^2 cells; the pure meshes get about as many
nt = int(round(n * (1.5 / 2.0) ** 0.5))
nq = int(round(n * 1.5 ** 0.5))
for mname, make in (("tria", lambda: ctx.mesh_tp_tria(nt, nt)), ("quad", lambda: ctx.mesh_tp_quad(nq, nq)),
                    ("hybrid", lambda: ctx.mesh_hybrid(n, 0.2, 12345))):
    mesh = make()
    pat = mesh.dofmap_lagrange(1).symbolic(major=lf.ROW_MAJOR)
    vals = ctx.empty(pat.nnz)
    xy = mesh.qp_coords(1, STRIDE).to_host().reshape(mesh.n_cells, STRIDE, 2)
    r2 = xy[..., 0] ** 2 + xy[..., 1] ** 2
    a_qp, g_qp = ctx.to_device(1.0 + r2), ctx.to_device(1.0 / (1.0 + r2))
    a_cell, g_cell = ctx.to_device(np.ascontiguousarray(1.0 + r2[:, 0])), ctx.to_device(np.ascontiguousarray(1.0 / (1.0 + r2[:, 0])))
    coeffs = {
        "const": (lf.Coeff.const(1.5), lf.Coeff.const(0.5)),
        "per_cell": (lf.Coeff.per_cell(a_cell), lf.Coeff.per_cell(g_cell)),
        "per_qp": (lf.Coeff.per_qp(a_qp, STRIDE), lf.Coeff.per_qp(g_qp, STRIDE)),
    }
    rec = {"cells": mesh.n_cells, "tria": mesh.n_tria, "quad": mesh.n_quad, "nodes": mesh.n_nodes, "nnz": pat.nnz, "ms": {}}
    for cname, (alpha, gamma) in coeffs.items():
        ref = None
        for aname, algo in (("auto", lf.ALGO_AUTO), ("gather", lf.ALGO_GATHER), ("atomic", lf.ALGO_ATOMIC)):
            try:
                for _ in range(WARMUP):
                    pat.assemble_reaction_diffusion(1, alpha, gamma, out=vals, algo=algo)
                e0, e1 = ctx.event(), ctx.event()
                ctx.record(e0)
                for _ in range(STEPS):
                    pat.assemble_reaction_diffusion(1, alpha, gamma, out=vals, algo=algo)
                ctx.record(e1)
                ms = ctx.elapsed_ms(e0, e1) / STEPS
            except lf.LfgpuError as e:
                rec["ms"][f"{cname}/{aname}"] = {"error": str(e)[:120]}
                continue
            v = vals.to_host()  # ALGO_ATOMIC zero-fills the array itself (beta = 0), its time includes that memset
            if ref is None:
                ref = v.copy()
            # algorithmic bytes: cell corners (16 B ids) + node coordinates once + coefficients + values once
            nco = {"const": 0, "per_cell": 2, "per_qp": 2 * (3 * mesh.n_tria + 4 * mesh.n_quad) / max(mesh.n_cells, 1)}[cname]
            alg = 16.0 * mesh.n_cells + 16.0 * mesh.n_nodes + 8.0 * nco * mesh.n_cells + 8.0 * pat.nnz
            rec["ms"][f"{cname}/{aname}"] = {"ms": ms, "cells_per_s": mesh.n_cells / ms * 1e3, "alg_GBs": alg / ms / 1e6,
                                             "rel_diff_vs_first": float(np.abs(v - ref).max() / np.abs(ref).max())}
    out["meshes"][mname] = rec
    del pat, vals, mesh
print(json.dumps(out))
