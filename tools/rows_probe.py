"""Times the numeric pass of degree p on a TP-triangle mesh with the row kernels (ALGO_FAN) and the item kernel (ALGO_GATHER):
CUDA events on the ctx stream, warm-up 3, 10 steps each.  usage: rows_probe.py <degree> <n> [rows|items|''] [refinement levels]
Prints one JSON line (kept under profiles/)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lehrfempp_b200 as lf  # noqa: E402

degree = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1448
only = sys.argv[3] if len(sys.argv) > 3 else ""
levels = int(sys.argv[4]) if len(sys.argv) > 4 else 0   # RefineRegular steps on top of the n x n builder mesh (config C4's numbering)
ctx = lf.Context(0)
mesh = ctx.mesh_tp_tria(n, n)
for _ in range(levels):
    mesh = mesh.refine_regular()
pat = mesh.dofmap_lagrange(degree).symbolic(major=lf.ROW_MAJOR)
vals = ctx.empty(pat.nnz)
out = {"degree": degree, "cells": mesh.n_cells, "nnz": pat.nnz, "refined": levels}
alpha, gamma = lf.Coeff.const(1.0), lf.Coeff.const(1.0 if degree == 3 else 0.0)  # config C4: stiffness + mass
nldof = {1: 3, 2: 6, 3: 10}[degree]
res = {}
for name, algo in (("rows", lf.ALGO_FAN), ("items", lf.ALGO_GATHER)):
    if only and name != only:
        continue
    for _ in range(3):
        pat.assemble_reaction_diffusion(degree, alpha, gamma, out=vals, algo=algo)
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(10):
        pat.assemble_reaction_diffusion(degree, alpha, gamma, out=vals, algo=algo)
    ctx.record(e1)
    ms = ctx.elapsed_ms(e0, e1) / 10
    res[name] = vals.to_host()
    alg_bytes = 4 * nldof * mesh.n_cells + 16 * mesh.n_nodes + 8 * pat.nnz
    out[name] = {"ms": ms, "cells_per_s": mesh.n_cells / ms * 1e3, "alg_GBs": alg_bytes / ms / 1e6}
if only != "items" and degree >= 2:
    # the row classes one by one through the row-range call: vertex rows, edge-dof rows, (P3) cell rows -- each range also
    # runs the generic kernel on its irregular rows (boundary), so the parts add up to a little more than the full pass
    n_rows = pat.download()[0].size - 1
    nn = mesh.n_nodes
    cuts = [0, nn, n_rows - mesh.n_cells, n_rows] if degree == 3 else [0, nn, n_rows]
    parts = {}
    for name, r0, r1 in zip(("vertex_rows", "edge_rows", "cell_rows"), cuts[:-1], cuts[1:]):
        for _ in range(3):
            pat.assemble_reaction_diffusion_range(degree, alpha, gamma, r0, r1 - r0, out=vals, algo=lf.ALGO_FAN)
        e0, e1 = ctx.event(), ctx.event()
        ctx.record(e0)
        for _ in range(10):
            pat.assemble_reaction_diffusion_range(degree, alpha, gamma, r0, r1 - r0, out=vals, algo=lf.ALGO_FAN)
        ctx.record(e1)
        parts[name] = {"rows": int(r1 - r0), "ms": ctx.elapsed_ms(e0, e1) / 10}
    out["parts"] = parts
    out["env"] = {k: v for k, v in os.environ.items() if k.startswith("LFGPU_")}
if not only:
    out["rel_diff"] = float(np.abs(res["rows"] - res["items"]).max() / np.abs(res["items"]).max())
print(json.dumps(out))
