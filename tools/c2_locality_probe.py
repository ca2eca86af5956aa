"""Config C2 (hybrid P1, coefficients tabulated per quadrature point) on (a) the hybrid builder's numbering (nodes row by row, cells
column by column: the 32 rows of a warp read their cells' coefficient records from 32 different lines) and (b) the same mesh with
the CELLS renumbered row by row (tables permuted with them).  How much of the kernel's time is that gather pattern?
usage: c2_locality_probe.py [n]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lehrfempp_b200 as lf  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1633
ctx = lf.Context(0)
builder = ctx.mesh_hybrid(n, 0.2, 12345)
d = builder.download()
xy, cn = d["node_coords"], d["cell_nodes"]
nv = np.where(cn[:, 3] == 0xFFFFFFFF, 3, 4)
first = xy[cn[:, 0]]
cen = first.copy()
for k in range(1, 4):
    sel = nv > k
    cen[sel] += xy[cn[sel, k]]
cen /= nv[:, None]
gi = np.clip((cen[:, 0] * n).astype(np.int64), 0, n - 1)
gj = np.clip((cen[:, 1] * n).astype(np.int64), 0, n - 1)
order = np.lexsort((np.arange(len(cn)), gi, gj))      # row by row, cells of one grid square together
rowmajor = ctx.mesh_upload(xy, np.ascontiguousarray(cn[order]))
out = {"n": n, "cells": int(builder.n_cells)}
for label, mesh in (("builder", builder), ("cells_row_major", rowmajor)):
    pat = mesh.dofmap_lagrange(1).symbolic(major=lf.ROW_MAJOR)
    q = mesh.qp_coords(1, 4).to_host().reshape(mesh.n_cells, 4, 2)
    r2 = q[..., 0] ** 2 + q[..., 1] ** 2
    alpha, gamma = lf.Coeff.per_qp(ctx.to_device(1.0 + r2), 4), lf.Coeff.per_qp(ctx.to_device(1.0 / (1.0 + r2)), 4)
    vals = ctx.empty(pat.nnz)
    res = {}
    for name, algo in (("rows", lf.ALGO_AUTO), ("generic", lf.ALGO_GATHER)):
        for _ in range(3):
            pat.assemble_reaction_diffusion(1, alpha, gamma, out=vals, algo=algo)
        e0, e1 = ctx.event(), ctx.event()
        ctx.record(e0)
        for _ in range(10):
            pat.assemble_reaction_diffusion(1, alpha, gamma, out=vals, algo=algo)
        ctx.record(e1)
        res[name + "_ms"] = ctx.elapsed_ms(e0, e1) / 10
        res[name + "_sum"] = float(vals.to_host().sum())
    out[label] = res
print(json.dumps(out))
