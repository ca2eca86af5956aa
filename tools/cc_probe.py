"""Times the P2 / P3 numeric pass on a mesh that carries per-cell corner coordinates (cell_coords: every cell's corners moved by
1e-9 of its size so that they are not bitwise the node positions) with the row kernels (ALGO_FAN, CC instantiations) and with the
item kernel (ALGO_GATHER), and the same mesh without cell_coords.  usage: cc_probe.py <degree> <n>.  One JSON line."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lehrfempp_b200 as lf  # noqa: E402

degree = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ctx = lf.Context(0)
plain = ctx.mesh_tp_tria(n, n)
d = plain.download(topology=True)
xy, cn = d["node_coords"], d["cell_nodes"]
cc = np.zeros((len(cn), 4, 2))
cc[:, :3] = xy[cn[:, :3]]
cc[:, :3] += 1e-9 / n * (np.random.default_rng(1).random((len(cn), 3, 2)) - 0.5)
mesh = ctx.mesh_upload(xy, cn, cc)
mesh.build_topology(d["edge_nodes"])
alpha, gamma = lf.Coeff.const(1.0), lf.Coeff.const(1.0 if degree == 3 else 0.0)
out = {"degree": degree, "cells": int(mesh.n_cells)}
res = {}
for label, m in (("cell_coords", mesh), ("node_coords", plain)):
    pat = m.dofmap_lagrange(degree).symbolic(major=lf.ROW_MAJOR)
    vals = ctx.empty(pat.nnz)
    for name, algo in (("rows", lf.ALGO_FAN), ("items", lf.ALGO_GATHER)):
        for _ in range(3):
            pat.assemble_reaction_diffusion(degree, alpha, gamma, out=vals, algo=algo)
        e0, e1 = ctx.event(), ctx.event()
        ctx.record(e0)
        for _ in range(10):
            pat.assemble_reaction_diffusion(degree, alpha, gamma, out=vals, algo=algo)
        ctx.record(e1)
        out["%s/%s_ms" % (label, name)] = ctx.elapsed_ms(e0, e1) / 10
        res[label, name] = vals.to_host()
    out["%s/rel_diff_rows_items" % label] = float(np.abs(res[label, "rows"] - res[label, "items"]).max() / np.abs(res[label, "items"]).max())
print(json.dumps(out))
