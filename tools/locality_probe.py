"""How much do the P2 / P3 edge-row kernels lose to the builder's numbering?  TPTriagMeshBuilder numbers nodes row by row but cells
and edges column by column, so the 32 edge rows of a warp gather their nodes from 32 different 128-byte lines.  This probe times
the row classes on (a) the builder's mesh and (b) the same mesh with the NODES renumbered column by column (cells unchanged, edges
numbered by the device in cell order), where a warp's gathers fall into a few lines.  usage: locality_probe.py <degree> <n>"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lehrfempp_b200 as lf  # noqa: E402

degree = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2828
ctx = lf.Context(0)
builder = ctx.mesh_tp_tria(n, n)
d = builder.download()
xy, cn = d["node_coords"], d["cell_nodes"]
ids = np.arange((n + 1) * (n + 1), dtype=np.int64)
i, j = ids // (n + 1), ids % (n + 1)          # builder: node = i * (n + 1) + j, i = mesh row
new_of_old = (j * (n + 1) + i).astype(np.uint32)   # column by column
xy2 = np.empty_like(xy)
xy2[new_of_old] = xy
cn2 = cn.copy()
m = cn2 != 0xFFFFFFFF
cn2[m] = new_of_old[cn2[m]]
colmajor = ctx.mesh_upload(xy2, cn2)
colmajor.build_topology(None)
alpha, gamma = lf.Coeff.const(1.0), lf.Coeff.const(1.0 if degree == 3 else 0.0)
out = {"degree": degree, "n": n, "cells": int(builder.n_cells)}
for label, mesh in (("builder", builder), ("nodes_column_major", colmajor)):
    pat = mesh.dofmap_lagrange(degree).symbolic(major=lf.ROW_MAJOR)
    vals = ctx.empty(pat.nnz)
    n_rows = pat.rows
    nn = mesh.n_nodes
    cuts = [0, nn, n_rows - mesh.n_cells, n_rows] if degree == 3 else [0, nn, n_rows]
    res = {}
    for name, r0, r1 in [("all", 0, n_rows)] + list(zip(("vertex_rows", "edge_rows", "cell_rows"), cuts[:-1], cuts[1:])):
        for _ in range(3):
            pat.assemble_reaction_diffusion_range(degree, alpha, gamma, r0, r1 - r0, out=vals, algo=lf.ALGO_FAN)
        e0, e1 = ctx.event(), ctx.event()
        ctx.record(e0)
        for _ in range(10):
            pat.assemble_reaction_diffusion_range(degree, alpha, gamma, r0, r1 - r0, out=vals, algo=lf.ALGO_FAN)
        ctx.record(e1)
        res[name] = ctx.elapsed_ms(e0, e1) / 10
    ref = pat.assemble_reaction_diffusion(degree, alpha, gamma, algo=lf.ALGO_GATHER).to_host()
    v = pat.assemble_reaction_diffusion(degree, alpha, gamma, algo=lf.ALGO_FAN).to_host()
    res["rel_diff_vs_generic"] = float(np.abs(v - ref).max() / np.abs(ref).max())
    out[label] = res
print(json.dumps(out))
