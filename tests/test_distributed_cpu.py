"""World-size-2 CPU tests (gloo) of the multi-GPU host logic: Morton partition, row ownership, message layout and the
owner-adds exchange.  The per-rank numeric kernels are replaced by the oracle (assembly restricted to the rank's cells)."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, degree, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lehrfempp_b200.distributed import PartitionPlan, morton_partition
        from oracle import lfo
        om = lfo.Mesh.tp_tria(13, 9) if kind == "tria" else lfo.Mesh.hybrid(10, 0.2, 7)
        ex = om.export()
        dofs, nl = om.cell_dofs(degree)
        a, g = lfo.coeff.const(1.5), lfo.coeff.const(0.5)
        outer, inner, vals, shape, _ = om.assemble_rd(degree, a, g, csr=True)
        n_rows = shape[0]
        # gather lists as the symbolic pass builds them: items sorted by dof, cells ascending
        cells = np.repeat(np.arange(om.n_cells), dofs.shape[1])
        d = dofs.ravel()
        keep = d >= 0
        order = np.argsort(d[keep], kind="stable")
        adj_cell = cells[keep][order]
        adj_ptr = np.searchsorted(d[keep][order], np.arange(n_rows + 1))
        cn = ex["cell_nodes"].astype(np.int64)
        cn[cn == 0xFFFFFFFF] = -1
        part = morton_partition(torch.from_numpy(ex["node_coords"]), torch.from_numpy(cn), world)
        counts = np.bincount(part.numpy(), minlength=world)
        assert counts.max() - counts.min() <= world  # balanced contiguous Morton ranges
        plan = PartitionPlan(part, torch.from_numpy(adj_ptr), torch.from_numpy(adj_cell), torch.from_numpy(outer.astype(np.int64)), rank, world)
        # ownership is a partition of the rows, identical on every rank
        owners = [torch.zeros_like(plan.owner) for _ in range(world)]
        dist.all_gather(owners, plan.owner)
        assert all(torch.equal(o, plan.owner) for o in owners)
        assert (plan.owner >= 0).all()
        assert plan.interior_rows.numel() + plan.iface_rows.numel() > 0
        # my partial assembly: only my cells are active (what the GPU kernels compute with the activity mask)
        po, pi, pv, _, _ = om.assemble_rd(degree, a, g, csr=True, active=(part.numpy() == rank).astype(np.uint8))
        P = sp.csr_matrix((pv, pi, po), shape=shape)
        rows = np.repeat(np.arange(n_rows), np.diff(outer))
        partial = np.asarray(P[rows, inner]).ravel()
        v = torch.from_numpy(partial.copy())
        plan.exchange_torch(v)
        mine = np.zeros(vals.size, dtype=bool)
        for r in plan.owned_rows.numpy():
            mine[outer[r]:outer[r + 1]] = True
        err = np.abs(v.numpy()[mine] - vals[mine]).max() / np.abs(vals).max()
        # rows touched by me only are complete without any exchange
        inter = np.zeros(vals.size, dtype=bool)
        for r in plan.interior_rows.numpy():
            inter[outer[r]:outer[r + 1]] = True
        err_i = np.abs(partial[inter] - vals[inter]).max() / np.abs(vals).max() if inter.any() else 0.0
        n_owned = torch.tensor([plan.owned_rows.numel()])
        dist.all_reduce(n_owned)
        q.put((rank, float(err), float(err_i), int(n_owned.item()), n_rows, plan.n_send, plan.n_recv))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,degree", [("tria", 1), ("tria", 2), ("hybrid", 1), ("hybrid", 3)])
def test_partitioned_exchange_world2(kind, degree):
    world = 2
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, kind, degree, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, err_i, n_owned, n_rows, n_send, n_recv in res:
        assert err <= 1e-13 and err_i <= 1e-13
        assert n_owned == n_rows  # every row has exactly one owner
    # what rank 1 sends is what rank 0 receives (the owner is the lowest rank touching a row)
    by = {r[0]: r for r in res}
    assert by[1][5] == by[0][6] and by[0][5] == 0 and by[1][6] == 0 and by[1][5] > 0
