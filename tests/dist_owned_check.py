"""Multi-GPU parity check of the distributed-ownership mode (lehrfempp_b200.distributed.OwnedAssembler), launched with torchrun
(one process per GPU, NCCL only for the verdict): every rank extracts ITS sub-problem, runs its own symbolic and numeric pass,
and compares the rows it owns with the oracle's matrix -- pattern bit-exact after local -> global, values within 1e-12; the
owned rows of all ranks must cover every row exactly once.  Prints one line per case and rank, DIST_OWNED_OK at the end."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import threading
    import time

    def _watchdog():
        time.sleep(280)
        os._exit(3)
    threading.Thread(target=_watchdog, daemon=True).start()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import lehrfempp_b200 as lf
    from lehrfempp_b200.distributed import OwnedAssembler
    from oracle import lfo
    ctx = lf.Context(local)
    ok = True
    cases = [("tria", 1), ("tria", 2), ("hybrid", 1), ("hybrid", 3), ("refined", 3), ("tria_big", 1), ("refined_big", 3)]
    for kind, degree in cases:
        if kind == "tria":
            gm, om = ctx.mesh_tp_tria(37, 29), lfo.Mesh.tp_tria(37, 29)
        elif kind == "tria_big":
            gm, om = ctx.mesh_tp_tria(300, 280), lfo.Mesh.tp_tria(300, 280)
        elif kind.startswith("refined"):
            # BASELINE config 4's mesh family: 2 x 2 x 2 builder mesh, regular refinements with the reference's numbering
            levels = 3 if kind == "refined" else 6  # 512 / 32 768 cells
            gm, om = ctx.mesh_tp_tria(2, 2), lfo.Mesh.tp_tria(2, 2)
            for _ in range(levels):
                gm, om = gm.refine_regular(), om.refine_regular()
        else:
            gm, om = ctx.mesh_hybrid(24, 0.2, 12345), lfo.Mesh.hybrid(24, 0.2, 12345)
        dm = gm.dofmap_lagrange(degree)
        n_dofs = dm.num_dofs
        asm = OwnedAssembler(ctx, gm, dm, degree, rank, world)
        del dm, gm
        a, g = lf.Coeff.const(1.5), lf.Coeff.const(0.5)
        values = asm.assemble(a, g)
        values = asm.assemble(a, g, values)
        ctx.synchronize()
        o_outer, o_inner, o_vals, _, _ = om.assemble_rd(degree, lfo.coeff.const(1.5), lfo.coeff.const(0.5), csr=True)
        l_outer, l_inner = asm.pattern.download()
        h = values.to_host()
        rows_g, rows_l = asm.owned_rows_global()
        l2g = asm.sub.l2g_dofs()
        lens = l_outer[rows_l + 1] - l_outer[rows_l]
        same_len = np.array_equal(lens, o_outer[rows_g + 1] - o_outer[rows_g])
        idx_l = np.repeat(l_outer[rows_l] - np.cumsum(lens) + lens, lens) + np.arange(lens.sum())
        idx_g = np.repeat(o_outer[rows_g] - np.cumsum(lens) + lens, lens) + np.arange(lens.sum())
        pattern_ok = same_len and np.array_equal(l2g[l_inner[idx_l]], o_inner[idx_g])
        err = np.abs(h[idx_l] - o_vals[idx_g]).max() / np.abs(o_vals).max() if pattern_ok else np.inf
        cover = torch.zeros(n_dofs, dtype=torch.int32, device="cuda")
        cover[torch.as_tensor(rows_g.astype(np.int64), device="cuda")] = 1
        dist.all_reduce(cover)
        full = bool((cover == 1).all().item())
        print("rank %d %s P%d: pattern %s err %.2e owned rows %d of %d local (%d cells, %d stored values) cover %s" % (
            rank, kind, degree, "bit-exact" if pattern_ok else "DIFFERS", err, rows_g.size, asm.pattern.rows, asm.mesh.n_cells,
            asm.pattern.nnz, full), flush=True)
        ok = ok and pattern_ok and err <= 1e-12 and full
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_OWNED_OK" if t.item() == 1 else "DIST_OWNED_FAILED", flush=True)
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
