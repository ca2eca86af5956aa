"""Multi-GPU parity check, launched by tests/test_gpu_multi.py with torchrun (one process per GPU, NCCL).

Every rank assembles its part with DistributedAssembler; the owned rows of all ranks together must reproduce the
oracle's matrix (pattern bit-exact, values within 1e-12)."""
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
        time.sleep(240)
        os._exit(3)
    threading.Thread(target=_watchdog, daemon=True).start()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import lehrfempp_b200 as lf
    from lehrfempp_b200.distributed import DistributedAssembler
    from oracle import lfo
    ctx = lf.Context(local)
    ok = True
    for kind, degree in (("tria", 1), ("tria", 2), ("hybrid", 1), ("hybrid", 3), ("tria_big", 1)):
        if kind == "tria":
            gm, om = ctx.mesh_tp_tria(37, 29), lfo.Mesh.tp_tria(37, 29)
        elif kind == "tria_big":
            gm, om = ctx.mesh_tp_tria(300, 280), lfo.Mesh.tp_tria(300, 280)
        else:
            gm, om = ctx.mesh_hybrid(24, 0.2, 12345), lfo.Mesh.hybrid(24, 0.2, 12345)
        dm = gm.dofmap_lagrange(degree)
        pat = dm.symbolic(major=lf.ROW_MAJOR)
        asm = DistributedAssembler(ctx, gm, pat, degree, mode=os.environ.get("LFGPU_DIST_MODE", "exchange"))
        values = ctx.zeros(pat.nnz)
        a, g = lf.Coeff.const(1.5), lf.Coeff.const(0.5)
        for _ in range(2):  # twice: buffers and events are reused
            asm.assemble(a, g, values)
        if kind == "tria_big" and os.environ.get("LFGPU_TEST_GRAPH", "0") == "1":  # opt-in: captured CUDA graph of the step
            asm.capture(a, g, values)
            ctx.check(ctx.L.lfgpu_memset(ctx.h, values.ptr, 0, values.nbytes))
            asm.replay()
            asm.replay()
        ctx.synchronize()
        torch.cuda.synchronize()
        o_outer, o_inner, o_vals, _, _ = om.assemble_rd(degree, lfo.coeff.const(1.5), lfo.coeff.const(0.5), csr=True)
        outer, inner = pat.download()
        assert np.array_equal(outer, o_outer) and np.array_equal(inner, o_inner)
        h = values.to_host()
        mine = asm.owned_value_mask(outer)
        err = np.abs(h[mine] - o_vals[mine]).max() / np.abs(o_vals).max()
        cover = torch.tensor(mine.astype(np.int32), device="cuda")
        dist.all_reduce(cover)
        full = bool((cover == 1).all().item())
        p = asm.plan
        print("rank %d %s P%d: err %.2e owned rows %d interior %d iface %d send %d recv %d cover %s" % (
            rank, kind, degree, err, p.owned_rows.numel(), p.interior_rows.numel(), p.iface_rows.numel(), p.n_send, p.n_recv, full), flush=True)
        ok = ok and err <= 1e-12 and full
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK_OK" if t.item() == 1 else "DIST_CHECK_FAILED", flush=True)
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
