#!/usr/bin/env python3
"""bench.py -- cells assembled per second into CSR (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A step is one numeric pass of the hot path (lf::assemble::AssembleMatrixLocally with
ReactionDiffusionElementMatrixProvider) over the whole mesh of the workload.  Default workload = BASELINE.json
config 5 at its largest size: P1 Laplacian on the 7071 x 7071 x 2 structured triangle mesh (1.0e8 cells), CSR.
`value` is timed on the device with CUDA events, inputs resident in HBM; `e2e` goes through the same C-ABI calls with
the per-step input (node coordinates) coming from pinned host memory and the CSR values returned to pinned host
memory inside the timed region.  Rank 0 prints ONE JSON line.

--impl reference times the CPU restatement of the reference path (oracle/, single thread as the reference is
serial) on a bounded sample of the same workload family.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

AUTO_ROW_DEGREES = (1, 2, 3)  # degrees for which LFGPU_ALGO_AUTO runs the row kernels (lehrfempp_b200/csrc/assemble.cu)

WORKLOADS = {
    # name: (description, kind, n, degree)
    "c5_1e8": ("C5: P1 Laplacian (alpha=1, gamma=0), TP-triangle mesh n=7071 (1.0e8 cells), CSR", "tp_tria", 7071, 1),
    "c5_1e7": ("C5: P1 Laplacian, TP-triangle mesh n=2236 (1.0e7 cells), CSR", "tp_tria", 2236, 1),
    "c5_1e6": ("C5: P1 Laplacian, TP-triangle mesh n=707 (1.0e6 cells), CSR", "tp_tria", 707, 1),
    "c1": ("C1: P1 Laplacian, TP-triangle mesh 256x256x2 (131072 cells), CSR", "tp_tria", 256, 1),
    "c2": ("C2: P1 reaction-diffusion, alpha=1+|x|^2, gamma=1/(1+|x|^2) per quadrature point, hybrid tri/quad mesh n=1633 (4.0e6 cells), CSR", "hybrid", 1633, 1),
    "c3": ("C3: P2 Laplacian, TP-triangle mesh n=2828 (1.6e7 cells), CSR", "tp_tria", 2828, 2),
    "c4s": ("C4 (single-GPU size): P3 stiffness+mass, TP-triangle mesh n=1448 (4.2e6 cells), CSR", "tp_tria", 1448, 3),
    "c4": ("C4 (single-GPU size): P3 stiffness+mass on a MeshHierarchy-refined mesh: TP-triangle mesh n=181, 3 x RefineRegular "
           "(4.2e6 cells), CSR", "refined:3", 181, 3),
    # BASELINE config 4 names 3.2e7 triangles; P3 on them has 2.45e9 stored values, more than the int32 storage index of the
    # reference's Eigen::SparseMatrix (and of the replicated pattern here) can address -> LFGPU_ERR_OVERFLOW.  This is the
    # largest mesh of the family that fits (nnz 2.11e9); the full size needs the per-GPU row-block pattern.
    # not a BASELINE configuration: an unstructured mesh (Delaunay triangulation of seeded random points, hull slivers removed,
    # valences 3..10) for the row kernels on Gmsh-like input; LFGPU_P2_GENERAL=1 selects the general-valence P2 vertex kernel
    "u2": ("U2: P2 Laplacian on an unstructured mesh (Delaunay triangulation of 5.0e5 random points, about 1.0e6 triangles), CSR",
           "delaunay", 500000, 2),
    "c4_full": ("C4 at its configured size: P3 stiffness+mass on a MeshHierarchy-refined mesh: TP-triangle mesh 2x2 (8 cells), 11 x RefineRegular "
                "(3.36e7 cells, 2.5e9 stored values: needs the distributed pattern, i.e. --gpus >= 2), CSR", "refined:11", 2, 3),
    "c4_27m": ("C4: P3 stiffness+mass on a MeshHierarchy-refined mesh: TP-triangle mesh n=232, 4 x RefineRegular (2.76e7 cells, the "
               "largest of the family whose nnz fits the reference's int32 storage index), CSR", "refined:4", 232, 3),
}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 7:
                self.samples.append((time.time(), f))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sel = [f for (t, f) in self.samples if t0 <= t <= t1] or [f for (_, f) in self.samples]
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [float(f[0]) for f in sel if f[0].replace(".", "").isdigit()]
        mx = [float(f[1]) for f in sel if f[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for f in sel for i in range(4) if f[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sel)}


def cpu_reference_sample(n, repeats=1):
    """CPU restatement of the reference path, single thread: AssembleMatrixLocally -> COO, then makeSparse."""
    from oracle import lfo
    m = lfo.Mesh.tp_tria(n, n)
    best = None
    for _ in range(repeats):
        _, _, _, _, t = m.assemble_rd(1, lfo.coeff.const(1.0), lfo.coeff.const(0.0), csr=True)
        s = t["assemble_s"] + t["makesparse_s"]
        best = s if best is None else min(best, s)
    return m.n_cells, best


def cpu_all_cores_sample(n, max_threads=32):
    """the 'all cores' figure -- C independent copies of the serial reference path run concurrently (one mesh
    and one matrix per thread; the reference itself cannot use more than one core).  One pass per copy."""
    from oracle import lfo
    threads = max(1, min(os.cpu_count() or 1, max_threads))
    try:  # one copy holds about 1.2 GB at n = 707 (mesh objects, triplets, compressed matrix): stay well inside the free memory
        import psutil
        threads = max(1, min(threads, int(psutil.virtual_memory().available / 2.5e9)))
    except Exception:
        threads = min(threads, 8)
    meshes = [None] * threads
    start = [0.0] * threads
    done = [0.0] * threads
    gate = threading.Barrier(threads)

    def work(i):
        meshes[i] = lfo.Mesh.tp_tria(n, n)  # untimed, built concurrently (ctypes releases the GIL during the calls)
        gate.wait()
        start[i] = time.time()
        meshes[i].assemble_rd(1, lfo.coeff.const(1.0), lfo.coeff.const(0.0), csr=True)
        done[i] = time.time()

    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    wall = max(done) - min(start)
    return {"value": threads * meshes[0].n_cells / wall, "unit": "cells/s", "cores": threads,
            "what": "%d independent copies of the serial path (mesh n=%d each) run concurrently, wall %.2f s; labelled as such: the "
                    "reference has no threading" % (threads, n, wall)}


def run_reference(args, rank):
    if rank != 0:
        return
    import numpy as np  # noqa: F401
    from oracle import lfo
    # the largest mesh of config C5's family whose --steps + --warmup passes finish in about four minutes of wall time (measured:
    # 2.2 us per cell and pass at 1e7 cells incl. the copy-out of the result, 48 s to build that mesh): n = 1414 (4.0e6 cells) for
    # the driver's 20 + 5 passes, n = 707 (1.0e6, the smallest C5 size) for the default 50 + 5
    passes = args.warmup + args.steps
    n = next((c for c in (2236, 1414, 1000) if 2 * c * c * passes * 2.2e-6 <= 240.0), 707)
    n = int(os.environ.get("LFGPU_REF_N", n))  # the contract test of the script runs a small mesh
    t_build = time.time()
    m = lfo.Mesh.tp_tria(n, n)
    t_build = time.time() - t_build
    times = []
    for it in range(args.warmup + args.steps):
        _, _, _, _, t = m.assemble_rd(1, lfo.coeff.const(1.0), lfo.coeff.const(0.0), csr=True)
        if it >= args.warmup:
            times.append(t["assemble_s"] + t["makesparse_s"])
    sec = sum(times) / len(times)
    value = m.n_cells / sec
    sample = "P1 Laplacian on TP-triangle mesh n=%d (%d cells), AssembleMatrixLocally->COO + makeSparse per step" % (n, m.n_cells)
    all_cores = cpu_all_cores_sample(707)
    out = {
        "impl": "reference", "metric": "cells assembled/sec into CSR", "value": value, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][0], "sample": sample, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": 1, "kind": "port", "sample": sample,
                         "note": "CPU restatement of the reference path (oracle/); the reference is serial, so 1 thread is all it can use; "
                                 "host has %d cores; mesh construction (%.1f s) excluded" % (os.cpu_count(), t_build),
                         "all_cores": all_cores},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_result(out)


def print_result(obj):  # replaced in main() by a writer to the real stdout
    print(json.dumps(obj))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5_1e8", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="override the mesh parameter n (debugging)")
    ap.add_argument("--algo", default="auto", choices=["auto", "fan", "gather", "atomic"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short C2 / C3 / C4 lines of the default single-GPU run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 (NCCL prints its version banner
    # there) are sent to stderr for the duration of the run; emit() writes the result to the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())
    global print_result
    print_result = emit

    # watchdog: a rank that is stuck (e.g. in a collective whose peer died) must not keep the node busy
    def _watchdog(limit=float(os.environ.get("LFGPU_BENCH_LIMIT_S", "420"))):
        time.sleep(limit)
        sys.stderr.write("bench.py watchdog: rank %d exceeded %.0f s, aborting\n" % (rank, limit))
        sys.stderr.flush()
        os._exit(3)
    threading.Thread(target=_watchdog, daemon=True).start()

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np

    import lehrfempp_b200 as lf

    # torch is plumbing for N > 1 only (process group, NCCL); a single-GPU run does not pay its import
    torch = None
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))

    desc, kind, n, degree = WORKLOADS[args.workload]
    if args.n:
        n = args.n
        desc += " [n overridden to %d]" % n
    ctx = lf.Context(local_rank)
    algo = {"auto": lf.ALGO_AUTO, "fan": lf.ALGO_FAN, "gather": lf.ALGO_GATHER, "atomic": lf.ALGO_ATOMIC}[args.algo]
    kernel_name = kernel_of(args.workload, kind, degree, args.algo)

    # ---- setup (untimed, like mesh / DofHandler construction on the CPU side) ----------------------------------------------
    t_setup = time.time()
    mesh = build_mesh(ctx, np, kind, n)
    dm = mesh.dofmap_lagrange(degree)
    n_cells, n_nodes, n_dofs, n_tria, n_quad = mesh.n_cells, mesh.n_nodes, dm.num_dofs, mesh.n_tria, mesh.n_quad

    # N > 1.  Default "owned": distributed ownership -- Morton cell ranges, every rank keeps only its sub-problem (its cells + the
    # one-cell halo around the rows it owns, local indices), owner-computes, no collective in the data path.  The round-1 modes
    # on a replicated pattern stay selectable: owner_rows (contiguous row blocks), owner (Morton row lists), exchange (partial
    # interface rows to their owner by one NCCL all-to-all-v overlapped with the interior rows).
    dist_mode = os.environ.get("LFGPU_DIST_MODE", "owned") if world > 1 else None
    asm = None
    t_part = 0.0
    if dist_mode == "owned":
        from lehrfempp_b200.distributed import OwnedAssembler
        asm = OwnedAssembler(ctx, mesh, dm, degree, rank, world)
        t_part, t_sym = asm.partition_s, asm.symbolic_s
        del dm, mesh  # the global mesh and dof map are not needed any more
        mesh, pat = asm.mesh, asm.pattern
        t = torch.tensor([float(asm.owned_nnz)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        nnz_global = int(t.item())
    else:
        if args.workload == "c4_full":
            raise SystemExit("c4_full has 2.5e9 stored values: it needs the distributed pattern (--gpus >= 2, LFGPU_DIST_MODE=owned)")
        t_sym = time.time()
        pat = dm.symbolic(major=lf.ROW_MAJOR)
        ctx.synchronize()
        t_sym = time.time() - t_sym
        nnz_global = pat.nnz
    alpha, gamma, coef_bytes = make_coeffs(ctx, lf, args.workload, mesh, degree)  # per-cell tables live on the (sub-)mesh in use
    if dist_mode == "owned" and coef_bytes:
        t = torch.tensor([float(coef_bytes)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        coef_bytes = float(t.item())  # includes the halo cells (read twice): an upper bound of the algorithmic count
    values = ctx.empty(pat.nnz)
    t_setup = time.time() - t_setup

    if world > 1 and dist_mode != "owned":
        from lehrfempp_b200.distributed import DistributedAssembler
        t_part = time.time()
        asm = DistributedAssembler(ctx, mesh, pat, degree, mode=dist_mode)
        t_part = time.time() - t_part

    use_graph = asm is not None and dist_mode != "owned" and os.environ.get("LFGPU_BENCH_GRAPH", "0") == "1"
    if use_graph:
        asm.capture(alpha, gamma, values)  # the partitioned step (4 launches + 1 collective) as one CUDA graph

    def step():
        if use_graph:
            asm.replay()
        elif asm is not None:
            asm.assemble(alpha, gamma, values)
        else:
            pat.assemble_reaction_diffusion(degree, alpha, gamma, out=values, algo=algo)

    def barrier():
        ctx.synchronize()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    ev0, ev1 = ctx.event(), ctx.event()
    l0 = ctx.kernel_launches
    t_timed0 = time.time()
    ctx.record(ev0)
    for _ in range(args.steps):
        step()
    ctx.record(ev1)
    ms = ctx.elapsed_ms(ev0, ev1)
    barrier()
    t_timed1 = time.time()
    launches = ctx.kernel_launches - l0
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    # keep the GPU busy a little longer so that the clock sampler sees the kernel under load (not part of any number)
    # (the step contains a collective for N > 1, so every rank runs the same, pre-agreed number of extra steps)
    if ms < 1000.0:
        n_probe = int(min(2000, max(5, 1000.0 / max(ms_per_step, 1e-3))))
        for _ in range(n_probe):
            step()
        barrier()
        t_timed1 = time.time()

    # ---- size-independent witness that the matrix left in `values` at the FULL size is the right operator (not part of any
    # timing): one device SpMV with the constant vector -- constants are in the kernel of the stiffness part, 1^T M 1 = |Omega|.
    # N > 1: every rank checks the rows it owns; max / sum over the ranks.
    check = None
    try:
        if dist_mode == "owned":
            wmax, wsum = asm.witness(values)
        elif asm is not None:
            y = pat.spmv(values, ctx.to_device(np.ones(n_dofs))).to_host()
            own = asm.plan.owned_rows.cpu().numpy()
            wmax, wsum = (float(np.abs(y[own]).max()) if own.size else 0.0), float(y[own].sum())
            del y
        else:
            y = pat.spmv(values, ctx.to_device(np.ones(n_dofs))).to_host()
            wmax, wsum = float(np.abs(y).max()), float(y.sum())
            del y
        if dist is not None:
            t = torch.tensor([wmax], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wmax = float(t.item())
            t = torch.tensor([wsum], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            wsum = float(t.item())
        check = witness_record(args.workload, wmax, wsum, world)
    except Exception as e:  # a witness must never cost the bench line
        check = {"error": str(e)[:200]}

    # ---- end to end through the C ABI with host buffers ---------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        h_xy = ctx.pinned(2 * mesh.n_nodes)
        d = mesh.download()
        h_xy[:] = d["node_coords"].ravel()
        del d
        h2d_rank = 16 * mesh.n_nodes
        e2e_blocks = int(os.environ.get("LFGPU_E2E_BLOCKS", "16"))
        range_mode = False
        if asm is None or dist_mode == "owned":
            # one rank (or every rank on its sub-problem): pinned coordinates of the (sub-)mesh in, pinned CSR values out
            h_vals = ctx.pinned(pat.nnz)
            d2h_bytes = 8 * pat.nnz
        else:
            # replicated pattern: every rank returns the rows it owns: pack their segments, then one D2H copy
            pl = asm.plan
            outer_t = torch.as_tensor(pat.download()[0], device="cuda").to(torch.int64)
            own = pl.owned_rows.to(torch.int64)
            lens = outer_t[own + 1] - outer_t[own]
            own_off = (torch.cumsum(lens, 0) - lens).contiguous()
            n_own = int(lens.sum().item())
            own_buf = ctx.empty(max(n_own, 1))
            h_vals = ctx.pinned(max(n_own, 1))

            def d2h():
                ctx.check(ctx.L.lfgpu_rows_pack(ctx.h, pat.h, pl.owned_rows.data_ptr(), pl.owned_rows.numel(), own_off.data_ptr(),
                                                values.ptr, own_buf.ptr))
                ctx.check(ctx.L.lfgpu_memcpy_d2h(ctx.h, h_vals.ctypes.data, own_buf.ptr, 8 * n_own))
            d2h_bytes = 8 * n_own
            # owner_rows on the fan kernel: every rank runs the host-buffer call for ITS row block -- uploads only the coordinate
            # window its rows refer to and downloads only its rows (lfgpu_assemble_reaction_diffusion_host_range)
            range_mode = (asm.mode == "owner_rows" and getattr(asm, "_range_ok", False) and args.algo in ("auto", "fan") and degree == 1)
            if range_mode:
                inner_t = torch.as_tensor(pat.download()[1], device="cuda")
                seg = inner_t[int(outer_t[asm.row0].item()):int(outer_t[asm.row0 + asm.n_rows].item())]
                h2d_rank = 16 * (int(seg.max().item()) + 1 - int(seg.min().item())) if seg.numel() > 0 else 0
                del inner_t, seg
        e2e_steps = max(3, min(args.steps, 10))
        host_call = asm is None or dist_mode == "owned"

        def e2e_step():
            if range_mode:
                pat.assemble_reaction_diffusion_host_range(degree, alpha, gamma, h_xy, h_vals, asm.row0, asm.n_rows, out=values, algo=algo,
                                                           n_blocks=max(2, e2e_blocks // world))
                return
            if host_call:
                # the host-buffer C-ABI call: pinned coordinates in, pinned CSR values out; upload, kernel and download
                # are pipelined over row blocks inside the call (csrc/hostpipe.cu); returns when h_vals is complete
                pat.assemble_reaction_diffusion_host(degree, alpha, gamma, h_xy, h_vals, out=values, algo=algo, n_blocks=e2e_blocks)
                return
            mesh.update_node_coords(h_xy.reshape(-1, 2))   # H2D of this step's input (16 B per node)
            step()                                           # numeric pass
            d2h()                                            # D2H of the CSR values
            ctx.synchronize()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        if dist is not None:
            t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        h2d_b = h2d_rank
        if dist is not None:
            t = torch.tensor([float(d2h_bytes), float(h2d_rank)], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            d2h_bytes = int(t[0].item())
            h2d_b = int(t[1].item())
        e2e = {"value": n_cells / e2e_s, "unit": "cells/s", "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_bytes),
               "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "what": ("per step: one lfgpu_assemble_reaction_diffusion_host call = H2D node coordinates (pinned) -> kernel -> D2H CSR "
                        "values (pinned), pipelined over %d row blocks" % e2e_blocks) if asm is None else
                       ("per step and rank: one lfgpu_assemble_reaction_diffusion_host call on the rank's sub-problem = H2D of its node "
                        "coordinates (pinned) -> kernel -> D2H of its CSR rows (pinned; incl. the few halo rows), pipelined over %d row "
                        "blocks" % e2e_blocks) if dist_mode == "owned" else
                       ("per step and rank: one lfgpu_assemble_reaction_diffusion_host_range call = H2D of the coordinate window of the "
                        "rank's row block (pinned) -> kernel -> D2H of its CSR rows (pinned), pipelined") if range_mode else
                       "per step: H2D node coordinates (pinned) -> partitioned assembly -> D2H of the owned CSR rows (pinned)"}
    t_end = time.time()
    if rank == 0:
        sampler.stop()

    def teardown():
        # release the captured graph (it references NCCL kernels and the ctx stream) before the communicator goes away
        if asm is not None:
            barrier()
            asm.graph = None
            torch.cuda.synchronize()
        sys.stdout.flush()
        if dist is not None:
            dist.destroy_process_group()

    # per-rank facts the line reports for N > 1 (max over the ranks)
    rank_facts = None
    if dist is not None:
        t = torch.tensor([t_sym, t_part, float(pat.nnz), float(mesh.n_cells)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rank_facts = {"symbolic_pass_s_max": round(float(t[0].item()), 3), "partition_s_max": round(float(t[1].item()), 3),
                      "stored_values_per_rank_max": int(t[2].item()), "cells_per_rank_max": int(t[3].item())}

    if rank != 0:
        teardown()
        return

    peak, peak_src = measured_peak()
    # algorithmic bytes (SURVEY.md 8d / DESIGN.md): int32 connectivity + each vertex coordinate once + coefficients +
    # each stored value written once
    nsf_t, nsf_q = {1: (3, 4), 2: (6, 9), 3: (10, 16)}[degree]
    conn = 4.0 * (nsf_t * n_tria + nsf_q * n_quad)
    alg_bytes = conn + 16.0 * n_nodes + coef_bytes + 8.0 * nnz_global
    # per launch (= per rank for N > 1) the kernel covers 1/world of the rows
    alg_bytes_launch = alg_bytes / world
    achieved = alg_bytes_launch / (ms_per_step * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp)).get(args.workload + ":" + args.algo)
            if tj:
                traffic = tj["dram_bytes_per_cell"] * n_cells / world
        except Exception:
            pass
    parallelism = {None: "1 GPU",
                   "owned": "distributed ownership: Morton cell ranges x%d, every rank holds only its cells + one-cell halo (local indices, "
                            "own symbolic pass), owner-computes rows, no data-path collective" % world,
                   "exchange": "Morton cell partition x%d on a replicated pattern, interface rows to owner by one NCCL all-to-all-v "
                               "overlapped with interior rows%s" % (world, ", step replayed as a CUDA graph" if use_graph else ""),
                   "owner": "Morton cell partition x%d on a replicated pattern, owner-computes rows (halo cells recomputed, no data-path "
                            "collective)" % world,
                   "owner_rows": "%d contiguous row blocks of equal nnz on a replicated pattern, owner-computes (halo cells recomputed, no "
                                 "data-path collective)" % world}[dist_mode]
    out = {
        "metric": "cells assembled/sec into CSR", "value": n_cells / (ms_per_step * 1e-3), "unit": "cells/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "cells": n_cells, "dofs": n_dofs, "nnz": nnz_global, "degree": degree,
                   "algo": args.algo, "l2": "inputs+outputs per step (%.2f GB) exceed the 126 MB L2; no explicit flush" % (alg_bytes / 1e9),
                   "parallelism": parallelism,
                   "symbolic_pass_s": round(t_sym, 3), "setup_s": round(t_setup, 3), "partition_s": round(t_part, 3)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_cell": alg_bytes / n_cells,
                     "kernel": kernel_name},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(t_timed0, max(t_timed1, t_end)),
    }
    if rank_facts is not None:
        out["config"]["per_rank"] = rank_facts
    if check is not None:
        out["check"] = check
    if e2e is not None:
        out["e2e"] = e2e
    # the other BASELINE configurations that fit one GPU, each as a short measured line (ms, roofline fraction, witness)
    if world == 1 and args.workload == "c5_1e8" and args.algo == "auto" and not args.no_other_configs:
        del values, pat, dm, mesh
        others = {}
        for name in ("c2", "c3", "c4"):
            try:
                others[name] = short_run(ctx, lf, np, name, peak)
            except Exception as e:
                others[name] = {"error": str(e)[:200]}
        out["other_configs"] = others
    if not args.no_cpu_baseline:
        ncpu = 1000
        cells, sec = cpu_reference_sample(ncpu)
        out["cpu_baseline"] = {"value": cells / sec, "unit": "cells/s", "cores": 1, "kind": "port",
                               "sample": "same operator on TP-triangle mesh n=%d (%d cells): AssembleMatrixLocally->COO + makeSparse, "
                                         "%.2f s; host has %d cores, the reference is serial" % (ncpu, cells, sec, os.cpu_count())}
    print_result(out)
    teardown()


def kernel_of(workload, kind, degree, algo):
    """Name of the kernel that dominates the step (what LFGPU_ALGO_AUTO runs, lehrfempp_b200/csrc/assemble.cu)."""
    tri = kind == "tp_tria" or kind.startswith("refined:") or kind == "delaunay"
    row_kernels = {1: "k_assemble_p1_fan", 2: "k_p2_vertex_rows + k_p2_edge_rows", 3: "k_p3_vertex_rows + k_p3_edge_rows + k_p3_cell_rows"}
    if algo in ("auto", "fan"):
        if tri and degree in AUTO_ROW_DEGREES:
            return row_kernels[degree]
        return "k_assemble_p1_rows" if degree == 1 else "k_assemble_items"
    return {"gather": "k_assemble_items", "atomic": "k_assemble_atomic"}[algo]


def build_mesh(ctx, np, kind, n):
    if kind == "tp_tria":
        return ctx.mesh_tp_tria(n, n)
    if kind.startswith("refined:"):
        # MeshHierarchy-refined mesh (BASELINE config 4): builder mesh + regular refinement steps with the reference's numbering
        mesh = ctx.mesh_tp_tria(n, n)
        for _ in range(int(kind.split(":")[1])):
            mesh = mesh.refine_regular()
        return mesh
    if kind == "delaunay":
        from scipy.spatial import Delaunay
        pts = np.random.default_rng(12345).random((n, 2))
        tri = Delaunay(pts).simplices
        a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
        area2 = np.abs((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1]))
        longest2 = np.maximum.reduce([((b - a) ** 2).sum(1), ((c - b) ** 2).sum(1), ((a - c) ** 2).sum(1)])
        tri = tri[area2 > 0.05 * longest2]  # slivers along the convex hull
        used = np.unique(tri)
        remap = np.full(n, -1, dtype=np.int64)
        remap[used] = np.arange(used.size)
        cn = np.full((tri.shape[0], 4), 0xFFFFFFFF, dtype=np.uint32)
        cn[:, :3] = remap[tri]
        return ctx.mesh_upload(pts[used], cn)
    return ctx.mesh_hybrid(n, 0.2, 12345)


def make_coeffs(ctx, lf, workload, mesh, degree):
    """(alpha, gamma, bytes of coefficient data one pass reads) of the workload on `mesh` (the whole mesh or a rank's share)."""
    if workload == "c2":
        stride = 4
        xy = mesh.qp_coords(degree, stride).to_host().reshape(mesh.n_cells, stride, 2)
        r2 = xy[..., 0] ** 2 + xy[..., 1] ** 2
        return (lf.Coeff.per_qp(ctx.to_device(1.0 + r2), stride), lf.Coeff.per_qp(ctx.to_device(1.0 / (1.0 + r2)), stride),
                2 * 8.0 * (3 * mesh.n_tria + 4 * mesh.n_quad))
    if workload in ("c4s", "c4", "c4_27m", "c4_full"):
        return lf.Coeff.const(1.0), lf.Coeff.const(1.0), 0.0
    return lf.Coeff.const(1.0), lf.Coeff.const(0.0), 0.0


def witness_record(workload, wmax, wsum, world):
    where = "" if world == 1 else " (every rank on the rows it owns; max / sum over %d ranks)" % world
    if workload in ("c4s", "c4", "c4_27m", "c4_full"):
        return {"what": "1^T A 1 for stiffness + mass with alpha = gamma = 1 on the unit square" + where, "value": wsum, "expected": 1.0}
    if workload == "c2":
        from scipy.integrate import dblquad
        expected = dblquad(lambda yy, xx: 1.0 / (1.0 + xx * xx + yy * yy), 0.0, 1.0, 0.0, 1.0)[0]
        return {"what": "1^T A 1 = integral of gamma = 1 / (1 + |x|^2) over the unit square (quadrature error O(h^2))" + where,
                "value": wsum, "expected": expected}
    return {"what": "max |A 1| for the Laplacian (constants are in its kernel; entries are O(1))" + where, "value": wmax, "expected": 0.0}


def short_run(ctx, lf, np, workload, peak, steps=20, warmup=3):
    """One of the other single-GPU BASELINE configurations, measured like the main one (CUDA events on the ctx stream, inputs
    resident, working set far above L2) but with few steps: ms per pass, fraction of the HBM roofline, witness."""
    desc, kind, n, degree = WORKLOADS[workload]
    mesh = build_mesh(ctx, np, kind, n)
    dm = mesh.dofmap_lagrange(degree)
    pat = dm.symbolic(major=lf.ROW_MAJOR)
    alpha, gamma, coef_bytes = make_coeffs(ctx, lf, workload, mesh, degree)
    values = ctx.empty(pat.nnz)
    for _ in range(warmup):
        pat.assemble_reaction_diffusion(degree, alpha, gamma, out=values)
    ctx.synchronize()
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(steps):
        pat.assemble_reaction_diffusion(degree, alpha, gamma, out=values)
    ctx.record(e1)
    ms = ctx.elapsed_ms(e0, e1) / steps
    y = pat.spmv(values, ctx.to_device(np.ones(dm.num_dofs))).to_host()
    check = witness_record(workload, float(np.abs(y).max()), float(y.sum()), 1)
    nsf_t, nsf_q = {1: (3, 4), 2: (6, 9), 3: (10, 16)}[degree]
    alg = 4.0 * (nsf_t * mesh.n_tria + nsf_q * mesh.n_quad) + 16.0 * mesh.n_nodes + coef_bytes + 8.0 * pat.nnz
    achieved = alg / (ms * 1e-3) / 1e9
    return {"workload": desc, "cells": mesh.n_cells, "nnz": pat.nnz, "ms_per_step": ms, "steps": steps, "value": mesh.n_cells / (ms * 1e-3),
            "unit": "cells/s", "roofline": {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                            "algorithmic_bytes_per_cell": alg / mesh.n_cells,
                                            "kernel": kernel_of(workload, kind, degree, "auto")},
            "check": check}


if __name__ == "__main__":
    main()
