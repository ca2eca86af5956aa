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
    n = 707  # 1.0e6 triangles: the smallest size of config C5, a bounded sample of the 1e8 workload
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
    sample = "P1 Laplacian on TP-triangle mesh n=707 (%d cells), AssembleMatrixLocally->COO + makeSparse per step" % m.n_cells
    all_cores = cpu_all_cores_sample(n)
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
    structured = kind == "tp_tria" or kind.startswith("refined:") or kind == "delaunay"  # triangle meshes: the row kernels apply
    # kernels that own matrix rows in registers (LFGPU_ALGO_AUTO takes them on triangle meshes with constant coefficients)
    row_kernels = {1: "k_assemble_p1_fan", 2: "k_p2_vertex_rows + k_p2_edge_rows", 3: "k_p3_vertex_rows + k_p3_edge_rows + k_p3_cell_rows"}
    if args.algo == "auto":
        kernel_name = row_kernels[degree] if (structured and degree in AUTO_ROW_DEGREES) else "k_assemble_items"
    elif args.algo == "fan":
        kernel_name = row_kernels[degree]
    else:
        kernel_name = {"gather": "k_assemble_items", "atomic": "k_assemble_atomic"}[args.algo]

    # ---- setup (untimed, like mesh / DofHandler construction on the CPU side) ----------------------------------------------
    t_setup = time.time()
    if kind == "tp_tria":
        mesh = ctx.mesh_tp_tria(n, n)
    elif kind.startswith("refined:"):
        # MeshHierarchy-refined mesh (BASELINE config 4): builder mesh + regular refinement steps with the reference's numbering
        mesh = ctx.mesh_tp_tria(n, n)
        for _ in range(int(kind.split(":")[1])):
            mesh = mesh.refine_regular()
    elif kind == "delaunay":
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
        mesh = ctx.mesh_upload(pts[used], cn)
        del pts, tri, a, b, c, area2, longest2, cn
    else:
        mesh = ctx.mesh_hybrid(n, 0.2, 12345)
    dm = mesh.dofmap_lagrange(degree)
    t_sym = time.time()
    pat = dm.symbolic(major=lf.ROW_MAJOR)
    ctx.synchronize()
    t_sym = time.time() - t_sym
    if args.workload == "c2":
        stride = 4
        xy = mesh.qp_coords(degree, stride).to_host().reshape(mesh.n_cells, stride, 2)
        r2 = xy[..., 0] ** 2 + xy[..., 1] ** 2
        alpha = lf.Coeff.per_qp(ctx.to_device(1.0 + r2), stride)
        gamma = lf.Coeff.per_qp(ctx.to_device(1.0 / (1.0 + r2)), stride)
        coef_bytes = 2 * 8.0 * (3 * mesh.n_tria + 4 * mesh.n_quad)
    elif args.workload in ("c4s", "c4", "c4_27m"):
        alpha, gamma, coef_bytes = lf.Coeff.const(1.0), lf.Coeff.const(1.0), 0.0
    else:
        alpha, gamma, coef_bytes = lf.Coeff.const(1.0), lf.Coeff.const(0.0), 0.0
    values = ctx.empty(pat.nnz)
    t_setup = time.time() - t_setup

    # N > 1: Morton partition of the cells; every rank assembles the contributions of its cells, partial sums of
    # interface rows go to the row owner in one all-to-all-v (NCCL) overlapped with the interior rows
    asm = None
    t_part = 0.0
    if world > 1:
        from lehrfempp_b200.distributed import DistributedAssembler
        t_part = time.time()
        # default: owner-computes (faster: one launch per rank, no collective); LFGPU_DIST_MODE=exchange selects the
        # contributions-to-owner variant with the NCCL all-to-all-v (both are parity-tested by tests/dist_gpu_check.py)
        # owner_rows (default): the same owner-computes scheme over contiguous row blocks -- one range launch per rank
        dist_mode = os.environ.get("LFGPU_DIST_MODE", "owner_rows")
        asm = DistributedAssembler(ctx, mesh, pat, degree, mode=dist_mode)
        t_part = time.time() - t_part

    use_graph = asm is not None and os.environ.get("LFGPU_BENCH_GRAPH", "0") == "1"
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

    # ---- end to end through the C ABI with host buffers ---------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        h_xy = ctx.pinned(2 * mesh.n_nodes)
        d = mesh.download()
        h_xy[:] = d["node_coords"].ravel()
        del d
        if asm is None:
            h_vals = ctx.pinned(pat.nnz)
            d2h = lambda: ctx.d2h_async(h_vals, values)  # noqa: E731
            d2h_bytes = 8 * pat.nnz
        else:
            # every rank returns the rows it owns: pack their segments, then one D2H copy
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
        e2e_steps = max(3, min(args.steps, 10))

        e2e_blocks = int(os.environ.get("LFGPU_E2E_BLOCKS", "16"))
        # owner_rows on the fan kernel: every rank runs the host-buffer call for ITS row block -- uploads only the coordinate
        # window its rows refer to and downloads only its rows (lfgpu_assemble_reaction_diffusion_host_range)
        # (P1 only: the P2 / P3 row kernels also take row ranges, but the host-buffer range call knows the coordinate window of
        # the fan kernel only -- they go through the generic upload / step / download sequence below)
        range_mode = (asm is not None and asm.mode == "owner_rows" and getattr(asm, "_range_ok", False) and args.algo in ("auto", "fan")
                      and degree == 1)
        h2d_rank = 16 * mesh.n_nodes
        if range_mode:
            inner_t = torch.as_tensor(pat.download()[1], device="cuda")
            seg = inner_t[int(outer_t[asm.row0].item()):int(outer_t[asm.row0 + asm.n_rows].item())]
            h2d_rank = 16 * (int(seg.max().item()) + 1 - int(seg.min().item())) if seg.numel() > 0 else 0
            del inner_t, seg

        def e2e_step():
            if range_mode:
                pat.assemble_reaction_diffusion_host_range(degree, alpha, gamma, h_xy, h_vals, asm.row0, asm.n_rows, out=values, algo=algo,
                                                           n_blocks=max(2, e2e_blocks // world))
                return
            if asm is None:
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
        e2e = {"value": mesh.n_cells / e2e_s, "unit": "cells/s", "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_bytes),
               "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "what": ("per step: one lfgpu_assemble_reaction_diffusion_host call = H2D node coordinates (pinned) -> kernel -> D2H CSR "
                        "values (pinned), pipelined over %d row blocks" % e2e_blocks) if asm is None else
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

    if rank != 0:
        teardown()
        return

    # size-independent witness that the matrix left in `values` at the FULL size is the right operator (not part of any timing):
    # one device SpMV with the constant vector -- constants are in the kernel of the stiffness part, 1^T M 1 = |Omega| = 1
    check = None
    if asm is None:
        try:
            y = pat.spmv(values, ctx.to_device(np.ones(dm.num_dofs))).to_host()
            if args.workload in ("c4s", "c4", "c4_27m"):
                check = {"what": "1^T A 1 for stiffness + mass with alpha = gamma = 1 on the unit square", "value": float(y.sum()), "expected": 1.0}
            elif args.workload == "c2":
                from scipy.integrate import dblquad
                expected = dblquad(lambda yy, xx: 1.0 / (1.0 + xx * xx + yy * yy), 0.0, 1.0, 0.0, 1.0)[0]
                check = {"what": "1^T A 1 = integral of gamma = 1 / (1 + |x|^2) over the unit square (quadrature error O(h^2))",
                         "value": float(y.sum()), "expected": expected}
            else:
                check = {"what": "max |A 1| for the Laplacian (constants are in its kernel; entries are O(1))", "value": float(np.abs(y).max()),
                         "expected": 0.0}
            del y
        except Exception as e:  # a witness must never cost the bench line
            check = {"error": str(e)[:200]}

    peak, peak_src = measured_peak()
    # algorithmic bytes (SURVEY.md 8d / DESIGN.md): int32 connectivity + each vertex coordinate once + coefficients +
    # each stored value written once
    nldof_sum = dm.stride * mesh.n_cells if mesh.n_quad == 0 or mesh.n_tria == 0 else None
    nsf_t, nsf_q = {1: (3, 4), 2: (6, 9), 3: (10, 16)}[degree]
    conn = 4.0 * (nsf_t * mesh.n_tria + nsf_q * mesh.n_quad)
    alg_bytes = conn + 16.0 * mesh.n_nodes + coef_bytes + 8.0 * pat.nnz
    # per launch (= per rank for N > 1) the kernel covers 1/world of the rows
    alg_bytes_launch = alg_bytes / world
    achieved = alg_bytes_launch / (ms_per_step * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp)).get(args.workload + ":" + args.algo)
            if tj:
                traffic = tj["dram_bytes_per_cell"] * mesh.n_cells / world
        except Exception:
            pass
    out = {
        "metric": "cells assembled/sec into CSR", "value": mesh.n_cells / (ms_per_step * 1e-3), "unit": "cells/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "cells": mesh.n_cells, "dofs": dm.num_dofs, "nnz": pat.nnz, "degree": degree,
                   "algo": args.algo, "l2": "inputs+outputs per step (%.2f GB) exceed the 126 MB L2; no explicit flush" % (alg_bytes / 1e9),
                   "parallelism": "1 GPU" if world == 1 else (
                       "Morton cell partition x%d, interface rows to owner by one NCCL all-to-all-v overlapped with interior rows%s" % (world, ", step replayed as a CUDA graph" if use_graph else "")
                       if asm.mode == "exchange" else
                       "Morton cell partition x%d, owner-computes rows (halo cells recomputed, no data-path collective)" % world
                       if asm.mode == "owner" else
                       "%d contiguous row blocks of equal nnz, owner-computes (halo cells recomputed, no data-path collective)" % world),
                   "symbolic_pass_s": round(t_sym, 3), "setup_s": round(t_setup, 3), "partition_s": round(t_part, 3)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_cell": alg_bytes / mesh.n_cells,
                     "kernel": kernel_name},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(t_timed0, max(t_timed1, t_end)),
    }
    if check is not None:
        out["check"] = check
    if e2e is not None:
        out["e2e"] = e2e
    if not args.no_cpu_baseline:
        ncpu = 1000
        cells, sec = cpu_reference_sample(ncpu)
        out["cpu_baseline"] = {"value": cells / sec, "unit": "cells/s", "cores": 1, "kind": "port",
                               "sample": "same operator on TP-triangle mesh n=%d (%d cells): AssembleMatrixLocally->COO + makeSparse, "
                                         "%.2f s; host has %d cores, the reference is serial" % (ncpu, cells, sec, os.cpu_count())}
    print_result(out)
    teardown()


if __name__ == "__main__":
    main()
