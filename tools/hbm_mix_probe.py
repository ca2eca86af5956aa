"""HBM calibration for DESIGN.md: write-only, read-only and copy rates on this GPU (torch kernels, CUDA events).

The P1 fan kernel writes 2.75 GB and reads 1.8 GB per launch; this shows what the memory system gives for such mixes."""
import json
import torch


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


def main():
    n = 350_000_000  # doubles = 2.8 GB, the value array of the headline workload
    a = torch.empty(n, dtype=torch.float64, device="cuda")
    b = torch.empty(n, dtype=torch.float64, device="cuda")
    c = torch.empty(n * 2 // 3, dtype=torch.float64, device="cuda")  # 1.87 GB, about what the kernel reads
    out = {}
    t = timeit(lambda: a.fill_(1.0))
    out["write_only_GBps"] = 8 * n / t / 1e6
    t = timeit(lambda: torch.sum(a))
    out["read_only_GBps"] = 8 * n / t / 1e6
    t = timeit(lambda: b.copy_(a))
    out["copy_GBps"] = 16 * n / t / 1e6
    # 2 : 3 read : write mix like the fan kernel: read c (1.87 GB), write a (2.8 GB) -- two back-to-back kernels cannot
    # overlap, so emulate with one kernel: a[:m] = c * 2 (1:1) plus fill of the rest in the same launch is not expressible
    # in torch; report the serial lower bound instead
    m = c.numel()
    t = timeit(lambda: (torch.mul(c, 2.0, out=a[:m]), a[m:].fill_(1.0)))
    out["mix_read1.87_write2.8_serial_ms"] = t
    out["mix_GBps"] = (8 * m + 8 * n) / t / 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
