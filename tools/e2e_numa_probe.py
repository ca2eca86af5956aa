"""Why does the end-to-end step barely scale with the number of GPUs (round 1: 56 / 46 / 45 / 33 ms at 1 / 2 / 4 / 8)?  Every rank
copies 0.35 GB device -> pinned host and 0.1 GB back, all ranks at once (torchrun, one process per GPU), three ways: as the
process starts; after binding the process to the CPUs local to its GPU (sysfs local_cpulist) BEFORE the pinned buffer is
allocated (first touch puts it on that NUMA node); and with the ranks taking turns (what one link can do alone).  Prints one
JSON line per rank."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
props = torch.cuda.get_device_properties(local)
bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
sysd = "/sys/bus/pci/devices/" + bus
info = {"rank": rank, "bus": bus}
for f in ("numa_node", "local_cpulist"):
    try:
        info[f] = open(os.path.join(sysd, f)).read().strip()
    except OSError as e:
        info[f] = "n/a (%s)" % e.__class__.__name__
info["affinity_at_start"] = len(os.sched_getaffinity(0))
N = 44_000_000  # doubles: 0.35 GB
dev = torch.empty(N, dtype=torch.float64, device="cuda")


def measure(tag, exclusive=False):
    host = torch.empty(N, dtype=torch.float64, pin_memory=True)
    host.zero_()
    torch.cuda.synchronize()
    for turn in range(world if exclusive else 1):
        dist.barrier()
        if exclusive and turn != rank:
            continue
        t0 = time.perf_counter()
        for _ in range(5):
            host.copy_(dev, non_blocking=True)
        torch.cuda.synchronize()
        d2h = 5 * N * 8 / (time.perf_counter() - t0) / 1e9
        t0 = time.perf_counter()
        for _ in range(5):
            dev[: N // 4].copy_(host[: N // 4], non_blocking=True)
        torch.cuda.synchronize()
        h2d = 5 * (N // 4) * 8 / (time.perf_counter() - t0) / 1e9
        info[tag] = {"d2h_GBs": round(d2h, 1), "h2d_GBs": round(h2d, 1)}
    dist.barrier()
    del host


measure("concurrent_default")
measure("one_at_a_time", exclusive=True)
try:
    cpus = set()
    for part in info["local_cpulist"].split(","):
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    cpus &= os.sched_getaffinity(0)
    if cpus:
        os.sched_setaffinity(0, cpus)
        info["bound_to"] = len(cpus)
        measure("concurrent_numa_bound")
except Exception as e:  # no sysfs topology in this VM
    info["bind_error"] = str(e)[:100]
print(json.dumps(info), flush=True)
dist.destroy_process_group()
