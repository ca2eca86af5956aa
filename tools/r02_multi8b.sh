#!/bin/bash
# second 8-GPU call of round 2: exchange-mode parity after the unpack-add fix, e2e / NUMA probe, bench lines after the row restriction
set -u
N=8
out=gpurun_out
mkdir -p $out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
LFGPU_DIST_MODE=exchange run 29612 tests/dist_gpu_check.py > $out/r02_dist_check_exchange_n$N.log 2>&1; tail -n 1 $out/r02_dist_check_exchange_n$N.log
run 29630 tools/e2e_numa_probe.py > $out/r02_e2e_numa_probe_n$N.log 2> $out/e2e_probe.err; grep '^{' $out/r02_e2e_numa_probe_n$N.log | cut -c1-600
run 29620 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline > $out/r02_bench_n${N}_owned.json 2> $out/bench_n${N}_owned.err; tail -c 1200 $out/r02_bench_n${N}_owned.json; echo
LFGPU_DIST_MODE=exchange run 29623 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > $out/r02_bench_n${N}_exchange.json 2> $out/bench_n${N}_exchange.err; tail -c 600 $out/r02_bench_n${N}_exchange.json; echo
LFGPU_DIST_MODE=exchange LFGPU_BENCH_GRAPH=1 run 29624 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > $out/r02_bench_n${N}_exchange_graph.json 2> $out/bench_n${N}_exchange_graph.err; tail -c 600 $out/r02_bench_n${N}_exchange_graph.json; echo
