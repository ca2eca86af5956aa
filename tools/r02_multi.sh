#!/bin/bash
# Multi-GPU evidence of round 2 (run under `gpurun --gpus N -- bash tools/r02_multi.sh N`): parity of every partitioned mode
# against the oracle, then the bench lines.  Logs go to gpurun_out/ (copied to profiles/ by hand).
set -u
N=${1:-2}
out=gpurun_out
mkdir -p $out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
# 1. parity: distributed ownership, then the three modes on a replicated pattern
run 29601 tests/dist_owned_check.py > $out/r02_dist_owned_n$N.log 2>&1; tail -1 $out/r02_dist_owned_n$N.log
p=29610
for mode in owner_rows owner exchange; do
  LFGPU_DIST_MODE=$mode run $p tests/dist_gpu_check.py > $out/r02_dist_check_${mode}_n$N.log 2>&1; tail -1 $out/r02_dist_check_${mode}_n$N.log
  p=$((p+1))
done
# 2. bench lines: headline workload in the new default mode and in the round-1 default, then the other configurations
run 29620 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline > $out/r02_bench_n${N}_owned.json 2> $out/bench_n${N}_owned.err; tail -c 1500 $out/r02_bench_n${N}_owned.json; echo
LFGPU_DIST_MODE=owner_rows run 29621 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > $out/r02_bench_n${N}_owner_rows.json 2> $out/bench_n${N}_owner_rows.err; tail -c 700 $out/r02_bench_n${N}_owner_rows.json; echo
for w in c4_full c3 c2; do
  run 29622 bench.py --gpus $N --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $out/r02_bench_n${N}_$w.json 2> $out/bench_n${N}_$w.err; tail -c 1200 $out/r02_bench_n${N}_$w.json; echo; tail -3 $out/bench_n${N}_$w.err
done
