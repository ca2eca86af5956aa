#!/bin/bash
# 8-GPU evidence of round 2 (gpurun --gpus 8 -- bash tools/r02_multi8.sh): parity of every partitioned mode, headline bench in
# the distributed-ownership mode and in the round-1 default, BASELINE config 4 at its full size.
set -u
N=8
out=gpurun_out
mkdir -p $out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29601 tests/dist_owned_check.py > $out/r02_dist_owned_n$N.log 2>&1; tail -1 $out/r02_dist_owned_n$N.log
p=29610
for mode in owner_rows owner exchange; do
  LFGPU_DIST_MODE=$mode run $p tests/dist_gpu_check.py > $out/r02_dist_check_${mode}_n$N.log 2>&1; tail -1 $out/r02_dist_check_${mode}_n$N.log
  p=$((p+1))
done
run 29620 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline > $out/r02_bench_n${N}_owned.json 2> $out/bench_n${N}_owned.err; tail -c 1500 $out/r02_bench_n${N}_owned.json; echo
run 29622 bench.py --gpus $N --workload c4_full --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $out/r02_bench_n${N}_c4_full.json 2> $out/bench_n${N}_c4_full.err; tail -c 1500 $out/r02_bench_n${N}_c4_full.json; echo; tail -3 $out/bench_n${N}_c4_full.err
LFGPU_DIST_MODE=owner_rows run 29621 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > $out/r02_bench_n${N}_owner_rows.json 2> $out/bench_n${N}_owner_rows.err; tail -c 700 $out/r02_bench_n${N}_owner_rows.json; echo
# several GPUs from ONE process through the C ABI (lfgpu_multi_*), device lists [0,1] and [0,1,2,3]
timeout 200 python -m pytest tests/test_gpu_multi_capi.py -q 2>&1 | tail -3 > $out/r02_multi_capi_n$N.log; cat $out/r02_multi_capi_n$N.log
