#!/bin/bash
# Multi-GPU sanity at the end of round 2 (gpurun --gpus N -- bash tools/r02_multi_final.sh N): the 2-GPU tests of the suite, parity of
# the default mode against the oracle, bench lines of the default workload (with e2e), C3 and C4 at its configured size.
set -u
N=${1:-2}
out=gpurun_out
mkdir -p $out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multi_capi.py -x -q 2>&1 | tail -3 | tee $out/r02_multi_tests_final_n$N.log
fi
if [ "$N" != "8" ]; then run 29601 tests/dist_owned_check.py > $out/r02_dist_owned_final_n$N.log 2>&1; tail -1 $out/r02_dist_owned_final_n$N.log; fi  # (8 GPUs: profiles/r02_dist_owned_n8.log)
run 29620 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline > $out/r02_bench_final_n${N}_owned.json 2> $out/bench_final_n${N}.err
python - <<PY
import json
d=json.load(open("$out/r02_bench_final_n${N}_owned.json"))
print("default", d["n_gpus"], d["ms_per_step"], d["value"], d["roofline"]["frac"], d["check"]["value"], d.get("e2e",{}).get("ms_per_step"))
PY
for w in c4_full $([ "$N" != "8" ] && echo c3); do
  run 29622 bench.py --gpus $N --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $out/r02_bench_final_n${N}_$w.json 2> $out/bench_final_n${N}_$w.err
  python - <<PY
import json
try:
    d=json.load(open("$out/r02_bench_final_n${N}_$w.json"))
    print("$w", d["n_gpus"], d["ms_per_step"], d["value"], d["roofline"]["frac"], d["check"]["value"])
except Exception as e:
    print("$w failed", e)
PY
done
