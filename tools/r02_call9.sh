#!/bin/bash
# last single-GPU check of round 2: the whole GPU suite and the default bench line with the shipped defaults
set -u
out=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $out/r02_gpu_tests_final2.log
timeout 400 python bench.py > $out/r02_bench_default_final2.json 2> $out/bench_default_final2.err
python - <<PY
import json
d=json.load(open("$out/r02_bench_default_final2.json"))
print(d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["clocks"])
for k,v in d["other_configs"].items(): print(k, v["ms_per_step"], v["roofline"]["frac"], v["check"]["value"])
PY
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
