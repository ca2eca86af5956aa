#!/bin/bash
# closing check of round 2 on one GPU: whole suite, smoke, the largest single-GPU P3 pattern (2.1e9 values), load probe P1 constant source
set -u
out=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $out/r02_gpu_tests_final3.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --workload c4_27m --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $out/r02_bench_c4_27m_final.json 2> $out/c4_27m_final.err
python - <<PY
import json
d=json.load(open("$out/r02_bench_c4_27m_final.json"))
print("c4_27m", d["ms_per_step"], d["roofline"]["frac"], d["config"]["nnz"], d["check"]["value"])
PY
timeout 120 python tools/load_probe.py 7071 1 const > $out/r02_load_probe_final.json 2>/dev/null; cat $out/r02_load_probe_final.json
