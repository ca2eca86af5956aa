#!/bin/bash
# Round 2, fourth call: edge rows on their own first-use-ordered coordinate copy -- parity (all P2 / P3 suites, multi-GPU helpers on one
# device), timings of the row classes, bench lines of C3 / C4 / C4 builder mesh / u2.
set -u
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_p23_cell_corners.py tests/test_gpu_p2_rows.py tests/test_gpu_p3_rows.py tests/test_gpu_zz_plan_variants.py \
  tests/test_gpu_zz_p2_general.py tests/test_gpu_zz_row_ranges.py tests/test_gpu_partition.py tests/test_gpu_hostpipe.py tests/test_gpu_zz_mixed_orientation.py \
  -x -q 2>&1 | tail -5 | tee $out/r02_edge_order_tests.log
timeout 90 python tools/rows_probe.py 2 2828 rows > $out/r02_p2_rows_edge_order.json 2>/dev/null; cat $out/r02_p2_rows_edge_order.json
timeout 90 python tools/rows_probe.py 3 1448 rows > $out/r02_p3_rows_edge_order.json 2>/dev/null; cat $out/r02_p3_rows_edge_order.json
LFGPU_P2_COMPACT=1 timeout 90 python tools/rows_probe.py 2 2828 rows > $out/r02_p2_rows_edge_order_compact.json 2>/dev/null; cat $out/r02_p2_rows_edge_order_compact.json
for w in c3 c4 c4s u2; do
  timeout 120 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $out/r02_bench_${w}_edge_order.json 2> $out/bench_${w}_eo.err
  python - <<PY
import json
d=json.load(open("$out/r02_bench_${w}_edge_order.json"))
print("$w", d["ms_per_step"], d["roofline"]["frac"], d["config"].get("symbolic_pass_s"), d["check"]["value"])
PY
done
