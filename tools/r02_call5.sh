#!/bin/bash
# Round 2, fifth call: final defaults (compact P2 plans + edge rows on their own coordinate copy): variants test, u2 A/B, refined-mesh row
# classes, ncu --set full captures for the DRAM bytes of C3 / C4 (profiles/traffic.json).
set -u
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_zz_plan_variants.py tests/test_gpu_p2_rows.py -x -q 2>&1 | tail -3
for v in "" "LFGPU_P2_COMPACT=0" "LFGPU_EDGE_ORDER=0" "LFGPU_P2_COMPACT=0 LFGPU_EDGE_ORDER=0"; do
  env $v timeout 120 python bench.py --workload u2 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $out/tmp_u2.json 2>/dev/null
  python -c "import json;d=json.load(open('$out/tmp_u2.json'));print('u2 [$v]', d['ms_per_step'])"
done
timeout 90 python tools/rows_probe.py 3 181 rows 3 > $out/r02_p3_rows_refined.json 2>/dev/null; cat $out/r02_p3_rows_refined.json
LFGPU_EDGE_ORDER=0 timeout 90 python tools/rows_probe.py 3 181 rows 3 > $out/r02_p3_rows_refined_noorder.json 2>/dev/null; cat $out/r02_p3_rows_refined_noorder.json
timeout 90 python tools/rows_probe.py 2 2828 rows > $out/r02_p2_rows_final.json 2>/dev/null; cat $out/r02_p2_rows_final.json
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:k_p2_(vertex|edge)_rows" -c 2 -f -o $out/r02_p2_rows_edge_order \
  python tools/rows_probe.py 2 2828 rows > $out/ncu_p2_eo.log 2>&1
tail -1 $out/ncu_p2_eo.log
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:k_p3_(vertex|edge|cell)_rows" -c 3 -f -o $out/r02_p3_rows_edge_order \
  python tools/rows_probe.py 3 181 rows 3 > $out/ncu_p3_eo.log 2>&1
tail -1 $out/ncu_p3_eo.log
