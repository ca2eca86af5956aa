#!/bin/bash
# any-role first-use order: parity subset, then compact edge plan on / off at C3, P3 builder mesh, u2
set -u
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_p2_rows.py tests/test_gpu_p3_rows.py tests/test_gpu_zz_plan_variants.py tests/test_gpu_p23_cell_corners.py -x -q 2>&1 | tail -3
timeout 90 python tools/rows_probe.py 2 2828 rows > $out/r02_p2_rows_anyrole_compact1.json 2>/dev/null; cat $out/r02_p2_rows_anyrole_compact1.json
LFGPU_P2_COMPACT=v timeout 90 python tools/rows_probe.py 2 2828 rows > $out/r02_p2_rows_anyrole_compactv.json 2>/dev/null; cat $out/r02_p2_rows_anyrole_compactv.json
timeout 90 python tools/rows_probe.py 3 1448 rows > $out/r02_p3_rows_anyrole.json 2>/dev/null; cat $out/r02_p3_rows_anyrole.json
timeout 120 python bench.py --workload u2 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $out/tmp_u2.json 2>/dev/null
python -c "import json;d=json.load(open('$out/tmp_u2.json'));print('u2', d['ms_per_step'])"
timeout 200 ncu --set full --clock-control none -k "regex:k_p2_(vertex|edge)_rows" -c 2 -f -o $out/r02_p2_rows_shipped2 \
  python tools/rows_probe.py 2 2828 rows > $out/ncu_p2_shipped2.log 2>&1
tail -1 $out/ncu_p2_shipped2.log
