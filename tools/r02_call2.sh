#!/bin/bash
# Round 2, compact P2 plan + L2 eviction hints: parity of every variant, then A/B of the four switch combinations at config C3
# (tools/rows_probe.py prints vertex / edge rows separately), one ncu capture of the shipped variant for the DRAM bytes.
set -u
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_p2_rows.py tests/test_gpu_zz_plan_variants.py tests/test_gpu_zz_p2_general.py -x -q 2>&1 | tail -5 | tee $out/r02_p2_compact_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for c in 1 0; do for h in 1 0; do
  LFGPU_P2_COMPACT=$c LFGPU_L2_HINTS=$h timeout 90 python tools/rows_probe.py 2 2828 rows > $out/r02_p2_rows_compact${c}_hints${h}.json 2>/dev/null
  echo "compact=$c hints=$h"; cat $out/r02_p2_rows_compact${c}_hints${h}.json
done; done
timeout 120 python bench.py --workload c3 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $out/r02_bench_c3_compact.json 2> $out/bench_c3_compact.err
tail -c 700 $out/r02_bench_c3_compact.json; echo
timeout 120 python bench.py --workload u2 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $out/r02_bench_u2_compact.json 2>> $out/bench_c3_compact.err
tail -c 400 $out/r02_bench_u2_compact.json; echo
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:k_p2_(vertex|edge)_rows" -c 2 -f -o $out/r02_p2_rows_compact \
  python tools/rows_probe.py 2 2828 rows > $out/ncu_p2_compact.log 2>&1
tail -2 $out/ncu_p2_compact.log
