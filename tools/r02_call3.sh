#!/bin/bash
# Round 2, third call: P2 / P3 row kernels on meshes with per-cell corners (CC instantiations): parity, then timings; regression
# check of the default kernels (the P2 vertex kernel body was restructured); the two remaining compact-plan variants.
set -u
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_p23_cell_corners.py tests/test_gpu_p2_rows.py tests/test_gpu_p3_rows.py tests/test_gpu_zz_plan_variants.py \
  tests/test_gpu_zz_p2_general.py tests/test_gpu_zz_row_ranges.py -x -q 2>&1 | tail -5 | tee $out/r02_cc_tests.log
timeout 90 python tools/rows_probe.py 2 2828 rows > $out/r02_p2_rows_after_cc.json 2>/dev/null; cat $out/r02_p2_rows_after_cc.json
timeout 90 python tools/rows_probe.py 3 1448 rows > $out/r02_p3_rows_after_cc.json 2>/dev/null; cat $out/r02_p3_rows_after_cc.json
LFGPU_P2_COMPACT=e LFGPU_EDGE_PFC=50 timeout 90 python tools/rows_probe.py 2 2828 rows > $out/r02_p2_rows_compact_e_pf.json 2>/dev/null; cat $out/r02_p2_rows_compact_e_pf.json
LFGPU_P2_COMPACT=v timeout 90 python tools/rows_probe.py 2 2828 rows > $out/r02_p2_rows_compact_v.json 2>/dev/null; cat $out/r02_p2_rows_compact_v.json
timeout 120 python tools/cc_probe.py 2 2000 > $out/r02_cc_probe_p2.json 2>$out/cc_probe.err; cat $out/r02_cc_probe_p2.json
timeout 120 python tools/cc_probe.py 3 1448 > $out/r02_cc_probe_p3.json 2>>$out/cc_probe.err; cat $out/r02_cc_probe_p3.json
tail -3 $out/cc_probe.err
