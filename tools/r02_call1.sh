#!/bin/bash
# First GPU call of the next round (run under gpurun from the repo root):  gpurun --timeout 600 -- 'bash tools/r02_first_call.sh'
# Everything the end of round 1 could not measure any more, outputs under gpurun_out/.
set -u
out=gpurun_out
mkdir -p $out
# 2. bench lines of the configurations that changed kernels (P2 / P3 row kernels on AUTO) and of the weakest one (C2)
for w in c3 c4 c4s c2; do
  timeout 120 python bench.py --workload $w --steps 30 --warmup 5 > $out/r02_bench_$w.json 2> $out/bench_$w.err
  tail -c 600 $out/r02_bench_$w.json; echo
done
# 2b. unstructured mesh: valence-6 plan + generic kernel for the other vertex rows vs the general-valence vertex kernel
timeout 120 python bench.py --workload u2 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $out/r02_bench_u2_default.json 2> $out/bench_u2.err
LFGPU_P2_GENERAL=1 timeout 120 python bench.py --workload u2 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $out/r02_bench_u2_general.json 2>> $out/bench_u2.err
tail -c 400 $out/r02_bench_u2_default.json; echo; tail -c 400 $out/r02_bench_u2_general.json; echo
# 3. full ncu capture of the three P3 row kernels (none exists yet) and of the C2 item kernel
timeout 150 ncu --set full --clock-control none --import-source on -k "regex:k_p3_(vertex|edge|cell)_rows" -c 3 -f -o $out/r02_p3_rows \
  python tools/rows_probe.py 3 1448 rows > $out/ncu_p3.log 2>&1
tail -2 $out/ncu_p3.log
timeout 150 ncu --set full --clock-control none --import-source on -k "regex:k_assemble_items" -c 1 -f -o $out/r02_c2_items \
  python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $out/ncu_c2.log 2>&1
tail -2 $out/ncu_c2.log
# 4. launch lists (kernel shares of a step) for C3 and C4
for w in c3 c4; do
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_p[23]_|k_assemble" -c 40 --csv --log-file $out/r02_launches_$w.csv \
    python bench.py --workload $w --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
done
# 5. probes: both P3 / P2 kernels in one process, load vector
timeout 60 python tools/rows_probe.py 3 1448 > $out/r02_p3_rows_vs_items.json 2>/dev/null; cat $out/r02_p3_rows_vs_items.json
timeout 60 python tools/rows_probe.py 2 2828 > $out/r02_p2_rows_vs_items.json 2>/dev/null; cat $out/r02_p2_rows_vs_items.json
timeout 60 python tools/load_probe.py > $out/r02_load_probe.json 2>/dev/null; cat $out/r02_load_probe.json
# 6. opt-in experiment: coordinate prefetch of the edge rows through the plan (LFGPU_EDGE_PFC = percent of the plan distance);
#    rows_probe prints the row classes one by one ("parts"), so the effect on the edge kernels is visible directly
for pfc in 0 25 50 100; do
  LFGPU_EDGE_PFC=$pfc timeout 60 python tools/rows_probe.py 2 2828 rows > $out/r02_p2_rows_pfc$pfc.json 2>/dev/null; cat $out/r02_p2_rows_pfc$pfc.json
  LFGPU_EDGE_PFC=$pfc timeout 60 python tools/rows_probe.py 3 1448 rows > $out/r02_p3_rows_pfc$pfc.json 2>/dev/null; cat $out/r02_p3_rows_pfc$pfc.json
done
# 7. opt-in experiment: P3 vertex rows with stiffness + mass (MODE 1, config C4) at 128 registers / 4 CTAs per SM instead of 168 / 3
#    (rel_diff against the item kernel in the same line: the variant is a different ptxas schedule of the same source)
LFGPU_P3_VOCC=4 timeout 60 python tools/rows_probe.py 3 1448 > $out/r02_p3_rows_vocc4.json 2>/dev/null; cat $out/r02_p3_rows_vocc4.json
# 8. config C2 (7.6 % of the roofline in round 1): P1 pass on triangle / quadrilateral / hybrid meshes x constant / per-cell /
#    per-point coefficients x every algorithm -- separates the cost of the quadrature loop, the non-affine geometry and the mixed warps
timeout 150 python tools/c2_probe.py > $out/r02_c2_probe.json 2>$out/c2_probe.err; cat $out/r02_c2_probe.json
# 8b. opt-in experiment: item kernel of the P1 quadrature route at 3 CTAs per SM (80 registers, 12 B of spills instead of 64 / 80 B)
LFGPU_ITEMS_OCC=3 timeout 150 python tools/c2_probe.py > $out/r02_c2_probe_occ3.json 2>>$out/c2_probe.err; cat $out/r02_c2_probe_occ3.json
