#!/bin/bash
# Round 2, final single-GPU call: the whole GPU suite, the default bench line (+ reference arm), the launch list of the bench command,
# ncu --set full of the shipped P2 kernels and of the C2 kernel (DRAM bytes for profiles/traffic.json).
set -u
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $out/r02_gpu_tests_final.log
timeout 400 python bench.py > $out/r02_bench_default_final.json 2> $out/bench_default_final.err
tail -c 1500 $out/r02_bench_default_final.json; echo
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $out/r02_bench_reference_final.json 2>> $out/bench_default_final.err
tail -c 600 $out/r02_bench_reference_final.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02_launches_bench_default_final.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:k_p2_(vertex|edge)_rows" -c 2 -f -o $out/r02_p2_rows_shipped \
  python tools/rows_probe.py 2 2828 rows > $out/ncu_p2_shipped.log 2>&1
tail -1 $out/ncu_p2_shipped.log
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:k_assemble_p1_rows" -c 1 -f -o $out/r02_c2_shipped \
  python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $out/ncu_c2_shipped.log 2>&1
tail -1 $out/ncu_c2_shipped.log
