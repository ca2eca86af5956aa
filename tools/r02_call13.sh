#!/bin/bash
set -u
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_load_fan.py tests/test_gpu_zz_plan_variants.py -x -q -k "load" 2>&1 | tail -6
timeout 200 python tools/load_probe.py 7071 1 per_qp > $out/r02_load_probe_p1_perqp_ordered.json 2>$out/load_probe.err; cat $out/r02_load_probe_p1_perqp_ordered.json
tail -2 $out/load_probe.err
