#!/bin/bash
set -u
out=gpurun_out
timeout 300 python -m pytest tests/test_gpu_load_fan.py -x -q 2>&1 | tail -2
timeout 200 python tools/load_probe.py 7071 1 per_qp > $out/r02_load_probe_p1_perqp_ordered_v4.json 2>/dev/null; cat $out/r02_load_probe_p1_perqp_ordered_v4.json
LFGPU_LOAD_ROWORDER=0 timeout 200 python tools/load_probe.py 7071 1 per_qp > $out/r02_load_probe_p1_perqp_unordered_v4.json 2>/dev/null; cat $out/r02_load_probe_p1_perqp_unordered_v4.json
