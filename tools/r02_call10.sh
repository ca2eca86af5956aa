#!/bin/bash
set -u
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_load_fan.py tests/test_gpu_dynamic_gmsh_loadgather.py tests/test_gpu_parity.py -x -q -k "load" 2>&1 | tail -15
timeout 200 python tools/load_probe.py 7071 1 per_qp > $out/r02_load_probe_p1_perqp_ring.json 2>$out/load_probe.err; cat $out/r02_load_probe_p1_perqp_ring.json
tail -2 $out/load_probe.err
