#!/bin/bash
set -u
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_load_fan.py tests/test_gpu_dynamic_gmsh_loadgather.py tests/test_gpu_nodal_coeff.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 120 python tools/load_probe.py 7071 1 const > $out/r02_load_probe_final.json 2>$out/load_probe.err; cat $out/r02_load_probe_final.json
timeout 200 python tools/load_probe.py 7071 1 per_qp > $out/r02_load_probe_p1_perqp_final.json 2>>$out/load_probe.err; cat $out/r02_load_probe_p1_perqp_final.json
timeout 120 python tools/load_probe.py 2828 2 per_qp > $out/r02_load_probe_p2_perqp_final.json 2>>$out/load_probe.err; cat $out/r02_load_probe_p2_perqp_final.json
timeout 120 python tools/load_probe.py 1448 3 per_qp > $out/r02_load_probe_p3_perqp_final.json 2>>$out/load_probe.err; cat $out/r02_load_probe_p3_perqp_final.json
tail -2 $out/load_probe.err
