#!/bin/bash
# two-pass load vector + cell-order view: parity, then gather / atomic / auto with a tabulated source at C3 (P2), C4's size (P3), 1e8 (P1)
set -u
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_dynamic_gmsh_loadgather.py tests/test_gpu_load_fan.py tests/test_gpu_parity.py tests/test_gpu_nodal_coeff.py tests/test_gpu_multi_capi.py -x -q 2>&1 | tail -3
timeout 120 python tools/load_probe.py 2828 2 per_qp > $out/r02_load_probe_p2_perqp.json 2>$out/load_probe.err; cat $out/r02_load_probe_p2_perqp.json
timeout 120 python tools/load_probe.py 1448 3 per_qp > $out/r02_load_probe_p3_perqp.json 2>>$out/load_probe.err; cat $out/r02_load_probe_p3_perqp.json
timeout 200 python tools/load_probe.py 7071 1 per_qp > $out/r02_load_probe_p1_perqp.json 2>>$out/load_probe.err; cat $out/r02_load_probe_p1_perqp.json
LFGPU_CELL_ORDER=0 timeout 200 python tools/load_probe.py 7071 1 per_qp > $out/r02_load_probe_p1_perqp_nocellorder.json 2>>$out/load_probe.err; cat $out/r02_load_probe_p1_perqp_nocellorder.json
tail -2 $out/load_probe.err
