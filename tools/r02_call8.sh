#!/bin/bash
# two-pass load vector: parity, then gather / atomic / auto at C3 (P2) and C4-size (P3) with a tabulated source, and P1 per-point at 1e8
set -u
out=gpurun_out
true
timeout 120 python tools/load_probe.py 2828 2 per_qp > $out/r02_load_probe_p2_perqp.json 2>$out/load_probe.err; cat $out/r02_load_probe_p2_perqp.json
timeout 120 python tools/load_probe.py 1448 3 per_qp > $out/r02_load_probe_p3_perqp.json 2>>$out/load_probe.err; cat $out/r02_load_probe_p3_perqp.json
true
timeout 200 python tools/load_probe.py 7071 1 per_qp > $out/r02_load_probe_p1_perqp.json 2>>$out/load_probe.err; cat $out/r02_load_probe_p1_perqp.json
tail -2 $out/load_probe.err
