"""bench.py's host logic against a STUB device layer (tests/stubs/lehrfempp_b200): every workload x algo combination must get
through argument handling, setup, the timing loop and the roofline arithmetic and print exactly ONE JSON line with the keys
of the driver's contract.  No GPU, no numbers of any meaning -- this only guards the script the driver runs at round end
against Python-level mistakes (one such mistake, a dict built with integer keywords, was found on the GPU box the hard way)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(ROOT, "tests", "stubs")
REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "roofline", "gpu_launches", "clocks"]


def run_bench(*args):
    env = dict(os.environ, PYTHONPATH=STUBS + os.pathsep + os.environ.get("PYTHONPATH", ""), LFGPU_BENCH_STUB="1")
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    # the stub directory must win over the repo root, which bench.py puts first on sys.path: run through -c
    code = ("import sys, runpy; sys.path.insert(0, %r); sys.argv = ['bench.py'] + %r; import lehrfempp_b200 as s; "
            "assert 'stubs' in s.__file__; runpy.run_path(%r, run_name='__main__')" % (STUBS, list(args), os.path.join(ROOT, "bench.py")))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout
    return json.loads(lines[0])


KERNEL = {"c5_1e8": "k_assemble_p1_fan", "c1": "k_assemble_p1_fan", "c2": "k_assemble_p1_rows", "c3": "k_p2_vertex_rows + k_p2_edge_rows",
          "c4": "k_p3_vertex_rows + k_p3_edge_rows + k_p3_cell_rows", "c4s": "k_p3_vertex_rows + k_p3_edge_rows + k_p3_cell_rows",
          "c4_27m": "k_p3_vertex_rows + k_p3_edge_rows + k_p3_cell_rows"}


def test_unstructured_workload():
    out = run_bench("--workload", "u2", "--n", "3000", "--steps", "3", "--no-cpu-baseline", "--no-e2e")
    check_line(out, 3)
    assert out["roofline"]["kernel"] == "k_p2_vertex_rows + k_p2_edge_rows" and 5000 < out["config"]["cells"] < 6100


def check_line(out, steps):
    for key in REQUIRED:
        assert key in out, key
    assert out["n_gpus"] == 1 and out["steps"] == steps and out["higher_is_better"] is True
    assert out["metric"].startswith("cells assembled/sec") and out["unit"] == "cells/s" and out["dtype"] == "f64"
    r = out["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert out["gpu_launches"] >= steps
    assert "workload" in out["config"]
    assert "error" not in out["check"] and {"what", "value", "expected"} <= set(out["check"])


@pytest.mark.parametrize("workload", sorted(KERNEL))
def test_bench_line_contract(workload):
    out = run_bench("--workload", workload, "--steps", "3", "--warmup", "3", "--no-cpu-baseline")
    check_line(out, 3)
    assert out["roofline"]["kernel"] == KERNEL[workload]  # the kernel LFGPU_ALGO_AUTO runs for this workload
    e = out["e2e"]
    assert e["unit"] == "cells/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0


@pytest.mark.parametrize("algo,kernel", [("fan", "k_p2_vertex_rows + k_p2_edge_rows"), ("gather", "k_assemble_items"), ("atomic", "k_assemble_atomic")])
def test_bench_algo_switch(algo, kernel):
    out = run_bench("--workload", "c3", "--algo", algo, "--steps", "4", "--no-cpu-baseline", "--no-e2e")
    check_line(out, 4)
    assert out["roofline"]["kernel"] == kernel and "e2e" not in out


def test_reference_arm_line():
    env = dict(os.environ, LFGPU_REF_N="300")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["value"] > 0 and out["cpu_baseline"]["kind"] == "port" and out["cpu_baseline"]["cores"] == 1
    assert out["e2e"] == {"value": out["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert out["cpu_baseline"]["all_cores"]["cores"] >= 1


def test_default_line_carries_the_other_configs():
    out = run_bench("--steps", "3", "--no-cpu-baseline", "--no-e2e")
    check_line(out, 3)
    oc = out["other_configs"]
    assert set(oc) == {"c2", "c3", "c4"}
    for name, rec in oc.items():
        assert "error" not in rec, rec
        assert rec["ms_per_step"] > 0 and 0 < rec["roofline"]["frac"] and {"what", "value", "expected"} <= set(rec["check"])
    assert oc["c2"]["roofline"]["kernel"] == "k_assemble_p1_rows"
