"""P2 / P3 row kernels under every plan format / copy-out / L2-hint switch (read once per process, hence subprocesses):
tests/plan_variants_check.py against the oracle and the generic kernel."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [
    {},  # defaults: compact plans, edge rows on their own coordinate copy, bulk copy-out, no L2 hints
    {"LFGPU_EDGE_ORDER": "0", "LFGPU_P2_COMPACT": "0"},
    {"LFGPU_P2_COMPACT": "0"},
    {"LFGPU_P2_COMPACT": "1", "LFGPU_L2_HINTS": "1"},
    {"LFGPU_P2_COMPACT": "v", "LFGPU_EDGE_PFC": "50", "LFGPU_EDGE_ORDER": "2"},  # 2: the coordinate copy on every mesh
    {"LFGPU_P2_BULK": "0"},
    {"LFGPU_P2_BULK": "0", "LFGPU_P2_COMPACT": "e", "LFGPU_EDGE_ORDER": "0"},
]


@pytest.mark.parametrize("variant", VARIANTS, ids=lambda v: ",".join("%s=%s" % kv for kv in sorted(v.items())) or "defaults")
def test_plan_variants(variant):
    env = dict(os.environ, **variant)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "plan_variants_check.py")], capture_output=True, text=True, timeout=600, env=env)
    assert "PLAN_VARIANTS_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


LOAD_VARIANTS = [{}, {"LFGPU_LOAD_FAN": "0"}, {"LFGPU_LOAD_FAN": "0", "LFGPU_LOAD_TWOPASS": "0"}]


@pytest.mark.parametrize("variant", LOAD_VARIANTS, ids=lambda v: ",".join("%s=%s" % kv for kv in sorted(v.items())) or "defaults")
def test_load_vector_variants(variant):
    env = dict(os.environ, **variant)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "load_variants_check.py")], capture_output=True, text=True, timeout=600, env=env)
    assert "LOAD_VARIANTS_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
