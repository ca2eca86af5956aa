"""P2 / P3 row kernels with the general-valence vertex plans (LFGPU_P2_GENERAL=1, LFGPU_P3_GENERAL=1; rows_p2_core.h,
rows_p3_core.h) on unstructured meshes, in a subprocess because the switch is read once per process.  The arithmetic is checked on the CPU
(tests/test_p2_rows_core.py, tests/test_p3_rows_core.py); this covers the CUDA wrapper, which has not run on a B200 yet (written after the round's GPU
minutes were spent) -- hence opt-in and last in the alphabet."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_p2_general_valence_rows():
    env = dict(os.environ, LFGPU_P2_GENERAL="1", LFGPU_P3_GENERAL="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "p2_general_check.py")], capture_output=True, text=True, timeout=300, env=env)
    assert "P2_GENERAL_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
