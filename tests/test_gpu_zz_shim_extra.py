"""The shim checks written after the round's GPU minutes were spent (`shim_test extra`: returning forms of AssembleMatrixLocally /
AssembleVectorLocally, lf::fe providers, FixSolutionComponentsLse) -- separate from tests/test_gpu_shim.py and last in the
alphabet so that their first run on a B200 cannot hide results that are already established."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_extra_checks():
    exe = os.path.join(ROOT, "tests", "cpp", "shim_test")
    subprocess.check_call(["make", "-C", os.path.dirname(exe), "-s"])
    out = subprocess.run([exe, "extra"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "SHIM_TEST_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
