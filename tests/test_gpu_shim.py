"""Runs the C++ test of the header-only shim (include/lf_gpu_shim.hpp) on the GPU; on CPU only checks it builds."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")


def _build():
    import lehrfempp_b200 as lf
    lf.build_library()
    subprocess.check_call(["make", "-C", CPP, "-s"])
    return os.path.join(CPP, "shim_test")


def test_shim_builds_and_fails_loudly_without_gpu():
    import torch
    exe = _build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 77 and "lfgpu needs a CUDA device" in out.stdout  # no CPU fallback


@pytest.mark.gpu
def test_shim_matches_reference_call_sequence():
    exe = _build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "SHIM_TEST_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
