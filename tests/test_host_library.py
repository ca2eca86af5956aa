"""CPU-side checks of the product library: it loads, exports every symbol of include/lfgpu.h, fails loudly without a
GPU, and its host-side reference-element tables agree with the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

import lehrfempp_b200 as lf
from lehrfempp_b200 import api
from oracle import lfo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    lf.build_library()
    return api._lib()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "lfgpu.h")).read()
    names = sorted(set(re.findall(r"\b(lfgpu_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 40
    raw = ctypes.CDLL(lf.library_path())
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, missing
    # and the binding declares a prototype for each of them
    assert set(names) <= set(lib._exported) | {"lfgpu_pattern_outer_device", "lfgpu_pattern_inner_device"}


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lf.LfgpuError) as e:
        lf.Context(0)
    assert e.value.code == -3  # LFGPU_ERR_NO_DEVICE


def test_product_does_not_reference_oracle():
    # the product path must never import / link the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lehrfempp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"liblfo_oracle|^\s*from oracle|^\s*import oracle|#include\s+\"[^\"]*(lfo_|oracle)", txt, re.M), f
    txt = open(os.path.join(ROOT, "include", "lfgpu.h")).read()
    assert "oracle" not in txt


@pytest.mark.parametrize("degree", [1, 2, 3])
@pytest.mark.parametrize("cell_type", [3, 4])
def test_fe_tables_match_oracle(lib, degree, cell_type):
    for qdeg in (None, 2, 4, 6, 8):
        qr = None if qdeg is None else lf.QuadRule(*lfo.quad_rule(cell_type, qdeg))
        phi, grad = lf.fe_tabulate(degree, cell_type, qr)
        pts = qr.points if qr is not None else lfo.quad_rule(cell_type, 2 * degree)[0]
        ophi, ograd, _ = lfo.eval_fe(degree, cell_type, pts)
        assert phi.shape == ophi.shape
        assert np.abs(phi - ophi).max() < 5e-15
        assert np.abs(grad - ograd).max() < 5e-14


@pytest.mark.parametrize("cell_type", [3, 4])
@pytest.mark.parametrize("degree", list(range(0, 13)))
def test_default_rules_match_oracle(lib, cell_type, degree):
    if cell_type == 4 and degree > 10:
        pytest.skip("capacity")
    q = lf.default_quad_rule(cell_type, degree)
    p, w = lfo.quad_rule(cell_type, degree)
    if cell_type == 3:
        assert np.array_equal(q.points, p) and np.array_equal(q.weights, w)  # literal tables: bitwise
    else:
        # Gauss-Legendre from two independent Newton iterations: equal up to the last bit
        assert np.abs(q.points - p).max() <= 2.3e-16 and np.abs(q.weights - w).max() <= 2.3e-16


def test_row_ranges_cover_and_balance():
    """Row-block partition of the owner_rows multi-GPU mode: blocks cover all rows exactly once with ~equal nnz."""
    import torch
    from lehrfempp_b200.distributed import row_ranges
    rng = np.random.default_rng(3)
    lens = rng.integers(0, 12, size=10007)
    outer = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32)
    for world in (1, 2, 3, 8):
        b = row_ranges(outer, world)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == len(lens)
        assert all(b[k] <= b[k + 1] for k in range(world))
        nnz = [int(outer[b[k + 1]] - outer[b[k]]) for k in range(world)]
        assert sum(nnz) == int(outer[-1])
        assert max(nnz) - min(nnz) <= 2 * 12
    # degenerate: fewer rows than ranks
    b = row_ranges(torch.tensor([0, 3, 5], dtype=torch.int32), 4)
    assert b[0] == 0 and b[-1] == 2 and all(b[k] <= b[k + 1] for k in range(4))


def test_c_abi_header_is_plain_c():
    """include/lfgpu.h is the FFI boundary: it must compile as C99 (no C++ in the signatures) and stand alone."""
    import subprocess
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "lfgpu.h")
    out = subprocess.run(["gcc", "-x", "c", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", hdr], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
