"""GPU parity: lfgpu_fix_flagged_solution_components against the oracle's FixFlaggedSolutionComponents + makeSparse
(assemble/fix_dof.h:86-138).  Compacted index arrays bit-exact, values and right-hand side within 1e-12 (max-norm)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import lfo
from tests.helpers import rel_max_err

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(scope="module")
def lf():
    import lehrfempp_b200 as lf
    return lf


@pytest.fixture(scope="module")
def ctx(lf):
    c = lf.Context(0)
    yield c
    c.close()


def meshes(ctx, kind):
    if kind == "hybrid":
        return lfo.Mesh.hybrid(9, 0.2, 12345), ctx.mesh_hybrid(9, 0.2, 12345)
    if kind == "tp_tria":
        return lfo.Mesh.tp_tria(12, 9), ctx.mesh_tp_tria(12, 9)
    return lfo.Mesh.tp_quad(7, 11), ctx.mesh_tp_quad(7, 11)


def gpu_system(ctx, lf, gm, degree, major, alpha, gamma, f):
    dm = gm.dofmap_lagrange(degree)
    pat = dm.symbolic(major=major)
    vals = pat.assemble_reaction_diffusion(degree, lf.Coeff.const(alpha), lf.Coeff.const(gamma))
    rhs = dm.assemble_load(degree, lf.Coeff.const(f))
    return dm, pat, vals, rhs


@pytest.mark.parametrize("kind", ["hybrid", "tp_tria", "tp_quad"])
@pytest.mark.parametrize("degree", [1, 2, 3])
@pytest.mark.parametrize("csr", [False, True])
def test_fix_compact_matches_oracle(ctx, lf, kind, degree, csr):
    om, gm = meshes(ctx, kind)
    n = om.num_dofs(degree)
    rng = np.random.default_rng(100 * degree + len(kind))
    fixed = (rng.random(n) < 0.3).astype(np.uint8)
    xhat = rng.standard_normal(n)
    o_outer, o_inner, o_vals, o_rhs = om.assemble_fixed(degree, 1.5, 0.5, 2.0, fixed, xhat, csr=csr)
    dm, pat, vals, rhs = gpu_system(ctx, lf, gm, degree, lf.ROW_MAJOR if csr else lf.COL_MAJOR, 1.5, 0.5, 2.0)
    assert dm.num_dofs == n
    outer, inner, cvals, kept = pat.fix_flagged_solution_components(vals, rhs, ctx.to_device(fixed), ctx.to_device(xhat), compact=True)
    assert kept == len(o_vals) and kept < pat.nnz
    assert np.array_equal(outer.to_host(), o_outer)
    assert np.array_equal(inner.to_host()[:kept], o_inner)
    assert rel_max_err(cvals.to_host()[:kept], o_vals) <= TOL
    assert rel_max_err(rhs.to_host(), o_rhs) <= TOL
    # the in-place matrix is the same operator with explicit zeros in the erased slots
    p_outer, p_inner = pat.download()
    fmt = sp.csr_matrix if csr else sp.csc_matrix
    A_inplace = fmt((vals.to_host(), p_inner, p_outer), shape=(n, n))
    A_oracle = fmt((o_vals, o_inner, o_outer), shape=(n, n))
    assert abs(A_inplace - A_oracle).max() <= TOL * np.abs(o_vals).max()


@pytest.mark.parametrize("degree", [1, 3])
@pytest.mark.parametrize("csr", [False, True])
def test_fix_alt_matches_oracle(ctx, lf, degree, csr):
    # FixFlaggedSolutionCompAlt (fix_dof.h:181-218): unit rows only
    om, gm = meshes(ctx, "hybrid")
    n = om.num_dofs(degree)
    rng = np.random.default_rng(17 + degree)
    fixed = (rng.random(n) < 0.3).astype(np.uint8)
    xhat = rng.standard_normal(n)
    o_outer, o_inner, o_vals, o_rhs = om.assemble_fixed(degree, 1.5, 0.5, 2.0, fixed, xhat, csr=csr, alt=True)
    dm, pat, vals, rhs = gpu_system(ctx, lf, gm, degree, lf.ROW_MAJOR if csr else lf.COL_MAJOR, 1.5, 0.5, 2.0)
    outer, inner, cvals, kept = pat.fix_flagged_solution_components(vals, rhs, ctx.to_device(fixed), ctx.to_device(xhat), compact=True,
                                                                    alt=True)
    assert kept == len(o_vals)
    assert np.array_equal(outer.to_host(), o_outer)
    assert np.array_equal(inner.to_host()[:kept], o_inner)
    assert rel_max_err(cvals.to_host()[:kept], o_vals) <= TOL
    assert rel_max_err(rhs.to_host(), o_rhs) <= TOL


def test_fix_in_place_only(ctx, lf):
    om, gm = meshes(ctx, "hybrid")
    n = om.num_dofs(2)
    rng = np.random.default_rng(5)
    fixed = (rng.random(n) < 0.2).astype(np.uint8)
    xhat = rng.standard_normal(n)
    o_outer, o_inner, o_vals, o_rhs = om.assemble_fixed(2, 1.0, 0.0, 1.0, fixed, xhat, csr=True)
    dm, pat, vals, rhs = gpu_system(ctx, lf, gm, 2, lf.ROW_MAJOR, 1.0, 0.0, 1.0)
    assert pat.fix_flagged_solution_components(vals, rhs, ctx.to_device(fixed), ctx.to_device(xhat)) is None
    p_outer, p_inner = pat.download()
    A = sp.csr_matrix((vals.to_host(), p_inner, p_outer), shape=(n, n))
    assert abs(A - sp.csr_matrix((o_vals, o_inner, o_outer), shape=(n, n))).max() <= TOL * np.abs(o_vals).max()
    assert rel_max_err(rhs.to_host(), o_rhs) <= TOL
    # fix_dof contract: the solve reproduces the prescribed values
    x = spla.spsolve(A.tocsc(), rhs.to_host())
    assert np.abs(x[fixed == 1] - xhat[fixed == 1]).max() <= 1e-12


def test_nothing_and_everything_fixed(ctx, lf):
    om, gm = meshes(ctx, "tp_tria")
    n = om.num_dofs(1)
    dm, pat, vals, rhs = gpu_system(ctx, lf, gm, 1, lf.COL_MAJOR, 1.0, 1.0, 1.0)
    v0, b0 = vals.to_host(), rhs.to_host()
    outer, inner, cvals, kept = pat.fix_flagged_solution_components(vals, rhs, ctx.zeros(n, np.uint8), ctx.zeros(n), compact=True)
    p_outer, p_inner = pat.download()
    assert kept == pat.nnz and np.array_equal(outer.to_host(), p_outer) and np.array_equal(inner.to_host(), p_inner)
    assert np.array_equal(cvals.to_host(), v0) and np.array_equal(rhs.to_host(), b0) and np.array_equal(vals.to_host(), v0)
    xhat = np.linspace(-1.0, 1.0, n)
    outer, inner, cvals, kept = pat.fix_flagged_solution_components(vals, rhs, ctx.to_device(np.ones(n, np.uint8)), ctx.to_device(xhat),
                                                                    compact=True)
    assert kept == n
    assert np.array_equal(outer.to_host(), np.arange(n + 1, dtype=np.int32))
    assert np.array_equal(inner.to_host()[:n], np.arange(n, dtype=np.int32))
    assert np.array_equal(cvals.to_host()[:n], np.ones(n)) and np.array_equal(rhs.to_host(), xhat)


def test_rectangular_pattern_rejected(ctx, lf):
    # fix_dof.h:90 "Matrix must be square!"
    gm = ctx.mesh_tp_tria(3, 3)
    trial, test = gm.dofmap_lagrange(1), gm.dofmap_lagrange(2)
    try:
        pat = test.symbolic(trial=trial)
    except TypeError:
        pytest.skip("binding has no rectangular symbolic pass")
    n = max(pat.rows, pat.cols)
    with pytest.raises(lf.LfgpuError):
        pat.fix_flagged_solution_components(ctx.zeros(pat.nnz), ctx.zeros(n), ctx.zeros(n, np.uint8), ctx.zeros(n))
