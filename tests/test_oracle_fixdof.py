"""CPU checks of the oracle's FixFlaggedSolutionComponents restatement (assemble/fix_dof.h:86-138).

The reference's own test for it (lib/lf/assemble/test/fix_dof_tests.cc) solves a system and checks that the fixed
components come out with their prescribed values; the same property is checked here, plus the algebraic definition
A' = [[A_ff, 0], [0, I]], b' = [b_f - A_fd xhat_d, xhat_d] evaluated with dense numpy.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import lfo


def random_fixed(n, seed, frac=0.25):
    rng = np.random.default_rng(seed)
    fixed = (rng.random(n) < frac).astype(np.uint8)
    vals = rng.standard_normal(n)
    return fixed, vals


@pytest.mark.parametrize("degree", [1, 2, 3])
@pytest.mark.parametrize("csr", [False, True])
def test_fix_matches_dense_definition(degree, csr):
    om = lfo.Mesh.hybrid(5, 0.2, 12345)
    n = om.num_dofs(degree)
    fixed, xhat = random_fixed(n, 7 + degree)
    outer, inner, vals, rhs = om.assemble_fixed(degree, 1.5, 0.5, 2.0, fixed, xhat, csr=csr)
    A_fix = (sp.csr_matrix if csr else sp.csc_matrix)((vals, inner, outer), shape=(n, n)).toarray()
    o0, i0, v0, shape, _ = om.assemble_rd(degree, lfo.coeff.const(1.5), lfo.coeff.const(0.5), csr=csr)
    A0 = (sp.csr_matrix if csr else sp.csc_matrix)((v0, i0, o0), shape=shape).toarray()
    b0, _ = om.assemble_load(degree, lfo.coeff.const(2.0))
    f = fixed.astype(bool)
    xh = np.where(f, xhat, 0.0)
    A_exp = A0.copy()
    A_exp[f, :] = 0.0
    A_exp[:, f] = 0.0
    A_exp[f, f] = 1.0
    b_exp = b0 - A0 @ xh
    b_exp[f] = xhat[f]
    assert np.abs(A_fix - A_exp).max() <= 1e-13 * np.abs(A0).max()
    assert np.abs(rhs - b_exp).max() <= 1e-12 * max(1.0, np.abs(b_exp).max())
    # erased triplets are gone from the compressed matrix: fixed rows / columns hold the unit diagonal only
    M = (sp.csr_matrix if csr else sp.csc_matrix)((np.ones_like(vals), inner, outer), shape=(n, n)).tocoo()
    in_fixed = f[M.row] | f[M.col]
    assert np.all(M.row[in_fixed] == M.col[in_fixed])
    assert in_fixed.sum() == f.sum()
    assert len(vals) < len(v0)


def test_solution_takes_prescribed_values():
    # fix_dof_tests.cc: after the elimination the solve returns the prescribed values on the fixed dofs
    om = lfo.Mesh.tp_tria(6, 5)
    ex = om.export()
    xy = ex["node_coords"]
    bd = (np.isclose(xy[:, 0], 0) | np.isclose(xy[:, 0], 1) | np.isclose(xy[:, 1], 0) | np.isclose(xy[:, 1], 1)).astype(np.uint8)
    g = 1.0 + xy[:, 0]
    outer, inner, vals, rhs = om.assemble_fixed(1, 1.0, 0.0, 1.0, bd, g)
    n = om.num_dofs(1)
    x = spla.spsolve(sp.csc_matrix((vals, inner, outer), shape=(n, n)), rhs)
    assert np.abs(x[bd == 1] - g[bd == 1]).max() <= 1e-14
    assert np.all(x[bd == 0] > 1.0)  # -Laplace u = 1 >= 0 with boundary data >= 1: maximum principle


def tridiag_case():
    """The reference's known-answer test (assemble/test/coomatrix_tests.cc:181-237): 10x10 tridiag(-1, 2, -1), b = 1..10,
    components 2, 4, 8 fixed to -1, -2, -3; expected solution listed at coomatrix_tests.cc:235."""
    n = 10
    rows, cols, vals = [], [], []
    for k in range(n):
        if k > 0:
            rows.append(k), cols.append(k - 1), vals.append(-1.0)
        if k < n - 1:
            rows.append(k), cols.append(k + 1), vals.append(-1.0)
        rows.append(k), cols.append(k), vals.append(2.0)
    b = np.arange(1.0, n + 1)
    fixed = np.zeros(n, np.uint8)
    fixed[[2, 4, 8]] = 1
    xhat = np.ones(n)
    xhat[[2, 4, 8]] = [-1.0, -2.0, -3.0]
    exact = np.array([1, 1, -1, 0.5, -2, 7.75, 11.5, 8.25, -3, 3.5])
    return n, rows, cols, vals, b, fixed, xhat, exact


def test_reference_known_answer():
    n, rows, cols, vals, b, fixed, xhat, exact = tridiag_case()
    outer, inner, v, rhs = lfo.fix_coo(n, rows, cols, vals, fixed, xhat, b)
    x = spla.spsolve(sp.csc_matrix((v, inner, outer), shape=(n, n)), rhs)
    assert np.linalg.norm(x - exact) <= 1e-12  # the reference's own tolerance
    assert len(v) == 28 - 4 * 3  # 28 entries; each fixed dof loses its 2 row and 2 column off-diagonals, keeps a unit diagonal


def test_reference_known_answer_alt():
    # coomatrix_tests.cc:123-179: the row-only variant has the same solution
    n, rows, cols, vals, b, fixed, xhat, exact = tridiag_case()
    outer, inner, v, rhs = lfo.fix_coo(n, rows, cols, vals, fixed, xhat, b, alt=True)
    x = spla.spsolve(sp.csc_matrix((v, inner, outer), shape=(n, n)), rhs)
    assert np.linalg.norm(x - exact) <= 1e-12
    assert len(v) == 28 - 2 * 3  # only the two off-diagonals of each fixed ROW go
    assert np.array_equal(rhs[fixed == 0], b[fixed == 0])


def test_alt_matches_dense_definition():
    om = lfo.Mesh.hybrid(5, 0.2, 12345)
    n = om.num_dofs(2)
    fixed, xhat = random_fixed(n, 3)
    outer, inner, vals, rhs = om.assemble_fixed(2, 1.5, 0.5, 2.0, fixed, xhat, alt=True)
    o0, i0, v0, shape, _ = om.assemble_rd(2, lfo.coeff.const(1.5), lfo.coeff.const(0.5))
    A0 = sp.csc_matrix((v0, i0, o0), shape=shape).toarray()
    b0, _ = om.assemble_load(2, lfo.coeff.const(2.0))
    f = fixed.astype(bool)
    A0[f, :] = 0.0
    A0[f, f] = 1.0
    b0[f] = xhat[f]
    assert np.abs(sp.csc_matrix((vals, inner, outer), shape=(n, n)).toarray() - A0).max() <= 1e-13 * np.abs(v0).max()
    assert np.array_equal(rhs, b0)


def test_nothing_fixed_is_identity():
    om = lfo.Mesh.tp_quad(4, 3)
    n = om.num_dofs(2)
    outer, inner, vals, rhs = om.assemble_fixed(2, 1.0, 1.0, 1.0, np.zeros(n, np.uint8), np.zeros(n))
    o0, i0, v0, _, _ = om.assemble_rd(2, lfo.coeff.const(1.0), lfo.coeff.const(1.0))
    b0, _ = om.assemble_load(2, lfo.coeff.const(1.0))
    assert np.array_equal(outer, o0) and np.array_equal(inner, i0) and np.array_equal(vals, v0) and np.array_equal(rhs, b0)


def test_reference_known_answer_lse():
    """coomatrix_tests.cc:78-122 (fix_dof_test): FixSolutionComponentsLse with {2: -1, 4: -2, 8: -3} on tridiag(-1, 2, -1),
    b = 1..10 -> x = (1, 1, -1, 0.5, -2, 7.75, 11.5, 8.25, -3, 3.5) to 1e-12; values of a repeated index add up
    (fix_dof.h:268)."""
    n, rows, cols, vals, b, fixed, xhat, exact = tridiag_case()
    outer, inner, v, rhs = lfo.fix_coo_lse(n, rows, cols, vals, [(2, -1.0), (4, -2.0), (8, -3.0)], b)
    x = spla.spsolve(sp.csc_matrix((v, inner, outer), shape=(n, n)), rhs)
    assert np.linalg.norm(x - exact) <= 1e-12
    o2, i2, v2, rhs2 = lfo.fix_coo(n, rows, cols, vals, fixed, xhat, b, alt=True)  # same as the row-only variant with flags
    assert np.array_equal(outer, o2) and np.array_equal(inner, i2) and np.array_equal(v, v2) and np.array_equal(rhs, rhs2)
    _, _, _, rhs3 = lfo.fix_coo_lse(n, rows, cols, vals, [(2, -0.25), (4, -2.0), (2, -0.75), (8, -3.0)], b)
    assert np.array_equal(rhs3, rhs)
