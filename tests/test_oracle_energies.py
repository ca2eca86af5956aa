"""Pins the oracle's whole P1 pipeline with VARIABLE coefficients on regularly refined HYBRID meshes against the reference's
known answers: lib/lf/uscalfe/test/full_gal_tests.cc:49-236 (tests a_dir_dbg_1 .. 9).

There the energy v^T A v of the piecewise linear interpolant of v is computed on GenerateHybrid2DTestMesh(0, 1/3) (the unit
square) after REFLEV = 6 uniform refinements, with A from ReactionDiffusionElementMatrixProvider (default rules, coefficients
as MeshFunctionGlobal lambdas evaluated at the quadrature points), and compared with the exact energy to 1e-4 / 2e-3.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import lfo

REFLEV = 6

V_LIN = lambda x, y: 2.0 * x + y  # noqa: E731
V_EXP = lambda x, y: np.exp(x * y)  # noqa: E731
ONE = lambda x, y: np.ones_like(x)  # noqa: E731
ZERO = lambda x, y: np.zeros_like(x)  # noqa: E731
R2 = lambda x, y: 1.0 + x * x + y * y  # noqa: E731

# (id, v, alpha, gamma, expected energy, tolerance) -- full_gal_tests.cc:49-218
CASES = [
    ("1: v = 1, alpha = 1, gamma = 0", ONE, ONE, ZERO, 0.0, 1e-10),
    ("2: alpha = 1 + x", V_LIN, lambda x, y: 1.0 + x, ZERO, 7.5, 1e-4),
    ("3: alpha = 1 + x^2 + y^2", V_LIN, R2, ZERO, 25.0 / 3.0, 1e-4),
    ("4: alpha = 0, gamma = 1", V_LIN, ZERO, ONE, 8.0 / 3.0, 1e-4),
    ("5: gamma = 1 / (1 + x^2 + y^2)", V_LIN, ZERO, lambda x, y: 1.0 / R2(x, y), 1.42447, 1e-4),
    ("6: v = exp(xy), alpha = 1", V_EXP, ONE, ZERO, 1.59726, 2e-3),
    ("7: v = exp(xy), alpha = 1 + x^2 + y^2", V_EXP, R2, ZERO, 3.39057, 2e-3),
    ("8: v = exp(xy), alpha and gamma variable", V_EXP, R2, lambda x, y: 1.0 / R2(x, y), 4.44757, 2e-3),
]


@pytest.fixture(scope="module")
def finest(golden_meshes):
    m = lfo.Mesh.from_golden(golden_meshes["0"], 1.0 / 3.0)
    for _ in range(REFLEV):
        m = m.refine_regular()
    assert m.n_cells == 9 * 4 ** REFLEV
    return m, m.qp_coords(2, 2), m.export()["node_coords"]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_energy_of_interpolant(finest, case):
    _, v, alpha, gamma, expected, tol = case
    m, qp, nodes = finest
    a_tab = alpha(qp[..., 0], qp[..., 1])
    g_tab = gamma(qp[..., 0], qp[..., 1])
    outer, inner, vals, shape, _ = m.assemble_rd(1, lfo.coeff.table(a_tab), lfo.coeff.table(g_tab))
    A = sp.csc_matrix((vals, inner, outer), shape=shape)
    vv = v(nodes[:, 0], nodes[:, 1])  # P1: dof = node, nodal interpolation
    energy = vv @ (A @ vv)
    assert abs(energy - expected) <= tol


def test_energy_with_tensor_coefficient(finest):
    """a_dir_dbg_9: alpha = diag(1 + y, 1 + x), v = 2x + y -> 7.5 (2e-3)"""
    m, qp, nodes = finest
    alpha = lfo.coeff.callback2x2(lambda x, y: [[1.0 + y, 0.0], [0.0, 1.0 + x]])  # a MeshFunctionGlobal lambda, called per point
    outer, inner, vals, shape, _ = m.assemble_rd(1, alpha, lfo.coeff.const(0.0))
    A = sp.csc_matrix((vals, inner, outer), shape=shape)
    vv = V_LIN(nodes[:, 0], nodes[:, 1])
    assert abs(vv @ (A @ vv) - 7.5) <= 2e-3
