"""Pins the oracle (oracle/, CPU restatement) against the reference's own literal goldens and known answers.

Sources (paths relative to the reference checkout):
  lib/lf/assemble/test/assembly_tests.cc:219-286   10x10 golden + vector (TestAssembler / TestVectorAssembler)
  lib/lf/assemble/test/assembly_tests.cc:323-458   36x36 golden, two dofs per edge with orientation reversal
  lib/lf/uscalfe/test/lagr_fe_tests.cc:774-882     a^T A b = 7911/8, 81, 1996731/280 (O1/O2/O3)
  lib/lf/uscalfe/test/loc_comp_test.cc:46-152      mass row sums, default-vs-explicit quadrature rules
  lib/lf/uscalfe/test/lagr_fe_tests.cc:58-370      cardinality of shape functions at evaluation nodes
  lib/lf/quad/test/make_quad_rule_tests.cc:30-135  exactness of the quadrature rules
  lib/lf/mesh/utils/test/tp_triag_mesh_builder_tests.cc:21-41  entity counts of the structured builder
  lib/lf/uscalfe/test/bvp_fe_tests.cc:30-60        Neumann matrix: row and column sums vanish
"""
import math

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import lfo


def mesh0(golden_meshes, scale=1.0):
    return lfo.Mesh.from_golden(golden_meshes["0"], scale)


def test_mesh0_counts(golden_meshes):
    m = mesh0(golden_meshes)
    assert (m.n_nodes, m.n_edges, m.n_cells) == (10, 18, 9)


def test_golden_10x10_matrix_and_vector(golden_meshes, assembly_goldens):
    m = mesh0(golden_meshes)
    dofh = lfo.DofHandler(m, n_pt=1)
    assert dofh.num_dofs == 10
    codim, idx = dofh.dof_entities()
    assert np.all(codim == 2)
    g = next(e for e in assembly_goldens["ref_mat_10"] if e["line"] < 300)
    ref = np.array(g["row_major"]).reshape(10, 10)
    A = dofh.test_matrix(0)
    assert np.array_equal(A, ref[np.ix_(idx, idx)])
    v = dofh.test_vector()
    assert np.array_equal(v, np.diag(ref)[idx])


def test_golden_36x36_edge_dofs(golden_meshes, assembly_goldens):
    m = mesh0(golden_meshes)
    dofh = lfo.DofHandler(m, n_seg=2)
    assert dofh.num_dofs == 36
    codim, _ = dofh.dof_entities()
    assert np.all(codim == 1)
    ref = np.array(assembly_goldens["ref_mat_36"][0]["row_major"]).reshape(36, 36)
    A = dofh.test_matrix(1)
    assert np.array_equal(A, ref)


def bilinear(m, degree, alpha, gamma, a, b, qr):
    outer, inner, vals, shape, _ = m.assemble_rd(degree, alpha, gamma, qr_tria=qr, qr_quad=qr)
    A = sp.csc_matrix((vals, inner, outer), shape=shape)
    av = m.nodal_projection(degree, a)
    bv = m.nodal_projection(degree, b)
    return av @ (A @ bv)


def test_bilinear_form_known_answers(golden_meshes):
    m = mesh0(golden_meshes)
    c = lfo.coeff
    # O1: tensor alpha = [1 x; y xy], gamma = xy, a = 1+x+2y, b = 3x, qr degree 4 -> 7911/8
    v1 = bilinear(m, 1, c.builtin(100), c.builtin(4), c.builtin(9), c.builtin(10), 4)
    assert abs(v1 - 7911.0 / 8.0) < 1e-10 * 1000
    assert m.num_dofs(1) == 10
    # O2: alpha = x, gamma = xy, a = x^2+y^2, b = x^2-y^2, qr degree 6 -> 81
    v2 = bilinear(m, 2, c.builtin(5), c.builtin(4), c.builtin(8), c.builtin(7), 6)
    assert abs(v2 - 81.0) < 1e-10 * 100
    assert m.num_dofs(2) == 30
    # O3: alpha = y, gamma = xy, a = x^3+y^3, b = x y^2, qr degree 8 -> 1996731/280
    v3 = bilinear(m, 3, c.builtin(6), c.builtin(4), c.builtin(11), c.builtin(12), 8)
    assert abs(v3 - 1996731.0 / 280.0) < 1e-10 * 1e4
    assert m.num_dofs(3) == 61


def test_callback_coefficients_match_builtin(golden_meshes):
    m = mesh0(golden_meshes)
    c = lfo.coeff
    ref = m.assemble_rd(2, c.builtin(100), c.builtin(4))
    got = m.assemble_rd(2, c.callback2x2(lambda x, y: [[1, x], [y, x * y]]), c.callback(lambda x, y: x * y))
    assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1])
    assert np.allclose(ref[2], got[2], rtol=0, atol=1e-15 * np.abs(ref[2]).max())


@pytest.mark.parametrize("degree", [1, 2, 3])
def test_mass_matrix_row_sums_equal_volume(golden_meshes, degree):
    # loc_comp_test.cc:46-84 : alpha = gamma = 1 => sum of all entries of the element MASS part = |K|.
    m = mesh0(golden_meshes)
    c = lfo.coeff
    E = m.element_matrices(degree, c.const(0.0), c.const(1.0))
    ex = m.export()
    for k in range(m.n_cells):
        xy = ex["cell_coords"][k]
        nv = 3 if ex["cell_type"][k] == 3 else 4
        x, y = xy[:nv, 0], xy[:nv, 1]
        area = 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
        assert abs(E[k].sum() - area) < 1e-13 * max(1.0, area)


def test_default_rule_equals_explicit_rule(golden_meshes):
    # loc_comp_test.cc:86-152: default rules for O1 are make_QuadRule(., 2)
    m = mesh0(golden_meshes)
    c = lfo.coeff
    a, g = c.builtin(8), c.builtin(2)
    E0 = m.element_matrices(1, a, g)
    E1 = m.element_matrices(1, a, g, qr_tria=2, qr_quad=2)
    assert np.array_equal(E0, E1)
    f = c.builtin(7)
    v0, _ = m.assemble_load(1, f)
    v1, _ = m.assemble_load(1, f, qr_tria=2, qr_quad=2)
    assert np.array_equal(v0, v1)


def test_missing_rule_throws(golden_meshes):
    # loc_comp_test.cc:154-183: only a triangle rule given, mesh 0 has quads -> Eval throws LfException
    m = mesh0(golden_meshes)
    c = lfo.coeff
    with pytest.raises(lfo.OracleError):
        m.assemble_rd(1, c.const(1.0), c.const(0.0), qr_tria=2, qr_quad=-1)


@pytest.mark.parametrize("degree", [1, 2, 3])
@pytest.mark.parametrize("ref_el", [3, 4])
def test_shape_functions_cardinal_and_gradients(degree, ref_el):
    _, _, nodes = lfo.eval_fe(degree, ref_el, np.zeros((2, 1)))
    phi, _, _ = lfo.eval_fe(degree, ref_el, nodes)
    assert np.allclose(phi, np.eye(phi.shape[0]), atol=1e-14)
    rng = np.random.default_rng(1)
    pts = rng.random((2, 7)) * 0.5
    phi, grad, _ = lfo.eval_fe(degree, ref_el, pts)
    assert np.allclose(phi.sum(axis=0), 1.0, atol=1e-13)  # partition of unity
    h = 1e-6
    for d in range(2):
        e = np.zeros((2, 1))
        e[d] = h
        fd = (lfo.eval_fe(degree, ref_el, pts + e)[0] - lfo.eval_fe(degree, ref_el, pts - e)[0]) / (2 * h)
        assert np.allclose(grad[:, d::2], fd, atol=1e-7)


def test_o1_literal_values():
    # lagr_fe_tests.cc: P1 triangle shape functions at the barycentre and gradients
    phi, grad, _ = lfo.eval_fe(1, 3, np.array([[1 / 3.0], [1 / 3.0]]))
    assert np.allclose(phi[:, 0], 1 / 3.0)
    assert np.array_equal(grad, np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]]))


@pytest.mark.parametrize("degree", list(range(1, 13)))
def test_tria_rules_exact(degree):
    pts, w = lfo.quad_rule(3, degree)
    assert abs(w.sum() - 0.5) < 1e-14
    for i in range(degree + 1):
        for j in range(degree + 1 - i):
            exact = math.factorial(i) * math.factorial(j) / math.factorial(i + j + 2)
            assert abs(np.dot(w, pts[0] ** i * pts[1] ** j) - exact) < 1e-14


def test_tria_degree3_is_degree4_rule():
    p3, w3 = lfo.quad_rule(3, 3)
    p4, w4 = lfo.quad_rule(3, 4)
    assert np.array_equal(p3, p4) and np.array_equal(w3, w4)


@pytest.mark.parametrize("degree", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_quad_rules_tensor_gauss(degree):
    pts, w = lfo.quad_rule(4, degree)
    n = degree // 2 + 1
    assert w.size == n * n
    x, wx = np.polynomial.legendre.leggauss(n)
    x, wx = 0.5 * (x + 1), 0.5 * wx
    # point i*n + j has x0 = p_i, x1 = p_j (make_quad_rule.cc:28-37)
    assert np.allclose(pts[0].reshape(n, n), np.repeat(x[:, None], n, 1), atol=2e-16)
    assert np.allclose(pts[1].reshape(n, n), np.repeat(x[None, :], n, 0), atol=2e-16)
    assert np.allclose(w.reshape(n, n), np.outer(wx, wx), atol=2e-16)
    for i in range(2 * n):
        for j in range(2 * n):
            assert abs(np.dot(w, pts[0] ** i * pts[1] ** j) - 1.0 / ((i + 1) * (j + 1))) < 1e-14


def test_tp_builder_counts():
    m = lfo.Mesh.tp_tria(2, 2)
    assert (m.n_cells, m.n_edges, m.n_nodes) == (8, 16, 9)
    ex = m.export()
    # all edges are supplied explicitly: horizontal (i outer, j inner), vertical, diagonal
    assert ex["edge_nodes"][0].tolist() == [0, 1] and ex["edge_nodes"][1].tolist() == [3, 4]
    assert ex["cell_nodes"][0].tolist() == [0, 4, 3, lfo.NIL]
    assert ex["cell_nodes"][1].tolist() == [0, 1, 4, lfo.NIL]
    q = lfo.Mesh.tp_quad(3, 2)
    assert (q.n_cells, q.n_edges, q.n_nodes) == (6, 17, 12)


def test_neumann_matrix_row_col_sums_zero(golden_meshes):
    for deg in (1, 2, 3):
        m = mesh0(golden_meshes)
        outer, inner, vals, shape, _ = m.assemble_rd(deg, lfo.coeff.const(1.0), lfo.coeff.const(0.0))
        A = sp.csc_matrix((vals, inner, outer), shape=shape)
        assert np.abs(A.sum(axis=0)).max() < 1e-10 and np.abs(A.sum(axis=1)).max() < 1e-10


def test_make_sparse_semantics_vs_scipy():
    # independent second opinion on Eigen's setFromTriplets semantics: sorted inner indices, duplicates summed,
    # explicit zeros kept (the P1 Laplacian on right triangles has cancelling diagonal-edge entries)
    m = lfo.Mesh.tp_tria(5, 4)
    outer, inner, vals, shape, _ = m.assemble_rd(1, lfo.coeff.const(1.0), lfo.coeff.const(0.0))
    d, nl = m.cell_dofs(1)
    rows = np.repeat(d[:, :3], 3, axis=1).ravel()
    cols = np.tile(d[:, :3], (1, 3)).ravel()
    S = sp.coo_matrix((np.ones(rows.size), (rows, cols)), shape=shape).tocsc()
    S.sum_duplicates()
    S.sort_indices()
    assert np.array_equal(S.indptr, outer) and np.array_equal(S.indices, inner)
    assert (vals == 0.0).sum() == 2 * 5 * 4  # both ends of every diagonal edge
    assert outer[-1] == inner.size


def test_csr_is_transpose_of_csc(golden_meshes):
    m = mesh0(golden_meshes)
    c = lfo.coeff
    o1, i1, v1, shape, _ = m.assemble_rd(2, c.builtin(100), c.builtin(4))
    o2, i2, v2, _, _ = m.assemble_rd(2, c.builtin(100), c.builtin(4), csr=True)
    A = sp.csc_matrix((v1, i1, o1), shape=shape)
    B = sp.csr_matrix((v2, i2, o2), shape=shape)
    assert abs(A - B).max() == 0.0
    assert np.array_equal(o1, o2) and np.array_equal(i1, i2)  # symmetric pattern


def test_accumulate_semantics(golden_meshes):
    # assembler.h:84-88: the matrix is not zeroed -- assembling twice into one COO doubles the values
    m = mesh0(golden_meshes)
    c = lfo.coeff
    a = m.assemble_rd(1, c.const(1.0), c.const(2.0))
    b = m.assemble_rd(1, c.const(1.0), c.const(2.0), repeat=2)
    assert np.array_equal(a[1], b[1])
    assert np.allclose(2 * a[2], b[2], rtol=1e-15, atol=0)


def test_p3_edge_dof_reversal_consistency(golden_meshes):
    # dofhandler.cc:245-260: two dofs on an edge appear in reversed order in the cell with negative orientation
    m = mesh0(golden_meshes)
    ex = m.export()
    d, nl = m.cell_dofs(3)
    nn, ne = m.n_nodes, m.n_edges
    for c in range(m.n_cells):
        nv = 3 if ex["cell_type"][c] == 3 else 4
        for e in range(nv):
            a, b = d[c, nv + 2 * e], d[c, nv + 2 * e + 1]
            eidx = ex["cell_edges"][c, e]
            lo = nn + 2 * eidx
            if ex["cell_edge_ori"][c, e] > 0:
                assert (a, b) == (lo, lo + 1)
            else:
                assert (a, b) == (lo + 1, lo)


@pytest.mark.parametrize("sel", ["1", "2", "3", "4", "5", "6", "7", "8"])
def test_all_reference_test_meshes_build(golden_meshes, sel):
    m = lfo.Mesh.from_golden(golden_meshes[sel])
    ex = m.export()
    # Euler characteristic of a disc-like mesh: V - E + F = 1
    assert m.n_nodes - m.n_edges + m.n_cells == 1
    # every edge of a cell joins local vertices (j, j+1)
    for c in range(m.n_cells):
        nv = 3 if ex["cell_type"][c] == 3 else 4
        for j in range(nv):
            en = set(ex["edge_nodes"][ex["cell_edges"][c, j]].tolist())
            assert en == {int(ex["cell_nodes"][c, j]), int(ex["cell_nodes"][c, (j + 1) % nv])}
    if sel == "4":
        assert (m.n_cells, m.n_edges, m.n_nodes) == (18, 33, 16)


# ---- lagr_fe_tests.cc:497-532 (lf_fe_ellbvp): alpha = 1, gamma = 0 against the closed-form LinearFELaplaceElementMatrix -------
def test_p1_laplace_element_matrices_closed_form(golden_meshes):
    """uscalfe/lin_fe.cc:39-160: triangles -> |K| grad(lambda_i).grad(lambda_j) with constant barycentric gradients;
    quadrilaterals -> 2x2 Gauss rule on the bilinear parametrisation.  The reference asks 1e-2 in the Frobenius norm; the
    two computations are the same integrals (the default rule of degree 2 is exact / is that Gauss rule), so 1e-12 here."""
    m = mesh0(golden_meshes)
    ex = m.export()
    mats = m.element_matrices(1, lfo.coeff.const(1.0), lfo.coeff.const(0.0))
    g = 0.5 / np.sqrt(3.0)
    gauss = [0.5 - g, 0.5 + g]
    for c in range(m.n_cells):
        nv = 3 if ex["cell_type"][c] == 3 else 4
        p = ex["cell_coords"][c, :nv]
        if nv == 3:
            area2 = (p[1, 0] - p[0, 0]) * (p[2, 1] - p[0, 1]) - (p[2, 0] - p[0, 0]) * (p[1, 1] - p[0, 1])
            grads = np.array([[p[(i + 1) % 3, 1] - p[(i + 2) % 3, 1], p[(i + 2) % 3, 0] - p[(i + 1) % 3, 0]] for i in range(3)]) / area2
            ref = 0.5 * abs(area2) * grads @ grads.T
        else:
            ref = np.zeros((4, 4))
            for x0 in gauss:
                for x1 in gauss:
                    dphi = np.array([[-(1 - x1), -(1 - x0)], [1 - x1, -x0], [x1, x0], [-x1, 1 - x0]])  # bilinear shape functions
                    J = p.T @ dphi
                    G = dphi @ np.linalg.inv(J)
                    ref += 0.25 * abs(np.linalg.det(J)) * G @ G.T
        A = mats[c][:nv, :nv]
        assert np.linalg.norm(A - ref) <= 1e-12 * max(1.0, np.linalg.norm(ref))
