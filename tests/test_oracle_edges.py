"""CPU checks of the oracle's edge (codim-1) path against the reference's own known answers:

  lib/lf/assemble/test/assembly_tests.cc:491-590   10x10 matrix of the boundary assembly (AssembleMatrixLocally, codim 1)
  lib/lf/uscalfe/test/lagr_fe_tests.cc:571-601     P1 edge mass matrix = |e| [[1/3, 1/6], [1/6, 1/3]]
  lib/lf/uscalfe/test/lagr_fe_tests.cc:696-722     P1 edge load vector = |e| [1/2, 1/2]
  lib/lf/uscalfe/test/lagr_fe_tests.cc:916-963     1^T M_e 1 = |e|, 1^T v_e = |e| for every degree
"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import lfo


def mesh0(golden_meshes):
    return lfo.Mesh.from_golden(golden_meshes["0"])


def edge_lengths(m):
    ex = m.export()
    p = ex["node_coords"][ex["edge_nodes"]]
    return np.linalg.norm(p[:, 1] - p[:, 0], axis=1)


def test_golden_boundary_assembly(golden_meshes, assembly_goldens):
    m = mesh0(golden_meshes)
    dofh = lfo.DofHandler(m, n_pt=1)
    assert dofh.num_dofs == 10
    _, idx = dofh.dof_entities()
    g = next(e for e in assembly_goldens["ref_mat_10"] if e["line"] > 500)
    ref = np.array(g["row_major"]).reshape(10, 10)
    A = dofh.boundary_test_matrix()
    assert np.array_equal(A, ref[np.ix_(idx, idx)])
    # the boundary of the test mesh: edges 11..17 with exactly one adjacent cell
    assert np.flatnonzero(m.boundary_edges()).tolist() == [11, 12, 13, 14, 15, 16, 17]


def test_p1_edge_mass_and_load_known_answers(golden_meshes):
    m = mesh0(golden_meshes)
    L = edge_lengths(m)
    M = m.edge_matrices(1, lfo.coeff.const(1.0))
    ref = np.array([[1 / 3, 1 / 6], [1 / 6, 1 / 3]])
    assert np.abs(M - L[:, None, None] * ref).max() <= 1e-14  # the reference asks 1e-6
    v = m.edge_vectors(1, lfo.coeff.const(1.0))
    assert np.abs(v - L[:, None] * 0.5).max() <= 1e-14


@pytest.mark.parametrize("degree", [1, 2, 3])
def test_edge_sums_are_lengths(golden_meshes, degree):
    m = mesh0(golden_meshes)
    L = edge_lengths(m)
    M = m.edge_matrices(degree, lfo.coeff.const(1.0))
    v = m.edge_vectors(degree, lfo.coeff.const(1.0))
    assert np.abs(M.sum(axis=(1, 2)) - L).max() <= 1e-14  # reference: 1e-3
    assert np.abs(v.sum(axis=1) - L).max() <= 1e-14
    assert np.allclose(M, M.transpose(0, 2, 1), rtol=0, atol=1e-16)
    # the default rule (degree 2p, p + 1 Gauss points) integrates the mass matrix exactly: compare with a richer rule
    assert np.abs(M - m.edge_matrices(degree, lfo.coeff.const(1.0), qr_degree=12)).max() <= 1e-15


@pytest.mark.parametrize("degree", [1, 2, 3])
def test_rd_plus_edge_matrix_is_sum_of_parts(degree):
    m = lfo.Mesh.hybrid(6, 0.2, 4)
    c = lfo.coeff
    bd = m.boundary_edges()
    o, i, v = m.assemble_rd_edge(degree, c.const(1.0), c.const(0.5), c.builtin(1), edge_mask=bd)
    o0, i0, v0, shape, _ = m.assemble_rd(degree, c.const(1.0), c.const(0.5))
    assert np.array_equal(o, o0) and np.array_equal(i, i0)  # edge entries live inside the cell pattern
    D = sp.csc_matrix((v, i, o), shape=shape) - sp.csc_matrix((v0, i0, o0), shape=shape)
    # the difference is the boundary mass matrix: symmetric, rows of interior dofs empty, total = int_boundary eta ds
    assert abs(D - D.T).max() <= 1e-15
    ones = np.ones(shape[0])
    # eta = 1 + |x|^2 on the boundary of [0,1]^2: 4 + (1/3 + 1/3 + 4/3 + 4/3) = 22/3
    assert abs(ones @ (D @ ones) - 22.0 / 3.0) <= 1e-12
    # load vector with g = 1: sum = perimeter
    b = m.assemble_edge_load(degree, c.const(1.0), edge_mask=bd)
    assert abs(b.sum() - 4.0) <= 1e-13
