"""Pins the oracle's DynamicFEDofHandler restatement (oracle/lfo_assemble.h) against the reference's own tests.

  lib/lf/assemble/test/assembly_tests.cc:180-217   dynamic_dof_index_test: layout {1,2,3,4} "should produce the same output"
                                                   as the UniformFEDofHandler of dof_index_test
  lib/lf/assemble/test/assembly_tests.cc:305-322   dynamic_dof_test   : the 10x10 golden through a dynamic handler
  lib/lf/assemble/test/assembly_tests.cc:476-490   edge_dof_dynamic   : the 36x36 golden through a dynamic handler
"""
import numpy as np
import pytest

from oracle import lfo


def mesh0(golden_meshes):
    return lfo.Mesh.from_golden(golden_meshes["0"])


def test_dynamic_equals_uniform_layout_1234(golden_meshes):
    m = mesh0(golden_meshes)
    ex = m.export()
    uni = lfo.DofHandler(m, n_pt=1, n_seg=2, n_tria=3, n_quad=4)
    dyn = lfo.DofHandler.dynamic(m, np.full(m.n_nodes, 1), np.full(m.n_edges, 2), np.where(ex["cell_type"] == 3, 3, 4))
    assert dyn.num_dofs == uni.num_dofs == 10 + 2 * 18 + sum(3 if t == 3 else 4 for t in ex["cell_type"])
    ud, unl = uni.cell_dofs()
    dd, dnl = dyn.cell_dofs()
    assert np.array_equal(unl, dnl) and np.array_equal(ud, dd)
    # output_entities_dofs (assembly_tests.cc:131-153): every dof belongs to the entity it is reported for
    assert all(np.array_equal(a, b) for a, b in zip(uni.dof_entities(), dyn.dof_entities()))


def test_golden_10x10_through_dynamic_handler(golden_meshes, assembly_goldens):
    m = mesh0(golden_meshes)
    dyn = lfo.DofHandler.dynamic(m, n_int_node=np.ones(m.n_nodes))
    assert dyn.num_dofs == 10
    g = next(e for e in assembly_goldens["ref_mat_10"] if e["line"] < 300)
    ref = np.array(g["row_major"]).reshape(10, 10)
    _, idx = dyn.dof_entities()
    assert np.array_equal(dyn.test_matrix(0), ref[np.ix_(idx, idx)])
    assert np.array_equal(dyn.test_vector(), np.diag(ref)[idx])


def test_golden_36x36_through_dynamic_handler(golden_meshes, assembly_goldens):
    m = mesh0(golden_meshes)
    dyn = lfo.DofHandler.dynamic(m, n_int_edge=np.full(m.n_edges, 2))
    assert dyn.num_dofs == 36
    ref = np.array(assembly_goldens["ref_mat_36"][0]["row_major"]).reshape(36, 36)
    assert np.array_equal(dyn.test_matrix(1), ref)


def variable_layout(m, seed):
    rng = np.random.default_rng(seed)
    ex = m.export()
    return rng.integers(0, 3, m.n_nodes), rng.integers(0, 4, m.n_edges), rng.integers(0, 3, m.n_cells), ex


@pytest.mark.parametrize("sel", ["0", "1", "5"])
def test_variable_layout_invariants(golden_meshes, sel):
    """hp-style layout: numbering by codimension blocks, shared edge dofs reversed exactly where the orientation is negative."""
    m = lfo.Mesh.from_golden(golden_meshes[sel])
    nn, ne, nc, ex = variable_layout(m, 7)
    dyn = lfo.DofHandler.dynamic(m, nn, ne, nc)
    assert dyn.num_dofs == nn.sum() + ne.sum() + nc.sum()
    codim, idx = dyn.dof_entities()
    # nodes first (index order), then edges, then cells
    expect = np.concatenate([np.repeat(np.arange(m.n_nodes), nn), np.repeat(np.arange(m.n_edges), ne), np.repeat(np.arange(m.n_cells), nc)])
    assert np.array_equal(idx, expect)
    assert np.array_equal(codim, np.concatenate([np.full(nn.sum(), 2), np.full(ne.sum(), 1), np.full(nc.sum(), 0)]))
    d, nl = dyn.cell_dofs()
    node_off = np.concatenate([[0], np.cumsum(nn)])
    edge_off = nn.sum() + np.concatenate([[0], np.cumsum(ne)])
    cell_off = nn.sum() + ne.sum() + np.concatenate([[0], np.cumsum(nc)])
    for c in range(m.n_cells):
        nv = 3 if ex["cell_type"][c] == 3 else 4
        want = []
        for l in range(nv):
            v = ex["cell_nodes"][c, l]
            want += list(range(node_off[v], node_off[v + 1]))
        for l in range(nv):
            e = ex["cell_edges"][c, l]
            r = list(range(edge_off[e], edge_off[e + 1]))
            want += r if ex["cell_edge_ori"][c, l] > 0 else r[::-1]
        want += list(range(cell_off[c], cell_off[c + 1]))
        assert nl[c] == len(want) and list(d[c, : nl[c]]) == want and np.all(d[c, nl[c]:] == -1)
