"""CPU checks of the oracle's regular refinement (lib/lf/refinement/mesh_hierarchy.cc:368-1262, rp_regular only).

The reference's refinement tests (lib/lf/refinement/test/*.cc) check relations, not literal index tables, so the numbering
is checked here against the rules read off the reference code, plus geometric / topological invariants."""
import numpy as np
import pytest

from oracle import lfo


def areas(ex):
    out = []
    for c, t in enumerate(ex["cell_type"]):
        p = ex["cell_coords"][c]
        n = 3 if t == 3 else 4
        x, y = p[:n, 0], p[:n, 1]
        out.append(0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))))
    return np.array(out)


@pytest.mark.parametrize("kind", ["tria", "quad", "hybrid", "golden"])
def test_regular_refinement_numbering_rules(kind, golden_meshes):
    if kind == "tria":
        m = lfo.Mesh.tp_tria(4, 3, 0.0, 0.0, 2.0, 1.0)
    elif kind == "quad":
        m = lfo.Mesh.tp_quad(3, 4)
    elif kind == "hybrid":
        m = lfo.Mesh.hybrid(5, 0.2, 3)
    else:
        m = lfo.Mesh.from_golden(golden_meshes["0"])
    f = m.refine_regular()
    ex, fx = m.export(), f.export()
    nn, ne, nc = m.n_nodes, m.n_edges, m.n_cells
    # counts (Euler): nodes + edge midpoints + quad centres; 2 children per edge + 3 / 4 interior edges; 4 children per cell
    assert f.n_nodes == nn + ne + m.n_quad
    assert f.n_edges == 2 * ne + 3 * m.n_tria + 4 * m.n_quad
    assert (f.n_cells, f.n_tria, f.n_quad) == (4 * nc, 4 * m.n_tria, 4 * m.n_quad)
    assert f.n_nodes - f.n_edges + f.n_cells == m.n_nodes - m.n_edges + m.n_cells
    # nodes: copies keep their index, midpoint of parent edge e is node nn + e
    assert np.array_equal(fx["node_coords"][:nn], ex["node_coords"])
    p = ex["node_coords"][ex["edge_nodes"]]
    assert np.allclose(fx["node_coords"][nn:nn + ne], 0.5 * (p[:, 0] + p[:, 1]), rtol=0, atol=1e-15)
    # edges: children of parent edge e are 2e = (p0, mid), 2e + 1 = (mid, p1), direction kept
    mids = nn + np.arange(ne)
    assert np.array_equal(fx["edge_nodes"][0:2 * ne:2], np.stack([ex["edge_nodes"][:, 0], mids], 1))
    assert np.array_equal(fx["edge_nodes"][1:2 * ne:2], np.stack([mids, ex["edge_nodes"][:, 1]], 1))
    # cells: children of parent cell c are 4c .. 4c + 3 with the node lists of the reference
    centre = nn + ne
    for c in range(nc):
        v = ex["cell_nodes"][c]
        mm = nn + ex["cell_edges"][c]
        ch = fx["cell_nodes"][4 * c:4 * c + 4]
        if ex["cell_type"][c] == 3:
            exp = [[v[0], mm[0], mm[2]], [v[1], mm[0], mm[1]], [v[2], mm[2], mm[1]], [mm[0], mm[1], mm[2]]]
            assert np.array_equal(ch[:, :3], np.array(exp, dtype=np.uint32)) and np.all(ch[:, 3] == lfo.NIL)
        else:
            exp = [[v[0], mm[0], centre, mm[3]], [v[1], mm[1], centre, mm[0]], [v[2], mm[1], centre, mm[2]], [v[3], mm[2], centre, mm[3]]]
            assert np.array_equal(ch, np.array(exp, dtype=np.uint32))
            centre += 1
    # geometry: children tile the parent, corner coordinates are the node positions here
    assert np.allclose(areas(fx).reshape(nc, 4).sum(axis=1), areas(ex), rtol=1e-13, atol=0)
    fc = fx["cell_coords"]
    for c in range(f.n_cells):
        n = 3 if fx["cell_type"][c] == 3 else 4
        assert np.allclose(fc[c, :n], fx["node_coords"][fx["cell_nodes"][c, :n]], rtol=0, atol=1e-15)
    # every edge of a child cell joins its local vertices (j, j + 1) and is a registered edge
    for c in range(f.n_cells):
        n = 3 if fx["cell_type"][c] == 3 else 4
        for j in range(n):
            en = set(fx["edge_nodes"][fx["cell_edges"][c, j]].tolist())
            assert en == {int(fx["cell_nodes"][c, j]), int(fx["cell_nodes"][c, (j + 1) % n])}


def test_two_levels_and_p3_dofs():
    m = lfo.Mesh.tp_tria(3, 2)
    f2 = m.refine_regular().refine_regular()
    assert f2.n_cells == 16 * m.n_cells
    # the Lagrange dof count of the refined mesh: nodes + 2 per edge + 1 per triangle
    assert f2.num_dofs(3) == f2.n_nodes + 2 * f2.n_edges + f2.n_cells
    outer, inner, vals, shape, _ = f2.assemble_rd(3, lfo.coeff.const(1.0), lfo.coeff.const(1.0))
    import scipy.sparse as sp
    A = sp.csc_matrix((vals, inner, outer), shape=shape)
    ones = np.ones(shape[0])
    assert abs(ones @ (A @ ones) - 1.0) <= 1e-13  # stiffness annihilates constants, mass sums to |Omega| = 1
