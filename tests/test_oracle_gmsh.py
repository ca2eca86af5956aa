"""Pins the oracle's GmshReader restatement (oracle/lfo_gmsh.py) against the reference's reader tests and input files.

  lib/lf/io/test/gmsh_reader_tests.cc:23-140   checkTwoElementMesh      (6 files: 2.2 / 4.1, text / binary, 2nd order)
  lib/lf/io/test/gmsh_reader_tests.cc:184-188  readLectureDemoMesh      (trailing blank at the end of a line)
  lib/lf/io/test/gmsh_reader_tests.cc:193-230  secondOrderMesh          (first-order circle: |area - pi| > 0.3)
  lib/lf/io/test/gmsh_reader_tests.cc:232-300  checkPieceOfCake         (partitioned 4.1 file with periodic links)
  lib/lf/io/test/gmsh_reader_tests.cc:302-311  curvedSquareTests        (files can be read)
The files are the reference's own test inputs (tests/golden/msh, copied by oracle/tools/extract_reference_data.py).
"""
import os

import numpy as np
import pytest

from oracle import lfo
from oracle.lfo_gmsh import GmshError, GmshReader

MSH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "msh")
NIL = 0xFFFFFFFF


def reader(name):
    return GmshReader(os.path.join(MSH, name))


def build_mesh(r):
    xy, en, cn, _ = r.arrays()
    return lfo.Mesh.from_arrays(xy, cn, edge_nodes=en), xy, en, cn


TWO_ELEMENT = ["two_element_hybrid_2d.msh", "two_element_hybrid_2d_binary.msh", "two_element_hybrid_2d_v4.msh",
               "two_element_hybrid_2d_v4_binary.msh", "two_element_hybrid_2d_second_order.msh",
               "two_element_hybrid_2d_second_order_v4.msh"]


@pytest.mark.parametrize("name", TWO_ELEMENT)
def test_two_element_mesh(name):
    r = reader(name)
    m, xy, en, cn = build_mesh(r)
    assert (m.n_cells, m.n_edges, m.n_nodes) == (2, 6, 5)
    # codim 2: the origin carries physical numbers [1, 2], no other node has any
    origin = int(np.argmin((xy ** 2).sum(axis=1)))
    assert (xy[origin] ** 2).sum() < 1e-10
    assert r.physical_entity_nr(2, origin) == [1, 2]
    assert all(r.physical_entity_nr(2, i) == [] for i in range(5) if i != origin)
    assert r.nr2name(1, 2) == "physicalEntity1" and r.nr2name(2, 2) == "physicalEntity2" and r.nr2name(2) == "physicalEntity2"
    with pytest.raises(GmshError):
        r.nr2name(1)
    assert r.name2nr("physicalEntity1", 2) == 1 and r.name2nr("physicalEntity2", 2) == 2 and r.name2nr("physicalEntity2") == 2
    for bad in (lambda: r.name2nr("physicalEntity1"), lambda: r.nr2name(100), lambda: r.name2nr("gugus")):
        with pytest.raises(GmshError):
            bad()
    assert r.physical_entities(2) == [(1, "physicalEntity1"), (2, "physicalEntity2")]
    # codim 1: exactly one explicit edge, the diagonal (length sqrt 2), with physical number "diagonal" = 4
    ex = m.export()
    length = np.linalg.norm(xy[ex["edge_nodes"][:, 1]] - xy[ex["edge_nodes"][:, 0]], axis=1)
    diag = np.flatnonzero(length > 1.1)
    assert list(diag) == [0] and len(en) == 1  # an explicitly added edge keeps its insertion index
    nr = r.name2nr("diagonal")
    assert nr == 4 and r.nr2name(nr) == "diagonal" and r.physical_entity_nr(1, 0) == [4]
    assert all(r.physical_entity_nr(1, e) == [] for e in range(1, 6))
    assert r.physical_entities(1) == [(4, "diagonal")]
    # codim 0
    types = ex["cell_type"]
    square, tria = int(np.flatnonzero(types == 4)[0]), int(np.flatnonzero(types == 3)[0])
    assert r.name2nr("square") == 5 and r.physical_entity_nr(0, square) == [5]
    assert r.name2nr("physicalEntity1", 0) == 1 and r.name2nr("physicalEntity3") == 3
    assert r.nr2name(1, 0) == "physicalEntity1" and r.nr2name(3) == "physicalEntity3" and r.nr2name(3, 0) == "physicalEntity3"
    with pytest.raises(GmshError):
        r.nr2name(3, 1)
    assert r.physical_entity_nr(0, tria) == [1, 3]
    assert r.physical_entities(0) == [(1, "physicalEntity1"), (3, "physicalEntity3"), (5, "square")]
    assert r.arrays()[3] == (2 if "second_order" in name else 1)


def test_numbering_follows_the_file():
    """2.2 lists the triangle before the square, 4.1 the surface of the square first: cell indices differ accordingly."""
    c2 = reader("two_element_hybrid_2d.msh").arrays()[2]
    c4 = reader("two_element_hybrid_2d_v4.msh").arrays()[2]
    assert c2[0, 3] == NIL and c2[1, 3] != NIL
    assert c4[0, 3] != NIL and c4[1, 3] == NIL
    # text and binary variants of one file are the same mesh
    for a, b in (("two_element_hybrid_2d.msh", "two_element_hybrid_2d_binary.msh"),
                 ("two_element_hybrid_2d_v4.msh", "two_element_hybrid_2d_v4_binary.msh")):
        for u, v in zip(reader(a).arrays(), reader(b).arrays()):
            assert np.array_equal(u, v)


def test_lecture_demo_mesh_with_trailing_blank():
    r = reader("lecturedemomesh.msh")
    m, xy, en, cn = build_mesh(r)
    assert (m.n_nodes, m.n_cells) == (8, 5) and len(en) == 7  # 8 boundary segments, one listed twice in a row (merged)
    assert r.physical_entity_nr(1, 2) == [2, 4]


def area(xy, cn):
    tot = 0.0
    for c in cn:
        v = xy[c[c != NIL]]
        x, y = v[:, 0], v[:, 1]
        tot += 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
    return tot


@pytest.mark.parametrize("name", ["circle_first_order.msh", "circle_first_order_v4.msh"])
def test_first_order_circle_area(name):
    r = reader(name)
    xy, en, cn, order = r.arrays()
    assert order == 1
    assert abs(area(xy, cn) - np.pi) > 0.3
    build_mesh(r)


@pytest.mark.parametrize("name", ["circle_second_order.msh", "circle_second_order_v4.msh", "circle_second_order_quad.msh",
                                  "circle_second_order_quad_v4.msh", "curved_square_quads_2nd_order.msh",
                                  "curved_square_trias_2nd_order.msh", "curved_square_quads_2nd_order_v4.msh",
                                  "curved_square_trias_2nd_order_v4.msh"])
def test_second_order_files_can_be_read(name):
    r = reader(name)
    xy, en, cn, order = r.arrays()
    assert order == 2
    # auxiliary (mid-side) nodes are not mesh nodes: every mesh node is the vertex of a cell
    used = np.unique(cn[cn != NIL])
    assert np.array_equal(used, np.arange(len(xy)))
    build_mesh(r)


@pytest.mark.parametrize("name", ["piece_of_cake.msh", "piece_of_cake_binary.msh"])
def test_piece_of_cake(name):
    r = reader(name)
    m, xy, en, cn = build_mesh(r)
    assert (m.n_cells, m.n_edges, m.n_nodes) == (2, 5, 4)
    origin = int(np.argmin((xy ** 2).sum(axis=1)))
    assert r.physical_entity_nr(2, origin) == [1] and r.is_physical_entity(2, origin, 1)
    ex = m.export()
    for e in range(m.n_edges):
        p = xy[ex["edge_nodes"][e]]
        on_arc = abs(np.linalg.norm(p[0]) - 1) < 1e-6 and abs(np.linalg.norm(p[1]) - 1) < 1e-6
        if on_arc:
            assert r.physical_entity_nr(1, e) == [2]
        assert r.is_physical_entity(1, e, 2) == on_arc
    # the partition interface (partitioned curve 8, child of surface 1) inherits the surface's physical number
    assert sorted(map(tuple, (r.physical_entity_nr(1, e) for e in range(m.n_edges)))) == [(), (), (2,), (2,), (3,)]
    for c in range(2):
        assert r.physical_entity_nr(0, c) == [3] and r.is_physical_entity(0, c, 3)
    assert r.name2nr("origin") == 1 and r.name2nr("arc") == 2
    assert r.nr2name(1) == "origin" and r.nr2name(1, 2) == "origin" and r.nr2name(2) == "arc" and r.nr2name(2, 1) == "arc"
    for bad in (lambda: r.nr2name(3), lambda: r.nr2name(3, 1)):
        with pytest.raises(GmshError):
            bad()
    assert r.physical_entities(0) == [] and r.physical_entities(1) == [(2, "arc")] and r.physical_entities(2) == [(1, "origin")]
