"""Seeded random MSH files in all four encodings (2.2 / 4.1, text / binary, binary also big-endian for 2.2): the product
reader (C ABI, host code) and the oracle's GmshReader restatement must agree bit for bit, and both must reproduce what the
generator put in (file-order node numbering restricted to element vertices, explicit edges first, cells in file order,
consecutive repetitions merged).  Complements the reference's own 19 input files, of which only three are binary."""
import struct

import numpy as np
import pytest

import lehrfempp_b200 as lf
from oracle.lfo_gmsh import GmshReader as OracleReader

NIL = 0xFFFFFFFF


def random_mesh(seed):
    """n x m grid of unit squares, each a quad or two triangles; node tags are a random injective relabelling with gaps;
    some auxiliary (unused) nodes; a few explicit boundary edges and point elements with physical numbers; some elements
    listed twice in a row with another physical number."""
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(1, 5)), int(rng.integers(1, 5))
    idx = lambda i, j: i + j * (n + 1)  # noqa: E731
    n_used = (n + 1) * (m + 1)
    n_aux = int(rng.integers(0, 4))
    tags = rng.permutation(np.arange(1, 3 * (n_used + n_aux) + 1))[: n_used + n_aux]
    order = rng.permutation(n_used + n_aux)  # file order of the nodes
    xy = np.zeros((n_used + n_aux, 2))
    for j in range(m + 1):
        for i in range(n + 1):
            xy[idx(i, j)] = (i + 0.1 * rng.random(), j + 0.1 * rng.random())
    xy[n_used:] = 50.0 + rng.random((n_aux, 2))
    elements = []  # (type, [node ids (mesh-level, before tagging)], physical)
    for k in rng.choice(n_used, size=min(3, n_used), replace=False):
        elements.append((15, [int(k)], int(rng.integers(1, 9))))
        if rng.random() < 0.4:
            elements.append((15, [int(k)], int(rng.integers(1, 9))))
    for i in range(n):  # bottom boundary as explicit edges
        if rng.random() < 0.7:
            e = [idx(i, 0), idx(i + 1, 0)] if rng.random() < 0.5 else [idx(i + 1, 0), idx(i, 0)]
            elements.append((1, e, int(rng.integers(1, 9))))
            if rng.random() < 0.3:
                elements.append((1, e, int(rng.integers(1, 9))))
    for j in range(m):
        for i in range(n):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            if rng.random() < 0.5:
                elements.append((3, [a, b, c, d], int(rng.integers(1, 9))))
            else:
                elements.append((2, [a, b, c], int(rng.integers(1, 9))))
                if rng.random() < 0.3:
                    elements.append((2, [a, b, c], int(rng.integers(1, 9))))
                elements.append((2, [a, c, d], int(rng.integers(1, 9))))
    return dict(tags=tags, order=order, xy=xy, elements=elements, n_used=n_used)


def expected(mesh):
    """what InitGmshFile hands to the MeshFactory (gmsh_reader.cc:121-340)"""
    used = set()
    for t, nodes, _ in mesh["elements"]:
        if t != 15:
            used.update(nodes)
    node_index, xy = {}, []
    for k in mesh["order"]:
        if int(k) in used:
            node_index[int(k)] = len(xy)
            xy.append(mesh["xy"][k])
    edges, cells, phys = [], [], {0: [], 1: [], 2: {}}
    prev = None
    max_pt = -1
    for t, nodes, p in mesh["elements"]:
        key = (t, tuple(nodes))
        if key == prev:
            if t == 15:
                if max_pt >= 0:
                    phys[2][max_pt].append(p)
            else:
                phys[0 if t != 1 else 1][-1].append(p)
            continue
        prev = key
        if t == 15:
            if nodes[0] in node_index:
                phys[2].setdefault(node_index[nodes[0]], []).append(p)
                max_pt = max(phys[2])
        elif t == 1:
            edges.append([node_index[v] for v in nodes])
            phys[1].append([p])
        else:
            cells.append([node_index[v] for v in nodes] + [NIL] * (4 - len(nodes)))
            phys[0].append([p])
    return (np.array(xy).reshape(-1, 2), np.array(edges, dtype=np.uint32).reshape(-1, 2), np.array(cells, dtype=np.uint32).reshape(-1, 4), phys)


def write_v2(mesh, binary, big_endian=False):
    en = ">" if big_endian else "<"
    tags, order, xy = mesh["tags"], mesh["order"], mesh["xy"]
    out = [b"$MeshFormat\n2.2 %d 8\n" % (1 if binary else 0)]
    if binary:
        out.append(struct.pack(en + "i", 1) + b"\n")
    out.append(b"$EndMeshFormat\n$PhysicalNames\n2\n2 1 \"one\"\n1 2 \"two\"\n$EndPhysicalNames\n$Nodes\n%d\n" % len(order))
    for k in order:
        if binary:
            out.append(struct.pack(en + "iddd", int(tags[k]), xy[k, 0], xy[k, 1], 0.0))
        else:
            out.append(b"%d %r %r 0\n" % (int(tags[k]), float(xy[k, 0]), float(xy[k, 1])))
    out.append(b"\n$EndNodes\n" if binary else b"$EndNodes\n")
    out.append(b"$Elements\n%d\n" % len(mesh["elements"]))
    for num, (t, nodes, p) in enumerate(mesh["elements"], 1):
        ids = [int(tags[v]) for v in nodes]
        if binary:
            out.append(struct.pack(en + "iii", t, 1, 2) + struct.pack(en + "i" * (3 + len(ids)), num, p, 7, *ids))
        else:
            out.append((" ".join(map(str, [num, t, 2, p, 7] + ids)) + "\n").encode())
    out.append(b"\n$EndElements\n" if binary else b"$EndElements\n")
    return b"".join(out)


def write_v4(mesh, binary):
    tags, order, xy = mesh["tags"], mesh["order"], mesh["xy"]
    dim_of = {15: 0, 1: 1, 2: 2, 3: 2}
    # one gmsh entity per (dimension, physical number); entity tag = physical number
    ents = {0: set(), 1: set(), 2: set()}
    for t, _, p in mesh["elements"]:
        ents[dim_of[t]].add(p)
    out = [b"$MeshFormat\n4.1 %d 8\n" % (1 if binary else 0)]
    if binary:
        out.append(struct.pack("<i", 1) + b"\n")
    out.append(b"$EndMeshFormat\n$Entities\n")
    if binary:
        buf = struct.pack("<QQQQ", len(ents[0]), len(ents[1]), len(ents[2]), 0)
        for p in sorted(ents[0]):
            buf += struct.pack("<idddQi", p, 0.0, 0.0, 0.0, 1, p)
        for d in (1, 2):
            for p in sorted(ents[d]):
                buf += struct.pack("<iddddddQiQ", p, 0.0, 0.0, 0.0, 1.0, 1.0, 0.0, 1, p, 0)
        out.append(buf + b"\n")
    else:
        out.append(b"%d %d %d 0\n" % (len(ents[0]), len(ents[1]), len(ents[2])))
        for p in sorted(ents[0]):
            out.append(b"%d 0 0 0 1 %d\n" % (p, p))
        for d in (1, 2):
            for p in sorted(ents[d]):
                out.append(b"%d 0 0 0 1 1 0 1 %d 0\n" % (p, p))
    out.append(b"$EndEntities\n$Nodes\n")
    # nodes in file order, split into blocks of up to 3
    blocks = [order[i:i + 3] for i in range(0, len(order), 3)]
    tmin, tmax = int(tags.min()), int(tags.max())
    if binary:
        buf = struct.pack("<QQQQ", len(blocks), len(order), tmin, tmax)
        for b in blocks:
            buf += struct.pack("<iiiQ", 2, 1, 0, len(b)) + struct.pack("<" + "Q" * len(b), *[int(tags[k]) for k in b])
            for k in b:
                buf += struct.pack("<ddd", xy[k, 0], xy[k, 1], 0.0)
        out.append(buf + b"\n")
    else:
        out.append(b"%d %d %d %d\n" % (len(blocks), len(order), tmin, tmax))
        for b in blocks:
            out.append(b"2 1 0 %d\n" % len(b))
            for k in b:
                out.append(b"%d\n" % int(tags[k]))
            for k in b:
                out.append(b"%r %r 0\n" % (float(xy[k, 0]), float(xy[k, 1])))
    out.append(b"$EndNodes\n$Elements\n")
    # one element block per maximal run of equal (type, physical number) in file order
    eblocks = []
    for t, nodes, p in mesh["elements"]:
        if eblocks and eblocks[-1][0] == (t, p):
            eblocks[-1][1].append(nodes)
        else:
            eblocks.append(((t, p), [nodes]))
    n_el = len(mesh["elements"])
    if binary:
        buf = struct.pack("<QQQQ", len(eblocks), n_el, 1, n_el)
        num = 1
        for (t, p), els in eblocks:
            buf += struct.pack("<iiiQ", dim_of[t], p, t, len(els))
            for nodes in els:
                buf += struct.pack("<" + "Q" * (1 + len(nodes)), num, *[int(tags[v]) for v in nodes])
                num += 1
        out.append(buf + b"\n")
    else:
        out.append(b"%d %d 1 %d\n" % (len(eblocks), n_el, n_el))
        num = 1
        for (t, p), els in eblocks:
            out.append(b"%d %d %d %d\n" % (dim_of[t], p, t, len(els)))
            for nodes in els:
                out.append((" ".join(map(str, [num] + [int(tags[v]) for v in nodes])) + " \n").encode())
                num += 1
    out.append(b"$EndElements\n")
    return b"".join(out)


def check_pair(data):
    o, g = OracleReader(data), lf.GmshReader(data)
    oxy, oen, ocn, _ = o.arrays()
    gxy, gen, gcn = g.arrays()
    assert np.array_equal(gxy, oxy) and np.array_equal(gen, oen) and np.array_equal(gcn, ocn)
    for cd, n in ((2, g.n_nodes), (1, g.n_explicit_edges), (0, g.n_cells)):
        for i in range(n + 1):
            assert g.physical_entity_nr(cd, i) == o.physical_entity_nr(cd, i)
    return g, (gxy, gen, gcn)


@pytest.mark.parametrize("seed", range(40))
def test_random_files_v2(seed):
    mesh = random_mesh(seed)
    exy, een, ecn, ephys = expected(mesh)
    for data in (write_v2(mesh, False), write_v2(mesh, True), write_v2(mesh, True, big_endian=True)):
        g, (gxy, gen, gcn) = check_pair(data)
        assert np.array_equal(gxy, exy) and np.array_equal(gen, een) and np.array_equal(gcn, ecn)
        for i, p in enumerate(ephys[0]):
            assert g.physical_entity_nr(0, i) == p
        for i, p in enumerate(ephys[1]):
            assert g.physical_entity_nr(1, i) == p
        for i in range(g.n_nodes):
            assert g.physical_entity_nr(2, i) == ephys[2].get(i, [])


@pytest.mark.parametrize("seed", range(40))
def test_random_files_v4(seed):
    mesh = random_mesh(1000 + seed)
    text = check_pair(write_v4(mesh, False))
    binary = check_pair(write_v4(mesh, True))
    for a, b in zip(text[1], binary[1]):
        assert np.array_equal(a, b)
    # 4.1: mesh nodes are the vertices of the CELLS only (gmsh_reader.cc:384-397), in file order
    used = set()
    for t, nodes, _ in mesh["elements"]:
        if t in (2, 3):
            used.update(nodes)
    want = np.array([mesh["xy"][k] for k in mesh["order"] if int(k) in used]).reshape(-1, 2)
    assert np.array_equal(text[1][0], want)
    # cells: file order, a triangle listed twice in a row with ANOTHER physical number sits in another entity block in 4.1,
    # so it is a new entity there (repetitions are merged within one block only, gmsh_reader.cc:467-474)
    n_listed = sum(1 for t, _, _ in mesh["elements"] if t in (2, 3))
    n_merged = sum(1 for i, e in enumerate(mesh["elements"]) if e[0] in (2, 3) and i > 0 and mesh["elements"][i - 1] == e)
    assert text[0].n_cells == n_listed - n_merged
