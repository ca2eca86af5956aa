"""ORACLE (test infrastructure, NOT product code): lf::io::GmshReader restated in plain Python.

Follows (paths relative to lib/lf/io/ of the reference):
  gmsh_reader.cc:629-697      ReadGmshFile          header "$MeshFormat version is_binary sizeof(size_t) [int 1]", dispatch on 2.2 / 4.1
  gmsh_file_v2.h:27-268, gmsh_file_v2.cc:425-735   MSH 2.2 text + binary (sections PhysicalNames, Nodes, Elements, Periodic)
  gmsh_file_v4.h:29-540, gmsh_file_v4_text.cc, gmsh_file_v4_binary.cc   MSH 4.1 text + binary (PhysicalNames, Entities,
                              PartitionedEntities, Nodes, Elements, Periodic, GhostElements)
  gmsh_reader.cc:121-340      GmshReader::InitGmshFile(GMshFileV2)
  gmsh_reader.cc:343-627      GmshReader::InitGmshFile(GMshFileV4)
  gmsh_reader.cc:30-118       PhysicalEntityName2Nr / PhysicalEntityNr2Name / PhysicalEntities

What InitGmshFile hands to the MeshFactory, and therefore what this module returns:
  * nodes    : the MAIN nodes (vertices of elements; v2: of every non-point element, v4: of elements of dimension dim_mesh)
               in FILE order -> mesh node index = position among the main nodes (AddPoint order)
  * entities : every non-point element in file order, consecutive repetitions of the same element (same type, same node
               list) merged into one entity carrying several physical numbers -> AddEntity order = index of explicitly
               added edges (hybrid2d::MeshFactory keeps them first, in insertion order) and of cells
  * physical numbers per entity and the name <-> number tables

Pinned by the expectations of lib/lf/io/test/gmsh_reader_tests.cc (tests/test_oracle_gmsh.py).  Pure-Python loops: the
fixtures are a few hundred bytes.
"""
import struct

import numpy as np

# gmsh_file_v2.h:33-101 / gmsh_file_v4.h:35-103: element type -> (number of nodes, dimension)
ELEMENT_TYPES = {
    1: (2, 1), 2: (3, 2), 3: (4, 2), 4: (4, 3), 5: (8, 3), 6: (6, 3), 7: (5, 3), 8: (3, 1), 9: (6, 2), 10: (9, 2), 11: (10, 3),
    12: (27, 3), 13: (18, 3), 14: (14, 3), 15: (1, 0), 16: (8, 2), 17: (20, 3), 18: (15, 3), 19: (13, 3), 20: (9, 2), 21: (10, 2),
    22: (12, 2), 23: (15, 2), 24: (15, 2), 25: (21, 2), 26: (4, 1), 27: (5, 1), 28: (6, 1), 29: (20, 3), 30: (35, 3), 31: (56, 3),
    92: (64, 3), 93: (125, 3),
}
# element types GmshReader can turn into entities (gmsh_reader.cc:259-292): type -> (main nodes, geometry order)
SUPPORTED = {1: (2, 1), 8: (2, 2), 2: (3, 1), 9: (3, 2), 3: (4, 1), 16: (4, 2), 10: (4, 2)}


class GmshError(RuntimeError):
    pass


class _Cursor:
    """Byte cursor with the two reading modes of the reference parsers: whitespace-separated text and raw binary."""

    def __init__(self, data):
        self.d = data
        self.p = 0

    def skip_ws(self):
        d, p = self.d, self.p
        while p < len(d) and d[p:p + 1].isspace():
            p += 1
        self.p = p

    def at_end(self):
        self.skip_ws()
        return self.p >= len(self.d)

    def token(self):
        self.skip_ws()
        d, p = self.d, self.p
        q = p
        while q < len(d) and not d[q:q + 1].isspace():
            q += 1
        if q == p:
            raise GmshError("unexpected end of file")
        self.p = q
        return d[p:q]

    def peek(self):
        p = self.p
        try:
            return self.token()
        finally:
            self.p = p

    def integer(self):
        return int(self.token())

    def real(self):
        return float(self.token())

    def expect(self, word):
        t = self.token()
        if t != word:
            raise GmshError("expected %r, found %r" % (word, t))

    def quoted(self):
        self.skip_ws()
        if self.d[self.p:self.p + 1] != b'"':
            raise GmshError("expected a quoted string")
        q = self.d.index(b'"', self.p + 1)
        s = self.d[self.p + 1:q].decode()
        self.p = q + 1
        return s

    def eol(self):
        """consume the single line break that precedes a binary payload"""
        if self.d[self.p:self.p + 2] == b"\r\n":
            self.p += 2
        elif self.d[self.p:self.p + 1] == b"\n":
            self.p += 1
        else:
            raise GmshError("expected end of line before binary data")

    def raw(self, fmt):
        n = struct.calcsize(fmt)
        if self.p + n > len(self.d):
            raise GmshError("binary payload truncated")
        v = struct.unpack_from(fmt, self.d, self.p)
        self.p += n
        return v

    def skip_section(self, name):
        end = b"$End" + name[1:]
        q = self.d.find(end, self.p)
        if q < 0:
            raise GmshError("section %r is not closed" % name)
        self.p = q + len(end)


# ---- file level ------------------------------------------------------------------------------------------------------
def parse(data):
    """ReadGmshFile (gmsh_reader.cc:629-697): returns ("2.2", dict) or ("4.1", dict)."""
    c = _Cursor(data)
    c.expect(b"$MeshFormat")
    version = c.token().decode()
    binary = c.integer()
    size_t_size = c.integer()
    endian = "<"
    if binary == 1:
        c.eol()
        (one,) = c.raw("<i")
        if one != 1:
            endian = ">"
    elif binary != 0:
        raise GmshError("Could not read header")
    c.expect(b"$EndMeshFormat")
    if size_t_size != 8:
        raise GmshError("Size of std::size_t must be 8.")
    if version == "4.1":
        return version, _parse_v4(c, bool(binary), endian)
    if version == "2.2":
        return version, _parse_v2(c, bool(binary), endian)
    raise GmshError("GmshFiles with Version %s are not yet supported." % version)


def _physical_names(c):
    n = c.integer()
    out = []
    for _ in range(n):
        dim, nr = c.integer(), c.integer()
        out.append((dim, nr, c.quoted()))
    c.expect(b"$EndPhysicalNames")
    return out


def _parse_v2(c, binary, en):
    f = dict(physical=[], nodes=[], elements=[])
    while not c.at_end():
        sec = c.token()
        if sec == b"$PhysicalNames":
            f["physical"] = _physical_names(c)
        elif sec == b"$Nodes":
            n = c.integer()
            if binary:
                c.eol()
                for _ in range(n):
                    tag, x, y, z = c.raw(en + "iddd")
                    f["nodes"].append((tag, (x, y, z)))
            else:
                for _ in range(n):
                    tag = c.integer()
                    f["nodes"].append((tag, (c.real(), c.real(), c.real())))
            c.expect(b"$EndNodes")
        elif sec == b"$Elements":
            n = c.integer()
            if binary:
                c.eol()
                done = 0
                while done < n:
                    etype, count, ntags = c.raw(en + "iii")
                    nn = ELEMENT_TYPES[etype][0]
                    for _ in range(count):
                        vals = c.raw(en + "i" * (1 + ntags + nn))
                        f["elements"].append(_v2_element(vals[0], etype, vals[1:1 + ntags], vals[1 + ntags:]))
                    done += count
            else:
                for _ in range(n):
                    number, etype, ntags = c.integer(), c.integer(), c.integer()
                    tags = [c.integer() for _ in range(ntags)]
                    nn = ELEMENT_TYPES[etype][0]
                    f["elements"].append(_v2_element(number, etype, tags, [c.integer() for _ in range(nn)]))
            c.expect(b"$EndElements")
        elif sec.startswith(b"$"):
            c.skip_section(sec)  # $Periodic is read and ignored by GmshReader (gmsh_reader.cc:333-339); comments
        else:
            raise GmshError("Could not parse file")
    return f


def _v2_element(number, etype, tags, nodes):
    # gmsh_file_v2.h:124-158: tags = physical entity, elementary entity [, number of partitions, partitions ...]
    if len(tags) < 2:
        raise GmshError("element %d has fewer than two tags" % number)
    return dict(number=number, type=etype, physical=tags[0], elementary=tags[1], partitions=list(tags[3:]), nodes=list(nodes))


def _parse_v4(c, binary, en):
    f = dict(physical=[], entities=[{}, {}, {}, {}], num_partitions=0, part_entities=[{}, {}, {}, {}], node_blocks=[], min_node_tag=0,
             max_node_tag=0, element_blocks=[])

    def ints(n):
        return list(c.raw(en + "i" * n)) if binary else [c.integer() for _ in range(n)]

    def sizes(n):
        return list(c.raw(en + "Q" * n)) if binary else [c.integer() for _ in range(n)]

    def reals(n):
        return list(c.raw(en + "d" * n)) if binary else [c.real() for _ in range(n)]

    while not c.at_end():
        sec = c.token()
        if sec == b"$PhysicalNames":
            f["physical"] = _physical_names(c)
        elif sec == b"$Entities":
            if binary:
                c.eol()
            counts = sizes(4)
            for dim in range(4):
                for _ in range(counts[dim]):
                    (tag,) = ints(1)
                    reals(3 if dim == 0 else 6)
                    phys = ints(sizes(1)[0])
                    if dim > 0:
                        ints(sizes(1)[0])  # bounding entities
                    f["entities"][dim][tag] = phys
            c.expect(b"$EndEntities")
        elif sec == b"$PartitionedEntities":
            if binary:
                c.eol()
            f["num_partitions"] = sizes(1)[0]
            for _ in range(sizes(1)[0]):
                ints(2)  # ghost entity tag, partition
            counts = sizes(4)
            for dim in range(4):
                for _ in range(counts[dim]):
                    tag, _parent_dim, _parent_tag = ints(3)
                    ints(sizes(1)[0])  # partitions
                    reals(3 if dim == 0 else 6)
                    phys = ints(sizes(1)[0])
                    if dim > 0:
                        ints(sizes(1)[0])
                    f["part_entities"][dim][tag] = phys
            c.expect(b"$EndPartitionedEntities")
        elif sec == b"$Nodes":
            if binary:
                c.eol()
            nblocks, _total, f["min_node_tag"], f["max_node_tag"] = sizes(4)
            for _ in range(nblocks):
                dim, _etag, parametric = ints(3)
                n = sizes(1)[0]
                tags = sizes(n)
                block = []
                for k in range(n):
                    xyz = reals(3)
                    if parametric:
                        reals(dim)
                    block.append((tags[k], tuple(xyz)))
                f["node_blocks"].append(block)
            c.expect(b"$EndNodes")
        elif sec == b"$Elements":
            if binary:
                c.eol()
            nblocks = sizes(4)[0]
            for _ in range(nblocks):
                dim, etag, etype = ints(3)
                n = sizes(1)[0]
                nn = ELEMENT_TYPES[etype][0]
                elems = []
                for _ in range(n):
                    v = sizes(1 + nn)
                    elems.append((v[0], v[1:]))
                f["element_blocks"].append(dict(dim=dim, entity_tag=etag, type=etype, elements=elems))
            c.expect(b"$EndElements")
        elif sec.startswith(b"$"):
            c.skip_section(sec)
        else:
            raise GmshError("Could not parse file")
    return f


# ---- GmshReader ------------------------------------------------------------------------------------------------------
class GmshReader:
    """GmshReader(std::make_unique<hybrid2d::MeshFactory>(dim_world), file): 2D meshes (dim_mesh = 2)."""

    def __init__(self, data, dim_world=2):
        if isinstance(data, str):
            with open(data, "rb") as fh:
                data = fh.read()
        self.version, f = parse(data)
        self.dim_mesh = 2
        self.dim_world = dim_world
        self.node_xy = []          # AddPoint order
        self.entities = {1: [], 0: []}   # codim -> list of (type, main node indices (mesh), all node coordinates) in AddEntity order
        self.physical = {2: {}, 1: [], 0: []}  # codim -> physical numbers per entity (nodes: dict index -> list)
        if self.version == "2.2":
            self._init_v2(f)
        else:
            self._init_v4(f)
        # gmsh_reader.cc:321-331 / 607-615
        self.names = [(nr, name, self.dim_mesh - dim) for dim, nr, name in f["physical"]]

    # -- gmsh_reader.cc:121-340 ------------------------------------------------------------------------------------------
    def _init_v2(self, f):
        main = set()
        n_top = 0
        for e in f["elements"]:
            nn, dim = ELEMENT_TYPES[e["type"]]
            if dim > self.dim_mesh:
                raise GmshError("msh-file contains entities with dimension %d" % dim)
            n_top += dim == self.dim_mesh
            if e["type"] != 15:
                if e["type"] not in SUPPORTED:
                    raise GmshError("Gmsh element type %d not (yet) supported by GmshReader." % e["type"])
                main.update(e["nodes"][:SUPPORTED[e["type"]][0]])
        if n_top == 0:
            raise GmshError("MshFile contains no elements with dimension %d" % self.dim_mesh)
        gi2mi, coords = {}, {}
        for tag, xyz in f["nodes"]:
            coords[tag] = xyz
            if tag in main:
                gi2mi[tag] = self._add_point(xyz)
        prev = None
        for e in f["elements"]:
            key = (e["type"], tuple(e["nodes"]))
            if key == prev:  # "This entity appears more than once" (:224-229)
                self._last_physical.append(e["physical"])
                continue
            prev = key
            self._insert(e["type"], e["nodes"], gi2mi, coords, [e["physical"]])

    # -- gmsh_reader.cc:343-627 ------------------------------------------------------------------------------------------
    def _init_v4(self, f):
        main = set()
        n_top = 0
        for b in f["element_blocks"]:
            nn, dim = ELEMENT_TYPES[b["type"]]
            if b["dim"] > self.dim_mesh:
                raise GmshError("msh-file contains entities with dimension %d" % b["dim"])
            if b["dim"] != dim:
                raise GmshError("error in GmshFile: Mismatch between entity block type and dimension")
            if b["dim"] == self.dim_mesh:
                n_top += len(b["elements"])
                if b["type"] not in SUPPORTED:
                    raise GmshError("Gmsh element type %d not (yet) supported by GmshReader." % b["type"])
                for _, nodes in b["elements"]:
                    main.update(nodes[:SUPPORTED[b["type"]][0]])
        if n_top == 0:
            raise GmshError("MshFile contains no elements with dimension %d" % self.dim_mesh)
        gi2mi, coords = {}, {}
        for block in f["node_blocks"]:
            for tag, xyz in block:
                coords[tag] = xyz
                if tag in main:
                    gi2mi[tag] = self._add_point(xyz)
        ent = f["part_entities"] if f["num_partitions"] != 0 else f["entities"]
        for b in f["element_blocks"]:
            phys = list(ent[b["dim"]].get(b["entity_tag"], []))
            prev = None
            for _, nodes in b["elements"]:
                key = tuple(nodes)
                if key == prev:
                    self._last_physical.extend(phys)
                    continue
                prev = key
                self._insert(b["type"], nodes, gi2mi, coords, list(phys))

    def _add_point(self, xyz):
        if self.dim_world == 2 and xyz[2] != 0:
            raise GmshError("In a 2D GmshMesh, the z-coordinate of every node must be zero")
        self.node_xy.append(xyz[:self.dim_world])
        return len(self.node_xy) - 1

    def _insert(self, etype, nodes, gi2mi, coords, phys):
        if etype == 15:
            mi = gi2mi.get(nodes[0])
            if mi is None:  # auxiliary nodes are not part of the mesh (:235-245)
                self._last_physical = []
                return
            self.physical[2].setdefault(mi, []).extend(phys)
            # a repeated point element is appended to mi2gi[dim_mesh].back() (:224-229 / :470-474), i.e. to the node with
            # the HIGHEST mesh index registered so far
            self._last_physical = self.physical[2].setdefault(max(self.physical[2]), [])
            return
        if etype not in SUPPORTED:
            raise GmshError("Gmsh element type %d not (yet) supported by GmshReader." % etype)
        n_main, _order = SUPPORTED[etype]
        codim = self.dim_mesh - ELEMENT_TYPES[etype][1]
        xy = [coords[t][:self.dim_world] for t in nodes]
        self.entities[codim].append((etype, [gi2mi[t] for t in nodes[:n_main]], xy))
        self.physical[codim].append(phys)
        self._last_physical = phys

    # -- what the MeshFactory received -------------------------------------------------------------------------------------
    def arrays(self):
        """node_xy [n][2], edge_nodes uint32 [n_explicit][2], cell_nodes uint32 [n_cells][4] (0xFFFFFFFF pad), order (1|2)."""
        xy = np.array(self.node_xy, dtype=np.float64).reshape(-1, self.dim_world)
        en = np.array([m for _, m, _ in self.entities[1]], dtype=np.uint32).reshape(-1, 2)
        cn = np.full((len(self.entities[0]), 4), 0xFFFFFFFF, dtype=np.uint32)
        for i, (_, m, _) in enumerate(self.entities[0]):
            cn[i, :len(m)] = m
        order = max([SUPPORTED[t][1] for t, _, _ in self.entities[0] + self.entities[1]] or [1])
        return xy, en, cn, order

    # -- physical entities (gmsh_reader.cc:17-118) ---------------------------------------------------------------------------
    def physical_entity_nr(self, codim, index):
        if codim == 2:
            return list(self.physical[2].get(index, []))
        return list(self.physical[codim][index]) if index < len(self.physical[codim]) else []

    def is_physical_entity(self, codim, index, nr):
        return nr in self.physical_entity_nr(codim, index)

    def physical_entities(self, codim):
        return sorted((nr, name) for nr, name, cd in self.names if cd == codim)

    def name2nr(self, name, codim=None):
        hits = [(nr, cd) for nr, n, cd in self.names if n == name]
        if not hits:
            raise GmshError("No Physical Entity with this name found.")
        if codim is None:
            if len(hits) > 1:
                raise GmshError("There are multiple physical entities with the name " + name + ", please specify also the codimension.")
            return hits[0][0]
        for nr, cd in hits:
            if cd == codim:
                return nr
        raise GmshError("Physical Entity with name='%s' and codimension=%d' not found." % (name, codim))

    def nr2name(self, nr, codim=None):
        hits = [(n, cd) for k, n, cd in self.names if k == nr]
        if not hits:
            raise GmshError("Physical entity with number %d not found." % nr)
        if codim is None:
            if len(hits) > 1:
                raise GmshError("There are multiple physical entities with the Number %d, please specify also the codimension" % nr)
            return hits[0][0]
        for n, cd in hits:
            if cd == codim:
                return n
        raise GmshError("Physical entity with number=%d, codim=%d not found." % (nr, codim))
