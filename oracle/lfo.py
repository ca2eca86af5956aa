"""ORACLE bindings (test infrastructure, NOT product code).

ctypes wrapper over oracle/liblfo_oracle.so, the CPU restatement of the LehrFEM++ assembly path
(see oracle/lfo_base.h for scope and reference citations).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
NIL = 0xFFFFFFFF


def build(force=False):
    so = os.path.join(_HERE, "liblfo_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".h", ".cc", ".inc"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


class Coeff(C.Structure):
    _fields_ = [("kind", C.c_int), ("c", C.c_double * 4), ("table", C.c_void_p), ("stride", C.c_long),
                ("fn", C.c_void_p), ("fn2", C.c_void_p)]


_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_double)
_FN2 = C.CFUNCTYPE(None, C.c_double, C.c_double, C.POINTER(C.c_double))


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.lfo_last_error.restype = C.c_char_p
        for name in ("lfo_mesh_tp_tria", "lfo_mesh_tp_quad"):
            f = getattr(L, name)
            f.restype = C.c_void_p
            f.argtypes = [C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_double]
        L.lfo_mesh_hybrid.restype = C.c_void_p
        L.lfo_mesh_hybrid.argtypes = [C.c_uint, C.c_double, C.c_uint64]
        L.lfo_mesh_from_arrays.restype = C.c_void_p
        L.lfo_mesh_from_arrays.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int64, C.c_void_p]
        L.lfo_mesh_refine_regular.restype = C.c_void_p
        L.lfo_mesh_refine_regular.argtypes = [C.c_void_p]
        L.lfo_mesh_free.argtypes = [C.c_void_p]
        L.lfo_mesh_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 5
        L.lfo_mesh_export.argtypes = [C.c_void_p] * 8
        L.lfo_dofh_create.restype = C.c_void_p
        L.lfo_dofh_create.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint]
        L.lfo_dofh_create_dynamic.restype = C.c_void_p
        L.lfo_dofh_create_dynamic.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.lfo_dofh_free.argtypes = [C.c_void_p]
        L.lfo_dofh_num_dofs.restype = C.c_int64
        L.lfo_dofh_num_dofs.argtypes = [C.c_void_p]
        L.lfo_dofh_stride.argtypes = [C.c_void_p]
        L.lfo_dofh_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.lfo_dofh_dof_entities.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.lfo_assemble_test_matrix.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.lfo_assemble_test_vector.argtypes = [C.c_void_p, C.c_void_p]
        L.lfo_assemble_rd.restype = C.c_void_p
        L.lfo_assemble_rd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Coeff), C.POINTER(Coeff),
                                      C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.lfo_assemble_fixed.restype = C.c_void_p
        L.lfo_assemble_fixed.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int, C.c_void_p]
        L.lfo_boundary_edges.argtypes = [C.c_void_p, C.c_void_p]
        L.lfo_assemble_boundary_test_matrix.argtypes = [C.c_void_p, C.c_void_p]
        L.lfo_edge_matrices.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(Coeff), C.c_void_p, C.c_int]
        L.lfo_edge_vectors.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(Coeff), C.c_void_p, C.c_int]
        L.lfo_assemble_rd_edge.restype = C.c_void_p
        L.lfo_assemble_rd_edge.argtypes = [C.c_void_p, C.c_int, C.POINTER(Coeff), C.POINTER(Coeff), C.POINTER(Coeff), C.c_int, C.c_void_p,
                                           C.c_int]
        L.lfo_assemble_edge_load.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(Coeff), C.c_void_p, C.c_void_p]
        L.lfo_fix_coo.restype = C.c_void_p
        L.lfo_fix_coo.argtypes = [C.c_int64, C.c_int64] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
        L.lfo_fix_coo_lse.restype = C.c_void_p
        L.lfo_fix_coo_lse.argtypes = [C.c_int64, C.c_int64] + [C.c_void_p] * 3 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.lfo_cm_sizes.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 3
        L.lfo_cm_export.argtypes = [C.c_void_p] * 4
        L.lfo_cm_free.argtypes = [C.c_void_p]
        L.lfo_assemble_load.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Coeff), C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_double)]
        L.lfo_element_matrices.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Coeff), C.POINTER(Coeff),
                                           C.c_void_p, C.c_int]
        L.lfo_fe_element_matrices.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(Coeff), C.c_void_p, C.c_int]
        L.lfo_fespace_num_dofs.restype = C.c_int64
        L.lfo_fespace_num_dofs.argtypes = [C.c_void_p, C.c_int]
        L.lfo_fespace_cell_dofs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.lfo_nodal_projection.argtypes = [C.c_void_p, C.c_int, C.POINTER(Coeff), C.c_void_p]
        L.lfo_quad_rule.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.lfo_eval_fe.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.lfo_qp_coords.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.lfo_builtin_scalar.restype = C.c_double
        L.lfo_builtin_scalar.argtypes = [C.c_int, C.c_double, C.c_double]
        _LIB = L
    return _LIB


class OracleError(RuntimeError):
    pass


def _check(ok):
    if not ok:
        raise OracleError(lib().lfo_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- coefficient descriptors ----------------------------------------------------------------------------------
class coeff:
    """Factory for coefficient descriptors understood by the oracle."""

    @staticmethod
    def const(v):
        c = Coeff(kind=0)
        c.c[0] = float(v)
        return c

    @staticmethod
    def const2x2(m):
        c = Coeff(kind=1)
        m = np.asarray(m, dtype=np.float64).reshape(4)
        for i in range(4):
            c.c[i] = m[i]
        return c

    @staticmethod
    def builtin(fid):
        c = Coeff(kind=3 if fid >= 100 else 2)
        c.c[0] = float(fid)
        return c

    @staticmethod
    def table(arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        c = Coeff(kind=4)
        c.table = arr.ctypes.data
        c.stride = 1 if arr.ndim == 1 else arr.shape[1]
        c._keep = arr
        return c

    @staticmethod
    def callback(f):
        cb = _FN(lambda x, y: float(f(x, y)))
        c = Coeff(kind=5)
        c.fn = C.cast(cb, C.c_void_p)
        c._keep = cb
        return c

    @staticmethod
    def callback2x2(f):
        def g(x, y, out):
            m = np.asarray(f(x, y), dtype=np.float64).reshape(4)
            for i in range(4):
                out[i] = m[i]
        cb = _FN2(g)
        c = Coeff(kind=6)
        c.fn2 = C.cast(cb, C.c_void_p)
        c._keep = cb
        return c


# ---- mesh --------------------------------------------------------------------------------------------------------
class Mesh:
    def __init__(self, handle):
        _check(handle)
        self.h = handle
        v = [C.c_int64() for _ in range(5)]
        lib().lfo_mesh_counts(self.h, *[C.byref(x) for x in v])
        self.n_nodes, self.n_edges, self.n_cells, self.n_tria, self.n_quad = [x.value for x in v]

    def __del__(self):
        if getattr(self, "h", None):
            lib().lfo_mesh_free(self.h)
            self.h = None

    @staticmethod
    def tp_tria(nx, ny, x0=0.0, y0=0.0, x1=1.0, y1=1.0):
        return Mesh(lib().lfo_mesh_tp_tria(nx, ny, x0, y0, x1, y1))

    @staticmethod
    def tp_quad(nx, ny, x0=0.0, y0=0.0, x1=1.0, y1=1.0):
        return Mesh(lib().lfo_mesh_tp_quad(nx, ny, x0, y0, x1, y1))

    @staticmethod
    def hybrid(n, jitter=0.2, seed=12345):
        return Mesh(lib().lfo_mesh_hybrid(n, jitter, seed))

    def refine_regular(self):
        """MeshHierarchy::RefineRegular(): the regularly refined mesh with the reference's numbering."""
        return Mesh(lib().lfo_mesh_refine_regular(self.h))

    @staticmethod
    def from_arrays(node_xy, cell_nodes, cell_coords=None, cell_geo=None, edge_nodes=None):
        node_xy = np.ascontiguousarray(node_xy, dtype=np.float64)
        cell_nodes = np.ascontiguousarray(cell_nodes, dtype=np.uint32)
        assert cell_nodes.ndim == 2 and cell_nodes.shape[1] == 4
        if cell_coords is not None:
            cell_coords = np.ascontiguousarray(cell_coords, dtype=np.float64)
        if cell_geo is not None:
            cell_geo = np.ascontiguousarray(cell_geo, dtype=np.uint8)
        ne = 0
        if edge_nodes is not None:
            edge_nodes = np.ascontiguousarray(edge_nodes, dtype=np.uint32)
            ne = edge_nodes.shape[0]
        return Mesh(lib().lfo_mesh_from_arrays(node_xy.shape[0], _p(node_xy), cell_nodes.shape[0], _p(cell_nodes),
                                               _p(cell_coords), _p(cell_geo), ne, _p(edge_nodes)))

    @staticmethod
    def from_golden(entry, scale=1.0):
        """Build GenerateHybrid2DTestMesh(selector, scale) from tests/golden/test_meshes.json."""
        if "builder" in entry:
            c = entry["corners"]
            return Mesh.tp_tria(entry["nx"], entry["ny"], c[0] * scale, c[1] * scale, c[2] * scale, c[3] * scale)
        xy = np.array(entry["nodes"], dtype=np.float64) * scale
        nc = len(entry["cells"])
        cn = np.full((nc, 4), NIL, dtype=np.uint32)
        cc = np.zeros((nc, 4, 2))
        geo = np.zeros(nc, dtype=np.uint8)
        for i, c in enumerate(entry["cells"]):
            cn[i, : len(c["nodes"])] = c["nodes"]
            if c["coords"] is not None:
                cc[i, : len(c["nodes"])] = c["coords"]
                geo[i] = 2 if c["geometry"] == "Parallelogram" else 1
        return Mesh.from_arrays(xy, cn, cc, geo)

    def export(self):
        nc, ne, nn = self.n_cells, self.n_edges, self.n_nodes
        out = dict(cell_type=np.zeros(nc, np.uint8), cell_nodes=np.zeros((nc, 4), np.uint32),
                   cell_coords=np.zeros((nc, 4, 2)), cell_edges=np.zeros((nc, 4), np.uint32),
                   cell_edge_ori=np.zeros((nc, 4), np.int8), edge_nodes=np.zeros((ne, 2), np.uint32),
                   node_coords=np.zeros((nn, 2)))
        rc = lib().lfo_mesh_export(self.h, _p(out["cell_type"]), _p(out["cell_nodes"]), _p(out["cell_coords"]),
                                   _p(out["cell_edges"]), _p(out["cell_edge_ori"]), _p(out["edge_nodes"]),
                                   _p(out["node_coords"]))
        _check(rc == 0)
        return out

    # ---- FE space level ------------------------------------------------------------------------------------
    def num_dofs(self, degree):
        n = lib().lfo_fespace_num_dofs(self.h, degree)
        _check(n >= 0)
        return n

    def cell_dofs(self, degree):
        stride = lib().lfo_fespace_cell_dofs(self.h, degree, None, None)
        _check(stride >= 0)
        d = np.zeros((self.n_cells, stride), np.int64)
        nl = np.zeros(self.n_cells, np.uint8)
        _check(lib().lfo_fespace_cell_dofs(self.h, degree, _p(d), _p(nl)) >= 0)
        return d, nl

    def assemble_rd(self, degree, alpha, gamma, qr_tria=-1, qr_quad=-1, active=None, csr=False, repeat=1):
        """Returns (outer, inner, values, n, timings) -- Eigen column-major arrays (csr=False) or CSR (csr=True)."""
        if active is not None:
            active = np.ascontiguousarray(active, dtype=np.uint8)
        ta, tm = C.c_double(), C.c_double()
        h = lib().lfo_assemble_rd(self.h, degree, qr_tria, qr_quad, C.byref(alpha), C.byref(gamma), _p(active),
                                  1 if csr else 0, repeat, C.byref(ta), C.byref(tm))
        _check(h)
        r, c, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        lib().lfo_cm_sizes(h, C.byref(r), C.byref(c), C.byref(nnz))
        outer = np.zeros(c.value + 1, np.int32)
        inner = np.zeros(nnz.value, np.int32)
        vals = np.zeros(nnz.value)
        lib().lfo_cm_export(h, _p(outer), _p(inner), _p(vals))
        lib().lfo_cm_free(h)
        return outer, inner, vals, (r.value, c.value), dict(assemble_s=ta.value, makesparse_s=tm.value)

    def assemble_fixed(self, degree, alpha, gamma, f, fixed, fixed_vals, csr=False, alt=False):
        """Matrix + load vector with constant coefficients, then FixFlaggedSolutionComponents, then makeSparse.
        Returns (outer, inner, values, rhs)."""
        fixed = np.ascontiguousarray(fixed, dtype=np.uint8)
        fixed_vals = np.ascontiguousarray(fixed_vals, dtype=np.float64)
        n = self.num_dofs(degree)
        rhs = np.zeros(n)
        h = lib().lfo_assemble_fixed(self.h, degree, alpha, gamma, f, _p(fixed), _p(fixed_vals), 1 if csr else 0, 1 if alt else 0, _p(rhs))
        _check(h)
        r, c, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        lib().lfo_cm_sizes(h, C.byref(r), C.byref(c), C.byref(nnz))
        outer = np.zeros(c.value + 1, np.int32)
        inner = np.zeros(nnz.value, np.int32)
        vals = np.zeros(nnz.value)
        lib().lfo_cm_export(h, _p(outer), _p(inner), _p(vals))
        lib().lfo_cm_free(h)
        return outer, inner, vals, rhs

    # ---- edge (codim-1) contributions ---------------------------------------------------------------------------
    def boundary_edges(self):
        f = np.zeros(self.n_edges, np.uint8)
        _check(lib().lfo_boundary_edges(self.h, _p(f)) == 0)
        return f

    def edge_matrices(self, degree, eta, qr_degree=-1):
        """MassEdgeMatrixProvider::Eval of every edge -> [edge][row][col]"""
        s = degree + 1
        out = np.zeros((self.n_edges, s, s))
        _check(lib().lfo_edge_matrices(self.h, degree, qr_degree, C.byref(eta), _p(out), s) == 0)
        return out.transpose(0, 2, 1).copy()

    def edge_vectors(self, degree, g, qr_degree=-1):
        s = degree + 1
        out = np.zeros((self.n_edges, s))
        _check(lib().lfo_edge_vectors(self.h, degree, qr_degree, C.byref(g), _p(out), s) == 0)
        return out

    def assemble_rd_edge(self, degree, alpha, gamma, eta, edge_mask=None, qr_degree=-1, csr=False):
        """cell reaction-diffusion matrix + edge mass matrix of the flagged edges in one COO, makeSparse -> (outer, inner, values)"""
        if edge_mask is not None:
            edge_mask = np.ascontiguousarray(edge_mask, dtype=np.uint8)
        h = lib().lfo_assemble_rd_edge(self.h, degree, C.byref(alpha), C.byref(gamma), C.byref(eta), qr_degree, _p(edge_mask),
                                       1 if csr else 0)
        _check(h)
        r, c, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        lib().lfo_cm_sizes(h, C.byref(r), C.byref(c), C.byref(nnz))
        outer = np.zeros(c.value + 1, np.int32)
        inner = np.zeros(nnz.value, np.int32)
        vals = np.zeros(nnz.value)
        lib().lfo_cm_export(h, _p(outer), _p(inner), _p(vals))
        lib().lfo_cm_free(h)
        return outer, inner, vals

    def assemble_edge_load(self, degree, g, edge_mask=None, qr_degree=-1, out=None):
        if out is None:
            out = np.zeros(self.num_dofs(degree))
        if edge_mask is not None:
            edge_mask = np.ascontiguousarray(edge_mask, dtype=np.uint8)
        _check(lib().lfo_assemble_edge_load(self.h, degree, qr_degree, C.byref(g), _p(edge_mask), _p(out)) == 0)
        return out

    def assemble_load(self, degree, f, qr_tria=-1, qr_quad=-1, active=None, out=None):
        n = self.num_dofs(degree)
        if out is None:
            out = np.zeros(n)
        if active is not None:
            active = np.ascontiguousarray(active, dtype=np.uint8)
        t = C.c_double()
        _check(lib().lfo_assemble_load(self.h, degree, qr_tria, qr_quad, C.byref(f), _p(active), _p(out), C.byref(t)) == 0)
        return out, t.value

    def element_matrices(self, degree, alpha, gamma, qr_tria=-1, qr_quad=-1):
        stride = {1: 4, 2: 9, 3: 16}[degree]
        out = np.zeros((self.n_cells, stride, stride))
        _check(lib().lfo_element_matrices(self.h, degree, qr_tria, qr_quad, C.byref(alpha), C.byref(gamma), _p(out),
                                          stride) == 0)
        return out.transpose(0, 2, 1).copy()  # -> [cell][row][col]

    def fe_element_matrices(self, degree, which, coeff_):
        """lf::fe::DiffusionElementMatrixProvider (which="diffusion") / MassElementMatrixProvider (which="mass") per cell."""
        stride = {1: 4, 2: 9, 3: 16}[degree]
        out = np.zeros((self.n_cells, stride, stride))
        _check(lib().lfo_fe_element_matrices(self.h, degree, 1 if which == "mass" else 0, C.byref(coeff_), _p(out), stride) == 0)
        return out.transpose(0, 2, 1).copy()

    def nodal_projection(self, degree, u):
        out = np.zeros(self.num_dofs(degree))
        _check(lib().lfo_nodal_projection(self.h, degree, C.byref(u), _p(out)) == 0)
        return out

    def qp_coords(self, qr_tria, qr_quad):
        nt, nq = quad_rule(3, qr_tria)[1].size, quad_rule(4, qr_quad)[1].size
        nqm = max(nt, nq)
        out = np.zeros((self.n_cells, nqm, 2))
        _check(lib().lfo_qp_coords(self.h, qr_tria, qr_quad, nqm, _p(out)) == 0)
        return out


class DofHandler:
    """UniformFEDofHandler(mesh, {Point: n_pt, Segment: n_seg, Tria: n_tria, Quad: n_quad})."""

    def __init__(self, mesh, n_pt=0, n_seg=0, n_tria=0, n_quad=0, _handle=None):
        self.mesh = mesh
        self.h = _handle if _handle is not None else lib().lfo_dofh_create(mesh.h, n_pt, n_seg, n_tria, n_quad)
        _check(self.h)
        self.num_dofs = lib().lfo_dofh_num_dofs(self.h)
        self.stride = lib().lfo_dofh_stride(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().lfo_dofh_free(self.h)
            self.h = None

    @staticmethod
    def dynamic(mesh, n_int_node=None, n_int_edge=None, n_int_cell=None):
        """DynamicFEDofHandler(mesh, locdof) with locdof tabulated per node / edge / cell (None = 0 everywhere)."""
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.uint32) for a in (n_int_node, n_int_edge, n_int_cell)]
        for a, n in zip(arrs, (mesh.n_nodes, mesh.n_edges, mesh.n_cells)):
            assert a is None or a.shape == (n,)
        h = lib().lfo_dofh_create_dynamic(mesh.h, *[_p(a) for a in arrs])
        return DofHandler(mesh, _handle=h)

    def cell_dofs(self):
        d = np.zeros((self.mesh.n_cells, self.stride), np.int64)
        nl = np.zeros(self.mesh.n_cells, np.uint8)
        _check(lib().lfo_dofh_export(self.h, _p(d), _p(nl)) == 0)
        return d, nl

    def dof_entities(self):
        cd = np.zeros(self.num_dofs, np.uint8)
        ix = np.zeros(self.num_dofs, np.uint32)
        _check(lib().lfo_dofh_dof_entities(self.h, _p(cd), _p(ix)) == 0)
        return cd, ix

    def test_matrix(self, kind):
        out = np.zeros((self.num_dofs, self.num_dofs))
        _check(lib().lfo_assemble_test_matrix(self.h, kind, _p(out)) == 0)
        return out

    def test_vector(self):
        out = np.zeros(self.num_dofs)
        _check(lib().lfo_assemble_test_vector(self.h, _p(out)) == 0)
        return out

    def boundary_test_matrix(self):
        """AssembleMatrixLocally(1, dofh, BoundaryAssembler) of assemble/test/assembly_tests.cc:491-590 (dense)"""
        out = np.zeros((self.num_dofs, self.num_dofs))
        _check(lib().lfo_assemble_boundary_test_matrix(self.h, _p(out)) == 0)
        return out


def quad_rule(ref_el_id, degree):
    n = lib().lfo_quad_rule(ref_el_id, degree, None, None, 0)
    _check(n >= 0)
    dim = 1 if ref_el_id == 2 else 2
    pts = np.zeros((dim, n))
    w = np.zeros(n)
    _check(lib().lfo_quad_rule(ref_el_id, degree, _p(pts), _p(w), n) == n)
    return pts, w


def eval_fe(degree, ref_el_id, pts):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    n = pts.shape[1]
    nsf = lib().lfo_eval_fe(degree, ref_el_id, 0, None, None, None, None)
    _check(nsf > 0)
    phi = np.zeros((nsf, n))
    grad = np.zeros((nsf, 2 * n))
    nodes = np.zeros((2, nsf))
    _check(lib().lfo_eval_fe(degree, ref_el_id, n, _p(pts), _p(phi), _p(grad), _p(nodes)) == nsf)
    return phi, grad, nodes


def builtin_scalar(fid, x, y):
    return lib().lfo_builtin_scalar(fid, x, y)


def fix_coo(n, rows, cols, vals, fixed, fixed_vals, rhs, alt=False):
    """FixFlaggedSolutionComponents on an n x n triplet list; returns (outer, inner, values) of makeSparse() (column-major)
    and the modified right-hand side."""
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    fixed = np.ascontiguousarray(fixed, dtype=np.uint8)
    fixed_vals = np.ascontiguousarray(fixed_vals, dtype=np.float64)
    rhs = np.array(rhs, dtype=np.float64)
    h = lib().lfo_fix_coo(n, len(vals), _p(rows), _p(cols), _p(vals), _p(fixed), _p(fixed_vals), 1 if alt else 0, _p(rhs))
    _check(h)
    r, c, nnz = C.c_int64(), C.c_int64(), C.c_int64()
    lib().lfo_cm_sizes(h, C.byref(r), C.byref(c), C.byref(nnz))
    outer = np.zeros(c.value + 1, np.int32)
    inner = np.zeros(nnz.value, np.int32)
    out = np.zeros(nnz.value)
    lib().lfo_cm_export(h, _p(outer), _p(inner), _p(out))
    lib().lfo_cm_free(h)
    return outer, inner, out, rhs


def fix_coo_lse(n, rows, cols, vals, pairs, rhs):
    """FixSolutionComponentsLse on an n x n triplet list with prescribed components [(index, value), ...]; returns
    (outer, inner, values) of makeSparse() (column-major) and the modified right-hand side."""
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    idx = np.ascontiguousarray([p[0] for p in pairs], dtype=np.int64)
    val = np.ascontiguousarray([p[1] for p in pairs], dtype=np.float64)
    rhs = np.array(rhs, dtype=np.float64)
    h = lib().lfo_fix_coo_lse(n, len(vals), _p(rows), _p(cols), _p(vals), len(idx), _p(idx), _p(val), _p(rhs))
    _check(h)
    r, c, nnz = C.c_int64(), C.c_int64(), C.c_int64()
    lib().lfo_cm_sizes(h, C.byref(r), C.byref(c), C.byref(nnz))
    outer = np.zeros(c.value + 1, np.int32)
    inner = np.zeros(nnz.value, np.int32)
    out = np.zeros(nnz.value)
    lib().lfo_cm_export(h, _p(outer), _p(inner), _p(out))
    lib().lfo_cm_free(h)
    return outer, inner, out, rhs
