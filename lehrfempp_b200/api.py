"""ctypes binding of include/lfgpu.h (host plumbing; no numerics happen in Python)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

COL_MAJOR, ROW_MAJOR = 0, 1
ALGO_AUTO, ALGO_ATOMIC, ALGO_GATHER, ALGO_FAN = 0, 1, 2, 3
NIL = 0xFFFFFFFF

_STATUS = {-1: "INVALID", -2: "CUDA", -3: "NO_DEVICE", -4: "MISSING_RULE", -5: "DEGENERATE", -6: "OVERFLOW",
           -7: "UNSUPPORTED", -8: "NCCL"}


class LfgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("lfgpu error %s (%d): %s" % (_STATUS.get(code, "?"), code, msg))
        self.code = code


def library_path():
    return os.path.join(_HERE, "liblfgpu.so")


def build_library(force=False):
    """Compile liblfgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc"), "-s", "-j8"]
    if force:
        subprocess.check_call(args + ["clean"])
    subprocess.check_call(args)
    return library_path()


class _CQuad(C.Structure):
    _fields_ = [("n", C.c_int), ("points", C.c_void_p), ("weights", C.c_void_p)]


class _CCoeff(C.Structure):
    _fields_ = [("kind", C.c_int), ("c", C.c_double * 4), ("data", C.c_void_p), ("stride", C.c_int64)]


def _lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise ImportError("liblfgpu.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
                          "`make -C lehrfempp_b200/csrc`. There is no CPU fallback." % path)
    L = C.CDLL(path)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
    pp = C.POINTER(C.c_void_p)
    sig = {
        "lfgpu_ctx_create": (i32, [i32, pp]),
        "lfgpu_ctx_destroy": (None, [vp]),
        "lfgpu_last_error": (C.c_char_p, [vp]),
        "lfgpu_ctx_synchronize": (i32, [vp]),
        "lfgpu_ctx_stream": (vp, [vp]),
        "lfgpu_ctx_kernel_launches": (i64, [vp]),
        "lfgpu_version": (C.c_char_p, []),
        "lfgpu_event_create": (i32, [vp, pp]),
        "lfgpu_event_record": (i32, [vp, vp]),
        "lfgpu_event_elapsed_ms": (i32, [vp, vp, vp, C.POINTER(dbl)]),
        "lfgpu_event_destroy": (i32, [vp, vp]),
        "lfgpu_malloc": (i32, [vp, i64, pp]),
        "lfgpu_free": (i32, [vp, vp]),
        "lfgpu_memset": (i32, [vp, vp, i32, i64]),
        "lfgpu_memcpy_h2d": (i32, [vp, vp, vp, i64]),
        "lfgpu_memcpy_d2h": (i32, [vp, vp, vp, i64]),
        "lfgpu_host_alloc_pinned": (i32, [vp, i64, pp]),
        "lfgpu_host_free_pinned": (i32, [vp, vp]),
        "lfgpu_mesh_upload": (i32, [vp, i64, vp, i64, vp, vp, pp]),
        "lfgpu_mesh_tp_tria": (i32, [vp, C.c_uint32, C.c_uint32, dbl, dbl, dbl, dbl, pp]),
        "lfgpu_mesh_tp_quad": (i32, [vp, C.c_uint32, C.c_uint32, dbl, dbl, dbl, dbl, pp]),
        "lfgpu_mesh_hybrid": (i32, [vp, C.c_uint32, dbl, C.c_uint64, pp]),
        "lfgpu_mesh_build_topology": (i32, [vp, vp, i64, vp, vp]),
        "lfgpu_mesh_refine_regular": (i32, [vp, vp, pp]),
        "lfgpu_mesh_counts": (i32, [vp] + [C.POINTER(i64)] * 5),
        "lfgpu_mesh_download": (i32, [vp] * 9),
        "lfgpu_mesh_update_node_coords": (i32, [vp, vp, vp]),
        "lfgpu_mesh_destroy": (None, [vp]),
        "lfgpu_gmsh_read_file": (i32, [C.c_char_p, i32, pp]),
        "lfgpu_gmsh_read_memory": (i32, [vp, i64, i32, pp]),
        "lfgpu_gmsh_destroy": (None, [vp]),
        "lfgpu_gmsh_counts": (i32, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]),
        "lfgpu_gmsh_arrays": (i32, [vp, vp, vp, vp]),
        "lfgpu_gmsh_physical_entity_nr": (i32, [vp, i32, i64, i32, vp]),
        "lfgpu_gmsh_physical_flags": (i32, [vp, i32, C.c_uint32, i64, vp]),
        "lfgpu_gmsh_physical_name": (i32, [vp, i32, C.POINTER(C.c_uint32), C.POINTER(i32), C.c_char_p, i32]),
        "lfgpu_gmsh_physical_name2nr": (i32, [vp, C.c_char_p, i32, C.POINTER(C.c_uint32)]),
        "lfgpu_gmsh_physical_nr2name": (i32, [vp, C.c_uint32, i32, C.c_char_p, i32]),
        "lfgpu_gmsh_mesh": (i32, [vp, vp, pp]),
        "lfgpu_dofmap_upload": (i32, [vp, vp, i64, i32, vp, vp, pp]),
        "lfgpu_dofmap_uniform": (i32, [vp, vp, i32, i32, i32, i32, pp]),
        "lfgpu_dofmap_lagrange": (i32, [vp, vp, i32, pp]),
        "lfgpu_dofmap_dynamic": (i32, [vp, vp, vp, vp, vp, pp]),
        "lfgpu_dofmap_num_dofs": (i64, [vp]),
        "lfgpu_dofmap_stride": (i32, [vp]),
        "lfgpu_dofmap_download": (i32, [vp, vp, vp, vp]),
        "lfgpu_dofmap_destroy": (None, [vp]),
        "lfgpu_symbolic": (i32, [vp, vp, vp, vp, i32, pp]),
        "lfgpu_pattern_nnz": (i64, [vp]),
        "lfgpu_pattern_rows": (i64, [vp]),
        "lfgpu_pattern_cols": (i64, [vp]),
        "lfgpu_pattern_download": (i32, [vp, vp, vp, vp]),
        "lfgpu_pattern_outer_device": (vp, [vp]),
        "lfgpu_pattern_inner_device": (vp, [vp]),
        "lfgpu_pattern_destroy": (None, [vp]),
        "lfgpu_pattern_restrict_rows": (i32, [vp, vp, vp]),
        "lfgpu_assemble_reaction_diffusion": (i32, [vp, vp, vp, i32, C.POINTER(_CQuad), C.POINTER(_CQuad), C.POINTER(_CCoeff),
                                                    C.POINTER(_CCoeff), vp, dbl, vp, i32]),
        "lfgpu_assemble_reaction_diffusion_rows": (i32, [vp, vp, vp, i32, C.POINTER(_CQuad), C.POINTER(_CQuad), C.POINTER(_CCoeff),
                                                         C.POINTER(_CCoeff), vp, dbl, vp, i32, vp, i64]),
        "lfgpu_assemble_reaction_diffusion_range": (i32, [vp, vp, vp, i32, C.POINTER(_CQuad), C.POINTER(_CQuad), C.POINTER(_CCoeff),
                                                          C.POINTER(_CCoeff), dbl, vp, i32, i64, i64]),
        "lfgpu_assemble_load": (i32, [vp, vp, vp, i32, C.POINTER(_CQuad), C.POINTER(_CQuad), C.POINTER(_CCoeff), vp, dbl, vp, i32]),
        "lfgpu_assemble_edge_mass": (i32, [vp, vp, vp, vp, i32, C.POINTER(_CQuad), C.POINTER(_CCoeff), vp, vp]),
        "lfgpu_assemble_edge_load": (i32, [vp, vp, vp, i32, C.POINTER(_CQuad), C.POINTER(_CCoeff), vp, vp]),
        "lfgpu_assemble_segment_mass": (i32, [vp, vp, i32, C.POINTER(_CQuad), i64, vp, vp, C.POINTER(_CCoeff), vp]),
        "lfgpu_assemble_segment_load": (i32, [vp, i32, C.POINTER(_CQuad), i64, vp, vp, C.POINTER(_CCoeff), i64, vp]),
        "lfgpu_edge_qp_coords": (i32, [vp, vp, i32, C.POINTER(_CQuad), i32, vp]),
        "lfgpu_mesh_boundary_edges": (i32, [vp, vp, vp]),
        "lfgpu_mesh_boundary_nodes": (i32, [vp, vp, vp]),
        "lfgpu_dofmap_boundary_dofs": (i32, [vp, vp, vp, vp]),
        "lfgpu_dofmap_edge_dof_flags": (i32, [vp, vp, vp, vp, vp]),
        "lfgpu_dofmap_dof_coords": (i32, [vp, vp, vp, i32, i32, vp]),
        "lfgpu_fix_flagged_solution_components": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "lfgpu_fix_flagged_solution_comp_alt": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "lfgpu_assemble_reaction_diffusion_host": (i32, [vp, vp, vp, i32, C.POINTER(_CQuad), C.POINTER(_CQuad), C.POINTER(_CCoeff),
                                                         C.POINTER(_CCoeff), vp, vp, vp, i32, i32]),
        "lfgpu_assemble_reaction_diffusion_host_range": (i32, [vp, vp, vp, i32, C.POINTER(_CQuad), C.POINTER(_CQuad), C.POINTER(_CCoeff),
                                                               C.POINTER(_CCoeff), vp, vp, vp, i32, i32, i64, i64]),
        "lfgpu_spmv": (i32, [vp, vp, vp, vp, vp]),
        "lfgpu_cg_solve": (i32, [vp, vp, vp, vp, vp, dbl, i32, i32, vp, vp]),
        "lfgpu_rows_pack": (i32, [vp, vp, vp, i64, vp, vp, vp]),
        "lfgpu_rows_unpack_add": (i32, [vp, vp, vp, i64, vp, vp, vp]),
        "lfgpu_pattern_adj_ptr_device": (vp, [vp]),
        "lfgpu_pattern_adj_device": (vp, [vp]),
        "lfgpu_pattern_num_items": (i64, [vp]),
        "lfgpu_mesh_node_coords_device": (vp, [vp]),
        "lfgpu_mesh_cell_nodes_device": (vp, [vp]),
        "lfgpu_ctx_wait_event": (i32, [vp, vp]),
        "lfgpu_partition_morton": (i32, [vp, vp, i32, vp]),
        "lfgpu_partition_dof_owner": (i32, [vp, vp, vp, vp]),
        "lfgpu_partition_select_cells": (i32, [vp, vp, vp, vp, i32, i32, vp]),
        "lfgpu_submesh_extract": (i32, [vp, vp, vp, vp, pp]),
        "lfgpu_submesh_destroy": (None, [vp]),
        "lfgpu_submesh_mesh": (vp, [vp]),
        "lfgpu_submesh_dofmap": (vp, [vp]),
        "lfgpu_submesh_counts": (i32, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
        "lfgpu_submesh_l2g_cells_device": (vp, [vp]),
        "lfgpu_submesh_l2g_nodes_device": (vp, [vp]),
        "lfgpu_submesh_l2g_dofs_device": (vp, [vp]),
        "lfgpu_submesh_owned_dofs": (i32, [vp, vp, vp, i32, vp]),
        "lfgpu_multi_create": (i32, [vp, i32, pp]),
        "lfgpu_multi_destroy": (None, [vp]),
        "lfgpu_multi_num_devices": (i32, [vp]),
        "lfgpu_multi_ctx": (vp, [vp, i32]),
        "lfgpu_multi_last_error": (C.c_char_p, [vp]),
        "lfgpu_multi_setup": (i32, [vp, i64, vp, i64, vp, vp, i64, i32, vp, vp, i32]),
        "lfgpu_multi_set_zero": (i32, [vp]),
        "lfgpu_multi_assemble_reaction_diffusion": (i32, [vp, i32, C.POINTER(_CQuad), C.POINTER(_CQuad), C.POINTER(_CCoeff),
                                                          C.POINTER(_CCoeff), i32]),
        "lfgpu_multi_part_sizes": (i32, [vp, i32] + [C.POINTER(i64)] * 5),
        "lfgpu_multi_part_download": (i32, [vp, i32, vp, vp, vp, vp]),
        "lfgpu_qp_coords": (i32, [vp, vp, i32, C.POINTER(_CQuad), C.POINTER(_CQuad), i32, vp]),
        "lfgpu_fe_tabulate": (i32, [i32, i32, C.POINTER(_CQuad), vp, vp]),
        "lfgpu_default_quad_rule": (i32, [i32, i32, i32, vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    L._exported = sorted(sig)
    _LIB = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class QuadRule:
    """lf::quad::QuadRule: points [2][n], weights [n] (host)."""

    def __init__(self, points, weights):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.weights = np.ascontiguousarray(weights, dtype=np.float64)
        n = self.weights.size
        assert self.points.shape in ((2, n), (1, n), (n,))  # cells: [2][n]; segment rules: [n]
        self._c = _CQuad(self.weights.size, self.points.ctypes.data, self.weights.ctypes.data)

    def ref(self):
        return C.byref(self._c)


def _qref(q):
    return None if q is None else q.ref()


def default_quad_rule(cell_type, degree):
    """make_QuadRule(ref_el, degree) as tabulated inside liblfgpu (host data)."""
    L = _lib()
    n = L.lfgpu_default_quad_rule(cell_type, degree, 0, None, None)
    if n < 0:
        raise LfgpuError(n, "no such rule")
    pts = np.zeros(n) if cell_type == 2 else np.zeros((2, n))  # 2 = segment (RefEl::Id)
    w = np.zeros(n)
    L.lfgpu_default_quad_rule(cell_type, degree, n, _p(pts), _p(w))
    return QuadRule(pts, w)


def fe_tabulate(degree, cell_type, qr=None):
    L = _lib()
    nsf = {3: (3, 6, 10), 4: (4, 9, 16)}[cell_type][degree - 1]
    nq = qr.weights.size if qr is not None else default_quad_rule(cell_type, 2 * degree).weights.size
    phi = np.zeros((nsf, nq))
    grad = np.zeros((nsf, 2 * nq))
    rc = L.lfgpu_fe_tabulate(degree, cell_type, _qref(qr), _p(phi), _p(grad))
    if rc < 0:
        raise LfgpuError(rc, L.lfgpu_last_error(None).decode())
    return phi, grad


class Coeff:
    """Coefficient descriptor (a MeshFunction evaluated at the quadrature points)."""

    def __init__(self, kind, c=(0, 0, 0, 0), data=None, stride=0):
        self._c = _CCoeff(kind=kind)
        for i in range(4):
            self._c.c[i] = float(c[i])
        self._data = data  # keeps the DeviceArray (or, for the multi-device calls, the host array) alive
        if data is None:
            self._c.data = None
        elif isinstance(data, np.ndarray):  # HOST table over the cells of the whole mesh (lfgpu_multi_assemble_reaction_diffusion)
            self._c.data = data.ctypes.data
        else:
            self._c.data = data.ptr
        self._c.stride = stride

    @staticmethod
    def const(v):
        return Coeff(0, (v, 0, 0, 0))

    @staticmethod
    def const2x2(m):
        return Coeff(1, tuple(np.asarray(m, dtype=np.float64).reshape(4)))

    @staticmethod
    def per_cell(dev):
        return Coeff(2, data=dev, stride=1)

    @staticmethod
    def per_qp(dev, stride):
        return Coeff(3, data=dev, stride=stride)

    @staticmethod
    def per_qp_2x2(dev, stride):
        return Coeff(4, data=dev, stride=stride)

    @staticmethod
    def nodal(dev):
        """Continuous piecewise (bi)linear coefficient given by its values at the mesh nodes (LFGPU_COEFF_NODAL)."""
        return Coeff(5, data=dev, stride=1)

    def ref(self):
        return C.byref(self._c)


class Context:
    """One lfgpu_ctx: a GPU, a stream, error state."""

    def __init__(self, device=0):
        L = _lib()
        h = C.c_void_p()
        rc = L.lfgpu_ctx_create(device, C.byref(h))
        if rc != 0:
            raise LfgpuError(rc, L.lfgpu_last_error(None).decode())
        self.h = h
        self.L = L

    def close(self):
        if getattr(self, "h", None):
            for p in getattr(self, "_pinned", []):
                self.L.lfgpu_host_free_pinned(self.h, p)
            self._pinned = []
            self.L.lfgpu_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def check(self, rc):
        if rc != 0:
            raise LfgpuError(rc, self.L.lfgpu_last_error(self.h).decode())

    def synchronize(self):
        self.check(self.L.lfgpu_ctx_synchronize(self.h))

    @property
    def stream(self):
        return self.L.lfgpu_ctx_stream(self.h)

    @property
    def kernel_launches(self):
        return self.L.lfgpu_ctx_kernel_launches(self.h)

    # ---- timing -----------------------------------------------------------------------------------------------------
    def event(self):
        e = C.c_void_p()
        self.check(self.L.lfgpu_event_create(self.h, C.byref(e)))
        return e

    def record(self, ev):
        self.check(self.L.lfgpu_event_record(self.h, ev))

    def elapsed_ms(self, start, stop):
        ms = C.c_double()
        self.check(self.L.lfgpu_event_elapsed_ms(self.h, start, stop, C.byref(ms)))
        return ms.value

    def pinned(self, n, dtype=np.float64):
        """numpy view of a page-locked host buffer (freed with the context)."""
        dtype = np.dtype(dtype)
        p = C.c_void_p()
        self.check(self.L.lfgpu_host_alloc_pinned(self.h, int(n) * dtype.itemsize, C.byref(p)))
        buf = (C.c_char * (int(n) * dtype.itemsize)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(n))
        self._pinned = getattr(self, "_pinned", []) + [p]
        return arr

    def h2d_async(self, dev, host):
        self.check(self.L.lfgpu_memcpy_h2d(self.h, dev.ptr, _p(host), host.nbytes))

    def d2h_async(self, host, dev):
        self.check(self.L.lfgpu_memcpy_d2h(self.h, _p(host), dev.ptr, host.nbytes))

    # ---- memory -----------------------------------------------------------------------------------------------------
    def empty(self, n, dtype=np.float64):
        return DeviceArray(self, n, dtype)

    def zeros(self, n, dtype=np.float64):
        a = DeviceArray(self, n, dtype)
        self.check(self.L.lfgpu_memset(self.h, a.ptr, 0, a.nbytes))
        return a

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        a = DeviceArray(self, arr.size, arr.dtype)
        self.check(self.L.lfgpu_memcpy_h2d(self.h, a.ptr, _p(arr), arr.nbytes))
        self.synchronize()
        return a

    # ---- meshes -------------------------------------------------------------------------------------------------------
    def mesh_upload(self, node_coords, cell_nodes, cell_coords=None):
        node_coords = np.ascontiguousarray(node_coords, dtype=np.float64)
        cell_nodes = np.ascontiguousarray(cell_nodes, dtype=np.uint32)
        assert cell_nodes.ndim == 2 and cell_nodes.shape[1] == 4
        if cell_coords is not None:
            cell_coords = np.ascontiguousarray(cell_coords, dtype=np.float64)
        h = C.c_void_p()
        self.check(self.L.lfgpu_mesh_upload(self.h, node_coords.shape[0], _p(node_coords), cell_nodes.shape[0], _p(cell_nodes),
                                            _p(cell_coords), C.byref(h)))
        return Mesh(self, h)

    def mesh_tp_tria(self, nx, ny, x0=0.0, y0=0.0, x1=1.0, y1=1.0):
        h = C.c_void_p()
        self.check(self.L.lfgpu_mesh_tp_tria(self.h, nx, ny, x0, y0, x1, y1, C.byref(h)))
        return Mesh(self, h)

    def mesh_tp_quad(self, nx, ny, x0=0.0, y0=0.0, x1=1.0, y1=1.0):
        h = C.c_void_p()
        self.check(self.L.lfgpu_mesh_tp_quad(self.h, nx, ny, x0, y0, x1, y1, C.byref(h)))
        return Mesh(self, h)

    def mesh_hybrid(self, n, jitter=0.2, seed=12345):
        h = C.c_void_p()
        self.check(self.L.lfgpu_mesh_hybrid(self.h, n, jitter, seed, C.byref(h)))
        return Mesh(self, h)


class GmshReader:
    """lf::io::GmshReader over the C ABI (host side; `mesh(ctx)` gives reader.mesh() on the device)."""

    def __init__(self, source, dim_world=2):
        L = _lib()
        self.L = L
        h = C.c_void_p()
        if isinstance(source, (bytes, bytearray)):
            buf = bytes(source)
            rc = L.lfgpu_gmsh_read_memory(buf, len(buf), dim_world, C.byref(h))
        else:
            rc = L.lfgpu_gmsh_read_file(os.fsencode(source), dim_world, C.byref(h))
        self.h = h if rc == 0 else None
        self._check(rc)
        v = [C.c_int64() for _ in range(3)]
        order, nnames = C.c_int(), C.c_int()
        self._check(L.lfgpu_gmsh_counts(self.h, *[C.byref(x) for x in v], C.byref(order), C.byref(nnames)))
        self.n_nodes, self.n_explicit_edges, self.n_cells = [x.value for x in v]
        self.geometry_order = order.value
        self.n_physical_names = nnames.value

    def _check(self, rc):
        if rc < 0:
            raise LfgpuError(rc, (self.L.lfgpu_last_error(None) or b"").decode(errors="replace"))
        return rc

    def __del__(self):
        if getattr(self, "h", None):
            self.L.lfgpu_gmsh_destroy(self.h)
            self.h = None

    def arrays(self):
        """(node_coords [n][2], edge_nodes uint32 [n_explicit][2], cell_nodes uint32 [n_cells][4]) in AddPoint / AddEntity order."""
        xy = np.zeros((self.n_nodes, 2))
        en = np.zeros((self.n_explicit_edges, 2), np.uint32)
        cn = np.zeros((self.n_cells, 4), np.uint32)
        self._check(self.L.lfgpu_gmsh_arrays(self.h, _p(xy), _p(en), _p(cn)))
        return xy, en, cn

    def physical_entity_nr(self, codim, index):
        out = np.zeros(16, np.uint32)
        n = self._check(self.L.lfgpu_gmsh_physical_entity_nr(self.h, codim, index, out.size, _p(out)))
        if n > out.size:
            out = np.zeros(n, np.uint32)
            self._check(self.L.lfgpu_gmsh_physical_entity_nr(self.h, codim, index, n, _p(out)))
        return [int(x) for x in out[:n]]

    def physical_flags(self, codim, nr, n):
        """uint8 [n]: IsPhysicalEntity(entity i of the codimension, nr)."""
        out = np.zeros(n, np.uint8)
        self._check(self.L.lfgpu_gmsh_physical_flags(self.h, codim, nr, n, _p(out)))
        return out

    def physical_entities(self, codim):
        res = []
        for i in range(self.n_physical_names):
            nr, cd = C.c_uint32(), C.c_int()
            buf = C.create_string_buffer(256)
            self._check(self.L.lfgpu_gmsh_physical_name(self.h, i, C.byref(nr), C.byref(cd), buf, 256))
            if cd.value == codim:
                res.append((nr.value, buf.value.decode(errors="replace")))
        return sorted(res)

    def name2nr(self, name, codim=-1):
        nr = C.c_uint32()
        self._check(self.L.lfgpu_gmsh_physical_name2nr(self.h, name.encode(), codim, C.byref(nr)))
        return nr.value

    def nr2name(self, nr, codim=-1):
        buf = C.create_string_buffer(256)
        self._check(self.L.lfgpu_gmsh_physical_nr2name(self.h, nr, codim, buf, 256))
        return buf.value.decode(errors="replace")

    def mesh(self, ctx):
        h = C.c_void_p()
        ctx.check(self.L.lfgpu_gmsh_mesh(ctx.h, self.h, C.byref(h)))
        return Mesh(ctx, h)


class DeviceArray:
    def __init__(self, ctx, n, dtype=np.float64):
        self.ctx = ctx
        self.n = int(n)
        self.dtype = np.dtype(dtype)
        self.nbytes = self.n * self.dtype.itemsize
        p = C.c_void_p()
        ctx.check(ctx.L.lfgpu_malloc(ctx.h, self.nbytes, C.byref(p)))
        self.ptr = p

    def __del__(self):
        if getattr(self, "ptr", None) and getattr(self.ctx, "h", None):
            self.ctx.L.lfgpu_free(self.ctx.h, self.ptr)
            self.ptr = None

    def to_host(self, out=None):
        if out is None:
            out = np.empty(self.n, self.dtype)
        self.ctx.check(self.ctx.L.lfgpu_memcpy_d2h(self.ctx.h, _p(out), self.ptr, self.nbytes))
        self.ctx.synchronize()
        return out

    def copy_from_host(self, arr):
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        assert arr.size == self.n
        self.ctx.check(self.ctx.L.lfgpu_memcpy_h2d(self.ctx.h, self.ptr, _p(arr), self.nbytes))
        self.ctx.synchronize()


class Mesh:
    def __init__(self, ctx, handle, owner=None):
        self.ctx = ctx
        self.h = handle
        self._owner = owner  # a SubMesh that owns the handle (then this object must not destroy it)
        self._refresh()

    def _refresh(self):
        v = [C.c_int64() for _ in range(5)]
        self.ctx.check(self.ctx.L.lfgpu_mesh_counts(self.h, *[C.byref(x) for x in v]))
        self.n_nodes, self.n_edges, self.n_cells, self.n_tria, self.n_quad = [x.value for x in v]

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None) and getattr(self, "_owner", None) is None:
            self.ctx.L.lfgpu_mesh_destroy(self.h)
            self.h = None

    def build_topology(self, edge_nodes=None, cell_has_geometry=None):
        n = 0
        if edge_nodes is not None:
            edge_nodes = np.ascontiguousarray(edge_nodes, dtype=np.uint32)
            n = edge_nodes.shape[0]
        if cell_has_geometry is not None:
            cell_has_geometry = np.ascontiguousarray(cell_has_geometry, dtype=np.uint8)
            assert cell_has_geometry.size == self.n_cells
        self.ctx.check(self.ctx.L.lfgpu_mesh_build_topology(self.ctx.h, self.h, n, _p(edge_nodes), _p(cell_has_geometry)))
        self._refresh()

    def download(self, topology=False):
        nc, nn = self.n_cells, self.n_nodes
        out = dict(cell_type=np.zeros(nc, np.uint8), cell_nodes=np.zeros((nc, 4), np.uint32), cell_coords=np.zeros((nc, 4, 2)),
                   node_coords=np.zeros((nn, 2)))
        ce = co = en = None
        if topology:
            # edge count is known only after numbering: two-step download
            self.ctx.check(self.ctx.L.lfgpu_mesh_download(self.ctx.h, self.h, None, None, None, _p(np.zeros((nc, 4), np.uint32)),
                                                          None, None, None))
            self._refresh()
            ce = np.zeros((nc, 4), np.uint32)
            co = np.zeros((nc, 4), np.int8)
            en = np.zeros((self.n_edges, 2), np.uint32)
            out.update(cell_edges=ce, cell_edge_ori=co, edge_nodes=en)
        self.ctx.check(self.ctx.L.lfgpu_mesh_download(self.ctx.h, self.h, _p(out["cell_type"]), _p(out["cell_nodes"]),
                                                      _p(out["cell_coords"]), _p(ce), _p(co), _p(en), _p(out["node_coords"])))
        return out

    def refine_regular(self):
        """MeshHierarchy::RefineRegular(): the regularly refined mesh with the reference's numbering, built on the device."""
        h = C.c_void_p()
        self.ctx.check(self.ctx.L.lfgpu_mesh_refine_regular(self.ctx.h, self.h, C.byref(h)))
        self._refresh()  # the parent's edges are numbered now
        return Mesh(self.ctx, h)

    def boundary_edges(self):
        """DeviceArray(uint8)[n_edges]: edges with exactly one adjacent cell (flagEntitiesOnBoundary(mesh, 1))."""
        if self.n_edges == 0:
            self.build_topology()  # edges discovered from the cells (hybrid2d/mesh.cc:378-404)
        out = self.ctx.empty(max(self.n_edges, 1), np.uint8)
        self.ctx.check(self.ctx.L.lfgpu_mesh_boundary_edges(self.ctx.h, self.h, out.ptr))
        return out

    def boundary_nodes(self):
        """DeviceArray(uint8)[n_nodes]: endpoints of boundary edges (flagEntitiesOnBoundary(mesh, 2))."""
        if self.n_edges == 0:
            self.build_topology()
        out = self.ctx.empty(max(self.n_nodes, 1), np.uint8)
        self.ctx.check(self.ctx.L.lfgpu_mesh_boundary_nodes(self.ctx.h, self.h, out.ptr))
        return out

    def edge_qp_coords(self, degree, qr_segment=None, nq_stride=None):
        """[n_edges][nq_stride][2] global coordinates of the edge quadrature points (host array)."""
        nq = qr_segment.weights.size if qr_segment is not None else degree + 1
        nq_stride = nq_stride or nq
        if self.n_edges == 0:
            self.build_topology()
        out = self.ctx.empty(max(self.n_edges, 1) * nq_stride * 2)
        self.ctx.check(self.ctx.L.lfgpu_edge_qp_coords(self.ctx.h, self.h, degree, _qref(qr_segment), nq_stride, out.ptr))
        return out.to_host()[: self.n_edges * nq_stride * 2].reshape(self.n_edges, nq_stride, 2)

    def update_node_coords(self, xy):
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        assert xy.shape == (self.n_nodes, 2)
        self.ctx.check(self.ctx.L.lfgpu_mesh_update_node_coords(self.ctx.h, self.h, _p(xy)))

    # ---- dof maps ---------------------------------------------------------------------------------------------------
    def dofmap_uniform(self, n_pt=0, n_seg=0, n_tria=0, n_quad=0):
        h = C.c_void_p()
        self.ctx.check(self.ctx.L.lfgpu_dofmap_uniform(self.ctx.h, self.h, n_pt, n_seg, n_tria, n_quad, C.byref(h)))
        self._refresh()
        return DofMap(self, h)

    def dofmap_lagrange(self, degree):
        h = C.c_void_p()
        self.ctx.check(self.ctx.L.lfgpu_dofmap_lagrange(self.ctx.h, self.h, degree, C.byref(h)))
        self._refresh()
        return DofMap(self, h)

    def dofmap_dynamic(self, n_int_node=None, n_int_edge=None, n_int_cell=None):
        """DynamicFEDofHandler(mesh, locdof), locdof tabulated per node / edge / cell (uint32 host arrays, None = 0)."""
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.uint32) for a in (n_int_node, n_int_edge, n_int_cell)]
        if arrs[1] is not None and self.n_edges == 0:
            self.build_topology()
        for a, n in zip(arrs, (self.n_nodes, self.n_edges, self.n_cells)):
            assert a is None or a.shape == (n,), "one count per entity"
        h = C.c_void_p()
        self.ctx.check(self.ctx.L.lfgpu_dofmap_dynamic(self.ctx.h, self.h, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), C.byref(h)))
        self._refresh()
        return DofMap(self, h)

    def dofmap_upload(self, n_dofs, cell_dofs, n_ldof=None):
        cell_dofs = np.ascontiguousarray(cell_dofs, dtype=np.int64)
        assert cell_dofs.shape[0] == self.n_cells
        if n_ldof is not None:
            n_ldof = np.ascontiguousarray(n_ldof, dtype=np.uint8)
        h = C.c_void_p()
        self.ctx.check(self.ctx.L.lfgpu_dofmap_upload(self.ctx.h, self.h, n_dofs, cell_dofs.shape[1], _p(cell_dofs), _p(n_ldof),
                                                      C.byref(h)))
        return DofMap(self, h)

    def qp_coords(self, degree, nq_stride, qr_tria=None, qr_quad=None):
        out = self.ctx.empty(self.n_cells * nq_stride * 2)
        self.ctx.check(self.ctx.L.lfgpu_qp_coords(self.ctx.h, self.h, degree, _qref(qr_tria), _qref(qr_quad), nq_stride, out.ptr))
        return out


class DofMap:
    def __init__(self, mesh, handle, owner=None):
        self.mesh = mesh
        self.ctx = mesh.ctx
        self.h = handle
        self._owner = owner
        self.num_dofs = self.ctx.L.lfgpu_dofmap_num_dofs(self.h)
        self.stride = self.ctx.L.lfgpu_dofmap_stride(self.h)

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None) and getattr(self, "_owner", None) is None:
            self.ctx.L.lfgpu_dofmap_destroy(self.h)
            self.h = None

    # ---- distributed ownership (include/lfgpu.h "distributed ownership") --------------------------------------------------
    def partition_morton(self, n_parts):
        """(cell_part, dof_owner): DeviceArray(uint8) [n_cells], [n_dofs] -- Morton cell ranges, dof owned by the lowest part touching it."""
        part = self.ctx.empty(self.mesh.n_cells, np.uint8)
        self.ctx.check(self.ctx.L.lfgpu_partition_morton(self.ctx.h, self.mesh.h, int(n_parts), part.ptr))
        owner = self.ctx.empty(self.num_dofs, np.uint8)
        self.ctx.check(self.ctx.L.lfgpu_partition_dof_owner(self.ctx.h, self.h, part.ptr, owner.ptr))
        return part, owner

    def submesh(self, cell_part, dof_owner, rank, halo=True):
        """The sub-problem of part `rank`: its cells (+ the one-cell halo around the dofs it owns when halo) as a SubMesh."""
        sel = self.ctx.empty(self.mesh.n_cells, np.uint8)
        self.ctx.check(self.ctx.L.lfgpu_partition_select_cells(self.ctx.h, self.h, cell_part.ptr, dof_owner.ptr, int(rank), 1 if halo else 0,
                                                               sel.ptr))
        h = C.c_void_p()
        self.ctx.check(self.ctx.L.lfgpu_submesh_extract(self.ctx.h, self.mesh.h, self.h, sel.ptr, C.byref(h)))
        return SubMesh(self.ctx, h, dof_owner, rank)

    def download(self):
        d = np.zeros((self.mesh.n_cells, self.stride), np.int64)
        nl = np.zeros(self.mesh.n_cells, np.uint8)
        self.ctx.check(self.ctx.L.lfgpu_dofmap_download(self.ctx.h, self.h, _p(d), _p(nl)))
        return d, nl

    def symbolic(self, trial=None, major=ROW_MAJOR):
        """Pattern for test = self, trial = `trial` (default: same space)."""
        trial = trial or self
        h = C.c_void_p()
        self.ctx.check(self.ctx.L.lfgpu_symbolic(self.ctx.h, self.mesh.h, self.h, trial.h, major, C.byref(h)))
        return Pattern(self.mesh, h, major)

    def boundary_dofs(self):
        """DeviceArray(uint8)[n_dofs]: dofs whose entity lies on the boundary (for device-numbered uniform layouts)."""
        if self.mesh.n_edges == 0:
            self.mesh.build_topology()
        out = self.ctx.empty(self.num_dofs, np.uint8)
        self.ctx.check(self.ctx.L.lfgpu_dofmap_boundary_dofs(self.ctx.h, self.mesh.h, self.h, out.ptr))
        return out

    def edge_dof_flags(self, edge_sel):
        """DeviceArray(uint8)[n_dofs]: dofs of the selected edges incl. their end points (InitEssentialConditionFromFunction, step 1)."""
        out = self.ctx.empty(self.num_dofs, np.uint8)
        self.ctx.check(self.ctx.L.lfgpu_dofmap_edge_dof_flags(self.ctx.h, self.mesh.h, self.h, edge_sel.ptr, out.ptr))
        return out

    def dof_coords(self, degree):
        """[n_dofs][2] host array: interpolation node of every dof of the degree-`degree` Lagrange layout."""
        n_tria, n_quad = {1: (0, 0), 2: (0, 1), 3: (1, 4)}[degree]
        out = self.ctx.empty(2 * self.num_dofs)
        self.ctx.check(self.ctx.L.lfgpu_dofmap_dof_coords(self.ctx.h, self.mesh.h, self.h, n_tria, n_quad, out.ptr))
        return out.to_host().reshape(-1, 2)

    def assemble_edge_load(self, degree, g, qr_segment=None, active_edges=None, out=None):
        """AssembleVectorLocally(1, dofh, ScalarLoadEdgeVectorProvider(fe_space, g, edge_sel), vec): accumulates into out."""
        if out is None:
            out = self.ctx.zeros(self.num_dofs)
        self.ctx.check(self.ctx.L.lfgpu_assemble_edge_load(self.ctx.h, self.mesh.h, self.h, degree, _qref(qr_segment), g.ref(),
                                                           active_edges.ptr if active_edges is not None else None, out.ptr))
        return out

    def assemble_load(self, degree, f, qr_tria=None, qr_quad=None, active=None, beta=0.0, out=None, algo=ALGO_AUTO):
        """AssembleVectorLocally(0, dofh, ScalarLoadElementVectorProvider(fe_space, f), vec)."""
        if out is None:
            out = self.ctx.zeros(self.num_dofs)
        self.ctx.check(self.ctx.L.lfgpu_assemble_load(self.ctx.h, self.mesh.h, self.h, degree, _qref(qr_tria), _qref(qr_quad),
                                                      f.ref(), active.ptr if active is not None else None, beta, out.ptr, algo))
        return out


class MultiAssembler:
    """Several GPUs from one process through the C ABI (lfgpu_multi_*): the flattened problem is cut into one sub-problem per
    listed device (distributed ownership, owner-computes); `parts()` returns every device's owned rows with global indices."""

    def __init__(self, device_ids):
        self.L = _lib()
        ids = (C.c_int * len(device_ids))(*device_ids)
        self.h = C.c_void_p()
        rc = self.L.lfgpu_multi_create(ids, len(device_ids), C.byref(self.h))
        if rc != 0:
            raise LfgpuError(rc, self.L.lfgpu_last_error(None).decode())
        self.n_dev = len(device_ids)

    def check(self, rc):
        if rc != 0:
            raise LfgpuError(rc, self.L.lfgpu_multi_last_error(self.h).decode())

    def setup(self, node_coords, cell_nodes, n_dofs, cell_dofs, n_ldof=None, cell_coords=None, major=ROW_MAJOR):
        xy = np.ascontiguousarray(node_coords, dtype=np.float64)
        cn = np.ascontiguousarray(cell_nodes, dtype=np.uint32)
        cd = np.ascontiguousarray(cell_dofs, dtype=np.int64)
        nl = None if n_ldof is None else np.ascontiguousarray(n_ldof, dtype=np.uint8)
        cc = None if cell_coords is None else np.ascontiguousarray(cell_coords, dtype=np.float64)
        self.check(self.L.lfgpu_multi_setup(self.h, xy.shape[0], _p(xy), cn.shape[0], _p(cn), _p(cc), int(n_dofs), cd.shape[1], _p(cd), _p(nl),
                                            major))

    def assemble_reaction_diffusion(self, degree, alpha, gamma, qr_tria=None, qr_quad=None, accumulate=False):
        self.check(self.L.lfgpu_multi_assemble_reaction_diffusion(self.h, degree, _qref(qr_tria), _qref(qr_quad), alpha.ref(), gamma.ref(),
                                                                  1 if accumulate else 0))

    def set_zero(self):
        self.check(self.L.lfgpu_multi_set_zero(self.h))

    def part_sizes(self, k):
        v = [C.c_int64() for _ in range(5)]
        self.check(self.L.lfgpu_multi_part_sizes(self.h, k, *[C.byref(x) for x in v]))
        return dict(zip(("rows", "nnz", "local_cells", "local_rows", "local_nnz"), [x.value for x in v]))

    def part(self, k):
        """(rows, row_ptr, cols, values) of the rows device k owns, global indices."""
        sz = self.part_sizes(k)
        rows, ptr = np.zeros(sz["rows"], np.int64), np.zeros(sz["rows"] + 1, np.int64)
        cols, vals = np.zeros(sz["nnz"], np.int32), np.zeros(sz["nnz"], np.float64)
        self.check(self.L.lfgpu_multi_part_download(self.h, k, _p(rows), _p(ptr), _p(cols), _p(vals)))
        return rows, ptr, cols, vals

    def close(self):
        if getattr(self, "h", None):
            self.L.lfgpu_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


class SubMesh:
    """One GPU's share of a partitioned problem: mesh + dof map with local indices, local -> global lists, owned-dof flags."""

    def __init__(self, ctx, handle, dof_owner, rank):
        self.ctx, self.h, self.rank = ctx, handle, rank
        L = ctx.L
        v = [C.c_int64() for _ in range(3)]
        ctx.check(L.lfgpu_submesh_counts(self.h, *[C.byref(x) for x in v]))
        self.n_cells, self.n_nodes, self.n_dofs = [x.value for x in v]
        self.mesh = Mesh(ctx, C.c_void_p(L.lfgpu_submesh_mesh(self.h)), owner=self)
        self.dofmap = DofMap(self.mesh, C.c_void_p(L.lfgpu_submesh_dofmap(self.h)), owner=self)
        self.owned = ctx.empty(self.n_dofs, np.uint8)
        ctx.check(L.lfgpu_submesh_owned_dofs(ctx.h, self.h, dof_owner.ptr, int(rank), self.owned.ptr))

    def _list(self, fn, n):
        out = np.zeros(n, np.int32)
        self.ctx.check(self.ctx.L.lfgpu_memcpy_d2h(self.ctx.h, _p(out), C.c_void_p(fn(self.h)), 4 * n))
        self.ctx.synchronize()
        return out

    def l2g_cells(self):
        return self._list(self.ctx.L.lfgpu_submesh_l2g_cells_device, self.n_cells)

    def l2g_nodes(self):
        return self._list(self.ctx.L.lfgpu_submesh_l2g_nodes_device, self.n_nodes)

    def l2g_dofs(self):
        return self._list(self.ctx.L.lfgpu_submesh_l2g_dofs_device, self.n_dofs)

    def l2g_cells_ptr(self):
        return C.c_void_p(self.ctx.L.lfgpu_submesh_l2g_cells_device(self.h))

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.mesh.h = None
            self.dofmap.h = None
            self.ctx.L.lfgpu_submesh_destroy(self.h)
            self.h = None


class Pattern:
    def __init__(self, mesh, handle, major):
        self.mesh = mesh
        self.ctx = mesh.ctx
        self.h = handle
        self.major = major
        L = self.ctx.L
        self.nnz = L.lfgpu_pattern_nnz(self.h)
        self.rows = L.lfgpu_pattern_rows(self.h)
        self.cols = L.lfgpu_pattern_cols(self.h)

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.L.lfgpu_pattern_destroy(self.h)
            self.h = None

    def restrict_rows(self, keep):
        """Only the outer indices flagged in `keep` (DeviceArray uint8) have to be produced by later numeric passes."""
        self.ctx.check(self.ctx.L.lfgpu_pattern_restrict_rows(self.ctx.h, self.h, keep.ptr))

    def download(self):
        n_outer = self.rows if self.major == ROW_MAJOR else self.cols
        outer = np.zeros(n_outer + 1, np.int32)
        inner = np.zeros(self.nnz, np.int32)
        self.ctx.check(self.ctx.L.lfgpu_pattern_download(self.ctx.h, self.h, _p(outer), _p(inner)))
        return outer, inner

    def assemble_reaction_diffusion(self, degree, alpha, gamma, qr_tria=None, qr_quad=None, active=None, beta=0.0, out=None,
                                    algo=ALGO_AUTO, rows=None):
        """AssembleMatrixLocally(0, dofh, dofh, ReactionDiffusionElementMatrixProvider(fe_space, alpha, gamma[, rules]), M).

        rows: optional DeviceArray(int32) of outer indices to compute (row partition of a multi-GPU run)."""
        if out is None:
            out = self.ctx.zeros(self.nnz)
        self.ctx.check(self.ctx.L.lfgpu_assemble_reaction_diffusion_rows(
            self.ctx.h, self.mesh.h, self.h, degree, _qref(qr_tria), _qref(qr_quad), alpha.ref(), gamma.ref(),
            active.ptr if active is not None else None, beta, out.ptr, algo, rows.ptr if rows is not None else None,
            rows.n if rows is not None else 0))
        return out

    def assemble_edge_mass(self, dofmap, degree, gamma, values, qr_segment=None, active_edges=None):
        """AssembleMatrixLocally(1, dofh, dofh, MassEdgeMatrixProvider(fe_space, gamma, edge_sel), A): ADDS to `values`."""
        self.ctx.check(self.ctx.L.lfgpu_assemble_edge_mass(self.ctx.h, self.mesh.h, dofmap.h, self.h, degree, _qref(qr_segment), gamma.ref(),
                                                           active_edges.ptr if active_edges is not None else None, values.ptr))
        return values

    def assemble_reaction_diffusion_range(self, degree, alpha, gamma, row0, n_rows, qr_tria=None, qr_quad=None, beta=0.0, out=None,
                                          algo=ALGO_AUTO):
        """Only the contiguous outer range [row0, row0 + n_rows) (fan kernel only: LfgpuError UNSUPPORTED otherwise)."""
        if out is None:
            out = self.ctx.zeros(self.nnz)
        self.ctx.check(self.ctx.L.lfgpu_assemble_reaction_diffusion_range(
            self.ctx.h, self.mesh.h, self.h, degree, _qref(qr_tria), _qref(qr_quad), alpha.ref(), gamma.ref(), beta, out.ptr, algo,
            int(row0), int(n_rows)))
        return out

    def assemble_reaction_diffusion_host(self, degree, alpha, gamma, h_node_coords, h_values, out=None, qr_tria=None, qr_quad=None,
                                         algo=ALGO_AUTO, n_blocks=0):
        """Host-buffer form: node coordinates from the numpy array h_node_coords (None = keep), values into the numpy array
        h_values (None = no download); pinned arrays (Context.pinned) make upload, kernel and download overlap.
        Returns the device values; synchronous."""
        if out is None:
            out = self.ctx.empty(self.nnz)
        if h_node_coords is not None:
            assert h_node_coords.dtype == np.float64 and h_node_coords.size == 2 * self.mesh.n_nodes and h_node_coords.flags.c_contiguous
        if h_values is not None:
            assert h_values.dtype == np.float64 and h_values.size == self.nnz and h_values.flags.c_contiguous
        self.ctx.check(self.ctx.L.lfgpu_assemble_reaction_diffusion_host(
            self.ctx.h, self.mesh.h, self.h, degree, _qref(qr_tria), _qref(qr_quad), alpha.ref(), gamma.ref(), _p(h_node_coords), out.ptr,
            _p(h_values), algo, n_blocks))
        return out

    def assemble_reaction_diffusion_host_range(self, degree, alpha, gamma, h_node_coords, h_values_range, row0, n_rows, out=None,
                                               qr_tria=None, qr_quad=None, algo=ALGO_AUTO, n_blocks=0):
        """Host-buffer form for the outer range [row0, row0 + n_rows): uploads only the coordinate window the range refers to
        (h_node_coords is the full array), downloads only the range's values into h_values_range."""
        if out is None:
            out = self.ctx.empty(self.nnz)
        if h_node_coords is not None:
            assert h_node_coords.dtype == np.float64 and h_node_coords.size == 2 * self.mesh.n_nodes and h_node_coords.flags.c_contiguous
        if h_values_range is not None:
            assert h_values_range.dtype == np.float64 and h_values_range.flags.c_contiguous
        self.ctx.check(self.ctx.L.lfgpu_assemble_reaction_diffusion_host_range(
            self.ctx.h, self.mesh.h, self.h, degree, _qref(qr_tria), _qref(qr_quad), alpha.ref(), gamma.ref(), _p(h_node_coords), out.ptr,
            _p(h_values_range), algo, n_blocks, int(row0), int(n_rows)))
        return out

    def spmv(self, values, x, out=None):
        """y = A x on the device."""
        if out is None:
            out = self.ctx.empty(self.rows)
        self.ctx.check(self.ctx.L.lfgpu_spmv(self.ctx.h, self.h, values.ptr, x.ptr, out.ptr))
        return out

    def cg_solve(self, values, rhs, x=None, rel_tol=1e-10, max_iter=10000, jacobi=True):
        """Conjugate gradients on the assembled matrix; returns (x, iterations, relative residual)."""
        if x is None:
            x = self.ctx.zeros(self.rows)
        it, res = C.c_int(0), C.c_double(0.0)
        self.ctx.check(self.ctx.L.lfgpu_cg_solve(self.ctx.h, self.h, values.ptr, rhs.ptr, x.ptr, rel_tol, max_iter, 1 if jacobi else 0,
                                                 C.byref(it), C.byref(res)))
        return x, it.value, res.value

    def fix_flagged_solution_components(self, values, rhs, fixed, fixed_values, compact=False, alt=False):
        """FixFlaggedSolutionComponents(selectvals, A, b) (assemble/fix_dof.h:86-138) on the device;
        alt=True: FixFlaggedSolutionCompAlt (fix_dof.h:181-218), unit rows only.

        values [nnz] / rhs [n] are edited in place; fixed: DeviceArray(uint8) [n]; fixed_values: DeviceArray(float64) [n].
        compact=True additionally returns (outer, inner, values) DeviceArrays + nnz of the matrix with the erased
        entries removed -- makeSparse() of the reference's edited triplet list."""
        fn = self.ctx.L.lfgpu_fix_flagged_solution_comp_alt if alt else self.ctx.L.lfgpu_fix_flagged_solution_components
        if not compact:
            self.ctx.check(fn(self.ctx.h, self.h, values.ptr, rhs.ptr, fixed.ptr, fixed_values.ptr, None, None, None, None))
            return None
        n_outer = self.rows if self.major == ROW_MAJOR else self.cols
        outer = self.ctx.empty(n_outer + 1, np.int32)
        inner = self.ctx.empty(max(self.nnz, 1), np.int32)
        vals = self.ctx.empty(max(self.nnz, 1), np.float64)
        kept = C.c_int64(0)
        self.ctx.check(fn(self.ctx.h, self.h, values.ptr, rhs.ptr, fixed.ptr, fixed_values.ptr, outer.ptr, inner.ptr, vals.ptr,
                          C.byref(kept)))
        return outer, inner, vals, kept.value
