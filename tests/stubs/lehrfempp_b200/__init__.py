"""STUB of the device layer (test infrastructure, never shipped): just enough of lehrfempp_b200's Python API for
tests/test_bench_contract.py to run bench.py's host logic -- argument handling, workload table, timing loop, roofline
arithmetic, JSON contract -- on a machine without a GPU.  It computes NOTHING; numbers in the JSON line are fake."""
import numpy as np

ALGO_AUTO, ALGO_ATOMIC, ALGO_GATHER, ALGO_FAN = 0, 1, 2, 3
COL_MAJOR, ROW_MAJOR = 0, 1


class LfgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class Coeff:
    def __init__(self, kind):
        self.kind = kind

    @staticmethod
    def const(v):
        return Coeff(0)

    @staticmethod
    def per_qp(dev, stride):
        return Coeff(3)


class _Dev:
    def __init__(self, n, dtype=np.float64):
        self.a = np.zeros(min(int(n), 1024), dtype)
        self.ptr = 0

    def to_host(self):
        return self.a


class _Pattern:
    def __init__(self, mesh, degree):
        self.mesh = mesh
        self.h = None
        per_cell = {1: 3.5, 2: 23.0, 3: 76.5}[degree]
        self.nnz = int(per_cell * mesh.n_cells)
        self.calls = 0

    def download(self):
        return np.zeros(8, np.int32), np.zeros(8, np.int32)

    def assemble_reaction_diffusion(self, degree, alpha, gamma, out=None, algo=ALGO_AUTO, **kw):
        self.mesh.ctx.kernel_launches += 1
        return out

    def spmv(self, values, x, out=None):
        return _Dev(self.mesh.n_cells)

    def assemble_reaction_diffusion_host(self, degree, alpha, gamma, h_xy, h_vals, out=None, algo=ALGO_AUTO, n_blocks=16, **kw):
        self.mesh.ctx.kernel_launches += 1


class _DofMap:
    def __init__(self, mesh, degree):
        self.mesh, self.degree = mesh, degree
        self.stride = {1: 4, 2: 9, 3: 16}[degree] if mesh.n_quad else {1: 3, 2: 6, 3: 10}[degree]
        self.num_dofs = int({1: 0.5, 2: 2.0, 3: 4.5}[degree] * mesh.n_cells)

    def symbolic(self, major=ROW_MAJOR):
        return _Pattern(self.mesh, self.degree)


class _Mesh:
    def __init__(self, ctx, n_tria, n_quad, n_nodes):
        self.ctx, self.n_tria, self.n_quad, self.n_nodes = ctx, n_tria, n_quad, n_nodes
        self.n_cells = n_tria + n_quad
        self.n_edges = 0

    def refine_regular(self):
        return _Mesh(self.ctx, 4 * self.n_tria, 4 * self.n_quad, 4 * self.n_nodes)

    def dofmap_lagrange(self, degree):
        return _DofMap(self, degree)

    def qp_coords(self, degree, stride):
        m = self

        class _Q:
            def to_host(self):
                return np.zeros(m.n_cells * stride * 2)
        return _Q()

    def download(self):
        return {"node_coords": np.zeros((self.n_nodes, 2))}

    def update_node_coords(self, xy):
        pass


class Context:
    def __init__(self, device=0):
        self.h = None
        self.L = None
        self.kernel_launches = 0
        self._t = 0.0

    def mesh_tp_tria(self, nx, ny, *a):
        n = min(nx, 64)  # the stub keeps arrays small whatever the workload asks for
        return _Mesh(self, 2 * n * n, 0, (n + 1) * (n + 1))

    def mesh_upload(self, node_coords, cell_nodes, cell_coords=None):
        return _Mesh(self, int(cell_nodes.shape[0]), 0, int(node_coords.shape[0]))

    def mesh_hybrid(self, n, jitter, seed):
        n = min(n, 64)
        return _Mesh(self, n * n, n * n // 2, (n + 1) * (n + 1))

    def synchronize(self):
        pass

    def event(self):
        return object()

    def record(self, ev):
        self._t += 1.0

    def elapsed_ms(self, a, b):
        return 1.0

    def pinned(self, n, dtype=np.float64):
        return np.zeros(int(n), dtype)

    def empty(self, n, dtype=np.float64):
        return _Dev(n, dtype)

    def to_device(self, arr):
        return _Dev(arr.size)

    def check(self, rc):
        return rc

    def close(self):
        pass
