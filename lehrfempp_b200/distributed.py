"""Multi-GPU assembly: Morton partition of the cells, owner-adds exchange of interface rows (host plumbing).

The reference is serial; this module is the one place where the path crosses devices:

  * cells are sorted by the Morton code of their centroid and cut into `world` contiguous ranges (one per GPU);
  * a matrix row is OWNED by the lowest rank that has a cell touching it; rows touched by one rank only are
    "interior", the others are "interface" rows;
  * every rank assembles the contributions of ITS cells (interface rows first), sends the partial sums of the interface
    rows it does not own to their owners with ONE all-to-all-v over NCCL / NVLink, assembles its interior rows while
    the messages fly, and finally adds what it received.  Results stay distributed: a rank holds the final values of
    the rows it owns, in the positions of the GLOBAL reference pattern (bit-exact numbering is never changed).

The list construction is plain torch and device agnostic, so the same code runs on CPU tensors with the gloo backend in
the unit tests; the kernels (assembly, pack, unpack-add) are liblfgpu's.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def _part1by1(v):
    v = v & 0xFFFF
    v = (v | (v << 8)) & 0x00FF00FF
    v = (v | (v << 4)) & 0x0F0F0F0F
    v = (v | (v << 2)) & 0x33333333
    v = (v | (v << 1)) & 0x55555555
    return v


def morton_partition(node_xy, cell_nodes, world):
    """cell -> rank (uint8 tensor): cells sorted by the Morton code of their centroid, equal-count contiguous ranges.

    node_xy: (n_nodes, 2) float64 tensor; cell_nodes: (n_cells, 4) int64 tensor with -1 / 0xFFFFFFFF in slot 3 of triangles."""
    n_cells = cell_nodes.shape[0]
    tri = (cell_nodes[:, 3] < 0) | (cell_nodes[:, 3] >= node_xy.shape[0])
    idx = cell_nodes.clone()
    idx[tri, 3] = idx[tri, 0]
    pts = node_xy[idx.reshape(-1)].reshape(n_cells, 4, 2)
    cnt = torch.where(tri, 3.0, 4.0).to(node_xy.dtype)
    s = pts[:, :3].sum(dim=1) + torch.where(tri[:, None], torch.zeros_like(pts[:, 3]), pts[:, 3])
    cen = s / cnt[:, None]
    lo = node_xy.min(dim=0).values
    hi = node_xy.max(dim=0).values
    q = ((cen - lo) / (hi - lo).clamp_min(1e-300) * 65535.0).clamp(0, 65535).to(torch.int64)
    code = _part1by1(q[:, 0]) | (_part1by1(q[:, 1]) << 1)
    order = torch.argsort(code, stable=True)
    part = torch.empty(n_cells, dtype=torch.uint8, device=node_xy.device)
    per = -(-n_cells // world)
    part[order] = (torch.arange(n_cells, device=node_xy.device) // per).to(torch.uint8)
    return part


def row_ranges(outer, world):
    """Contiguous row blocks with (nearly) equal numbers of stored values: boundaries [world + 1] as a python list.
    Block k = rows [b[k], b[k+1]).  `outer` is the row-pointer array of the pattern (any device)."""
    n_rows = outer.numel() - 1
    nnz = int(outer[-1].item())
    targets = torch.tensor([nnz * k // world for k in range(1, world)], dtype=outer.dtype, device=outer.device)
    cuts = torch.searchsorted(outer[:-1].contiguous(), targets, right=False).tolist() if world > 1 else []
    bounds = [0] + [min(max(int(c), 0), n_rows) for c in cuts] + [n_rows]
    for k in range(1, len(bounds)):  # monotone even for degenerate inputs
        bounds[k] = max(bounds[k], bounds[k - 1])
    return bounds


class PartitionPlan:
    """Row classification and message layout of one rank (device-agnostic torch tensors)."""

    def __init__(self, cell_part, adj_ptr, adj_cell, outer, rank, world):
        dev = adj_ptr.device
        n_rows = adj_ptr.numel() - 1
        self.rank, self.world, self.n_rows = rank, world, n_rows
        counts = (adj_ptr[1:] - adj_ptr[:-1]).to(torch.int64)
        row_of_item = torch.repeat_interleave(torch.arange(n_rows, device=dev), counts)
        item_part = cell_part[adj_cell.to(torch.int64)].to(torch.int64)
        mask = torch.zeros(n_rows, dtype=torch.int64, device=dev)
        owner = torch.full((n_rows,), -1, dtype=torch.int64, device=dev)
        touched = []
        for s in range(world):
            t = torch.zeros(n_rows, dtype=torch.bool, device=dev)
            t[row_of_item[item_part == s]] = True
            touched.append(t)
            mask |= t.to(torch.int64) << s
        for s in reversed(range(world)):
            owner = torch.where(touched[s], torch.full_like(owner, s), owner)
        del row_of_item, item_part
        mine = touched[rank]
        only_me = mask == (1 << rank)
        self.owner = owner
        self.interior_rows = torch.nonzero(only_me).flatten().to(torch.int32)
        self.iface_rows = torch.nonzero(mine & ~only_me).flatten().to(torch.int32)
        self.owned_rows = torch.nonzero(owner == rank).flatten().to(torch.int32)
        self.active = (cell_part == rank).to(torch.uint8)
        lens = (outer[1:] - outer[:-1]).to(torch.int64)
        outer64 = outer.to(torch.int64)

        def layout(lists):
            rows = torch.cat(lists) if lists else torch.zeros(0, dtype=torch.int64, device=dev)
            l = lens[rows]
            off = torch.cumsum(l, 0) - l
            splits = [int(lens[x].sum().item()) for x in lists]
            # flat positions inside the value array of every buffer element (torch fallback of pack / unpack-add)
            if rows.numel() > 0:
                idx = torch.repeat_interleave(outer64[rows] - off, l) + torch.arange(int(l.sum().item()), device=dev)
            else:
                idx = torch.zeros(0, dtype=torch.int64, device=dev)
            return rows.to(torch.int32), off, splits, idx

        send_lists = [torch.nonzero(mine & (owner == s)).flatten() if s != rank else torch.zeros(0, dtype=torch.int64, device=dev)
                      for s in range(world)]
        recv_lists = [torch.nonzero((owner == rank) & touched[s]).flatten() if s != rank else torch.zeros(0, dtype=torch.int64, device=dev)
                      for s in range(world)]
        self.send_rows, self.send_off, self.send_splits, self.send_idx = layout(send_lists)
        self.recv_rows, self.recv_off, self.recv_splits, self.recv_idx = layout(recv_lists)
        self.n_send, self.n_recv = sum(self.send_splits), sum(self.recv_splits)

    def exchange_torch(self, values, group=None):
        """Reference implementation of the exchange with torch ops only (used by the CPU / gloo tests)."""
        send = values[self.send_idx].contiguous()
        recv = torch.empty(self.n_recv, dtype=values.dtype, device=values.device)
        dist.all_to_all_single(recv, send, self.recv_splits, self.send_splits, group=group)
        values.index_add_(0, self.recv_idx, recv)
        return values


class _CudaView:
    """Zero-copy torch view of a raw device pointer owned by liblfgpu."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_view(ptr, n, dtype, device):
    typestr = {torch.int32: "<i4", torch.uint32: "<u4", torch.float64: "<f8", torch.uint8: "|u1", torch.int64: "<i8"}[dtype]
    if dtype == torch.uint32:  # torch has no uint32 arithmetic: reinterpret as int32 (values < 2^31 here)
        typestr, dtype = "<i4", torch.int32
    return torch.as_tensor(_CudaView(ptr, n, typestr), device=device)


class DistributedAssembler:
    """Partitioned AssembleMatrixLocally over the GPUs of one node (one process per GPU, torch.distributed / NCCL)."""

    def __init__(self, ctx, mesh, pattern, degree, group=None, mode="exchange"):
        """mode = "exchange": every rank assembles the contributions of its cells, interface rows are summed at their
        owner (one all-to-all-v per step).  mode = "owner": the owner of a row assembles it completely, re-computing the
        few halo cells of its neighbours (the mesh is replicated at setup) -- no data-path collective at all."""
        import lehrfempp_b200 as lf
        assert mode in ("exchange", "owner", "owner_rows")
        self.mode = mode
        if mode == "owner_rows":
            # owner-computes over contiguous ROW BLOCKS with equal numbers of stored values: no Morton plan, no lists --
            # the rank's share is one range launch of the fastest kernel (the fan kernel with its L2 prefetch needs
            # contiguous rows); halo cells are re-computed from the replicated geometry as in "owner"
            self.graph = None
            self.lf, self.ctx, self.mesh, self.pattern, self.degree, self.group = lf, ctx, mesh, pattern, degree, group
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
            dev = torch.device("cuda", torch.cuda.current_device())
            self.dev = dev
            ctx.synchronize()
            n_rows = pattern.rows if pattern.major == lf.ROW_MAJOR else pattern.cols
            outer = device_view(ctx.L.lfgpu_pattern_outer_device(pattern.h), n_rows + 1, torch.int32, dev)
            self.bounds = row_ranges(outer, self.world)
            self.row0, self.n_rows = self.bounds[self.rank], self.bounds[self.rank + 1] - self.bounds[self.rank]

            class _P:  # the part of PartitionPlan the callers look at
                pass
            self.plan = _P()
            self.plan.owned_rows = torch.arange(self.row0, self.row0 + self.n_rows, dtype=torch.int32, device=dev)
            self.plan.interior_rows = self.plan.iface_rows = torch.zeros(0, dtype=torch.int32, device=dev)
            self.plan.n_send = self.plan.n_recv = 0
            self._range_ok = True
            self.main = torch.cuda.ExternalStream(ctx.stream, device=dev)
            return
        self.graph = None
        self.lf, self.ctx, self.mesh, self.pattern, self.degree, self.group = lf, ctx, mesh, pattern, degree, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        L = ctx.L
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        ctx.synchronize()
        n_rows = pattern.rows if pattern.major == lf.ROW_MAJOR else pattern.cols
        n_items = L.lfgpu_pattern_num_items(pattern.h)
        adj_ptr = device_view(L.lfgpu_pattern_adj_ptr_device(pattern.h), n_rows + 1, torch.int32, dev)
        adj = device_view(L.lfgpu_pattern_adj_device(pattern.h), n_items, torch.uint32, dev)
        outer = device_view(L.lfgpu_pattern_outer_device(pattern.h), n_rows + 1, torch.int32, dev)
        xy = device_view(L.lfgpu_mesh_node_coords_device(mesh.h), 2 * mesh.n_nodes, torch.float64, dev).reshape(-1, 2)
        cn = device_view(L.lfgpu_mesh_cell_nodes_device(mesh.h), 4 * mesh.n_cells, torch.uint32, dev).reshape(-1, 4).to(torch.int64)
        self.cell_part = morton_partition(xy, cn, self.world)
        del cn
        adj_cell = (adj.to(torch.int64) & 0xFFFFFFFF) >> 4
        self.plan = PartitionPlan(self.cell_part, adj_ptr, adj_cell, outer, self.rank, self.world)
        del adj_cell
        p = self.plan
        self.send_buf = torch.zeros(max(p.n_send, 1), dtype=torch.float64, device=dev)
        self.recv_buf = torch.zeros(max(p.n_recv, 1), dtype=torch.float64, device=dev)
        self.main = torch.cuda.ExternalStream(ctx.stream, device=dev)
        self.comm = torch.cuda.Stream(device=dev, priority=-1)  # high priority: the collective must not queue behind the interior kernel
        self.ev_packed = torch.cuda.Event()
        self.ev_received = torch.cuda.Event()
        torch.cuda.synchronize()

    class _Rows:
        def __init__(self, t):
            self.ptr, self.n, self._keep = C.c_void_p(t.data_ptr()), t.numel(), t

    class _Mask:
        def __init__(self, t):
            self.ptr, self._keep = C.c_void_p(t.data_ptr()), t

    def assemble(self, alpha, gamma, values, qr_tria=None, qr_quad=None):
        """One partitioned numeric pass; afterwards this rank's OWNED rows of `values` are final."""
        lf, ctx, p, L = self.lf, self.ctx, self.plan, self.ctx.L
        pat = self.pattern
        if self.mode == "owner_rows":
            if self._range_ok:
                try:
                    pat.assemble_reaction_diffusion_range(self.degree, alpha, gamma, self.row0, self.n_rows, qr_tria, qr_quad, out=values)
                    return values
                except lf.LfgpuError as e:
                    if e.code != -7:  # LFGPU_ERR_UNSUPPORTED: not a fan-kernel call -> generic kernel over the same rows as a list
                        raise
                    self._range_ok = False
            pat.assemble_reaction_diffusion(self.degree, alpha, gamma, qr_tria, qr_quad, out=values, algo=lf.ALGO_AUTO,
                                            rows=self._Rows(p.owned_rows))
            return values
        if self.mode == "owner" and self.world > 1:
            # owner-computes: one launch over the rows I own, all adjacent cells (mine or halo) contribute
            pat.assemble_reaction_diffusion(self.degree, alpha, gamma, qr_tria, qr_quad, out=values, algo=lf.ALGO_AUTO,
                                            rows=self._Rows(p.owned_rows))
            return values
        if self.world > 1 and p.iface_rows.numel() > 0:
            # 1. interface rows: only my cells contribute (activity mask), generic owner-computes kernel
            pat.assemble_reaction_diffusion(self.degree, alpha, gamma, qr_tria, qr_quad, active=self._Mask(p.active), out=values,
                                            algo=lf.ALGO_GATHER, rows=self._Rows(p.iface_rows))
            # 2. pack the partial sums I do not own
            if p.n_send > 0:
                ctx.check(L.lfgpu_rows_pack(ctx.h, pat.h, p.send_rows.data_ptr(), p.send_rows.numel(), p.send_off.data_ptr(),
                                            values.ptr, self.send_buf.data_ptr()))
        if self.world > 1:
            # 3. one all-to-all-v on a side stream, overlapped with the interior rows
            self.ev_packed.record(self.main)
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(self.ev_packed)
                dist.all_to_all_single(self.recv_buf[: p.n_recv], self.send_buf[: p.n_send], p.recv_splits, p.send_splits, group=self.group)
                self.ev_received.record(self.comm)
        # 4. interior rows: every adjacent cell is mine -> the fastest kernel that applies, no mask
        if p.interior_rows.numel() > 0:
            pat.assemble_reaction_diffusion(self.degree, alpha, gamma, qr_tria, qr_quad, out=values, algo=lf.ALGO_AUTO,
                                            rows=self._Rows(p.interior_rows))
        if self.world > 1:
            # 5. owner adds what the other ranks computed for its interface rows
            self.main.wait_event(self.ev_received)
            if p.n_recv > 0:
                ctx.check(L.lfgpu_rows_unpack_add(ctx.h, pat.h, p.recv_rows.data_ptr(), p.recv_rows.numel(), p.recv_off.data_ptr(),
                                                  self.recv_buf.data_ptr(), values.ptr))
        return values

    # ---- CUDA graph of one step: the partitioned step is a handful of short launches plus a collective, so at 4-8 GPUs
    # the host launch path, not the GPUs, would set the pace.  Capturing the step once makes it one launch.
    def capture(self, alpha, gamma, values, qr_tria=None, qr_quad=None):
        for _ in range(2):  # warm-up outside the capture: plan building, table upload, NCCL channel setup
            self.assemble(alpha, gamma, values, qr_tria, qr_quad)
        self.ctx.synchronize()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.main, capture_error_mode="thread_local"):
            self.assemble(alpha, gamma, values, qr_tria, qr_quad)
        self.graph = g
        self._graph_keep = (alpha, gamma, values, qr_tria, qr_quad)
        return g

    def replay(self):
        with torch.cuda.stream(self.main):
            self.graph.replay()

    def owned_value_mask(self, outer_host):
        """bool numpy mask over the value array: entries of rows this rank owns (for checks / gathers)."""
        owned = self.plan.owned_rows.cpu().numpy()
        m = np.zeros(outer_host[-1], dtype=bool)
        for r in owned:
            m[outer_host[r]:outer_host[r + 1]] = True
        return m


class OwnedAssembler:
    """Distributed ownership (BASELINE.json north star: "Morton-ordered cell ranges, each GPU owning the matrix rows for its
    cells"): this rank keeps ONLY its sub-problem -- its cells plus the one-cell halo around the dofs it owns, with local indices
    (lehrfempp_b200/csrc/partition.cu) -- and runs symbolic pass, plans and numeric kernels on it unchanged.  Owner-computes: the
    rows a rank owns are complete without any exchange (halo cells are recomputed), so the data path has no collective; setup has
    none either, because the partition is a deterministic function of the replicated mesh.  Memory and symbolic time per GPU fall
    like 1/N, and int32 indices only have to address 1/N of the matrix (config 4: 2.5e9 stored values globally).

    After construction the caller may drop the global mesh / dof map: nothing here refers to them."""

    def __init__(self, ctx, mesh, dofmap, degree, rank, world, major=None, halo=True):
        import time

        import lehrfempp_b200 as lf
        self.lf, self.ctx, self.degree, self.rank, self.world = lf, ctx, degree, rank, world
        self.mode = "owned"
        self.global_cells, self.global_dofs = mesh.n_cells, dofmap.num_dofs
        t0 = time.time()
        part, owner = dofmap.partition_morton(world)
        self.sub = dofmap.submesh(part, owner, rank, halo=halo)
        del part, owner
        ctx.synchronize()
        self.partition_s = time.time() - t0
        t0 = time.time()
        self.pattern = self.sub.dofmap.symbolic(major=lf.ROW_MAJOR if major is None else major)
        self.pattern.restrict_rows(self.sub.owned)  # halo rows are incomplete anyway: no generic pass for those outside the fast plans
        ctx.synchronize()
        self.symbolic_s = time.time() - t0
        own = self.sub.owned.to_host().astype(bool)
        self.n_owned_rows = int(own.sum())
        outer = self.pattern.download()[0]
        self.owned_nnz = int((np.diff(outer)[own]).sum())
        self._own = own
        self.mesh = self.sub.mesh
        self.graph = None

    def assemble(self, alpha, gamma, values=None, qr_tria=None, qr_quad=None):
        """One numeric pass over the sub-problem; the rows flagged in `self.sub.owned` are final."""
        return self.pattern.assemble_reaction_diffusion(self.degree, alpha, gamma, qr_tria, qr_quad, out=values)

    def owned_rows_global(self):
        """(global row ids, local row ids) of the rows this rank owns, ascending."""
        l2g = self.sub.l2g_dofs()
        rows_l = np.nonzero(self._own)[0]
        return l2g[rows_l], rows_l

    def witness(self, values):
        """(max |A 1| over the owned rows, sum of A 1 over the owned rows): size-independent checks of the distributed matrix."""
        y = self.pattern.spmv(values, self.ctx.to_device(np.ones(self.pattern.cols))).to_host()[self._own]
        return (float(np.abs(y).max()) if y.size else 0.0), float(y.sum())
