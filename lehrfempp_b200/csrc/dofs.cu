// Dof maps on the device (product code).
//
// Stands in for lf::assemble::DofHandler / UniformFEDofHandler (lib/lf/assemble/dofhandler.h:112-228,260-503,
// dofhandler.cc:86-338).  Numbering rule restated from dofhandler.cc:141-284:
//   dofs of nodes first (node index order, n_pt each), then edge-interior dofs (edge index order, n_seg each), then
//   cell-interior dofs (cell index order, n_tria or n_quad depending on the cell type);
//   a cell lists: vertex dofs in local vertex order | for each local edge its interior dofs, REVERSED when the edge's
//   relative orientation is negative (:245-260) | its own interior dofs.  Table stride = max(tria, quad) (:138).
// One thread per cell writes its row of the table; the only serial dependency of the reference loop (the running
// interior-dof counter) becomes an exclusive prefix sum over the cells.
#include <cub/cub.cuh>

#include "lfgpu_internal.cuh"

namespace lfgpu {
int ensure_topology(lfgpu_ctx* ctx, lfgpu_mesh* m);

namespace {
constexpr int kThreads = 256;

__global__ void k_interior_counts(int64_t n_cells, const uint32_t* __restrict__ cell_nodes, int n_tria, int n_quad,
                                  int64_t* __restrict__ counts) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  counts[c] = (cell_nodes[4 * c + 3] == LFGPU_IDX_NIL) ? n_tria : n_quad;
}

__global__ void k_uniform_dofs(int64_t n_cells, const uint32_t* __restrict__ cell_nodes, const uint32_t* __restrict__ cell_edges,
                               const int8_t* __restrict__ cell_edge_ori, const int64_t* __restrict__ interior_offset,
                               int n_pt, int n_seg, int n_tria, int n_quad, int64_t edge_dof_base, int64_t cell_dof_base,
                               int stride, int32_t* __restrict__ cell_dofs, uint8_t* __restrict__ n_ldof) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[c];
  const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
  const int nv = (v.w == LFGPU_IDX_NIL) ? 3 : 4;
  int32_t* out = cell_dofs + c * stride;
  int k = 0;
  for (int l = 0; l < nv; ++l) {
    for (int j = 0; j < n_pt; ++j) out[k++] = static_cast<int32_t>(static_cast<int64_t>(vv[l]) * n_pt + j);
  }
  if (n_seg > 0) {
    for (int l = 0; l < nv; ++l) {
      const int64_t base = edge_dof_base + static_cast<int64_t>(cell_edges[4 * c + l]) * n_seg;
      if (cell_edge_ori[4 * c + l] > 0) {
        for (int j = 0; j < n_seg; ++j) out[k++] = static_cast<int32_t>(base + j);
      } else {
        for (int j = n_seg - 1; j >= 0; --j) out[k++] = static_cast<int32_t>(base + j);
      }
    }
  }
  const int n_int = (nv == 3) ? n_tria : n_quad;
  const int64_t ibase = cell_dof_base + (interior_offset ? interior_offset[c] : c * static_cast<int64_t>(n_int));
  for (int j = 0; j < n_int; ++j) out[k++] = static_cast<int32_t>(ibase + j);
  n_ldof[c] = static_cast<uint8_t>(k);
  for (; k < stride; ++k) out[k] = -1;
}

__global__ void k_convert_dofs(int64_t n, int stride, const int64_t* __restrict__ in, const uint8_t* __restrict__ nl_in,
                               int64_t n_dofs, int32_t* __restrict__ out, uint8_t* __restrict__ nl_out, int* __restrict__ flags) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n) return;
  int cnt = 0;
  const int lim = nl_in ? nl_in[c] : stride;
  for (int k = 0; k < stride; ++k) {
    const int64_t d = in[c * stride + k];
    const bool used = (k < lim) && d >= 0;
    if (used) {
      if (d >= n_dofs) flags[0] = 1;
      if (cnt != k) flags[1] = 1;  // holes in the list are not allowed
      ++cnt;
    }
    out[c * stride + k] = used ? static_cast<int32_t>(d) : -1;
  }
  nl_out[c] = static_cast<uint8_t>(cnt);
}

// ---- DynamicFEDofHandler (dofhandler.h:514-789): per-entity numbers of interior dofs ---------------------------------
// counts widened to int64 with a trailing 0, so that an exclusive scan over n + 1 entries also yields the total
__global__ void k_widen_counts(int64_t n, const uint32_t* __restrict__ in, int64_t* __restrict__ out, int* __restrict__ flags) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i > n) return;
  const uint32_t v = (i < n && in != nullptr) ? in[i] : 0U;
  if (v > static_cast<uint32_t>(kMaxNsf)) flags[0] = 1;
  out[i] = v;
}

// length of every cell's dof list and its maximum (= row length of the table)
__global__ void k_dynamic_lengths(int64_t n_cells, const uint32_t* __restrict__ cell_nodes, const uint32_t* __restrict__ cell_edges,
                                  const int64_t* __restrict__ node_off, const int64_t* __restrict__ edge_off,
                                  const int64_t* __restrict__ cell_off, int* __restrict__ max_len) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int len = 0;
  if (c < n_cells) {
    const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[c];
    const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
    const int nv = (v.w == LFGPU_IDX_NIL) ? 3 : 4;
    for (int l = 0; l < nv; ++l) {
      len += static_cast<int>(node_off[vv[l] + 1] - node_off[vv[l]]);
      if (edge_off != nullptr) {
        const uint32_t e = cell_edges[4 * c + l];
        len += static_cast<int>(edge_off[e + 1] - edge_off[e]);
      }
    }
    len += static_cast<int>(cell_off[c + 1] - cell_off[c]);
  }
  len = __reduce_max_sync(0xffffffffU, len);
  if ((threadIdx.x & 31) == 0) atomicMax(max_len, len);
}

// one thread per cell writes its list: vertex dofs | edge-interior dofs per local edge, reversed for a negative relative
// orientation (dofhandler.h:669-688) | own interior dofs
__global__ void k_dynamic_dofs(int64_t n_cells, const uint32_t* __restrict__ cell_nodes, const uint32_t* __restrict__ cell_edges,
                               const int8_t* __restrict__ cell_edge_ori, const int64_t* __restrict__ node_off,
                               const int64_t* __restrict__ edge_off, const int64_t* __restrict__ cell_off, int64_t edge_dof_base,
                               int64_t cell_dof_base, int stride, int32_t* __restrict__ cell_dofs, uint8_t* __restrict__ n_ldof) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[c];
  const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
  const int nv = (v.w == LFGPU_IDX_NIL) ? 3 : 4;
  int32_t* out = cell_dofs + c * stride;
  int k = 0;
  for (int l = 0; l < nv; ++l) {
    const int64_t b = node_off[vv[l]], e = node_off[vv[l] + 1];
    for (int64_t d = b; d < e && k < stride; ++d) out[k++] = static_cast<int32_t>(d);
  }
  if (edge_off != nullptr) {
    for (int l = 0; l < nv; ++l) {
      const uint32_t ed = cell_edges[4 * c + l];
      const int64_t b = edge_dof_base + edge_off[ed], e = edge_dof_base + edge_off[ed + 1];
      if (cell_edge_ori[4 * c + l] > 0) {
        for (int64_t d = b; d < e && k < stride; ++d) out[k++] = static_cast<int32_t>(d);
      } else {
        for (int64_t d = e - 1; d >= b && k < stride; --d) out[k++] = static_cast<int32_t>(d);
      }
    }
  }
  {
    const int64_t b = cell_dof_base + cell_off[c], e = cell_dof_base + cell_off[c + 1];
    for (int64_t d = b; d < e && k < stride; ++d) out[k++] = static_cast<int32_t>(d);
  }
  n_ldof[c] = static_cast<uint8_t>(k);
  for (; k < stride; ++k) out[k] = -1;
}

// ---- gather plan of the load vector ------------------------------------------------------------------------------------
__global__ void k_plan_items(int64_t n_cells, int stride, const int32_t* __restrict__ dofs, const uint8_t* __restrict__ nldof,
                             int32_t invalid_key, int32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n_cells * stride) return;
  const int64_t c = t / stride;
  const int a = static_cast<int>(t - c * stride);
  keys[t] = (a < nldof[c]) ? dofs[t] : invalid_key;  // unused slots sort behind every dof
  vals[t] = (static_cast<uint32_t>(c) << 4) | static_cast<uint32_t>(a);
}

// ptr[r] = number of sorted keys < r, r = 0 .. n_dofs
__global__ void k_plan_ptr(int64_t n_dofs, int64_t n_keys, const int32_t* __restrict__ keys, int32_t* __restrict__ ptr) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r > n_dofs) return;
  int64_t lo = 0, hi = n_keys;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < r) lo = mid + 1; else hi = mid;
  }
  ptr[r] = static_cast<int32_t>(lo);
}

// ---- positions of the dofs of a Lagrange layout (interpolation nodes: lagr_fe.h EvaluationNodes of O1 / O2 / O3) ----------
__global__ void k_dof_xy_nodes(int64_t n_nodes, int n_pt, const double* __restrict__ node_coords, double* __restrict__ out) {
  const int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= n_nodes) return;
  for (int j = 0; j < n_pt; ++j) {
    out[2 * (v * n_pt + j)] = node_coords[2 * v];
    out[2 * (v * n_pt + j) + 1] = node_coords[2 * v + 1];
  }
}
// interior dof j of an edge sits at t = (j + 1) / (n_seg + 1) along the edge's own direction (FeLagrangeO{2,3}Segment)
__global__ void k_dof_xy_edges(int64_t n_edges, int n_seg, int64_t base, const uint32_t* __restrict__ edge_nodes,
                               const double* __restrict__ node_coords, double* __restrict__ out) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const uint32_t a = edge_nodes[2 * e], b = edge_nodes[2 * e + 1];
  const double ax = node_coords[2 * a], ay = node_coords[2 * a + 1], bx = node_coords[2 * b], by = node_coords[2 * b + 1];
  for (int j = 0; j < n_seg; ++j) {
    const double t = static_cast<double>(j + 1) / static_cast<double>(n_seg + 1);
    out[2 * (base + e * n_seg + j)] = ax * (1.0 - t) + bx * t;  // SegmentO1::Global (geometry/segment_o1.cc:9-11)
    out[2 * (base + e * n_seg + j) + 1] = ay * (1.0 - t) + by * t;
  }
}
// cell-interior dofs: triangle centroid (O3), quadrilateral centre (O2) or the 2 x 2 interior lattice (O3) -- read off the table
__global__ void k_dof_xy_cells(int64_t n_cells, int stride, int n_pt, int n_seg, int n_tria, int n_quad,
                               const uint32_t* __restrict__ cell_nodes, const double* __restrict__ node_coords,
                               const double* __restrict__ cell_coords, const int32_t* __restrict__ cell_dofs, double* __restrict__ out) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[c];
  const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
  const int nv = (v.w == LFGPU_IDX_NIL) ? 3 : 4;
  double x[4], y[4];
  for (int k = 0; k < nv; ++k) {
    x[k] = cell_coords ? cell_coords[8 * c + 2 * k] : node_coords[2 * vv[k]];
    y[k] = cell_coords ? cell_coords[8 * c + 2 * k + 1] : node_coords[2 * vv[k] + 1];
  }
  const int n_int = nv == 3 ? n_tria : n_quad;
  const int first = nv * n_pt + nv * n_seg;
  for (int j = 0; j < n_int; ++j) {
    double X, Y;
    if (nv == 3) {  // TriaO1::Global at (1/3, 1/3) (tria_o1.cc:70-74)
      const double t = 1.0 / 3.0, l0 = 1.0 - t - t;
      X = x[0] * l0 + x[1] * t + x[2] * t;
      Y = y[0] * l0 + y[1] * t + y[2] * t;
    } else {  // QuadO1::Global (quad_o1.cc:68-83)
      const double third = 1.0 / 3.0;
      const double rx[4] = {third, 2 * third, 2 * third, third}, ry[4] = {third, third, 2 * third, 2 * third};
      const double x0 = n_int == 1 ? 0.5 : rx[j], x1 = n_int == 1 ? 0.5 : ry[j];
      const double a = (1.0 - x0) * (1.0 - x1), b = x0 * (1.0 - x1), cc = x0 * x1, d = (1.0 - x0) * x1;
      X = x[0] * a + x[1] * b + x[2] * cc + x[3] * d;
      Y = y[0] * a + y[1] * b + y[2] * cc + y[3] * d;
    }
    const int32_t dof = cell_dofs[c * stride + first + j];
    out[2 * static_cast<int64_t>(dof)] = X;
    out[2 * static_cast<int64_t>(dof) + 1] = Y;
  }
}

// dofs of the selected edges: their interior dofs and the dofs of their end points (fe/fe_tools.h:320-353 visits the
// selected edges and flags GlobalDofIndices(edge))
__global__ void k_edge_dof_flags(int64_t n_edges, int n_pt, int n_seg, int64_t edge_base, const uint32_t* __restrict__ edge_nodes,
                                 const uint8_t* __restrict__ edge_sel, uint8_t* __restrict__ flags) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges || edge_sel[e] == 0) return;
  for (int k = 0; k < 2; ++k)
    for (int j = 0; j < n_pt; ++j) flags[static_cast<int64_t>(edge_nodes[2 * e + k]) * n_pt + j] = 1;  // benign: every writer stores 1
  for (int j = 0; j < n_seg; ++j) flags[edge_base + e * n_seg + j] = 1;
}

__global__ void k_dofs_to_i64(int64_t n, const int32_t* __restrict__ in, int64_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

}  // namespace
}  // namespace lfgpu

// Items (dof, cell << 4 | a) sorted stably by dof: the slots are generated in (cell, a) order, so the items of a dof end
// up in ascending cell order -- the order of the reference's cell loop.
int lfgpu::dofmap_gather_plan(lfgpu_ctx* ctx, const lfgpu_dofmap* dc) {
  lfgpu_dofmap* d = const_cast<lfgpu_dofmap*>(dc);
  if (d->g_state == 1) return LFGPU_OK;
  if (d->n_cells >= (1LL << 28)) LFGPU_FAIL(ctx, LFGPU_ERR_OVERFLOW, "gather plan: more than 2^28 cells");
  cudaStream_t st = ctx->stream;
  const int64_t n_slots = d->n_cells * d->stride;
  int32_t *keys_in = nullptr, *keys_out = nullptr;
  uint32_t *vals_in = nullptr, *vals_out = nullptr;
  int32_t* ptr = nullptr;
  void* tmp = nullptr;
  cudaError_t e = cudaMalloc(&keys_in, sizeof(int32_t) * n_slots);
  if (e == cudaSuccess) e = cudaMalloc(&keys_out, sizeof(int32_t) * n_slots);
  if (e == cudaSuccess) e = cudaMalloc(&vals_in, sizeof(uint32_t) * n_slots);
  if (e == cudaSuccess) e = cudaMalloc(&vals_out, sizeof(uint32_t) * n_slots);
  if (e == cudaSuccess) e = cudaMalloc(&ptr, sizeof(int32_t) * (d->n_dofs + 1));
  if (e == cudaSuccess) {
    k_plan_items<<<static_cast<unsigned>(cdiv(n_slots, kThreads)), kThreads, 0, st>>>(d->n_cells, d->stride, d->cell_dofs, d->n_ldof,
                                                                                     static_cast<int32_t>(d->n_dofs), keys_in, vals_in);
    ctx->launches++;
    int key_bits = 1;
    while ((1LL << key_bits) <= d->n_dofs) ++key_bits;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, keys_in, keys_out, vals_in, vals_out, n_slots, 0, key_bits, st);
    e = cudaMalloc(&tmp, tb);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tb, keys_in, keys_out, vals_in, vals_out, n_slots, 0, key_bits, st);
  }
  int32_t n_items = 0;
  if (e == cudaSuccess) {
    k_plan_ptr<<<static_cast<unsigned>(cdiv(d->n_dofs + 1, kThreads)), kThreads, 0, st>>>(d->n_dofs, n_slots, keys_out, ptr);
    ctx->launches++;
    e = cudaMemcpyAsync(&n_items, ptr + d->n_dofs, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(keys_in);
  cudaFree(keys_out);
  cudaFree(vals_in);
  cudaFree(tmp);
  if (e != cudaSuccess) {
    cudaFree(vals_out);
    cudaFree(ptr);
    LFGPU_FAIL(ctx, LFGPU_ERR_CUDA, std::string("dofmap gather plan: ") + cudaGetErrorString(e));
  }
  d->g_ptr = ptr;
  d->g_items = vals_out;  // the first n_items entries are the used slots
  d->g_n_items = n_items;
  d->g_state = 1;
  return LFGPU_OK;
}

using namespace lfgpu;

extern "C" {

void lfgpu_dofmap_destroy(lfgpu_dofmap* d) {
  if (d == nullptr) return;
  if (d->ctx) {
    cudaSetDevice(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
  }
  cudaFree(d->cell_dofs);
  cudaFree(d->n_ldof);
  cudaFree(d->g_ptr);
  cudaFree(d->g_items);
  cudaFree(d->lv_nbr);
  cudaFree(d->lv_nbr16);
  cudaFree(d->lv_info);
  cudaFree(d->lv_irregular);
  cudaFree(d->lv_cells);
  cudaFree(d->lv_ev);
  cudaFree(d->lv_pos);
  delete d;
}

int64_t lfgpu_dofmap_num_dofs(const lfgpu_dofmap* d) { return d ? d->n_dofs : -1; }
int lfgpu_dofmap_stride(const lfgpu_dofmap* d) { return d ? d->stride : -1; }

int lfgpu_dofmap_upload(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, int64_t n_dofs, int stride, const int64_t* cell_dofs,
                        const uint8_t* n_ldof, lfgpu_dofmap** out) {
  if (ctx == nullptr || mesh == nullptr || out == nullptr || cell_dofs == nullptr || stride < 1 || stride > kMaxNsf)
    return LFGPU_ERR_INVALID;
  *out = nullptr;
  if (n_dofs < 1 || n_dofs >= (1LL << 31)) LFGPU_FAIL(ctx, LFGPU_ERR_OVERFLOW, "n_dofs does not fit the int32 storage index");
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  const int64_t n = mesh->n_cells;
  auto* d = new lfgpu_dofmap;
  d->ctx = ctx;
  d->n_cells = n;
  d->n_dofs = n_dofs;
  d->stride = stride;
  d->max_ldof = stride;
  int64_t* d_in = nullptr;
  uint8_t* d_nl = nullptr;
  cudaError_t e = cudaMalloc(&d->cell_dofs, sizeof(int32_t) * n * stride);
  if (e == cudaSuccess) e = cudaMalloc(&d->n_ldof, n);
  if (e == cudaSuccess) e = cudaMalloc(&d_in, sizeof(int64_t) * n * stride);
  if (e == cudaSuccess && n_ldof) e = cudaMalloc(&d_nl, n);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, cell_dofs, sizeof(int64_t) * n * stride, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && n_ldof) e = cudaMemcpyAsync(d_nl, n_ldof, n, cudaMemcpyHostToDevice, ctx->stream);
  int h_flags[2] = {0, 0};
  if (e == cudaSuccess) {
    int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 64);
    cudaMemsetAsync(d_flags, 0, 64, ctx->stream);
    k_convert_dofs<<<static_cast<unsigned>(cdiv(n, kThreads)), kThreads, 0, ctx->stream>>>(n, stride, d_in, d_nl, n_dofs, d->cell_dofs, d->n_ldof, d_flags);
    ctx->launches++;
    e = cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(d_in);
  cudaFree(d_nl);
  if (e != cudaSuccess) {
    lfgpu_dofmap_destroy(d);
    LFGPU_FAIL(ctx, LFGPU_ERR_CUDA, std::string("dofmap upload: ") + cudaGetErrorString(e));
  }
  if (h_flags[0] || h_flags[1]) {
    lfgpu_dofmap_destroy(d);
    LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, h_flags[0] ? "cell_dofs entry >= n_dofs" : "cell_dofs rows must be packed (no holes)");
  }
  *out = d;
  return LFGPU_OK;
}

int lfgpu_dofmap_uniform(lfgpu_ctx* ctx, lfgpu_mesh* mesh, int n_pt, int n_seg, int n_tria, int n_quad, lfgpu_dofmap** out) {
  if (ctx == nullptr || mesh == nullptr || out == nullptr || n_pt < 0 || n_seg < 0 || n_tria < 0 || n_quad < 0) return LFGPU_ERR_INVALID;
  *out = nullptr;
  const int tria_total = 3 * n_pt + 3 * n_seg + n_tria, quad_total = 4 * n_pt + 4 * n_seg + n_quad;
  const int stride = tria_total > quad_total ? tria_total : quad_total;  // dofhandler.cc:138
  if (stride < 1 || stride > kMaxNsf) LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "local dof count must be 1..16");
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  if (n_seg > 0) {
    const int rc = ensure_topology(ctx, mesh);
    if (rc != LFGPU_OK) return rc;
  }
  const int64_t n = mesh->n_cells;
  const int64_t n_edges = n_seg > 0 ? mesh->n_edges : 0;
  const int64_t edge_base = mesh->n_nodes * n_pt;
  const int64_t cell_base = edge_base + n_edges * n_seg;
  const int64_t n_dofs = cell_base + mesh->n_tria * n_tria + mesh->n_quad * n_quad;
  if (n_dofs < 1 || n_dofs >= (1LL << 31)) LFGPU_FAIL(ctx, LFGPU_ERR_OVERFLOW, "n_dofs does not fit the int32 storage index");
  auto* d = new lfgpu_dofmap;
  d->ctx = ctx;
  d->n_cells = n;
  d->n_dofs = n_dofs;
  d->stride = stride;
  d->max_ldof = stride;
  d->n_pt = n_pt;
  d->n_seg = n_seg;
  d->n_nodes = mesh->n_nodes;
  int64_t *counts = nullptr, *offsets = nullptr;
  void* tmp = nullptr;
  cudaError_t e = cudaMalloc(&d->cell_dofs, sizeof(int32_t) * n * stride);
  if (e == cudaSuccess) e = cudaMalloc(&d->n_ldof, n);
  const bool mixed = (n_tria != n_quad) && mesh->n_tria > 0 && mesh->n_quad > 0;
  if (e == cudaSuccess && mixed) {
    e = cudaMalloc(&counts, sizeof(int64_t) * n);
    if (e == cudaSuccess) e = cudaMalloc(&offsets, sizeof(int64_t) * n);
    if (e == cudaSuccess) {
      k_interior_counts<<<static_cast<unsigned>(cdiv(n, kThreads)), kThreads, 0, ctx->stream>>>(n, mesh->cell_nodes, n_tria, n_quad, counts);
      ctx->launches++;
      size_t tb = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb, counts, offsets, n, ctx->stream);
      e = cudaMalloc(&tmp, tb);
      if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tb, counts, offsets, n, ctx->stream);
    }
  }
  if (e == cudaSuccess) {
    k_uniform_dofs<<<static_cast<unsigned>(cdiv(n, kThreads)), kThreads, 0, ctx->stream>>>(
        n, mesh->cell_nodes, mesh->cell_edges, mesh->cell_edge_ori, mixed ? offsets : nullptr, n_pt, n_seg, n_tria, n_quad,
        edge_base, cell_base, stride, d->cell_dofs, d->n_ldof);
    ctx->launches++;
    e = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(counts);
  cudaFree(offsets);
  cudaFree(tmp);
  if (e != cudaSuccess) {
    lfgpu_dofmap_destroy(d);
    LFGPU_FAIL(ctx, LFGPU_ERR_CUDA, std::string("dofmap_uniform: ") + cudaGetErrorString(e));
  }
  *out = d;
  return LFGPU_OK;
}

int lfgpu_dofmap_dynamic(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const uint32_t* n_int_node, const uint32_t* n_int_edge,
                         const uint32_t* n_int_cell, lfgpu_dofmap** out) {
  if (ctx == nullptr || mesh == nullptr || out == nullptr) return LFGPU_ERR_INVALID;
  *out = nullptr;
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  if (n_int_edge != nullptr) {
    const int rc = ensure_topology(ctx, mesh);
    if (rc != LFGPU_OK) return rc;
  }
  cudaStream_t st = ctx->stream;
  const int64_t n_ent[3] = {mesh->n_nodes, n_int_edge != nullptr ? mesh->n_edges : 0, mesh->n_cells};
  const uint32_t* h_cnt[3] = {n_int_node, n_int_edge, n_int_cell};
  uint32_t* d_cnt[3] = {nullptr, nullptr, nullptr};
  int64_t* d_wide[3] = {nullptr, nullptr, nullptr};
  int64_t* d_off[3] = {nullptr, nullptr, nullptr};
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  lfgpu_dofmap* d = nullptr;
  auto cleanup = [&]() {
    for (int k = 0; k < 3; ++k) {
      cudaFree(d_cnt[k]);
      cudaFree(d_wide[k]);
      cudaFree(d_off[k]);
    }
    cudaFree(tmp);
  };
#define DYN_CHECK(expr)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      set_last_error(ctx, std::string("dofmap_dynamic: ") + cudaGetErrorString(_e));     \
      cleanup();                                                                         \
      lfgpu_dofmap_destroy(d);                                                           \
      return LFGPU_ERR_CUDA;                                                             \
    }                                                                                    \
  } while (0)
  int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 64);
  DYN_CHECK(cudaMemsetAsync(d_flags, 0, 64, st));
  int64_t totals[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) {
    const int64_t n = n_ent[k];
    if (h_cnt[k] != nullptr && n > 0) {
      DYN_CHECK(cudaMalloc(&d_cnt[k], sizeof(uint32_t) * n));
      DYN_CHECK(cudaMemcpyAsync(d_cnt[k], h_cnt[k], sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    }
    DYN_CHECK(cudaMalloc(&d_wide[k], sizeof(int64_t) * (n + 1)));
    DYN_CHECK(cudaMalloc(&d_off[k], sizeof(int64_t) * (n + 1)));
    k_widen_counts<<<static_cast<unsigned>(cdiv(n + 1, kThreads)), kThreads, 0, st>>>(n, d_cnt[k], d_wide[k], d_flags);
    ctx->launches++;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, d_wide[k], d_off[k], n + 1, st);
    if (tb > tmp_bytes) {
      DYN_CHECK(cudaStreamSynchronize(st));
      cudaFree(tmp);
      tmp = nullptr;
      DYN_CHECK(cudaMalloc(&tmp, tb));
      tmp_bytes = tb;
    }
    DYN_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_wide[k], d_off[k], n + 1, st));
    DYN_CHECK(cudaMemcpyAsync(&totals[k], d_off[k] + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  }
  k_dynamic_lengths<<<static_cast<unsigned>(cdiv(mesh->n_cells, kThreads)), kThreads, 0, st>>>(
      mesh->n_cells, mesh->cell_nodes, mesh->cell_edges, d_off[0], n_int_edge != nullptr ? d_off[1] : nullptr, d_off[2], d_flags + 1);
  ctx->launches++;
  int h_flags[2] = {0, 0};
  DYN_CHECK(cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
  DYN_CHECK(cudaStreamSynchronize(st));
  const int stride = h_flags[1];
  if (h_flags[0] != 0 || stride > kMaxNsf) {
    cleanup();
    LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "a cell may carry at most 16 local dofs");
  }
  const int64_t n_dofs = totals[0] + totals[1] + totals[2];
  if (n_dofs < 1 || stride < 1) {
    cleanup();
    LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "the layout assigns no dofs");
  }
  if (n_dofs >= (1LL << 31)) {
    cleanup();
    LFGPU_FAIL(ctx, LFGPU_ERR_OVERFLOW, "n_dofs does not fit the int32 storage index");
  }
  d = new lfgpu_dofmap;
  d->ctx = ctx;
  d->n_cells = mesh->n_cells;
  d->n_dofs = n_dofs;
  d->stride = stride;
  d->max_ldof = stride;
  d->n_nodes = mesh->n_nodes;
  DYN_CHECK(cudaMalloc(&d->cell_dofs, sizeof(int32_t) * mesh->n_cells * stride));
  DYN_CHECK(cudaMalloc(&d->n_ldof, mesh->n_cells));
  k_dynamic_dofs<<<static_cast<unsigned>(cdiv(mesh->n_cells, kThreads)), kThreads, 0, st>>>(
      mesh->n_cells, mesh->cell_nodes, mesh->cell_edges, mesh->cell_edge_ori, d_off[0], n_int_edge != nullptr ? d_off[1] : nullptr,
      d_off[2], totals[0], totals[0] + totals[1], stride, d->cell_dofs, d->n_ldof);
  ctx->launches++;
  DYN_CHECK(cudaGetLastError());
  DYN_CHECK(cudaStreamSynchronize(st));
#undef DYN_CHECK
  cleanup();
  *out = d;
  return LFGPU_OK;
}

int lfgpu_dofmap_lagrange(lfgpu_ctx* ctx, lfgpu_mesh* mesh, int degree, lfgpu_dofmap** out) {
  // uniform_scalar_fe_space.h:334-341 with the interior dof counts of FeLagrangeO{1,2,3}: {1,0,0,0} / {1,1,0,1} / {1,2,1,4}
  switch (degree) {
    case 1: return lfgpu_dofmap_uniform(ctx, mesh, 1, 0, 0, 0, out);
    case 2: return lfgpu_dofmap_uniform(ctx, mesh, 1, 1, 0, 1, out);
    case 3: return lfgpu_dofmap_uniform(ctx, mesh, 1, 2, 1, 4, out);
    default:
      if (ctx) set_last_error(ctx, "degree must be 1, 2 or 3");
      return LFGPU_ERR_INVALID;
  }
}

int lfgpu_dofmap_dof_coords(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_dofmap* d, int n_tria, int n_quad, double* d_xy) {
  if (ctx == nullptr || mesh == nullptr || d == nullptr || d_xy == nullptr) return LFGPU_ERR_INVALID;
  if (d->n_pt < 0 || d->n_seg < 0 || d->n_nodes != mesh->n_nodes || d->n_cells != mesh->n_cells)
    LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "dof positions need a dof map built by lfgpu_dofmap_uniform / _lagrange on this mesh");
  if (d->n_pt > 1 || d->n_seg > 2 || n_tria < 0 || n_tria > 1 || (n_quad != 0 && n_quad != 1 && n_quad != 4))
    LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "dof positions are defined for the Lagrange layouts of degree 1..3");
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  if (d->n_seg > 0) {
    const int rc = ensure_topology(ctx, mesh);
    if (rc != LFGPU_OK) return rc;
  }
  cudaStream_t st = ctx->stream;
  if (d->n_pt > 0) {
    k_dof_xy_nodes<<<static_cast<unsigned>(cdiv(mesh->n_nodes, kThreads)), kThreads, 0, st>>>(mesh->n_nodes, d->n_pt, mesh->node_coords, d_xy);
    ctx->launches++;
  }
  const int64_t edge_base = mesh->n_nodes * d->n_pt;
  if (d->n_seg > 0 && mesh->n_edges > 0) {
    k_dof_xy_edges<<<static_cast<unsigned>(cdiv(mesh->n_edges, kThreads)), kThreads, 0, st>>>(mesh->n_edges, d->n_seg, edge_base, mesh->edge_nodes,
                                                                                             mesh->node_coords, d_xy);
    ctx->launches++;
  }
  if (n_tria > 0 || n_quad > 0) {
    k_dof_xy_cells<<<static_cast<unsigned>(cdiv(mesh->n_cells, kThreads)), kThreads, 0, st>>>(mesh->n_cells, d->stride, d->n_pt, d->n_seg, n_tria, n_quad,
                                                                                             mesh->cell_nodes, mesh->node_coords, mesh->cell_coords,
                                                                                             d->cell_dofs, d_xy);
    ctx->launches++;
  }
  LFGPU_CUDA_CHECK(ctx, cudaGetLastError());
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
  return LFGPU_OK;
}

int lfgpu_dofmap_edge_dof_flags(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_dofmap* d, const uint8_t* d_edge_sel, uint8_t* d_flags) {
  if (ctx == nullptr || mesh == nullptr || d == nullptr || d_edge_sel == nullptr || d_flags == nullptr) return LFGPU_ERR_INVALID;
  if (d->n_pt < 0 || d->n_seg < 0 || d->n_nodes != mesh->n_nodes)
    LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "edge dof flags need a dof map built by lfgpu_dofmap_uniform / _lagrange on this mesh");
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  const int rc = ensure_topology(ctx, mesh);
  if (rc != LFGPU_OK) return rc;
  LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_flags, 0, d->n_dofs, ctx->stream));
  if (mesh->n_edges > 0) {
    k_edge_dof_flags<<<static_cast<unsigned>(cdiv(mesh->n_edges, kThreads)), kThreads, 0, ctx->stream>>>(
        mesh->n_edges, d->n_pt, d->n_seg, mesh->n_nodes * d->n_pt, mesh->edge_nodes, d_edge_sel, d_flags);
    ctx->launches++;
  }
  LFGPU_CUDA_CHECK(ctx, cudaGetLastError());
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return LFGPU_OK;
}

int lfgpu_dofmap_download(lfgpu_ctx* ctx, const lfgpu_dofmap* d, int64_t* cell_dofs, uint8_t* n_ldof) {
  if (ctx == nullptr || d == nullptr) return LFGPU_ERR_INVALID;
  const int64_t total = d->n_cells * d->stride;
  if (cell_dofs) {
    int64_t* tmp = nullptr;
    LFGPU_CUDA_CHECK(ctx, cudaMalloc(&tmp, sizeof(int64_t) * total));
    k_dofs_to_i64<<<static_cast<unsigned>(cdiv(total, kThreads)), kThreads, 0, ctx->stream>>>(total, d->cell_dofs, tmp);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(cell_dofs, tmp, sizeof(int64_t) * total, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    LFGPU_CUDA_CHECK(ctx, e);
  }
  if (n_ldof) {
    LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(n_ldof, d->n_ldof, d->n_cells, cudaMemcpyDeviceToHost, ctx->stream));
    LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return LFGPU_OK;
}

}  // extern "C"
