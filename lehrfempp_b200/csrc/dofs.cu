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

__global__ void k_dofs_to_i64(int64_t n, const int32_t* __restrict__ in, int64_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

}  // namespace
}  // namespace lfgpu

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
