// Host-buffer entry points of the numeric pass (product code): what a CPU-side caller of AssembleMatrixLocally sees --
// node coordinates of this step in host memory, the assembled values back in host memory.
//
// For a 1.0e8-triangle P1 matrix the kernel takes 1 ms, the PCIe copies 65 ms (0.8 GB in, 2.8 GB out), so the call is
// organised around the copies: the outer indices are cut into blocks; block b is computed as soon as the leading part
// of the coordinate array it needs has arrived, and its values leave on a second copy stream while the next blocks'
// coordinates are still arriving -- H2D and D2H share the link in both directions instead of taking turns.
// Which coordinates a block needs is read off the pattern once (for the nodal P1 tables of the fan kernel the stored
// columns of a row ARE the nodes it reads) and cached as a running maximum, so the only requirement on the numbering
// is locality (any numbering is CORRECT; a numbering without locality simply degrades to "upload everything first").
// The _range form does the same for a contiguous block of rows -- the share of one GPU in a multi-GPU run by row
// blocks: only the window of coordinates those rows refer to is uploaded, only their values are downloaded.
// Calls that do not run in the fan kernel take the plain sequence upload -> assemble -> download on the context stream.
#include <algorithm>

#include "lfgpu_internal.cuh"

namespace lfgpu {
namespace {
constexpr int kThreads = 256;

// hi[b] = 1 + largest, *lo = smallest node index the fan kernel reads for the rows of block b of [row0, row0 + n_rows).
// The fan kernel runs on nodal P1 tables only, where the stored columns of row r are exactly node r and its ring.
__global__ void k_block_need(int64_t row0, int64_t n_rows, const int32_t* __restrict__ outer, const int32_t* __restrict__ inner,
                             int64_t rows_per_block, unsigned long long* __restrict__ hi, unsigned long long* __restrict__ lo) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool act = t < n_rows;
  unsigned long long m = 0ULL, l = ~0ULL;
  long long b = -1;
  if (act) {
    const int64_t r = row0 + t;
    b = t / rows_per_block;
    m = l = static_cast<unsigned long long>(r);
    for (int32_t k = outer[r]; k < outer[r + 1]; ++k) {
      const unsigned long long c = static_cast<unsigned long long>(inner[k]);
      m = max(m, c);
      l = min(l, c);
    }
  }
  // one atomic per warp (the rows of a warp almost always share the block): 5e7 threads on 17 addresses otherwise
  const unsigned mask = __ballot_sync(0xffffffffU, act);
  if (mask == 0U) return;
  const int leader = __ffs(mask) - 1;
  const long long b0 = __shfl_sync(0xffffffffU, b, leader);
  if (__all_sync(0xffffffffU, !act || b == b0)) {
    for (int o = 16; o > 0; o >>= 1) {
      m = max(m, __shfl_xor_sync(0xffffffffU, m, o));
      l = min(l, __shfl_xor_sync(0xffffffffU, l, o));
    }
    if ((threadIdx.x & 31) == leader) {
      atomicMax(hi + b0, m + 1);
      atomicMin(lo, l);
    }
  } else if (act) {
    atomicMax(hi + b, m + 1);
    atomicMin(lo, l);
  }
}

int ensure_pipe(lfgpu_ctx* ctx, size_t n_events) {
  if (ctx->s_h2d == nullptr) LFGPU_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
  if (ctx->s_d2h == nullptr) LFGPU_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
  while (ctx->pipe_events.size() < n_events) {
    cudaEvent_t e;
    LFGPU_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->pipe_events.push_back(e);
  }
  return LFGPU_OK;
}

int build_plan(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, lfgpu_pattern* p, int nb, int64_t row0, int64_t n_rows) {
  if (p->hp_blocks == nb && p->hp_row0 == row0 && p->hp_rows == n_rows) return LFGPU_OK;
  const int64_t rpb = cdiv(n_rows, nb);
  unsigned long long* d_need = nullptr;  // [nb] hi, [nb] = lo
  LFGPU_CUDA_CHECK(ctx, cudaMalloc(&d_need, sizeof(unsigned long long) * (nb + 1)));
  std::vector<unsigned long long> h(nb + 1, 0ULL);
  h[nb] = ~0ULL;
  cudaError_t e = cudaMemcpyAsync(d_need, h.data(), sizeof(unsigned long long) * (nb + 1), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    k_block_need<<<static_cast<unsigned>(cdiv(n_rows, kThreads)), kThreads, 0, ctx->stream>>>(row0, n_rows, p->outer, p->inner, rpb, d_need,
                                                                                             d_need + nb);
    ctx->launches++;
    e = cudaGetLastError();
  }
  std::vector<int32_t> ob(nb + 1);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), d_need, sizeof(unsigned long long) * (nb + 1), cudaMemcpyDeviceToHost, ctx->stream);
  for (int b = 0; b <= nb && e == cudaSuccess; ++b) {
    const int64_t r = row0 + std::min<int64_t>(static_cast<int64_t>(b) * rpb, n_rows);
    e = cudaMemcpyAsync(&ob[b], p->outer + r, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_need);
  LFGPU_CUDA_CHECK(ctx, e);
  p->hp_need.assign(nb, 0);
  p->hp_val.assign(nb + 1, 0);
  int64_t run = 0;
  for (int b = 0; b < nb; ++b) {
    run = std::max<int64_t>(run, static_cast<int64_t>(h[b]));
    p->hp_need[b] = std::min<int64_t>(run, mesh->n_nodes);
  }
  p->hp_lo = std::min<int64_t>(static_cast<int64_t>(std::min<unsigned long long>(h[nb], static_cast<unsigned long long>(mesh->n_nodes))), run);
  for (int b = 0; b <= nb; ++b) p->hp_val[b] = ob[b];
  p->hp_blocks = nb;
  p->hp_row0 = row0;
  p->hp_rows = n_rows;
  return LFGPU_OK;
}

// whole = all rows, every coordinate uploaded, h_values indexed like the value array;
// !whole = rows [row0, row0 + n_rows), only the coordinate window they refer to, h_values = the range's values
int host_impl(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_pattern* pattern, int degree, const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad,
              const lfgpu_coeff* alpha, const lfgpu_coeff* gamma, const double* h_node_coords, double* d_values, double* h_values, int algo,
              int n_blocks, bool whole, int64_t row0, int64_t n_rows) {
  if (ctx == nullptr || mesh == nullptr || pattern == nullptr || d_values == nullptr) return LFGPU_ERR_INVALID;
  if (h_node_coords != nullptr && mesh->cell_coords != nullptr)
    LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "mesh carries explicit cell corner coordinates");
  if (!whole && (row0 < 0 || n_rows < 0 || row0 + n_rows > pattern->n_outer)) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "row range outside the matrix");
  if (whole) {
    row0 = 0;
    n_rows = pattern->n_outer;
  }
  if (n_rows == 0) return LFGPU_OK;
  if (n_blocks <= 0) n_blocks = 16;
  n_blocks = static_cast<int>(std::min<int64_t>(n_blocks, std::max<int64_t>(1, n_rows / 4096)));
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  auto* p = const_cast<lfgpu_pattern*>(pattern);
  int fan = 0;
  int rc = assemble_rd_impl(ctx, mesh, p, degree, qr_tria, qr_quad, alpha, gamma, nullptr, 0.0, d_values, algo, nullptr, 0, -1, &fan);
  if (rc != LFGPU_OK) return rc;
  if (!fan && !whole) LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "contiguous row ranges are a fan-kernel feature");
  if (!fan || (whole && (n_blocks < 2 || (h_node_coords == nullptr && h_values == nullptr)))) {
    // plain sequence on the context stream
    if (h_node_coords != nullptr) {
      LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(mesh->node_coords, h_node_coords, sizeof(double) * 2 * mesh->n_nodes, cudaMemcpyHostToDevice, ctx->stream));
      mesh->coords_version++;
    }
    rc = assemble_rd_impl(ctx, mesh, p, degree, qr_tria, qr_quad, alpha, gamma, nullptr, 0.0, d_values, algo, nullptr, 0, -1, nullptr);
    if (rc != LFGPU_OK) return rc;
    if (h_values != nullptr)
      LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(h_values, d_values, sizeof(double) * p->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    // positions that came from the caller are checked like those of lfgpu_mesh_upload (degenerate cell -> LFGPU_ERR_DEGENERATE)
    if (h_node_coords != nullptr && (rc = queue_geometry_check(ctx, mesh)) != LFGPU_OK) return rc;
    return lfgpu_ctx_synchronize(ctx);
  }
  // pipelined: H2D stream -> compute stream -> D2H stream, one event per block and hop
  const int nb = n_blocks;
  if ((rc = build_plan(ctx, mesh, p, nb, row0, n_rows)) != LFGPU_OK) return rc;
  if ((rc = ensure_pipe(ctx, 2 * static_cast<size_t>(nb) + 2)) != LFGPU_OK) return rc;
  const int64_t rpb = cdiv(n_rows, nb);
  cudaEvent_t ev_start = ctx->pipe_events[2 * nb];
  LFGPU_CUDA_CHECK(ctx, cudaEventRecord(ev_start, ctx->stream));  // earlier work on the context stream may still read the coordinates
  LFGPU_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->s_h2d, ev_start, 0));
  int64_t uploaded = whole ? 0 : p->hp_lo;
  const int64_t val_base = whole ? 0 : p->hp_val[0];  // h_values[0] is this value
  for (int b = 0; b < nb; ++b) {
    if (h_node_coords != nullptr && p->hp_need[b] > uploaded) {
      LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(mesh->node_coords + 2 * uploaded, h_node_coords + 2 * uploaded,
                                            sizeof(double) * 2 * (p->hp_need[b] - uploaded), cudaMemcpyHostToDevice, ctx->s_h2d));
      uploaded = p->hp_need[b];
      mesh->coords_version++;
      LFGPU_CUDA_CHECK(ctx, cudaEventRecord(ctx->pipe_events[b], ctx->s_h2d));
      LFGPU_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pipe_events[b], 0));
    }
    const int64_t off = static_cast<int64_t>(b) * rpb, rows = std::min<int64_t>(rpb, n_rows - off);
    if (rows <= 0) continue;
    rc = assemble_rd_impl(ctx, mesh, p, degree, qr_tria, qr_quad, alpha, gamma, nullptr, 0.0, d_values, algo, nullptr, rows, row0 + off, nullptr);
    if (rc != LFGPU_OK) break;
    if (h_values != nullptr && p->hp_val[b + 1] > p->hp_val[b]) {
      LFGPU_CUDA_CHECK(ctx, cudaEventRecord(ctx->pipe_events[nb + b], ctx->stream));
      LFGPU_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->s_d2h, ctx->pipe_events[nb + b], 0));
      LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(h_values + (p->hp_val[b] - val_base), d_values + p->hp_val[b],
                                            sizeof(double) * (p->hp_val[b + 1] - p->hp_val[b]), cudaMemcpyDeviceToHost, ctx->s_d2h));
    }
  }
  if (rc == LFGPU_OK && whole && h_node_coords != nullptr && uploaded < mesh->n_nodes) {  // nodes no row refers to
    LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(mesh->node_coords + 2 * uploaded, h_node_coords + 2 * uploaded,
                                          sizeof(double) * 2 * (mesh->n_nodes - uploaded), cudaMemcpyHostToDevice, ctx->s_h2d));
    mesh->coords_version++;
  }
  // the uploaded positions are checked behind the last block's kernel (hidden by the last download); whole-mesh form only: a
  // row-range call sees a window of the coordinates
  const bool check = rc == LFGPU_OK && whole && h_node_coords != nullptr;
  cudaError_t e1 = cudaStreamSynchronize(ctx->s_h2d);
  if (check && e1 == cudaSuccess) rc = queue_geometry_check(ctx, mesh);
  cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
  cudaError_t e3 = cudaStreamSynchronize(ctx->s_d2h);
  if (rc != LFGPU_OK) return rc;
  LFGPU_CUDA_CHECK(ctx, e1);
  LFGPU_CUDA_CHECK(ctx, e2);
  LFGPU_CUDA_CHECK(ctx, e3);
  return lfgpu_ctx_synchronize(ctx);  // reads the flag of the check
}
}  // namespace
}  // namespace lfgpu

using namespace lfgpu;

extern "C" int lfgpu_assemble_reaction_diffusion_host(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_pattern* pattern, int degree,
                                                       const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad, const lfgpu_coeff* alpha,
                                                       const lfgpu_coeff* gamma, const double* h_node_coords, double* d_values,
                                                       double* h_values, int algo, int n_blocks) {
  return host_impl(ctx, mesh, pattern, degree, qr_tria, qr_quad, alpha, gamma, h_node_coords, d_values, h_values, algo, n_blocks, true, 0, 0);
}

extern "C" int lfgpu_assemble_reaction_diffusion_host_range(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_pattern* pattern, int degree,
                                                             const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad, const lfgpu_coeff* alpha,
                                                             const lfgpu_coeff* gamma, const double* h_node_coords, double* d_values,
                                                             double* h_values_range, int algo, int n_blocks, int64_t row0,
                                                             int64_t n_rows) {
  return host_impl(ctx, mesh, pattern, degree, qr_tria, qr_quad, alpha, gamma, h_node_coords, d_values, h_values_range, algo, n_blocks, false,
                   row0, n_rows);
}
