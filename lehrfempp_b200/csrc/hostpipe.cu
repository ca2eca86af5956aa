// Host-buffer entry point of the numeric pass (product code): what a CPU-side caller of AssembleMatrixLocally sees --
// node coordinates of this step in host memory, the assembled values back in host memory.
//
// For a 1.0e8-triangle P1 matrix the kernel takes 1 ms, the PCIe copies 65 ms (0.8 GB in, 2.8 GB out), so the call is
// organised around the copies: the outer indices are cut into blocks; block b is computed as soon as the leading part
// of the coordinate array it needs has arrived, and its values leave on a second copy stream while the next blocks'
// coordinates are still arriving -- H2D and D2H share the link in both directions instead of taking turns.
// Which coordinates a block needs is read off the vertex-ring plan of the fan kernel (assemble_p1.cu) once and cached
// as a running maximum, so the only requirement on the numbering is locality (any numbering is CORRECT; a numbering
// without locality simply degrades to "upload everything first").  Calls that do not run in the fan kernel take the
// plain sequence upload -> assemble -> download on the context stream.
#include <algorithm>

#include "lfgpu_internal.cuh"

namespace lfgpu {
namespace {
constexpr int kThreads = 256;

// need[b] = 1 + largest node index the fan kernel reads for the rows of block b.  The fan kernel runs on nodal P1
// tables only, where the stored columns of row r are exactly node r and its ring (and the pattern is symmetric).
__global__ void k_block_need(int64_t n_rows, const int32_t* __restrict__ outer, const int32_t* __restrict__ inner,
                             int64_t rows_per_block, unsigned long long* __restrict__ need) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  unsigned long long m = static_cast<unsigned long long>(r);
  for (int32_t k = outer[r]; k < outer[r + 1]; ++k) m = max(m, static_cast<unsigned long long>(inner[k]));
  atomicMax(need + r / rows_per_block, m + 1);
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

int build_plan(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, lfgpu_pattern* p, int nb) {
  if (p->hp_blocks == nb) return LFGPU_OK;
  const int64_t N = p->n_outer;
  const int64_t rpb = cdiv(N, nb);
  unsigned long long* d_need = nullptr;
  LFGPU_CUDA_CHECK(ctx, cudaMalloc(&d_need, sizeof(unsigned long long) * nb));
  cudaError_t e = cudaMemsetAsync(d_need, 0, sizeof(unsigned long long) * nb, ctx->stream);
  if (e == cudaSuccess) {
    k_block_need<<<static_cast<unsigned>(cdiv(N, kThreads)), kThreads, 0, ctx->stream>>>(N, p->outer, p->inner, rpb, d_need);
    ctx->launches++;
    e = cudaGetLastError();
  }
  std::vector<unsigned long long> h(nb);
  std::vector<int32_t> ob(nb + 1);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), d_need, sizeof(unsigned long long) * nb, cudaMemcpyDeviceToHost, ctx->stream);
  for (int b = 0; b <= nb && e == cudaSuccess; ++b) {
    const int64_t r = std::min<int64_t>(static_cast<int64_t>(b) * rpb, N);
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
  for (int b = 0; b <= nb; ++b) p->hp_val[b] = ob[b];
  p->hp_blocks = nb;
  return LFGPU_OK;
}
}  // namespace
}  // namespace lfgpu

using namespace lfgpu;

extern "C" int lfgpu_assemble_reaction_diffusion_host(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_pattern* pattern, int degree,
                                                       const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad, const lfgpu_coeff* alpha,
                                                       const lfgpu_coeff* gamma, const double* h_node_coords, double* d_values,
                                                       double* h_values, int algo, int n_blocks) {
  if (ctx == nullptr || mesh == nullptr || pattern == nullptr || d_values == nullptr) return LFGPU_ERR_INVALID;
  if (h_node_coords != nullptr && mesh->cell_coords != nullptr)
    LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "mesh carries explicit cell corner coordinates");
  if (n_blocks <= 0) n_blocks = 16;
  n_blocks = std::min<int64_t>(n_blocks, std::max<int64_t>(1, pattern->n_outer / 4096));
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  auto* p = const_cast<lfgpu_pattern*>(pattern);
  int fan = 0;
  int rc = assemble_rd_impl(ctx, mesh, p, degree, qr_tria, qr_quad, alpha, gamma, nullptr, 0.0, d_values, algo, nullptr, 0, -1, &fan);
  if (rc != LFGPU_OK) return rc;
  const size_t coord_bytes = sizeof(double) * 2 * mesh->n_nodes;
  if (!fan || n_blocks < 2 || (h_node_coords == nullptr && h_values == nullptr)) {
    // plain sequence on the context stream
    if (h_node_coords != nullptr)
      LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(mesh->node_coords, h_node_coords, coord_bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = assemble_rd_impl(ctx, mesh, p, degree, qr_tria, qr_quad, alpha, gamma, nullptr, 0.0, d_values, algo, nullptr, 0, -1, nullptr);
    if (rc != LFGPU_OK) return rc;
    if (h_values != nullptr)
      LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(h_values, d_values, sizeof(double) * p->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return LFGPU_OK;
  }
  // pipelined: H2D stream -> compute stream -> D2H stream, one event per block and hop
  const int nb = n_blocks;
  if ((rc = build_plan(ctx, mesh, p, nb)) != LFGPU_OK) return rc;
  if ((rc = ensure_pipe(ctx, 2 * static_cast<size_t>(nb) + 2)) != LFGPU_OK) return rc;
  const int64_t N = p->n_outer, rpb = cdiv(N, nb);
  cudaEvent_t ev_start = ctx->pipe_events[2 * nb];
  LFGPU_CUDA_CHECK(ctx, cudaEventRecord(ev_start, ctx->stream));  // earlier work on the context stream may still read the coordinates
  LFGPU_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->s_h2d, ev_start, 0));
  int64_t uploaded = 0;
  for (int b = 0; b < nb; ++b) {
    if (h_node_coords != nullptr && p->hp_need[b] > uploaded) {
      LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(mesh->node_coords + 2 * uploaded, h_node_coords + 2 * uploaded,
                                            sizeof(double) * 2 * (p->hp_need[b] - uploaded), cudaMemcpyHostToDevice, ctx->s_h2d));
      uploaded = p->hp_need[b];
      LFGPU_CUDA_CHECK(ctx, cudaEventRecord(ctx->pipe_events[b], ctx->s_h2d));
      LFGPU_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pipe_events[b], 0));
    }
    const int64_t row0 = static_cast<int64_t>(b) * rpb, rows = std::min<int64_t>(rpb, N - row0);
    if (rows <= 0) continue;
    rc = assemble_rd_impl(ctx, mesh, p, degree, qr_tria, qr_quad, alpha, gamma, nullptr, 0.0, d_values, algo, nullptr, rows, row0, nullptr);
    if (rc != LFGPU_OK) break;
    if (h_values != nullptr && p->hp_val[b + 1] > p->hp_val[b]) {
      LFGPU_CUDA_CHECK(ctx, cudaEventRecord(ctx->pipe_events[nb + b], ctx->stream));
      LFGPU_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->s_d2h, ctx->pipe_events[nb + b], 0));
      LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(h_values + p->hp_val[b], d_values + p->hp_val[b],
                                            sizeof(double) * (p->hp_val[b + 1] - p->hp_val[b]), cudaMemcpyDeviceToHost, ctx->s_d2h));
    }
  }
  if (rc == LFGPU_OK && h_node_coords != nullptr && uploaded < mesh->n_nodes) {  // nodes no row refers to
    LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(mesh->node_coords + 2 * uploaded, h_node_coords + 2 * uploaded,
                                          sizeof(double) * 2 * (mesh->n_nodes - uploaded), cudaMemcpyHostToDevice, ctx->s_h2d));
  }
  cudaError_t e1 = cudaStreamSynchronize(ctx->s_h2d);
  cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
  cudaError_t e3 = cudaStreamSynchronize(ctx->s_d2h);
  if (rc != LFGPU_OK) return rc;
  LFGPU_CUDA_CHECK(ctx, e1);
  LFGPU_CUDA_CHECK(ctx, e2);
  LFGPU_CUDA_CHECK(ctx, e3);
  return LFGPU_OK;
}
