// Consumer side of the assembled matrix (product code) -- SURVEY.md section 8f row 4: sparse matrix-vector product and a
// (Jacobi-preconditioned) conjugate-gradient solve on the compressed arrays the numeric pass produced, so that the chain
// assemble -> edge terms -> Dirichlet elimination -> solve never leaves the device.  The reference hands
// COOMatrix::makeSparse() to an Eigen solver on the host (examples/ellbvp_linfe/homDir_linfe_demo.cc:166-175).
//
// SpMV: a group of LANES threads per outer index (LANES = 1..32, chosen from the mean segment length) walks the
// segment with coalesced loads and reduces with shuffles; the compressed-column layout (Eigen's default) scatters with
// FP64 atomics instead.  Dot products are two-stage with a fixed grid, hence bitwise repeatable.
#include <cstdlib>

#include "lfgpu_internal.cuh"

namespace lfgpu {
namespace {
constexpr int kThreads = 256;
constexpr int kDotBlocks = 592;  // 4 per SM

template <int LANES>
__global__ void __launch_bounds__(kThreads) k_spmv_rows(int64_t n, const int32_t* __restrict__ outer, const int32_t* __restrict__ inner,
                                                        const double* __restrict__ values, const double* __restrict__ x,
                                                        double* __restrict__ y) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t r = t / LANES;
  const int l = static_cast<int>(t % LANES);
  double s = 0.0;
  if (r < n) {
    const int32_t e = outer[r + 1];
    for (int32_t k = outer[r] + l; k < e; k += LANES) s += values[k] * __ldg(x + inner[k]);
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffU, s, o, LANES);
  if (r < n && l == 0) y[r] = s;
}
__global__ void k_spmv_cols(int64_t n, const int32_t* __restrict__ outer, const int32_t* __restrict__ inner,
                            const double* __restrict__ values, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const double xc = x[c];
  for (int32_t k = outer[c]; k < outer[c + 1]; ++k) atomicAdd(y + inner[k], values[k] * xc);
}
// dinv[r] = 1 / A(r, r) (1 if the diagonal is missing or zero); the diagonal sits in segment r in both layouts
__global__ void k_diag_inv(int64_t n, const int32_t* __restrict__ outer, const int32_t* __restrict__ inner,
                           const double* __restrict__ values, double* __restrict__ dinv) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double d = 0.0;
  for (int32_t k = outer[r]; k < outer[r + 1]; ++k)
    if (inner[k] == r) d = values[k];
  dinv[r] = d != 0.0 ? 1.0 / d : 1.0;
}

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double s[kThreads / 32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffU, v, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  v = threadIdx.x < kThreads / 32 ? s[threadIdx.x] : 0.0;
  if (threadIdx.x < 32)
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffU, v, o);
  __syncthreads();
  return v;
}
// partial[b] = sum over the block's grid-stride share of a[i] * b[i]
__global__ void __launch_bounds__(kThreads) k_dot_partial(int64_t n, const double* __restrict__ a, const double* __restrict__ b,
                                                          double* __restrict__ partial) {
  double s = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    s += a[i] * b[i];
  s = block_sum(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void __launch_bounds__(kThreads) k_sum_partials(int n_partial, const double* __restrict__ partial, double* __restrict__ out) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n_partial; i += kThreads) s += partial[i];
  s = block_sum(s);
  if (threadIdx.x == 0) *out = s;
}
// x += alpha p;  r -= alpha q;  z = dinv .* r (or r);  partial[b] = sum r .* z
__global__ void __launch_bounds__(kThreads) k_cg_update(int64_t n, double alpha, const double* __restrict__ p, const double* __restrict__ q,
                                                        const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r,
                                                        double* __restrict__ z, double* __restrict__ partial_rz,
                                                        double* __restrict__ partial_rr) {
  double s = 0.0, t = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * q[i];
    r[i] = ri;
    const double zi = dinv != nullptr ? dinv[i] * ri : ri;
    z[i] = zi;
    s += ri * zi;
    t += ri * ri;
  }
  s = block_sum(s);
  t = block_sum(t);
  if (threadIdx.x == 0) {
    partial_rz[blockIdx.x] = s;
    partial_rr[blockIdx.x] = t;
  }
}
// p = z + beta p
__global__ void k_cg_direction(int64_t n, double beta, const double* __restrict__ z, double* __restrict__ p) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = z[i] + beta * p[i];
}
// r = b - q;  z = dinv .* r
__global__ void k_cg_residual(int64_t n, const double* __restrict__ b, const double* __restrict__ q, const double* __restrict__ dinv,
                              double* __restrict__ r, double* __restrict__ z, double* __restrict__ p) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ri = b[i] - q[i];
  r[i] = ri;
  const double zi = dinv != nullptr ? dinv[i] * ri : ri;
  z[i] = zi;
  p[i] = zi;
}

int spmv(lfgpu_ctx* ctx, const lfgpu_pattern* p, const double* d_values, const double* d_x, double* d_y, bool treat_outer_as_rows) {
  cudaStream_t st = ctx->stream;
  const int64_t n = p->n_outer;
  if (p->major == LFGPU_ROW_MAJOR || treat_outer_as_rows) {
    const double mean = n > 0 ? static_cast<double>(p->nnz) / static_cast<double>(n) : 0.0;
    static const int lanes_env = [] { const char* e = std::getenv("LFGPU_SPMV_LANES"); return e != nullptr ? std::atoi(e) : 0; }();
    // measured at 7 entries per row (3.5e8 entries): 1 lane 0.85 ms, 2 lanes 0.97 ms, 4 lanes 1.55 ms, 8 lanes 2.53 ms
    int lanes = mean <= 10 ? 1 : (mean <= 20 ? 2 : (mean <= 40 ? 4 : (mean <= 80 ? 8 : (mean <= 160 ? 16 : 32))));
    if (lanes_env == 1 || lanes_env == 2 || lanes_env == 4 || lanes_env == 8 || lanes_env == 16 || lanes_env == 32) lanes = lanes_env;
    const unsigned grid = static_cast<unsigned>(cdiv(n * lanes, kThreads));
    switch (lanes) {
      case 1: k_spmv_rows<1><<<grid, kThreads, 0, st>>>(n, p->outer, p->inner, d_values, d_x, d_y); break;
      case 2: k_spmv_rows<2><<<grid, kThreads, 0, st>>>(n, p->outer, p->inner, d_values, d_x, d_y); break;
      case 4: k_spmv_rows<4><<<grid, kThreads, 0, st>>>(n, p->outer, p->inner, d_values, d_x, d_y); break;
      case 8: k_spmv_rows<8><<<grid, kThreads, 0, st>>>(n, p->outer, p->inner, d_values, d_x, d_y); break;
      case 16: k_spmv_rows<16><<<grid, kThreads, 0, st>>>(n, p->outer, p->inner, d_values, d_x, d_y); break;
      default: k_spmv_rows<32><<<grid, kThreads, 0, st>>>(n, p->outer, p->inner, d_values, d_x, d_y); break;
    }
  } else {
    LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_y, 0, sizeof(double) * p->n_inner, st));
    k_spmv_cols<<<static_cast<unsigned>(cdiv(n, kThreads)), kThreads, 0, st>>>(n, p->outer, p->inner, d_values, d_x, d_y);
  }
  LFGPU_LAUNCH_CHECK(ctx);
  return LFGPU_OK;
}
}  // namespace
}  // namespace lfgpu

using namespace lfgpu;

extern "C" {

int lfgpu_spmv(lfgpu_ctx* ctx, const lfgpu_pattern* p, const double* d_values, const double* d_x, double* d_y) {
  if (ctx == nullptr || p == nullptr || d_values == nullptr || d_x == nullptr || d_y == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  return spmv(ctx, p, d_values, d_x, d_y, false);
}

int lfgpu_cg_solve(lfgpu_ctx* ctx, const lfgpu_pattern* p, const double* d_values, const double* d_rhs, double* d_x, double rel_tol,
                   int max_iter, int jacobi, int* iters_out, double* rel_res_out) {
  if (ctx == nullptr || p == nullptr || d_values == nullptr || d_rhs == nullptr || d_x == nullptr || max_iter < 0) return LFGPU_ERR_INVALID;
  if (p->n_outer != p->n_inner) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "Matrix must be square!");
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int64_t n = p->n_outer;
  double *r = nullptr, *z = nullptr, *d = nullptr, *q = nullptr, *dinv = nullptr, *part = nullptr, *scal = nullptr;
  auto cleanup = [&]() { cudaFree(r); cudaFree(z); cudaFree(d); cudaFree(q); cudaFree(dinv); cudaFree(part); cudaFree(scal); };
#define CG_CHECK(expr)                                                        \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) {                                                  \
      set_last_error(ctx, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
      cleanup();                                                              \
      return LFGPU_ERR_CUDA;                                                  \
    }                                                                         \
  } while (0)
  const size_t vb = sizeof(double) * static_cast<size_t>(n > 0 ? n : 1);
  CG_CHECK(cudaMalloc(&r, vb));
  CG_CHECK(cudaMalloc(&z, vb));
  CG_CHECK(cudaMalloc(&d, vb));
  CG_CHECK(cudaMalloc(&q, vb));
  if (jacobi) CG_CHECK(cudaMalloc(&dinv, vb));
  CG_CHECK(cudaMalloc(&part, sizeof(double) * 2 * kDotBlocks));
  CG_CHECK(cudaMalloc(&scal, sizeof(double) * 4));
  const unsigned gn = static_cast<unsigned>(cdiv(n, kThreads));
  auto dot = [&](const double* a, const double* b, double* host_out) -> cudaError_t {
    k_dot_partial<<<kDotBlocks, kThreads, 0, st>>>(n, a, b, part);
    k_sum_partials<<<1, kThreads, 0, st>>>(kDotBlocks, part, scal);
    ctx->launches += 2;
    cudaError_t e = cudaMemcpyAsync(host_out, scal, sizeof(double), cudaMemcpyDeviceToHost, st);
    return e == cudaSuccess ? cudaStreamSynchronize(st) : e;
  };
  // the matrix is symmetric (CG): segment r is row r in either layout
  if (jacobi) {
    k_diag_inv<<<gn, kThreads, 0, st>>>(n, p->outer, p->inner, d_values, dinv);
    ctx->launches++;
  }
  int rc = spmv(ctx, p, d_values, d_x, q, true);
  if (rc != LFGPU_OK) {
    cleanup();
    return rc;
  }
  k_cg_residual<<<gn, kThreads, 0, st>>>(n, d_rhs, q, dinv, r, z, d);
  ctx->launches++;
  double bb = 0.0, rz = 0.0, rr = 0.0;
  CG_CHECK(dot(d_rhs, d_rhs, &bb));
  CG_CHECK(dot(r, z, &rz));
  CG_CHECK(dot(r, r, &rr));
  const double bnorm = bb > 0.0 ? sqrt(bb) : 1.0;
  int it = 0;
  while (it < max_iter && sqrt(rr) > rel_tol * bnorm) {
    rc = spmv(ctx, p, d_values, d, q, true);
    if (rc != LFGPU_OK) {
      cleanup();
      return rc;
    }
    double dq = 0.0;
    CG_CHECK(dot(d, q, &dq));
    if (!(dq > 0.0)) {  // not positive definite along this direction (or breakdown)
      cleanup();
      LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "conjugate gradients: the matrix is not symmetric positive definite");
    }
    const double alpha = rz / dq;
    k_cg_update<<<kDotBlocks, kThreads, 0, st>>>(n, alpha, d, q, dinv, d_x, r, z, part, part + kDotBlocks);
    k_sum_partials<<<1, kThreads, 0, st>>>(kDotBlocks, part, scal);
    k_sum_partials<<<1, kThreads, 0, st>>>(kDotBlocks, part + kDotBlocks, scal + 1);
    ctx->launches += 3;
    double h[2] = {0.0, 0.0};
    CG_CHECK(cudaMemcpyAsync(h, scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CG_CHECK(cudaStreamSynchronize(st));
    const double beta = h[0] / rz;
    rz = h[0];
    rr = h[1];
    k_cg_direction<<<gn, kThreads, 0, st>>>(n, beta, z, d);
    ctx->launches++;
    ++it;
  }
  CG_CHECK(cudaStreamSynchronize(st));
#undef CG_CHECK
  cleanup();
  if (iters_out != nullptr) *iters_out = it;
  if (rel_res_out != nullptr) *rel_res_out = sqrt(rr) / bnorm;
  return LFGPU_OK;
}

}  // extern "C"
