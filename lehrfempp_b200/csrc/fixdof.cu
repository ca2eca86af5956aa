// Dirichlet elimination on the compressed matrix (product code) -- first of the "next" rows of SURVEY.md section 8f.
//
// Stands in for lf::assemble::FixFlaggedSolutionComponents (lib/lf/assemble/fix_dof.h:86-138), which edits the COO
// triplet list:   b <- b - A * xhat (xhat = prescribed values, 0 elsewhere);  b[fixed] <- xhat;  every triplet in a
// fixed row or column is erased (COOMatrix::setZero(pred), coomatrix.h:108-115);  a unit diagonal triplet is appended
// for every fixed dof.  makeSparse() of the result therefore has a SMALLER pattern: the erased entries are gone.
//
// On the device the same happens on the compressed arrays: one pass updates the right-hand side (row-wise gather for
// CSR, column-wise FP64 atomics for the Eigen column-major layout), one pass flags the surviving entries and rewrites the
// fixed diagonals to 1, an exclusive scan + compaction produce the new index arrays -- bit-identical to makeSparse() of
// the edited COO matrix -- or, if the caller wants to keep the pattern of the symbolic pass, the erased entries are left
// in place as explicit zeros.
#include <cub/cub.cuh>

#include "lfgpu_internal.cuh"

namespace lfgpu {
namespace {
constexpr int kThreads = 256;

// b[outer] -= sum_k A[outer, inner_k] * xhat[inner_k]   (row-major: outer = row)
__global__ void k_rhs_rows(int64_t n, const int32_t* __restrict__ outer, const int32_t* __restrict__ inner,
                           const double* __restrict__ values, const uint8_t* __restrict__ fixed, const double* __restrict__ xhat,
                           double* __restrict__ b) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double s = 0.0;
  for (int32_t k = outer[r]; k < outer[r + 1]; ++k) {
    const int32_t c = inner[k];
    if (fixed[c]) s += values[k] * xhat[c];
  }
  b[r] -= s;
}
// column-major: outer = column; only fixed columns contribute
__global__ void k_rhs_cols(int64_t n, const int32_t* __restrict__ outer, const int32_t* __restrict__ inner,
                           const double* __restrict__ values, const uint8_t* __restrict__ fixed, const double* __restrict__ xhat,
                           double* __restrict__ b) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n || !fixed[c]) return;
  const double x = xhat[c];
  for (int32_t k = outer[c]; k < outer[c + 1]; ++k) atomicAdd(b + inner[k], -values[k] * x);
}
__global__ void k_rhs_set(int64_t n, const uint8_t* __restrict__ fixed, const double* __restrict__ xhat, double* __restrict__ b) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r < n && fixed[r]) b[r] = xhat[r];
}
// one thread per outer index: rewrite values, flag survivors; missing diagonal of a fixed dof is an error
__global__ void k_mark(int64_t n, const int32_t* __restrict__ outer, const int32_t* __restrict__ inner, double* __restrict__ values,
                       const uint8_t* __restrict__ fixed, uint8_t* __restrict__ keep, int* __restrict__ flags, bool by_outer,
                       bool by_inner) {
  const int64_t o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (o >= n) return;
  const bool fo = fixed[o] != 0;
  bool diag_seen = false;
  for (int32_t k = outer[o]; k < outer[o + 1]; ++k) {
    const int32_t i = inner[k];
    const bool erased = (by_outer && fo) || (by_inner && fixed[i]);
    if (erased) {
      const bool diag = (i == o);
      values[k] = diag ? 1.0 : 0.0;
      if (keep) keep[k] = diag ? 1 : 0;
      diag_seen |= diag;
    } else if (keep) {
      keep[k] = 1;
    }
  }
  if (fo && !diag_seen) flags[0] = 1;
}
__global__ void k_new_outer(int64_t n_plus_1, const int32_t* __restrict__ outer, const int32_t* __restrict__ keep_scan, int64_t nnz,
                            int32_t total_kept, int32_t* __restrict__ outer_out) {
  const int64_t o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (o >= n_plus_1) return;
  const int32_t k = outer[o];
  outer_out[o] = (k < nnz) ? keep_scan[k] : total_kept;
}
__global__ void k_widen(int64_t n, const uint8_t* __restrict__ in, int32_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// rows_only = FixFlaggedSolutionCompAlt (fix_dof.h:181-218): unit ROWS for the fixed dofs, columns and the other
// right-hand-side entries untouched.
int fix_impl(lfgpu_ctx* ctx, const lfgpu_pattern* p, double* d_values, double* d_rhs, const uint8_t* d_fixed,
             const double* d_fixed_values, int32_t* d_outer_out, int32_t* d_inner_out, double* d_values_out, int64_t* nnz_out,
             bool rows_only) {
  if (ctx == nullptr || p == nullptr || d_values == nullptr || d_rhs == nullptr || d_fixed == nullptr || d_fixed_values == nullptr)
    return LFGPU_ERR_INVALID;
  if (p->n_outer != p->n_inner) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "Matrix must be square!");  // fix_dof.h:90
  const bool compact = d_outer_out != nullptr || d_inner_out != nullptr || d_values_out != nullptr;
  if (compact && (d_outer_out == nullptr || d_inner_out == nullptr || d_values_out == nullptr || nnz_out == nullptr))
    LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "compaction needs all three output arrays and nnz_out");
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int64_t N = p->n_outer, nnz = p->nnz;
  const unsigned gn = static_cast<unsigned>(cdiv(N, kThreads));
  if (N == 0) {  // nothing to fix: no launch with an empty grid
    if (nnz_out != nullptr) *nnz_out = 0;
    return LFGPU_OK;
  }
  // 1. right-hand side
  if (!rows_only) {
    if (p->major == LFGPU_ROW_MAJOR) {
      k_rhs_rows<<<gn, kThreads, 0, st>>>(N, p->outer, p->inner, d_values, d_fixed, d_fixed_values, d_rhs);
    } else {
      k_rhs_cols<<<gn, kThreads, 0, st>>>(N, p->outer, p->inner, d_values, d_fixed, d_fixed_values, d_rhs);
    }
    LFGPU_LAUNCH_CHECK(ctx);
  }
  k_rhs_set<<<gn, kThreads, 0, st>>>(N, d_fixed, d_fixed_values, d_rhs);
  LFGPU_LAUNCH_CHECK(ctx);
  // 2. matrix
  uint8_t* keep = nullptr;
  int32_t *keep32 = nullptr, *scan = nullptr;
  void* tmp = nullptr;
  int64_t* d_num = nullptr;
  auto cleanup = [&]() { cudaFree(keep); cudaFree(keep32); cudaFree(scan); cudaFree(tmp); cudaFree(d_num); };
  int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 512);
  cudaError_t e = cudaMemsetAsync(d_flags, 0, 16, st);
  if (e == cudaSuccess && compact) e = cudaMalloc(&keep, nnz > 0 ? nnz : 1);
  if (e != cudaSuccess) {
    cleanup();
    LFGPU_CUDA_CHECK(ctx, e);
  }
  const bool by_outer = !rows_only || p->major == LFGPU_ROW_MAJOR;  // outer index is a row
  const bool by_inner = !rows_only || p->major != LFGPU_ROW_MAJOR;  // inner index is a row
  k_mark<<<gn, kThreads, 0, st>>>(N, p->outer, p->inner, d_values, d_fixed, keep, d_flags, by_outer, by_inner);
  ctx->launches++;
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    cleanup();
    LFGPU_CUDA_CHECK(ctx, e);
  }
  int h_flags[2] = {0, 0};
  if (compact) {
    e = cudaMalloc(&keep32, sizeof(int32_t) * (nnz + 1));
    if (e == cudaSuccess) e = cudaMalloc(&scan, sizeof(int32_t) * (nnz + 1));
    if (e == cudaSuccess) e = cudaMalloc(&d_num, sizeof(int64_t));
    if (e == cudaSuccess) {
      if (nnz > 0) {
        k_widen<<<static_cast<unsigned>(cdiv(nnz, kThreads)), kThreads, 0, st>>>(nnz, keep, keep32);
        ctx->launches++;
        e = cudaGetLastError();
      }
      size_t tb = 0, tb2 = 0, tb3 = 0;
      if (e == cudaSuccess) cub::DeviceScan::ExclusiveSum(nullptr, tb, keep32, scan, nnz, st);
      cub::DeviceSelect::Flagged(nullptr, tb2, p->inner, keep, d_inner_out, d_num, nnz, st);
      cub::DeviceSelect::Flagged(nullptr, tb3, d_values, keep, d_values_out, d_num, nnz, st);
      tb = tb > tb2 ? tb : tb2;
      tb = tb > tb3 ? tb : tb3;
      if (e == cudaSuccess) e = cudaMalloc(&tmp, tb > 0 ? tb : 1);
      if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tb, keep32, scan, nnz, st);
      if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(tmp, tb, p->inner, keep, d_inner_out, d_num, nnz, st);
      if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(tmp, tb, d_values, keep, d_values_out, d_num, nnz, st);
      int64_t kept = 0;
      if (e == cudaSuccess) e = cudaMemcpyAsync(&kept, d_num, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e == cudaSuccess) {
        k_new_outer<<<static_cast<unsigned>(cdiv(N + 1, kThreads)), kThreads, 0, st>>>(N + 1, p->outer, scan, nnz, static_cast<int32_t>(kept), d_outer_out);
        ctx->launches++;
        e = cudaGetLastError();
        *nnz_out = kept;
      }
    }
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  LFGPU_CUDA_CHECK(ctx, e);
  if (h_flags[0]) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "a fixed dof has no diagonal entry in the pattern");
  return LFGPU_OK;
}
}  // namespace
}  // namespace lfgpu

extern "C" int lfgpu_fix_flagged_solution_components(lfgpu_ctx* ctx, const lfgpu_pattern* p, double* d_values, double* d_rhs,
                                                      const uint8_t* d_fixed, const double* d_fixed_values, int32_t* d_outer_out,
                                                      int32_t* d_inner_out, double* d_values_out, int64_t* nnz_out) {
  return lfgpu::fix_impl(ctx, p, d_values, d_rhs, d_fixed, d_fixed_values, d_outer_out, d_inner_out, d_values_out, nnz_out, false);
}
extern "C" int lfgpu_fix_flagged_solution_comp_alt(lfgpu_ctx* ctx, const lfgpu_pattern* p, double* d_values, double* d_rhs,
                                                    const uint8_t* d_fixed, const double* d_fixed_values, int32_t* d_outer_out,
                                                    int32_t* d_inner_out, double* d_values_out, int64_t* nnz_out) {
  return lfgpu::fix_impl(ctx, p, d_values, d_rhs, d_fixed, d_fixed_values, d_outer_out, d_inner_out, d_values_out, nnz_out, true);
}
