// Symbolic pass on the device (product code): sparsity pattern, per-row gather lists, per-cell scatter map.
//
// Stands in for the structure that  COOMatrix::AddToEntry (lib/lf/assemble/coomatrix.h:87-91)  +
// COOMatrix::makeSparse -> Eigen::SparseMatrix::setFromTriplets (coomatrix.h:172-180)  produce for the triplets
// emitted by AssembleMatrixLocally (assembler.h:166-179): one stored entry per distinct (row dof, col dof) pair that
// shares a cell (explicit zeros kept), inner indices ascending per outer index, int32 indices.
//
// "outer" = the compressed dimension: trial (column) dofs for LFGPU_COL_MAJOR (Eigen's default), test (row) dofs for
// LFGPU_ROW_MAJOR (CSR).  Pipeline (all CUB radix sorts / scans, no host loop):
//   1. items (outer dof, cell<<4 | local index) stably sorted by outer dof -> gather lists, cells ascending per dof,
//      i.e. the reference's summation order
//   2. every item expands to its cell's inner dofs -> keys (outer<<32 | inner) -> sort -> unique -> pattern
//   3. every (cell, local outer a, local inner b) finds its slot in the outer segment by binary search -> scatter map
#include <cub/cub.cuh>

#include "lfgpu_internal.cuh"

namespace lfgpu {
namespace {
constexpr int kThreads = 256;

__global__ void k_make_items(int64_t n_cells, int stride, const int32_t* __restrict__ dofs, const uint8_t* __restrict__ nldof,
                             int32_t invalid_key, int32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n_cells * stride) return;
  const int64_t c = t / stride;
  const int a = static_cast<int>(t % stride);
  const bool used = a < nldof[c];
  keys[t] = used ? dofs[t] : invalid_key;
  vals[t] = (static_cast<uint32_t>(c) << 4) | static_cast<uint32_t>(a);
}

// ptr[r] = first position with key >= r, r = 0..n  (keys sorted ascending)
template <typename K>
__global__ void k_lower_bounds(int64_t n_rows_plus_1, int64_t n_keys, const K* __restrict__ keys, int shift,
                               int32_t* __restrict__ ptr, int* __restrict__ max_len) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_rows_plus_1) return;
  const K target = static_cast<K>(r) << shift;
  int64_t lo = 0, hi = n_keys;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  ptr[r] = static_cast<int32_t>(lo);
}

__global__ void k_max_diff(int64_t n, const int32_t* __restrict__ ptr, int* __restrict__ max_len) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int len = (r < n) ? (ptr[r + 1] - ptr[r]) : 0;
  len = __reduce_max_sync(0xffffffffU, len);
  if ((threadIdx.x & 31) == 0) atomicMax(max_len, len);
}

// longest value range owned by a block of `block` consecutive outer indices
__global__ void k_max_block_nnz(int64_t n, int block, const int32_t* __restrict__ ptr, int* __restrict__ out) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int len = 0;
  if (b * block < n) {
    const int64_t e = (b + 1) * block < n ? (b + 1) * block : n;
    len = ptr[e] - ptr[b * block];
  }
  len = __reduce_max_sync(0xffffffffU, len);
  if ((threadIdx.x & 31) == 0) atomicMax(out, len);
}

// item-parallel plan: first outer index of block b = first row whose first item is >= b * items_per_block
__global__ void k_block_rows(int64_t n_blocks, int64_t n_rows, int items_per_block, const int32_t* __restrict__ adj_ptr,
                             int32_t* __restrict__ blk_rows) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b > n_blocks) return;
  if (b == n_blocks) {
    blk_rows[b] = static_cast<int32_t>(n_rows);
    return;
  }
  const int64_t target = b * items_per_block;
  int64_t lo = 0, hi = n_rows;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (adj_ptr[mid] < target) lo = mid + 1; else hi = mid;
  }
  blk_rows[b] = static_cast<int32_t>(lo);
}
__global__ void k_block_nnz(int64_t n_blocks, const int32_t* __restrict__ blk_rows, const int32_t* __restrict__ outer,
                            const int32_t* __restrict__ adj_ptr, int* __restrict__ out /*[0] max nnz, [1] max items, [2] max rows*/) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int nnz = 0, items = 0, rows = 0;
  if (b < n_blocks) {
    nnz = outer[blk_rows[b + 1]] - outer[blk_rows[b]];
    items = adj_ptr[blk_rows[b + 1]] - adj_ptr[blk_rows[b]];
    rows = blk_rows[b + 1] - blk_rows[b];
  }
  nnz = __reduce_max_sync(0xffffffffU, nnz);
  items = __reduce_max_sync(0xffffffffU, items);
  rows = __reduce_max_sync(0xffffffffU, rows);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, nnz);
    atomicMax(out + 1, items);
    atomicMax(out + 2, rows);
  }
}
// thread order of the item kernel inside a block: items sorted by (rank within their dof, local index, dof), so that in
// the k-th accumulation round the active threads are a contiguous range (whole warps work or idle)
__global__ void __launch_bounds__(kItemThreads) k_item_perm(const int32_t* __restrict__ blk_rows, const int32_t* __restrict__ adj_ptr,
                                                   const int32_t* __restrict__ outer, const uint32_t* __restrict__ adj,
                                                   uint32_t* __restrict__ item_perm, uint2* __restrict__ item_sorted,
                                                   int4* __restrict__ blk_hdr) {
  using Sort = cub::BlockRadixSort<uint32_t, kItemThreads, 1, uint32_t>;
  __shared__ typename Sort::TempStorage tmp;
  __shared__ int32_t s_adj[kItemThreads + 1];
  __shared__ int s_max_rank;
  const int tid = threadIdx.x;
  if (tid == 0) s_max_rank = 0;
  const int32_t R0 = blk_rows[blockIdx.x], R1 = blk_rows[blockIdx.x + 1];
  const int nrows = R1 - R0;
  const int32_t adj0 = adj_ptr[R0];
  for (int j = tid; j <= nrows; j += kItemThreads) s_adj[j] = adj_ptr[R0 + j] - adj0;
  __syncthreads();
  const int n_items = s_adj[nrows];
  uint32_t key[1] = {0xffffffffU}, val[1] = {0};
  if (tid < n_items) {
    int lo = 0, hi = nrows;
    while (hi - lo > 1) {
      const int mid = lo + ((hi - lo) >> 1);
      if (s_adj[mid] <= tid) lo = mid; else hi = mid;
    }
    const uint32_t rank = static_cast<uint32_t>(tid - s_adj[lo]);
    // (rank, local index a, dof): same-rank items belong to different dofs, so their order is free -- grouping them by
    // the local index makes the reference-tensor reads of a warp hit one table row (shared-memory broadcast)
    const uint32_t a = adj[adj0 + tid] & 15U;
    key[0] = (rank << 12) | (a << 8) | static_cast<uint32_t>(lo);
    val[0] = static_cast<uint32_t>(tid) | (static_cast<uint32_t>(lo) << 8) | (rank << 16);
    atomicMax(&s_max_rank, static_cast<int>(rank));
  }
  Sort(tmp).Sort(key, val, 0, 20);
  const int32_t out0 = outer[R0];
  if (tid < n_items) {
    item_perm[adj0 + tid] = val[0];
    // thread-ordered copy of the items: (cell << 4 | a, offset of the dof's segment in the block image | rank << 16)
    // -- one coalesced 8-byte load per thread, nothing else to look up
    const uint32_t lo = (val[0] >> 8) & 255U;
    const uint32_t off = static_cast<uint32_t>(outer[R0 + lo] - out0);
    item_sorted[adj0 + tid] = make_uint2(adj[adj0 + (val[0] & 255U)], off | ((val[0] >> 16) << 16));
  }
  __syncthreads();
  // block header: everything the item kernel needs to start, in one 16-byte load
  if (tid == 0) blk_hdr[blockIdx.x] = make_int4(adj0, out0, n_items | (s_max_rank << 16), outer[R1] - out0);
}

template <typename P>
__global__ void k_pos_by_item(int64_t n_items, int o_stride, int pos_row, const uint2* __restrict__ item_sorted, const P* __restrict__ pos,
                              P* __restrict__ pos_item) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n_items * pos_row) return;
  const int64_t it = t / pos_row;
  const int b = static_cast<int>(t - it * pos_row);
  const uint32_t item = item_sorted[it].x;
  pos_item[t] = pos[(static_cast<int64_t>(item >> 4) * o_stride + (item & 15U)) * pos_row + b];
}

__global__ void k_item_counts(int64_t n_items, const uint32_t* __restrict__ adj, const uint8_t* __restrict__ i_nldof,
                              int64_t* __restrict__ counts) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n_items) return;
  counts[t] = i_nldof[adj[t] >> 4];
}

__global__ void k_expand(int64_t n_items, const int32_t* __restrict__ item_keys, const uint32_t* __restrict__ adj,
                         const int64_t* __restrict__ offsets, int i_stride, const int32_t* __restrict__ i_dofs,
                         const uint8_t* __restrict__ i_nldof, uint64_t* __restrict__ out) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n_items) return;
  const int64_t c = adj[t] >> 4;
  const uint64_t hi = static_cast<uint64_t>(static_cast<uint32_t>(item_keys[t])) << 32;
  const int n = i_nldof[c];
  uint64_t* o = out + offsets[t];
  for (int b = 0; b < n; ++b) o[b] = hi | static_cast<uint32_t>(i_dofs[c * i_stride + b]);
}

__global__ void k_low32(int64_t n, const uint64_t* __restrict__ keys, int32_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<int32_t>(keys[i] & 0xffffffffULL);
}

template <typename P>
__global__ void k_positions(int64_t n_cells, int o_stride, int i_stride, int pos_row, const int32_t* __restrict__ o_dofs,
                            const uint8_t* __restrict__ o_nldof, const int32_t* __restrict__ i_dofs,
                            const uint8_t* __restrict__ i_nldof, const int32_t* __restrict__ outer,
                            const int32_t* __restrict__ inner, P* __restrict__ pos, int* __restrict__ flags) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n_cells * o_stride) return;
  const int64_t c = t / o_stride;
  const int a = static_cast<int>(t % o_stride);
  P* out = pos + t * pos_row;
  if (a >= o_nldof[c]) {
    for (int b = 0; b < pos_row; ++b) out[b] = 0;
    return;
  }
  const int32_t r = o_dofs[c * o_stride + a];
  const int32_t s = outer[r], e = outer[r + 1];
  const int n = i_nldof[c];
  for (int b = 0; b < pos_row; ++b) {
    P p = 0;
    if (b < n) {
      const int32_t target = i_dofs[c * i_stride + b];
      int32_t lo = s, hi = e;
      while (lo < hi) {
        const int32_t mid = lo + ((hi - lo) >> 1);  // lo + hi overflows int32 once a pattern holds more than 2^30 values
        if (inner[mid] < target) lo = mid + 1; else hi = mid;
      }
      if (lo >= e || inner[lo] != target) flags[0] = 1;
      p = static_cast<P>(lo - s);
    }
    out[b] = p;
  }
}

}  // namespace
}  // namespace lfgpu

using namespace lfgpu;

extern "C" {

void lfgpu_pattern_destroy(lfgpu_pattern* p) {
  if (p == nullptr) return;
  if (p->ctx) {
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
  }
  cudaFree(p->outer);
  cudaFree(p->inner);
  cudaFree(p->adj_ptr);
  cudaFree(p->adj);
  cudaFree(p->pos);
  cudaFree(p->blk_rows);
  cudaFree(p->pos_item);
  cudaFree(p->item_perm);
  cudaFree(p->item_sorted);
  cudaFree(p->blk_hdr);
  cudaFree(p->cell_metric);
  cudaFree(p->row_keep);
  cudaFree(p->fan_nbr);
  cudaFree(p->fan_rowinfo);
  cudaFree(p->fan_nbr16);
  cudaFree(p->fan_info32);
  cudaFree(p->fan_irregular);
  cudaFree(p->p1h_qw);
  cudaFree(p->p1h_tw);
  cudaFree(p->p1h_rowinfo);
  cudaFree(p->p1h_irregular);
  cudaFree(p->p2v_nbr);
  cudaFree(p->p2v_cidx);
  cudaFree(p->p2e_newid);
  cudaFree(p->p2e_xy);
  cudaFree(p->p3e_newid);
  cudaFree(p->p3e_xy);
  cudaFree(p->p2v_slots);
  cudaFree(p->p2e_nbr);
  cudaFree(p->p2e_slots);
  cudaFree(p->p2_irregular);
  cudaFree(p->p2g_nbr);
  cudaFree(p->p2g_slots);
  cudaFree(p->p3v_nbr);
  cudaFree(p->p3v_slots);
  cudaFree(p->p3e_nbr);
  cudaFree(p->p3e_slots);
  cudaFree(p->p3c_slots);
  cudaFree(p->p3_irregular);
  cudaFree(p->p3g_nbr);
  cudaFree(p->p3g_slots);
  cudaFree(p->o_dofs);
  if (p->i_dofs != p->o_dofs) cudaFree(p->i_dofs);
  cudaFree(p->o_nldof);
  if (p->i_nldof != p->o_nldof) cudaFree(p->i_nldof);
  delete p;
}

int64_t lfgpu_pattern_nnz(const lfgpu_pattern* p) { return p ? p->nnz : -1; }
int64_t lfgpu_pattern_rows(const lfgpu_pattern* p) { return p ? (p->major == LFGPU_ROW_MAJOR ? p->n_outer : p->n_inner) : -1; }
int64_t lfgpu_pattern_cols(const lfgpu_pattern* p) { return p ? (p->major == LFGPU_ROW_MAJOR ? p->n_inner : p->n_outer) : -1; }
const int32_t* lfgpu_pattern_outer_device(const lfgpu_pattern* p) { return p ? p->outer : nullptr; }
const int32_t* lfgpu_pattern_inner_device(const lfgpu_pattern* p) { return p ? p->inner : nullptr; }

int lfgpu_pattern_download(lfgpu_ctx* ctx, const lfgpu_pattern* p, int32_t* outer, int32_t* inner) {
  if (ctx == nullptr || p == nullptr) return LFGPU_ERR_INVALID;
  if (outer) LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(outer, p->outer, sizeof(int32_t) * (p->n_outer + 1), cudaMemcpyDeviceToHost, ctx->stream));
  if (inner) LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(inner, p->inner, sizeof(int32_t) * p->nnz, cudaMemcpyDeviceToHost, ctx->stream));
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return LFGPU_OK;
}

int lfgpu_symbolic(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_dofmap* test, const lfgpu_dofmap* trial, int major,
                   lfgpu_pattern** out) {
  if (ctx == nullptr || mesh == nullptr || test == nullptr || trial == nullptr || out == nullptr) return LFGPU_ERR_INVALID;
  *out = nullptr;
  if (major != LFGPU_COL_MAJOR && major != LFGPU_ROW_MAJOR) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "major must be LFGPU_COL_MAJOR or LFGPU_ROW_MAJOR");
  if (test->n_cells != mesh->n_cells || trial->n_cells != mesh->n_cells)
    LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "Trial and test space must be defined on the same mesh");  // assembler.h:121-122
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const lfgpu_dofmap* O = (major == LFGPU_ROW_MAJOR) ? test : trial;
  const lfgpu_dofmap* I = (major == LFGPU_ROW_MAJOR) ? trial : test;
  const int64_t n_cells = mesh->n_cells;

  auto* p = new lfgpu_pattern;
  p->ctx = ctx;
  p->major = major;
  p->n_outer = O->n_dofs;
  p->n_inner = I->n_dofs;
  p->n_cells = n_cells;
  p->o_stride = O->stride;
  p->i_stride = I->stride;

  int32_t *keys_in = nullptr, *keys_out = nullptr;
  uint32_t *vals_in = nullptr;
  int64_t *counts = nullptr, *offsets = nullptr;
  uint64_t *ck_in = nullptr, *ck_out = nullptr;
  void* tmp = nullptr;
  int64_t* d_num = nullptr;
  auto cleanup = [&]() {
    cudaFree(keys_in); cudaFree(keys_out); cudaFree(vals_in); cudaFree(counts); cudaFree(offsets);
    cudaFree(ck_in); cudaFree(ck_out); cudaFree(tmp); cudaFree(d_num);
  };
#define SYM_CHECK(expr)                                                             \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      set_last_error(ctx, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
      cleanup();                                                                    \
      lfgpu_pattern_destroy(p);                                                     \
      return LFGPU_ERR_CUDA;                                                        \
    }                                                                               \
  } while (0)

  // own copies of the dof tables (the pattern outlives the dofmaps)
  SYM_CHECK(cudaMalloc(&p->o_dofs, sizeof(int32_t) * n_cells * O->stride));
  SYM_CHECK(cudaMalloc(&p->o_nldof, n_cells));
  SYM_CHECK(cudaMemcpyAsync(p->o_dofs, O->cell_dofs, sizeof(int32_t) * n_cells * O->stride, cudaMemcpyDeviceToDevice, st));
  SYM_CHECK(cudaMemcpyAsync(p->o_nldof, O->n_ldof, n_cells, cudaMemcpyDeviceToDevice, st));
  if (I == O) {
    p->i_dofs = p->o_dofs;
    p->i_nldof = p->o_nldof;
  } else {
    SYM_CHECK(cudaMalloc(&p->i_dofs, sizeof(int32_t) * n_cells * I->stride));
    SYM_CHECK(cudaMalloc(&p->i_nldof, n_cells));
    SYM_CHECK(cudaMemcpyAsync(p->i_dofs, I->cell_dofs, sizeof(int32_t) * n_cells * I->stride, cudaMemcpyDeviceToDevice, st));
    SYM_CHECK(cudaMemcpyAsync(p->i_nldof, I->n_ldof, n_cells, cudaMemcpyDeviceToDevice, st));
  }

  // ---- 1. gather lists ------------------------------------------------------------------------------------------
  const int64_t n_slots = n_cells * O->stride;
  SYM_CHECK(cudaMalloc(&keys_in, sizeof(int32_t) * n_slots));
  SYM_CHECK(cudaMalloc(&keys_out, sizeof(int32_t) * n_slots));
  SYM_CHECK(cudaMalloc(&vals_in, sizeof(uint32_t) * n_slots));
  SYM_CHECK(cudaMalloc(&p->adj, sizeof(uint32_t) * n_slots));
  k_make_items<<<static_cast<unsigned>(cdiv(n_slots, kThreads)), kThreads, 0, st>>>(n_cells, O->stride, p->o_dofs, p->o_nldof,
                                                                                    static_cast<int32_t>(p->n_outer), keys_in, vals_in);
  ctx->launches++;
  int key_bits = 1;
  while ((1LL << key_bits) <= p->n_outer) ++key_bits;
  size_t tb = 0, tb_max = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, keys_in, keys_out, vals_in, p->adj, n_slots, 0, key_bits, st);
  tb_max = tb;
  SYM_CHECK(cudaMalloc(&tmp, tb_max));
  SYM_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tb_max, keys_in, keys_out, vals_in, p->adj, n_slots, 0, key_bits, st));
  SYM_CHECK(cudaMalloc(&p->adj_ptr, sizeof(int32_t) * (p->n_outer + 1)));
  int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 64);
  SYM_CHECK(cudaMemsetAsync(d_flags, 0, 64, st));
  k_lower_bounds<int32_t><<<static_cast<unsigned>(cdiv(p->n_outer + 1, kThreads)), kThreads, 0, st>>>(p->n_outer + 1, n_slots, keys_out, 0, p->adj_ptr, nullptr);
  k_max_diff<<<static_cast<unsigned>(cdiv(p->n_outer, kThreads)), kThreads, 0, st>>>(p->n_outer, p->adj_ptr, d_flags + 4);
  ctx->launches += 2;
  int32_t n_items32 = 0;
  SYM_CHECK(cudaMemcpyAsync(&n_items32, p->adj_ptr + p->n_outer, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  SYM_CHECK(cudaStreamSynchronize(st));
  p->n_items = n_items32;
  cudaFree(keys_in); keys_in = nullptr;
  cudaFree(vals_in); vals_in = nullptr;

  // ---- 2. pattern -----------------------------------------------------------------------------------------------
  SYM_CHECK(cudaMalloc(&counts, sizeof(int64_t) * (p->n_items + 1)));
  SYM_CHECK(cudaMalloc(&offsets, sizeof(int64_t) * (p->n_items + 1)));
  SYM_CHECK(cudaMemsetAsync(counts + p->n_items, 0, sizeof(int64_t), st));
  k_item_counts<<<static_cast<unsigned>(cdiv(p->n_items, kThreads)), kThreads, 0, st>>>(p->n_items, p->adj, p->i_nldof, counts);
  ctx->launches++;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, counts, offsets, p->n_items + 1, st);
  if (tb > tb_max) {
    cudaFree(tmp); tmp = nullptr;
    tb_max = tb;
    SYM_CHECK(cudaMalloc(&tmp, tb_max));
  }
  SYM_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tb_max, counts, offsets, p->n_items + 1, st));
  int64_t n_cand = 0;
  SYM_CHECK(cudaMemcpyAsync(&n_cand, offsets + p->n_items, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  SYM_CHECK(cudaStreamSynchronize(st));
  cudaFree(counts); counts = nullptr;
  SYM_CHECK(cudaMalloc(&ck_in, sizeof(uint64_t) * n_cand));
  SYM_CHECK(cudaMalloc(&ck_out, sizeof(uint64_t) * n_cand));
  k_expand<<<static_cast<unsigned>(cdiv(p->n_items, kThreads)), kThreads, 0, st>>>(p->n_items, keys_out, p->adj, offsets, I->stride,
                                                                                   p->i_dofs, p->i_nldof, ck_in);
  ctx->launches++;
  cudaFree(offsets); offsets = nullptr;
  cudaFree(keys_out); keys_out = nullptr;
  cub::DeviceRadixSort::SortKeys(nullptr, tb, ck_in, ck_out, n_cand, 0, 32 + key_bits, st);
  if (tb > tb_max) {
    cudaFree(tmp); tmp = nullptr;
    tb_max = tb;
    SYM_CHECK(cudaMalloc(&tmp, tb_max));
  }
  SYM_CHECK(cub::DeviceRadixSort::SortKeys(tmp, tb_max, ck_in, ck_out, n_cand, 0, 32 + key_bits, st));
  SYM_CHECK(cudaMalloc(&d_num, sizeof(int64_t)));
  cub::DeviceSelect::Unique(nullptr, tb, ck_out, ck_in, d_num, n_cand, st);
  if (tb > tb_max) {
    cudaFree(tmp); tmp = nullptr;
    tb_max = tb;
    SYM_CHECK(cudaMalloc(&tmp, tb_max));
  }
  SYM_CHECK(cub::DeviceSelect::Unique(tmp, tb_max, ck_out, ck_in, d_num, n_cand, st));
  int64_t nnz = 0;
  SYM_CHECK(cudaMemcpyAsync(&nnz, d_num, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  SYM_CHECK(cudaStreamSynchronize(st));
  cudaFree(ck_out); ck_out = nullptr;
  if (nnz >= (1LL << 31)) {
    cleanup();
    lfgpu_pattern_destroy(p);
    LFGPU_FAIL(ctx, LFGPU_ERR_OVERFLOW, "nnz does not fit the reference's int32 storage index (Eigen::SparseMatrix<double>)");
  }
  p->nnz = nnz;
  SYM_CHECK(cudaMalloc(&p->inner, sizeof(int32_t) * (nnz > 0 ? nnz : 1)));
  SYM_CHECK(cudaMalloc(&p->outer, sizeof(int32_t) * (p->n_outer + 1)));
  k_low32<<<static_cast<unsigned>(cdiv(nnz, kThreads)), kThreads, 0, st>>>(nnz, ck_in, p->inner);
  k_lower_bounds<uint64_t><<<static_cast<unsigned>(cdiv(p->n_outer + 1, kThreads)), kThreads, 0, st>>>(p->n_outer + 1, nnz, ck_in, 32, p->outer, nullptr);
  k_max_diff<<<static_cast<unsigned>(cdiv(p->n_outer, kThreads)), kThreads, 0, st>>>(p->n_outer, p->outer, d_flags + 5);
  k_max_block_nnz<<<static_cast<unsigned>(cdiv(cdiv(p->n_outer, 128), kThreads)), kThreads, 0, st>>>(p->n_outer, 128, p->outer, d_flags + 6);
  ctx->launches += 4;
  int h_flags[8] = {0};
  SYM_CHECK(cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
  SYM_CHECK(cudaStreamSynchronize(st));
  p->max_items = h_flags[4];
  p->max_row_len = h_flags[5];
  p->max_block_nnz = h_flags[6];
  cudaFree(ck_in); ck_in = nullptr;

  // ---- 3. scatter map -------------------------------------------------------------------------------------------
  p->pos_row = (I->stride + 3) & ~3;
  p->pos_bytes = (p->max_row_len <= 256) ? 1 : 2;
  SYM_CHECK(cudaMalloc(&p->pos, static_cast<size_t>(p->pos_bytes) * n_slots * p->pos_row));
  SYM_CHECK(cudaMemsetAsync(d_flags, 0, 16, st));
  if (p->pos_bytes == 1) {
    k_positions<uint8_t><<<static_cast<unsigned>(cdiv(n_slots, kThreads)), kThreads, 0, st>>>(
        n_cells, O->stride, I->stride, p->pos_row, p->o_dofs, p->o_nldof, p->i_dofs, p->i_nldof, p->outer, p->inner,
        static_cast<uint8_t*>(p->pos), d_flags);
  } else {
    k_positions<uint16_t><<<static_cast<unsigned>(cdiv(n_slots, kThreads)), kThreads, 0, st>>>(
        n_cells, O->stride, I->stride, p->pos_row, p->o_dofs, p->o_nldof, p->i_dofs, p->i_nldof, p->outer, p->inner,
        static_cast<uint16_t*>(p->pos), d_flags);
  }
  ctx->launches++;
  SYM_CHECK(cudaMemcpyAsync(h_flags, d_flags, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
  SYM_CHECK(cudaStreamSynchronize(st));
  // ---- 4. item-parallel plan -----------------------------------------------------------------------------------
  if (p->max_items >= 1 && p->max_items <= 32 && p->n_items > 0) {
    const int items_per_block = kItemThreads - p->max_items;
    p->n_item_blocks = cdiv(p->n_items, items_per_block);
    SYM_CHECK(cudaMalloc(&p->blk_rows, sizeof(int32_t) * (p->n_item_blocks + 1)));
    k_block_rows<<<static_cast<unsigned>(cdiv(p->n_item_blocks + 1, kThreads)), kThreads, 0, st>>>(p->n_item_blocks, p->n_outer, items_per_block,
                                                                                                 p->adj_ptr, p->blk_rows);
    int* d_blk = d_flags + 8;
    SYM_CHECK(cudaMemsetAsync(d_blk, 0, 16, st));
    k_block_nnz<<<static_cast<unsigned>(cdiv(p->n_item_blocks, kThreads)), kThreads, 0, st>>>(p->n_item_blocks, p->blk_rows, p->outer, p->adj_ptr, d_blk);
    SYM_CHECK(cudaMalloc(&p->item_perm, sizeof(uint32_t) * p->n_items));
    ctx->launches += 2;
    int h_blk[4] = {0, 0, 0, 0};
    SYM_CHECK(cudaMemcpyAsync(h_blk, d_blk, sizeof(h_blk), cudaMemcpyDeviceToHost, st));
    SYM_CHECK(cudaStreamSynchronize(st));
    p->max_item_block_nnz = h_blk[0];
    if (h_blk[1] > kItemThreads || h_blk[2] > kItemThreads || h_blk[0] >= 65536) {  // rows without items would break the bound: keep the row-parallel kernel
      cudaFree(p->blk_rows);
      cudaFree(p->item_perm);
      p->blk_rows = nullptr;
      p->item_perm = nullptr;
      p->n_item_blocks = 0;
    } else {
      SYM_CHECK(cudaMalloc(&p->item_sorted, sizeof(uint2) * p->n_items));
      SYM_CHECK(cudaMalloc(&p->blk_hdr, sizeof(int4) * p->n_item_blocks));
      k_item_perm<<<static_cast<unsigned>(p->n_item_blocks), kItemThreads, 0, st>>>(p->blk_rows, p->adj_ptr, p->outer, p->adj, p->item_perm,
                                                                           static_cast<uint2*>(p->item_sorted), static_cast<int4*>(p->blk_hdr));
      ctx->launches++;
    const int64_t n_pos = p->n_items * p->pos_row;
      SYM_CHECK(cudaMalloc(&p->pos_item, static_cast<size_t>(p->pos_bytes) * n_pos));
      if (p->pos_bytes == 1) {
        k_pos_by_item<uint8_t><<<static_cast<unsigned>(cdiv(n_pos, kThreads)), kThreads, 0, st>>>(p->n_items, O->stride, p->pos_row, static_cast<const uint2*>(p->item_sorted),
                                                                                                 static_cast<const uint8_t*>(p->pos), static_cast<uint8_t*>(p->pos_item));
      } else {
        k_pos_by_item<uint16_t><<<static_cast<unsigned>(cdiv(n_pos, kThreads)), kThreads, 0, st>>>(p->n_items, O->stride, p->pos_row, static_cast<const uint2*>(p->item_sorted),
                                                                                                  static_cast<const uint16_t*>(p->pos), static_cast<uint16_t*>(p->pos_item));
      }
      ctx->launches++;
      SYM_CHECK(cudaStreamSynchronize(st));
    }
  }
  cleanup();
#undef SYM_CHECK
  if (h_flags[0]) {
    lfgpu_pattern_destroy(p);
    LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "internal error: scatter slot not found in pattern");
  }
  if (p->max_row_len > 65536) {
    lfgpu_pattern_destroy(p);
    LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "row longer than 65536 entries");
  }
  *out = p;
  return LFGPU_OK;
}

}  // extern "C"
