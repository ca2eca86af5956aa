// Dictionary coding of per-row plan words (product code; used by the P2 / P3 row-kernel plans).
//
// The row kernels read, for every matrix row, the slots of its columns inside the row (5 or 4 bits per entry, several 32-bit
// words per row).  On the meshes the reference's builders and refinement produce these words take a handful of distinct
// values (the relative order of the dof numbers around a vertex or an edge repeats), so the plan stores a 16-bit index into
// a table of the distinct word tuples instead of the words: the ncu captures of round 2 show both P2 kernels moving 1.23 x
// the algorithmic bytes, the difference being the plan.  Exact: the rows are ordered by their whole tuples
// (one stable radix sort per word, least significant word first), no hashing.  More than 65535 distinct tuples -> n_dict = -1,
// the caller keeps the uncoded plan.
#include <algorithm>
#include <string>

#include <cub/cub.cuh>

#include "lfgpu_internal.cuh"

namespace lfgpu {
namespace {

__global__ void k_iota(int64_t n, int32_t* __restrict__ v) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) v[i] = static_cast<int32_t>(i);
}

__global__ void k_gather_word(int64_t n, const uint32_t* __restrict__ word, const int32_t* __restrict__ perm, uint32_t* __restrict__ keys) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = word[perm[i]];
}

template <int W>
__global__ void k_heads(int64_t n, const uint32_t* __restrict__ words, const int32_t* __restrict__ perm, int32_t* __restrict__ head) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool differs = (i == 0);
  if (!differs) {
    const int32_t a = perm[i], b = perm[i - 1];
#pragma unroll
    for (int k = 0; k < W; ++k) differs = differs || (words[k * n + a] != words[k * n + b]);
  }
  head[i] = differs ? 1 : 0;
}

// rank[i] = inclusive sum of head: tuple number of sorted position i, 1-based
template <int W>
__global__ void k_assign(int64_t n, const uint32_t* __restrict__ words, const int32_t* __restrict__ perm, const int32_t* __restrict__ head,
                         const int32_t* __restrict__ rank, uint16_t* __restrict__ idx, uint4* __restrict__ dict) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t r = perm[i];
  const int32_t j = rank[i] - 1;
  idx[r] = static_cast<uint16_t>(j);
  if (head[i]) {
    uint32_t w[4] = {0U, 0U, 0U, 0U};
#pragma unroll
    for (int k = 0; k < W; ++k) w[k] = words[k * n + r];
    dict[j] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

template <int W>
int build_dict_w(lfgpu_ctx* ctx, int64_t n, const uint32_t* words, uint16_t* idx, uint4** dict_out, int* n_dict) {
  cudaStream_t st = ctx->stream;
  // perm / perm2 and keys / keys2 are the double buffers of the sorts; keys / keys2 serve as head / rank afterwards
  int32_t *perm = nullptr, *perm2 = nullptr;
  uint32_t *keys = nullptr, *keys2 = nullptr;
  void* tmp = nullptr;
  uint4* dict = nullptr;
  auto cleanup = [&]() { cudaFree(perm); cudaFree(perm2); cudaFree(keys); cudaFree(keys2); cudaFree(tmp); };
#define DICT_CHECK(expr)                                                          \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      set_last_error(ctx, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
      cleanup();                                                                  \
      cudaFree(dict);                                                             \
      return LFGPU_ERR_CUDA;                                                      \
    }                                                                             \
  } while (0)
  DICT_CHECK(cudaMalloc(&perm, sizeof(int32_t) * n));
  DICT_CHECK(cudaMalloc(&perm2, sizeof(int32_t) * n));
  DICT_CHECK(cudaMalloc(&keys, sizeof(uint32_t) * n));
  DICT_CHECK(cudaMalloc(&keys2, sizeof(uint32_t) * n));
  const unsigned grid = static_cast<unsigned>(cdiv(n, 256));
  k_iota<<<grid, 256, 0, st>>>(n, perm);
  ctx->launches++;
  size_t tb_sort = 0, tb_scan = 0;
  DICT_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb_sort, keys, keys2, perm, perm2, n, 0, 32, st));
  DICT_CHECK(cub::DeviceScan::InclusiveSum(nullptr, tb_scan, reinterpret_cast<int32_t*>(keys), reinterpret_cast<int32_t*>(keys2), n, st));
  const size_t tb = std::max<size_t>(std::max(tb_sort, tb_scan), 16);
  DICT_CHECK(cudaMalloc(&tmp, tb));
  for (int k = W - 1; k >= 0; --k) {  // stable sorts, least significant word first: perm ends up ordered by the whole tuple
    k_gather_word<<<grid, 256, 0, st>>>(n, words + static_cast<size_t>(k) * n, perm, keys);
    ctx->launches++;
    size_t tb_use = tb;
    DICT_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tb_use, keys, keys2, perm, perm2, n, 0, 32, st));
    std::swap(perm, perm2);
  }
  int32_t* head = reinterpret_cast<int32_t*>(keys);
  int32_t* rank = reinterpret_cast<int32_t*>(keys2);
  k_heads<W><<<grid, 256, 0, st>>>(n, words, perm, head);
  ctx->launches++;
  size_t tb_use = tb;
  DICT_CHECK(cub::DeviceScan::InclusiveSum(tmp, tb_use, head, rank, n, st));
  int32_t total = 0;
  DICT_CHECK(cudaMemcpyAsync(&total, rank + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  DICT_CHECK(cudaStreamSynchronize(st));
  if (total > 65535) {  // index 0xFFFF is the callers' mark for "not a planned row"
    cleanup();
    *n_dict = -1;
    *dict_out = nullptr;
    return LFGPU_OK;
  }
  DICT_CHECK(cudaMalloc(&dict, sizeof(uint4) * std::max<int32_t>(total, 1)));
  k_assign<W><<<grid, 256, 0, st>>>(n, words, perm, head, rank, idx, dict);
  ctx->launches++;
  DICT_CHECK(cudaGetLastError());
  DICT_CHECK(cudaStreamSynchronize(st));
#undef DICT_CHECK
  cleanup();
  *n_dict = total;
  *dict_out = dict;
  return LFGPU_OK;
}

}  // namespace

// ---- node order of an edge plan ---------------------------------------------------------------------------------------------
// The edge-row kernels gather four node positions per row.  The reference's mesh builders number nodes row by row but edges (and
// cells) column by column, so the 32 consecutive edge rows of a warp read from 32 different 128-byte lines per gather; with the
// nodes renumbered in the order the edge rows use them the same kernels ran 30 % (P2) / 11 % (P3) faster on B200
// (profiles/r02_locality_p2.json).  The matrix keeps the reference's numbering -- only the kernels' private copy of the
// coordinates is stored in "first use" order: rank of a node = the first edge row that touches it.
namespace {
__global__ void k_first_use(int64_t ne, const int32_t* __restrict__ enb, uint32_t* __restrict__ key) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const int32_t p = enb[e];
  if (p < 0) return;  // not a planned row
  // rank = first row that uses the node in ANY role (P before Q before o inside a row).  (Ranking by first use as P alone gives the
  // same locality on the builder's numbering -- 20.5 against 20.9 lines per warp -- but sends the nodes that are never a P to the
  // end of the order, and the compact plan's 16-bit differences then overflow.)
  for (int k = 0; k < 4; ++k) atomicMin(key + enb[k * ne + e], static_cast<uint32_t>(4 * e + k));
}
// locality of the gathers of consecutive rows: how often a row's node k lies in another 128-byte line (8 positions) than the
// previous planned row's node k.  map == nullptr: the numbers as they are.
__global__ void k_line_changes(int64_t ne, const int32_t* __restrict__ enb, const uint32_t* __restrict__ map, unsigned long long* __restrict__ count) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int c = 0;
  if (e > 0 && e < ne && enb[e] >= 0 && enb[e - 1] >= 0) {
    for (int k = 0; k < 4; ++k) {
      uint32_t a = static_cast<uint32_t>(enb[k * ne + e]), b = static_cast<uint32_t>(enb[k * ne + e - 1]);
      if (map != nullptr) {
        a = map[a];
        b = map[b];
      }
      c += (a >> 3) != (b >> 3);
    }
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffU, c, o);
  if ((threadIdx.x & 31) == 0 && c > 0) atomicAdd(count, static_cast<unsigned long long>(c));
}
__global__ void k_rank_of(int64_t n, const int32_t* __restrict__ order, uint32_t* __restrict__ new_id) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) new_id[order[i]] = static_cast<uint32_t>(i);
}
__global__ void k_remap_ids(int64_t ne, int32_t* __restrict__ enb, const uint32_t* __restrict__ new_id) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= ne || enb[e] < 0) return;
  for (int k = 0; k < 4; ++k) enb[k * ne + e] = static_cast<int32_t>(new_id[enb[k * ne + e]]);
}
__global__ void k_permute_xy(int64_t n, const uint32_t* __restrict__ new_id, const double2* __restrict__ xy, double2* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[new_id[i]] = xy[i];
}
}  // namespace

// enb: device [4][ne] node numbers of an edge plan (P, Q, o_1, o_2; P < 0: row not planned), rewritten to the new numbers.
// *new_id_out: device [nn] (cudaMalloc, the caller frees): position of every node in the kernels' coordinate copy.
int edge_node_order(lfgpu_ctx* ctx, int64_t nn, int64_t ne, int32_t* enb, uint32_t** new_id_out) {
  *new_id_out = nullptr;
  if (nn < 8 || ne <= 0 || ne >= (1LL << 30) || nn >= (1LL << 31)) return LFGPU_OK;
  cudaStream_t st = ctx->stream;
  uint32_t *key = nullptr, *key2 = nullptr, *new_id = nullptr;
  int32_t *ids = nullptr, *ids2 = nullptr;
  void* tmp = nullptr;
  auto cleanup = [&]() { cudaFree(key); cudaFree(key2); cudaFree(ids); cudaFree(ids2); cudaFree(tmp); };
#define ORD_CHECK(expr)                                                           \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      set_last_error(ctx, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
      cleanup();                                                                  \
      cudaFree(new_id);                                                           \
      return LFGPU_ERR_CUDA;                                                      \
    }                                                                             \
  } while (0)
  ORD_CHECK(cudaMalloc(&key, sizeof(uint32_t) * nn));
  ORD_CHECK(cudaMalloc(&key2, sizeof(uint32_t) * nn));
  ORD_CHECK(cudaMalloc(&ids, sizeof(int32_t) * nn));
  ORD_CHECK(cudaMalloc(&ids2, sizeof(int32_t) * nn));
  ORD_CHECK(cudaMalloc(&new_id, sizeof(uint32_t) * nn));
  ORD_CHECK(cudaMemsetAsync(key, 0xFF, sizeof(uint32_t) * nn, st));  // nodes no planned row uses come last, in their old order
  k_first_use<<<static_cast<unsigned>(cdiv(ne, 256)), 256, 0, st>>>(ne, enb, key);
  ctx->launches++;
  k_iota<<<static_cast<unsigned>(cdiv(nn, 256)), 256, 0, st>>>(nn, ids);
  ctx->launches++;
  size_t tb = 0;
  ORD_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tb, key, key2, ids, ids2, nn, 0, 32, st));
  ORD_CHECK(cudaMalloc(&tmp, std::max<size_t>(tb, 16)));
  ORD_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tb, key, key2, ids, ids2, nn, 0, 32, st));  // stable
  k_rank_of<<<static_cast<unsigned>(cdiv(nn, 256)), 256, 0, st>>>(nn, ids2, new_id);
  ctx->launches++;
  // adopted only where it pays: at least a quarter fewer line changes between consecutive rows than the mesh's own numbering has
  // (builder meshes: 4.0 -> 0.6 per row; MeshHierarchy-refined and Morton-ordered meshes are as local as they get: measured 1 % either
  // way on config C4's mesh, 4 % slower on the Delaunay-numbered workload u2 -- those keep the mesh's array, and the memory)
  unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(key);  // key is sorted out: reuse its first 16 bytes
  unsigned long long h_cnt[2] = {0ULL, 0ULL};
  ORD_CHECK(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), st));
  k_line_changes<<<static_cast<unsigned>(cdiv(ne, 256)), 256, 0, st>>>(ne, enb, nullptr, d_cnt);
  k_line_changes<<<static_cast<unsigned>(cdiv(ne, 256)), 256, 0, st>>>(ne, enb, new_id, d_cnt + 1);
  ctx->launches += 2;
  ORD_CHECK(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
  ORD_CHECK(cudaStreamSynchronize(st));
  static const bool force = [] { const char* e = std::getenv("LFGPU_EDGE_ORDER"); return e != nullptr && e[0] == '2'; }();
  if (!force && 4 * h_cnt[1] > 3 * h_cnt[0]) {
    cleanup();
    cudaFree(new_id);
    return LFGPU_OK;  // *new_id_out stays null: the plan keeps the mesh's numbers
  }
  k_remap_ids<<<static_cast<unsigned>(cdiv(ne, 256)), 256, 0, st>>>(ne, enb, new_id);
  ctx->launches++;
  ORD_CHECK(cudaGetLastError());
  ORD_CHECK(cudaStreamSynchronize(st));
#undef ORD_CHECK
  cleanup();
  *new_id_out = new_id;
  return LFGPU_OK;
}

// xy_perm[new_id[i]] = xy[i] on the context stream
int permute_node_coords(lfgpu_ctx* ctx, int64_t nn, const uint32_t* new_id, const double* xy, double* xy_perm) {
  k_permute_xy<<<static_cast<unsigned>(cdiv(nn, 256)), 256, 0, ctx->stream>>>(nn, new_id, reinterpret_cast<const double2*>(xy),
                                                                             reinterpret_cast<double2*>(xy_perm));
  LFGPU_LAUNCH_CHECK(ctx);
  return LFGPU_OK;
}

// words: device [n_words][n] (slot-major), n_words in 1..4.  idx: device [n], receives the tuple number of every row; *dict_out:
// device uint4 [*n_dict] (cudaMalloc, the caller frees), unused components zero.  *n_dict = -1: too many distinct tuples, idx untouched.
int build_row_dict(lfgpu_ctx* ctx, int n_words, int64_t n, const uint32_t* words, uint16_t* idx, void** dict_out, int* n_dict) {
  *dict_out = nullptr;
  *n_dict = -1;
  if (n <= 0 || n >= (1LL << 31)) return LFGPU_OK;
  uint4* dict = nullptr;
  int rc = LFGPU_ERR_INVALID;
  switch (n_words) {
    case 1: rc = build_dict_w<1>(ctx, n, words, idx, &dict, n_dict); break;
    case 2: rc = build_dict_w<2>(ctx, n, words, idx, &dict, n_dict); break;
    case 3: rc = build_dict_w<3>(ctx, n, words, idx, &dict, n_dict); break;
    case 4: rc = build_dict_w<4>(ctx, n, words, idx, &dict, n_dict); break;
    default: break;
  }
  *dict_out = dict;
  return rc;
}

}  // namespace lfgpu
