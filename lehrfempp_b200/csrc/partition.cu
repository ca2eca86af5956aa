// Distributed ownership (product code): Morton partition of the cells, dof ownership, and extraction of one GPU's
// sub-problem -- its cells plus a one-cell halo -- as an ordinary lfgpu_mesh / lfgpu_dofmap pair with LOCAL indices.
//
// The reference is serial; BASELINE.json's north star partitions the cell loop of
// lf::assemble::AssembleMatrixLocally (assemble/assembler.h:125-180) over the GPUs of one node: "Morton-ordered cell ranges,
// each GPU owning the matrix rows for its cells".  What makes that cheap here: the local numbering of a sub-problem is the
// ORDER-PRESERVING restriction of the global one (local index = rank of the global index among the indices present), so
//   * the cells of a sub-problem keep their relative order -> the additions to a matrix entry happen in the reference's order,
//   * the columns of a local row, mapped through local -> global, are ascending -> the local pattern of an owned row IS the
//     global pattern of that row (bit-exact after the mapping; rows of halo dofs are incomplete and are never handed out),
//   * the dof layouts the row kernels rely on (node dofs first and equal to the node index, then edge dofs, then one interior
//     dof per cell in cell order) hold locally because they hold globally,
// hence symbolic pass, plans and numeric kernels run unchanged on the sub-problem, with int32 indices that only have to address
// 1/N of the matrix: the global number of stored values may exceed 2^31 (BASELINE config 4: 2.5e9).
#include <cub/cub.cuh>

#include "lfgpu_internal.cuh"

struct lfgpu_submesh {
  lfgpu_ctx* ctx = nullptr;
  lfgpu_mesh* mesh = nullptr;      // owned
  lfgpu_dofmap* dofmap = nullptr;  // owned
  int64_t n_cells = 0, n_nodes = 0, n_dofs = 0;
  int32_t* l2g_cells = nullptr;  // [n_cells] ascending
  int32_t* l2g_nodes = nullptr;  // [n_nodes] ascending
  int32_t* l2g_dofs = nullptr;   // [n_dofs] ascending
};

namespace lfgpu {
namespace {

constexpr int kThreads = 256;
constexpr uint32_t kNil = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t part1by1(uint32_t v) {
  v &= 0xFFFFu;
  v = (v | (v << 8)) & 0x00FF00FFu;
  v = (v | (v << 4)) & 0x0F0F0F0Fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}

// bounding box of the node positions: lo / hi as ordered integers (atomicMin / atomicMax on the bit patterns of doubles)
__device__ __forceinline__ unsigned long long ordered(double x) {
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(x));
  return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double unordered(unsigned long long o) {
  const unsigned long long b = (o & 0x8000000000000000ULL) ? (o & 0x7FFFFFFFFFFFFFFFULL) : ~o;
  return __longlong_as_double(static_cast<long long>(b));
}
__global__ void k_bbox(int64_t n_nodes, const double* __restrict__ xy, unsigned long long* __restrict__ box) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  unsigned long long lx = ~0ULL, ly = ~0ULL, hx = 0ULL, hy = 0ULL;
  if (i < n_nodes) {
    lx = hx = ordered(xy[2 * i]);
    ly = hy = ordered(xy[2 * i + 1]);
  }
  for (int d = 16; d > 0; d >>= 1) {
    lx = min(lx, __shfl_xor_sync(0xffffffffU, lx, d));
    ly = min(ly, __shfl_xor_sync(0xffffffffU, ly, d));
    hx = max(hx, __shfl_xor_sync(0xffffffffU, hx, d));
    hy = max(hy, __shfl_xor_sync(0xffffffffU, hy, d));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(box, lx);
    atomicMin(box + 1, ly);
    atomicMax(box + 2, hx);
    atomicMax(box + 3, hy);
  }
}

// Morton code of the cell centroid on a 65536 x 65536 grid over the bounding box, cell index as the tie-breaker
__global__ void k_morton(int64_t n_cells, const uint32_t* __restrict__ cell_nodes, const double* __restrict__ xy,
                         const unsigned long long* __restrict__ box, uint64_t* __restrict__ keys) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const double lx = unordered(box[0]), ly = unordered(box[1]), hx = unordered(box[2]), hy = unordered(box[3]);
  const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[c];
  const double2* p = reinterpret_cast<const double2*>(xy);
  double sx = p[v.x].x + p[v.y].x + p[v.z].x, sy = p[v.x].y + p[v.y].y + p[v.z].y, cnt = 3.0;
  if (v.w != kNil) {
    sx += p[v.w].x;
    sy += p[v.w].y;
    cnt = 4.0;
  }
  const double wx = hx - lx > 0.0 ? hx - lx : 1.0, wy = hy - ly > 0.0 ? hy - ly : 1.0;
  const double qx = fmin(fmax((sx / cnt - lx) / wx * 65535.0, 0.0), 65535.0), qy = fmin(fmax((sy / cnt - ly) / wy * 65535.0, 0.0), 65535.0);
  const uint32_t code = part1by1(static_cast<uint32_t>(qx)) | (part1by1(static_cast<uint32_t>(qy)) << 1);
  keys[c] = (static_cast<uint64_t>(code) << 32) | static_cast<uint64_t>(c);
}

// position in the sorted order -> part (equal cell counts), scattered back to the cell
__global__ void k_assign_parts(int64_t n_cells, const uint64_t* __restrict__ sorted, int64_t per, uint8_t* __restrict__ part) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_cells) return;
  part[sorted[i] & 0xFFFFFFFFULL] = static_cast<uint8_t>(i / per);
}

// a dof is owned by the LOWEST part with a cell touching it
__global__ void k_owner_min(int64_t n_cells, int stride, const int32_t* __restrict__ cell_dofs, const uint8_t* __restrict__ part,
                            uint32_t* __restrict__ owner32) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n_cells * stride) return;
  const int32_t d = cell_dofs[t];
  if (d >= 0) atomicMin(owner32 + d, static_cast<uint32_t>(part[t / stride]));
}
__global__ void k_narrow_owner(int64_t n, const uint32_t* __restrict__ owner32, uint8_t* __restrict__ owner) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) owner[i] = static_cast<uint8_t>(owner32[i] > 255U ? 255U : owner32[i]);
}

// sel[c] = 1 if cell c belongs to part `rank` (halo == 0) or touches a dof owned by `rank` (halo != 0)
__global__ void k_select_cells(int64_t n_cells, int stride, const int32_t* __restrict__ cell_dofs, const uint8_t* __restrict__ part,
                               const uint8_t* __restrict__ owner, int rank, int halo, uint8_t* __restrict__ sel) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  bool s = part[c] == rank;
  if (halo && !s) {
    for (int k = 0; k < stride; ++k) {
      const int32_t d = cell_dofs[c * stride + k];
      if (d >= 0 && owner[d] == rank) s = true;
    }
  }
  sel[c] = s ? 1 : 0;
}

// marks the nodes / dofs the selected cells refer to
__global__ void k_mark_used(int64_t n_sel, const int32_t* __restrict__ cells, int stride, const int32_t* __restrict__ cell_dofs,
                            const uint32_t* __restrict__ cell_nodes, uint8_t* __restrict__ dof_used, uint8_t* __restrict__ node_used) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_sel) return;
  const int64_t c = cells[i];
  for (int k = 0; k < stride; ++k) {
    const int32_t d = cell_dofs[c * stride + k];
    if (d >= 0) dof_used[d] = 1;
  }
  const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[c];
  node_used[v.x] = 1;
  node_used[v.y] = 1;
  node_used[v.z] = 1;
  if (v.w != kNil) node_used[v.w] = 1;
}

// the local tables: global indices replaced by their rank among the used ones (g2l = exclusive scan of the marks)
__global__ void k_local_tables(int64_t n_sel, const int32_t* __restrict__ cells, int stride, const int32_t* __restrict__ cell_dofs,
                               const uint8_t* __restrict__ n_ldof, const uint32_t* __restrict__ cell_nodes, const double* __restrict__ cell_coords,
                               const int32_t* __restrict__ g2l_dof, const int32_t* __restrict__ g2l_node, int32_t* __restrict__ l_dofs,
                               uint8_t* __restrict__ l_nldof, uint32_t* __restrict__ l_nodes, double* __restrict__ l_cell_coords) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_sel) return;
  const int64_t c = cells[i];
  for (int k = 0; k < stride; ++k) {
    const int32_t d = cell_dofs[c * stride + k];
    l_dofs[i * stride + k] = d >= 0 ? g2l_dof[d] : -1;
  }
  l_nldof[i] = n_ldof[c];
  const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[c];
  uint4 w;
  w.x = static_cast<uint32_t>(g2l_node[v.x]);
  w.y = static_cast<uint32_t>(g2l_node[v.y]);
  w.z = static_cast<uint32_t>(g2l_node[v.z]);
  w.w = v.w != kNil ? static_cast<uint32_t>(g2l_node[v.w]) : kNil;
  reinterpret_cast<uint4*>(l_nodes)[i] = w;
  if (cell_coords != nullptr) {
    const double4* src = reinterpret_cast<const double4*>(cell_coords) + 2 * c;
    double4* dst = reinterpret_cast<double4*>(l_cell_coords) + 2 * i;
    dst[0] = src[0];
    dst[1] = src[1];
  }
}

__global__ void k_gather_coords(int64_t n, const int32_t* __restrict__ l2g, const double* __restrict__ xy, double* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) reinterpret_cast<double2*>(out)[i] = reinterpret_cast<const double2*>(xy)[l2g[i]];
}

__global__ void k_count_quads_local(int64_t n, const uint32_t* __restrict__ cell_nodes, unsigned long long* __restrict__ cnt) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const unsigned q = (i < n && cell_nodes[4 * i + 3] != kNil) ? 1U : 0U;
  const unsigned b = __ballot_sync(0xffffffffU, q);
  if ((threadIdx.x & 31) == 0 && b != 0) atomicAdd(cnt, static_cast<unsigned long long>(__popc(b)));
}

__global__ void k_owned_flags(int64_t n, const int32_t* __restrict__ l2g, const uint8_t* __restrict__ owner, int rank, uint8_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = owner[l2g[i]] == rank ? 1 : 0;
}

struct Scratch {  // frees whatever is still registered when it goes out of scope
  std::vector<void*> ptrs;
  ~Scratch() {
    for (void* p : ptrs) cudaFree(p);
  }
  template <typename T>
  cudaError_t alloc(T** p, size_t n) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), n > 0 ? n : 1);
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
  void release(void* p) {
    for (auto& q : ptrs)
      if (q == p) q = nullptr;
  }
};

// marks -> (g2l, l2g, count): exclusive scan + flagged select of the index sequence
int compact_marks(lfgpu_ctx* ctx, Scratch& sc, int64_t n, const uint8_t* marks, int32_t** g2l, int32_t** l2g, int64_t* count) {
  cudaStream_t st = ctx->stream;
  LFGPU_CUDA_CHECK(ctx, sc.alloc(g2l, sizeof(int32_t) * n));
  int32_t* tmp_l2g = nullptr;
  LFGPU_CUDA_CHECK(ctx, sc.alloc(&tmp_l2g, sizeof(int32_t) * n));
  int64_t* d_num = nullptr;
  LFGPU_CUDA_CHECK(ctx, sc.alloc(&d_num, sizeof(int64_t)));
  size_t tb1 = 0, tb2 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb1, marks, *g2l, n, st);
  cub::CountingInputIterator<int32_t> it(0);
  cub::DeviceSelect::Flagged(nullptr, tb2, it, marks, tmp_l2g, d_num, n, st);
  void* tmp = nullptr;
  LFGPU_CUDA_CHECK(ctx, sc.alloc(&tmp, tb1 > tb2 ? tb1 : tb2));
  LFGPU_CUDA_CHECK(ctx, cub::DeviceScan::ExclusiveSum(tmp, tb1, marks, *g2l, n, st));
  LFGPU_CUDA_CHECK(ctx, cub::DeviceSelect::Flagged(tmp, tb2, it, marks, tmp_l2g, d_num, n, st));
  LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(count, d_num, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
  // keep an exactly sized copy of the list
  LFGPU_CUDA_CHECK(ctx, cudaMalloc(reinterpret_cast<void**>(l2g), sizeof(int32_t) * (*count > 0 ? *count : 1)));
  LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(*l2g, tmp_l2g, sizeof(int32_t) * *count, cudaMemcpyDeviceToDevice, st));
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
  return LFGPU_OK;
}

}  // namespace
}  // namespace lfgpu

using namespace lfgpu;

extern "C" {

int lfgpu_partition_morton(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, int n_parts, uint8_t* d_cell_part) {
  if (ctx == nullptr || mesh == nullptr || d_cell_part == nullptr || n_parts < 1 || n_parts > 255) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int64_t nc = mesh->n_cells;
  Scratch sc;
  unsigned long long* box = nullptr;
  uint64_t *keys = nullptr, *sorted = nullptr;
  LFGPU_CUDA_CHECK(ctx, sc.alloc(&box, 4 * sizeof(unsigned long long)));
  LFGPU_CUDA_CHECK(ctx, sc.alloc(&keys, sizeof(uint64_t) * nc));
  LFGPU_CUDA_CHECK(ctx, sc.alloc(&sorted, sizeof(uint64_t) * nc));
  const unsigned long long init[4] = {~0ULL, ~0ULL, 0ULL, 0ULL};
  LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(box, init, sizeof(init), cudaMemcpyHostToDevice, st));
  k_bbox<<<static_cast<unsigned>(cdiv(mesh->n_nodes, kThreads)), kThreads, 0, st>>>(mesh->n_nodes, mesh->node_coords, box);
  LFGPU_LAUNCH_CHECK(ctx);
  k_morton<<<static_cast<unsigned>(cdiv(nc, kThreads)), kThreads, 0, st>>>(nc, mesh->cell_nodes, mesh->node_coords, box, keys);
  LFGPU_LAUNCH_CHECK(ctx);
  size_t tb = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tb, keys, sorted, nc, 0, 64, st);
  void* tmp = nullptr;
  LFGPU_CUDA_CHECK(ctx, sc.alloc(&tmp, tb));
  LFGPU_CUDA_CHECK(ctx, cub::DeviceRadixSort::SortKeys(tmp, tb, keys, sorted, nc, 0, 64, st));
  const int64_t per = cdiv(nc, n_parts);
  k_assign_parts<<<static_cast<unsigned>(cdiv(nc, kThreads)), kThreads, 0, st>>>(nc, sorted, per, d_cell_part);
  LFGPU_LAUNCH_CHECK(ctx);
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
  return LFGPU_OK;
}

int lfgpu_partition_dof_owner(lfgpu_ctx* ctx, const lfgpu_dofmap* dofmap, const uint8_t* d_cell_part, uint8_t* d_dof_owner) {
  if (ctx == nullptr || dofmap == nullptr || d_cell_part == nullptr || d_dof_owner == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  Scratch sc;
  uint32_t* o32 = nullptr;
  LFGPU_CUDA_CHECK(ctx, sc.alloc(&o32, sizeof(uint32_t) * dofmap->n_dofs));
  LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(o32, 0xFF, sizeof(uint32_t) * dofmap->n_dofs, st));
  const int64_t n = dofmap->n_cells * dofmap->stride;
  k_owner_min<<<static_cast<unsigned>(cdiv(n, kThreads)), kThreads, 0, st>>>(dofmap->n_cells, dofmap->stride, dofmap->cell_dofs, d_cell_part, o32);
  LFGPU_LAUNCH_CHECK(ctx);
  k_narrow_owner<<<static_cast<unsigned>(cdiv(dofmap->n_dofs, kThreads)), kThreads, 0, st>>>(dofmap->n_dofs, o32, d_dof_owner);
  LFGPU_LAUNCH_CHECK(ctx);
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
  return LFGPU_OK;
}

int lfgpu_partition_select_cells(lfgpu_ctx* ctx, const lfgpu_dofmap* dofmap, const uint8_t* d_cell_part, const uint8_t* d_dof_owner, int rank,
                                 int halo, uint8_t* d_cell_sel) {
  if (ctx == nullptr || dofmap == nullptr || d_cell_part == nullptr || d_cell_sel == nullptr || (halo && d_dof_owner == nullptr))
    return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  k_select_cells<<<static_cast<unsigned>(cdiv(dofmap->n_cells, kThreads)), kThreads, 0, ctx->stream>>>(
      dofmap->n_cells, dofmap->stride, dofmap->cell_dofs, d_cell_part, d_dof_owner, rank, halo, d_cell_sel);
  LFGPU_LAUNCH_CHECK(ctx);
  return LFGPU_OK;
}

int lfgpu_submesh_extract(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_dofmap* dofmap, const uint8_t* d_cell_sel, lfgpu_submesh** out) {
  if (ctx == nullptr || mesh == nullptr || dofmap == nullptr || d_cell_sel == nullptr || out == nullptr) return LFGPU_ERR_INVALID;
  *out = nullptr;
  if (dofmap->n_cells != mesh->n_cells) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "dofmap was built for another mesh");
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int64_t nc = mesh->n_cells, nn = mesh->n_nodes, nd = dofmap->n_dofs;
  const int stride = dofmap->stride;
  Scratch sc;
  auto* sub = new lfgpu_submesh;
  sub->ctx = ctx;
  struct Guard {
    lfgpu_submesh* s;
    ~Guard() {
      if (s != nullptr) lfgpu_submesh_destroy(s);
    }
  } guard{sub};
  // cells
  int32_t* g2l_cells = nullptr;
  int rc = compact_marks(ctx, sc, nc, d_cell_sel, &g2l_cells, &sub->l2g_cells, &sub->n_cells);
  if (rc != LFGPU_OK) return rc;
  if (sub->n_cells == 0) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "no cell selected");
  // nodes and dofs the cells refer to
  uint8_t *dof_used = nullptr, *node_used = nullptr;
  LFGPU_CUDA_CHECK(ctx, sc.alloc(&dof_used, nd));
  LFGPU_CUDA_CHECK(ctx, sc.alloc(&node_used, nn));
  LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(dof_used, 0, nd, st));
  LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(node_used, 0, nn, st));
  const unsigned gs = static_cast<unsigned>(cdiv(sub->n_cells, kThreads));
  k_mark_used<<<gs, kThreads, 0, st>>>(sub->n_cells, sub->l2g_cells, stride, dofmap->cell_dofs, mesh->cell_nodes, dof_used, node_used);
  LFGPU_LAUNCH_CHECK(ctx);
  int32_t *g2l_dof = nullptr, *g2l_node = nullptr;
  if ((rc = compact_marks(ctx, sc, nd, dof_used, &g2l_dof, &sub->l2g_dofs, &sub->n_dofs)) != LFGPU_OK) return rc;
  if ((rc = compact_marks(ctx, sc, nn, node_used, &g2l_node, &sub->l2g_nodes, &sub->n_nodes)) != LFGPU_OK) return rc;
  // local mesh and dof table
  auto* m = new lfgpu_mesh;
  sub->mesh = m;
  m->ctx = ctx;
  m->n_nodes = sub->n_nodes;
  m->n_cells = sub->n_cells;
  auto* d = new lfgpu_dofmap;
  sub->dofmap = d;
  d->ctx = ctx;
  d->n_cells = sub->n_cells;
  d->n_dofs = sub->n_dofs;
  d->stride = stride;
  d->max_ldof = dofmap->max_ldof;
  d->n_nodes = sub->n_nodes;
  LFGPU_CUDA_CHECK(ctx, cudaMalloc(&m->node_coords, sizeof(double) * 2 * sub->n_nodes));
  LFGPU_CUDA_CHECK(ctx, cudaMalloc(&m->cell_nodes, sizeof(uint32_t) * 4 * sub->n_cells));
  if (mesh->cell_coords != nullptr) LFGPU_CUDA_CHECK(ctx, cudaMalloc(&m->cell_coords, sizeof(double) * 8 * sub->n_cells));
  LFGPU_CUDA_CHECK(ctx, cudaMalloc(&d->cell_dofs, sizeof(int32_t) * sub->n_cells * stride));
  LFGPU_CUDA_CHECK(ctx, cudaMalloc(&d->n_ldof, sub->n_cells));
  k_local_tables<<<gs, kThreads, 0, st>>>(sub->n_cells, sub->l2g_cells, stride, dofmap->cell_dofs, dofmap->n_ldof, mesh->cell_nodes,
                                           mesh->cell_coords, g2l_dof, g2l_node, d->cell_dofs, d->n_ldof, m->cell_nodes, m->cell_coords);
  LFGPU_LAUNCH_CHECK(ctx);
  k_gather_coords<<<static_cast<unsigned>(cdiv(sub->n_nodes, kThreads)), kThreads, 0, st>>>(sub->n_nodes, sub->l2g_nodes, mesh->node_coords,
                                                                                             m->node_coords);
  LFGPU_LAUNCH_CHECK(ctx);
  unsigned long long* d_cnt = static_cast<unsigned long long*>(ctx->d_scratch);
  LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), st));
  k_count_quads_local<<<gs, kThreads, 0, st>>>(sub->n_cells, m->cell_nodes, d_cnt);
  LFGPU_LAUNCH_CHECK(ctx);
  unsigned long long h_cnt = 0;
  LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(&h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
  m->n_quad = static_cast<int64_t>(h_cnt);
  m->n_tria = m->n_cells - m->n_quad;
  guard.s = nullptr;
  *out = sub;
  return LFGPU_OK;
}

void lfgpu_submesh_destroy(lfgpu_submesh* s) {
  if (s == nullptr) return;
  if (s->ctx != nullptr) {
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
  }
  lfgpu_mesh_destroy(s->mesh);
  lfgpu_dofmap_destroy(s->dofmap);
  cudaFree(s->l2g_cells);
  cudaFree(s->l2g_nodes);
  cudaFree(s->l2g_dofs);
  delete s;
}

lfgpu_mesh* lfgpu_submesh_mesh(lfgpu_submesh* s) { return s ? s->mesh : nullptr; }
lfgpu_dofmap* lfgpu_submesh_dofmap(lfgpu_submesh* s) { return s ? s->dofmap : nullptr; }
int lfgpu_submesh_counts(const lfgpu_submesh* s, int64_t* n_cells, int64_t* n_nodes, int64_t* n_dofs) {
  if (s == nullptr) return LFGPU_ERR_INVALID;
  if (n_cells) *n_cells = s->n_cells;
  if (n_nodes) *n_nodes = s->n_nodes;
  if (n_dofs) *n_dofs = s->n_dofs;
  return LFGPU_OK;
}
const int32_t* lfgpu_submesh_l2g_cells_device(const lfgpu_submesh* s) { return s ? s->l2g_cells : nullptr; }
const int32_t* lfgpu_submesh_l2g_nodes_device(const lfgpu_submesh* s) { return s ? s->l2g_nodes : nullptr; }
const int32_t* lfgpu_submesh_l2g_dofs_device(const lfgpu_submesh* s) { return s ? s->l2g_dofs : nullptr; }

int lfgpu_submesh_owned_dofs(lfgpu_ctx* ctx, const lfgpu_submesh* s, const uint8_t* d_dof_owner, int rank, uint8_t* d_owned) {
  if (ctx == nullptr || s == nullptr || d_dof_owner == nullptr || d_owned == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  k_owned_flags<<<static_cast<unsigned>(cdiv(s->n_dofs, kThreads)), kThreads, 0, ctx->stream>>>(s->n_dofs, s->l2g_dofs, d_dof_owner, rank, d_owned);
  LFGPU_LAUNCH_CHECK(ctx);
  return LFGPU_OK;
}

}  // extern "C"
