// Mesh handling on the device (product code): upload, structured generators with the reference's numbering,
// edge numbering / orientation as lf::mesh::hybrid2d::Mesh assigns them.
//
// Reference behaviour restated here (paths relative to lib/lf/mesh/):
//   hybrid2d/mesh.cc:178-810       edges keyed by (min,max) endpoint; supplied edges keep index and direction; new edges
//                                  are numbered in ascending key order after them and point along the local direction
//                                  of the first (lowest-index) adjacent cell
//   hybrid2d/triangle.cc:70-77     relative orientation of local edge j is positive iff edge.first == cell vertex j
//   utils/tp_triag_mesh_builder.cc:18-178, utils/tp_quad_mesh_builder.cc:19-95   structured builders
// The reference walks a std::map on one core; here the same order falls out of one stable radix sort of
// (key, record) pairs, segment heads and two prefix sums.
#include <cub/cub.cuh>

#include "lfgpu_internal.cuh"

namespace lfgpu {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

// nodes i + j (nx+1) at (x0 + i hx, y0 + j hy)   (tp_triag_mesh_builder.cc:48-61)
__global__ void k_tp_nodes(uint32_t nx, uint32_t ny, double x0, double y0, double hx, double hy, double jitter,
                           uint64_t seed, double* __restrict__ xy) {
  const int64_t n = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(nx + 1) * (ny + 1);
  if (n >= total) return;
  const uint32_t i = static_cast<uint32_t>(n % (nx + 1)), j = static_cast<uint32_t>(n / (nx + 1));
  // mult and add rounded separately, as the host builders do (no FMA contraction) -> bitwise equal coordinates
  double x = __dadd_rn(x0, __dmul_rn(static_cast<double>(i), hx)), y = __dadd_rn(y0, __dmul_rn(static_cast<double>(j), hy));
  if (jitter != 0.0 && i > 0 && i < nx && j > 0 && j < ny) {
    const double u0 = static_cast<double>(splitmix64(seed + 2 * static_cast<uint64_t>(n)) >> 11) * 0x1.0p-53;
    const double u1 = static_cast<double>(splitmix64(seed + 2 * static_cast<uint64_t>(n) + 1) >> 11) * 0x1.0p-53;
    // written as in the spec: x += jitter*h*(2u-1), no FMA contraction so that host and device agree
    x = __dadd_rn(x, __dmul_rn(__dmul_rn(jitter, hx), __dsub_rn(__dmul_rn(2.0, u0), 1.0)));
    y = __dadd_rn(y, __dmul_rn(__dmul_rn(jitter, hy), __dsub_rn(__dmul_rn(2.0, u1), 1.0)));
  }
  xy[2 * n] = x;
  xy[2 * n + 1] = y;
}

// two triangles per square, squares i outer / j inner, "upper" first (tp_triag_mesh_builder.cc:144-176)
__global__ void k_tp_tria_cells(uint32_t nx, uint32_t ny, uint32_t* __restrict__ cell_nodes) {
  const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= static_cast<int64_t>(nx) * ny) return;
  const uint32_t i = static_cast<uint32_t>(s / ny), j = static_cast<uint32_t>(s % ny);
  const uint32_t v00 = i + j * (nx + 1), v10 = v00 + 1, v01 = v00 + (nx + 1), v11 = v01 + 1;
  uint4* out = reinterpret_cast<uint4*>(cell_nodes) + 2 * s;
  out[0] = make_uint4(v00, v11, v01, LFGPU_IDX_NIL);
  out[1] = make_uint4(v00, v10, v11, LFGPU_IDX_NIL);
}

// explicit edge list of the triangle builder: horizontal, vertical, diagonal (tp_triag_mesh_builder.cc:68-136)
__global__ void k_tp_tria_edges(uint32_t nx, uint32_t ny, uint32_t* __restrict__ edge_nodes) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nh = static_cast<int64_t>(nx) * (ny + 1), nv = static_cast<int64_t>(nx + 1) * ny,
                nd = static_cast<int64_t>(nx) * ny;
  if (e >= nh + nv + nd) return;
  uint32_t a, b;
  if (e < nh) {
    const uint32_t i = static_cast<uint32_t>(e / (ny + 1)), j = static_cast<uint32_t>(e % (ny + 1));
    a = i + j * (nx + 1);
    b = a + 1;
  } else if (e < nh + nv) {
    const int64_t r = e - nh;
    const uint32_t i = static_cast<uint32_t>(r / ny), j = static_cast<uint32_t>(r % ny);
    a = i + j * (nx + 1);
    b = a + (nx + 1);
  } else {
    const int64_t r = e - nh - nv;
    const uint32_t i = static_cast<uint32_t>(r / ny), j = static_cast<uint32_t>(r % ny);
    a = i + j * (nx + 1);
    b = a + (nx + 1) + 1;
  }
  edge_nodes[2 * e] = a;
  edge_nodes[2 * e + 1] = b;
}

__global__ void k_tp_quad_cells(uint32_t nx, uint32_t ny, uint32_t* __restrict__ cell_nodes) {
  const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= static_cast<int64_t>(nx) * ny) return;
  const uint32_t i = static_cast<uint32_t>(s / ny), j = static_cast<uint32_t>(s % ny);
  const uint32_t v00 = i + j * (nx + 1), v10 = v00 + 1, v01 = v00 + (nx + 1), v11 = v01 + 1;
  reinterpret_cast<uint4*>(cell_nodes)[s] = make_uint4(v00, v10, v11, v01);  // tp_quad_mesh_builder.cc:70-72
}

// hybrid mesh (spec in DESIGN.md): square (i,j), i outer / j inner, is one quad if (i+j) even, else two triangles.
// cells before square s = i n + j:  s + (number of odd squares among the first s)
__device__ __forceinline__ int64_t hybrid_cells_before(uint32_t n, uint32_t i, uint32_t j) {
  // odd squares (i'+j' odd) in full columns i' < i: per column: n/2 if n even; if n odd: (n-1)/2 for even i', (n+1)/2 for odd i'
  int64_t odd;
  if (n % 2 == 0) {
    odd = static_cast<int64_t>(i) * (n / 2);
  } else {
    const int64_t even_cols = (i + 1) / 2, odd_cols = i / 2;
    odd = even_cols * ((n - 1) / 2) + odd_cols * ((n + 1) / 2);
  }
  // partial column i, rows j' < j: j' with (i + j') odd
  odd += (i % 2 == 0) ? (j / 2) : ((j + 1) / 2);
  return static_cast<int64_t>(i) * n + j + odd;
}
__global__ void k_hybrid_cells(uint32_t n, uint32_t* __restrict__ cell_nodes) {
  const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= static_cast<int64_t>(n) * n) return;
  const uint32_t i = static_cast<uint32_t>(s / n), j = static_cast<uint32_t>(s % n);
  const uint32_t v00 = i + j * (n + 1), v10 = v00 + 1, v01 = v00 + (n + 1), v11 = v01 + 1;
  const int64_t c = hybrid_cells_before(n, i, j);
  uint4* out = reinterpret_cast<uint4*>(cell_nodes) + c;
  if ((i + j) % 2 == 0) {
    out[0] = make_uint4(v00, v10, v11, v01);
  } else {
    out[0] = make_uint4(v00, v11, v01, LFGPU_IDX_NIL);
    out[1] = make_uint4(v00, v10, v11, LFGPU_IDX_NIL);
  }
}

__global__ void k_count_quads(int64_t n_cells, const uint32_t* __restrict__ cell_nodes, unsigned long long* count) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  unsigned q = (c < n_cells && cell_nodes[4 * c + 3] != LFGPU_IDX_NIL) ? 1U : 0U;
  q = __reduce_add_sync(0xffffffffU, q);
  if ((threadIdx.x & 31) == 0 && q) atomicAdd(count, static_cast<unsigned long long>(q));
}

// validity checks of the upload: node indices in range; geometry non-degenerate (tria_o1.cc:10-48, quad_o1.cc:14-59)
__global__ void k_validate_cells(int64_t n_cells, int64_t n_nodes, const uint32_t* __restrict__ cell_nodes,
                                 const double* __restrict__ node_coords, const double* __restrict__ cell_coords,
                                 int* __restrict__ flags) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[c];
  const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
  const int nv = (v.w == LFGPU_IDX_NIL) ? 3 : 4;
  double x[4], y[4];
  for (int k = 0; k < nv; ++k) {
    if (vv[k] >= n_nodes) {
      flags[0] = 1;
      return;
    }
    if (cell_coords != nullptr) {
      x[k] = cell_coords[8 * c + 2 * k];
      y[k] = cell_coords[8 * c + 2 * k + 1];
    } else {
      x[k] = node_coords[2 * static_cast<int64_t>(vv[k])];
      y[k] = node_coords[2 * static_cast<int64_t>(vv[k]) + 1];
    }
  }
  const double tol = 1.0e-8;
  double circum = 0.0, emin = 1e300;
  for (int k = 0; k < nv; ++k) {
    const int k1 = (k + 1) % nv;
    const double l2 = (x[k1] - x[k]) * (x[k1] - x[k]) + (y[k1] - y[k]) * (y[k1] - y[k]);
    circum += l2;
    emin = fmin(emin, l2);
  }
  double area = fabs((x[1] - x[0]) * (y[2] - y[0]) - (y[1] - y[0]) * (x[2] - x[0]));
  if (nv == 4) area += fabs((x[3] - x[0]) * (y[2] - y[0]) - (y[3] - y[0]) * (x[2] - x[0]));
  if (!(emin > tol * circum) || !(area > tol * circum)) flags[1] = 1;
}

// ---- topology -------------------------------------------------------------------------------------------------------
// record r < n_explicit: supplied edge r; record n_explicit + 4 c + j: local edge j of cell c (slot 3 of a triangle unused)
__global__ void k_edge_records(int64_t n_explicit, const uint32_t* __restrict__ explicit_nodes, int64_t n_cells,
                               const uint32_t* __restrict__ cell_nodes, uint64_t* __restrict__ keys,
                               uint32_t* __restrict__ recs) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = n_explicit + 4 * n_cells;
  if (r >= total) return;
  uint32_t a, b;
  if (r < n_explicit) {
    a = explicit_nodes[2 * r];
    b = explicit_nodes[2 * r + 1];
  } else {
    const int64_t q = r - n_explicit;
    const int64_t c = q >> 2;
    const int j = static_cast<int>(q & 3);
    const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[c];
    const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
    const int nv = (v.w == LFGPU_IDX_NIL) ? 3 : 4;
    if (j >= nv) {
      keys[r] = ~0ULL;  // sorts last, ignored
      recs[r] = static_cast<uint32_t>(r);
      return;
    }
    a = vv[j];
    b = vv[(j + 1) % nv];
  }
  const uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
  keys[r] = (static_cast<uint64_t>(lo) << 32) | hi;
  recs[r] = static_cast<uint32_t>(r);
}

// head flags of the sorted records; new (not supplied) edges get 1 in `is_new`
__global__ void k_edge_heads(int64_t total, int64_t n_explicit, const uint64_t* __restrict__ keys,
                             const uint32_t* __restrict__ recs, uint32_t* __restrict__ head, uint32_t* __restrict__ is_new) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= total) return;
  const uint64_t k = keys[r];
  const bool valid = (k != ~0ULL);
  const bool h = valid && (r == 0 || keys[r - 1] != k);
  head[r] = h ? 1U : 0U;
  is_new[r] = (h && recs[r] >= n_explicit) ? 1U : 0U;
}

// one thread per sorted record: the segment head decides the edge's index and direction
__global__ void k_edge_assign(int64_t total, int64_t n_explicit, const uint64_t* __restrict__ keys,
                              const uint32_t* __restrict__ recs, const uint32_t* __restrict__ head,
                              const uint32_t* __restrict__ head_scan /*inclusive*/, const uint32_t* __restrict__ new_scan /*exclusive*/,
                              const int64_t* __restrict__ head_pos /*position of the head of segment s*/,
                              const uint32_t* __restrict__ explicit_nodes, const uint32_t* __restrict__ cell_nodes,
                              const uint8_t* __restrict__ cell_geo, uint32_t* __restrict__ edge_nodes, uint32_t* __restrict__ cell_edges,
                              int8_t* __restrict__ cell_edge_ori, int* __restrict__ flags) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= total) return;
  const uint64_t k = keys[r];
  if (k == ~0ULL) return;
  const uint32_t seg = head_scan[r] - 1;
  const int64_t hp = head_pos[seg];
  const uint32_t hrec = recs[hp];
  // index and direction from the head record
  uint32_t eidx, first, second;
  if (hrec < n_explicit) {
    eidx = hrec;
    first = explicit_nodes[2 * static_cast<int64_t>(hrec)];
    second = explicit_nodes[2 * static_cast<int64_t>(hrec) + 1];
  } else {
    eidx = static_cast<uint32_t>(n_explicit) + new_scan[hp];
    const int64_t q = static_cast<int64_t>(hrec) - n_explicit;
    const int64_t c = q >> 2;
    const int j = static_cast<int>(q & 3);
    const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[c];
    const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
    const int nv = (v.w == LFGPU_IDX_NIL) ? 3 : 4;
    first = vv[j];
    second = vv[(j + 1) % nv];
    // mesh.cc:413-428: an edge created by a cell WITHOUT geometry takes its geometry from the first later cell that
    // has one; if that cell runs along the edge the other way round the edge is reversed
    if (cell_geo != nullptr && cell_geo[c] == 0) {
      for (int64_t s = hp + 1; s < total && keys[s] == k; ++s) {
        const int64_t q2 = static_cast<int64_t>(recs[s]) - n_explicit;
        const int64_t c2 = q2 >> 2;
        if (cell_geo[c2] != 0) {
          if (cell_nodes[4 * c2 + (q2 & 3)] != first) {
            const uint32_t t = first;
            first = second;
            second = t;
          }
          break;
        }
      }
    }
  }
  const uint32_t rec = recs[r];
  if (head[r]) {
    edge_nodes[2 * static_cast<int64_t>(eidx)] = first;
    edge_nodes[2 * static_cast<int64_t>(eidx) + 1] = second;
  }
  if (rec < n_explicit) {
    if (!head[r]) flags[2] = 1;  // duplicate supplied edge (mesh.cc:268-270)
    return;
  }
  const int64_t q = static_cast<int64_t>(rec) - n_explicit;
  const int64_t c = q >> 2;
  const int j = static_cast<int>(q & 3);
  const uint32_t vj = cell_nodes[4 * c + j];
  cell_edges[4 * c + j] = eidx;
  cell_edge_ori[4 * c + j] = (first == vj) ? 1 : -1;  // triangle.cc:70-77
}

__global__ void k_head_positions(int64_t total, const uint32_t* __restrict__ head, const uint32_t* __restrict__ head_scan,
                                 int64_t* __restrict__ head_pos) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= total) return;
  if (head[r]) head_pos[head_scan[r] - 1] = r;
}

__global__ void k_fill_u32(int64_t n, uint32_t* p, uint32_t v) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_fill_i8(int64_t n, int8_t* p, int8_t v) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int finish_mesh(lfgpu_ctx* ctx, lfgpu_mesh* m) {
  // count quads, validate
  unsigned long long* d_cnt = static_cast<unsigned long long*>(ctx->d_scratch);
  int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 64);
  LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_scratch, 0, 256, ctx->stream));
  const unsigned grid = static_cast<unsigned>(cdiv(m->n_cells, kThreads));
  k_count_quads<<<grid, kThreads, 0, ctx->stream>>>(m->n_cells, m->cell_nodes, d_cnt);
  LFGPU_LAUNCH_CHECK(ctx);
  k_validate_cells<<<grid, kThreads, 0, ctx->stream>>>(m->n_cells, m->n_nodes, m->cell_nodes, m->node_coords, m->cell_coords, d_flags);
  LFGPU_LAUNCH_CHECK(ctx);
  unsigned long long h_cnt = 0;
  int h_flags[4] = {0, 0, 0, 0};
  LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(&h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, ctx->stream));
  LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, ctx->stream));
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  m->n_quad = static_cast<int64_t>(h_cnt);
  m->n_tria = m->n_cells - m->n_quad;
  if (h_flags[0]) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "cell references a node index >= n_nodes");
  if (h_flags[1]) LFGPU_FAIL(ctx, LFGPU_ERR_DEGENERATE, "degenerate cell geometry (collapsed edge or zero area)");
  return LFGPU_OK;
}

int alloc_mesh(lfgpu_ctx* ctx, int64_t n_nodes, int64_t n_cells, bool with_cell_coords, lfgpu_mesh** out) {
  if (n_nodes <= 0 || n_cells <= 0) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "empty mesh");
  if (n_cells >= (1LL << 28) || n_nodes >= (1LL << 31)) LFGPU_FAIL(ctx, LFGPU_ERR_OVERFLOW, "mesh too large for 28-bit cell / 31-bit node indices");
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  auto* m = new lfgpu_mesh;
  m->ctx = ctx;
  m->n_nodes = n_nodes;
  m->n_cells = n_cells;
  cudaError_t e = cudaMalloc(&m->node_coords, sizeof(double) * 2 * n_nodes);
  if (e == cudaSuccess) e = cudaMalloc(&m->cell_nodes, sizeof(uint32_t) * 4 * n_cells);
  if (e == cudaSuccess && with_cell_coords) e = cudaMalloc(&m->cell_coords, sizeof(double) * 8 * n_cells);
  if (e != cudaSuccess) {
    lfgpu_mesh_destroy(m);
    LFGPU_FAIL(ctx, LFGPU_ERR_CUDA, std::string("cudaMalloc mesh: ") + cudaGetErrorString(e));
  }
  *out = m;
  return LFGPU_OK;
}

}  // namespace

// New node positions are checked like those of lfgpu_mesh_upload (the reference asserts on a degenerate cell, tria_o1.cc:10-48;
// the kernels' 1 / det has no slow path), asynchronously on the context stream: the flag is read by the next lfgpu_ctx_synchronize.
int queue_geometry_check(lfgpu_ctx* ctx, const lfgpu_mesh* mesh) {
  int* d_flag = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 1024);
  if (!ctx->geom_check_pending) LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_flag, 0, 16, ctx->stream));
  k_validate_cells<<<static_cast<unsigned>(cdiv(mesh->n_cells, kThreads)), kThreads, 0, ctx->stream>>>(mesh->n_cells, mesh->n_nodes, mesh->cell_nodes,
                                                                                                    mesh->node_coords, mesh->cell_coords, d_flag);
  LFGPU_LAUNCH_CHECK(ctx);
  ctx->geom_check_pending = true;
  return LFGPU_OK;
}

}  // namespace lfgpu

using namespace lfgpu;

extern "C" {

void lfgpu_mesh_destroy(lfgpu_mesh* m) {
  if (m == nullptr) return;
  if (m->ctx) {
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
  }
  cudaFree(m->node_coords);
  cudaFree(m->cell_nodes);
  cudaFree(m->cell_coords);
  cudaFree(m->edge_nodes);
  cudaFree(m->cell_edges);
  cudaFree(m->cell_edge_ori);
  delete m;
}

int lfgpu_mesh_upload(lfgpu_ctx* ctx, int64_t n_nodes, const double* node_coords, int64_t n_cells,
                      const uint32_t* cell_nodes, const double* cell_coords, lfgpu_mesh** out) {
  if (ctx == nullptr || out == nullptr || node_coords == nullptr || cell_nodes == nullptr) return LFGPU_ERR_INVALID;
  *out = nullptr;
  lfgpu_mesh* m = nullptr;
  int rc = alloc_mesh(ctx, n_nodes, n_cells, cell_coords != nullptr, &m);
  if (rc != LFGPU_OK) return rc;
  cudaError_t e = cudaMemcpyAsync(m->node_coords, node_coords, sizeof(double) * 2 * n_nodes, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->cell_nodes, cell_nodes, sizeof(uint32_t) * 4 * n_cells, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && cell_coords) e = cudaMemcpyAsync(m->cell_coords, cell_coords, sizeof(double) * 8 * n_cells, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    lfgpu_mesh_destroy(m);
    LFGPU_FAIL(ctx, LFGPU_ERR_CUDA, std::string("mesh upload: ") + cudaGetErrorString(e));
  }
  rc = finish_mesh(ctx, m);
  if (rc != LFGPU_OK) {
    lfgpu_mesh_destroy(m);
    return rc;
  }
  *out = m;
  return LFGPU_OK;
}

int lfgpu_mesh_update_node_coords(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const double* node_coords) {
  if (ctx == nullptr || mesh == nullptr || node_coords == nullptr) return LFGPU_ERR_INVALID;
  if (mesh->cell_coords != nullptr) LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "mesh carries explicit cell corner coordinates");
  LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(mesh->node_coords, node_coords, sizeof(double) * 2 * mesh->n_nodes, cudaMemcpyHostToDevice, ctx->stream));
  mesh->coords_version++;
  return queue_geometry_check(ctx, mesh);
}

static int tp_common(lfgpu_ctx* ctx, uint32_t nx, uint32_t ny, double x0, double y0, double x1, double y1, int64_t n_cells,
                     double jitter, uint64_t seed, lfgpu_mesh** out) {
  if (ctx == nullptr || out == nullptr) return LFGPU_ERR_INVALID;
  *out = nullptr;
  if (nx == 0 || ny == 0 || !(x1 - x0 > 0.0) || !(y1 - y0 > 0.0)) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "empty tensor-product mesh");
  const int64_t n_nodes = static_cast<int64_t>(nx + 1) * (ny + 1);
  lfgpu_mesh* m = nullptr;
  const int rc = alloc_mesh(ctx, n_nodes, n_cells, false, &m);
  if (rc != LFGPU_OK) return rc;
  const double hx = (x1 - x0) / nx, hy = (y1 - y0) / ny;
  k_tp_nodes<<<static_cast<unsigned>(cdiv(n_nodes, kThreads)), kThreads, 0, ctx->stream>>>(nx, ny, x0, y0, hx, hy, jitter, seed, m->node_coords);
  ctx->launches++;
  *out = m;
  return LFGPU_OK;
}

int lfgpu_mesh_tp_tria(lfgpu_ctx* ctx, uint32_t nx, uint32_t ny, double x0, double y0, double x1, double y1, lfgpu_mesh** out) {
  const int64_t squares = static_cast<int64_t>(nx) * ny;
  int rc = tp_common(ctx, nx, ny, x0, y0, x1, y1, 2 * squares, 0.0, 0, out);
  if (rc != LFGPU_OK) return rc;
  lfgpu_mesh* m = *out;
  k_tp_tria_cells<<<static_cast<unsigned>(cdiv(squares, kThreads)), kThreads, 0, ctx->stream>>>(nx, ny, m->cell_nodes);
  ctx->launches++;
  rc = finish_mesh(ctx, m);
  if (rc != LFGPU_OK) {
    lfgpu_mesh_destroy(m);
    *out = nullptr;
    return rc;
  }
  // the builder supplies all edges explicitly; remember the recipe, build lazily (only P2/P3 dof maps need edges)
  m->tp_nx = nx;
  m->tp_ny = ny;
  m->n_edges = static_cast<int64_t>(nx) * (ny + 1) + static_cast<int64_t>(nx + 1) * ny + squares;
  return LFGPU_OK;
}

int lfgpu_mesh_tp_quad(lfgpu_ctx* ctx, uint32_t nx, uint32_t ny, double x0, double y0, double x1, double y1, lfgpu_mesh** out) {
  const int64_t squares = static_cast<int64_t>(nx) * ny;
  int rc = tp_common(ctx, nx, ny, x0, y0, x1, y1, squares, 0.0, 0, out);
  if (rc != LFGPU_OK) return rc;
  lfgpu_mesh* m = *out;
  k_tp_quad_cells<<<static_cast<unsigned>(cdiv(squares, kThreads)), kThreads, 0, ctx->stream>>>(nx, ny, m->cell_nodes);
  ctx->launches++;
  rc = finish_mesh(ctx, m);
  if (rc != LFGPU_OK) {
    lfgpu_mesh_destroy(m);
    *out = nullptr;
  }
  return rc;
}

int lfgpu_mesh_hybrid(lfgpu_ctx* ctx, uint32_t n, double jitter, uint64_t seed, lfgpu_mesh** out) {
  if (n == 0) return LFGPU_ERR_INVALID;
  const int64_t squares = static_cast<int64_t>(n) * n;
  const int64_t odd = squares / 2;  // squares with (i+j) odd
  int rc = tp_common(ctx, n, n, 0.0, 0.0, 1.0, 1.0, squares + odd, jitter, seed, out);
  if (rc != LFGPU_OK) return rc;
  lfgpu_mesh* m = *out;
  k_hybrid_cells<<<static_cast<unsigned>(cdiv(squares, kThreads)), kThreads, 0, ctx->stream>>>(n, m->cell_nodes);
  ctx->launches++;
  rc = finish_mesh(ctx, m);
  if (rc != LFGPU_OK) {
    lfgpu_mesh_destroy(m);
    *out = nullptr;
  }
  return rc;
}

}  // extern "C"

namespace lfgpu {
// d_explicit_in: device array [n_explicit][2] or nullptr; if tp_recipe the explicit list of the triangle builder is generated
static int build_topology_impl(lfgpu_ctx* ctx, lfgpu_mesh* m, int64_t n_explicit, const uint32_t* edge_nodes_host, bool tp_recipe,
                               const uint8_t* cell_geo_host, const uint32_t* d_explicit_given = nullptr) {
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int64_t total = n_explicit + 4 * m->n_cells;
  if (total >= (1LL << 32)) LFGPU_FAIL(ctx, LFGPU_ERR_OVERFLOW, "too many edge records");
  uint32_t* d_explicit = nullptr;
  uint8_t* d_geo = nullptr;
  uint64_t *keys_in = nullptr, *keys_out = nullptr;
  uint32_t *recs_in = nullptr, *recs_out = nullptr, *head = nullptr, *is_new = nullptr, *head_scan = nullptr, *new_scan = nullptr;
  int64_t* head_pos = nullptr;
  void* tmp = nullptr;
  int rc = LFGPU_OK;
  auto cleanup = [&]() {
    cudaFree(d_geo);
    cudaFree(d_explicit); cudaFree(keys_in); cudaFree(keys_out); cudaFree(recs_in); cudaFree(recs_out);
    cudaFree(head); cudaFree(is_new); cudaFree(head_scan); cudaFree(new_scan); cudaFree(head_pos); cudaFree(tmp);
  };
#define TOPO_CHECK(expr)                                                                    \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      set_last_error(ctx, std::string(#expr) + ": " + cudaGetErrorString(_e));              \
      cleanup();                                                                            \
      return LFGPU_ERR_CUDA;                                                                \
    }                                                                                       \
  } while (0)
  if (n_explicit > 0) {
    TOPO_CHECK(cudaMalloc(&d_explicit, sizeof(uint32_t) * 2 * n_explicit));
    if (tp_recipe) {
      k_tp_tria_edges<<<static_cast<unsigned>(cdiv(n_explicit, kThreads)), kThreads, 0, st>>>(m->tp_nx, m->tp_ny, d_explicit);
      ctx->launches++;
    } else if (d_explicit_given != nullptr) {
      TOPO_CHECK(cudaMemcpyAsync(d_explicit, d_explicit_given, sizeof(uint32_t) * 2 * n_explicit, cudaMemcpyDeviceToDevice, st));
    } else {
      TOPO_CHECK(cudaMemcpyAsync(d_explicit, edge_nodes_host, sizeof(uint32_t) * 2 * n_explicit, cudaMemcpyHostToDevice, st));
    }
  }
  if (cell_geo_host != nullptr) {
    TOPO_CHECK(cudaMalloc(&d_geo, m->n_cells));
    TOPO_CHECK(cudaMemcpyAsync(d_geo, cell_geo_host, m->n_cells, cudaMemcpyHostToDevice, st));
  }
  TOPO_CHECK(cudaMalloc(&keys_in, sizeof(uint64_t) * total));
  TOPO_CHECK(cudaMalloc(&keys_out, sizeof(uint64_t) * total));
  TOPO_CHECK(cudaMalloc(&recs_in, sizeof(uint32_t) * total));
  TOPO_CHECK(cudaMalloc(&recs_out, sizeof(uint32_t) * total));
  TOPO_CHECK(cudaMalloc(&head, sizeof(uint32_t) * total));
  TOPO_CHECK(cudaMalloc(&is_new, sizeof(uint32_t) * total));
  TOPO_CHECK(cudaMalloc(&head_scan, sizeof(uint32_t) * total));
  TOPO_CHECK(cudaMalloc(&new_scan, sizeof(uint32_t) * total));
  const unsigned grid = static_cast<unsigned>(cdiv(total, kThreads));
  k_edge_records<<<grid, kThreads, 0, st>>>(n_explicit, d_explicit, m->n_cells, m->cell_nodes, keys_in, recs_in);
  ctx->launches++;
  size_t tmp_bytes = 0, tb2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, recs_in, recs_out, total, 0, 64, st);
  cub::DeviceScan::InclusiveSum(nullptr, tb2, head, head_scan, total, st);
  tmp_bytes = tmp_bytes > tb2 ? tmp_bytes : tb2;
  TOPO_CHECK(cudaMalloc(&tmp, tmp_bytes));
  TOPO_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, recs_in, recs_out, total, 0, 64, st));
  k_edge_heads<<<grid, kThreads, 0, st>>>(total, n_explicit, keys_out, recs_out, head, is_new);
  ctx->launches++;
  TOPO_CHECK(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, head, head_scan, total, st));
  TOPO_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, is_new, new_scan, total, st));
  uint32_t n_edges_u = 0;
  TOPO_CHECK(cudaMemcpyAsync(&n_edges_u, head_scan + (total - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  TOPO_CHECK(cudaStreamSynchronize(st));
  const int64_t n_edges = n_edges_u;
  TOPO_CHECK(cudaMalloc(&head_pos, sizeof(int64_t) * (n_edges > 0 ? n_edges : 1)));
  k_head_positions<<<grid, kThreads, 0, st>>>(total, head, head_scan, head_pos);
  ctx->launches++;
  cudaFree(m->edge_nodes); cudaFree(m->cell_edges); cudaFree(m->cell_edge_ori);
  m->edge_nodes = nullptr; m->cell_edges = nullptr; m->cell_edge_ori = nullptr;
  TOPO_CHECK(cudaMalloc(&m->edge_nodes, sizeof(uint32_t) * 2 * (n_edges > 0 ? n_edges : 1)));
  TOPO_CHECK(cudaMalloc(&m->cell_edges, sizeof(uint32_t) * 4 * m->n_cells));
  TOPO_CHECK(cudaMalloc(&m->cell_edge_ori, sizeof(int8_t) * 4 * m->n_cells));
  k_fill_u32<<<static_cast<unsigned>(cdiv(4 * m->n_cells, kThreads)), kThreads, 0, st>>>(4 * m->n_cells, m->cell_edges, LFGPU_IDX_NIL);
  k_fill_i8<<<static_cast<unsigned>(cdiv(4 * m->n_cells, kThreads)), kThreads, 0, st>>>(4 * m->n_cells, m->cell_edge_ori, 0);
  ctx->launches += 2;
  int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 64);
  TOPO_CHECK(cudaMemsetAsync(d_flags, 0, 64, st));
  k_edge_assign<<<grid, kThreads, 0, st>>>(total, n_explicit, keys_out, recs_out, head, head_scan, new_scan, head_pos, d_explicit,
                                           m->cell_nodes, d_geo, m->edge_nodes, m->cell_edges, m->cell_edge_ori, d_flags);
  ctx->launches++;
  int h_flags[4] = {0, 0, 0, 0};
  TOPO_CHECK(cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
  TOPO_CHECK(cudaStreamSynchronize(st));
  cleanup();
#undef TOPO_CHECK
  if (h_flags[2]) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "duplicate edge in the supplied edge list");
  m->n_edges = n_edges;
  m->has_topology = true;
  return rc;
}
}  // namespace lfgpu

extern "C" {

int lfgpu_mesh_build_topology(lfgpu_ctx* ctx, lfgpu_mesh* m, int64_t n_explicit, const uint32_t* edge_nodes_host,
                              const uint8_t* cell_has_geometry) {
  if (ctx == nullptr || m == nullptr || n_explicit < 0 || (n_explicit > 0 && edge_nodes_host == nullptr)) return LFGPU_ERR_INVALID;
  return build_topology_impl(ctx, m, n_explicit, edge_nodes_host, false, cell_has_geometry);
}

int lfgpu_mesh_counts(const lfgpu_mesh* m, int64_t* n_nodes, int64_t* n_edges, int64_t* n_cells, int64_t* n_tria, int64_t* n_quad) {
  if (m == nullptr) return LFGPU_ERR_INVALID;
  if (n_nodes) *n_nodes = m->n_nodes;
  if (n_edges) *n_edges = m->n_edges;
  if (n_cells) *n_cells = m->n_cells;
  if (n_tria) *n_tria = m->n_tria;
  if (n_quad) *n_quad = m->n_quad;
  return LFGPU_OK;
}

}  // extern "C"

namespace lfgpu {
// make sure edges are numbered (lazy: P1 never needs them)
int ensure_topology(lfgpu_ctx* ctx, lfgpu_mesh* m) {
  if (m->has_topology) return LFGPU_OK;
  if (m->tp_nx > 0) return build_topology_impl(ctx, m, m->n_edges, nullptr, true, nullptr);
  return build_topology_impl(ctx, m, 0, nullptr, false, nullptr);
}

__global__ void k_gather_cell_coords(int64_t n_cells, const uint32_t* __restrict__ cell_nodes, const double* __restrict__ node_coords,
                                     const double* __restrict__ cell_coords, double* __restrict__ out, uint8_t* __restrict__ cell_type) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[c];
  const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
  const int nv = (v.w == LFGPU_IDX_NIL) ? 3 : 4;
  if (cell_type) cell_type[c] = static_cast<uint8_t>(nv);
  if (out == nullptr) return;
  for (int k = 0; k < 4; ++k) {
    double x = 0.0, y = 0.0;
    if (k < nv) {
      if (cell_coords) {
        x = cell_coords[8 * c + 2 * k];
        y = cell_coords[8 * c + 2 * k + 1];
      } else {
        x = node_coords[2 * static_cast<int64_t>(vv[k])];
        y = node_coords[2 * static_cast<int64_t>(vv[k]) + 1];
      }
    }
    out[8 * c + 2 * k] = x;
    out[8 * c + 2 * k + 1] = y;
  }
}
}  // namespace lfgpu

extern "C" int lfgpu_mesh_download(lfgpu_ctx* ctx, const lfgpu_mesh* m, uint8_t* cell_type, uint32_t* cell_nodes, double* cell_coords,
                                   uint32_t* cell_edges, int8_t* cell_edge_ori, uint32_t* edge_nodes, double* node_coords) {
  if (ctx == nullptr || m == nullptr) return LFGPU_ERR_INVALID;
  cudaStream_t st = ctx->stream;
  if ((cell_edges || cell_edge_ori || edge_nodes)) {
    const int rc = ensure_topology(ctx, const_cast<lfgpu_mesh*>(m));
    if (rc != LFGPU_OK) return rc;
  }
  if (cell_nodes) LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(cell_nodes, m->cell_nodes, sizeof(uint32_t) * 4 * m->n_cells, cudaMemcpyDeviceToHost, st));
  if (node_coords) LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(node_coords, m->node_coords, sizeof(double) * 2 * m->n_nodes, cudaMemcpyDeviceToHost, st));
  if (cell_edges) LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(cell_edges, m->cell_edges, sizeof(uint32_t) * 4 * m->n_cells, cudaMemcpyDeviceToHost, st));
  if (cell_edge_ori) LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(cell_edge_ori, m->cell_edge_ori, sizeof(int8_t) * 4 * m->n_cells, cudaMemcpyDeviceToHost, st));
  if (edge_nodes) LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(edge_nodes, m->edge_nodes, sizeof(uint32_t) * 2 * m->n_edges, cudaMemcpyDeviceToHost, st));
  if (cell_coords || cell_type) {
    double* d_cc = nullptr;
    uint8_t* d_ct = nullptr;
    if (cell_coords) LFGPU_CUDA_CHECK(ctx, cudaMalloc(&d_cc, sizeof(double) * 8 * m->n_cells));
    if (cell_type) LFGPU_CUDA_CHECK(ctx, cudaMalloc(&d_ct, m->n_cells));
    k_gather_cell_coords<<<static_cast<unsigned>(cdiv(m->n_cells, 256)), 256, 0, st>>>(m->n_cells, m->cell_nodes, m->node_coords, m->cell_coords, d_cc, d_ct);
    ctx->launches++;
    if (cell_coords) cudaMemcpyAsync(cell_coords, d_cc, sizeof(double) * 8 * m->n_cells, cudaMemcpyDeviceToHost, st);
    if (cell_type) cudaMemcpyAsync(cell_type, d_ct, m->n_cells, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    cudaFree(d_cc);
    cudaFree(d_ct);
  }
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
  return LFGPU_OK;
}

// ---- regular refinement with the reference's numbering -----------------------------------------------------------------
// MeshHierarchy::RefineRegular (refinement/mesh_hierarchy.cc:72-114) -> PerformRefinement (:368-1262) for the all-regular
// case, as a device-side mesh generator (the input of BASELINE config 4 at scale).  Numbering = order of the reference's
// MeshFactory calls: nodes = copies, then edge midpoints in edge order, then quad centres in cell order; ALL edges
// explicit: (p0, mid), (mid, p1) per parent edge, then per parent cell its interior edges; four children per cell.
// Child corners are the parent map at lattice points / 6 (hybrid2d_refinement_pattern.cc, tria_o1.cc:99-151): for
// straight-sided parents whose corners are node positions these are bitwise the new node positions, so the refined mesh
// needs no separate cell corner array.  Products and sums are rounded separately like the host code (no FMA).
namespace lfgpu {
namespace {
__global__ void k_ref_is_quad(int64_t n_cells, const uint32_t* __restrict__ cell_nodes, int32_t* __restrict__ is_quad) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c < n_cells) is_quad[c] = cell_nodes[4 * c + 3] != LFGPU_IDX_NIL ? 1 : 0;
}
__global__ void k_ref_nodes_edges(int64_t nn, int64_t ne, const double* __restrict__ xy, const uint32_t* __restrict__ edge_nodes,
                                  double* __restrict__ xy_f, uint32_t* __restrict__ edges_f) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t < nn) {
    xy_f[2 * t] = xy[2 * t];
    xy_f[2 * t + 1] = xy[2 * t + 1];
  } else if (t < nn + ne) {
    const int64_t e = t - nn;
    const uint32_t p0 = edge_nodes[2 * e], p1 = edge_nodes[2 * e + 1];
    const double h = 1.0 / 6.0, s = h * 3.0;  // lattice constant 6, midpoint = lattice point 3
    // SegmentO1::Global (segment_o1.cc:9-11): col(1) * t + col(0) * (1 - t)
    xy_f[2 * t] = __dadd_rn(__dmul_rn(xy[2 * p1], s), __dmul_rn(xy[2 * p0], 1.0 - s));
    xy_f[2 * t + 1] = __dadd_rn(__dmul_rn(xy[2 * p1 + 1], s), __dmul_rn(xy[2 * p0 + 1], 1.0 - s));
    const uint32_t mid = static_cast<uint32_t>(t);
    edges_f[4 * e] = p0;
    edges_f[4 * e + 1] = mid;
    edges_f[4 * e + 2] = mid;
    edges_f[4 * e + 3] = p1;
  }
}
__global__ void k_ref_cells(int64_t nc, int64_t nn, int64_t ne, const uint32_t* __restrict__ cell_nodes, const uint32_t* __restrict__ cell_edges,
                            const int32_t* __restrict__ quads_before, const double* __restrict__ xy, double* __restrict__ xy_f,
                            uint32_t* __restrict__ edges_f, uint32_t* __restrict__ cells_f) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const uint32_t v[4] = {cell_nodes[4 * c], cell_nodes[4 * c + 1], cell_nodes[4 * c + 2], cell_nodes[4 * c + 3]};
  const bool quad = v[3] != LFGPU_IDX_NIL;
  const int64_t qb = quads_before[c], tb = c - qb;
  uint32_t m[4];
  for (int j = 0; j < (quad ? 4 : 3); ++j) m[j] = static_cast<uint32_t>(nn + cell_edges[4 * c + j]);
  uint32_t* ed = edges_f + 2 * (2 * ne + 3 * tb + 4 * qb);  // interior edges of this cell
  uint32_t* ch = cells_f + 16 * c;
  if (!quad) {
    // interior edges (m0,m2), (m0,m1), (m2,m1); children (v0,m0,m2), (v1,m0,m1), (v2,m2,m1), (m0,m1,m2)
    ed[0] = m[0]; ed[1] = m[2]; ed[2] = m[0]; ed[3] = m[1]; ed[4] = m[2]; ed[5] = m[1];
    const uint32_t t4[4][4] = {{v[0], m[0], m[2], LFGPU_IDX_NIL}, {v[1], m[0], m[1], LFGPU_IDX_NIL}, {v[2], m[2], m[1], LFGPU_IDX_NIL},
                               {m[0], m[1], m[2], LFGPU_IDX_NIL}};
    for (int k = 0; k < 4; ++k)
      for (int l = 0; l < 4; ++l) ch[4 * k + l] = t4[k][l];
  } else {
    const uint32_t ctr = static_cast<uint32_t>(nn + ne + qb);
    // QuadO1::Global at (1/2, 1/2) (quad_o1.cc:68-83): c0 (1-x0)(1-x1) + c1 x0 (1-x1) + c2 x0 x1 + c3 (1-x0) x1
    const double h = 1.0 / 6.0, x0 = h * 3.0, x1 = h * 3.0;
    const double w0 = __dmul_rn(1.0 - x0, 1.0 - x1), w1 = __dmul_rn(x0, 1.0 - x1), w2 = __dmul_rn(x0, x1), w3 = __dmul_rn(1.0 - x0, x1);
    for (int d = 0; d < 2; ++d) {
      double s = __dmul_rn(xy[2 * v[0] + d], w0);
      s = __dadd_rn(s, __dmul_rn(xy[2 * v[1] + d], w1));
      s = __dadd_rn(s, __dmul_rn(xy[2 * v[2] + d], w2));
      s = __dadd_rn(s, __dmul_rn(xy[2 * v[3] + d], w3));
      xy_f[2 * static_cast<int64_t>(ctr) + d] = s;
    }
    for (int k = 0; k < 4; ++k) {
      ed[2 * k] = m[k];
      ed[2 * k + 1] = ctr;
    }
    const uint32_t q4[4][4] = {{v[0], m[0], ctr, m[3]}, {v[1], m[1], ctr, m[0]}, {v[2], m[1], ctr, m[2]}, {v[3], m[2], ctr, m[3]}};
    for (int k = 0; k < 4; ++k)
      for (int l = 0; l < 4; ++l) ch[4 * k + l] = q4[k][l];
  }
}
}  // namespace
}  // namespace lfgpu

extern "C" int lfgpu_mesh_refine_regular(lfgpu_ctx* ctx, lfgpu_mesh* parent, lfgpu_mesh** out) {
  if (ctx == nullptr || parent == nullptr || out == nullptr) return LFGPU_ERR_INVALID;
  *out = nullptr;
  if (parent->cell_coords != nullptr) LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "parent carries explicit cell corner coordinates: refine on the host and upload");
  int rc = ensure_topology(ctx, parent);
  if (rc != LFGPU_OK) return rc;
  cudaStream_t st = ctx->stream;
  const int64_t nn = parent->n_nodes, ne = parent->n_edges, nc = parent->n_cells, nt = parent->n_tria, nq = parent->n_quad;
  const int64_t nn_f = nn + ne + nq, ne_f = 2 * ne + 3 * nt + 4 * nq, nc_f = 4 * nc;
  lfgpu_mesh* m = nullptr;
  if ((rc = alloc_mesh(ctx, nn_f, nc_f, false, &m)) != LFGPU_OK) return rc;
  int32_t *is_quad = nullptr, *quads_before = nullptr;
  uint32_t* edges_f = nullptr;
  void* tmp = nullptr;
  auto cleanup = [&]() { cudaFree(is_quad); cudaFree(quads_before); cudaFree(edges_f); cudaFree(tmp); };
  cudaError_t e = cudaMalloc(&is_quad, sizeof(int32_t) * nc);
  if (e == cudaSuccess) e = cudaMalloc(&quads_before, sizeof(int32_t) * nc);
  if (e == cudaSuccess) e = cudaMalloc(&edges_f, sizeof(uint32_t) * 2 * ne_f);
  if (e == cudaSuccess) {
    k_ref_is_quad<<<static_cast<unsigned>(cdiv(nc, kThreads)), kThreads, 0, st>>>(nc, parent->cell_nodes, is_quad);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, is_quad, quads_before, static_cast<int>(nc), st);
    e = cudaMalloc(&tmp, tb);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tb, is_quad, quads_before, static_cast<int>(nc), st);
  }
  if (e == cudaSuccess) {
    k_ref_nodes_edges<<<static_cast<unsigned>(cdiv(nn + ne, kThreads)), kThreads, 0, st>>>(nn, ne, parent->node_coords, parent->edge_nodes,
                                                                                          m->node_coords, edges_f);
    k_ref_cells<<<static_cast<unsigned>(cdiv(nc, kThreads)), kThreads, 0, st>>>(nc, nn, ne, parent->cell_nodes, parent->cell_edges, quads_before,
                                                                               parent->node_coords, m->node_coords, edges_f, m->cell_nodes);
    ctx->launches += 3;
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    cleanup();
    lfgpu_mesh_destroy(m);
    LFGPU_FAIL(ctx, LFGPU_ERR_CUDA, std::string("mesh_refine_regular: ") + cudaGetErrorString(e));
  }
  rc = finish_mesh(ctx, m);
  if (rc == LFGPU_OK) rc = build_topology_impl(ctx, m, ne_f, nullptr, false, nullptr, edges_f);
  cleanup();
  if (rc != LFGPU_OK) {
    lfgpu_mesh_destroy(m);
    return rc;
  }
  *out = m;
  return LFGPU_OK;
}
