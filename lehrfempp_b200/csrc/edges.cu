// Edge (codim-1) contributions (product code) -- SURVEY.md section 8f row 2: impedance / Robin / Neumann boundary terms.
//
// Stands in for (paths relative to lib/lf/):
//   uscalfe/loc_comp_ellbvp.h:367-529   MassEdgeMatrixProvider::Eval      M_ab = sum_k (w_k |e|) gamma_k phi_a(k) phi_b(k)
//   uscalfe/loc_comp_ellbvp.h:784-921   ScalarLoadEdgeVectorProvider::Eval  v_a = sum_k (w_k |e|) g_k phi_a(k)
//   geometry/segment_o1.cc:9-35         Global, IntegrationElement of a straight edge
//   assemble/assembler.h:125-182, 306-326 with codim = 1: the loop over edges and the local -> global scatter
//   mesh/utils (flagEntitiesOnBoundary / CountNumSuperEntities(mesh, 1, 1)): edges with exactly one adjacent cell
//
// The reference adds these triplets to the same COOMatrix as the cell contributions, so the entries are already part
// of the pattern of the symbolic pass (both dofs of an edge belong to the adjacent cell): the kernel finds the slot by
// binary search in the row and adds with an FP64 atomic.  Edge terms are O(sqrt(N)) work on a boundary -- nothing here
// is performance critical; the point is that the matrix never leaves the device between the cell pass and the solve.
// When two active edges share a dof the order of their two additions is not fixed (last-bit differences only).
#include <cmath>

#include <algorithm>

#include "lfgpu_internal.cuh"

namespace lfgpu {
namespace {
constexpr int kThreads = 128;

struct EdgeCoeff {
  int kind;  // LFGPU_COEFF_CONST, PER_CELL (= per edge), PER_QP (per edge and quadrature point)
  double c;
  const double* data;
  long long stride;
};

__device__ __forceinline__ double eval_edge_coeff(const EdgeCoeff& G, int64_t e, int k) {
  switch (G.kind) {
    case LFGPU_COEFF_CONST: return G.c;
    case LFGPU_COEFF_PER_CELL: return __ldg(G.data + e);
    default: return __ldg(G.data + e * G.stride + k);
  }
}

// UniformFEDofHandler::GlobalDofIndices(edge) (dofhandler.cc:286-338): dofs of endpoint 0, endpoint 1, interior dofs
__device__ __forceinline__ int32_t edge_dof(int a, uint32_t n0, uint32_t n1, int64_t e, int n_seg, int64_t edge_base) {
  return a == 0 ? static_cast<int32_t>(n0) : (a == 1 ? static_cast<int32_t>(n1) : static_cast<int32_t>(edge_base + e * n_seg + (a - 2)));
}

__device__ __forceinline__ int find_slot(const int32_t* __restrict__ inner, int32_t lo, int32_t hi, int32_t key) {
  while (lo < hi) {
    const int32_t mid = lo + ((hi - lo) >> 1);  // lo + hi overflows int32 above 2^30 stored values
    const int32_t v = __ldg(inner + mid);
    if (v == key) return mid;
    if (v < key) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  return -1;
}

__device__ __forceinline__ double edge_length(const double* __restrict__ xy, uint32_t n0, uint32_t n1) {
  const double dx = xy[2 * n1] - xy[2 * n0], dy = xy[2 * n1 + 1] - xy[2 * n0 + 1];
  return sqrt(dx * dx + dy * dy);  // segment_o1.cc:31-35
}

__global__ void k_edge_mass(int64_t n_edges, const uint32_t* __restrict__ edge_nodes, const double* __restrict__ xy, SegTable T, EdgeCoeff G,
                            const uint8_t* __restrict__ active, int n_seg, int64_t edge_base, bool row_major,
                            const int32_t* __restrict__ outer, const int32_t* __restrict__ inner, double* __restrict__ values,
                            int* __restrict__ flags) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  if (active != nullptr && active[e] == 0) return;
  const uint32_t n0 = edge_nodes[2 * e], n1 = edge_nodes[2 * e + 1];
  const double len = edge_length(xy, n0, n1);
  const int nsf = T.nsf;
  double m[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) m[a][b] = 0.0;
  for (int k = 0; k < T.nq; ++k) {
    const double w = (T.w[k] * len) * eval_edge_coeff(G, e, k);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (a < nsf && b < nsf) m[a][b] += (T.phi[a * kMaxSegNq + k] * T.phi[b * kMaxSegNq + k]) * w;
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (a >= nsf) break;
    const int32_t da = edge_dof(a, n0, n1, e, n_seg, edge_base);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (b >= nsf) break;
      const int32_t db = edge_dof(b, n0, n1, e, n_seg, edge_base);
      const int32_t o = row_major ? da : db, i = row_major ? db : da;  // entry (row da, column db)
      const int slot = find_slot(inner, outer[o], outer[o + 1], i);
      if (slot < 0) {
        flags[0] = 1;
      } else {
        atomicAdd(values + slot, m[a][b]);
      }
    }
  }
}

__global__ void k_edge_load(int64_t n_edges, const uint32_t* __restrict__ edge_nodes, const double* __restrict__ xy, SegTable T, EdgeCoeff G,
                            const uint8_t* __restrict__ active, int n_seg, int64_t edge_base, double* __restrict__ vec) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  if (active != nullptr && active[e] == 0) return;
  const uint32_t n0 = edge_nodes[2 * e], n1 = edge_nodes[2 * e + 1];
  const double len = edge_length(xy, n0, n1);
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k = 0; k < T.nq; ++k) {
    const double w = (T.w[k] * len) * eval_edge_coeff(G, e, k);
#pragma unroll
    for (int a = 0; a < 4; ++a)
      if (a < T.nsf) v[a] += T.phi[a * kMaxSegNq + k] * w;
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
    if (a < T.nsf) atomicAdd(vec + edge_dof(a, n0, n1, e, n_seg, edge_base), v[a]);
}

// The same two operations for an explicit list of straight segments with their dofs -- the form a host caller that
// owns the DofHandler uses (it passes the ACTIVE edges only, typically a boundary part): seg_xy [n][4] = x0 y0 x1 y1,
// seg_dofs [n][nsf] = GlobalDofIndices(edge).
__global__ void k_segment_mass(int64_t n, const double* __restrict__ seg_xy, const int32_t* __restrict__ seg_dofs, SegTable T, EdgeCoeff G,
                               bool row_major, const int32_t* __restrict__ outer, const int32_t* __restrict__ inner, int64_t n_outer,
                               int64_t n_inner, double* __restrict__ values, int* __restrict__ flags) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const double dx = seg_xy[4 * e + 2] - seg_xy[4 * e], dy = seg_xy[4 * e + 3] - seg_xy[4 * e + 1];
  const double len = sqrt(dx * dx + dy * dy);
  const int nsf = T.nsf;
  for (int a = 0; a < nsf; ++a) {
    const int32_t da = seg_dofs[e * nsf + a];
    for (int b = 0; b < nsf; ++b) {
      const int32_t db = seg_dofs[e * nsf + b];
      double m = 0.0;
      for (int k = 0; k < T.nq; ++k) m += (T.phi[a * kMaxSegNq + k] * T.phi[b * kMaxSegNq + k]) * ((T.w[k] * len) * eval_edge_coeff(G, e, k));
      const int32_t o = row_major ? da : db, i = row_major ? db : da;
      if (o < 0 || o >= n_outer || i < 0 || i >= n_inner) {  // a dof outside the matrix: reported like a missing entry
        flags[0] = 1;
        continue;
      }
      const int slot = find_slot(inner, outer[o], outer[o + 1], i);
      if (slot < 0) {
        flags[0] = 1;
      } else {
        atomicAdd(values + slot, m);
      }
    }
  }
}
__global__ void k_segment_load(int64_t n, const double* __restrict__ seg_xy, const int32_t* __restrict__ seg_dofs, SegTable T, EdgeCoeff G,
                               int64_t n_dofs, double* __restrict__ vec, int* __restrict__ flags) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const double dx = seg_xy[4 * e + 2] - seg_xy[4 * e], dy = seg_xy[4 * e + 3] - seg_xy[4 * e + 1];
  const double len = sqrt(dx * dx + dy * dy);
  for (int a = 0; a < T.nsf; ++a) {
    double v = 0.0;
    for (int k = 0; k < T.nq; ++k) v += T.phi[a * kMaxSegNq + k] * ((T.w[k] * len) * eval_edge_coeff(G, e, k));
    const int32_t d = seg_dofs[e * T.nsf + a];
    if (d < 0 || d >= n_dofs) {
      flags[0] = 1;
    } else {
      atomicAdd(vec + d, v);
    }
  }
}

// SegmentO1::Global (segment_o1.cc:9-11): x = p1 * t + p0 * (1 - t)
__global__ void k_edge_qp_coords(int64_t n_edges, const uint32_t* __restrict__ edge_nodes, const double* __restrict__ xy, SegTable T,
                                 int nq_stride, double* __restrict__ out) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const uint32_t n0 = edge_nodes[2 * e], n1 = edge_nodes[2 * e + 1];
  for (int k = 0; k < nq_stride; ++k) {
    double X = 0.0, Y = 0.0;
    if (k < T.nq) {
      const double t = T.x[k];
      X = __dadd_rn(__dmul_rn(xy[2 * n1], t), __dmul_rn(xy[2 * n0], 1.0 - t));
      Y = __dadd_rn(__dmul_rn(xy[2 * n1 + 1], t), __dmul_rn(xy[2 * n0 + 1], 1.0 - t));
    }
    out[(e * nq_stride + k) * 2] = X;
    out[(e * nq_stride + k) * 2 + 1] = Y;
  }
}

__global__ void k_count_edge_cells(int64_t n_cells, const uint32_t* __restrict__ cell_nodes, const uint32_t* __restrict__ cell_edges,
                                   unsigned* __restrict__ count) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const int nv = cell_nodes[4 * c + 3] == LFGPU_IDX_NIL ? 3 : 4;
  for (int j = 0; j < nv; ++j) atomicAdd(count + cell_edges[4 * c + j], 1U);
}
__global__ void k_flag_boundary(int64_t n_edges, const unsigned* __restrict__ count, uint8_t* __restrict__ flags) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e < n_edges) flags[e] = count[e] == 1U ? 1 : 0;
}

// boundary flags of nodes (endpoints of boundary edges) and of the dofs of a uniform layout (dofhandler.cc:141-284: node
// dofs first, n_pt per node; then edge-interior dofs, n_seg per edge; cell-interior dofs are never on the boundary)
__global__ void k_boundary_nodes(int64_t n_edges, const uint32_t* __restrict__ edge_nodes, const uint8_t* __restrict__ edge_flags,
                                 uint8_t* __restrict__ node_flags) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges || edge_flags[e] == 0) return;
  node_flags[edge_nodes[2 * e]] = 1;  // several edges may write the same 1: benign
  node_flags[edge_nodes[2 * e + 1]] = 1;
}
__global__ void k_boundary_dofs(int64_t n_dofs, int64_t n_nodes, int64_t n_edges, int n_pt, int n_seg, const uint8_t* __restrict__ node_flags,
                                const uint8_t* __restrict__ edge_flags, uint8_t* __restrict__ dof_flags) {
  const int64_t d = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (d >= n_dofs) return;
  const int64_t node_dofs = n_nodes * n_pt, edge_dofs = n_edges * n_seg;
  uint8_t f = 0;
  if (d < node_dofs) f = node_flags[d / n_pt];
  else if (d < node_dofs + edge_dofs) f = edge_flags[(d - node_dofs) / n_seg];
  dof_flags[d] = f;
}

int prepare(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_dofmap* dofmap, int degree, const lfgpu_quad* qr, const lfgpu_coeff* coeff,
            SegTable* T, EdgeCoeff* G) {
  if (degree < 1 || degree > 3) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "degree must be 1, 2 or 3");
  if (mesh->cell_coords != nullptr) LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "edge terms need edge geometry = node positions (mesh carries explicit cell corners)");
  if (dofmap != nullptr) {
    if (dofmap->n_pt != 1 || dofmap->n_seg != degree - 1 || dofmap->n_nodes != mesh->n_nodes)
      LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "edge terms need the Lagrange dof layout of this degree built on the device (lfgpu_dofmap_lagrange)");
  }
  int rc = ensure_topology(ctx, mesh);
  if (rc != LFGPU_OK) return rc;
  std::string err;
  if ((rc = build_segment_table(degree, qr, T, &err)) != LFGPU_OK) LFGPU_FAIL(ctx, rc, err);
  if (coeff != nullptr) {
    if (coeff->kind != LFGPU_COEFF_CONST && coeff->kind != LFGPU_COEFF_PER_CELL && coeff->kind != LFGPU_COEFF_PER_QP)
      LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "edge coefficient must be CONST, PER_CELL (per edge) or PER_QP");
    if (coeff->kind != LFGPU_COEFF_CONST && coeff->data == nullptr) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "coefficient table missing");
    if (coeff->kind == LFGPU_COEFF_PER_QP && coeff->stride < T->nq) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "coefficient stride smaller than the number of quadrature points");
    G->kind = coeff->kind;
    G->c = coeff->c[0];
    G->data = coeff->data;
    G->stride = coeff->stride;
  }
  return LFGPU_OK;
}
}  // namespace
}  // namespace lfgpu

using namespace lfgpu;

extern "C" {

int lfgpu_assemble_edge_mass(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_dofmap* dofmap, const lfgpu_pattern* p, int degree,
                             const lfgpu_quad* qr_segment, const lfgpu_coeff* gamma, const uint8_t* active_edges, double* d_values) {
  if (ctx == nullptr || mesh == nullptr || dofmap == nullptr || p == nullptr || gamma == nullptr || d_values == nullptr) return LFGPU_ERR_INVALID;
  if (p->n_outer != dofmap->n_dofs || p->n_inner != dofmap->n_dofs) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "pattern does not belong to this dof map (square matrix expected)");
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  SegTable T;
  EdgeCoeff G{};
  int rc = prepare(ctx, mesh, dofmap, degree, qr_segment, gamma, &T, &G);
  if (rc != LFGPU_OK) return rc;
  if (mesh->n_edges == 0) return LFGPU_OK;
  int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 768);
  LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_flags, 0, sizeof(int), ctx->stream));
  k_edge_mass<<<static_cast<unsigned>(cdiv(mesh->n_edges, kThreads)), kThreads, 0, ctx->stream>>>(
      mesh->n_edges, mesh->edge_nodes, mesh->node_coords, T, G, active_edges, dofmap->n_seg, mesh->n_nodes, p->major == LFGPU_ROW_MAJOR,
      p->outer, p->inner, d_values, d_flags);
  LFGPU_LAUNCH_CHECK(ctx);
  int h = 0;
  LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(&h, d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (h != 0) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "an edge entry is missing from the pattern (pattern built from another dof map?)");
  return LFGPU_OK;
}

int lfgpu_assemble_edge_load(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_dofmap* dofmap, int degree, const lfgpu_quad* qr_segment,
                             const lfgpu_coeff* g, const uint8_t* active_edges, double* d_vec) {
  if (ctx == nullptr || mesh == nullptr || dofmap == nullptr || g == nullptr || d_vec == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  SegTable T;
  EdgeCoeff G{};
  int rc = prepare(ctx, mesh, dofmap, degree, qr_segment, g, &T, &G);
  if (rc != LFGPU_OK) return rc;
  if (mesh->n_edges == 0) return LFGPU_OK;
  k_edge_load<<<static_cast<unsigned>(cdiv(mesh->n_edges, kThreads)), kThreads, 0, ctx->stream>>>(
      mesh->n_edges, mesh->edge_nodes, mesh->node_coords, T, G, active_edges, dofmap->n_seg, mesh->n_nodes, d_vec);
  LFGPU_LAUNCH_CHECK(ctx);
  return LFGPU_OK;
}

int lfgpu_edge_qp_coords(lfgpu_ctx* ctx, lfgpu_mesh* mesh, int degree, const lfgpu_quad* qr_segment, int nq_stride, double* d_out) {
  if (ctx == nullptr || mesh == nullptr || d_out == nullptr || nq_stride < 1) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  SegTable T;
  int rc = prepare(ctx, mesh, nullptr, degree, qr_segment, nullptr, &T, nullptr);
  if (rc != LFGPU_OK) return rc;
  if (mesh->n_edges == 0) return LFGPU_OK;
  k_edge_qp_coords<<<static_cast<unsigned>(cdiv(mesh->n_edges, kThreads)), kThreads, 0, ctx->stream>>>(mesh->n_edges, mesh->edge_nodes,
                                                                                                      mesh->node_coords, T, nq_stride, d_out);
  LFGPU_LAUNCH_CHECK(ctx);
  return LFGPU_OK;
}

static int prepare_segments(lfgpu_ctx* ctx, int degree, const lfgpu_quad* qr, const lfgpu_coeff* coeff, SegTable* T, EdgeCoeff* G) {
  std::string err;
  const int rc = build_segment_table(degree, qr, T, &err);
  if (rc != LFGPU_OK) LFGPU_FAIL(ctx, rc, err);
  if (coeff->kind != LFGPU_COEFF_CONST && coeff->kind != LFGPU_COEFF_PER_CELL && coeff->kind != LFGPU_COEFF_PER_QP)
    LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "segment coefficient must be CONST, PER_CELL (per segment) or PER_QP");
  if (coeff->kind != LFGPU_COEFF_CONST && coeff->data == nullptr) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "coefficient table missing");
  if (coeff->kind == LFGPU_COEFF_PER_QP && coeff->stride < T->nq) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "coefficient stride smaller than the number of quadrature points");
  G->kind = coeff->kind;
  G->c = coeff->c[0];
  G->data = coeff->data;
  G->stride = coeff->stride;
  return LFGPU_OK;
}

int lfgpu_assemble_segment_mass(lfgpu_ctx* ctx, const lfgpu_pattern* p, int degree, const lfgpu_quad* qr_segment, int64_t n_segments,
                                const double* d_seg_xy, const int32_t* d_seg_dofs, const lfgpu_coeff* gamma, double* d_values) {
  if (ctx == nullptr || p == nullptr || gamma == nullptr || d_values == nullptr || n_segments < 0) return LFGPU_ERR_INVALID;
  if (n_segments == 0) return LFGPU_OK;
  if (d_seg_xy == nullptr || d_seg_dofs == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  SegTable T;
  EdgeCoeff G{};
  const int rc = prepare_segments(ctx, degree, qr_segment, gamma, &T, &G);
  if (rc != LFGPU_OK) return rc;
  int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 768);
  LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_flags, 0, sizeof(int), ctx->stream));
  k_segment_mass<<<static_cast<unsigned>(cdiv(n_segments, kThreads)), kThreads, 0, ctx->stream>>>(
      n_segments, d_seg_xy, d_seg_dofs, T, G, p->major == LFGPU_ROW_MAJOR, p->outer, p->inner, p->n_outer, p->n_inner, d_values, d_flags);
  LFGPU_LAUNCH_CHECK(ctx);
  int h = 0;
  LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(&h, d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (h != 0) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "a segment entry is missing from the pattern");
  return LFGPU_OK;
}

int lfgpu_assemble_segment_load(lfgpu_ctx* ctx, int degree, const lfgpu_quad* qr_segment, int64_t n_segments, const double* d_seg_xy,
                                const int32_t* d_seg_dofs, const lfgpu_coeff* g, int64_t n_dofs, double* d_vec) {
  if (ctx == nullptr || g == nullptr || d_vec == nullptr || n_segments < 0) return LFGPU_ERR_INVALID;
  if (n_segments == 0) return LFGPU_OK;
  if (d_seg_xy == nullptr || d_seg_dofs == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  SegTable T;
  EdgeCoeff G{};
  const int rc = prepare_segments(ctx, degree, qr_segment, g, &T, &G);
  if (rc != LFGPU_OK) return rc;
  int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 768);
  LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_flags, 0, sizeof(int), ctx->stream));
  k_segment_load<<<static_cast<unsigned>(cdiv(n_segments, kThreads)), kThreads, 0, ctx->stream>>>(n_segments, d_seg_xy, d_seg_dofs, T, G, n_dofs,
                                                                                                 d_vec, d_flags);
  LFGPU_LAUNCH_CHECK(ctx);
  int h = 0;
  LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(&h, d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  if (h != 0) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "a segment dof is outside the vector");
  return LFGPU_OK;
}

int lfgpu_mesh_boundary_edges(lfgpu_ctx* ctx, lfgpu_mesh* mesh, uint8_t* d_flags) {
  if (ctx == nullptr || mesh == nullptr || d_flags == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  int rc = ensure_topology(ctx, mesh);
  if (rc != LFGPU_OK) return rc;
  if (mesh->n_edges == 0) return LFGPU_OK;
  unsigned* count = nullptr;
  LFGPU_CUDA_CHECK(ctx, cudaMalloc(&count, sizeof(unsigned) * mesh->n_edges));
  cudaError_t e = cudaMemsetAsync(count, 0, sizeof(unsigned) * mesh->n_edges, ctx->stream);
  if (e == cudaSuccess) {
    k_count_edge_cells<<<static_cast<unsigned>(cdiv(mesh->n_cells, kThreads)), kThreads, 0, ctx->stream>>>(mesh->n_cells, mesh->cell_nodes,
                                                                                                          mesh->cell_edges, count);
    k_flag_boundary<<<static_cast<unsigned>(cdiv(mesh->n_edges, kThreads)), kThreads, 0, ctx->stream>>>(mesh->n_edges, count, d_flags);
    ctx->launches += 2;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(count);
  LFGPU_CUDA_CHECK(ctx, e);
  return LFGPU_OK;
}

// d_node_flags [n_nodes]: nodes on the boundary = endpoints of boundary edges (flagEntitiesOnBoundary(mesh, 2))
int lfgpu_mesh_boundary_nodes(lfgpu_ctx* ctx, lfgpu_mesh* mesh, uint8_t* d_node_flags) {
  if (ctx == nullptr || mesh == nullptr || d_node_flags == nullptr) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  int rc = ensure_topology(ctx, mesh);
  if (rc != LFGPU_OK) return rc;
  LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_node_flags, 0, mesh->n_nodes, ctx->stream));
  if (mesh->n_edges == 0) return LFGPU_OK;
  uint8_t* d_edge = nullptr;
  LFGPU_CUDA_CHECK(ctx, cudaMalloc(&d_edge, mesh->n_edges));
  rc = lfgpu_mesh_boundary_edges(ctx, mesh, d_edge);
  if (rc == LFGPU_OK) {
    k_boundary_nodes<<<static_cast<unsigned>(cdiv(mesh->n_edges, kThreads)), kThreads, 0, ctx->stream>>>(mesh->n_edges, mesh->edge_nodes, d_edge,
                                                                                                        d_node_flags);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      set_last_error(ctx, std::string("boundary nodes: ") + cudaGetErrorString(e));
      rc = LFGPU_ERR_CUDA;
    }
  }
  cudaFree(d_edge);
  return rc;
}

int lfgpu_dofmap_boundary_dofs(lfgpu_ctx* ctx, lfgpu_mesh* mesh, const lfgpu_dofmap* dofmap, uint8_t* d_dof_flags) {
  if (ctx == nullptr || mesh == nullptr || dofmap == nullptr || d_dof_flags == nullptr) return LFGPU_ERR_INVALID;
  if (dofmap->n_pt < 0 || dofmap->n_seg < 0 || dofmap->n_nodes != mesh->n_nodes)
    LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "boundary dofs need a dof map built by lfgpu_dofmap_uniform / _lagrange on this mesh");
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  int rc = ensure_topology(ctx, mesh);
  if (rc != LFGPU_OK) return rc;
  uint8_t *d_edge = nullptr, *d_node = nullptr;
  LFGPU_CUDA_CHECK(ctx, cudaMalloc(&d_edge, std::max<int64_t>(mesh->n_edges, 1)));
  cudaError_t e = cudaMalloc(&d_node, std::max<int64_t>(mesh->n_nodes, 1));
  if (e != cudaSuccess) {
    cudaFree(d_edge);
    LFGPU_CUDA_CHECK(ctx, e);
  }
  rc = lfgpu_mesh_boundary_edges(ctx, mesh, d_edge);
  if (rc == LFGPU_OK) e = cudaMemsetAsync(d_node, 0, std::max<int64_t>(mesh->n_nodes, 1), ctx->stream);
  if (rc == LFGPU_OK && e == cudaSuccess && mesh->n_edges > 0) {
    k_boundary_nodes<<<static_cast<unsigned>(cdiv(mesh->n_edges, kThreads)), kThreads, 0, ctx->stream>>>(mesh->n_edges, mesh->edge_nodes, d_edge, d_node);
    ctx->launches++;
  }
  if (rc == LFGPU_OK && e == cudaSuccess) {
    const int64_t n_edges_with_dofs = dofmap->n_seg > 0 ? mesh->n_edges : 0;
    k_boundary_dofs<<<static_cast<unsigned>(cdiv(dofmap->n_dofs, kThreads)), kThreads, 0, ctx->stream>>>(
        dofmap->n_dofs, mesh->n_nodes, n_edges_with_dofs, dofmap->n_pt, dofmap->n_seg > 0 ? dofmap->n_seg : 1, d_node, d_edge, d_dof_flags);
    ctx->launches++;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(d_edge);
  cudaFree(d_node);
  if (rc != LFGPU_OK) return rc;
  LFGPU_CUDA_CHECK(ctx, e);
  return LFGPU_OK;
}

}  // extern "C"
