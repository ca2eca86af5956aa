// Numeric pass (product code): fused element computation + scatter into the compressed matrix / load vector.
//
// Stands in for (paths relative to lib/lf/):
//   uscalfe/loc_comp_ellbvp.h:266-339   ReactionDiffusionElementMatrixProvider::Eval
//        A_K = sum_k w_k |det J_k| [ G_k^T (alpha_k G_k) + gamma_k phi_k phi_k^T ],  G_k = J_k^{-T} grad_hat(Phi)_k^T
//   uscalfe/loc_comp_ellbvp.h:691-746   ScalarLoadElementVectorProvider::Eval
//   geometry/tria_o1.cc:50-74, quad_o1.cc:61-158   Jacobian, inverse transposed, |det|, Global
//   assemble/assembler.h:125-182, 306-326          the cell loop and the local -> global scatter
//
// Design (DESIGN.md has the long version): the unit of work is one ROW of one element matrix, (cell, local index a).
//   * LFGPU_ALGO_ATOMIC : one thread per (cell, a); the row is added into the values with FP64 atomics through the
//                         scatter map of the symbolic pass.
//   * LFGPU_ALGO_GATHER : one thread per OUTER index (matrix row for CSR) walks the (cell, a) items of that dof in
//                         ascending cell order (the reference's summation order), accumulates in a private
//                         shared-memory strip and writes every stored value exactly once: deterministic, no atomics,
//                         no zero-fill, no read-modify-write of the value array.
// Both use the same element-row routine.  Affine cells with cell-wise constant coefficients take the reference-tensor
// route (5 FMAs per entry); everything else integrates per quadrature point (3 FMAs per entry and point).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <utility>
#include <vector>

#include "lfgpu_internal.cuh"

namespace lfgpu {
namespace {

struct DevCoeff {
  int kind;
  double c[4];
  const double* data;
  long long stride;
};

// compact shared-memory image of the reference tables of one cell type
struct TabView {
  int nsf, nq;
  const double *w, *qx, *qy, *phi, *gx, *gy;            // per-qp tables, phi[a * nq + k]
  const double *k00, *k01, *k10, *k11, *m, *l;          // reference tensors, [a * nsf + b]
};

struct Tables {
  // global-memory blob: [header ints][doubles...]; built on the host per call (a few KB)
  int nsf[2], nq[2];      // index 0 = tria, 1 = quad; nsf = 0 -> no rule / shape functions for that type
  int off[2];             // offset (in doubles) of each type's block
  int total;              // total doubles
};

__device__ __forceinline__ int block_doubles(int nsf, int nq) { return 3 * nq + 3 * nsf * nq + 5 * nsf * nsf + nsf; }

__device__ __forceinline__ TabView make_view(const double* base, int nsf, int nq) {
  TabView v;
  v.nsf = nsf;
  v.nq = nq;
  v.w = base;
  v.qx = v.w + nq;
  v.qy = v.qx + nq;
  v.phi = v.qy + nq;
  v.gx = v.phi + nsf * nq;
  v.gy = v.gx + nsf * nq;
  v.k00 = v.gy + nsf * nq;
  v.k01 = v.k00 + nsf * nsf;
  v.k10 = v.k01 + nsf * nsf;
  v.k11 = v.k10 + nsf * nsf;
  v.m = v.k11 + nsf * nsf;
  v.l = v.m + nsf * nsf;
  return v;
}

struct MeshView {
  const double* node_coords;
  const uint32_t* cell_nodes;
  const double* cell_coords;
};

struct CellGeom {
  double x[4], y[4];
  bool quad;
};

__device__ __forceinline__ CellGeom load_geom(const MeshView& mv, int64_t cell) {
  CellGeom g;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(mv.cell_nodes) + cell);
  g.quad = (v.w != LFGPU_IDX_NIL);
  if (mv.cell_coords != nullptr) {
    const double2* cc = reinterpret_cast<const double2*>(mv.cell_coords) + 4 * cell;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double2 p = __ldg(cc + k);
      g.x[k] = p.x;
      g.y[k] = p.y;
    }
  } else {
    const double2* nc = reinterpret_cast<const double2*>(mv.node_coords);
    const double2 p0 = __ldg(nc + v.x), p1 = __ldg(nc + v.y), p2 = __ldg(nc + v.z);
    g.x[0] = p0.x; g.y[0] = p0.y; g.x[1] = p1.x; g.y[1] = p1.y; g.x[2] = p2.x; g.y[2] = p2.y;
    if (g.quad) {
      const double2 p3 = __ldg(nc + v.w);
      g.x[3] = p3.x; g.y[3] = p3.y;
    } else {
      g.x[3] = 0.0; g.y[3] = 0.0;
    }
  }
  return g;
}

// Jacobian J = [j00 j01; j10 j11] at reference point (x0, x1): tria_o1.cc:57, quad_o1.cc:114-117
__device__ __forceinline__ void jacobian(const CellGeom& g, double x0, double x1, double& j00, double& j01, double& j10, double& j11) {
  if (!g.quad) {
    j00 = g.x[1] - g.x[0]; j01 = g.x[2] - g.x[0];
    j10 = g.y[1] - g.y[0]; j11 = g.y[2] - g.y[0];
  } else {
    j00 = (g.x[1] - g.x[0]) * (1.0 - x1) + (g.x[2] - g.x[3]) * x1;
    j10 = (g.y[1] - g.y[0]) * (1.0 - x1) + (g.y[2] - g.y[3]) * x1;
    j01 = (g.x[3] - g.x[0]) * (1.0 - x0) + (g.x[2] - g.x[1]) * x0;
    j11 = (g.y[3] - g.y[0]) * (1.0 - x0) + (g.y[2] - g.y[1]) * x0;
  }
}

// Geometry::Global: tria_o1.cc:70-74, quad_o1.cc:68-83
__device__ __forceinline__ void global_point(const CellGeom& g, double x0, double x1, double& X, double& Y) {
  if (!g.quad) {
    const double l0 = 1.0 - x0 - x1;
    X = g.x[0] * l0 + g.x[1] * x0 + g.x[2] * x1;
    Y = g.y[0] * l0 + g.y[1] * x0 + g.y[2] * x1;
  } else {
    const double a = (1.0 - x0) * (1.0 - x1), b = x0 * (1.0 - x1), c = x0 * x1, d = (1.0 - x0) * x1;
    X = g.x[0] * a + g.x[1] * b + g.x[2] * c + g.x[3] * d;
    Y = g.y[0] * a + g.y[1] * b + g.y[2] * c + g.y[3] * d;
  }
}

// 2x2 diffusion tensor at (cell, qp) as used by the row routine: transposed for row-major output (see header)
__device__ __forceinline__ void eval_alpha(const DevCoeff& A, int64_t cell, int k, bool transpose, double& a00, double& a01, double& a10, double& a11) {
  switch (A.kind) {
    case LFGPU_COEFF_CONST:
      a00 = a11 = A.c[0]; a01 = a10 = 0.0;
      break;
    case LFGPU_COEFF_CONST_2X2:
      a00 = A.c[0]; a01 = A.c[1]; a10 = A.c[2]; a11 = A.c[3];
      break;
    case LFGPU_COEFF_PER_CELL:
      a00 = a11 = __ldg(A.data + cell); a01 = a10 = 0.0;
      break;
    case LFGPU_COEFF_PER_QP:
      a00 = a11 = __ldg(A.data + cell * A.stride + k); a01 = a10 = 0.0;
      break;
    default: {  // PER_QP_2X2
      const double* p = A.data + (cell * A.stride + k) * 4;
      a00 = __ldg(p); a01 = __ldg(p + 1); a10 = __ldg(p + 2); a11 = __ldg(p + 3);
    }
  }
  if (transpose) {
    const double t = a01;
    a01 = a10;
    a10 = t;
  }
}
__device__ __forceinline__ double eval_scalar(const DevCoeff& G, int64_t cell, int k) {
  switch (G.kind) {
    case LFGPU_COEFF_CONST: return G.c[0];
    case LFGPU_COEFF_PER_CELL: return __ldg(G.data + cell);
    default: return __ldg(G.data + cell * G.stride + k);  // PER_QP
  }
}
__device__ __forceinline__ bool cellwise_const(const DevCoeff& c) { return c.kind <= LFGPU_COEFF_PER_CELL; }
__device__ __forceinline__ double fast_rcp(double x);

// Row `a` of the element matrix of `cell` (or column a when alpha is passed untransposed, see header): acc[b], b < nsf
// Row `a` of the element matrix from the cell metric and the reference tensors of the rule:
//   acc[b] = sum_ij M_ij Khat^{ji}[a][b] + gm Mhat[a][b]        (5 FMA per entry; 3 when M is symmetric and gm = 0)
template <int NSF>
__device__ __forceinline__ void tensor_row(const TabView& T, int a, double m00, double m01, double m10, double m11, double gm, bool sym,
                                           double (&acc)[NSF]) {
  const int nsf = T.nsf;
  const int row = a * nsf;
  if (sym) {  // warp-uniform branches
    if (gm == 0.0) {
#pragma unroll
      for (int b = 0; b < NSF; ++b) {
        if (b < nsf) acc[b] = m00 * T.k00[row + b] + m01 * (T.k10[row + b] + T.k01[row + b]) + m11 * T.k11[row + b];
      }
    } else {
#pragma unroll
      for (int b = 0; b < NSF; ++b) {
        if (b < nsf) acc[b] = m00 * T.k00[row + b] + m01 * (T.k10[row + b] + T.k01[row + b]) + m11 * T.k11[row + b] + gm * T.m[row + b];
      }
    }
  } else {
#pragma unroll
    for (int b = 0; b < NSF; ++b) {
      if (b < nsf) acc[b] = m00 * T.k00[row + b] + m01 * T.k10[row + b] + m10 * T.k01[row + b] + m11 * T.k11[row + b] + gm * T.m[row + b];
    }
  }
}

// The same with the reference tensors in the kernel's PARAMETER block (constant bank).  ncu on the P2 workload showed the item
// kernel limited by the L1/shared-memory pipe (84 % busy): 24-40 table loads per item came from shared memory plus a per-block
// copy of the tables; threads of a warp share `a` (items are ordered by rank, local index, dof), so the constant cache serves
// them as broadcasts and the LSU pipe is left to the accumulation rounds.  A by-value parameter instead of a module-wide
// __constant__ array: every launch carries its own copy, so two contexts (or streams) assembling different degrees or rules
// on one device cannot overwrite each other's tables, and nothing is copied from a host stack buffer asynchronously.
constexpr int kParamTensorNsf = 10;  // triangles up to FeLagrangeO3Tria: 5 tensors x 100 doubles = 4000 bytes
struct KTensors {
  double k[5 * kParamTensorNsf * kParamTensorNsf];  // k00 | k01 | k10 | k11 | m, each [nsf * nsf] row-major, packed for the nsf in use
};

template <int NSF>
__device__ __forceinline__ void tensor_row_const(const KTensors& KT, int nsf, int a, double m00, double m01, double m10, double m11, double gm,
                                                 bool sym, double (&acc)[NSF]) {
  const int nn = nsf * nsf;
  const int r0 = a * nsf;  // k00; then k01, k10, k11, m at multiples of nn
  if (sym) {
    if (gm == 0.0) {
#pragma unroll
      for (int b = 0; b < NSF; ++b) {
        if (b < nsf) acc[b] = m00 * KT.k[r0 + b] + m01 * (KT.k[r0 + 2 * nn + b] + KT.k[r0 + nn + b]) + m11 * KT.k[r0 + 3 * nn + b];
      }
    } else {
#pragma unroll
      for (int b = 0; b < NSF; ++b) {
        if (b < nsf)
          acc[b] = m00 * KT.k[r0 + b] + m01 * (KT.k[r0 + 2 * nn + b] + KT.k[r0 + nn + b]) + m11 * KT.k[r0 + 3 * nn + b] + gm * KT.k[r0 + 4 * nn + b];
      }
    }
  } else {
#pragma unroll
    for (int b = 0; b < NSF; ++b) {
      if (b < nsf)
        acc[b] = m00 * KT.k[r0 + b] + m01 * KT.k[r0 + 2 * nn + b] + m10 * KT.k[r0 + nn + b] + m11 * KT.k[r0 + 3 * nn + b] + gm * KT.k[r0 + 4 * nn + b];
    }
  }
}

// TENSOR_ONLY: the caller guarantees affine cells with cell-wise constant coefficients (no quadrature loop is compiled)
template <int NSF, bool TENSOR_ONLY = false>
__device__ __forceinline__ void element_row(const CellGeom& g, const TabView& T, int a, const DevCoeff& alpha, const DevCoeff& gamma,
                                            int64_t cell, bool transpose_alpha, double (&acc)[NSF]) {
#pragma unroll
  for (int b = 0; b < NSF; ++b) acc[b] = 0.0;
  const int nsf = T.nsf, nq = T.nq;
  if (TENSOR_ONLY || (!g.quad && cellwise_const(alpha) && cellwise_const(gamma))) {
    // affine cell, cell-wise constant coefficients: A_K = sum_ij M_ij Khat^{ji} + gamma |det| Mhat
    double j00, j01, j10, j11;
    jacobian(g, 0.0, 0.0, j00, j01, j10, j11);
    const double det = j00 * j11 - j01 * j10;
    const double adet = fabs(det), idet = fast_rcp(det);
    // Jinv = idet [j11 -j01; -j10 j00]
    const double i00 = j11 * idet, i01 = -j01 * idet, i10 = -j10 * idet, i11 = j00 * idet;
    double a00, a01, a10, a11;
    eval_alpha(alpha, cell, 0, transpose_alpha, a00, a01, a10, a11);
    // M = |det| Jinv A Jinv^T
    const double t00 = i00 * a00 + i01 * a10, t01 = i00 * a01 + i01 * a11;
    const double t10 = i10 * a00 + i11 * a10, t11 = i10 * a01 + i11 * a11;
    const double m00 = adet * (t00 * i00 + t01 * i01), m01 = adet * (t00 * i10 + t01 * i11);
    const double m10 = adet * (t10 * i00 + t11 * i01), m11 = adet * (t10 * i10 + t11 * i11);
    const double gm = adet * eval_scalar(gamma, cell, 0);
    tensor_row<NSF>(T, a, m00, m01, m10, m11, gm, alpha.kind != LFGPU_COEFF_CONST_2X2, acc);
    return;
  }
  if (TENSOR_ONLY) return;
  double j00, j01, j10, j11;
  if (!g.quad) jacobian(g, 0.0, 0.0, j00, j01, j10, j11);
  for (int k = 0; k < nq; ++k) {
    if (g.quad) jacobian(g, T.qx[k], T.qy[k], j00, j01, j10, j11);
    const double det = j00 * j11 - j01 * j10;
    const double wd = T.w[k] * fabs(det), idet = fast_rcp(det);
    const double i00 = j11 * idet, i01 = -j01 * idet, i10 = -j10 * idet, i11 = j00 * idet;
    double a00, a01, a10, a11;
    eval_alpha(alpha, cell, k, transpose_alpha, a00, a01, a10, a11);
    const double gxa = T.gx[a * nq + k], gya = T.gy[a * nq + k];
    // G_a = Jinv^T ghat_a ; u = A G_a ; s = wd * Jinv u
    const double Gx = i00 * gxa + i10 * gya, Gy = i01 * gxa + i11 * gya;
    const double ux = a00 * Gx + a01 * Gy, uy = a10 * Gx + a11 * Gy;
    const double sx = wd * (i00 * ux + i01 * uy), sy = wd * (i10 * ux + i11 * uy);
    const double mm = wd * eval_scalar(gamma, cell, k) * T.phi[a * nq + k];
#pragma unroll
    for (int b = 0; b < NSF; ++b) {
      if (b < nsf) acc[b] += sx * T.gx[b * nq + k] + sy * T.gy[b * nq + k] + mm * T.phi[b * nq + k];
    }
  }
}

// per-cell metric of an affine cell with cell-wise constant coefficients: M = |det| Jinv A Jinv^T and gamma |det|
__device__ __forceinline__ void cell_metric(const CellGeom& g, const DevCoeff& alpha, const DevCoeff& gamma, int64_t cell,
                                            bool transpose_alpha, double& m00, double& m01, double& m10, double& m11, double& gm) {
  double j00, j01, j10, j11;
  jacobian(g, 0.0, 0.0, j00, j01, j10, j11);
  const double det = j00 * j11 - j01 * j10;
  const double adet = fabs(det), idet = fast_rcp(det);
  const double i00 = j11 * idet, i01 = -j01 * idet, i10 = -j10 * idet, i11 = j00 * idet;
  double a00, a01, a10, a11;
  eval_alpha(alpha, cell, 0, transpose_alpha, a00, a01, a10, a11);
  const double t00 = i00 * a00 + i01 * a10, t01 = i00 * a01 + i01 * a11;
  const double t10 = i10 * a00 + i11 * a10, t11 = i10 * a01 + i11 * a11;
  m00 = adet * (t00 * i00 + t01 * i01);
  m01 = adet * (t00 * i10 + t01 * i11);
  m10 = adet * (t10 * i00 + t11 * i01);
  m11 = adet * (t10 * i10 + t11 * i11);
  gm = adet * eval_scalar(gamma, cell, 0);
}

// cooperative copy of the table blob into shared memory; returns views.
// what: bit 0 = triangle tables, bit 1 = quadrilateral tables, bit 2 = the per-quadrature-point part is needed too
// (a kernel that only meets affine cells with cell-wise constant coefficients reads the reference tensors only)
__device__ __forceinline__ void load_tables(const Tables& hdr, const double* __restrict__ blob, double* smem, TabView& tt, TabView& tq,
                                            int what = 7) {
  for (int ty = 0; ty < 2; ++ty) {
    if (!(what & (1 << ty)) || hdr.nsf[ty] == 0) continue;
    const int nsf = hdr.nsf[ty], nq = hdr.nq[ty];
    const int qp_len = 3 * nq + 3 * nsf * nq;
    const int begin = hdr.off[ty] + ((what & 4) ? 0 : qp_len);
    const int end = hdr.off[ty] + block_doubles(nsf, nq);
    for (int i = begin + threadIdx.x; i < end; i += blockDim.x) smem[i] = blob[i];
  }
  __syncthreads();
  tt = make_view(smem + hdr.off[0], hdr.nsf[0], hdr.nq[0]);
  tq = make_view(smem + hdr.off[1], hdr.nsf[1], hdr.nq[1]);
}

// Bank swizzle of the shared-memory image of a block's value range: rows of equal length L start L doubles apart, and
// L = 16 (e.g. the edge dofs of cubic triangles) would put the same slot of 16 rows into ONE bank.  XOR-ing the low four
// index bits with the next four is a bijection inside every aligned group of 16 and spreads such columns over all banks.
__device__ __forceinline__ int swz(int k) { return k ^ ((k >> 4) & 15); }

// 1/x for x != 0 of moderate magnitude (Jacobian determinants; degenerate cells are rejected at upload): hardware seed
// (MUFU.RCP64H) + two Newton steps, relative error ~1e-16, no slow path
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

template <int NSF, typename P>
__global__ void __launch_bounds__(256) k_assemble_atomic(Tables hdr, const double* __restrict__ blob, MeshView mv, int64_t n_cells,
                                                         int o_stride, int pos_row, const int32_t* __restrict__ o_dofs,
                                                         const uint8_t* __restrict__ o_nldof, const int32_t* __restrict__ outer,
                                                         const P* __restrict__ pos, DevCoeff alpha, DevCoeff gamma,
                                                         const uint8_t* __restrict__ active, bool transpose_alpha,
                                                         double* __restrict__ values, int* __restrict__ flags) {
  extern __shared__ double smem[];
  TabView tt, tq;
  load_tables(hdr, blob, smem, tt, tq);
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n_cells * o_stride) return;
  const int64_t cell = t / o_stride;
  const int a = static_cast<int>(t - cell * o_stride);
  if (a >= o_nldof[cell]) return;
  if (active != nullptr && active[cell] == 0) return;
  const CellGeom g = load_geom(mv, cell);
  const TabView& T = g.quad ? tq : tt;
  if (T.nsf == 0) {
    flags[0] = 1;  // no rule / shape functions for this cell type (loc_comp_ellbvp.h:273-287)
    return;
  }
  double acc[NSF];
  element_row<NSF>(g, T, a, alpha, gamma, cell, transpose_alpha, acc);
  const int32_t r = o_dofs[t];
  double* dst = values + outer[r];
  const P* pp = pos + t * pos_row;
#pragma unroll
  for (int b = 0; b < NSF; ++b) {
    if (b < T.nsf) atomicAdd(dst + pp[b], acc[b]);
  }
}

// Owner-computes kernel: one thread per outer index (matrix row for CSR).  The thread walks the (cell, a) items of its
// dof in ascending cell order and accumulates into ITS segment of a block-wide shared-memory image of the value range
// the block owns (segment offsets = block scan of the row lengths, i.e. the image is laid out exactly like the output).
// The block then streams the image to HBM as one contiguous, fully coalesced copy.  Every value is written exactly
// once: no atomics, no zero-fill, no read-modify-write of the value array.
template <int NSF, typename P, int THREADS, bool TENSOR_ONLY>
__global__ void __launch_bounds__(THREADS, 6) k_assemble_gather(Tables hdr, const double* __restrict__ blob, int table_mask, MeshView mv,
                                                                int64_t n_rows, int o_stride, int pos_row,
                                                                const int32_t* __restrict__ outer, const int32_t* __restrict__ adj_ptr,
                                                                const uint32_t* __restrict__ adj, const P* __restrict__ pos,
                                                                DevCoeff alpha, DevCoeff gamma, const uint8_t* __restrict__ active,
                                                                bool transpose_alpha, double beta, const int32_t* __restrict__ row_list,
                                                                double* __restrict__ values, int* __restrict__ flags) {
  extern __shared__ double smem[];
  __shared__ int32_t s_warp[THREADS / 32 + 1];
  __shared__ int32_t s_first_v0;
  TabView tt, tq;
  load_tables(hdr, blob, smem, tt, tq, table_mask);
  double* image = smem + ((hdr.total + 1) & ~1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * THREADS + threadIdx.x;
  const bool in_range = t < n_rows;
  const int64_t r = in_range ? (row_list != nullptr ? row_list[t] : t) : 0;
  int32_t v0 = 0, len = 0;
  if (in_range) {
    v0 = __ldg(outer + r);
    len = __ldg(outer + r + 1) - v0;
  }
  // exclusive block scan of the row lengths
  int32_t inc = len;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int32_t up = __shfl_up_sync(0xffffffffU, inc, d);
    if (lane >= d) inc += up;
  }
  if (lane == 31) s_warp[warp] = inc;
  if (threadIdx.x == 0) s_first_v0 = v0;
  __syncthreads();
  int32_t base = 0;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) base += (w < warp) ? s_warp[w] : 0;
  const int32_t off = base + inc - len;
  if (in_range) {
    for (int s = 0; s < len; ++s) image[swz(off + s)] = 0.0;
    const int32_t it1 = __ldg(adj_ptr + r + 1);
    for (int32_t it = __ldg(adj_ptr + r); it < it1; ++it) {
      const uint32_t item = __ldg(adj + it);
      const int64_t cell = item >> 4;
      const int a = static_cast<int>(item & 15U);
      if (active != nullptr && active[cell] == 0) continue;
      const CellGeom g = load_geom(mv, cell);
      const TabView& T = g.quad ? tq : tt;
      // slots of the row's entries: pos_row bytes (or shorts), fetched as 32-bit words
      constexpr int kWords = (NSF * static_cast<int>(sizeof(P)) + 3) / 4;
      uint32_t pw[kWords];
      const uint32_t* pp = reinterpret_cast<const uint32_t*>(pos + (cell * o_stride + a) * pos_row);
#pragma unroll
      for (int w = 0; w < kWords; ++w) pw[w] = __ldg(pp + w);
      double acc[NSF];
      element_row<NSF, TENSOR_ONLY>(g, T, a, alpha, gamma, cell, transpose_alpha, acc);
#pragma unroll
      for (int b = 0; b < NSF; ++b) {
        if (b < T.nsf) {
          const int slot = sizeof(P) == 1 ? static_cast<int>((pw[b >> 2] >> (8 * (b & 3))) & 0xffU)
                                          : static_cast<int>((pw[b >> 1] >> (16 * (b & 1))) & 0xffffU);
          image[swz(off + slot)] += acc[b];
        }
      }
    }
  }
  // do the block's rows form one contiguous value range?  (always without a row list; mostly with a partition's list)
  const int contiguous = __syncthreads_and(!in_range || v0 == s_first_v0 + off);
  if (contiguous) {
    int32_t total = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) total += s_warp[w];
    double* dst = values + s_first_v0;
    if (beta == 0.0) {
      for (int k = threadIdx.x; k < total; k += THREADS) dst[k] = image[swz(k)];
    } else {
      for (int k = threadIdx.x; k < total; k += THREADS) dst[k] = fma(beta, dst[k], image[swz(k)]);
    }
  } else if (in_range) {
    double* dst = values + v0;
    if (beta == 0.0) {
      for (int s = 0; s < len; ++s) dst[s] = image[swz(off + s)];
    } else {
      for (int s = 0; s < len; ++s) dst[s] = fma(beta, dst[s], image[swz(off + s)]);
    }
  }
  (void)flags;
}

// Per-cell metric table for meshes of affine cells with cell-wise constant coefficients, M = |det J| J^-1 A J^-T:
// (M00, M01, M11, gamma |det|) = 32 B when A is a scalar (M symmetric; the row routine never reads M10 then), else
// (M00, M01, M10, M11, gamma |det|, 0) = 48 B.  Computed once per numeric pass so that the 6..10 items of a cell neither
// repeat the geometry nor chase cell -> vertices -> coordinates.  The records are GATHERED by the items (one line per
// lane in the worst case), so their size is what the L1 pipe of the item kernel pays for.
__global__ void __launch_bounds__(256) k_cell_metric(MeshView mv, int64_t n_cells, DevCoeff alpha, DevCoeff gamma, bool transpose_alpha,
                                                     bool sym, double* __restrict__ out) {
  const int64_t cell = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  const CellGeom g = load_geom(mv, cell);
  double m00, m01, m10, m11, gm;
  cell_metric(g, alpha, gamma, cell, transpose_alpha, m00, m01, m10, m11, gm);
  if (sym) {
    double2* o = reinterpret_cast<double2*>(out) + 2 * cell;
    o[0] = make_double2(m00, m01);
    o[1] = make_double2(m11, gm);
  } else {
    double2* o = reinterpret_cast<double2*>(out) + 3 * cell;
    o[0] = make_double2(m00, m01);
    o[1] = make_double2(m10, m11);
    o[2] = make_double2(gm, 0.0);
  }
}

// Item-parallel owner-computes kernel (the default without a row list): one thread per ITEM (cell, a) of the block's
// outer indices, so the element rows of all items are computed concurrently and the work is balanced whatever the
// valence of a dof.  The items of one dof are then folded into the block's shared-memory image of its value range in
// rounds: round k adds the k-th item of every dof (at most one item per dof and round -> plain read-modify-write, no
// atomics; ascending cell order per entry = the reference's summation order, bitwise repeatable).  The image is laid
// out exactly like the output and leaves as one coalesced copy.
// MINB > 0 (opt-in, LFGPU_ITEMS_OCC=3; P1 quadrature route only = config C2): 3 CTAs per SM instead of 4 -- 80 registers and
// 12 bytes of spill stores instead of 64 registers and 80 (the ncu capture shows the kernel bound by LSU work per item, and
// spill traffic is LSU work); not yet measured
template <int NSF, typename P, bool TENSOR_ONLY, int MINB = 0>
__global__ void __launch_bounds__(kItemThreads, MINB > 0 ? MINB : (TENSOR_ONLY ? (NSF <= 6 ? 6 : 5) : 4)) k_assemble_items(Tables hdr, const double* __restrict__ blob, int table_mask, MeshView mv,
                                                           const int4* __restrict__ blk_hdr, int pos_row, const P* __restrict__ pos_item,
                                                           const uint2* __restrict__ item_sorted, DevCoeff alpha, DevCoeff gamma,
                                                           const uint8_t* __restrict__ active, bool transpose_alpha, double beta,
                                                           const double* __restrict__ cell_metric_tab, double* __restrict__ values,
                                                           bool const_tables, const __grid_constant__ KTensors KT) {
  extern __shared__ double smem[];
  const int tid = threadIdx.x;
  const int4 bh = __ldg(blk_hdr + blockIdx.x);
  const int32_t adj0 = bh.x, out0 = bh.y;
  const int n_items = bh.z & 0xffff, max_rank = bh.z >> 16, total = bh.w;
  // metric route with the reference tensors in constant memory: no shared-memory tables at all
  const bool ctab = TENSOR_ONLY && const_tables && cell_metric_tab != nullptr;
  TabView tt, tq;
  if (!ctab) load_tables(hdr, blob, smem, tt, tq, table_mask);  // contains a __syncthreads()
  double* image = ctab ? smem : smem + ((hdr.total + 1) & ~1);
  for (int k = tid; k < ((total + 15) & ~15); k += kItemThreads) image[k] = 0.0;
  // my item (threads are ordered by rank-in-dof, then dof: see k_item_perm)
  bool valid = tid < n_items;
  int rank = 0, off = 0, nsf = 0;
  double acc[NSF];
  constexpr int kWords = (NSF * static_cast<int>(sizeof(P)) + 3) / 4;
  uint32_t pw[kWords];
  if (valid) {
    const uint2 w = __ldg(item_sorted + adj0 + tid);
    rank = static_cast<int>(w.y >> 16);
    off = static_cast<int>(w.y & 0xffffU);
    const uint32_t item = w.x;
    const int64_t cell = item >> 4;
    const int a = static_cast<int>(item & 15U);
    if (active != nullptr && active[cell] == 0) {
      valid = false;
    } else {
      const uint32_t* pp = reinterpret_cast<const uint32_t*>(pos_item + (static_cast<int64_t>(adj0) + tid) * pos_row);
#pragma unroll
      for (int w = 0; w < kWords; ++w) pw[w] = __ldg(pp + w);
      if (TENSOR_ONLY && cell_metric_tab != nullptr) {
        // affine cells, cell-wise constant coefficients: the metric of every cell was computed once by k_cell_metric
        const bool sym = alpha.kind != LFGPU_COEFF_CONST_2X2;
        double m00, m01, m10, m11, gm;
        if (sym) {
          const double2* mp = reinterpret_cast<const double2*>(cell_metric_tab) + 2 * cell;
          const double2 ma = __ldg(mp), mb = __ldg(mp + 1);
          m00 = ma.x; m01 = ma.y; m10 = ma.y; m11 = mb.x; gm = mb.y;
        } else {
          const double2* mp = reinterpret_cast<const double2*>(cell_metric_tab) + 3 * cell;
          const double2 ma = __ldg(mp), mb = __ldg(mp + 1), mc = __ldg(mp + 2);
          m00 = ma.x; m01 = ma.y; m10 = mb.x; m11 = mb.y; gm = mc.x;
        }
        nsf = hdr.nsf[0];
        if (ctab) {
          tensor_row_const<NSF>(KT, nsf, a, m00, m01, m10, m11, gm, sym, acc);
        } else {
          tensor_row<NSF>(tt, a, m00, m01, m10, m11, gm, sym, acc);
        }
      } else {
        const CellGeom g = load_geom(mv, cell);
        const TabView& T = g.quad ? tq : tt;
        nsf = T.nsf;
        element_row<NSF, TENSOR_ONLY>(g, T, a, alpha, gamma, cell, transpose_alpha, acc);
      }
    }
  }
  __syncthreads();  // image zeroed
  for (int k = 0; k <= max_rank; ++k) {
    if (k > 0) __syncthreads();
    if (valid && rank == k) {
#pragma unroll
      for (int b = 0; b < NSF; ++b) {
        if (b < nsf) {
          const int slot = sizeof(P) == 1 ? static_cast<int>((pw[b >> 2] >> (8 * (b & 3))) & 0xffU)
                                          : static_cast<int>((pw[b >> 1] >> (16 * (b & 1))) & 0xffffU);
          image[swz(off + slot)] += acc[b];
        }
      }
    }
  }
  __syncthreads();
  double* dst = values + out0;
  if (beta == 0.0) {
    for (int k = tid; k < total; k += kItemThreads) dst[k] = image[swz(k)];
  } else {
    for (int k = tid; k < total; k += kItemThreads) dst[k] = fma(beta, dst[k], image[swz(k)]);
  }
}

// load vector: one thread per cell, FP64 atomics (assembler.h:322-324)
// STORE: entry b of the element vector goes to vec[dofs[cell * stride + b]] by a plain store -- first pass of the two-pass load vector:
// `dofs` is then the position table of k_load_positions, `vec` the item-ordered array k_load_sum_items adds up
template <int NSF, bool STORE = false>
__global__ void __launch_bounds__(256) k_load_atomic(Tables hdr, const double* __restrict__ blob, MeshView mv, int64_t n_cells, int stride,
                                                     const int32_t* __restrict__ dofs, const uint8_t* __restrict__ nldof, DevCoeff f,
                                                     const uint8_t* __restrict__ active, double* __restrict__ vec, int* __restrict__ flags) {
  extern __shared__ double smem[];
  TabView tt, tq;
  load_tables(hdr, blob, smem, tt, tq);
  const int64_t cell = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  if (active != nullptr && active[cell] == 0) {
    if (STORE) {  // an inactive cell contributes zeros: the second pass adds whole item ranges
      const int n0 = nldof[cell];
#pragma unroll
      for (int b = 0; b < NSF; ++b)
        if (b < n0) vec[dofs[cell * stride + b]] = 0.0;
    }
    return;
  }
  const CellGeom g = load_geom(mv, cell);
  const TabView& T = g.quad ? tq : tt;
  if (T.nsf == 0) {
    flags[0] = 1;
    return;
  }
  double acc[NSF];
#pragma unroll
  for (int b = 0; b < NSF; ++b) acc[b] = 0.0;
  double j00, j01, j10, j11;
  if (!g.quad) jacobian(g, 0.0, 0.0, j00, j01, j10, j11);
  for (int k = 0; k < T.nq; ++k) {
    if (g.quad) jacobian(g, T.qx[k], T.qy[k], j00, j01, j10, j11);
    const double s = T.w[k] * fabs(j00 * j11 - j01 * j10) * eval_scalar(f, cell, k);
#pragma unroll
    for (int b = 0; b < NSF; ++b) {
      if (b < T.nsf) acc[b] += s * T.phi[b * T.nq + k];
    }
  }
  if (STORE) {  // dofs = position of every (cell, local index) in the item list of its dof, stride = NSF
    const int n0 = nldof[cell];
#pragma unroll
    for (int b = 0; b < NSF; ++b)
      if (b < n0) vec[dofs[cell * stride + b]] = acc[b];
    return;
  }
  const int n = nldof[cell];
#pragma unroll
  for (int b = 0; b < NSF; ++b) {
    if (b < n) atomicAdd(vec + dofs[cell * stride + b], acc[b]);
  }
}

// Two-pass load vector.  k_load_positions (once per dof map): where entry (cell, a) of an element vector goes -- the index of the item
// (cell << 4 | a) in the gather list of its dof, i.e. items of one dof are adjacent and in ascending cell order.  First pass
// (k_load_atomic<NSF, true>): every element vector once, one thread per cell, stored there.  Second pass (k_load_sum_items): one
// thread per dof adds its run of entries front to back -- result[dof] += elem_vec[a] of assembler.h:322-324 in the reference's
// order, no atomics, bitwise repeatable, contiguous reads.  (k_load_gather evaluates geometry, source and shape functions once per
// ITEM, i.e. every element vector once per local dof: 2.5 x slower at config C3, 3 x at C4's size; a first two-pass version that
// kept the vectors cell by cell and gathered 8-byte entries out of 32-byte sectors was only 1.5 - 1.8 x faster than that.)
__global__ void k_load_positions(int64_t n_dofs, const int32_t* __restrict__ ptr, const uint32_t* __restrict__ items, int nsf,
                                 int32_t* __restrict__ pos) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_dofs) return;
  for (int32_t t = ptr[r]; t < ptr[r + 1]; ++t) {
    const uint32_t item = items[t];
    pos[static_cast<int64_t>(item >> 4) * nsf + (item & 15U)] = t;
  }
}

__global__ void __launch_bounds__(256) k_load_sum_items(int64_t n_dofs, const int32_t* __restrict__ ptr, const double* __restrict__ ev,
                                                        double beta, double* __restrict__ vec) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_dofs) return;
  double sum = (beta == 0.0) ? 0.0 : beta * vec[r];
  const int32_t t1 = __ldg(ptr + r + 1);
  for (int32_t t = __ldg(ptr + r); t < t1; ++t) sum += __ldg(ev + t);
  vec[r] = sum;
}

// load vector, owner-computes: one thread per dof walks its (cell, local index) items in ascending cell order and adds
// the entries phi_K[a] = sum_k w_k |det J_k| f(x_k) phi_a(x_k) in that order -- result[dof] += elem_vec[a] of
// assembler.h:322-324 with the additions in the reference's order: no atomics, no zero-fill, bitwise repeatable.
// Every cell is visited once per local dof; the load vector is a small part of the traffic of the path.
__global__ void __launch_bounds__(256) k_load_gather(Tables hdr, const double* __restrict__ blob, MeshView mv, int64_t n_dofs,
                                                     const int32_t* __restrict__ ptr, const uint32_t* __restrict__ items, DevCoeff f,
                                                     const uint8_t* __restrict__ active, double beta, double* __restrict__ vec,
                                                     int* __restrict__ flags, const int32_t* __restrict__ row_list) {
  extern __shared__ double smem[];
  TabView tt, tq;
  load_tables(hdr, blob, smem, tt, tq);
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t0 >= n_dofs) return;  // n_dofs = number of listed rows when a list is given
  const int64_t r = row_list != nullptr ? row_list[t0] : t0;
  double sum = (beta == 0.0) ? 0.0 : beta * vec[r];
  const int32_t t1 = ptr[r + 1];
  for (int32_t t = ptr[r]; t < t1; ++t) {
    const uint32_t item = __ldg(items + t);
    const int64_t cell = item >> 4;
    const int a = static_cast<int>(item & 15U);
    if (active != nullptr && active[cell] == 0) continue;
    const CellGeom g = load_geom(mv, cell);
    const TabView& T = g.quad ? tq : tt;
    if (T.nsf == 0) {
      flags[0] = 1;
      continue;
    }
    double j00, j01, j10, j11;
    if (!g.quad) jacobian(g, 0.0, 0.0, j00, j01, j10, j11);
    double e = 0.0;
    for (int k = 0; k < T.nq; ++k) {
      if (g.quad) jacobian(g, T.qx[k], T.qy[k], j00, j01, j10, j11);
      const double s = T.w[k] * fabs(j00 * j11 - j01 * j10) * eval_scalar(f, cell, k);
      e += s * T.phi[a * T.nq + k];
    }
    sum += e;
  }
  vec[r] = sum;
}

// LFGPU_COEFF_NODAL (values at the mesh nodes, interpolated by the P1 shape functions: sum over the cell's vertices of N_a(xhat)
// data[node_a], N_a = FeLagrangeO1Tria / Quad, lagr_fe.h:110-128, 258-284 -- what MeshFunctionFE of a P1 space evaluates to) is turned
// into a per-point table on the device before the numeric kernels run: out[cell * stride + k] = value at quadrature point k.  A
// first version evaluated it inside the quadrature loops; carrying the node numbers and the branch cost EVERY coefficient kind
// registers (generic load kernels 64 -> 79, i.e. 2 resident blocks instead of 4: 2.76 -> 4.28 ms at 1.0e8 triangles; up to 130
// bytes more spills in the generic matrix kernels).
__global__ void k_nodal_tabulate(Tables hdr, const double* __restrict__ blob, const uint32_t* __restrict__ cell_nodes, int64_t n_cells, int stride,
                                 const double* __restrict__ nodal, double* __restrict__ out) {
  extern __shared__ double smem[];
  TabView tt, tq;
  load_tables(hdr, blob, smem, tt, tq);
  const int64_t cell = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(cell_nodes) + cell);
  const bool quad = v.w != LFGPU_IDX_NIL;
  const TabView& T = quad ? tq : tt;
  const double f0 = __ldg(nodal + v.x), f1 = __ldg(nodal + v.y), f2 = __ldg(nodal + v.z), f3 = quad ? __ldg(nodal + v.w) : 0.0;
  for (int k = 0; k < stride; ++k) {
    double val = 0.0;
    if (k < T.nq) {
      const double x0 = T.qx[k], x1 = T.qy[k];
      val = quad ? (1.0 - x0) * (1.0 - x1) * f0 + x0 * (1.0 - x1) * f1 + x0 * x1 * f2 + (1.0 - x0) * x1 * f3
                 : (1.0 - x0 - x1) * f0 + x0 * f1 + x1 * f2;
    }
    out[cell * stride + k] = val;
  }
}

__global__ void k_qp_coords(Tables hdr, const double* __restrict__ blob, MeshView mv, int64_t n_cells, int nq_stride, double* __restrict__ out) {
  extern __shared__ double smem[];
  TabView tt, tq;
  load_tables(hdr, blob, smem, tt, tq);
  const int64_t cell = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  const CellGeom g = load_geom(mv, cell);
  const TabView& T = g.quad ? tq : tt;
  for (int k = 0; k < nq_stride; ++k) {
    double X = 0.0, Y = 0.0;
    if (k < T.nq) global_point(g, T.qx[k], T.qy[k], X, Y);
    out[(cell * nq_stride + k) * 2] = X;
    out[(cell * nq_stride + k) * 2 + 1] = Y;
  }
}

__global__ void k_and_keep(int64_t n, uint8_t* __restrict__ flag, const uint8_t* __restrict__ keep) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n && keep[i] == 0) flag[i] = 0;
}

__global__ void k_scale(int64_t n, double beta, double* v) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) v[i] *= beta;
}

// host: pack the tables of both cell types into one blob
struct HostTables {
  Tables hdr;
  std::vector<double> blob;
};

void pack_type(const FeTable& t, std::vector<double>& blob) {
  FeTensors ten;
  build_fe_tensors(t, &ten);
  const int nsf = t.nsf, nq = t.nq;
  auto push = [&](const double* p, int n) { blob.insert(blob.end(), p, p + n); };
  push(t.w, nq);
  push(t.qx, nq);
  push(t.qy, nq);
  push(t.phi, nsf * nq);
  push(t.gx, nsf * nq);
  push(t.gy, nsf * nq);
  push(ten.k00, nsf * nsf);
  push(ten.k01, nsf * nsf);
  push(ten.k10, nsf * nsf);
  push(ten.k11, nsf * nsf);
  push(ten.m, nsf * nsf);
  push(ten.l, nsf);
}

// rule selection as in loc_comp_ellbvp.h:210-263: both null -> default rules for both types; otherwise a type without
// a supplied rule has no PrecomputedScalarReferenceFiniteElement and Eval fails for such cells
int make_tables(lfgpu_ctx* ctx, int degree, const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad, HostTables* out) {
  const bool defaults = (qr_tria == nullptr && qr_quad == nullptr);
  out->blob.clear();
  for (int ty = 0; ty < 2; ++ty) {
    const lfgpu_quad* qr = ty == 0 ? qr_tria : qr_quad;
    out->hdr.off[ty] = static_cast<int>(out->blob.size());
    if (!defaults && qr == nullptr) {
      out->hdr.nsf[ty] = 0;
      out->hdr.nq[ty] = 0;
      continue;
    }
    FeTable t;
    std::string err;
    const int rc = build_fe_table(degree, ty == 0 ? 3 : 4, qr, &t, &err);
    if (rc < 0) LFGPU_FAIL(ctx, rc, err);
    out->hdr.nsf[ty] = t.nsf;
    out->hdr.nq[ty] = t.nq;
    pack_type(t, out->blob);
  }
  out->hdr.total = static_cast<int>(out->blob.size());
  return LFGPU_OK;
}

int check_coeff(lfgpu_ctx* ctx, const lfgpu_coeff* c, bool allow_tensor, const HostTables& ht, DevCoeff* out) {
  if (c == nullptr) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "coefficient descriptor is null");
  const int maxq = ht.hdr.nq[0] > ht.hdr.nq[1] ? ht.hdr.nq[0] : ht.hdr.nq[1];
  switch (c->kind) {
    case LFGPU_COEFF_CONST: break;
    case LFGPU_COEFF_CONST_2X2:
      if (!allow_tensor) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "this coefficient must be scalar valued");
      break;
    case LFGPU_COEFF_PER_CELL:
      if (c->data == nullptr) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "PER_CELL coefficient without data");
      break;
    case LFGPU_COEFF_NODAL:
      if (c->data == nullptr) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "NODAL coefficient without data");
      break;
    case LFGPU_COEFF_PER_QP_2X2:
      if (!allow_tensor) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "this coefficient must be scalar valued");
      [[fallthrough]];
    case LFGPU_COEFF_PER_QP:
      if (c->data == nullptr || c->stride < maxq) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "PER_QP coefficient: data null or stride < number of quadrature points");
      break;
    default: LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "unknown coefficient kind");
  }
  out->kind = c->kind;
  for (int i = 0; i < 4; ++i) out->c[i] = c->c[i];
  out->data = c->data;
  out->stride = c->stride;
  return LFGPU_OK;
}

struct DeviceBlob {
  double* d = nullptr;  // owned by the ctx table cache
};

// The tables are a few KB and rarely change between calls: keep the last few on the device, keyed by content, so that a
// numeric call is a single kernel launch (no allocation, no synchronisation).
int upload_blob(lfgpu_ctx* ctx, const HostTables& ht, DeviceBlob* b) {
  for (auto& e : ctx->table_cache) {
    if (e.host == ht.blob) {
      b->d = e.dev;
      return LFGPU_OK;
    }
  }
  if (ctx->table_cache.size() >= 8) {
    LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& e : ctx->table_cache) cudaFree(e.dev);
    ctx->table_cache.clear();
  }
  double* d = nullptr;
  LFGPU_CUDA_CHECK(ctx, cudaMalloc(&d, sizeof(double) * ht.blob.size()));
  ctx->table_cache.push_back({ht.blob, d});
  // source = the cache's own copy, which outlives the asynchronous transfer
  LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(d, ctx->table_cache.back().host.data(), sizeof(double) * ht.blob.size(), cudaMemcpyHostToDevice, ctx->stream));
  b->d = d;
  return LFGPU_OK;
}

// loc_comp_ellbvp.h:273-287: a cell type that occurs in the mesh but has no rule / shape functions is an error.
// Decided on the host from the mesh's cell-type counts, so the numeric call never has to synchronise.
int check_rules(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const HostTables& ht) {
  if ((mesh->n_tria > 0 && ht.hdr.nsf[0] == 0) || (mesh->n_quad > 0 && ht.hdr.nsf[1] == 0))
    LFGPU_FAIL(ctx, LFGPU_ERR_MISSING_RULE, "No local shape function information or no quadrature rule for a reference element type present in the mesh");
  return LFGPU_OK;
}

// LFGPU_COEFF_NODAL -> LFGPU_COEFF_PER_QP: tabulate the interpolated values at the quadrature points of the rule in use into a
// buffer owned by the context (slot 0: diffusion / source, slot 1: reaction), see k_nodal_tabulate.  Stream-ordered: the table is
// written right before the kernels of this call read it.
int resolve_nodal(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const HostTables& ht, DevCoeff* c, int slot) {
  if (c->kind != LFGPU_COEFF_NODAL) return LFGPU_OK;
  DeviceBlob blob;
  int rc = upload_blob(ctx, ht, &blob);
  if (rc != LFGPU_OK) return rc;
  const int stride = std::max(1, std::max(ht.hdr.nq[0], ht.hdr.nq[1]));
  const size_t need = static_cast<size_t>(mesh->n_cells) * stride;
  if (ctx->nodal_cap[slot] < need) {
    LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));  // an earlier call may still read the old table
    cudaFree(ctx->nodal_tab[slot]);
    ctx->nodal_tab[slot] = nullptr;
    ctx->nodal_cap[slot] = 0;
    LFGPU_CUDA_CHECK(ctx, cudaMalloc(&ctx->nodal_tab[slot], sizeof(double) * std::max<size_t>(need, 1)));
    ctx->nodal_cap[slot] = need;
  }
  const size_t tab_bytes = sizeof(double) * ((ht.hdr.total + 1) & ~1);
  k_nodal_tabulate<<<static_cast<unsigned>(cdiv(mesh->n_cells, 256)), 256, tab_bytes, ctx->stream>>>(ht.hdr, blob.d, mesh->cell_nodes, mesh->n_cells,
                                                                                                stride, c->data, ctx->nodal_tab[slot]);
  LFGPU_LAUNCH_CHECK(ctx);
  c->kind = LFGPU_COEFF_PER_QP;
  c->data = ctx->nodal_tab[slot];
  c->stride = stride;
  return LFGPU_OK;
}

template <int NSF, typename P>
int launch_matrix(lfgpu_ctx* ctx, const HostTables& ht, const double* d_blob, const MeshView& mv, const lfgpu_pattern* p,
                  const DevCoeff& alpha, const DevCoeff& gamma, const uint8_t* active, double beta, double* d_values, int algo,
                  int* d_flags, const int32_t* row_list, int64_t n_rows, bool has_tria, bool has_quad) {
  const bool transpose_alpha = (p->major == LFGPU_ROW_MAJOR);
  const size_t tab_bytes = sizeof(double) * ((ht.hdr.total + 1) & ~1);
  if (algo == LFGPU_ALGO_ATOMIC) {
    if (beta == 0.0) {
      LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_values, 0, sizeof(double) * p->nnz, ctx->stream));
    } else if (beta != 1.0) {
      k_scale<<<static_cast<unsigned>(cdiv(p->nnz, 256)), 256, 0, ctx->stream>>>(p->nnz, beta, d_values);
      LFGPU_LAUNCH_CHECK(ctx);
    }
    const int64_t n_threads = p->n_cells * p->o_stride;
    k_assemble_atomic<NSF, P><<<static_cast<unsigned>(cdiv(n_threads, 256)), 256, tab_bytes, ctx->stream>>>(
        ht.hdr, d_blob, mv, p->n_cells, p->o_stride, p->pos_row, p->o_dofs, p->o_nldof, p->outer, static_cast<const P*>(p->pos),
        alpha, gamma, active, transpose_alpha, d_values, d_flags);
    LFGPU_LAUNCH_CHECK(ctx);
  } else {
    // which tables the kernel will touch: cell types present in the mesh; the per-point part only if quadrature is needed
    int table_mask = (has_tria ? 1 : 0) | (has_quad ? 2 : 0);
    const bool cellwise = alpha.kind <= LFGPU_COEFF_PER_CELL && gamma.kind <= LFGPU_COEFF_PER_CELL;
    const bool tensor_only = cellwise && !has_quad;
    if (!tensor_only) table_mask |= 4;
    if (row_list == nullptr && p->blk_rows != nullptr) {
      const size_t smem_i = tab_bytes + sizeof(double) * static_cast<size_t>((p->max_item_block_nnz + 15) & ~15);
      if (smem_i <= 200 * 1024) {
        auto ki = tensor_only ? k_assemble_items<NSF, P, true> : k_assemble_items<NSF, P, false>;
        if constexpr (NSF <= 4) {
          static const bool occ3_env = [] { const char* e = std::getenv("LFGPU_ITEMS_OCC"); return e != nullptr && e[0] == '3'; }();
          if (occ3_env && !tensor_only) ki = k_assemble_items<NSF, P, false, 3>;
        }
        LFGPU_CUDA_CHECK(ctx, cudaFuncSetAttribute(ki, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_i)));
        const double* metric = nullptr;
        if (tensor_only && NSF > 3) {
          // P2 / P3: each cell has 6 / 10 items -- compute its metric once (scratch owned by the pattern, 48 B per cell)
          lfgpu_pattern* pm = const_cast<lfgpu_pattern*>(p);
          if (pm->cell_metric == nullptr) LFGPU_CUDA_CHECK(ctx, cudaMalloc(&pm->cell_metric, sizeof(double) * 6 * p->n_cells));
          k_cell_metric<<<static_cast<unsigned>(cdiv(p->n_cells, 256)), 256, 0, ctx->stream>>>(mv, p->n_cells, alpha, gamma, transpose_alpha,
                                                                                              alpha.kind != LFGPU_COEFF_CONST_2X2, pm->cell_metric);
          LFGPU_LAUNCH_CHECK(ctx);
          metric = pm->cell_metric;
        }
        // metric route: reference tensors through the constant cache (LFGPU_CONST_TABLES=0 keeps them in shared memory)
        static const bool ctab_env = [] { const char* e = std::getenv("LFGPU_CONST_TABLES"); return e == nullptr || e[0] != '0'; }();
        const int nsf0 = ht.hdr.nsf[0];
        const bool const_tables = ctab_env && metric != nullptr && nsf0 <= kParamTensorNsf;
        KTensors KT;
        if (const_tables) {
          const double* k00 = ht.blob.data() + ht.hdr.off[0] + 3 * ht.hdr.nq[0] + 3 * nsf0 * ht.hdr.nq[0];
          std::copy(k00, k00 + 5 * nsf0 * nsf0, KT.k);
        }
        ki<<<static_cast<unsigned>(p->n_item_blocks), kItemThreads, smem_i, ctx->stream>>>(
            ht.hdr, d_blob, table_mask, mv, static_cast<const int4*>(p->blk_hdr), p->pos_row, static_cast<const P*>(p->pos_item),
            static_cast<const uint2*>(p->item_sorted), alpha, gamma, active, transpose_alpha, beta, metric, d_values, const_tables, KT);
        LFGPU_LAUNCH_CHECK(ctx);
        return LFGPU_OK;
      }
    }
    // shared-memory image of the block's value range: exact maximum over the blocks of consecutive rows (symbolic pass)
    // or, with a row list, the safe bound threads * longest row; long rows (a vertex of high valence) take fewer threads per
    // block instead of failing
    const int64_t rows = row_list != nullptr ? n_rows : p->n_outer;
    auto launch = [&](auto kern, int threads) -> int {
      const int64_t image_len = row_list != nullptr ? static_cast<int64_t>(threads) * p->max_row_len
                                                    : (threads == 128 ? static_cast<int64_t>(p->max_block_nnz) : static_cast<int64_t>(threads) * p->max_row_len);
      const size_t smem = tab_bytes + sizeof(double) * static_cast<size_t>((image_len + 15) & ~15);
      if (smem > 200 * 1024) return 1;
      LFGPU_CUDA_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      if (rows > 0) {
        kern<<<static_cast<unsigned>(cdiv(rows, threads)), threads, smem, ctx->stream>>>(
            ht.hdr, d_blob, table_mask, mv, rows, p->o_stride, p->pos_row, p->outer, p->adj_ptr, p->adj, static_cast<const P*>(p->pos), alpha,
            gamma, active, transpose_alpha, beta, row_list, d_values, d_flags);
      }
      LFGPU_LAUNCH_CHECK(ctx);
      return LFGPU_OK;
    };
    int lr = tensor_only ? launch(k_assemble_gather<NSF, P, 128, true>, 128) : launch(k_assemble_gather<NSF, P, 128, false>, 128);
    if (lr == 1) lr = tensor_only ? launch(k_assemble_gather<NSF, P, 32, true>, 32) : launch(k_assemble_gather<NSF, P, 32, false>, 32);
    if (lr == 1) LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "rows too long for the gather kernel; use LFGPU_ALGO_ATOMIC");
    if (lr != LFGPU_OK) return lr;
  }
  return LFGPU_OK;
}

}  // namespace

int and_row_keep(lfgpu_ctx* ctx, int64_t n, uint8_t* d_flag, const uint8_t* d_keep) {
  if (d_keep == nullptr || n <= 0) return LFGPU_OK;
  k_and_keep<<<static_cast<unsigned>(cdiv(n, 256)), 256, 0, ctx->stream>>>(n, d_flag, d_keep);
  LFGPU_LAUNCH_CHECK(ctx);
  return LFGPU_OK;
}
}  // namespace lfgpu

using namespace lfgpu;

extern "C" {

int lfgpu_pattern_restrict_rows(lfgpu_ctx* ctx, lfgpu_pattern* p, const uint8_t* d_keep) {
  if (ctx == nullptr || p == nullptr || d_keep == nullptr) return LFGPU_ERR_INVALID;
  if (p->fan_state != 0 || p->p1h_state != 0 || p->p2_state != 0 || p->p3_state != 0)
    LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "lfgpu_pattern_restrict_rows must be called before the first numeric pass on the pattern");
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  if (p->row_keep == nullptr) LFGPU_CUDA_CHECK(ctx, cudaMalloc(&p->row_keep, p->n_outer > 0 ? p->n_outer : 1));
  LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(p->row_keep, d_keep, p->n_outer, cudaMemcpyDeviceToDevice, ctx->stream));
  return LFGPU_OK;
}

int lfgpu_assemble_reaction_diffusion(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* p, int degree,
                                      const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad, const lfgpu_coeff* alpha,
                                      const lfgpu_coeff* gamma, const uint8_t* active, double beta, double* d_values, int algo) {
  return lfgpu_assemble_reaction_diffusion_rows(ctx, mesh, p, degree, qr_tria, qr_quad, alpha, gamma, active, beta, d_values, algo,
                                                nullptr, 0);
}

int lfgpu_assemble_reaction_diffusion_rows(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* p, int degree,
                                           const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad, const lfgpu_coeff* alpha,
                                           const lfgpu_coeff* gamma, const uint8_t* active, double beta, double* d_values,
                                           int algo, const int32_t* d_row_list, int64_t n_rows) {
  return lfgpu::assemble_rd_impl(ctx, mesh, p, degree, qr_tria, qr_quad, alpha, gamma, active, beta, d_values, algo, d_row_list, n_rows,
                                 -1, nullptr);
}

int lfgpu_assemble_reaction_diffusion_range(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* p, int degree,
                                            const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad, const lfgpu_coeff* alpha,
                                            const lfgpu_coeff* gamma, double beta, double* d_values, int algo, int64_t row0,
                                            int64_t n_rows) {
  if (p == nullptr || row0 < 0 || n_rows < 0 || row0 + n_rows > p->n_outer) return LFGPU_ERR_INVALID;
  if (n_rows == 0) return LFGPU_OK;
  return lfgpu::assemble_rd_impl(ctx, mesh, p, degree, qr_tria, qr_quad, alpha, gamma, nullptr, beta, d_values, algo, nullptr, n_rows, row0,
                                 nullptr);
}
}  // extern "C"

// row0 >= 0 (with d_row_list == nullptr): only the contiguous outer range [row0, row0 + n_rows) -- fan path only.
// fan_query != nullptr: nothing is launched; *fan_query tells whether this call would run entirely in the fan kernel.
int lfgpu::assemble_rd_impl(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* p, int degree, const lfgpu_quad* qr_tria,
                            const lfgpu_quad* qr_quad, const lfgpu_coeff* alpha, const lfgpu_coeff* gamma, const uint8_t* active,
                            double beta, double* d_values, int algo, const int32_t* d_row_list, int64_t n_rows, int64_t row0,
                            int* fan_query) {
  if (fan_query != nullptr) *fan_query = 0;
  if (ctx == nullptr || mesh == nullptr || p == nullptr || d_values == nullptr) return LFGPU_ERR_INVALID;
  if (degree < 1 || degree > 3) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "degree must be 1, 2 or 3");
  if (p->n_cells != mesh->n_cells) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "pattern was built for another mesh");
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  HostTables ht;
  int rc = make_tables(ctx, degree, qr_tria, qr_quad, &ht);
  if (rc != LFGPU_OK) return rc;
  // the provider's element matrix must cover the dof tables (assembler.h:143-148)
  // (the table stride is max(tria, quad) even on a pure triangle mesh, dofhandler.cc:138)
  const int need = nsf_of(degree, 4);
  if (p->o_stride > need || p->i_stride > need) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "dof tables have more local dofs than the element matrix has rows (nrows mismatch)");
  DevCoeff da, dg;
  if ((rc = check_coeff(ctx, alpha, true, ht, &da)) != LFGPU_OK) return rc;
  if ((rc = check_coeff(ctx, gamma, false, ht, &dg)) != LFGPU_OK) return rc;
  if ((rc = check_rules(ctx, mesh, ht)) != LFGPU_OK) return rc;
  lfgpu_coeff ra = *alpha, rg = *gamma;  // what the row kernels and the calls for irregular rows below are handed
  if (fan_query == nullptr) {  // (a query launches nothing; a node-interpolated coefficient never runs in the fan kernel)
    if ((rc = resolve_nodal(ctx, mesh, ht, &da, 0)) != LFGPU_OK) return rc;
    if ((rc = resolve_nodal(ctx, mesh, ht, &dg, 1)) != LFGPU_OK) return rc;
    if (alpha->kind == LFGPU_COEFF_NODAL) {
      ra.kind = da.kind;
      ra.data = da.data;
      ra.stride = static_cast<int>(da.stride);
    }
    if (gamma->kind == LFGPU_COEFF_NODAL) {
      rg.kind = dg.kind;
      rg.data = dg.data;
      rg.stride = static_cast<int>(dg.stride);
    }
  }
  if (algo == LFGPU_ALGO_AUTO || algo == LFGPU_ALGO_FAN) {
    // P1 vertex-fan kernel: triangles only, constant coefficients, every cell active, square nodal dof table
    bool ok = degree == 1 && active == nullptr && da.kind <= LFGPU_COEFF_CONST_2X2 && dg.kind == LFGPU_COEFF_CONST &&
              mesh->n_quad == 0 && mesh->cell_coords == nullptr;
    double wsum = 0.0, m_diag = 0.0, m_off = 0.0;
    if (ok) {
      // reference mass tensor of the rule must not depend on the local vertex numbering (true for every symmetric rule)
      const int nq = ht.hdr.nq[0];
      const double* base = ht.blob.data() + ht.hdr.off[0];
      for (int k = 0; k < nq; ++k) wsum += base[k];
      const double* m = base + 3 * nq + 3 * 3 * nq + 4 * 9;
      m_diag = m[0];
      m_off = m[1];
      if (dg.c[0] != 0.0) {
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b)
            if (std::fabs(m[a * 3 + b] - (a == b ? m_diag : m_off)) > 1e-15) ok = false;
      }
    }
    if (ok) {
      if ((rc = p1_fan_prepare(ctx, mesh, const_cast<lfgpu_pattern*>(p))) != LFGPU_OK) return rc;
      ok = p->fan_state == 1 && ((d_row_list == nullptr && row0 < 0) || p->n_irregular == 0);
    }
    if (fan_query != nullptr) {
      *fan_query = (ok && p->n_irregular == 0) ? 1 : 0;
      return LFGPU_OK;
    }
    if (ok) {
      const bool tr = (p->major == LFGPU_ROW_MAJOR);
      double a[4] = {da.c[0], da.c[1], da.c[2], da.c[3]};
      const int tensor = da.kind == LFGPU_COEFF_CONST_2X2;
      if (tensor && tr) std::swap(a[1], a[2]);
      if (p->n_irregular > 0) {
        // rows that are not a single fan: generic gather kernel on exactly those rows
        rc = lfgpu_assemble_reaction_diffusion_rows(ctx, mesh, p, degree, qr_tria, qr_quad, alpha, gamma, nullptr, beta, d_values,
                                                    LFGPU_ALGO_GATHER, p->fan_irregular, p->n_irregular);
        if (rc != LFGPU_OK) return rc;
      }
      return p1_fan_launch(ctx, mesh, p, a, tensor, dg.c[0], wsum, m_diag, m_off, beta, d_row_list, n_rows, d_values, row0);
    }
    // P1 row kernel (assemble_p1h.cu): quadrilaterals / hybrid meshes, coefficients per cell or per quadrature point, activity
    // masks, cell corners that are not bitwise the node positions; rules invariant under the rotations of the reference cell
    static const bool p1h_env = [] { const char* e = std::getenv("LFGPU_P1_ROWS"); return e == nullptr || e[0] != '0'; }();
    if (degree == 1 && fan_query == nullptr && p1h_env) {
      FeTable ft, fq;
      std::string err;
      const bool dflt = qr_tria == nullptr && qr_quad == nullptr;
      const bool have_t = mesh->n_tria > 0, have_q = mesh->n_quad > 0;
      bool okh = true;
      if (have_t) okh = okh && (dflt || qr_tria != nullptr) && build_fe_table(1, 3, qr_tria, &ft, &err) == LFGPU_OK;
      if (have_q) okh = okh && (dflt || qr_quad != nullptr) && build_fe_table(1, 4, qr_quad, &fq, &err) == LFGPU_OK;
      okh = okh && p1h_rules_ok(have_t ? &ft : nullptr, have_q ? &fq : nullptr);
      if (okh) {
        if ((rc = p1h_prepare(ctx, mesh, const_cast<lfgpu_pattern*>(p))) != LFGPU_OK) return rc;
        okh = p->p1h_state == 1 && (d_row_list == nullptr || p->n_p1h_irregular == 0);
      }
      if (okh) {
        const int64_t r0 = row0 >= 0 ? row0 : 0, r1 = row0 >= 0 ? row0 + n_rows : p->n_outer;
        const auto& irr = p->p1h_irregular_host;
        const int64_t i0 = std::lower_bound(irr.begin(), irr.end(), r0) - irr.begin(), i1 = std::lower_bound(irr.begin(), irr.end(), r1) - irr.begin();
        if (i1 > i0) {
          rc = lfgpu_assemble_reaction_diffusion_rows(ctx, mesh, p, degree, qr_tria, qr_quad, &ra, &rg, active, beta, d_values,
                                                      LFGPU_ALGO_GATHER, p->p1h_irregular + i0, i1 - i0);
          if (rc != LFGPU_OK) return rc;
        }
        return p1h_launch(ctx, mesh, p, have_t ? &ft : nullptr, have_q ? &fq : nullptr, &ra, &rg, active, beta, d_row_list, n_rows, row0,
                          d_values);
      }
    }
    // P2 row kernels (assemble_p2.cu): triangles, constant coefficients, the provider's default rule (exact for P2, hence
    // independent of the local vertex numbering), every cell active, all rows, overwrite.  LFGPU_ALGO_FAN asks for them
    // explicitly; LFGPU_ALGO_AUTO takes them unless LFGPU_P2_ROWS=0.
    static const bool p2_env = [] { const char* e = std::getenv("LFGPU_P2_ROWS"); return e == nullptr || e[0] != '0'; }();
    if (degree == 2 && fan_query == nullptr && (algo == LFGPU_ALGO_FAN || p2_env) && active == nullptr &&
        d_row_list == nullptr && qr_tria == nullptr && qr_quad == nullptr && da.kind <= LFGPU_COEFF_CONST_2X2 &&
        dg.kind == LFGPU_COEFF_CONST && mesh->n_quad == 0 && ht.hdr.nsf[0] == 6) {
      if ((rc = p2_rows_prepare(ctx, mesh, const_cast<lfgpu_pattern*>(p))) != LFGPU_OK) return rc;
      if (p->p2_state == 1 && p->p2_cc == (mesh->cell_coords != nullptr)) {
        const bool tr = (p->major == LFGPU_ROW_MAJOR);
        double a[4] = {da.c[0], da.c[1], da.c[2], da.c[3]};
        const int tensor = da.kind == LFGPU_COEFF_CONST_2X2;
        if (tensor && tr) std::swap(a[1], a[2]);
        // all rows, or the contiguous range [row0, row0 + n_rows) of one GPU of a row-block partition
        const int64_t r0 = row0 >= 0 ? row0 : 0, r1 = row0 >= 0 ? row0 + n_rows : p->n_outer;
        const auto& irr = p->p2_irregular_host;
        const int64_t i0 = std::lower_bound(irr.begin(), irr.end(), r0) - irr.begin(), i1 = std::lower_bound(irr.begin(), irr.end(), r1) - irr.begin();
        if (i1 > i0) {
          rc = lfgpu_assemble_reaction_diffusion_rows(ctx, mesh, p, degree, qr_tria, qr_quad, alpha, gamma, nullptr, beta, d_values,
                                                      LFGPU_ALGO_GATHER, p->p2_irregular + i0, i1 - i0);
          if (rc != LFGPU_OK) return rc;
        }
        const int nq = ht.hdr.nq[0];
        const double* k00 = ht.blob.data() + ht.hdr.off[0] + 3 * nq + 3 * 6 * nq;  // pack_type: w qx qy | phi gx gy | k00 k01 k10 k11 m
        return p2_rows_launch(ctx, mesh, p, a, tensor, dg.c[0], k00, k00 + 36, k00 + 72, k00 + 108, k00 + 144, d_values, r0, r1, beta);
      }
    }
    // P3 row kernels (assemble_p3.cu), same conditions (their host/device core is also checked against the oracle on the CPU);
    // LFGPU_P3_ROWS=0 keeps LFGPU_ALGO_AUTO on the item kernel.
    static const bool p3_env = [] { const char* e = std::getenv("LFGPU_P3_ROWS"); return e == nullptr || e[0] != '0'; }();
    if (degree == 3 && fan_query == nullptr && (algo == LFGPU_ALGO_FAN || p3_env) && active == nullptr &&
        d_row_list == nullptr && qr_tria == nullptr && qr_quad == nullptr && da.kind <= LFGPU_COEFF_CONST_2X2 &&
        dg.kind == LFGPU_COEFF_CONST && mesh->n_quad == 0 && ht.hdr.nsf[0] == 10) {
      if ((rc = p3_rows_prepare(ctx, mesh, const_cast<lfgpu_pattern*>(p))) != LFGPU_OK) return rc;
      if (p->p3_state == 1 && p->p3_cc == (mesh->cell_coords != nullptr)) {
        const bool tr = (p->major == LFGPU_ROW_MAJOR);
        double a[4] = {da.c[0], da.c[1], da.c[2], da.c[3]};
        const int tensor = da.kind == LFGPU_COEFF_CONST_2X2;
        if (tensor && tr) std::swap(a[1], a[2]);
        const int64_t r0 = row0 >= 0 ? row0 : 0, r1 = row0 >= 0 ? row0 + n_rows : p->n_outer;
        const auto& irr = p->p3_irregular_host;
        const int64_t i0 = std::lower_bound(irr.begin(), irr.end(), r0) - irr.begin(), i1 = std::lower_bound(irr.begin(), irr.end(), r1) - irr.begin();
        if (i1 > i0) {
          rc = lfgpu_assemble_reaction_diffusion_rows(ctx, mesh, p, degree, qr_tria, qr_quad, alpha, gamma, nullptr, beta, d_values,
                                                      LFGPU_ALGO_GATHER, p->p3_irregular + i0, i1 - i0);
          if (rc != LFGPU_OK) return rc;
        }
        const int nq = ht.hdr.nq[0];
        const double* k00 = ht.blob.data() + ht.hdr.off[0] + 3 * nq + 3 * 10 * nq;
        return p3_rows_launch(ctx, mesh, p, a, tensor, dg.c[0], k00, k00 + 100, k00 + 200, k00 + 300, k00 + 400, d_values, r0, r1, beta);
      }
    }
    if (algo == LFGPU_ALGO_FAN) LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "LFGPU_ALGO_FAN needs P1, P2 or P3 on a triangle mesh with constant coefficients and no activity mask");
    algo = LFGPU_ALGO_GATHER;
  }
  if (fan_query != nullptr) return LFGPU_OK;
  if (row0 >= 0) LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED, "contiguous row ranges are a fan-kernel feature");
  if (algo != LFGPU_ALGO_ATOMIC && algo != LFGPU_ALGO_GATHER) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "unknown algo");
  if (d_row_list != nullptr && algo != LFGPU_ALGO_GATHER) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "a row list needs LFGPU_ALGO_GATHER");
  if ((rc = check_rules(ctx, mesh, ht)) != LFGPU_OK) return rc;
  DeviceBlob blob;
  if ((rc = upload_blob(ctx, ht, &blob)) != LFGPU_OK) return rc;
  int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 128);
  const MeshView mv{mesh->node_coords, mesh->cell_nodes, mesh->cell_coords};
  const bool has_quads = mesh->n_quad > 0;
#define LFGPU_DISPATCH(NSF)                                                                                                      \
  rc = (p->pos_bytes == 1) ? launch_matrix<NSF, uint8_t>(ctx, ht, blob.d, mv, p, da, dg, active, beta, d_values, algo, d_flags, d_row_list, n_rows, mesh->n_tria > 0, mesh->n_quad > 0)  \
                           : launch_matrix<NSF, uint16_t>(ctx, ht, blob.d, mv, p, da, dg, active, beta, d_values, algo, d_flags, d_row_list, n_rows, mesh->n_tria > 0, mesh->n_quad > 0)
  switch (degree) {
    case 1: if (has_quads) { LFGPU_DISPATCH(4); } else { LFGPU_DISPATCH(3); } break;
    case 2: if (has_quads) { LFGPU_DISPATCH(9); } else { LFGPU_DISPATCH(6); } break;
    default: if (has_quads) { LFGPU_DISPATCH(16); } else { LFGPU_DISPATCH(10); } break;
  }
#undef LFGPU_DISPATCH
  return rc;
}

extern "C" {

int lfgpu_assemble_load(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_dofmap* dofmap, int degree, const lfgpu_quad* qr_tria,
                        const lfgpu_quad* qr_quad, const lfgpu_coeff* f, const uint8_t* active, double beta, double* d_vec, int algo) {
  if (ctx == nullptr || mesh == nullptr || dofmap == nullptr || d_vec == nullptr) return LFGPU_ERR_INVALID;
  if (degree < 1 || degree > 3) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "degree must be 1, 2 or 3");
  if (dofmap->n_cells != mesh->n_cells) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "dofmap was built for another mesh");
  if (algo != LFGPU_ALGO_AUTO && algo != LFGPU_ALGO_ATOMIC && algo != LFGPU_ALGO_GATHER)
    LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "load vector: algo must be AUTO, ATOMIC or GATHER");
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  HostTables ht;
  int rc = make_tables(ctx, degree, qr_tria, qr_quad, &ht);
  if (rc != LFGPU_OK) return rc;
  DevCoeff df;
  if ((rc = check_coeff(ctx, f, false, ht, &df)) != LFGPU_OK) return rc;
  if ((rc = check_rules(ctx, mesh, ht)) != LFGPU_OK) return rc;
  if ((rc = resolve_nodal(ctx, mesh, ht, &df, 0)) != LFGPU_OK) return rc;
  DeviceBlob blob;
  if ((rc = upload_blob(ctx, ht, &blob)) != LFGPU_OK) return rc;
  int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 128);
  if (algo == LFGPU_ALGO_GATHER) {
    // deterministic owner-computes variant (AUTO stays on the atomic kernel until the two are measured side by side)
    if ((rc = dofmap_gather_plan(ctx, dofmap)) != LFGPU_OK) return rc;
    const MeshView mvg{mesh->node_coords, mesh->cell_nodes, mesh->cell_coords};
    const size_t tbg = sizeof(double) * ((ht.hdr.total + 1) & ~1);
    k_load_gather<<<static_cast<unsigned>(cdiv(dofmap->n_dofs, 256)), 256, tbg, ctx->stream>>>(
        ht.hdr, blob.d, mvg, dofmap->n_dofs, dofmap->g_ptr, dofmap->g_items, df, active, beta, d_vec, d_flags, nullptr);
    LFGPU_LAUNCH_CHECK(ctx);
    return LFGPU_OK;
  }
  // P1 triangles, constant source, every cell active: the vertex-ring kernel (assemble_p1.cu), deterministic, no atomics;
  // rows that are not a single fan go through the gather kernel.  LFGPU_LOAD_FAN=0 keeps the atomic kernel.
  static const bool lfan_env = [] { const char* e = std::getenv("LFGPU_LOAD_FAN"); return e == nullptr || e[0] != '0'; }();
  const bool tabulated = df.kind == LFGPU_COEFF_PER_CELL || df.kind == LFGPU_COEFF_PER_QP;
  if (algo == LFGPU_ALGO_AUTO && lfan_env && degree == 1 && (df.kind == LFGPU_COEFF_CONST || tabulated) && active == nullptr &&
      mesh->n_quad == 0 && mesh->cell_coords == nullptr && ht.hdr.nsf[0] == 3) {
    const int nq = ht.hdr.nq[0];
    const double* l = ht.blob.data() + ht.hdr.off[0] + 3 * nq + 3 * 3 * nq + 5 * 9;  // pack_type: ... k00 k01 k10 k11 m | l
    if (tabulated ? nq <= 4 : (std::fabs(l[1] - l[0]) <= 1e-15 && std::fabs(l[2] - l[0]) <= 1e-15)) {
      int handled = 0;
      if (tabulated) {
        // w_q phi_a(x_q) of the rule in use (pack_type: w qx qy | phi[a][q] ...)
        const double* w = ht.blob.data() + ht.hdr.off[0];
        const double* phi = w + 3 * nq;
        double wtab[12];
        for (int a = 0; a < 3; ++a)
          for (int q = 0; q < nq; ++q) wtab[a * nq + q] = w[q] * phi[a * nq + q];
        rc = p1_load_fan(ctx, mesh, dofmap, 0.0, beta, d_vec, &handled, df.data, df.kind == LFGPU_COEFF_PER_CELL ? 1 : df.stride, nq, wtab);
      } else {
        rc = p1_load_fan(ctx, mesh, dofmap, df.c[0] * l[0], beta, d_vec, &handled);
      }
      if (rc != LFGPU_OK) return rc;
      if (handled) {
        if (dofmap->n_lv_irregular > 0) {
          const MeshView mvg{mesh->node_coords, mesh->cell_nodes, mesh->cell_coords};
          const size_t tbg = sizeof(double) * ((ht.hdr.total + 1) & ~1);
          k_load_gather<<<static_cast<unsigned>(cdiv(dofmap->n_lv_irregular, 256)), 256, tbg, ctx->stream>>>(
              ht.hdr, blob.d, mvg, dofmap->n_lv_irregular, dofmap->g_ptr, dofmap->g_items, df, active, beta, d_vec, d_flags,
              dofmap->lv_irregular);
          LFGPU_LAUNCH_CHECK(ctx);
        }
        return LFGPU_OK;
      }
    }
  }
  const MeshView mv{mesh->node_coords, mesh->cell_nodes, mesh->cell_coords};
  const size_t tab_bytes = sizeof(double) * ((ht.hdr.total + 1) & ~1);
  const unsigned grid = static_cast<unsigned>(cdiv(mesh->n_cells, 256));
  const bool has_quads = mesh->n_quad > 0;
  // AUTO: two passes (see k_load_positions) -- every element vector once, one thread per cell, stored in item order; then one thread per
  // dof adds its run: deterministic like GATHER, every cell evaluated once like ATOMIC.  LFGPU_LOAD_TWOPASS=0: atomics.
  static const bool twopass_env = [] { const char* e = std::getenv("LFGPU_LOAD_TWOPASS"); return e == nullptr || e[0] != '0'; }();
  if (algo == LFGPU_ALGO_AUTO && twopass_env) {
    const int nsf = degree == 1 ? (has_quads ? 4 : 3) : degree == 2 ? (has_quads ? 9 : 6) : (has_quads ? 16 : 10);
    if ((rc = dofmap_gather_plan(ctx, dofmap)) != LFGPU_OK) return rc;
    auto* dmut = const_cast<lfgpu_dofmap*>(dofmap);
    if (dmut->lv_ev == nullptr || dmut->lv_ev_stride != nsf) {
      cudaFree(dmut->lv_ev);
      cudaFree(dmut->lv_pos);
      dmut->lv_ev = nullptr;
      dmut->lv_pos = nullptr;
      int32_t n_items = 0;
      LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(&n_items, dofmap->g_ptr + dofmap->n_dofs, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
      LFGPU_CUDA_CHECK(ctx, cudaMalloc(&dmut->lv_ev, sizeof(double) * std::max<int64_t>(n_items, 1)));
      LFGPU_CUDA_CHECK(ctx, cudaMalloc(&dmut->lv_pos, sizeof(int32_t) * static_cast<size_t>(nsf) * mesh->n_cells));
      k_load_positions<<<static_cast<unsigned>(cdiv(dofmap->n_dofs, 256)), 256, 0, ctx->stream>>>(dofmap->n_dofs, dofmap->g_ptr, dofmap->g_items, nsf,
                                                                                                 dmut->lv_pos);
      LFGPU_LAUNCH_CHECK(ctx);
      dmut->lv_ev_stride = nsf;
    }
#define LFGPU_LOAD_EV(NSF)                                                                                                   \
  k_load_atomic<NSF, true><<<grid, 256, tab_bytes, ctx->stream>>>(ht.hdr, blob.d, mv, mesh->n_cells, NSF, dmut->lv_pos, dofmap->n_ldof, df, active, dmut->lv_ev, d_flags)
    switch (nsf) {
      case 3: LFGPU_LOAD_EV(3); break;
      case 4: LFGPU_LOAD_EV(4); break;
      case 6: LFGPU_LOAD_EV(6); break;
      case 9: LFGPU_LOAD_EV(9); break;
      case 10: LFGPU_LOAD_EV(10); break;
      default: LFGPU_LOAD_EV(16); break;
    }
#undef LFGPU_LOAD_EV
    LFGPU_LAUNCH_CHECK(ctx);
    k_load_sum_items<<<static_cast<unsigned>(cdiv(dofmap->n_dofs, 256)), 256, 0, ctx->stream>>>(dofmap->n_dofs, dofmap->g_ptr, dmut->lv_ev, beta, d_vec);
    LFGPU_LAUNCH_CHECK(ctx);
    return LFGPU_OK;
  }
  if (beta == 0.0) {
    LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_vec, 0, sizeof(double) * dofmap->n_dofs, ctx->stream));
  } else if (beta != 1.0) {
    k_scale<<<static_cast<unsigned>(cdiv(dofmap->n_dofs, 256)), 256, 0, ctx->stream>>>(dofmap->n_dofs, beta, d_vec);
    LFGPU_LAUNCH_CHECK(ctx);
  }
#define LFGPU_LOAD(NSF)                                                                                                         \
  k_load_atomic<NSF, false><<<grid, 256, tab_bytes, ctx->stream>>>(ht.hdr, blob.d, mv, mesh->n_cells, dofmap->stride, dofmap->cell_dofs, dofmap->n_ldof, df, active, d_vec, d_flags)
  switch (degree) {
    case 1: if (has_quads) { LFGPU_LOAD(4); } else { LFGPU_LOAD(3); } break;
    case 2: if (has_quads) { LFGPU_LOAD(9); } else { LFGPU_LOAD(6); } break;
    default: if (has_quads) { LFGPU_LOAD(16); } else { LFGPU_LOAD(10); } break;
  }
#undef LFGPU_LOAD
  LFGPU_LAUNCH_CHECK(ctx);
  return LFGPU_OK;
}

int lfgpu_qp_coords(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, int degree, const lfgpu_quad* qr_tria, const lfgpu_quad* qr_quad,
                    int nq_stride, double* d_out) {
  if (ctx == nullptr || mesh == nullptr || d_out == nullptr || nq_stride < 1) return LFGPU_ERR_INVALID;
  LFGPU_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  HostTables ht;
  int rc = make_tables(ctx, degree, qr_tria, qr_quad, &ht);
  if (rc != LFGPU_OK) return rc;
  DeviceBlob blob;
  if ((rc = upload_blob(ctx, ht, &blob)) != LFGPU_OK) return rc;
  const MeshView mv{mesh->node_coords, mesh->cell_nodes, mesh->cell_coords};
  const size_t tab_bytes = sizeof(double) * ((ht.hdr.total + 1) & ~1);
  k_qp_coords<<<static_cast<unsigned>(cdiv(mesh->n_cells, 256)), 256, tab_bytes, ctx->stream>>>(ht.hdr, blob.d, mv, mesh->n_cells, nq_stride, d_out);
  LFGPU_LAUNCH_CHECK(ctx);
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
  return LFGPU_OK;
}

}  // extern "C"
