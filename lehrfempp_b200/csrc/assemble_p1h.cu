// P1 (FeLagrangeO1Tria / FeLagrangeO1Quad) row kernel for everything the vertex-fan kernel (assemble_p1.cu) does not take:
// quadrilaterals and hybrid meshes (QuadO1: Jacobian per quadrature point), coefficients that vary per cell or per
// quadrature point, activity masks, cells whose corners are not bitwise their node positions (product code).
//
// Same mathematics as the generic path (uscalfe/loc_comp_ellbvp.h:266-339 with lagr_fe.h:56-263, geometry/tria_o1.cc:50-74,
// geometry/quad_o1.cc:106-158, mesh/utils/mesh_function_global.h:77-88), same output.  One thread owns one matrix row (mesh node i)
// and walks the cells around the node: first the quadrilaterals, then the triangles, so that the lanes of a warp run the same
// code although the rings of a hybrid mesh mix both (a fan-ordered walk would diverge at every ring position).  Every cell is
// taken with node i as local vertex 0 -- a ROTATION of its own vertex order -- so only row 0 of the element matrix is
// computed and every table entry (reference gradients, shape functions, weights) is a compile-time position of the kernel's
// parameter block, i.e. an operand of the FP64 instruction, not a load.  The rotation only matters for coefficients given per
// quadrature point: point k of the rotated cell is point perm[rot][k] of the cell's own numbering, which exists for every rule
// that is invariant under the rotations of the reference cell (checked on the host: the provider's default rules are; any
// other rule keeps the generic kernels).  Same integrals as the reference, summed in another order: last-bit differences.
//
// Plan per row (slot-major word arrays, so the 32 rows of a warp read full lines): per quadrilateral 4 words
// (cell | rot << 28, then the three other corners in the cell's cyclic order, each node | slot-in-row << 28), per triangle 3 words;
// one byte per row (slot of the diagonal, 0xFF = row left to the generic kernel: more cells than the plan holds, row longer than 16).
// The values of the 32 consecutive rows of a warp are accumulated in a shared-memory image of their contiguous value range
// (bank-swizzled) and leave as full 128-byte lines.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include <cub/cub.cuh>

#include "lfgpu_internal.cuh"

namespace lfgpu {
namespace {

constexpr uint32_t kNil = 0xFFFFFFFFu;
constexpr int kMaxQ = 4, kMaxT = 8;  // most quadrilaterals / triangles around one node the plan holds
constexpr int kMaxLen = 16;          // longest row (4-bit slots)

// ---- plan construction ---------------------------------------------------------------------------------------------------
// pass 0: counts per row (quadrilaterals, triangles); rows that cannot be planned get 0xFF in both
template <typename P>
__global__ void k_p1h_count(int64_t n_rows, const int32_t* __restrict__ outer, const int32_t* __restrict__ adj_ptr,
                            const uint32_t* __restrict__ adj, const uint32_t* __restrict__ cell_nodes, int* __restrict__ maxima,
                            uint8_t* __restrict__ rowinfo) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int nq = 0, nt = 0;
  if (r < n_rows) {
    const int len = outer[r + 1] - outer[r];
    const int32_t t1 = adj_ptr[r + 1];
    for (int32_t t = adj_ptr[r]; t < t1; ++t) {
      const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[adj[t] >> 4];
      if (v.w != kNil) ++nq; else ++nt;
    }
    const bool ok = len <= kMaxLen && nq <= kMaxQ && nt <= kMaxT;
    rowinfo[r] = ok ? 0 : 0xFF;
    if (!ok) nq = nt = 0;
  }
  nq = __reduce_max_sync(0xffffffffU, nq);
  nt = __reduce_max_sync(0xffffffffU, nt);
  if ((threadIdx.x & 31) == 0) {
    if (nq > 0) atomicMax(maxima, nq);
    if (nt > 0) atomicMax(maxima + 1, nt);
  }
}

// pass 1: the item words.  qw [4 * KQ][n_rows], tw [3 * KT][n_rows]; rowinfo[r] = slot of the diagonal (or 0xFF)
template <typename P>
__global__ void k_p1h_fill(int64_t n_rows, int KQ, int KT, int o_stride, int pos_row, const int32_t* __restrict__ adj_ptr,
                           const uint32_t* __restrict__ adj, const uint32_t* __restrict__ cell_nodes, const P* __restrict__ pos,
                           uint32_t* __restrict__ qw, uint32_t* __restrict__ tw, uint8_t* __restrict__ rowinfo, int* __restrict__ bad) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int nq = 0, nt = 0, diag = -1;
  if (rowinfo[r] != 0xFF) {
    const int32_t t1 = adj_ptr[r + 1];
    for (int32_t t = adj_ptr[r]; t < t1; ++t) {
      const uint32_t item = adj[t];
      const int64_t cell = item >> 4;
      const int a = static_cast<int>(item & 15U);
      const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[cell];
      const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
      const int nv = v.w != kNil ? 4 : 3;
      const P* pp = pos + (cell * o_stride + a) * pos_row;
      if (vv[a] != static_cast<uint32_t>(r)) *bad = 1;  // dof table is not the vertex table
      const int d = pp[a];
      if (diag >= 0 && d != diag) *bad = 1;
      diag = d;
      uint32_t* dst = nv == 4 ? qw + static_cast<int64_t>(4 * nq) * n_rows + r : tw + static_cast<int64_t>(3 * nt) * n_rows + r;
      dst[0] = static_cast<uint32_t>(cell) | (static_cast<uint32_t>(a) << 28);
      for (int m = 1; m < nv; ++m) {
        const int b = (a + m) % nv;
        dst[static_cast<int64_t>(m) * n_rows] = vv[b] | (static_cast<uint32_t>(pp[b]) << 28);
      }
      if (nv == 4) ++nq; else ++nt;
    }
    rowinfo[r] = static_cast<uint8_t>(diag < 0 ? 0xFE : diag);  // 0xFE: a node without cells (nothing to write)
  }
  for (int k = nq; k < KQ; ++k)
    for (int m = 0; m < 4; ++m) qw[static_cast<int64_t>(4 * k + m) * n_rows + r] = kNil;
  for (int k = nt; k < KT; ++k)
    for (int m = 0; m < 3; ++m) tw[static_cast<int64_t>(3 * k + m) * n_rows + r] = kNil;
}

__global__ void k_p1h_flag_irregular(int64_t n_rows, const uint8_t* __restrict__ rowinfo, uint8_t* __restrict__ flag) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r < n_rows) flag[r] = rowinfo[r] == 0xFF ? 1 : 0;
}

// dof table == vertex table (P1: dof == node index, dofhandler.cc:147-164)?
__global__ void k_p1h_check_nodal(int64_t n_cells, int stride, const int32_t* __restrict__ dofs, const uint8_t* __restrict__ nldof,
                                  const uint32_t* __restrict__ cell_nodes, int* __restrict__ bad) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[c];
  const int nv = v.w != kNil ? 4 : 3;
  const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
  if (nldof[c] != nv) *bad = 1;
  for (int m = 0; m < nv; ++m)
    if (dofs[c * stride + m] != static_cast<int32_t>(vv[m])) *bad = 1;
}

// ---- the kernel ------------------------------------------------------------------------------------------------------------
struct RowCoeff {
  int kind;
  int vec;  // PER_QP table with stride 4 and 32-byte aligned base: one 256-bit load fetches all points of a cell
  double c[4];
  const double* data;
  long long stride;
};

struct P1HParams {
  // triangles, cell taken as (i, p1, p2): weights, mass products ct[b][k] = w_k phi_0(k) phi_b(k)
  double wt[3];
  double ct[3][3];
  // quadrilaterals, cell taken as (i, p1, p2, p3): reference points, weights, reference gradients and shape functions
  double qx[4], qy[4], wq[4];
  double gx[4][4], gy[4][4], ph[4][4];  // [b][k]
  uint32_t perm_t, perm_q;  // byte `rot`: 2 bits per point k = index of that point in the cell's own numbering
  RowCoeff alpha, gamma;
  int has_mass;     // gamma is not the constant 0
  int transpose;    // tensor alpha given per point: swap the off-diagonal entries (row-major output)
  double beta;
};

__device__ __forceinline__ double rcp_fast(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

__device__ __forceinline__ void prefetch_l2(const void* a) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a)); }
__device__ __forceinline__ int swz(int k) { return k ^ ((k >> 4) & 15); }

// Coefficient kinds as compile-time facts where it pays (ncu / SASS of the first version: the run-time kind switches cost
// predicated-off loads, branches and ~230 register moves per row): MODE 0 reads the kinds from the parameter block (any
// combination), MODE 1 = alpha and gamma per quadrature point from 32-byte aligned tables of stride 4 and equal triangle
// weights (config C2), MODE 2 = both constant.
template <int MODE>
struct Kinds {
  static __device__ __forceinline__ int a(const P1HParams& P) { return MODE == 1 ? LFGPU_COEFF_PER_QP : (MODE == 2 ? LFGPU_COEFF_CONST : P.alpha.kind); }
  static __device__ __forceinline__ int g(const P1HParams& P) { return MODE == 1 ? LFGPU_COEFF_PER_QP : (MODE == 2 ? LFGPU_COEFF_CONST : P.gamma.kind); }
  static __device__ __forceinline__ bool avec(const P1HParams& P) { return MODE == 1 ? true : P.alpha.vec != 0; }
  static __device__ __forceinline__ bool gvec(const P1HParams& P) { return MODE == 1 ? true : P.gamma.vec != 0; }
  static __device__ __forceinline__ bool mass(const P1HParams& P) { return MODE == 1 ? true : P.has_mass != 0; }
  static __device__ __forceinline__ bool tri_eqw(const P1HParams& P) { return MODE == 1 ? true : false; }
};

// Per-cell coefficient data is GATHERED by the row-owner threads, and a gather costs one L1 wavefront per distinct line whatever
// its width: the values of all points of a cell come with ONE 256-bit load per coefficient (tables of stride 4), in the cell's
// own point numbering; the rotation is applied afterwards by register selects.
struct Raw4 {
  double v0, v1, v2, v3;
};
__device__ __forceinline__ Raw4 coeff_issue(const RowCoeff& C, int kind, bool vec, uint32_t cell, bool four) {
  Raw4 r;
  if (kind <= LFGPU_COEFF_PER_CELL) {
    r.v0 = kind == LFGPU_COEFF_CONST ? C.c[0] : __ldg(C.data + cell);
    r.v1 = r.v2 = r.v3 = r.v0;
  } else if (kind == LFGPU_COEFF_PER_QP) {
    const double* base = C.data + static_cast<long long>(cell) * C.stride;
    if (vec) {
      asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v0), "=d"(r.v1), "=d"(r.v2), "=d"(r.v3) : "l"(base));
    } else {
      r.v0 = __ldg(base); r.v1 = __ldg(base + 1); r.v2 = __ldg(base + 2);
      r.v3 = four ? __ldg(base + 3) : 0.0;
    }
  } else {
    r.v0 = r.v1 = r.v2 = r.v3 = 0.0;  // tensor kinds are read where they are used
  }
  return r;
}
__device__ __forceinline__ double pick4(int k, const Raw4& r) {
  const double lo = (k & 1) ? r.v1 : r.v0, hi = (k & 1) ? r.v3 : r.v2;
  return (k & 2) ? hi : lo;
}
template <int NQ>
__device__ __forceinline__ void coeff_points(int kind, const Raw4& r, const int (&kk)[NQ], double (&out)[NQ]) {
  if (kind != LFGPU_COEFF_PER_QP) {
#pragma unroll
    for (int k = 0; k < NQ; ++k) out[k] = r.v0;
  } else {
#pragma unroll
    for (int k = 0; k < NQ; ++k) out[k] = pick4(kk[k], r);
  }
}

__device__ __forceinline__ void coeff_tensor(const RowCoeff& C, uint32_t cell, int kk, bool tr, double& a00, double& a01, double& a10, double& a11) {
  switch (C.kind) {
    case LFGPU_COEFF_CONST: a00 = a11 = C.c[0]; a01 = a10 = 0.0; break;
    case LFGPU_COEFF_CONST_2X2: a00 = C.c[0]; a01 = C.c[1]; a10 = C.c[2]; a11 = C.c[3]; break;  // already transposed by the host
    case LFGPU_COEFF_PER_CELL: a00 = a11 = __ldg(C.data + cell); a01 = a10 = 0.0; break;
    case LFGPU_COEFF_PER_QP: a00 = a11 = __ldg(C.data + static_cast<long long>(cell) * C.stride + kk); a01 = a10 = 0.0; break;
    default: {
      const double2* p = reinterpret_cast<const double2*>(C.data + (static_cast<long long>(cell) * C.stride + kk) * 4);
      const double2 u = __ldg(p), v = __ldg(p + 1);
      a00 = u.x; a01 = tr ? v.x : u.y; a10 = tr ? u.y : v.x; a11 = v.y;
    }
  }
}

// row 0 of the element matrix of the triangle (p0, p1, p2), a = p1 - p0, b = p2 - p0: e0 (diagonal), e1, e2
template <bool TENSOR, int MODE>
__device__ __forceinline__ void tri_row(const P1HParams& P, uint32_t cell, int rot, const Raw4& ra, const Raw4& rg, double ax, double ay,
                                        double bx, double by, double& e0, double& e1, double& e2) {
  const double det = ax * by - ay * bx;
  const double adet = fabs(det);
  const double ridet = rcp_fast(adet);
  const int k0 = (P.perm_t >> (8 * rot)) & 3, k1 = (P.perm_t >> (8 * rot + 2)) & 3, k2 = (P.perm_t >> (8 * rot + 4)) & 3;
  const int kk[3] = {k0, k1, k2};
  if (!TENSOR) {
    // grad phi_b constant on the cell: sum_k w_k alpha_k |det| G_0 . G_b = (sum_k w_k alpha_k) / |det| * (N^T g_0) . (N^T g_b)
    double abar;
    if (Kinds<MODE>::tri_eqw(P)) {
      abar = P.wt[0] * ((ra.v0 + ra.v1) + ra.v2);  // equal weights: the sum over the points does not see the rotation
    } else {
      double al[3];
      coeff_points<3>(Kinds<MODE>::a(P), ra, kk, al);
      abar = P.wt[0] * al[0];
      abar = fma(P.wt[1], al[1], abar);
      abar = fma(P.wt[2], al[2], abar);
    }
    const double aa = ax * ax + ay * ay, bb = bx * bx + by * by, ab = ax * bx + ay * by;
    const double s = abar * ridet;
    e1 = s * (ab - bb);
    e2 = s * (ab - aa);
    e0 = -(e1 + e2);
  } else {
    double a00, a01, a10, a11;
    if (P.alpha.kind != LFGPU_COEFF_PER_QP_2X2) {
      coeff_tensor(P.alpha, cell, 0, false, a00, a01, a10, a11);
      const double ws = P.wt[0] + P.wt[1] + P.wt[2];
      a00 *= ws; a01 *= ws; a10 *= ws; a11 *= ws;
    } else {
      double b00, b01, b10, b11;
      coeff_tensor(P.alpha, cell, k0, P.transpose != 0, b00, b01, b10, b11);
      a00 = P.wt[0] * b00; a01 = P.wt[0] * b01; a10 = P.wt[0] * b10; a11 = P.wt[0] * b11;
      coeff_tensor(P.alpha, cell, k1, P.transpose != 0, b00, b01, b10, b11);
      a00 = fma(P.wt[1], b00, a00); a01 = fma(P.wt[1], b01, a01); a10 = fma(P.wt[1], b10, a10); a11 = fma(P.wt[1], b11, a11);
      coeff_tensor(P.alpha, cell, k2, P.transpose != 0, b00, b01, b10, b11);
      a00 = fma(P.wt[2], b00, a00); a01 = fma(P.wt[2], b01, a01); a10 = fma(P.wt[2], b10, a10); a11 = fma(P.wt[2], b11, a11);
    }
    // N = adj(J) = [by -bx; -ay ax]; u = N^T g_0 with g_0 = (-1, -1); t = A u; entry b = (N t) . g_b / |det|
    const double ux = ay - by, uy = bx - ax;
    const double tx = a00 * ux + a01 * uy, ty = a10 * ux + a11 * uy;
    const double vx = by * tx - bx * ty, vy = ax * ty - ay * tx;
    e1 = vx * ridet;
    e2 = vy * ridet;
    e0 = -(e1 + e2);
  }
  if (Kinds<MODE>::mass(P)) {
    double gv[3];
    coeff_points<3>(Kinds<MODE>::g(P), rg, kk, gv);
    const double g0 = adet * gv[0], g1 = adet * gv[1], g2 = adet * gv[2];
    e0 = fma(P.ct[0][0], g0, e0); e0 = fma(P.ct[0][1], g1, e0); e0 = fma(P.ct[0][2], g2, e0);
    e1 = fma(P.ct[1][0], g0, e1); e1 = fma(P.ct[1][1], g1, e1); e1 = fma(P.ct[1][2], g2, e1);
    e2 = fma(P.ct[2][0], g0, e2); e2 = fma(P.ct[2][1], g1, e2); e2 = fma(P.ct[2][2], g2, e2);
  }
}

// everything one cell of the row needs, loaded ahead of its use (ncu on the first version: 16 warps per SM, each waiting on
// plan word -> coordinates / coefficients -> arithmetic per cell in turn; long-scoreboard stalls 8 per issue)
struct ItemData {
  double2 p0, p1, p2, p3;
  Raw4 ra, rg;
  uint32_t cell;
  uint32_t meta;  // rot | slot_1 << 4 | slot_2 << 8 | slot_3 << 12 | on << 16
};

// row 0 of the element matrices of NB quadrilaterals (p0, p1, p2, p3) at once -- the NB cells share every table operand
// of a quadrature point while it sits in a uniform register, and their dependency chains interleave.
// With e0v = p1 - p0, e1v = p3 - p0, d = p2 - p3 - e0v:  J(xhat) = [e0v + d xhat_1, e1v + d xhat_0]  (quad_o1.cc:114-117)
template <int NB, bool TENSOR, bool CC, int MODE>
__device__ __forceinline__ void quad_rows(const P1HParams& P, const ItemData* d, double2 xi, double (&r0)[NB], double (&r1)[NB], double (&r2)[NB],
                                          double (&r3)[NB]) {
  double e0x[NB], e0y[NB], e1x[NB], e1y[NB], dx[NB], dy[NB], al[NB][4], gv[NB][4];
  int kq[NB][4];
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const double2 p0 = CC ? d[q].p0 : xi;
    e0x[q] = d[q].p1.x - p0.x; e0y[q] = d[q].p1.y - p0.y; e1x[q] = d[q].p3.x - p0.x; e1y[q] = d[q].p3.y - p0.y;
    dx[q] = (d[q].p2.x - d[q].p3.x) - e0x[q]; dy[q] = (d[q].p2.y - d[q].p3.y) - e0y[q];
    const uint32_t pq = P.perm_q >> (8 * (d[q].meta & 15U));
    kq[q][0] = static_cast<int>(pq & 3); kq[q][1] = static_cast<int>((pq >> 2) & 3); kq[q][2] = static_cast<int>((pq >> 4) & 3);
    kq[q][3] = static_cast<int>((pq >> 6) & 3);
    coeff_points<4>(Kinds<MODE>::a(P), d[q].ra, kq[q], al[q]);
    coeff_points<4>(Kinds<MODE>::g(P), d[q].rg, kq[q], gv[q]);
    r0[q] = r1[q] = r2[q] = r3[q] = 0.0;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const double c0x = fma(dx[q], P.qy[k], e0x[q]), c0y = fma(dy[q], P.qy[k], e0y[q]);
      const double c1x = fma(dx[q], P.qx[k], e1x[q]), c1y = fma(dy[q], P.qx[k], e1y[q]);
      const double det = c0x * c1y - c0y * c1x;
      const double adet = fabs(det);
      const double ridet = rcp_fast(adet);
      // u = N^T ghat_0, N = adj(J) = [c1y -c1x; -c0y c0x]
      const double ux = c1y * P.gx[0][k] - c0y * P.gy[0][k];
      const double uy = c0x * P.gy[0][k] - c1x * P.gx[0][k];
      double tx, ty;
      if (!TENSOR) {
        const double s = P.wq[k] * al[q][k] * ridet;
        tx = s * ux;
        ty = s * uy;
      } else {
        double a00, a01, a10, a11;
        coeff_tensor(P.alpha, d[q].cell, kq[q][k], P.transpose != 0, a00, a01, a10, a11);
        const double s = P.wq[k] * ridet;
        tx = s * (a00 * ux + a01 * uy);
        ty = s * (a10 * ux + a11 * uy);
      }
      const double vx = c1y * tx - c1x * ty, vy = c0x * ty - c0y * tx;
      r0[q] = fma(vx, P.gx[0][k], r0[q]); r0[q] = fma(vy, P.gy[0][k], r0[q]);
      r1[q] = fma(vx, P.gx[1][k], r1[q]); r1[q] = fma(vy, P.gy[1][k], r1[q]);
      r2[q] = fma(vx, P.gx[2][k], r2[q]); r2[q] = fma(vy, P.gy[2][k], r2[q]);
      r3[q] = fma(vx, P.gx[3][k], r3[q]); r3[q] = fma(vy, P.gy[3][k], r3[q]);
      if (Kinds<MODE>::mass(P)) {
        const double mm = P.wq[k] * adet * gv[q][k] * P.ph[0][k];
        r0[q] = fma(mm, P.ph[0][k], r0[q]);
        r1[q] = fma(mm, P.ph[1][k], r1[q]);
        r2[q] = fma(mm, P.ph[2][k], r2[q]);
        r3[q] = fma(mm, P.ph[3][k], r3[q]);
      }
    }
  }
}

template <bool QUAD, bool CC, int MODE>
__device__ __forceinline__ void item_issue(ItemData& d, const P1HParams& P, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int32_t r,
                                           const double2* __restrict__ nc, const double2* __restrict__ cc, const uint8_t* __restrict__ active) {
  const bool valid = w0 != kNil;
  const uint32_t cell = valid ? (w0 & 0x0fffffffU) : 0U;
  const int rot = valid ? static_cast<int>(w0 >> 28) : 0;
  uint32_t on = valid ? 1U : 0U;
  if (active != nullptr) on &= __ldg(active + cell);
  d.cell = cell;
  d.meta = static_cast<uint32_t>(rot) | ((w1 >> 28) << 4) | ((w2 >> 28) << 8) | ((QUAD ? (w3 >> 28) : 0U) << 12) | (on << 16);
  if (CC) {
    const double2* c4 = cc + 4 * static_cast<size_t>(cell);
    d.p0 = __ldg(c4 + rot);
    if (QUAD) {
      d.p1 = __ldg(c4 + ((rot + 1) & 3)); d.p2 = __ldg(c4 + ((rot + 2) & 3)); d.p3 = __ldg(c4 + ((rot + 3) & 3));
    } else {
      d.p1 = __ldg(c4 + (rot + 1) % 3); d.p2 = __ldg(c4 + (rot + 2) % 3);
    }
  } else {
    d.p1 = __ldg(nc + (valid ? (w1 & 0x0fffffffU) : static_cast<uint32_t>(r)));
    d.p2 = __ldg(nc + (valid ? (w2 & 0x0fffffffU) : static_cast<uint32_t>(r)));
    if (QUAD) d.p3 = __ldg(nc + (valid ? (w3 & 0x0fffffffU) : static_cast<uint32_t>(r)));
  }
  d.ra = coeff_issue(P.alpha, Kinds<MODE>::a(P), Kinds<MODE>::avec(P), cell, QUAD);
  d.rg = coeff_issue(P.gamma, Kinds<MODE>::g(P), Kinds<MODE>::gvec(P), cell, QUAD);
}

__device__ __forceinline__ void stage_add(double* __restrict__ stage, int off, uint32_t meta, int shift, double v) {
  stage[swz(off + static_cast<int>((meta >> shift) & 15U))] += v;
}

// NB cells of one kind at once (quadrilaterals in pairs, triangles singly)
template <bool QUAD, int NB, bool TENSOR, bool CC, int MODE>
__device__ __forceinline__ void group_consume(const ItemData* d, const P1HParams& P, double2 xi, double* __restrict__ stage, int off, double& diag) {
  bool any = false;
#pragma unroll
  for (int q = 0; q < NB; ++q) any = any || (d[q].meta >> 16) != 0;
  if (!__any_sync(0xffffffffU, any)) return;  // slots no row of the warp uses
  if (QUAD) {
    double r0[NB], r1[NB], r2[NB], r3[NB];
    quad_rows<NB, TENSOR, CC, MODE>(P, d, xi, r0, r1, r2, r3);
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      if ((d[q].meta >> 16) != 0) {
        diag += r0[q];
        stage_add(stage, off, d[q].meta, 4, r1[q]);
        stage_add(stage, off, d[q].meta, 8, r2[q]);
        stage_add(stage, off, d[q].meta, 12, r3[q]);
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const double2 p0 = CC ? d[q].p0 : xi;
      double e0, e1, e2;
      tri_row<TENSOR, MODE>(P, d[q].cell, static_cast<int>(d[q].meta & 15U), d[q].ra, d[q].rg, d[q].p1.x - p0.x, d[q].p1.y - p0.y,
                            d[q].p2.x - p0.x, d[q].p2.y - p0.y, e0, e1, e2);
      if ((d[q].meta >> 16) != 0) {
        diag += e0;
        stage_add(stage, off, d[q].meta, 4, e1);
        stage_add(stage, off, d[q].meta, 8, e2);
      }
    }
  }
}

// CC: cell corners come from the mesh's cell_coords array (cells whose geometry is not bitwise the node positions).
// The cells of a row are processed in groups -- KQ / 2 pairs of quadrilaterals, then KT triangles -- and the loads of group
// g + DEPTH are issued before group g is computed.
template <int KQ, int KT, bool TENSOR, bool CC, int DEPTH, int MODE, int MINB = 0>
__global__ void __launch_bounds__(128, MINB > 0 ? MINB : (KQ == 4 ? 3 : 4)) k_assemble_p1_rows(int n_rows, int n_total_rows, const uint32_t* __restrict__ qw,
                                                             const uint32_t* __restrict__ tw, const uint8_t* __restrict__ rowinfo,
                                                             const double* __restrict__ node_coords, const double* __restrict__ cell_coords,
                                                             const int32_t* __restrict__ outer, const uint8_t* __restrict__ active,
                                                             const int32_t* __restrict__ row_list, int row0, int pf_dist,
                                                             const __grid_constant__ P1HParams P, double* __restrict__ values) {
  extern __shared__ double stage_all[];
  static_assert(KQ % 2 == 0, "quadrilaterals are processed in pairs");
  constexpr int NI = KQ + KT, NG = KQ / 2 + KT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = t < n_rows;
  const int32_t r = in_range ? (row_list != nullptr ? row_list[t] : t + row0) : 0;
  if (pf_dist > 0 && row_list == nullptr && warp == 0) {
    // pull the plan, row-pointer and coordinate lines of the CTA one wave ahead into L2 (see assemble_p1.cu)
    const int tp = blockIdx.x * blockDim.x + pf_dist;
    if (tp + 128 <= n_rows) {
      const size_t rp = static_cast<size_t>(tp) + row0;
      constexpr int nq_lines = 16 * KQ, nt_lines = 12 * KT;
      for (int L = lane; L < nq_lines + nt_lines + 21; L += 32) {
        const char* a;
        if (L < nq_lines) {
          a = reinterpret_cast<const char*>(qw + static_cast<size_t>(L >> 2) * n_total_rows + rp) + (L & 3) * 128;
        } else if (L < nq_lines + nt_lines) {
          const int M = L - nq_lines;
          a = reinterpret_cast<const char*>(tw + static_cast<size_t>(M >> 2) * n_total_rows + rp) + (M & 3) * 128;
        } else if (L == nq_lines + nt_lines) {
          a = reinterpret_cast<const char*>(rowinfo + rp);
        } else if (L < nq_lines + nt_lines + 5) {
          a = reinterpret_cast<const char*>(outer + rp) + (L - nq_lines - nt_lines - 1) * 128;
        } else {
          a = reinterpret_cast<const char*>(node_coords + 2 * rp) + (L - nq_lines - nt_lines - 5) * 128;
        }
        prefetch_l2(a);
      }
    }
  }
  int32_t v0 = 0, v1 = 0;
  int info = 0xFE;
  if (in_range) {
    v0 = __ldg(outer + r);
    v1 = __ldg(outer + r + 1);
    info = __ldg(rowinfo + r);
  }
  const bool regular = in_range && info < 0xFE;
  // the plan words of the row, all at once
  uint32_t w[4 * KQ + 3 * KT + 1];
#pragma unroll
  for (int k = 0; k < 4 * KQ; ++k) w[k] = regular ? __ldg(qw + static_cast<size_t>(k) * n_total_rows + r) : kNil;
#pragma unroll
  for (int k = 0; k < 3 * KT; ++k) w[4 * KQ + k] = regular ? __ldg(tw + static_cast<size_t>(k) * n_total_rows + r) : kNil;
  // staging: the warp's rows are consecutive and none of them belongs to the generic kernel -> image of one value range
  const int32_t wbase = __shfl_sync(0xffffffffU, v0, 0);
  const int32_t r_first = __shfl_sync(0xffffffffU, r, 0);
  const bool consecutive = row_list == nullptr || !__any_sync(0xffffffffU, in_range && r != r_first + lane);
  const bool staged = consecutive && !__any_sync(0xffffffffU, in_range && info == 0xFF);
  double* stage = stage_all + warp * (32 * kMaxLen);
  const int off = staged ? v0 - wbase : lane * kMaxLen;
  const double2* nc = reinterpret_cast<const double2*>(node_coords);
  const double2* cc = reinterpret_cast<const double2*>(cell_coords);
  double2 xi = make_double2(0.0, 0.0);
  if (!CC) xi = __ldg(nc + r);
  if (regular) {
    const int len = v1 - v0;
    for (int s = 0; s < len; ++s) stage[swz(off + s)] = 0.0;
  }
  double diag = 0.0;
  ItemData D[NI];
#pragma unroll
  for (int s = 0; s < NG + DEPTH; ++s) {
    if (s < NG) {
      if (s < KQ / 2) {
        item_issue<true, CC, MODE>(D[2 * s], P, w[8 * s], w[8 * s + 1], w[8 * s + 2], w[8 * s + 3], r, nc, cc, active);
        item_issue<true, CC, MODE>(D[2 * s + 1], P, w[8 * s + 4], w[8 * s + 5], w[8 * s + 6], w[8 * s + 7], r, nc, cc, active);
      } else {
        const int i = s - KQ / 2;
        item_issue<false, CC, MODE>(D[KQ + i], P, w[4 * KQ + 3 * i], w[4 * KQ + 3 * i + 1], w[4 * KQ + 3 * i + 2], 0U, r, nc, cc, active);
      }
    }
    if (s >= DEPTH) {
      const int g = s - DEPTH;
      if (g < KQ / 2)
        group_consume<true, 2, TENSOR, CC, MODE>(&D[2 * g], P, xi, stage, off, diag);
      else
        group_consume<false, 1, TENSOR, CC, MODE>(&D[KQ + g - KQ / 2], P, xi, stage, off, diag);
    }
  }
  if (regular) stage[swz(off + info)] = diag;
  __syncwarp();
  if (staged) {
    const unsigned ballot = __ballot_sync(0xffffffffU, in_range);
    if (ballot == 0) return;
    const int total = __shfl_sync(0xffffffffU, v1, 31 - __clz(ballot)) - wbase;
    double* out = values + wbase;
    if (P.beta == 0.0) {
#pragma unroll
      for (int k = 0; k < kMaxLen; ++k) {
        const int idx = k * 32 + lane;
        if (idx < total) out[idx] = stage[swz(idx)];
      }
    } else {
#pragma unroll
      for (int k = 0; k < kMaxLen; ++k) {
        const int idx = k * 32 + lane;
        if (idx < total) out[idx] = fma(P.beta, out[idx], stage[swz(idx)]);
      }
    }
  } else if (regular) {
    const int len = v1 - v0;
    for (int k = 0; k < len; ++k) {
      const double v = stage[swz(off + k)];
      values[v0 + k] = P.beta == 0.0 ? v : fma(P.beta, values[v0 + k], v);
    }
  }
}

// Hybrid rows with the work of a row split over TWO threads in different warps of the CTA (mixed meshes: KQ > 0 and KT > 0):
// warps 0-1 take the quadrilaterals of 64 rows, warps 2-3 the triangles of the same rows.  ncu on the one-thread-per-row kernel
// at config C2: 16 warps per SM at 128 registers, each walking its row's six cells in turn -- issue slots 51 % busy, stall
// samples on the first use of every group of loads.  Two threads per row halve the serial chain of a row, every warp still runs one
// code path (no divergence), and each role keeps fewer values alive.  The quadrilateral warps accumulate into the shared image
// first; the triangle warps compute meanwhile and add after a CTA barrier; all four warps copy the two images out.
template <int KQ, int KT, bool TENSOR, bool CC, int MODE>
__global__ void __launch_bounds__(128, KQ == 4 ? 3 : 5) k_assemble_p1_rows_split(int n_rows, int n_total_rows, const uint32_t* __restrict__ qw,
                                                                   const uint32_t* __restrict__ tw, const uint8_t* __restrict__ rowinfo,
                                                                   const double* __restrict__ node_coords,
                                                                   const double* __restrict__ cell_coords, const int32_t* __restrict__ outer,
                                                                   const uint8_t* __restrict__ active, const int32_t* __restrict__ row_list,
                                                                   int row0, int pf_dist, const __grid_constant__ P1HParams P,
                                                                   double* __restrict__ values) {
  extern __shared__ double stage_all[];
  static_assert(KQ % 2 == 0 && KQ > 0 && KT > 0, "mixed rows only");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = warp & 1, role = warp >> 1;  // role 0: quadrilaterals, role 1: triangles
  const int t = blockIdx.x * 64 + half * 32 + lane;
  const bool in_range = t < n_rows;
  const int32_t r = in_range ? (row_list != nullptr ? row_list[t] : t + row0) : 0;
  if (pf_dist > 0 && row_list == nullptr && warp == 0) {
    const int tp = blockIdx.x * 64 + pf_dist;
    if (tp + 64 <= n_rows) {
      const size_t rp = static_cast<size_t>(tp) + row0;
      constexpr int nq_lines = 8 * KQ, nt_lines = 6 * KT;  // 2 lines of 128 B per plan array and CTA
      for (int L = lane; L < nq_lines + nt_lines + 11; L += 32) {
        const char* a;
        if (L < nq_lines) {
          a = reinterpret_cast<const char*>(qw + static_cast<size_t>(L >> 1) * n_total_rows + rp) + (L & 1) * 128;
        } else if (L < nq_lines + nt_lines) {
          const int M = L - nq_lines;
          a = reinterpret_cast<const char*>(tw + static_cast<size_t>(M >> 1) * n_total_rows + rp) + (M & 1) * 128;
        } else if (L == nq_lines + nt_lines) {
          a = reinterpret_cast<const char*>(rowinfo + rp);
        } else if (L < nq_lines + nt_lines + 3) {
          a = reinterpret_cast<const char*>(outer + rp) + (L - nq_lines - nt_lines - 1) * 128;
        } else {
          a = reinterpret_cast<const char*>(node_coords + 2 * rp) + (L - nq_lines - nt_lines - 3) * 128;
        }
        prefetch_l2(a);
      }
    }
  }
  int32_t v0 = 0, v1 = 0;
  int info = 0xFE;
  if (in_range) {
    v0 = __ldg(outer + r);
    v1 = __ldg(outer + r + 1);
    info = __ldg(rowinfo + r);
  }
  const bool regular = in_range && info < 0xFE;
  const int32_t wbase = __shfl_sync(0xffffffffU, v0, 0);
  const int32_t r_first = __shfl_sync(0xffffffffU, r, 0);
  const bool consecutive = row_list == nullptr || !__any_sync(0xffffffffU, in_range && r != r_first + lane);
  const bool staged = consecutive && !__any_sync(0xffffffffU, in_range && info == 0xFF);
  double* stage = stage_all + half * (32 * kMaxLen);
  const int off = staged ? v0 - wbase : lane * kMaxLen;
  const double2* nc = reinterpret_cast<const double2*>(node_coords);
  const double2* cc = reinterpret_cast<const double2*>(cell_coords);
  double2 xi = make_double2(0.0, 0.0);
  if (!CC) xi = __ldg(nc + r);
  double diag = 0.0;
  if (role == 0) {
    uint32_t w[4 * KQ];
#pragma unroll
    for (int k = 0; k < 4 * KQ; ++k) w[k] = regular ? __ldg(qw + static_cast<size_t>(k) * n_total_rows + r) : kNil;
    if (regular) {
      const int len = v1 - v0;
      for (int s = 0; s < len; ++s) stage[swz(off + s)] = 0.0;
    }
    ItemData D[KQ];
#pragma unroll
    for (int g = 0; g < KQ / 2 + 1; ++g) {  // pair g + 1 is loaded before pair g is computed
      if (g < KQ / 2) {
        item_issue<true, CC, MODE>(D[2 * g], P, w[8 * g], w[8 * g + 1], w[8 * g + 2], w[8 * g + 3], r, nc, cc, active);
        item_issue<true, CC, MODE>(D[2 * g + 1], P, w[8 * g + 4], w[8 * g + 5], w[8 * g + 6], w[8 * g + 7], r, nc, cc, active);
      }
      if (g >= 1) group_consume<true, 2, TENSOR, CC, MODE>(&D[2 * (g - 1)], P, xi, stage, off, diag);
    }
    if (regular) stage[swz(off + info)] = diag;
    __syncthreads();
  } else {
    uint32_t w[3 * KT];
#pragma unroll
    for (int k = 0; k < 3 * KT; ++k) w[k] = regular ? __ldg(tw + static_cast<size_t>(k) * n_total_rows + r) : kNil;
    ItemData D[KT];
    double e1[KT], e2[KT];
#pragma unroll
    for (int s = 0; s < KT + 2; ++s) {  // triangle s + 2 is loaded before triangle s is computed
      if (s < KT) item_issue<false, CC, MODE>(D[s], P, w[3 * s], w[3 * s + 1], w[3 * s + 2], 0U, r, nc, cc, active);
      if (s >= 2) {
        const int c = s - 2;
        const double2 p0 = CC ? D[c].p0 : xi;
        double e0;
        tri_row<TENSOR, MODE>(P, D[c].cell, static_cast<int>(D[c].meta & 15U), D[c].ra, D[c].rg, D[c].p1.x - p0.x, D[c].p1.y - p0.y,
                              D[c].p2.x - p0.x, D[c].p2.y - p0.y, e0, e1[c], e2[c]);
        if ((D[c].meta >> 16) != 0) diag += e0;
      }
    }
    __syncthreads();  // the quadrilateral warps have zeroed and filled the image
#pragma unroll
    for (int c = 0; c < KT; ++c) {
      if ((D[c].meta >> 16) != 0) {
        stage_add(stage, off, D[c].meta, 4, e1[c]);
        stage_add(stage, off, D[c].meta, 8, e2[c]);
      }
    }
    if (regular) stage[swz(off + info)] += diag;
  }
  __syncthreads();
  if (staged) {
    const unsigned ballot = __ballot_sync(0xffffffffU, in_range);
    if (ballot == 0) return;
    const int total = __shfl_sync(0xffffffffU, v1, 31 - __clz(ballot)) - wbase;
    double* out = values + wbase;
    // the two warps of a half share the copy of its image
    for (int idx = role * 32 + lane; idx < total; idx += 64) {
      const double v = stage[swz(idx)];
      out[idx] = P.beta == 0.0 ? v : fma(P.beta, out[idx], v);
    }
  } else if (regular && role == 0) {
    const int len = v1 - v0;
    for (int k = 0; k < len; ++k) {
      const double v = stage[swz(off + k)];
      values[v0 + k] = P.beta == 0.0 ? v : fma(P.beta, values[v0 + k], v);
    }
  }
}

// is the rule invariant under the rotations of its reference cell?  perm[rot] (2 bits per point) = index, in the cell's own
// numbering, of point k of the cell taken with its local vertex `rot` as vertex 0
bool rule_rotations(const FeTable& t, int nv, uint32_t* perm_out) {
  const int nq = t.nq;
  if (nq != nv) return false;  // the kernel is compiled for the default rules: 3 points on triangles, 2 x 2 on quadrilaterals
  uint32_t perm = 0;
  for (int rot = 0; rot < nv; ++rot) {
    for (int k = 0; k < nq; ++k) {
      // vertex weights of point k in the rotated frame ...
      double lam[4];
      if (nv == 3) {
        lam[0] = 1.0 - t.qx[k] - t.qy[k]; lam[1] = t.qx[k]; lam[2] = t.qy[k];
      } else {
        lam[0] = (1.0 - t.qx[k]) * (1.0 - t.qy[k]); lam[1] = t.qx[k] * (1.0 - t.qy[k]); lam[2] = t.qx[k] * t.qy[k]; lam[3] = (1.0 - t.qx[k]) * t.qy[k];
      }
      // ... belong to the cell's own vertices (rot + m) % nv
      double own[4];
      for (int m = 0; m < nv; ++m) own[(rot + m) % nv] = lam[m];
      const double x = nv == 3 ? own[1] : own[1] + own[2], y = nv == 3 ? own[2] : own[2] + own[3];
      int found = -1;
      for (int j = 0; j < nq; ++j)
        if (std::fabs(t.qx[j] - x) < 1e-13 && std::fabs(t.qy[j] - y) < 1e-13 && std::fabs(t.w[j] - t.w[k]) < 1e-15) found = j;
      if (found < 0) return false;
      perm |= static_cast<uint32_t>(found) << (8 * rot + 2 * k);
    }
  }
  *perm_out = perm;
  return true;
}

}  // namespace

// ---- host side ---------------------------------------------------------------------------------------------------------------
int p1h_prepare(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, lfgpu_pattern* p) {
  if (p->p1h_state != 0) return LFGPU_OK;
  p->p1h_state = -1;
  if (p->i_dofs != p->o_dofs || p->n_outer != mesh->n_nodes || p->n_outer >= (1LL << 28) - 1 || p->n_cells >= (1LL << 28) - 1 ||
      p->pos == nullptr)
    return LFGPU_OK;
  cudaStream_t st = ctx->stream;
  int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 256);
  LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_flags, 0, 16, st));
  const unsigned gc = static_cast<unsigned>(cdiv(p->n_cells, 256)), gr = static_cast<unsigned>(cdiv(p->n_outer, 256));
  k_p1h_check_nodal<<<gc, 256, 0, st>>>(p->n_cells, p->o_stride, p->o_dofs, p->o_nldof, mesh->cell_nodes, d_flags + 2);
  LFGPU_LAUNCH_CHECK(ctx);
  uint8_t* rowinfo = nullptr;
  uint32_t *qw = nullptr, *tw = nullptr;
  uint8_t* flag = nullptr;
  int32_t* iota = nullptr;
  int64_t* d_num = nullptr;
  void* tmp = nullptr;
  auto cleanup = [&]() { cudaFree(flag); cudaFree(iota); cudaFree(d_num); cudaFree(tmp); };
  auto fail = [&]() { cleanup(); cudaFree(rowinfo); cudaFree(qw); cudaFree(tw); };
#define P1H_CHECK(expr)                                                             \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      set_last_error(ctx, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
      fail();                                                                       \
      return LFGPU_ERR_CUDA;                                                        \
    }                                                                               \
  } while (0)
  P1H_CHECK(cudaMalloc(&rowinfo, p->n_outer + 4));
  if (p->pos_bytes == 1)
    k_p1h_count<uint8_t><<<gr, 256, 0, st>>>(p->n_outer, p->outer, p->adj_ptr, p->adj, mesh->cell_nodes, d_flags, rowinfo);
  else
    k_p1h_count<uint16_t><<<gr, 256, 0, st>>>(p->n_outer, p->outer, p->adj_ptr, p->adj, mesh->cell_nodes, d_flags, rowinfo);
  ctx->launches++;
  int h[3] = {0, 0, 0};
  P1H_CHECK(cudaMemcpyAsync(h, d_flags, sizeof(h), cudaMemcpyDeviceToHost, st));
  P1H_CHECK(cudaStreamSynchronize(st));
  if (h[2] != 0 || (h[0] == 0 && h[1] == 0)) {  // not the nodal P1 table, or no row can be planned
    fail();
    return LFGPU_OK;
  }
  // instantiated shapes: (KQ, KT) in {(0, 8), (4, 0), (2, 4), (4, 8)}
  int KQ, KT;
  if (h[0] == 0) { KQ = 0; KT = 8; }
  else if (h[1] == 0) { KQ = 4; KT = 0; }
  else if (h[0] <= 2 && h[1] <= 4) { KQ = 2; KT = 4; }
  else { KQ = 4; KT = 8; }
  if (KQ > 0) P1H_CHECK(cudaMalloc(&qw, sizeof(uint32_t) * 4 * KQ * static_cast<size_t>(p->n_outer)));
  if (KT > 0) P1H_CHECK(cudaMalloc(&tw, sizeof(uint32_t) * 3 * KT * static_cast<size_t>(p->n_outer)));
  P1H_CHECK(cudaMemsetAsync(d_flags, 0, 16, st));
  if (p->pos_bytes == 1)
    k_p1h_fill<uint8_t><<<gr, 256, 0, st>>>(p->n_outer, KQ, KT, p->o_stride, p->pos_row, p->adj_ptr, p->adj, mesh->cell_nodes,
                                           static_cast<const uint8_t*>(p->pos), qw, tw, rowinfo, d_flags);
  else
    k_p1h_fill<uint16_t><<<gr, 256, 0, st>>>(p->n_outer, KQ, KT, p->o_stride, p->pos_row, p->adj_ptr, p->adj, mesh->cell_nodes,
                                            static_cast<const uint16_t*>(p->pos), qw, tw, rowinfo, d_flags);
  ctx->launches++;
  // rows left to the generic kernel
  P1H_CHECK(cudaMalloc(&flag, p->n_outer));
  k_p1h_flag_irregular<<<gr, 256, 0, st>>>(p->n_outer, rowinfo, flag);
  ctx->launches++;
  P1H_CHECK(cudaMalloc(&iota, sizeof(int32_t) * p->n_outer));
  P1H_CHECK(cudaMalloc(&d_num, sizeof(int64_t)));
  cub::CountingInputIterator<int32_t> count_it(0);
  size_t tb = 0;
  if (and_row_keep(ctx, p->n_outer, flag, p->row_keep) != LFGPU_OK) return LFGPU_ERR_CUDA;  // rows nobody asks for need no generic kernel
  cub::DeviceSelect::Flagged(nullptr, tb, count_it, flag, iota, d_num, p->n_outer, st);
  P1H_CHECK(cudaMalloc(&tmp, tb));
  P1H_CHECK(cub::DeviceSelect::Flagged(tmp, tb, count_it, flag, iota, d_num, p->n_outer, st));
  int64_t n_irr = 0;
  int bad = 0;
  P1H_CHECK(cudaMemcpyAsync(&n_irr, d_num, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  P1H_CHECK(cudaMemcpyAsync(&bad, d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
  P1H_CHECK(cudaStreamSynchronize(st));
  if (bad != 0 || n_irr * 2 > p->n_outer) {  // mostly irregular: the generic kernels are the better choice
    fail();
    return LFGPU_OK;
  }
  if (n_irr > 0) {
    P1H_CHECK(cudaMalloc(&p->p1h_irregular, sizeof(int32_t) * n_irr));
    P1H_CHECK(cudaMemcpyAsync(p->p1h_irregular, iota, sizeof(int32_t) * n_irr, cudaMemcpyDeviceToDevice, st));
    p->p1h_irregular_host.resize(n_irr);
    P1H_CHECK(cudaMemcpyAsync(p->p1h_irregular_host.data(), iota, sizeof(int32_t) * n_irr, cudaMemcpyDeviceToHost, st));
    P1H_CHECK(cudaStreamSynchronize(st));
  }
#undef P1H_CHECK
  cleanup();
  p->n_p1h_irregular = n_irr;
  p->p1h_qw = qw;
  p->p1h_tw = tw;
  p->p1h_rowinfo = rowinfo;
  p->p1h_kq = KQ;
  p->p1h_kt = KT;
  p->p1h_state = 1;
  return LFGPU_OK;
}

// does the kernel take these tables?  (rules invariant under the rotations of the reference cells, 3 / 4 points)
bool p1h_rules_ok(const FeTable* tt, const FeTable* tq) {
  uint32_t perm;
  if (tt != nullptr && (tt->nsf != 3 || !rule_rotations(*tt, 3, &perm))) return false;
  if (tq != nullptr && (tq->nsf != 4 || !rule_rotations(*tq, 4, &perm))) return false;
  return true;
}

int p1h_launch(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* p, const FeTable* tt, const FeTable* tq, const lfgpu_coeff* alpha,
               const lfgpu_coeff* gamma, const uint8_t* active, double beta, const int32_t* row_list, int64_t n_rows, int64_t row0,
               double* d_values) {
  P1HParams P{};
  if (tt != nullptr) {
    rule_rotations(*tt, 3, &P.perm_t);
    for (int k = 0; k < 3; ++k) {
      P.wt[k] = tt->w[k];
      for (int b = 0; b < 3; ++b) P.ct[b][k] = tt->w[k] * tt->phi[0 * 3 + k] * tt->phi[b * 3 + k];
    }
  }
  if (tq != nullptr) {
    rule_rotations(*tq, 4, &P.perm_q);
    for (int k = 0; k < 4; ++k) {
      P.qx[k] = tq->qx[k];
      P.qy[k] = tq->qy[k];
      P.wq[k] = tq->w[k];
      for (int b = 0; b < 4; ++b) {
        P.gx[b][k] = tq->gx[b * 4 + k];
        P.gy[b][k] = tq->gy[b * 4 + k];
        P.ph[b][k] = tq->phi[b * 4 + k];
      }
    }
  }
  const bool tr = p->major == LFGPU_ROW_MAJOR;
  P.alpha.kind = alpha->kind;
  for (int i = 0; i < 4; ++i) P.alpha.c[i] = alpha->c[i];
  if (alpha->kind == LFGPU_COEFF_CONST_2X2 && tr) std::swap(P.alpha.c[1], P.alpha.c[2]);
  P.alpha.data = alpha->data;
  P.alpha.stride = alpha->stride;
  P.alpha.vec = (alpha->kind == LFGPU_COEFF_PER_QP && alpha->stride == 4 && (reinterpret_cast<uintptr_t>(alpha->data) & 31) == 0) ? 1 : 0;
  P.gamma.vec = (gamma->kind == LFGPU_COEFF_PER_QP && gamma->stride == 4 && (reinterpret_cast<uintptr_t>(gamma->data) & 31) == 0) ? 1 : 0;
  P.gamma.kind = gamma->kind;
  for (int i = 0; i < 4; ++i) P.gamma.c[i] = gamma->c[i];
  P.gamma.data = gamma->data;
  P.gamma.stride = gamma->stride;
  P.has_mass = !(gamma->kind == LFGPU_COEFF_CONST && gamma->c[0] == 0.0);
  P.transpose = tr ? 1 : 0;
  P.beta = beta;
  const bool tensor = alpha->kind == LFGPU_COEFF_CONST_2X2 || alpha->kind == LFGPU_COEFF_PER_QP_2X2;
  const bool cc = mesh->cell_coords != nullptr;
  const int64_t rows = (row_list != nullptr || row0 >= 0) ? n_rows : p->n_outer;
  if (rows <= 0) return LFGPU_OK;
  const int64_t first_row = (row_list == nullptr && row0 >= 0) ? row0 : 0;
  const int threads = 128;
  const unsigned grid = static_cast<unsigned>(cdiv(rows, threads));
  const size_t smem = sizeof(double) * (threads / 32) * 32 * kMaxLen;
  static const int pfd_env = [] { const char* e = std::getenv("LFGPU_P1H_PFD"); return e != nullptr ? std::atoi(e) : 100; }();
  const int ipf = cc ? 0 : static_cast<int>((static_cast<int64_t>(ctx->sm_count) * 4 * threads * pfd_env / 100) & ~static_cast<int64_t>(127));
  const int irows = static_cast<int>(rows), itotal = static_cast<int>(p->n_outer), ifirst = static_cast<int>(first_row);
  // groups of cells loaded this many groups ahead of their use (LFGPU_P1H_DEPTH = 1 or 2)
  static const int depth_env = [] { const char* e = std::getenv("LFGPU_P1H_DEPTH"); return e != nullptr ? std::atoi(e) : 1; }();
  static const int mode_env = [] { const char* e = std::getenv("LFGPU_P1H_MODE"); return e != nullptr ? std::atoi(e) : -1; }();
  // compile-time coefficient kinds where they apply (Kinds<> above); LFGPU_P1H_MODE=0 keeps the run-time switches
  int mode = 0;
  if (!tensor && !cc) {
    const bool eqw = tt == nullptr || (tt->w[0] == tt->w[1] && tt->w[1] == tt->w[2]);
    if (P.alpha.vec && P.gamma.vec && eqw) mode = 1;
    if (alpha->kind == LFGPU_COEFF_CONST && gamma->kind == LFGPU_COEFF_CONST) mode = 2;
  }
  if (mode_env == 0) mode = 0;
#define P1H_KERN(KQ, KT, D)                                                                                                          \
  (tensor ? (cc ? k_assemble_p1_rows<KQ, KT, true, true, D, 0> : k_assemble_p1_rows<KQ, KT, true, false, D, 0>)                       \
          : (cc ? k_assemble_p1_rows<KQ, KT, false, true, D, 0>                                                                      \
                : (mode == 1 ? k_assemble_p1_rows<KQ, KT, false, false, D, 1>                                                        \
                             : (mode == 2 ? k_assemble_p1_rows<KQ, KT, false, false, D, 2> : k_assemble_p1_rows<KQ, KT, false, false, D, 0>))))
#define P1H_LAUNCH(KQ, KT)                                                                                                         \
  do {                                                                                                                             \
    auto kern = depth_env == 2 ? P1H_KERN(KQ, KT, 2) : P1H_KERN(KQ, KT, 1);                                                         \
    kern<<<grid, threads, smem, ctx->stream>>>(irows, itotal, p->p1h_qw, p->p1h_tw, p->p1h_rowinfo, mesh->node_coords, mesh->cell_coords, \
                                               p->outer, active, row_list, ifirst, ipf, P, d_values);                                  \
  } while (0)
  // mixed rows, opt-in (LFGPU_P1H_SPLIT=1): two threads per row in warps of different roles.  Measured on config C2: 0.277 ms
  // against 0.273 ms with one thread per row -- the serial chain of a row is not what limits the kernel -- so it stays off.
  static const bool split_env = [] { const char* e = std::getenv("LFGPU_P1H_SPLIT"); return e != nullptr && e[0] == '1'; }();
#define P1H_SPLIT_KERN(KQ, KT)                                                                                                       \
  (tensor ? (cc ? k_assemble_p1_rows_split<KQ, KT, true, true, 0> : k_assemble_p1_rows_split<KQ, KT, true, false, 0>)                  \
          : (cc ? k_assemble_p1_rows_split<KQ, KT, false, true, 0>                                                                    \
                : (mode == 1 ? k_assemble_p1_rows_split<KQ, KT, false, false, 1>                                                      \
                             : (mode == 2 ? k_assemble_p1_rows_split<KQ, KT, false, false, 2> : k_assemble_p1_rows_split<KQ, KT, false, false, 0>))))
#define P1H_SPLIT_LAUNCH(KQ, KT)                                                                                                     \
  do {                                                                                                                             \
    auto kern = P1H_SPLIT_KERN(KQ, KT);                                                                                            \
    const unsigned grid2 = static_cast<unsigned>(cdiv(rows, 64));                                                                  \
    const size_t smem2 = sizeof(double) * 2 * 32 * kMaxLen;                                                                        \
    const int ipf2 = cc ? 0 : static_cast<int>((static_cast<int64_t>(ctx->sm_count) * 5 * 64 * pfd_env / 100) & ~static_cast<int64_t>(63)); \
    kern<<<grid2, 128, smem2, ctx->stream>>>(irows, itotal, p->p1h_qw, p->p1h_tw, p->p1h_rowinfo, mesh->node_coords, mesh->cell_coords,   \
                                             p->outer, active, row_list, ifirst, ipf2, P, d_values);                                  \
  } while (0)
  if (p->p1h_kq == 0) P1H_LAUNCH(0, 8);
  else if (p->p1h_kt == 0) P1H_LAUNCH(4, 0);
  else if (p->p1h_kq == 2) {
    static const int minb_env = [] { const char* e = std::getenv("LFGPU_P1H_MINB"); return e != nullptr ? std::atoi(e) : 0; }();
    if (minb_env >= 5 && mode == 1 && !tensor && !cc) {  // experiment: more resident CTAs at fewer registers (config C2's instantiation)
      auto kern = minb_env == 5 ? k_assemble_p1_rows<2, 4, false, false, 1, 1, 5> : k_assemble_p1_rows<2, 4, false, false, 1, 1, 6>;
      const int ipf5 = static_cast<int>((static_cast<int64_t>(ctx->sm_count) * minb_env * threads * pfd_env / 100) & ~static_cast<int64_t>(127));
      kern<<<grid, threads, smem, ctx->stream>>>(irows, itotal, p->p1h_qw, p->p1h_tw, p->p1h_rowinfo, mesh->node_coords, mesh->cell_coords,
                                                 p->outer, active, row_list, ifirst, ipf5, P, d_values);
    } else if (split_env) {
      P1H_SPLIT_LAUNCH(2, 4);
    } else {
      P1H_LAUNCH(2, 4);
    }
  }
  else { if (split_env) P1H_SPLIT_LAUNCH(4, 8); else P1H_LAUNCH(4, 8); }
#undef P1H_SPLIT_KERN
#undef P1H_SPLIT_LAUNCH
#undef P1H_KERN
#undef P1H_LAUNCH
  LFGPU_LAUNCH_CHECK(ctx);
  return LFGPU_OK;
}

}  // namespace lfgpu
