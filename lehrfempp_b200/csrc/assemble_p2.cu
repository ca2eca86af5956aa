// P2 (FeLagrangeO2Tria) fast path of the numeric pass: "row kernels" that own matrix rows in registers (product code).
//
// Same mathematics as the generic path (uscalfe/loc_comp_ellbvp.h:266-339 with FeLagrangeO2Tria, lagr_fe.h:399-546, and
// TriaO1, geometry/tria_o1.cc:50-74; affine cells with constant coefficients, i.e. A_K = sum_ij M_ij Khat^{ji} + gamma
// |det| Mhat with the reference tensors of the provider's default rule), same output (the values of the compressed matrix
// of the symbolic pass).  What changes is the decomposition: ncu showed the item kernel bound by
// shared-memory work and barriers per item, so -- as the vertex-fan kernel does for P1 -- one thread owns one ROW:
//   * a VERTEX row (dof = mesh node i, interior vertex of valence 6): ring n_0..n_5 of neighbour nodes in fan order; cell
//     k is the triangle (i, n_k, n_k+1) taken with i as local vertex 0, so only row 0 of the reference tensors is needed;
//     its six entries go to the diagonal, the two neighbour columns, the two spoke-edge columns (each summed over the two
//     cells sharing them: a rolling register) and the rim-edge column (one cell).  19 stored values.
//   * an EDGE row (dof = mesh edge e with two adjacent cells): endpoints p, q, opposite vertices o_1, o_2; both cells are
//     taken as (p, q, o) so only row 3 (local edge 0) of the reference tensors is needed.  9 stored values.
// Taking a cell with another local numbering than the mesh's is legitimate because the P2 Lagrange basis is invariant
// under vertex permutations and the default rule (degree 4) integrates stiffness and mass products exactly: the entries
// are the same integrals, rounded differently (parity bar 1e-12, observed ~1e-15).  The slot of every column inside
// the row comes from the scatter map of the symbolic pass and is stored in the plan (5 / 4 bits per entry).
// Each warp stages its 32 consecutive rows in shared memory and writes the values as full 128-byte lines.
// Every other row (boundary vertices and edges, valence != 6, ...) is listed as irregular and computed by the generic
// gather kernel (assemble.cu); the row kernels leave those rows untouched.
#include <algorithm>
#include <cstdlib>

#include <cub/cub.cuh>

#include "lfgpu_internal.cuh"
#include "rows_p2_core.h"

namespace lfgpu {
namespace {

constexpr int kRing = 6;         // valence handled by the vertex-row kernel
constexpr int kVertexRowLen = 19;  // 1 + 6 neighbours + 6 spokes + 6 rim edges
constexpr int kEdgeRowLen = 9;
constexpr uint32_t kNil = 0xFFFFFFFFu;

// ---- plan construction ------------------------------------------------------------------------------------------------
// dof table == [node ids | edge dofs >= n_nodes] and all cells are triangles with six local dofs?
__global__ void k_p2_check(int64_t n_cells, int stride, int64_t n_nodes, const int32_t* __restrict__ dofs, const uint8_t* __restrict__ nldof,
                           const uint32_t* __restrict__ cell_nodes, int* __restrict__ bad) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[c];
  const int32_t* d = dofs + c * stride;
  bool ok = v.w == kNil && nldof[c] == 6 && d[0] == static_cast<int32_t>(v.x) && d[1] == static_cast<int32_t>(v.y) &&
            d[2] == static_cast<int32_t>(v.z);
  if (ok) ok = d[3] >= n_nodes && d[4] >= n_nodes && d[5] >= n_nodes;
  if (!ok) *bad = 1;
}

// vertex row r < n_nodes: ring of exactly six cells closing around the node -> nbr[k][r] = n_k, slots[0..2][r]
// (5 bits per ring position: neighbour columns | spoke-edge columns | rim-edge columns); irregular rows get nbr[0][r] = -1
__global__ void k_p2_vertex_plan(int64_t n_nodes, int o_stride, int pos_row, const int32_t* __restrict__ adj_ptr,
                                 const uint32_t* __restrict__ adj, const uint32_t* __restrict__ cell_nodes,
                                 const uint8_t* __restrict__ pos, const int32_t* __restrict__ outer, int32_t* __restrict__ nbr,
                                 uint32_t* __restrict__ slots, uint8_t* __restrict__ irregular, int cc) {
  // cc != 0 (mesh with per-cell corner coordinates): nbr[k][r] = cell << 4 | (local index of n_k) << 2 | local index of n_k+1 of
  // ring cell k (the node itself is the third corner) instead of the node number n_k
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_nodes) return;
  const int32_t it0 = adj_ptr[r];
  const int m = adj_ptr[r + 1] - it0;
  bool ok = (m == kRing) && (outer[r + 1] - outer[r] == kVertexRowLen);
  uint32_t ja[kRing], ka[kRing], ring[kRing];
  int cell_a[kRing];       // local index of the node in the cell
  int64_t cell_id[kRing];
  int ord[kRing];          // adjacent cell at ring position k
  bool fwd[kRing];         // cell k runs n_k = next local vertex, n_k+1 = the one after (else the other way round)
  if (ok) {
    for (int t = 0; t < kRing; ++t) {
      const uint32_t item = adj[it0 + t];
      cell_id[t] = item >> 4;
      cell_a[t] = static_cast<int>(item & 15U);
      if (cell_a[t] > 2) {
        ok = false;
        cell_a[t] = 0;
      }
      const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[cell_id[t]];
      const uint32_t vv[3] = {v.x, v.y, v.z};
      ja[t] = vv[(cell_a[t] + 1) % 3];
      ka[t] = vv[(cell_a[t] + 2) % 3];
    }
  }
  if (ok) {
    unsigned used = 1U;
    ring[0] = ja[0];
    ord[0] = 0;
    fwd[0] = true;
    uint32_t cur = ka[0];
    for (int k = 1; k < kRing && ok; ++k) {
      int nxt = -1;
      for (int u = 0; u < kRing; ++u) {
        if (!(used & (1U << u)) && (ja[u] == cur || ka[u] == cur)) {
          nxt = u;
          break;
        }
      }
      if (nxt < 0) {
        ok = false;
        break;
      }
      used |= 1U << nxt;
      ring[k] = cur;
      ord[k] = nxt;
      fwd[k] = (ja[nxt] == cur);
      cur = fwd[k] ? ka[nxt] : ja[nxt];
    }
    if (ok && cur != ring[0]) ok = false;
    // six distinct neighbours, none of them the node itself
    for (int k = 0; k < kRing && ok; ++k) {
      if (ring[k] == static_cast<uint32_t>(r)) ok = false;
      for (int u = 0; u < k; ++u)
        if (ring[u] == ring[k]) ok = false;
    }
  }
  uint32_t w[3] = {0U, 0U, 0U};
  if (ok) {
    int sum = 0, sd = -1;
    for (int k = 0; k < kRing; ++k) {
      const int u = ord[k];
      const int a = cell_a[u], vb = (a + 1) % 3, vc = (a + 2) % 3;
      const uint8_t* prow = pos + (cell_id[u] * o_stride + a) * static_cast<int64_t>(pos_row);
      const int s_n = prow[fwd[k] ? vb : vc];       // column n_k
      const int s_s = prow[3 + (fwd[k] ? a : vc)];  // spoke edge (i, n_k)
      const int s_r = prow[3 + vb];                 // rim edge (n_k, n_k+1)
      const int s_d = prow[a];
      if (sd >= 0 && s_d != sd) ok = false;
      sd = s_d;
      if (s_n >= kVertexRowLen || s_s >= kVertexRowLen || s_r >= kVertexRowLen) ok = false;
      w[0] |= static_cast<uint32_t>(s_n & 31) << (5 * k);
      w[1] |= static_cast<uint32_t>(s_s & 31) << (5 * k);
      w[2] |= static_cast<uint32_t>(s_r & 31) << (5 * k);
      sum += s_n + s_s + s_r;
    }
    // the 19 slots must be a permutation of 0..18 (the kernel recovers the diagonal slot as 171 - sum of the others)
    if (ok && sd != 171 - sum) ok = false;
    if (ok) {
      unsigned seen = 1U << sd;
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < kRing; ++k) seen |= 1U << ((w[j] >> (5 * k)) & 31U);
      if (seen != (1U << kVertexRowLen) - 1U) ok = false;
    }
  }
  if (ok && cc) {
    for (int k = 0; k < kRing; ++k) {
      const int u = ord[k];
      const int a = cell_a[u], vb = (a + 1) % 3, vc = (a + 2) % 3;
      ring[k] = (static_cast<uint32_t>(cell_id[u]) << 4) | static_cast<uint32_t>((fwd[k] ? vb : vc) << 2) | static_cast<uint32_t>(fwd[k] ? vc : vb);
    }
  }
  for (int k = 0; k < kRing; ++k) nbr[static_cast<int64_t>(k) * n_nodes + r] = ok ? static_cast<int32_t>(ring[k]) : -1;
  for (int j = 0; j < 3; ++j) slots[static_cast<int64_t>(j) * n_nodes + r] = ok ? w[j] : 0U;
  irregular[r] = (!ok && m > 0) ? 1 : 0;
}

// the same for any closed ring of 3..8 cells (rows_p2_core.h): gnbr[k][r] = n_k (-1 beyond the ring / for an irregular row),
// gslots[0..5][r]; replaces the valence-6 plan on meshes with other valences (unstructured input)
__global__ void k_p2_vertex_plan_general(int64_t n_nodes, int o_stride, int pos_row, const int32_t* __restrict__ adj_ptr,
                                         const uint32_t* __restrict__ adj, const uint32_t* __restrict__ cell_nodes,
                                         const uint8_t* __restrict__ pos, const int32_t* __restrict__ outer, int32_t* __restrict__ gnbr,
                                         uint32_t* __restrict__ gslots, uint8_t* __restrict__ irregular) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_nodes) return;
  const int32_t it0 = adj_ptr[r];
  const int m = adj_ptr[r + 1] - it0;
  int32_t ring[p2::kMaxRing];
  uint32_t w[p2::kSlotWords];
  const bool ok = p2::vertex_plan_general(r, m, adj + it0, cell_nodes, pos, o_stride, pos_row, outer[r + 1] - outer[r], ring, w);
  for (int k = 0; k < p2::kMaxRing; ++k) gnbr[static_cast<int64_t>(k) * n_nodes + r] = ok ? ring[k] : -1;
  for (int j = 0; j < p2::kSlotWords; ++j) gslots[static_cast<int64_t>(j) * n_nodes + r] = ok ? w[j] : 0U;
  irregular[r] = (!ok && m > 0) ? 1 : 0;
}

// edge row r = n_nodes + e with exactly two adjacent cells: enb[0..3][e] = p, q, o_1, o_2; eslots[e] = 8 x 4 bits:
// columns p, q, o_1, o_2, edge (q, o_1), edge (o_1, p), edge (q, o_2), edge (o_2, p); the row's own slot is 36 - sum
__global__ void k_p2_edge_plan(int64_t n_nodes, int64_t n_edges, int o_stride, int pos_row, const int32_t* __restrict__ adj_ptr,
                               const uint32_t* __restrict__ adj, const uint32_t* __restrict__ cell_nodes,
                               const uint8_t* __restrict__ pos, const int32_t* __restrict__ outer, int32_t* __restrict__ enb,
                               uint32_t* __restrict__ eslots, uint8_t* __restrict__ irregular, int cc) {
  // cc != 0: enb[0..1][e] = cell << 4 | (local index of q) << 2 | local index of o for the two cells (p is the third corner: the
  // origin of the cell's frame), enb[2..3] unused
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int64_t r = n_nodes + e;
  const int32_t it0 = adj_ptr[r];
  const int m = adj_ptr[r + 1] - it0;
  bool ok = (m == 2) && (outer[r + 1] - outer[r] == kEdgeRowLen);
  uint32_t ids[4] = {0, 0, 0, 0};
  uint32_t w = 0;
  if (ok) {
    const uint32_t i1 = adj[it0], i2 = adj[it0 + 1];
    const int64_t c1 = i1 >> 4, c2 = i2 >> 4;
    const int a1 = static_cast<int>(i1 & 15U), a2 = static_cast<int>(i2 & 15U);
    if (a1 < 3 || a1 > 5 || a2 < 3 || a2 > 5) ok = false;
    if (ok) {
      const int j1 = a1 - 3, j2 = a2 - 3;
      const uint4 v1 = reinterpret_cast<const uint4*>(cell_nodes)[c1];
      const uint4 v2 = reinterpret_cast<const uint4*>(cell_nodes)[c2];
      const uint32_t n1[3] = {v1.x, v1.y, v1.z}, n2[3] = {v2.x, v2.y, v2.z};
      const uint32_t p = n1[j1], q = n1[(j1 + 1) % 3], o1 = n1[(j1 + 2) % 3], o2 = n2[(j2 + 2) % 3];
      const bool same = (n2[j2] == p && n2[(j2 + 1) % 3] == q);
      const bool opposite = (n2[j2] == q && n2[(j2 + 1) % 3] == p);
      if (!(same || opposite) || o1 == o2) ok = false;
      const uint8_t* r1 = pos + (c1 * o_stride + a1) * static_cast<int64_t>(pos_row);
      const uint8_t* r2 = pos + (c2 * o_stride + a2) * static_cast<int64_t>(pos_row);
      int s[8];
      s[0] = r1[j1];                // p
      s[1] = r1[(j1 + 1) % 3];      // q
      s[2] = r1[(j1 + 2) % 3];      // o_1
      s[3] = r2[(j2 + 2) % 3];      // o_2
      s[4] = r1[3 + (j1 + 1) % 3];  // edge (q, o_1)
      s[5] = r1[3 + (j1 + 2) % 3];  // edge (o_1, p)
      // in cell 2 local edge j2+1 starts at the second endpoint of the shared edge, local edge j2+2 ends at the first
      s[6] = r2[3 + (same ? (j2 + 1) % 3 : (j2 + 2) % 3)];  // edge (q, o_2)
      s[7] = r2[3 + (same ? (j2 + 2) % 3 : (j2 + 1) % 3)];  // edge (o_2, p)
      const int self = r1[a1];
      if (r2[a2] != self) ok = false;
      int sum = 0;
      unsigned seen = 1U << self;
      for (int k = 0; k < 8; ++k) {
        if (s[k] >= kEdgeRowLen) ok = false;
        sum += s[k];
        seen |= 1U << (s[k] & 15);
        w |= static_cast<uint32_t>(s[k] & 15) << (4 * k);
      }
      if (self != 36 - sum || seen != (1U << kEdgeRowLen) - 1U) ok = false;
      ids[0] = p; ids[1] = q; ids[2] = o1; ids[3] = o2;
      if (cc) {
        // cell 1 is (p, q, o_1) = local (j1, j1+1, j1+2); cell 2 lists the edge as (p, q) (same) or as (q, p)
        const int q2 = same ? (j2 + 1) % 3 : j2;
        ids[0] = (static_cast<uint32_t>(c1) << 4) | static_cast<uint32_t>(((j1 + 1) % 3) << 2) | static_cast<uint32_t>((j1 + 2) % 3);
        ids[1] = (static_cast<uint32_t>(c2) << 4) | static_cast<uint32_t>(q2 << 2) | static_cast<uint32_t>((j2 + 2) % 3);
        ids[2] = ids[3] = 0;
      }
    }
  }
  for (int k = 0; k < 4; ++k) enb[static_cast<int64_t>(k) * n_edges + e] = ok ? static_cast<int32_t>(ids[k]) : -1;
  eslots[e] = ok ? w : 0U;
  irregular[r] = (!ok && m > 0) ? 1 : 0;
}

// ---- compact plan (round 2) ----------------------------------------------------------------------------------------------
// Both kernels moved 1.23 x the algorithmic bytes, the difference being the plan: 36 B per vertex row, 20 B per edge row.  When
// every ring / edge neighbour lies within +-32767 of the row's reference node (structured, refined and Morton-ordered meshes) and the
// slot words take at most 65535 distinct values (plan_dict.cu), the plan shrinks to
//   vertex row r: cd[0..2][r] = six 16-bit differences n_k - r, cidx[r] = number of its slot triple in the table     (14 B)
//   edge row e:   ce[0][e] = p, ce[1][e] = (q - p) | (o_1 - p) << 16, ce[2][e] = (o_2 - p) | table index << 16          (12 B)
// index 0xFFFF marks a row that is not planned.  Otherwise the plan stays as it was built.
__global__ void k_p2_vertex_compact(int64_t nn, const int32_t* __restrict__ nbr, uint32_t* __restrict__ cd, uint16_t* __restrict__ cidx,
                                    int* __restrict__ overflow) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= nn) return;
  uint32_t w[3] = {0U, 0U, 0U};
  if (nbr[r] < 0) {
    cidx[r] = 0xFFFFU;
  } else {
    bool fits = true;
    for (int s = 0; s < kRing; ++s) {
      const int64_t d = static_cast<int64_t>(nbr[static_cast<int64_t>(s) * nn + r]) - r;
      if (d < -32768 || d > 32767) fits = false;
      w[s >> 1] |= (static_cast<uint32_t>(d) & 0xFFFFU) << (16 * (s & 1));
    }
    if (!fits) *overflow = 1;
  }
  for (int j = 0; j < 3; ++j) cd[static_cast<int64_t>(j) * nn + r] = w[j];
}

// A row whose differences do not fit is written as "not planned" and counted: if they are few (first-use order leaves a handful at
// the seams of the edge families) the caller hands them to the generic kernel (irregular[row] = 1) instead of giving the format up.
__global__ void k_p2_edge_compact(int64_t ne, const int32_t* __restrict__ enb, const uint16_t* __restrict__ idx, uint32_t* __restrict__ ce,
                                  int* __restrict__ overflow) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  const int32_t p = enb[e];
  uint32_t w0 = 0U, w1 = 0U, w2 = 0xFFFF0000U;
  if (p >= 0) {
    const int64_t dq = static_cast<int64_t>(enb[ne + e]) - p, d1 = static_cast<int64_t>(enb[2 * ne + e]) - p,
                  d2 = static_cast<int64_t>(enb[3 * ne + e]) - p;
    if (dq < -32768 || dq > 32767 || d1 < -32768 || d1 > 32767 || d2 < -32768 || d2 > 32767) {
      atomicAdd(overflow, 1);
    } else {
      w0 = static_cast<uint32_t>(p);
      w1 = (static_cast<uint32_t>(dq) & 0xFFFFU) | (static_cast<uint32_t>(d1) << 16);
      w2 = (static_cast<uint32_t>(d2) & 0xFFFFU) | (static_cast<uint32_t>(idx[e]) << 16);
    }
  }
  ce[e] = w0;
  ce[ne + e] = w1;
  ce[2 * ne + e] = w2;
}
// rows the full plan covers but the compact one does not -> generic kernel
__global__ void k_p2_edge_handback(int64_t ne, const int32_t* __restrict__ enb, const uint32_t* __restrict__ ce, uint8_t* __restrict__ irregular) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e < ne && enb[e] >= 0 && (ce[2 * ne + e] >> 16) == 0xFFFFU) irregular[e] = 1;
}

// ---- the kernels --------------------------------------------------------------------------------------------------------
struct P2Params {
  double a00, a01, a10, a11;  // diffusion tensor as the row routine of assemble.cu uses it (transposed for row-major output)
  double gamma;
  // rows 0 (vertex kernel) and 3 (edge kernel) of the reference tensors; MODE 0 reads k01 as k01 + k10
  double vk00[6], vk01[6], vk10[6], vk11[6], vm[6];
  double ek00[6], ek01[6], ek10[6], ek11[6], em[6];
};

__device__ __forceinline__ double rcp64(double x) {
  // as in assemble.cu: hardware seed + two Newton steps (Jacobian determinants: no denormals, no zeros)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// row `a` (0 or 3, through the tables handed in) of the element matrix of the triangle (x0, x0 + A, x0 + B):
// J = [A B], M = |det| J^-1 alpha J^-T, t[b] = sum_ij M_ij Khat^{ji}[a][b] + gamma |det| Mhat[a][b]
// MODE 0: scalar alpha, gamma = 0 (M symmetric: three products per entry)
template <int MODE>
__device__ __forceinline__ void p2_row(const P2Params& P, const double (&k00)[6], const double (&k01)[6], const double (&k10)[6],
                                       const double (&k11)[6], const double (&km)[6], double ax, double ay, double bx, double by,
                                       double (&t)[6]) {
  const double det = ax * by - ay * bx;
  if (MODE == 0) {
    const double s = P.a00 * rcp64(fabs(det));
    const double m00 = s * (bx * bx + by * by), m01 = -s * (ax * bx + ay * by), m11 = s * (ax * ax + ay * ay);
#pragma unroll
    for (int b = 0; b < 6; ++b) t[b] = m00 * k00[b] + m01 * k01[b] + m11 * k11[b];
  } else {
    const double adet = fabs(det), idet = rcp64(det);
    const double i00 = by * idet, i01 = -bx * idet, i10 = -ay * idet, i11 = ax * idet;
    const double t00 = i00 * P.a00 + i01 * P.a10, t01 = i00 * P.a01 + i01 * P.a11;
    const double t10 = i10 * P.a00 + i11 * P.a10, t11 = i10 * P.a01 + i11 * P.a11;
    const double m00 = adet * (t00 * i00 + t01 * i01), m01 = adet * (t00 * i10 + t01 * i11);
    const double m10 = adet * (t10 * i00 + t11 * i01), m11 = adet * (t10 * i10 + t11 * i11);
    const double gm = adet * P.gamma;
#pragma unroll
    for (int b = 0; b < 6; ++b) t[b] = m00 * k00[b] + m01 * k10[b] + m10 * k01[b] + m11 * k11[b] + gm * km[b];
  }
}

__device__ __forceinline__ void prefetch_l2(const void* a) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a)); }

// L2 eviction priorities (HINT variants of the kernels): the value stream is written once and never read again (evict first), the
// node coordinates are gathered by several rows from different CTAs (evict last) -- the ncu capture of the edge-row kernel showed
// the coordinate array being read three times from DRAM because the value stream pushes it out of L2
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
template <bool HINT>
__device__ __forceinline__ double2 ld_coords(const double2* a, uint64_t pol) {
  if (HINT) {
    double2 v;
    asm("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(a), "l"(pol));
    return v;
  }
  return __ldg(a);
}
// edge vectors of a cell from ITS corners (cell_coords [n_cells][4] points): word = cell << 4 | ia << 2 | ib, origin = third corner
__device__ __forceinline__ void cell_vectors(const double2* cc, uint32_t cw, double& ax, double& ay, double& bx, double& by) {
  const double2* c = cc + 4 * static_cast<size_t>(cw >> 4);
  const int ia = (cw >> 2) & 3, ib = cw & 3;
  const double2 x0 = __ldg(c + (3 - ia - ib)), xa = __ldg(c + ia), xb = __ldg(c + ib);
  ax = xa.x - x0.x; ay = xa.y - x0.y;
  bx = xb.x - x0.x; by = xb.y - x0.y;
}
__device__ __forceinline__ int32_t sext16(uint32_t v) { return static_cast<int32_t>(static_cast<int16_t>(v & 0xFFFFU)); }

// copy-out shared by both kernels: the warp's stage is the image of the contiguous value range of its 32 rows.
// BULK: the image leaves through ONE bulk copy of the TMA unit (cp.async.bulk shared -> global, SASS UBLKCP) issued by lane 0
// instead of LEN shared-memory loads + LEN global stores per lane -- ncu showed the edge-row kernel limited by the L1 / LSU pipe
// (82 % busy: 9 scattered STS, 9 LDS, 9 STG and 11 loads per row), and the copy-out was half of that work.  A bulk copy needs
// 16-byte aligned addresses and sizes: the image is kept at the parity of its first global element (stage[odd + k] <-> values[wbase
// + k], odd = wbase & 1), a leading / trailing odd element is stored by an ordinary lane.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int LEN, bool BULK, bool HINT = false>
__device__ __forceinline__ void write_rows(bool staged, bool regular, bool in_range, int lane, int32_t v0, int32_t v1, int32_t wbase,
                                           const double* __restrict__ stage, const double* __restrict__ dst, double* __restrict__ values,
                                           double beta) {
  if (BULK) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my shared-memory writes -> visible to the async proxy
  __syncwarp();
  if (staged) {
    const unsigned ballot = __ballot_sync(0xffffffffU, in_range);
    if (ballot == 0) return;
    const int total = __shfl_sync(0xffffffffU, v1, 31 - __clz(ballot)) - wbase;
    double* out = values + wbase;
    if (beta != 0.0) {  // accumulate (assembler.h:84-88): the lanes read what is there; image element k sits at stage[odd + k]
      const int odd = BULK ? (wbase & 1) : 0;
#pragma unroll
      for (int k = 0; k < LEN; ++k) {
        const int idx = k * 32 + lane;
        if (idx < total) out[idx] = fma(beta, out[idx], stage[odd + idx]);
      }
      return;
    }
    if (BULK) {
      const int odd = wbase & 1;
      const int n_bulk = (total - odd) & ~1;  // elements [odd, odd + n_bulk) start and end on 16-byte boundaries
      if (lane == 0 && n_bulk > 0) {
        if (HINT) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(out + odd),
                       "r"(smem_u32(stage + 2 * odd)), "r"(n_bulk * 8), "l"(l2_policy_evict_first())
                       : "memory");
        } else {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + odd), "r"(smem_u32(stage + 2 * odd)),
                       "r"(n_bulk * 8)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (lane == 1 && odd && total > 0) out[0] = stage[odd];
      if (lane == 2 && odd + n_bulk < total) out[total - 1] = stage[odd + total - 1];
      if (lane == 0 && n_bulk > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the stage may go away now
      return;
    }
#pragma unroll
    for (int k = 0; k < LEN; ++k) {
      const int idx = k * 32 + lane;
      if (idx < total) out[idx] = stage[idx];
    }
  } else if (regular) {
    for (int k = 0; k < LEN; ++k) values[v0 + k] = beta == 0.0 ? dst[k] : fma(beta, values[v0 + k], dst[k]);
  }
}

// COMPACT: nbr = the 16-bit differences cd[3][n_rows], slots = the table of slot triples (uint4), cidx[n_rows] = table index
// HINT: L2 eviction priorities on the coordinate gathers and the value stream
// CC: node_coords is the mesh's cell_coords array, nbr holds (cell, corner) words (k_p2_vertex_plan)
template <int MODE, bool BULK, bool COMPACT, bool HINT, bool CC = false>
__global__ void __launch_bounds__(128, 6) k_p2_vertex_rows(int n_rows, const int32_t* __restrict__ nbr,
                                                         const uint32_t* __restrict__ slots, const double* __restrict__ node_coords,
                                                         const int32_t* __restrict__ outer, int pf_dist, P2Params P,
                                                         double* __restrict__ values, int first, int end, double beta,
                                                         const uint16_t* __restrict__ cidx) {
  // rows [first, end) of the n_rows vertex rows (the whole range, or one GPU's share of it)
  extern __shared__ __align__(16) double stage_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = first + blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = r < end;
  if (pf_dist > 0 && warp == 0) {
    // pull the lines of the CTA that runs about one wave later into L2 (see assemble_p1.cu): 4 lines per plan array
    // (6 ring + 3 slot arrays), 4 of row pointers, 16 of coordinates; COMPACT: 3 arrays of differences, 2 lines of indices, coordinates
    const int rp = first + blockIdx.x * blockDim.x + pf_dist;
    if (rp + 128 <= end) {
      if (COMPACT) {
        // (the row pointers are not prefetched: one 32-byte sector per warp is read, and only the copy-out waits for it)
        if (lane < 30) {
          const char* a;
          if (lane < 12) a = reinterpret_cast<const char*>(nbr + static_cast<size_t>(lane >> 2) * n_rows + rp) + (lane & 3) * 128;
          else if (lane < 14) a = reinterpret_cast<const char*>(cidx + rp) + (lane - 12) * 128;
          else a = reinterpret_cast<const char*>(node_coords + 2 * static_cast<size_t>(rp)) + (lane - 14) * 128;
          prefetch_l2(a);
        }
      } else {
        for (int L = lane; L < (CC ? 40 : 56); L += 32) {  // CC: corners are indexed by cell, not by row
          const char* a;
          if (L < 24) a = reinterpret_cast<const char*>(nbr + static_cast<size_t>(L >> 2) * n_rows + rp) + (L & 3) * 128;
          else if (L < 36) a = reinterpret_cast<const char*>(slots + static_cast<size_t>((L - 24) >> 2) * n_rows + rp) + (L & 3) * 128;
          else if (L < 40) a = reinterpret_cast<const char*>(outer + rp) + (L - 36) * 128;
          else a = reinterpret_cast<const char*>(node_coords + 2 * static_cast<size_t>(rp)) + (L - 40) * 128;
          prefetch_l2(a);
        }
      }
    }
  }
  int32_t v0 = 0, v1 = 0;
  int32_t nid[kRing];
  uint32_t w0 = 0, w1 = 0, w2 = 0;
  bool regular;
  int32_t wbase;
  if (COMPACT) {
    uint32_t d0 = 0, d1 = 0, d2 = 0, ix = 0xFFFFU;
    if (in_range) {
      d0 = __ldg(reinterpret_cast<const uint32_t*>(nbr) + r);
      d1 = __ldg(reinterpret_cast<const uint32_t*>(nbr) + static_cast<size_t>(n_rows) + r);
      d2 = __ldg(reinterpret_cast<const uint32_t*>(nbr) + 2 * static_cast<size_t>(n_rows) + r);
      ix = __ldg(cidx + r);
    }
    regular = in_range && ix != 0xFFFFU;
    nid[0] = r + sext16(d0); nid[1] = r + sext16(d0 >> 16);
    nid[2] = r + sext16(d1); nid[3] = r + sext16(d1 >> 16);
    nid[4] = r + sext16(d2); nid[5] = r + sext16(d2 >> 16);
    if (regular) {
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(slots) + ix);
      w0 = t.x; w1 = t.y; w2 = t.z;
    }
    if (__all_sync(0xffffffffU, regular)) {
      // 32 planned rows: 19 values each, one after the other -- only the first row pointer is read
      int32_t b = 0;
      if (lane == 0) b = __ldg(outer + r);
      wbase = __shfl_sync(0xffffffffU, b, 0);
      v0 = wbase + kVertexRowLen * lane;
      v1 = v0 + kVertexRowLen;
    } else {
      if (in_range) {
        v0 = __ldg(outer + r);
        v1 = __ldg(outer + r + 1);
      }
      wbase = __shfl_sync(0xffffffffU, v0, 0);
    }
  } else {
#pragma unroll
    for (int s = 0; s < kRing; ++s) nid[s] = -1;
    if (in_range) {
      v0 = __ldg(outer + r);
      v1 = __ldg(outer + r + 1);
#pragma unroll
      for (int s = 0; s < kRing; ++s) nid[s] = __ldg(nbr + static_cast<size_t>(s) * n_rows + r);
      w0 = __ldg(slots + r);
      w1 = __ldg(slots + static_cast<size_t>(n_rows) + r);
      w2 = __ldg(slots + 2 * static_cast<size_t>(n_rows) + r);
    }
    regular = in_range && nid[0] >= 0;
    wbase = __shfl_sync(0xffffffffU, v0, 0);
  }
  const bool staged = !__any_sync(0xffffffffU, in_range && !regular && v1 > v0);
  double* stage = stage_all + warp * (32 * (kVertexRowLen + 1));
  double* dst = stage + (staged ? (v0 - wbase) + (BULK ? (wbase & 1) : 0) : lane * (kVertexRowLen + 1));
  if (regular) {
    const uint64_t keep = HINT ? l2_policy_evict_last() : 0ULL;
    const double2* nc = reinterpret_cast<const double2*>(node_coords);
    double dx[kRing], dy[kRing];
    if (!CC) {
      const double2 xi = __ldg(nc + r);
#pragma unroll
      for (int s = 0; s < kRing; ++s) {
        const double2 p = ld_coords<HINT>(nc + nid[s], keep);
        dx[s] = p.x - xi.x;
        dy[s] = p.y - xi.y;
      }
    }
    int ssum = 0;
#pragma unroll
    for (int s = 0; s < kRing; ++s) ssum += static_cast<int>((w0 >> (5 * s)) & 31U) + static_cast<int>((w1 >> (5 * s)) & 31U) +
                                            static_cast<int>((w2 >> (5 * s)) & 31U);
    double t[6];
    double ax, ay, bx, by;
    if (CC) cell_vectors(nc, static_cast<uint32_t>(nid[0]), ax, ay, bx, by);
    else { ax = dx[0]; ay = dy[0]; bx = dx[1]; by = dy[1]; }
    p2_row<MODE>(P, P.vk00, P.vk01, P.vk10, P.vk11, P.vm, ax, ay, bx, by, t);
    double diag = t[0];
    const double first_n = t[1], first_s = t[3];
    double carry_n = t[2], carry_s = t[5];
    dst[w2 & 31U] = t[4];
#pragma unroll
    for (int s = 1; s < kRing; ++s) {
      const int u = (s + 1 < kRing) ? s + 1 : 0;
      if (CC) cell_vectors(nc, static_cast<uint32_t>(nid[s]), ax, ay, bx, by);
      else { ax = dx[s]; ay = dy[s]; bx = dx[u]; by = dy[u]; }
      p2_row<MODE>(P, P.vk00, P.vk01, P.vk10, P.vk11, P.vm, ax, ay, bx, by, t);
      diag += t[0];
      dst[(w0 >> (5 * s)) & 31U] = carry_n + t[1];
      dst[(w1 >> (5 * s)) & 31U] = carry_s + t[3];
      dst[(w2 >> (5 * s)) & 31U] = t[4];
      carry_n = t[2];
      carry_s = t[5];
    }
    dst[w0 & 31U] = first_n + carry_n;
    dst[w1 & 31U] = first_s + carry_s;
    dst[171 - ssum] = diag;
  }
  write_rows<kVertexRowLen, BULK, HINT>(staged, regular, in_range, lane, v0, v1, wbase, stage, dst, values, beta);
}

// vertex rows with closed rings of 3..8 cells (rows_p2_core.h); rows of different lengths (1 + 3m) share a warp, the staged
// copy-out only needs their value ranges to be consecutive
template <int MODE>
__global__ void __launch_bounds__(128, 4) k_p2_vertex_rows_general(int first, int end, int n_rows, const int32_t* __restrict__ gnbr,
                                                                 const uint32_t* __restrict__ gslots, const double* __restrict__ node_coords,
                                                                 const int32_t* __restrict__ outer, p2::VertexParams P,
                                                                 double* __restrict__ values, double beta) {
  extern __shared__ double stage_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = first + blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = r < end;
  int32_t v0 = 0, v1 = 0;
  int32_t nid[p2::kMaxRing];
  uint32_t w[p2::kSlotWords];
#pragma unroll
  for (int s = 0; s < p2::kMaxRing; ++s) nid[s] = -1;
#pragma unroll
  for (int j = 0; j < p2::kSlotWords; ++j) w[j] = 0U;
  if (in_range) {
    v0 = __ldg(outer + r);
    v1 = __ldg(outer + r + 1);
#pragma unroll
    for (int s = 0; s < p2::kMaxRing; ++s) nid[s] = __ldg(gnbr + static_cast<size_t>(s) * n_rows + r);
#pragma unroll
    for (int j = 0; j < p2::kSlotWords; ++j) w[j] = __ldg(gslots + static_cast<size_t>(j) * n_rows + r);
  }
  const bool regular = in_range && nid[0] >= 0;
  const int32_t wbase = __shfl_sync(0xffffffffU, v0, 0);
  const bool staged = !__any_sync(0xffffffffU, in_range && !regular && v1 > v0);
  double* stage = stage_all + warp * (32 * (p2::kMaxVertexRowLen + 1));
  double* dst = stage + (staged ? v0 - wbase : lane * (p2::kMaxVertexRowLen + 1));
  if (regular) {
    const double2* nc = reinterpret_cast<const double2*>(node_coords);
    const double2 xi = __ldg(nc + r);
    double dx[p2::kMaxRing], dy[p2::kMaxRing];
#pragma unroll
    for (int s = 0; s < p2::kMaxRing; ++s) {
      const double2 q = __ldg(nc + (nid[s] >= 0 ? nid[s] : r));
      dx[s] = q.x - xi.x;
      dy[s] = q.y - xi.y;
    }
    p2::vertex_row_general<MODE>(P, dx, dy, w, dst);
  }
  __syncwarp();
  if (staged) {
    const unsigned ballot = __ballot_sync(0xffffffffU, in_range);
    if (ballot == 0) return;
    const int total = __shfl_sync(0xffffffffU, v1, 31 - __clz(ballot)) - wbase;
    double* out = values + wbase;
#pragma unroll
    for (int k = 0; k < p2::kMaxVertexRowLen; ++k) {
      const int idx = k * 32 + lane;
      if (idx < total) out[idx] = beta == 0.0 ? stage[idx] : fma(beta, out[idx], stage[idx]);
    }
  } else if (regular) {
    for (int k = 0; k < v1 - v0; ++k) values[v0 + k] = beta == 0.0 ? dst[k] : fma(beta, values[v0 + k], dst[k]);
  }
}

// 12 CTAs per SM = the occupancy of the measured kernel (40 registers; the row-range arguments had pushed ptxas to 46 -> 10 CTAs)
// COMPACT: enb = ce[3][n_edges] (p | differences | table index), eslots = the table of slot words (uint4, .x used)
// CC: node_coords is the mesh's cell_coords array, enb[0..1] hold the (cell, corner) words of the two cells
template <int MODE, bool BULK, bool COMPACT, bool HINT, bool CC = false>
__global__ void __launch_bounds__(128, 12) k_p2_edge_rows(int n_edges, int row0, const int32_t* __restrict__ enb,
                                                       const uint32_t* __restrict__ eslots, const double* __restrict__ node_coords,
                                                       const int32_t* __restrict__ outer, int pf_dist, int pfc_dist, P2Params P,
                                                       double* __restrict__ values, int first, int end, double beta) {
  // edge rows [first, end) of n_edges; edge e is matrix row row0 + e
  extern __shared__ __align__(16) double stage_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = first + blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = e < end;
  if (pf_dist > 0 && warp == 0) {
    const int ep = first + blockIdx.x * blockDim.x + pf_dist;
    if (COMPACT) {
      // 4 lines per plan word array; the row pointers (one sector per warp is read) only on request: pfc_dist != 0
      if (ep + 128 <= end && lane < (pfc_dist != 0 ? 16 : 12)) {
        const char* a;
        if (lane < 12) a = reinterpret_cast<const char*>(enb + static_cast<size_t>(lane >> 2) * n_edges + ep) + (lane & 3) * 128;
        else a = reinterpret_cast<const char*>(outer + row0 + ep) + (lane - 12) * 128;
        prefetch_l2(a);
      }
    } else if (ep + 128 <= end && lane < 24) {  // 4 lines per id array, 4 of slots, 4 of row pointers
      const char* a;
      if (lane < 16) a = reinterpret_cast<const char*>(enb + static_cast<size_t>(lane >> 2) * n_edges + ep) + (lane & 3) * 128;
      else if (lane < 20) a = reinterpret_cast<const char*>(eslots + ep) + (lane - 16) * 128;
      else a = reinterpret_cast<const char*>(outer + row0 + ep) + (lane - 20) * 128;
      prefetch_l2(a);
    }
  }
  // (a per-thread L2 prefetch of the coordinates of the row one wave ahead, as in k_p3_edge_rows, was measured here in round 2:
  // 0.53 -> 0.82 ms -- this kernel is bound by L1 / LSU work per row, and eight more memory instructions per row make it worse;
  // pfc_dist now only switches the row-pointer prefetch of the compact plan)
  int32_t v0 = 0, v1 = 0;
  int32_t ip = -1, iq = 0, io1 = 0, io2 = 0;
  uint32_t w = 0;
  bool regular;
  int32_t wbase;
  if (COMPACT) {
    uint32_t c0 = 0, c1 = 0, c2 = 0xFFFF0000U;
    if (in_range) {
      c0 = __ldg(reinterpret_cast<const uint32_t*>(enb) + e);
      c1 = __ldg(reinterpret_cast<const uint32_t*>(enb) + static_cast<size_t>(n_edges) + e);
      c2 = __ldg(reinterpret_cast<const uint32_t*>(enb) + 2 * static_cast<size_t>(n_edges) + e);
    }
    regular = in_range && (c2 >> 16) != 0xFFFFU;
    ip = static_cast<int32_t>(c0);
    iq = ip + sext16(c1);
    io1 = ip + sext16(c1 >> 16);
    io2 = ip + sext16(c2);
    if (regular) w = __ldg(eslots + 4 * static_cast<size_t>(c2 >> 16));
    if (__all_sync(0xffffffffU, regular)) {
      int32_t b = 0;
      if (lane == 0) b = __ldg(outer + row0 + e);
      wbase = __shfl_sync(0xffffffffU, b, 0);
      v0 = wbase + kEdgeRowLen * lane;
      v1 = v0 + kEdgeRowLen;
    } else {
      if (in_range) {
        v0 = __ldg(outer + row0 + e);
        v1 = __ldg(outer + row0 + e + 1);
      }
      wbase = __shfl_sync(0xffffffffU, v0, 0);
    }
  } else {
    if (in_range) {
      v0 = __ldg(outer + row0 + e);
      v1 = __ldg(outer + row0 + e + 1);
      ip = __ldg(enb + e);
      iq = __ldg(enb + static_cast<size_t>(n_edges) + e);
      if (!CC) {
        io1 = __ldg(enb + 2 * static_cast<size_t>(n_edges) + e);
        io2 = __ldg(enb + 3 * static_cast<size_t>(n_edges) + e);
      }
      w = __ldg(eslots + e);
    }
    regular = in_range && ip >= 0;
    wbase = __shfl_sync(0xffffffffU, v0, 0);
  }
  const bool staged = !__any_sync(0xffffffffU, in_range && !regular && v1 > v0);
  double* stage = stage_all + warp * (32 * (kEdgeRowLen + 1));
  double* dst = stage + (staged ? (v0 - wbase) + (BULK ? (wbase & 1) : 0) : lane * (kEdgeRowLen + 1));
  if (regular) {
    const uint64_t keep = HINT ? l2_policy_evict_last() : 0ULL;
    const double2* nc = reinterpret_cast<const double2*>(node_coords);
    double t1[6], t2[6];
    if (CC) {
      double ax, ay, bx, by;
      cell_vectors(nc, static_cast<uint32_t>(ip), ax, ay, bx, by);
      p2_row<MODE>(P, P.ek00, P.ek01, P.ek10, P.ek11, P.em, ax, ay, bx, by, t1);
      cell_vectors(nc, static_cast<uint32_t>(iq), ax, ay, bx, by);
      p2_row<MODE>(P, P.ek00, P.ek01, P.ek10, P.ek11, P.em, ax, ay, bx, by, t2);
    } else {
      const double2 xp = ld_coords<HINT>(nc + ip, keep), xq = ld_coords<HINT>(nc + iq, keep), x1 = ld_coords<HINT>(nc + io1, keep),
                    x2 = ld_coords<HINT>(nc + io2, keep);
      const double ax = xq.x - xp.x, ay = xq.y - xp.y;
      p2_row<MODE>(P, P.ek00, P.ek01, P.ek10, P.ek11, P.em, ax, ay, x1.x - xp.x, x1.y - xp.y, t1);
      p2_row<MODE>(P, P.ek00, P.ek01, P.ek10, P.ek11, P.em, ax, ay, x2.x - xp.x, x2.y - xp.y, t2);
    }
    int ssum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) ssum += static_cast<int>((w >> (4 * k)) & 15U);
    dst[w & 15U] = t1[0] + t2[0];
    dst[(w >> 4) & 15U] = t1[1] + t2[1];
    dst[(w >> 8) & 15U] = t1[2];
    dst[(w >> 12) & 15U] = t2[2];
    dst[(w >> 16) & 15U] = t1[4];
    dst[(w >> 20) & 15U] = t1[5];
    dst[(w >> 24) & 15U] = t2[4];
    dst[(w >> 28) & 15U] = t2[5];
    dst[36 - ssum] = t1[3] + t2[3];
  }
  write_rows<kEdgeRowLen, BULK, HINT>(staged, regular, in_range, lane, v0, v1, wbase, stage, dst, values, beta);
}

__global__ void k_count_flags(int64_t n, const uint8_t* __restrict__ flag, int* __restrict__ cnt) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const unsigned b = __ballot_sync(0xffffffffU, i < n && flag[i] != 0);
  if ((threadIdx.x & 31) == 0 && b != 0) atomicAdd(cnt, __popc(b));
}

}  // namespace

// ---- host side -----------------------------------------------------------------------------------------------------------
int p2_rows_prepare(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, lfgpu_pattern* p) {
  if (p->p2_state != 0) return LFGPU_OK;
  p->p2_state = -1;
  const int64_t nn = mesh->n_nodes, ne = p->n_outer - mesh->n_nodes;
  // cells with their own corner coordinates (cc): the plan carries (cell, corner) words instead of node numbers -- 27 bits of cell
  const int cc = mesh->cell_coords != nullptr ? 1 : 0;
  if (mesh->n_quad != 0 || p->i_dofs != p->o_dofs || ne <= 0 || p->pos_bytes != 1 || p->pos == nullptr || p->n_outer >= (1LL << 31) - 256 ||
      (cc && p->n_cells >= (1LL << 27)))
    return LFGPU_OK;
  p->p2_cc = cc != 0;
  cudaStream_t st = ctx->stream;
  int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 256);
  LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_flags, 0, 16, st));
  k_p2_check<<<static_cast<unsigned>(cdiv(p->n_cells, 256)), 256, 0, st>>>(p->n_cells, p->o_stride, nn, p->o_dofs, p->o_nldof, mesh->cell_nodes,
                                                                            d_flags);
  LFGPU_LAUNCH_CHECK(ctx);
  int bad = 1;
  LFGPU_CUDA_CHECK(ctx, cudaMemcpyAsync(&bad, d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
  LFGPU_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
  if (bad != 0) return LFGPU_OK;  // not the dof layout of FeSpaceLagrangeO2 on triangles: stay with the generic kernels
  uint8_t* flag = nullptr;
  int32_t* iota = nullptr;
  int64_t* d_num = nullptr;
  void* tmp = nullptr;
  auto cleanup = [&]() { cudaFree(flag); cudaFree(iota); cudaFree(d_num); cudaFree(tmp); };
  auto drop_plan = [&]() {
    cudaFree(p->p2v_nbr); cudaFree(p->p2v_slots); cudaFree(p->p2e_nbr); cudaFree(p->p2e_slots); cudaFree(p->p2_irregular);
    cudaFree(p->p2g_nbr); cudaFree(p->p2g_slots); cudaFree(p->p2v_cidx); cudaFree(p->p2e_newid); cudaFree(p->p2e_xy);
    p->p2e_newid = nullptr; p->p2e_xy = nullptr;
    p->p2v_cidx = nullptr; p->p2_compact_v = false; p->p2_compact_e = false;
    p->p2v_nbr = nullptr; p->p2v_slots = nullptr; p->p2e_nbr = nullptr; p->p2e_slots = nullptr; p->p2_irregular = nullptr;
    p->p2g_nbr = nullptr; p->p2g_slots = nullptr; p->p2_general = false;
  };
#define P2_CHECK(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      set_last_error(ctx, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
      cleanup();                                                                    \
      drop_plan();                                                                  \
      return LFGPU_ERR_CUDA;                                                        \
    }                                                                               \
  } while (0)
  // + 128 entries of slack: the prefetch of the kernels reads whole lines
  P2_CHECK(cudaMalloc(&p->p2v_nbr, sizeof(int32_t) * (kRing * static_cast<size_t>(nn) + 128)));
  P2_CHECK(cudaMalloc(&p->p2v_slots, sizeof(uint32_t) * (3 * static_cast<size_t>(nn) + 128)));
  P2_CHECK(cudaMalloc(&p->p2e_nbr, sizeof(int32_t) * (4 * static_cast<size_t>(ne) + 128)));
  P2_CHECK(cudaMalloc(&p->p2e_slots, sizeof(uint32_t) * (static_cast<size_t>(ne) + 128)));
  P2_CHECK(cudaMalloc(&flag, p->n_outer));
  k_p2_vertex_plan<<<static_cast<unsigned>(cdiv(nn, 128)), 128, 0, st>>>(nn, p->o_stride, p->pos_row, p->adj_ptr, p->adj, mesh->cell_nodes,
                                                                          static_cast<const uint8_t*>(p->pos), p->outer, p->p2v_nbr,
                                                                          p->p2v_slots, flag, cc);
  ctx->launches++;
  k_p2_edge_plan<<<static_cast<unsigned>(cdiv(ne, 128)), 128, 0, st>>>(nn, ne, p->o_stride, p->pos_row, p->adj_ptr, p->adj, mesh->cell_nodes,
                                                                        static_cast<const uint8_t*>(p->pos), p->outer, p->p2e_nbr,
                                                                        p->p2e_slots, flag, cc);
  ctx->launches++;
  P2_CHECK(cudaGetLastError());
  // Unstructured meshes: vertex rows whose ring is not exactly six cells would all go to the generic kernel.  On request
  // (LFGPU_P2_GENERAL=1; the kernel's host/device core is checked on the CPU, its CUDA wrapper has not been on a B200 yet)
  // the plan for closed rings of 3..8 cells replaces the valence-6 plan of the vertex rows.
  // automatic (default): the general plan when more than 5 % of the vertex rows miss the valence-6 plan (Gmsh / Delaunay meshes:
  // measured on workload u2, 1.0e6 triangles: 0.297 -> 0.153 ms); LFGPU_P2_GENERAL=1 forces it, =0 never builds it
  static const int general_env = [] { const char* e = std::getenv("LFGPU_P2_GENERAL"); return e == nullptr ? -1 : (e[0] == '1' ? 1 : 0); }();
  if (general_env != 0 && cc == 0) {  // (the general-valence kernel reads node positions)
    int* d_cnt = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 384);
    auto count_flags = [&](const uint8_t* fl, int* h) -> cudaError_t {
      cudaError_t e = cudaMemsetAsync(d_cnt, 0, sizeof(int), st);
      if (e != cudaSuccess) return e;
      k_count_flags<<<static_cast<unsigned>(cdiv(nn, 256)), 256, 0, st>>>(nn, fl, d_cnt);
      ctx->launches++;
      e = cudaMemcpyAsync(h, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st);
      return e == cudaSuccess ? cudaStreamSynchronize(st) : e;
    };
    int cnt6 = 0, cntg = 0;
    P2_CHECK(count_flags(flag, &cnt6));
    if (general_env == 1 || static_cast<int64_t>(cnt6) * 20 > nn) {
      // candidate: plan the vertex rows for closed rings of 3..8 cells; adopted if it takes more than 5 % of the vertex rows off
      // the generic kernel (boundary rows fail both plans, so a small structured mesh keeps the leaner valence-6 kernel)
      uint8_t* flag_g = nullptr;
      P2_CHECK(cudaMalloc(&flag_g, nn));
      cudaError_t eg = cudaMalloc(&p->p2g_nbr, sizeof(int32_t) * (p2::kMaxRing * static_cast<size_t>(nn) + 128));
      if (eg == cudaSuccess) eg = cudaMalloc(&p->p2g_slots, sizeof(uint32_t) * (p2::kSlotWords * static_cast<size_t>(nn) + 128));
      if (eg == cudaSuccess) {
        k_p2_vertex_plan_general<<<static_cast<unsigned>(cdiv(nn, 128)), 128, 0, st>>>(nn, p->o_stride, p->pos_row, p->adj_ptr, p->adj, mesh->cell_nodes,
                                                                         static_cast<const uint8_t*>(p->pos), p->outer, p->p2g_nbr, p->p2g_slots, flag_g);
        ctx->launches++;
        eg = cudaGetLastError();
      }
      if (eg == cudaSuccess) eg = count_flags(flag_g, &cntg);
      const bool adopt = eg == cudaSuccess && (general_env == 1 || static_cast<int64_t>(cnt6 - cntg) * 20 > nn);
      if (adopt) eg = cudaMemcpyAsync(flag, flag_g, nn, cudaMemcpyDeviceToDevice, st);
      if (eg == cudaSuccess) eg = cudaStreamSynchronize(st);
      cudaFree(flag_g);
      if (!adopt || eg != cudaSuccess) {
        cudaFree(p->p2g_nbr);
        cudaFree(p->p2g_slots);
        p->p2g_nbr = nullptr;
        p->p2g_slots = nullptr;
      }
      P2_CHECK(eg);
      p->p2_general = adopt;
    }
  }
  // how many rows the plans do not cover (decides whether the plan is worth refining; the exact list is made at the end)
  int64_t n_irr = 0;
  {
    int* d_cnt = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 384);
    int h_cnt = 0;
    P2_CHECK(cudaMemsetAsync(d_cnt, 0, sizeof(int), st));
    k_count_flags<<<static_cast<unsigned>(cdiv(p->n_outer, 256)), 256, 0, st>>>(p->n_outer, flag, d_cnt);
    ctx->launches++;
    P2_CHECK(cudaMemcpyAsync(&h_cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
    P2_CHECK(cudaStreamSynchronize(st));
    n_irr = h_cnt;
  }
  // the edge rows' own copy of the node positions, in the order the rows use them (plan_dict.cu: edge_node_order; LFGPU_EDGE_ORDER=0
  // keeps the mesh's array): on the builder's numbering the edge-row kernel is 30 % (P2) / 11 % (P3) faster with it
  static const bool order_env = [] { const char* e = std::getenv("LFGPU_EDGE_ORDER"); return e == nullptr || e[0] != '0'; }();
  if (order_env && cc == 0 && n_irr * 2 <= p->n_outer) {
    if (edge_node_order(ctx, nn, ne, p->p2e_nbr, &p->p2e_newid) != LFGPU_OK) {
      cleanup();
      drop_plan();
      return LFGPU_ERR_CUDA;
    }
    if (p->p2e_newid != nullptr) {
      P2_CHECK(cudaMalloc(&p->p2e_xy, sizeof(double) * (2 * static_cast<size_t>(nn) + 32)));
      p->p2e_xy_version = 0;
      p->p2e_xy_mesh = nullptr;
    }
  }
  // compact plan (see k_p2_vertex_compact): tried once the plan stands; any failure to fit leaves the arrays as they are
  // LFGPU_P2_COMPACT: 1 = both row classes, v = vertex rows only, e = edge rows only, 0 = off -- measured at config C3 (B200,
  // profiles/r02_p2_rows_compact*_hints*.json): DRAM traffic 4.25 -> 3.87 GB (1.23 -> 1.12 x algorithmic), vertex rows 0.263 -> 0.257 ms,
  // edge rows 0.524 -> 0.549 ms: the kernels wait on dependent loads (plan -> coordinates), not on DRAM bandwidth, and the table
  // lookup adds one more
  // The vertex rows gain 2 % (0.2626 -> 0.2571 / 0.2622 -> 0.2565 ms in two runs) and their plan shrinks from 36 to 14 B -> default 'v'.
  // The edge rows stay on the full plan also on their own coordinate copy (edge_node_order above): compact 0.419 ms / 769 MB read,
  // full 0.404 ms / 968 MB read (ncu: 5.9 against 6.75 TB/s -- the dependent table lookup and the unprefetched row pointer cost more
  // than 8 B per row save)
  static const char compact_env = [] { const char* e = std::getenv("LFGPU_P2_COMPACT"); return e == nullptr ? 'v' : e[0]; }();
  const bool compact_v = compact_env == '1' || compact_env == 'v', compact_e = compact_env == '1' || compact_env == 'e';
  if ((compact_v || compact_e) && cc == 0 && n_irr * 2 <= p->n_outer) {
    int* d_over = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 384);
    if (compact_v && !p->p2_general) {
      uint16_t* cidx = nullptr;
      uint32_t* cd = nullptr;
      void* dict = nullptr;
      int n_dict = -1, over = 1;
      P2_CHECK(cudaMalloc(&cidx, sizeof(uint16_t) * (static_cast<size_t>(nn) + 256)));
      const int rc = build_row_dict(ctx, 3, nn, p->p2v_slots, cidx, &dict, &n_dict);
      if (rc == LFGPU_OK && n_dict >= 0) {
        cudaError_t e = cudaMalloc(&cd, sizeof(uint32_t) * (3 * static_cast<size_t>(nn) + 128));
        if (e == cudaSuccess) e = cudaMemsetAsync(d_over, 0, sizeof(int), st);
        if (e == cudaSuccess) {
          k_p2_vertex_compact<<<static_cast<unsigned>(cdiv(nn, 256)), 256, 0, st>>>(nn, p->p2v_nbr, cd, cidx, d_over);
          ctx->launches++;
          e = cudaMemcpyAsync(&over, d_over, sizeof(int), cudaMemcpyDeviceToHost, st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) over = 1;
      }
      if (over == 0) {
        cudaFree(p->p2v_nbr);
        cudaFree(p->p2v_slots);
        p->p2v_nbr = reinterpret_cast<int32_t*>(cd);
        p->p2v_slots = static_cast<uint32_t*>(dict);
        p->p2v_cidx = cidx;
        p->p2_compact_v = true;
      } else {
        cudaFree(cidx);
        cudaFree(cd);
        cudaFree(dict);
        (void)cudaGetLastError();
      }
    }
    if (compact_e) {
      uint16_t* idx = nullptr;
      uint32_t* ce = nullptr;
      void* dict = nullptr;
      int n_dict = -1, over = 1;
      P2_CHECK(cudaMalloc(&idx, sizeof(uint16_t) * static_cast<size_t>(ne)));
      const int rc = build_row_dict(ctx, 1, ne, p->p2e_slots, idx, &dict, &n_dict);
      if (rc == LFGPU_OK && n_dict >= 0) {
        cudaError_t e = cudaMalloc(&ce, sizeof(uint32_t) * (3 * static_cast<size_t>(ne) + 128));
        if (e == cudaSuccess) e = cudaMemsetAsync(d_over, 0, sizeof(int), st);
        if (e == cudaSuccess) {
          k_p2_edge_compact<<<static_cast<unsigned>(cdiv(ne, 256)), 256, 0, st>>>(ne, p->p2e_nbr, idx, ce, d_over);
          ctx->launches++;
          e = cudaMemcpyAsync(&over, d_over, sizeof(int), cudaMemcpyDeviceToHost, st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) over = 1;
      }
      cudaFree(idx);
      if (over > 0 && static_cast<int64_t>(over) * 1000 <= ne) {  // a few rows do not fit: they go to the generic kernel
        k_p2_edge_handback<<<static_cast<unsigned>(cdiv(ne, 256)), 256, 0, st>>>(ne, p->p2e_nbr, ce, flag + nn);
        ctx->launches++;
        P2_CHECK(cudaStreamSynchronize(st));
        n_irr += over;
        over = 0;
      }
      if (over == 0) {
        cudaFree(p->p2e_nbr);
        cudaFree(p->p2e_slots);
        p->p2e_nbr = reinterpret_cast<int32_t*>(ce);
        p->p2e_slots = static_cast<uint32_t*>(dict);
        p->p2_compact_e = true;
      } else {
        cudaFree(ce);
        cudaFree(dict);
        (void)cudaGetLastError();
      }
    }
  }
  // rows left to the generic kernel (after the compact edge plan has handed back the few rows it cannot code)
  P2_CHECK(cudaMalloc(&iota, sizeof(int32_t) * p->n_outer));
  P2_CHECK(cudaMalloc(&d_num, sizeof(int64_t)));
  cub::CountingInputIterator<int32_t> count_it(0);
  size_t tb = 0;
  if (and_row_keep(ctx, p->n_outer, flag, p->row_keep) != LFGPU_OK) return LFGPU_ERR_CUDA;  // rows nobody asks for need no generic kernel
  cub::DeviceSelect::Flagged(nullptr, tb, count_it, flag, iota, d_num, p->n_outer, st);
  P2_CHECK(cudaMalloc(&tmp, tb));
  P2_CHECK(cub::DeviceSelect::Flagged(tmp, tb, count_it, flag, iota, d_num, p->n_outer, st));
  n_irr = 0;
  P2_CHECK(cudaMemcpyAsync(&n_irr, d_num, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  P2_CHECK(cudaStreamSynchronize(st));
  if (n_irr > 0) {
    P2_CHECK(cudaMalloc(&p->p2_irregular, sizeof(int32_t) * n_irr));
    P2_CHECK(cudaMemcpyAsync(p->p2_irregular, iota, sizeof(int32_t) * n_irr, cudaMemcpyDeviceToDevice, st));
    p->p2_irregular_host.resize(static_cast<size_t>(n_irr));  // ascending; lets a row range find its share of the list
    P2_CHECK(cudaMemcpyAsync(p->p2_irregular_host.data(), iota, sizeof(int32_t) * n_irr, cudaMemcpyDeviceToHost, st));
    P2_CHECK(cudaStreamSynchronize(st));
  }
#undef P2_CHECK
  cleanup();
  p->n_p2_irregular = n_irr;
  p->p2_nn = nn;
  // a mesh on which most rows are irregular gains nothing: keep the item kernel there
  if (n_irr * 2 > p->n_outer) {
    drop_plan();
    return LFGPU_OK;
  }
  p->p2_state = 1;
  return LFGPU_OK;
}

// tables: the reference tensors of FeLagrangeO2Tria for the rule in use, each [6 * 6] row-major (assemble.cu: pack_type)
// rows [r0, r1) of the matrix (the caller has already sent the irregular rows of the range through the generic kernel)
int p2_rows_launch(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* p, const double alpha[4], int tensor, double gamma,
                   const double* k00, const double* k01, const double* k10, const double* k11, const double* km, double* d_values,
                   int64_t r0, int64_t r1, double beta) {
  if (p->p2_cc != (mesh->cell_coords != nullptr)) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "the P2 plan was built for another kind of mesh geometry");
  P2Params P;
  P.a00 = alpha[0]; P.a01 = tensor ? alpha[1] : 0.0; P.a10 = tensor ? alpha[2] : 0.0; P.a11 = tensor ? alpha[3] : alpha[0];
  P.gamma = gamma;
  const bool simple = !tensor && gamma == 0.0;
  for (int b = 0; b < 6; ++b) {
    P.vk00[b] = k00[b]; P.vk01[b] = simple ? k01[b] + k10[b] : k01[b]; P.vk10[b] = k10[b]; P.vk11[b] = k11[b]; P.vm[b] = km[b];
    P.ek00[b] = k00[18 + b]; P.ek01[b] = simple ? k01[18 + b] + k10[18 + b] : k01[18 + b]; P.ek10[b] = k10[18 + b];
    P.ek11[b] = k11[18 + b]; P.em[b] = km[18 + b];
  }
  const int threads = 128;
  const int nn = static_cast<int>(p->p2_nn), ne = static_cast<int>(p->n_outer - p->p2_nn);
  static const int pfd_env = [] { const char* e = std::getenv("LFGPU_P2_PFD"); return e != nullptr ? std::atoi(e) : 100; }();
  // L2 prefetch distance: about one wave of resident CTAs (6 per SM for the vertex rows, 12 for the edge rows)
  const int ipf_v = static_cast<int>((static_cast<int64_t>(ctx->sm_count) * 6 * threads * pfd_env / 100) & ~static_cast<int64_t>(127));
  const int ipf_e = static_cast<int>((static_cast<int64_t>(ctx->sm_count) * 12 * threads * pfd_env / 100) & ~static_cast<int64_t>(127));
  // LFGPU_EDGE_PFC (percent of the plan distance, default 0 = off): coordinate prefetch of the edge rows through the plan
  static const int pfc_env = [] { const char* e = std::getenv("LFGPU_EDGE_PFC"); return e != nullptr ? std::atoi(e) : 0; }();
  // (with the compact edge plan the value only switches the prefetch of the row-pointer lines on)
  const int ipc_e = pfc_env > 0 && ipf_e > 0 ? std::max(128, static_cast<int>((static_cast<int64_t>(ipf_e) * pfc_env / 100) & ~static_cast<int64_t>(127))) : 0;
  // copy-out of the staged rows by the TMA unit (LFGPU_P2_BULK=0: by the lanes); needs a 16-byte aligned value array
  static const bool bulk_env = [] { const char* e = std::getenv("LFGPU_P2_BULK"); return e == nullptr || e[0] != '0'; }();
  const bool bulk = bulk_env && (reinterpret_cast<uintptr_t>(d_values) & 15) == 0;
  // L2 eviction priorities (LFGPU_L2_HINTS=1; default off): evict-first on the value stream (bulk copy-out only), evict-last on the
  // coordinate gathers.  Measured at config C3: 0.777 -> 0.791 ms, and the coordinate re-reads from DRAM did not go down
  static const bool hint_env = [] { const char* e = std::getenv("LFGPU_L2_HINTS"); return e != nullptr && e[0] == '1'; }();
  const bool hint = hint_env && bulk;
  const size_t smem_v = sizeof(double) * (threads / 32) * 32 * (kVertexRowLen + 1);
  const size_t smem_e = sizeof(double) * (threads / 32) * 32 * (kEdgeRowLen + 1);
  // the share of the range in the vertex rows [0, nn) and in the edge rows [nn, nn + ne)
  const int v_first = static_cast<int>(std::min<int64_t>(std::max<int64_t>(r0, 0), nn)), v_end = static_cast<int>(std::min<int64_t>(r1, nn));
  const int e_first = static_cast<int>(std::max<int64_t>(r0 - nn, 0)), e_end = static_cast<int>(std::min<int64_t>(std::max<int64_t>(r1 - nn, 0), ne));
  if (v_end > v_first && p->p2_general) {
    p2::VertexParams G;
    G.a00 = P.a00; G.a01 = P.a01; G.a10 = P.a10; G.a11 = P.a11; G.gamma = P.gamma;
    for (int b = 0; b < 6; ++b) {
      G.k00[b] = P.vk00[b]; G.k01[b] = P.vk01[b]; G.k10[b] = P.vk10[b]; G.k11[b] = P.vk11[b]; G.km[b] = P.vm[b];
    }
    const size_t smem_g = sizeof(double) * (threads / 32) * 32 * (p2::kMaxVertexRowLen + 1);
    const unsigned gg = static_cast<unsigned>(cdiv(v_end - v_first, threads));
    if (simple)
      k_p2_vertex_rows_general<0><<<gg, threads, smem_g, ctx->stream>>>(v_first, v_end, nn, p->p2g_nbr, p->p2g_slots, mesh->node_coords, p->outer, G, d_values, beta);
    else
      k_p2_vertex_rows_general<1><<<gg, threads, smem_g, ctx->stream>>>(v_first, v_end, nn, p->p2g_nbr, p->p2g_slots, mesh->node_coords, p->outer, G, d_values, beta);
    LFGPU_LAUNCH_CHECK(ctx);
  } else if (v_end > v_first && p->p2_cc) {
    const unsigned gv = static_cast<unsigned>(cdiv(v_end - v_first, threads));
    auto kv = simple ? (bulk ? k_p2_vertex_rows<0, true, false, false, true> : k_p2_vertex_rows<0, false, false, false, true>)
                     : (bulk ? k_p2_vertex_rows<1, true, false, false, true> : k_p2_vertex_rows<1, false, false, false, true>);
    kv<<<gv, threads, smem_v, ctx->stream>>>(nn, p->p2v_nbr, p->p2v_slots, mesh->cell_coords, p->outer, ipf_v, P, d_values, v_first, v_end, beta,
                                             nullptr);
    LFGPU_LAUNCH_CHECK(ctx);
  } else if (v_end > v_first) {
    const unsigned gv = static_cast<unsigned>(cdiv(v_end - v_first, threads));
    const bool cv = p->p2_compact_v;
#define P2_LAUNCH_V(MODE, BULK, COMPACT, HINT)                                                                                         \
  k_p2_vertex_rows<MODE, BULK, COMPACT, HINT><<<gv, threads, smem_v, ctx->stream>>>(nn, p->p2v_nbr, p->p2v_slots, mesh->node_coords, p->outer, \
                                                                                  ipf_v, P, d_values, v_first, v_end, beta, p->p2v_cidx)
#define P2_PICK_V(MODE)                                  \
  do {                                                   \
    if (!bulk) {                                         \
      if (cv) P2_LAUNCH_V(MODE, false, true, false);     \
      else P2_LAUNCH_V(MODE, false, false, false);       \
    } else if (hint) {                                   \
      if (cv) P2_LAUNCH_V(MODE, true, true, true);       \
      else P2_LAUNCH_V(MODE, true, false, true);         \
    } else {                                             \
      if (cv) P2_LAUNCH_V(MODE, true, true, false);      \
      else P2_LAUNCH_V(MODE, true, false, false);        \
    }                                                    \
  } while (0)
    if (simple) P2_PICK_V(0);
    else P2_PICK_V(1);
#undef P2_PICK_V
#undef P2_LAUNCH_V
    LFGPU_LAUNCH_CHECK(ctx);
  }
  if (e_end > e_first && p->p2_cc) {
    const unsigned ge = static_cast<unsigned>(cdiv(e_end - e_first, threads));
    auto ke = simple ? (bulk ? k_p2_edge_rows<0, true, false, false, true> : k_p2_edge_rows<0, false, false, false, true>)
                     : (bulk ? k_p2_edge_rows<1, true, false, false, true> : k_p2_edge_rows<1, false, false, false, true>);
    ke<<<ge, threads, smem_e, ctx->stream>>>(ne, nn, p->p2e_nbr, p->p2e_slots, mesh->cell_coords, p->outer, ipf_e, 0, P, d_values, e_first, e_end, beta);
    LFGPU_LAUNCH_CHECK(ctx);
  } else if (e_end > e_first) {
    const unsigned ge = static_cast<unsigned>(cdiv(e_end - e_first, threads));
    const bool ce = p->p2_compact_e;
    const double* exy = mesh->node_coords;
    if (p->p2e_newid != nullptr) {  // the plan's node numbers are positions in the rows' own coordinate copy: bring it up to date
      auto* pm = const_cast<lfgpu_pattern*>(p);
      if (pm->p2e_xy_version != mesh->coords_version || pm->p2e_xy_mesh != mesh) {
        const int rc = permute_node_coords(ctx, nn, p->p2e_newid, mesh->node_coords, pm->p2e_xy);
        if (rc != LFGPU_OK) return rc;
        pm->p2e_xy_version = mesh->coords_version;
        pm->p2e_xy_mesh = mesh;
      }
      exy = p->p2e_xy;
    }
#define P2_LAUNCH_E(MODE, BULK, COMPACT, HINT)                                                                                         \
  k_p2_edge_rows<MODE, BULK, COMPACT, HINT><<<ge, threads, smem_e, ctx->stream>>>(ne, nn, p->p2e_nbr, p->p2e_slots, exy, p->outer, \
                                                                                ipf_e, ipc_e, P, d_values, e_first, e_end, beta)
#define P2_PICK_E(MODE)                                  \
  do {                                                   \
    if (!bulk) {                                         \
      if (ce) P2_LAUNCH_E(MODE, false, true, false);     \
      else P2_LAUNCH_E(MODE, false, false, false);       \
    } else if (hint) {                                   \
      if (ce) P2_LAUNCH_E(MODE, true, true, true);       \
      else P2_LAUNCH_E(MODE, true, false, true);         \
    } else {                                             \
      if (ce) P2_LAUNCH_E(MODE, true, true, false);      \
      else P2_LAUNCH_E(MODE, true, false, false);        \
    }                                                    \
  } while (0)
    if (simple) P2_PICK_E(0);
    else P2_PICK_E(1);
#undef P2_PICK_E
#undef P2_LAUNCH_E
    LFGPU_LAUNCH_CHECK(ctx);
  }
  return LFGPU_OK;
}

}  // namespace lfgpu
