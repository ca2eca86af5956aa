// P3 (FeLagrangeO3Tria) row kernels -- the part that is plain arithmetic and index logic, written once for host and device
// (product code).  The CUDA kernels of assemble_p3.cu are thin wrappers around these functions; tests/cpp/p3_rows_emul.cc
// compiles the same functions with g++ so that plan construction and row arithmetic can be checked against the oracle
// without a GPU (tests/test_p3_rows_core.py).
//
// Decomposition (same idea as assemble_p2.cu): one thread owns one matrix ROW.
//   * VERTEX row (dof = node i, closed ring of six cells n_0..n_5): cell k is taken as the triangle (i, n_k, n_k+1) with i
//     as local vertex 0 -> row 0 of the reference tensors.  Its ten entries: diagonal | n_k | n_k+1 | the two dofs of the
//     spoke (i, n_k), nearer i first | the two dofs of the rim edge (n_k, n_k+1), nearer n_k first | the two dofs of the
//     spoke (n_k+1, i), nearer n_k+1 first | the cell's interior dof.  37 stored values.
//   * EDGE-DOF row (one of the two interior dofs of an edge with two adjacent cells): each cell is taken as (P, Q, o) with
//     P the endpoint NEARER to the dof -> the dof is local dof 3 (first dof of local edge 0) and row 3 of the reference
//     tensors serves both dofs of every edge.  16 stored values.
//   * CELL row (the interior dof of a cell): row 9 of the element matrix in the cell's own numbering.  10 stored values.
// Legitimate because the cubic Lagrange basis (lagr_fe.h:944-1140: vertices, points at 1/3 and 2/3 of every edge in the
// direction of the local edge, centroid) is invariant under vertex permutations and the provider's default rule (degree
// 6) integrates stiffness and mass products of cubics exactly: same integrals, rounded differently.
// Which global dof sits at which local position -- including the reversal of the two edge dofs for a negative relative
// orientation (dofhandler.cc:245-260) -- is NOT re-derived here: position b of a cell's dof list always carries local
// shape function b, and the scatter map of the symbolic pass holds the slot of list position b in every row of the
// cell; the plan functions below only translate "local position in the re-labelled triangle" into "list position of the
// actual cell".
#ifndef LFGPU_ROWS_P3_CORE_H
#define LFGPU_ROWS_P3_CORE_H

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define LFGPU_HD __host__ __device__ __forceinline__
#else
#define LFGPU_HD inline
#endif

namespace lfgpu {
namespace p3 {

constexpr int kRing = 6;
constexpr int kVertexRowLen = 37;  // 1 + 6 neighbours + 12 spoke dofs + 12 rim dofs + 6 interior dofs
constexpr int kEdgeRowLen = 16;    // P, Q, the edge's two dofs + per cell: o, 4 edge dofs, interior dof
constexpr int kCellRowLen = 10;
constexpr int kVertexSlotWords = 9;  // 36 bytes: [ring position k][n_k, spokeA_k, spokeB_k, rim1_k, rim2_k, interior_k]
constexpr int kEdgeSlotWords = 2;    // 16 nibbles: P, Q, self, sibling | cell 1: o, t5..t9 | cell 2: o, t5..t9

struct Params {
  double a00, a01, a10, a11;  // diffusion tensor as the row routine of assemble.cu uses it (transposed for row-major output)
  double gamma;
  // rows 0 (vertex), 3 (edge dof) and 9 (cell) of the reference tensors of the rule; MODE 0 reads k01 as k01 + k10
  double k00[3][10], k01[3][10], k10[3][10], k11[3][10], km[3][10];
};

// Position of value `k` of a warp's value range inside its shared-memory stage.  SWZ (device kernels): XOR of the low four
// index bits with the next four -- a bijection inside every aligned group of 16 that spreads the same slot of rows of 16
// (or 10) values over all banks (ncu on k_p3_edge_rows: 113 M of 139 M shared-memory wavefronts were bank conflicts, the
// L1 pipe 98 % busy).  The host emulation keeps the identity.
template <bool SWZ>
LFGPU_HD int stage_ix(int k) { return SWZ ? (k ^ ((k >> 4) & 15)) : k; }

LFGPU_HD double rcp(double x) {
#if defined(__CUDA_ARCH__)
  // hardware seed + two Newton steps (Jacobian determinants: no denormals, no zeros), as in assemble.cu
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
#else
  return 1.0 / x;
#endif
}

// Row (0, 3 or 9: WHICH = 0, 1, 2) of the element matrix of the triangle (x0, x0 + A, x0 + B):
// J = [A B], M = |det| J^-1 alpha J^-T, t[b] = sum_ij M_ij Khat^{ji}[row][b] + gamma |det| Mhat[row][b]
// MODE 0: scalar alpha, gamma = 0 (M symmetric: three products per entry)
template <int MODE, int WHICH>
LFGPU_HD void row(const Params& P, double ax, double ay, double bx, double by, double (&t)[10]) {
  const double det = ax * by - ay * bx;
  if (MODE == 0) {
    const double s = P.a00 * rcp(fabs(det));
    const double m00 = s * (bx * bx + by * by), m01 = -s * (ax * bx + ay * by), m11 = s * (ax * ax + ay * ay);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < 10; ++b) t[b] = m00 * P.k00[WHICH][b] + m01 * P.k01[WHICH][b] + m11 * P.k11[WHICH][b];
  } else {
    const double adet = fabs(det), idet = rcp(det);
    const double i00 = by * idet, i01 = -bx * idet, i10 = -ay * idet, i11 = ax * idet;
    const double t00 = i00 * P.a00 + i01 * P.a10, t01 = i00 * P.a01 + i01 * P.a11;
    const double t10 = i10 * P.a00 + i11 * P.a10, t11 = i10 * P.a01 + i11 * P.a11;
    const double m00 = adet * (t00 * i00 + t01 * i01), m01 = adet * (t00 * i10 + t01 * i11);
    const double m10 = adet * (t10 * i00 + t11 * i01), m11 = adet * (t10 * i10 + t11 * i11);
    const double gm = adet * P.gamma;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < 10; ++b)
      t[b] = m00 * P.k00[WHICH][b] + m01 * P.k10[WHICH][b] + m10 * P.k01[WHICH][b] + m11 * P.k11[WHICH][b] + gm * P.km[WHICH][b];
  }
}

LFGPU_HD int byte_at(const uint32_t* w, int i) { return static_cast<int>((w[i >> 2] >> (8 * (i & 3))) & 255U); }
LFGPU_HD int nibble_at(const uint32_t* w, int i) { return static_cast<int>((w[i >> 3] >> (4 * (i & 7))) & 15U); }

// ---- plans ------------------------------------------------------------------------------------------------------------
// items: the row's (cell << 4 | list position) entries, ascending in cell; pos: scatter map, pos[(cell * o_stride + a) *
// pos_row + b] = slot of list position b in the row of list position a; row_len: stored values of the row.

// Vertex row r: false unless exactly six cells close around the node and the 37 slots are a permutation of 0..36.
// cellw (optional, six entries): for meshes whose cells carry their own corner coordinates (lfgpu_mesh_upload: cell_coords) the
// kernels read the corners of ring cell k instead of node positions: cellw[k] = cell << 4 | (local index of n_k) << 2 | local
// index of n_k+1 (the node itself is the third corner).
LFGPU_HD bool vertex_plan(int64_t r, int m, const uint32_t* items, const uint32_t* cell_nodes, const uint8_t* pos, int o_stride,
                          int pos_row, int row_len, int32_t (&ring)[kRing], uint32_t (&words)[kVertexSlotWords],
                          uint32_t* cellw = nullptr) {
  if (m != kRing || row_len != kVertexRowLen) return false;
  uint32_t ja[kRing], ka[kRing];
  int la[kRing];
  int64_t cid[kRing];
  for (int t = 0; t < kRing; ++t) {
    cid[t] = items[t] >> 4;
    la[t] = static_cast<int>(items[t] & 15U);
    if (la[t] > 2) return false;
    const uint32_t* v = cell_nodes + 4 * cid[t];
    ja[t] = v[(la[t] + 1) % 3];
    ka[t] = v[(la[t] + 2) % 3];
  }
  int ord[kRing];
  bool fwd[kRing];
  uint32_t rg[kRing];
  unsigned used = 1U;
  rg[0] = ja[0];
  ord[0] = 0;
  fwd[0] = true;
  uint32_t cur = ka[0];
  for (int k = 1; k < kRing; ++k) {
    int nxt = -1;
    for (int u = 0; u < kRing; ++u) {
      if (!(used & (1U << u)) && (ja[u] == cur || ka[u] == cur)) {
        nxt = u;
        break;
      }
    }
    if (nxt < 0) return false;
    used |= 1U << nxt;
    rg[k] = cur;
    ord[k] = nxt;
    fwd[k] = (ja[nxt] == cur);
    cur = fwd[k] ? ka[nxt] : ja[nxt];
  }
  if (cur != rg[0]) return false;
  for (int k = 0; k < kRing; ++k) {
    if (rg[k] == static_cast<uint32_t>(r)) return false;
    for (int u = 0; u < k; ++u)
      if (rg[u] == rg[k]) return false;
  }
  uint8_t s[36];
  int sd = -1;
  for (int k = 0; k < kRing; ++k) {
    const int u = ord[k];
    const int a = la[u], vb = (a + 1) % 3, vc = (a + 2) % 3;
    const uint8_t* prow = pos + (cid[u] * o_stride + a) * static_cast<int64_t>(pos_row);
    if (sd >= 0 && prow[a] != sd) return false;
    sd = prow[a];
    if (fwd[k]) {
      // actual cell = (i, n_k, n_k+1) up to rotation: spoke k is the local edge a (i -> n_k), the rim the local edge vb
      s[6 * k + 0] = prow[vb];
      s[6 * k + 1] = prow[3 + 2 * a];      // spoke dof nearer i
      s[6 * k + 2] = prow[3 + 2 * a + 1];  // spoke dof nearer n_k
      s[6 * k + 3] = prow[3 + 2 * vb];     // rim dof nearer n_k
      s[6 * k + 4] = prow[3 + 2 * vb + 1];
    } else {
      // actual cell = (i, n_k+1, n_k) up to rotation: spoke k is the local edge vc (n_k -> i), the rim runs n_k+1 -> n_k
      s[6 * k + 0] = prow[vc];
      s[6 * k + 1] = prow[3 + 2 * vc + 1];
      s[6 * k + 2] = prow[3 + 2 * vc];
      s[6 * k + 3] = prow[3 + 2 * vb + 1];
      s[6 * k + 4] = prow[3 + 2 * vb];
    }
    s[6 * k + 5] = prow[9];
  }
  uint64_t seen = 1ULL << sd;
  int sum = 0;
  for (int j = 0; j < 36; ++j) {
    if (s[j] >= kVertexRowLen) return false;
    seen |= 1ULL << s[j];
    sum += s[j];
  }
  if (seen != (1ULL << kVertexRowLen) - 1ULL || sd != 666 - sum) return false;  // the kernel recovers the diagonal slot as 666 - sum
  for (int k = 0; k < kRing; ++k) ring[k] = static_cast<int32_t>(rg[k]);
  for (int j = 0; j < kVertexSlotWords; ++j)
    words[j] = static_cast<uint32_t>(s[4 * j]) | (static_cast<uint32_t>(s[4 * j + 1]) << 8) | (static_cast<uint32_t>(s[4 * j + 2]) << 16) |
               (static_cast<uint32_t>(s[4 * j + 3]) << 24);
  if (cellw != nullptr) {
    for (int k = 0; k < kRing; ++k) {
      const int u = ord[k];
      const int a = la[u], vb = (a + 1) % 3, vc = (a + 2) % 3;
      cellw[k] = (static_cast<uint32_t>(cid[u]) << 4) | static_cast<uint32_t>((fwd[k] ? vb : vc) << 2) | static_cast<uint32_t>(fwd[k] ? vc : vb);
    }
  }
  return true;
}

// ---- vertex rows for ANY closed ring of 3..8 cells (unstructured meshes) ------------------------------------------------
// Same decomposition; the loop runs over eight ring positions under `k < m` predicates so that register indices stay
// static.  Plan: ring[8] (-1 beyond the ring), 48 slot bytes [k][n_k, spokeA_k, spokeB_k, rim1_k, rim2_k, interior_k] in 12
// words and the ring length m in a 13th.
constexpr int kMaxRing = 8;
constexpr int kMaxVertexRowLen = 1 + 6 * kMaxRing;  // 49
constexpr int kGeneralSlotWords = 13;               // 48 slot bytes + the ring length

LFGPU_HD bool vertex_plan_general(int64_t r, int m, const uint32_t* items, const uint32_t* cell_nodes, const uint8_t* pos, int o_stride,
                                  int pos_row, int row_len, int32_t (&ring)[kMaxRing], uint32_t (&words)[kGeneralSlotWords]) {
  if (m < 3 || m > kMaxRing || row_len != 1 + 6 * m) return false;
  uint32_t ja[kMaxRing], ka[kMaxRing], rg[kMaxRing];
  int la[kMaxRing], ord[kMaxRing];
  int64_t cid[kMaxRing];
  bool fwd[kMaxRing];
  for (int t = 0; t < m; ++t) {
    cid[t] = items[t] >> 4;
    la[t] = static_cast<int>(items[t] & 15U);
    if (la[t] > 2) return false;
    const uint32_t* v = cell_nodes + 4 * cid[t];
    if (v[3] != 0xFFFFFFFFu) return false;  // a quadrilateral
    ja[t] = v[(la[t] + 1) % 3];
    ka[t] = v[(la[t] + 2) % 3];
  }
  unsigned used = 1U;
  rg[0] = ja[0];
  ord[0] = 0;
  fwd[0] = true;
  uint32_t cur = ka[0];
  for (int k = 1; k < m; ++k) {
    int nxt = -1;
    for (int u = 0; u < m; ++u) {
      if (!(used & (1U << u)) && (ja[u] == cur || ka[u] == cur)) {
        nxt = u;
        break;
      }
    }
    if (nxt < 0) return false;
    used |= 1U << nxt;
    rg[k] = cur;
    ord[k] = nxt;
    fwd[k] = (ja[nxt] == cur);
    cur = fwd[k] ? ka[nxt] : ja[nxt];
  }
  if (cur != rg[0]) return false;
  for (int k = 0; k < m; ++k) {
    if (rg[k] == static_cast<uint32_t>(r)) return false;
    for (int u = 0; u < k; ++u)
      if (rg[u] == rg[k]) return false;
  }
  uint8_t s[48];
  for (int j = 0; j < 48; ++j) s[j] = 0;
  int sd = -1;
  for (int k = 0; k < m; ++k) {
    const int u = ord[k];
    const int a = la[u], vb = (a + 1) % 3, vc = (a + 2) % 3;
    const uint8_t* prow = pos + (cid[u] * o_stride + a) * static_cast<int64_t>(pos_row);
    if (sd >= 0 && prow[a] != sd) return false;
    sd = prow[a];
    if (fwd[k]) {
      s[6 * k + 0] = prow[vb];
      s[6 * k + 1] = prow[3 + 2 * a];
      s[6 * k + 2] = prow[3 + 2 * a + 1];
      s[6 * k + 3] = prow[3 + 2 * vb];
      s[6 * k + 4] = prow[3 + 2 * vb + 1];
    } else {
      s[6 * k + 0] = prow[vc];
      s[6 * k + 1] = prow[3 + 2 * vc + 1];
      s[6 * k + 2] = prow[3 + 2 * vc];
      s[6 * k + 3] = prow[3 + 2 * vb + 1];
      s[6 * k + 4] = prow[3 + 2 * vb];
    }
    s[6 * k + 5] = prow[9];
  }
  uint64_t seen = 1ULL << sd;
  int sum = 0;
  for (int j = 0; j < 6 * m; ++j) {
    if (s[j] >= row_len) return false;
    seen |= 1ULL << s[j];
    sum += s[j];
  }
  if (seen != (1ULL << row_len) - 1ULL || sd != row_len * (row_len - 1) / 2 - sum) return false;
  for (int k = 0; k < kMaxRing; ++k) ring[k] = k < m ? static_cast<int32_t>(rg[k]) : -1;
  for (int j = 0; j < 12; ++j)
    words[j] = static_cast<uint32_t>(s[4 * j]) | (static_cast<uint32_t>(s[4 * j + 1]) << 8) | (static_cast<uint32_t>(s[4 * j + 2]) << 16) |
               (static_cast<uint32_t>(s[4 * j + 3]) << 24);
  words[12] = static_cast<uint32_t>(m);
  return true;
}

// Edge-dof row: false unless exactly two cells share the edge and the 16 slots are a permutation of 0..15.
// ids = P (endpoint nearer to the dof), Q, o_1, o_2.
// cellw (optional, two entries): cell << 4 | (local index of Q) << 2 | local index of o -- P, the origin of the cell's frame, is
// the third corner (as the node itself is for vertex_plan).
LFGPU_HD bool edge_plan(int m, const uint32_t* items, const uint32_t* cell_nodes, const uint8_t* pos, int o_stride, int pos_row, int row_len,
                        int32_t (&ids)[4], uint32_t (&words)[kEdgeSlotWords], uint32_t* cellw = nullptr) {
  if (m != 2 || row_len != kEdgeRowLen) return false;
  uint8_t s[16];
  uint32_t cw[2] = {0U, 0U};
  uint32_t P = 0, Q = 0, o[2] = {0, 0};
  for (int c = 0; c < 2; ++c) {
    const int64_t cell = items[c] >> 4;
    const int a = static_cast<int>(items[c] & 15U);
    if (a < 3 || a > 8) return false;
    const int j = (a - 3) >> 1, w = (a - 3) & 1;
    const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    const uint32_t* v = cell_nodes + 4 * cell;
    const uint8_t* prow = pos + (cell * o_stride + a) * static_cast<int64_t>(pos_row);
    const uint32_t p = w == 0 ? v[j] : v[j1], q = w == 0 ? v[j1] : v[j];
    if (c == 0) {
      P = p;
      Q = q;
      s[0] = prow[w == 0 ? j : j1];
      s[1] = prow[w == 0 ? j1 : j];
      s[2] = prow[a];
      s[3] = prow[w == 0 ? a + 1 : a - 1];  // the edge's other dof
    } else {
      if (p != P || q != Q) return false;
      if (prow[a] != s[2] || prow[w == 0 ? a + 1 : a - 1] != s[3]) return false;
    }
    o[c] = v[j2];
    cw[c] = (static_cast<uint32_t>(cell) << 4) | static_cast<uint32_t>((w == 0 ? j1 : j) << 2) | static_cast<uint32_t>(j2);
    uint8_t* sc = s + 4 + 6 * c;
    sc[0] = prow[j2];
    if (w == 0) {
      // (P, Q, o) = (v_j, v_j+1, v_j+2): a rotation of the cell; edge (Q, o) = local edge j+1, edge (o, P) = local edge j+2
      sc[1] = prow[3 + 2 * j1];      // t5: dof of (Q, o) nearer Q
      sc[2] = prow[3 + 2 * j1 + 1];  // t6
      sc[3] = prow[3 + 2 * j2];      // t7: dof of (o, P) nearer o
      sc[4] = prow[3 + 2 * j2 + 1];  // t8
    } else {
      // (P, Q, o) = (v_j+1, v_j, v_j+2): a reflection; edge (Q, o) = local edge j+2 run backwards, (o, P) = local edge j+1 backwards
      sc[1] = prow[3 + 2 * j2 + 1];
      sc[2] = prow[3 + 2 * j2];
      sc[3] = prow[3 + 2 * j1 + 1];
      sc[4] = prow[3 + 2 * j1];
    }
    sc[5] = prow[9];
  }
  if (o[0] == o[1]) return false;
  unsigned seen = 0;
  for (int k = 0; k < 16; ++k) {
    if (s[k] >= kEdgeRowLen) return false;
    seen |= 1U << s[k];
  }
  if (seen != 0xFFFFU) return false;
  ids[0] = static_cast<int32_t>(P);
  ids[1] = static_cast<int32_t>(Q);
  ids[2] = static_cast<int32_t>(o[0]);
  ids[3] = static_cast<int32_t>(o[1]);
  words[0] = words[1] = 0U;
  for (int k = 0; k < 16; ++k) words[k >> 3] |= static_cast<uint32_t>(s[k]) << (4 * (k & 7));
  if (cellw != nullptr) {
    cellw[0] = cw[0];
    cellw[1] = cw[1];
  }
  return true;
}

// ---- rows -------------------------------------------------------------------------------------------------------------
// dx, dy: ring coordinates relative to the node; w: the 36 slot bytes; dst: the row's 37 values (any order of writes)
template <int MODE>
LFGPU_HD void vertex_row(const Params& P, const double (&dx)[kRing], const double (&dy)[kRing], const uint32_t (&w)[kVertexSlotWords],
                         double* dst) {
  int ssum = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int j = 0; j < kVertexSlotWords; ++j) ssum += static_cast<int>(__vsadu4(w[j], 0U));
#else
  for (int j = 0; j < 36; ++j) ssum += byte_at(w, j);
#endif
  double t[10];
  row<MODE, 0>(P, dx[0], dy[0], dx[1], dy[1], t);
  double diag = t[0];
  const double first_n = t[1], first_a = t[3], first_b = t[4];
  double carry_n = t[2], carry_b = t[7], carry_a = t[8];
  dst[byte_at(w, 3)] = t[5];
  dst[byte_at(w, 4)] = t[6];
  dst[byte_at(w, 5)] = t[9];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int s = 1; s < kRing; ++s) {
    const int u = (s + 1 < kRing) ? s + 1 : 0;
    row<MODE, 0>(P, dx[s], dy[s], dx[u], dy[u], t);
    diag += t[0];
    dst[byte_at(w, 6 * s + 0)] = carry_n + t[1];
    dst[byte_at(w, 6 * s + 1)] = carry_a + t[3];
    dst[byte_at(w, 6 * s + 2)] = carry_b + t[4];
    dst[byte_at(w, 6 * s + 3)] = t[5];
    dst[byte_at(w, 6 * s + 4)] = t[6];
    dst[byte_at(w, 6 * s + 5)] = t[9];
    carry_n = t[2];
    carry_b = t[7];
    carry_a = t[8];
  }
  dst[byte_at(w, 0)] = first_n + carry_n;
  dst[byte_at(w, 1)] = first_a + carry_a;
  dst[byte_at(w, 2)] = first_b + carry_b;
  dst[666 - ssum] = diag;
}

// The same row with the edge vectors of every ring cell supplied by the caller: cv(k, ax, ay, bx, by) returns A = n_k - i and
// B = n_k+1 - i as cell k sees them (meshes with per-cell corner coordinates: the corners of two cells at one node may differ in
// the last bit, and the reference computes every element matrix from the cell's own geometry object, tria_o1.cc:50-74).
// (A separate function: vertex_row above is the measured kernel body and stays as it is.)
template <int MODE, class CELLVEC>
LFGPU_HD void vertex_row_cv(const Params& P, const CELLVEC& cv, const uint32_t (&w)[kVertexSlotWords], double* dst) {
  int ssum = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int j = 0; j < kVertexSlotWords; ++j) ssum += static_cast<int>(__vsadu4(w[j], 0U));
#else
  for (int j = 0; j < 36; ++j) ssum += byte_at(w, j);
#endif
  double t[10];
  double ax, ay, bx, by;
  cv(0, ax, ay, bx, by);
  row<MODE, 0>(P, ax, ay, bx, by, t);
  double diag = t[0];
  const double first_n = t[1], first_a = t[3], first_b = t[4];
  double carry_n = t[2], carry_b = t[7], carry_a = t[8];
  dst[byte_at(w, 3)] = t[5];
  dst[byte_at(w, 4)] = t[6];
  dst[byte_at(w, 5)] = t[9];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int s = 1; s < kRing; ++s) {
    cv(s, ax, ay, bx, by);
    row<MODE, 0>(P, ax, ay, bx, by, t);
    diag += t[0];
    dst[byte_at(w, 6 * s + 0)] = carry_n + t[1];
    dst[byte_at(w, 6 * s + 1)] = carry_a + t[3];
    dst[byte_at(w, 6 * s + 2)] = carry_b + t[4];
    dst[byte_at(w, 6 * s + 3)] = t[5];
    dst[byte_at(w, 6 * s + 4)] = t[6];
    dst[byte_at(w, 6 * s + 5)] = t[9];
    carry_n = t[2];
    carry_b = t[7];
    carry_a = t[8];
  }
  dst[byte_at(w, 0)] = first_n + carry_n;
  dst[byte_at(w, 1)] = first_a + carry_a;
  dst[byte_at(w, 2)] = first_b + carry_b;
  dst[666 - ssum] = diag;
}

// general ring: dx, dy entries k >= m unused; dst: the row's 1 + 6m values
template <int MODE>
LFGPU_HD void vertex_row_general(const Params& P, const double (&dx)[kMaxRing], const double (&dy)[kMaxRing],
                                 const uint32_t (&w)[kGeneralSlotWords], double* dst) {
  const int m = static_cast<int>(w[12]);
  const int len = 1 + 6 * m;
  int ssum = 0;
  double t[10];
  row<MODE, 0>(P, dx[0], dy[0], dx[1], dy[1], t);  // m >= 3: cell 0 is (i, n_0, n_1)
  double diag = t[0];
  const double first_n = t[1], first_a = t[3], first_b = t[4];
  double carry_n = t[2], carry_b = t[7], carry_a = t[8];
  dst[byte_at(w, 3)] = t[5];
  dst[byte_at(w, 4)] = t[6];
  dst[byte_at(w, 5)] = t[9];
  ssum += byte_at(w, 0) + byte_at(w, 1) + byte_at(w, 2) + byte_at(w, 3) + byte_at(w, 4) + byte_at(w, 5);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int s = 1; s < kMaxRing; ++s) {
    if (s < m) {
      const bool last = (s + 1 == m);  // the last cell closes the ring with n_0
      const double nx = last ? dx[0] : dx[(s + 1) & (kMaxRing - 1)], ny = last ? dy[0] : dy[(s + 1) & (kMaxRing - 1)];
      row<MODE, 0>(P, dx[s], dy[s], nx, ny, t);
      diag += t[0];
      dst[byte_at(w, 6 * s + 0)] = carry_n + t[1];
      dst[byte_at(w, 6 * s + 1)] = carry_a + t[3];
      dst[byte_at(w, 6 * s + 2)] = carry_b + t[4];
      dst[byte_at(w, 6 * s + 3)] = t[5];
      dst[byte_at(w, 6 * s + 4)] = t[6];
      dst[byte_at(w, 6 * s + 5)] = t[9];
      ssum += byte_at(w, 6 * s) + byte_at(w, 6 * s + 1) + byte_at(w, 6 * s + 2) + byte_at(w, 6 * s + 3) + byte_at(w, 6 * s + 4) +
              byte_at(w, 6 * s + 5);
      carry_n = t[2];
      carry_b = t[7];
      carry_a = t[8];
    }
  }
  dst[byte_at(w, 0)] = first_n + carry_n;
  dst[byte_at(w, 1)] = first_a + carry_a;
  dst[byte_at(w, 2)] = first_b + carry_b;
  dst[len * (len - 1) / 2 - ssum] = diag;
}

// (ax, ay) = Q - P, (b1x, b1y) = o_1 - P, (b2x, b2y) = o_2 - P; value k of the row goes to dst[stage_ix<SWZ>(off + k)]
template <int MODE, bool SWZ = false>
LFGPU_HD void edge_row(const Params& P, double ax, double ay, double b1x, double b1y, double b2x, double b2y,
                       const uint32_t (&w)[kEdgeSlotWords], double* dst, int off = 0) {
  double t1[10], t2[10];
  row<MODE, 1>(P, ax, ay, b1x, b1y, t1);
  row<MODE, 1>(P, ax, ay, b2x, b2y, t2);
  dst[stage_ix<SWZ>(off + nibble_at(w, 0))] = t1[0] + t2[0];
  dst[stage_ix<SWZ>(off + nibble_at(w, 1))] = t1[1] + t2[1];
  dst[stage_ix<SWZ>(off + nibble_at(w, 2))] = t1[3] + t2[3];
  dst[stage_ix<SWZ>(off + nibble_at(w, 3))] = t1[4] + t2[4];
  dst[stage_ix<SWZ>(off + nibble_at(w, 4))] = t1[2];
  dst[stage_ix<SWZ>(off + nibble_at(w, 10))] = t2[2];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 5; ++k) {
    dst[stage_ix<SWZ>(off + nibble_at(w, 5 + k))] = t1[5 + k];
    dst[stage_ix<SWZ>(off + nibble_at(w, 11 + k))] = t2[5 + k];
  }
}

// The same row with each cell's own edge vectors: (a1, b1) = (Q - P, o_1 - P) from the corners of cell 1, (a2, b2) from cell 2
template <int MODE, bool SWZ = false>
LFGPU_HD void edge_row2(const Params& P, double a1x, double a1y, double b1x, double b1y, double a2x, double a2y, double b2x, double b2y,
                        const uint32_t (&w)[kEdgeSlotWords], double* dst, int off = 0) {
  double t1[10], t2[10];
  row<MODE, 1>(P, a1x, a1y, b1x, b1y, t1);
  row<MODE, 1>(P, a2x, a2y, b2x, b2y, t2);
  dst[stage_ix<SWZ>(off + nibble_at(w, 0))] = t1[0] + t2[0];
  dst[stage_ix<SWZ>(off + nibble_at(w, 1))] = t1[1] + t2[1];
  dst[stage_ix<SWZ>(off + nibble_at(w, 2))] = t1[3] + t2[3];
  dst[stage_ix<SWZ>(off + nibble_at(w, 3))] = t1[4] + t2[4];
  dst[stage_ix<SWZ>(off + nibble_at(w, 4))] = t1[2];
  dst[stage_ix<SWZ>(off + nibble_at(w, 10))] = t2[2];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 5; ++k) {
    dst[stage_ix<SWZ>(off + nibble_at(w, 5 + k))] = t1[5 + k];
    dst[stage_ix<SWZ>(off + nibble_at(w, 11 + k))] = t2[5 + k];
  }
}

// the cell in its own numbering: (ax, ay) = v1 - v0, (bx, by) = v2 - v0; pw: the slots of the ten list positions in the row of
// list position 9, one byte each (NIB = false: the first bytes of the scatter-map row) or one nibble each (NIB = true: the
// compact plan word pair of the device kernel)
template <int MODE, bool SWZ = false, bool NIB = false>
LFGPU_HD void cell_row(const Params& P, double ax, double ay, double bx, double by, const uint32_t (&pw)[3], double* dst, int off = 0) {
  double t[10];
  row<MODE, 2>(P, ax, ay, bx, by, t);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int b = 0; b < 10; ++b) {
    const int slot = NIB ? static_cast<int>((pw[b >> 3] >> (4 * (b & 7))) & 15U) : byte_at(pw, b);
    dst[stage_ix<SWZ>(off + slot)] = t[b];
  }
}

}  // namespace p3
}  // namespace lfgpu
#endif
