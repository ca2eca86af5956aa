// P2 (FeLagrangeO2Tria) vertex rows for ANY closed ring of 3..8 cells -- arithmetic and index logic written once for host
// and device (product code).  assemble_p2.cu's vertex kernel is specialised for the valence-6 rings of structured meshes
// (everything static, 80 registers); on unstructured meshes (Gmsh input: valences 4..8) most vertex rows would fall back
// to the generic gather kernel.  The functions below cover those rings with the same decomposition: cell k
// is the triangle (i, n_k, n_k+1 mod m) taken with i as local vertex 0, row 0 of the reference tensors, neighbour and
// spoke-edge columns summed over two consecutive cells in rolling registers, rim-edge columns from one cell; the loop runs
// over the 8 ring positions with `k < m` predicates, so all register indices stay static.
// tests/cpp/p2_rows_emul.cc compiles this file with g++ and tests/test_p2_rows_core.py compares the rows with the oracle
// on a Gmsh mesh without a GPU; the CUDA wrapper is k_p2_vertex_rows_general in assemble_p2.cu.
#ifndef LFGPU_ROWS_P2_CORE_H
#define LFGPU_ROWS_P2_CORE_H

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define LFGPU_P2_HD __host__ __device__ __forceinline__
#else
#define LFGPU_P2_HD inline
#endif

namespace lfgpu {
namespace p2 {

constexpr int kMaxRing = 8;
constexpr int kMaxVertexRowLen = 1 + 3 * kMaxRing;  // 25
constexpr int kSlotWords = 6;  // neighbour slots in words 0-1, spoke slots in 2-3, rim slots in 4-5: 6 x 5 bits in the even
                               // word, 2 x 5 bits in the odd one; bits 10..13 of word 1 hold the ring length m

struct VertexParams {
  double a00, a01, a10, a11;  // diffusion tensor as the row routine of assemble.cu uses it (transposed for row-major output)
  double gamma;
  double k00[6], k01[6], k10[6], k11[6], km[6];  // row 0 of the reference tensors; MODE 0 reads k01 as k01 + k10
};

LFGPU_P2_HD double rcp(double x) {
#if defined(__CUDA_ARCH__)
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

// row 0 of the element matrix of the triangle (x0, x0 + A, x0 + B), as p2_row in assemble_p2.cu
template <int MODE>
LFGPU_P2_HD void row0(const VertexParams& P, double ax, double ay, double bx, double by, double (&t)[6]) {
  const double det = ax * by - ay * bx;
  if (MODE == 0) {
    const double s = P.a00 * rcp(fabs(det));
    const double m00 = s * (bx * bx + by * by), m01 = -s * (ax * bx + ay * by), m11 = s * (ax * ax + ay * ay);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < 6; ++b) t[b] = m00 * P.k00[b] + m01 * P.k01[b] + m11 * P.k11[b];
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
    for (int b = 0; b < 6; ++b) t[b] = m00 * P.k00[b] + m01 * P.k10[b] + m10 * P.k01[b] + m11 * P.k11[b] + gm * P.km[b];
  }
}

// slot of ring position k in the word pair (w[2 * kind], w[2 * kind + 1]); k is a compile-time constant after unrolling
LFGPU_P2_HD int slot_at(const uint32_t* w, int kind, int k) {
  return k < 6 ? static_cast<int>((w[2 * kind] >> (5 * k)) & 31U) : static_cast<int>((w[2 * kind + 1] >> (5 * (k - 6))) & 31U);
}
LFGPU_P2_HD int ring_length(const uint32_t* w) { return static_cast<int>((w[1] >> 10) & 15U); }

// Vertex row r: false unless 3..8 cells close around the node in one fan and the 1 + 3m slots are a permutation of
// 0 .. 3m.  items: the row's (cell << 4 | list position) entries; pos: the scatter map of the symbolic pass.
LFGPU_P2_HD bool vertex_plan_general(int64_t r, int m, const uint32_t* items, const uint32_t* cell_nodes, const uint8_t* pos, int o_stride,
                                     int pos_row, int row_len, int32_t (&ring)[kMaxRing], uint32_t (&words)[kSlotWords]) {
  if (m < 3 || m > kMaxRing || row_len != 1 + 3 * m) return false;
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
  for (int j = 0; j < kSlotWords; ++j) words[j] = 0U;
  int sd = -1, sum = 0;
  uint32_t seen = 0U;
  for (int k = 0; k < m; ++k) {
    const int u = ord[k];
    const int a = la[u], vb = (a + 1) % 3, vc = (a + 2) % 3;
    const uint8_t* prow = pos + (cid[u] * o_stride + a) * static_cast<int64_t>(pos_row);
    const int s[3] = {prow[fwd[k] ? vb : vc],       // column n_k
                      prow[3 + (fwd[k] ? a : vc)],  // spoke edge (i, n_k)
                      prow[3 + vb]};                // rim edge (n_k, n_k+1)
    if (sd >= 0 && prow[a] != sd) return false;
    sd = prow[a];
    for (int kind = 0; kind < 3; ++kind) {
      if (s[kind] >= row_len) return false;
      seen |= 1U << s[kind];
      sum += s[kind];
      if (k < 6) words[2 * kind] |= static_cast<uint32_t>(s[kind]) << (5 * k);
      else words[2 * kind + 1] |= static_cast<uint32_t>(s[kind]) << (5 * (k - 6));
    }
  }
  seen |= 1U << sd;
  if (seen != (1U << row_len) - 1U || sd != row_len * (row_len - 1) / 2 - sum) return false;
  words[1] |= static_cast<uint32_t>(m) << 10;
  for (int k = 0; k < kMaxRing; ++k) ring[k] = k < m ? static_cast<int32_t>(rg[k]) : -1;
  return true;
}

// dx, dy: ring coordinates relative to the node (entries k >= m unused); dst: the row's 1 + 3m values
template <int MODE>
LFGPU_P2_HD void vertex_row_general(const VertexParams& P, const double (&dx)[kMaxRing], const double (&dy)[kMaxRing],
                                    const uint32_t (&w)[kSlotWords], double* dst) {
  const int m = ring_length(w);
  const int len = 1 + 3 * m;
  int ssum = 0;
  double t[6];
  row0<MODE>(P, dx[0], dy[0], dx[1], dy[1], t);  // m >= 3: cell 0 is (i, n_0, n_1)
  double diag = t[0];
  const double first_n = t[1], first_s = t[3];
  double carry_n = t[2], carry_s = t[5];
  dst[slot_at(w, 2, 0)] = t[4];
  ssum += slot_at(w, 0, 0) + slot_at(w, 1, 0) + slot_at(w, 2, 0);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int s = 1; s < kMaxRing; ++s) {
    if (s < m) {
      // cell s is (i, n_s, n_s+1), the last one closes the ring with n_0
      const bool last = (s + 1 == m);
      const double nx = last ? dx[0] : dx[(s + 1) & (kMaxRing - 1)], ny = last ? dy[0] : dy[(s + 1) & (kMaxRing - 1)];
      row0<MODE>(P, dx[s], dy[s], nx, ny, t);
      diag += t[0];
      dst[slot_at(w, 0, s)] = carry_n + t[1];
      dst[slot_at(w, 1, s)] = carry_s + t[3];
      dst[slot_at(w, 2, s)] = t[4];
      ssum += slot_at(w, 0, s) + slot_at(w, 1, s) + slot_at(w, 2, s);
      carry_n = t[2];
      carry_s = t[5];
    }
  }
  dst[slot_at(w, 0, 0)] = first_n + carry_n;
  dst[slot_at(w, 1, 0)] = first_s + carry_s;
  dst[len * (len - 1) / 2 - ssum] = diag;
}

}  // namespace p2
}  // namespace lfgpu
#endif
