// P1 (FeLagrangeO1Tria) fast path of the numeric pass: the "vertex fan" kernel (product code).
//
// Same mathematics as the generic path (uscalfe/loc_comp_ellbvp.h:266-339 with FeLagrangeO1Tria, lagr_fe.h:110-128,
// and TriaO1, geometry/tria_o1.cc:50-74), same output (the values of the compressed matrix of the symbolic pass), but
// organised around the only data an affine P1 row really needs: for matrix row i (= mesh node i) the ring of its
// neighbour nodes n_0, n_1, ... in fan order.  Consecutive neighbours (n_t, n_t+1) span one adjacent triangle, so
//   * the ring IS the connectivity (4 bytes per adjacent cell, instead of cell id + 3 vertex ids + scatter slots),
//   * every off-diagonal entry (i, n_t) is the sum of exactly two consecutive cells -> a rolling register, no
//     accumulator array, no atomics,
//   * the slot of column n_t inside row i is its rank among the ring's ids (+1 if i < n_t): integer compares in
//     registers instead of a stored scatter map.
// HBM traffic per cell: ring 12 B + vertex coordinates 8 B + row pointer 2 B + values 28 B = 50 B (algorithmic: 48 B).
// Each warp stages its 32 consecutive rows in shared memory and writes the values as full 128-byte lines.
//
// Rows whose cells do not form a single fan (non-manifold vertices) or with more cells than the ring holds are listed
// as "irregular" and go through the generic gather kernel (assemble.cu) first; this kernel then leaves them untouched.
#include <algorithm>
#include <cstdlib>

#include <cub/cub.cuh>

#include "lfgpu_internal.cuh"

namespace lfgpu {
namespace {

constexpr int kMaxFan = 12;  // longest ring the builder considers
constexpr uint32_t kNil = 0xFFFFFFFFu;

// ---- plan construction --------------------------------------------------------------------------------------------
// One thread per row: order the adjacent cells into a fan.  Returns the ring in `ring` (ids), its length, closed flag;
// false if the cells do not form exactly one fan.
// cell_item (optional): the adjacency item (cell << 4 | local index of the row's node) of the cell between ring[s] and ring[s + 1]
__device__ bool build_ring(int32_t row, int m, const uint32_t* __restrict__ adj, int64_t it0, const uint32_t* __restrict__ cell_nodes,
                           uint32_t* ring, int& len, bool& closed, uint32_t* cell_item = nullptr) {
  uint32_t ja[kMaxFan], ka[kMaxFan];
  for (int t = 0; t < m; ++t) {
    const uint32_t item = adj[it0 + t];
    const int64_t cell = item >> 4;
    const int a = static_cast<int>(item & 15U);
    const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[cell];
    const uint32_t vv[3] = {v.x, v.y, v.z};
    ja[t] = vv[(a + 1) % 3];
    ka[t] = vv[(a + 2) % 3];
  }
  // endpoints of an open fan: ids that occur exactly once
  int start = -1;
  bool start_is_j = true;
  int n_single = 0;
  for (int t = 0; t < m; ++t) {
    for (int e = 0; e < 2; ++e) {
      const uint32_t id = e == 0 ? ja[t] : ka[t];
      int cnt = 0;
      for (int u = 0; u < m; ++u) cnt += (ja[u] == id) + (ka[u] == id);
      if (cnt == 1) {
        ++n_single;
        if (start < 0) {
          start = t;
          start_is_j = (e == 0);
        }
      } else if (cnt != 2) {
        return false;  // an edge shared by more than two cells
      }
    }
  }
  if (n_single != 0 && n_single != 2) return false;
  closed = (n_single == 0);
  unsigned used = 0;
  uint32_t cur;
  if (closed) {
    start = 0;
    ring[0] = ja[0];
    cur = ka[0];
  } else {
    ring[0] = start_is_j ? ja[start] : ka[start];
    cur = start_is_j ? ka[start] : ja[start];
  }
  used |= 1U << start;
  len = 1;
  if (cell_item != nullptr) cell_item[0] = adj[it0 + start];
  for (int step = 1; step < m; ++step) {
    ring[len++] = cur;
    int nxt = -1;
    for (int u = 0; u < m; ++u) {
      if (!(used & (1U << u)) && (ja[u] == cur || ka[u] == cur)) {
        nxt = u;
        break;
      }
    }
    if (nxt < 0) return false;  // chain broken: more than one fan
    used |= 1U << nxt;
    if (cell_item != nullptr) cell_item[step] = adj[it0 + nxt];
    cur = (ja[nxt] == cur) ? ka[nxt] : ja[nxt];
  }
  if (closed) {
    if (cur != ring[0]) return false;
  } else {
    ring[len++] = cur;
  }
  (void)row;
  return true;
}

__global__ void k_fan_lengths(int64_t n_rows, const int32_t* __restrict__ adj_ptr, const uint32_t* __restrict__ adj,
                              const uint32_t* __restrict__ cell_nodes, uint8_t* __restrict__ ring_len, int* __restrict__ max_len) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int len = 0;
  if (r < n_rows) {
    const int32_t it0 = adj_ptr[r];
    const int m = adj_ptr[r + 1] - it0;
    if (m >= 1 && m <= kMaxFan - 1) {
      uint32_t ring[kMaxFan + 1];
      bool closed;
      if (!build_ring(static_cast<int32_t>(r), m, adj, it0, cell_nodes, ring, len, closed)) len = 0;
    }
    ring_len[r] = static_cast<uint8_t>(len);  // 0 = irregular
  }
  len = __reduce_max_sync(0xffffffffU, len);
  if ((threadIdx.x & 31) == 0) atomicMax(max_len, len);
}

// ring slot = node id | (slot of that column inside row r) << 28; rowinfo[r] = slot of the diagonal | closed << 7.
// The slot of column c in row r is its rank among the row's column ids = ring ids and r itself.
__global__ void k_fan_fill(int64_t n_rows, int W, const int32_t* __restrict__ adj_ptr, const uint32_t* __restrict__ adj,
                           const uint32_t* __restrict__ cell_nodes, const uint8_t* __restrict__ ring_len, uint32_t* __restrict__ nbr,
                           uint8_t* __restrict__ rowinfo) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  uint32_t ring[kMaxFan + 1];
  int len = 0;
  bool closed = false;
  if (ring_len[r] != 0 && ring_len[r] <= W) {
    const int32_t it0 = adj_ptr[r];
    build_ring(static_cast<int32_t>(r), adj_ptr[r + 1] - it0, adj, it0, cell_nodes, ring, len, closed);
  }
  int posd = 0;
  for (int s = 0; s < len; ++s) posd += (ring[s] < static_cast<uint32_t>(r)) ? 1 : 0;
  for (int s = 0; s < W; ++s) {
    uint32_t v = kNil;
    if (s < len) {
      uint32_t rank = (static_cast<uint32_t>(r) < ring[s]) ? 1U : 0U;
      for (int u = 0; u < len; ++u) rank += (ring[u] < ring[s]) ? 1U : 0U;
      v = ring[s] | (rank << 28);
    }
    nbr[static_cast<int64_t>(s) * n_rows + r] = v;
  }
  rowinfo[r] = static_cast<uint8_t>(posd | (closed ? 0x80 : 0));
}

__global__ void k_flag_irregular(int64_t n_rows, int W, const uint8_t* __restrict__ ring_len, const int32_t* __restrict__ adj_ptr,
                                 uint8_t* __restrict__ flag) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const bool has_cells = adj_ptr[r + 1] > adj_ptr[r];
  flag[r] = (has_cells && (ring_len[r] == 0 || ring_len[r] > W)) ? 1 : 0;
}

// dof table == vertex table and all cells are triangles?
__global__ void k_check_nodal(int64_t n_cells, int stride, const int32_t* __restrict__ dofs, const uint8_t* __restrict__ nldof,
                              const uint32_t* __restrict__ cell_nodes, int* __restrict__ bad) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[c];
  if (v.w != kNil || nldof[c] != 3 || dofs[c * stride] != static_cast<int32_t>(v.x) || dofs[c * stride + 1] != static_cast<int32_t>(v.y) ||
      dofs[c * stride + 2] != static_cast<int32_t>(v.z))
    *bad = 1;
}

// ---- the kernel -----------------------------------------------------------------------------------------------------
struct FanParams {
  double a00, a01, a10, a11;  // effective diffusion tensor (already transposed for row-major output)
  double gamma;
  double wsum;                // sum of the rule's weights (1/2 for every rule of the reference)
  double m_diag, m_off;       // reference mass tensor of the rule
  double beta;
};

__device__ __forceinline__ double fast_rcp(double x) {
  // 1/x for x = |det J| > 0 (cell areas: far from the double range limits, degenerate cells are rejected at upload):
  // hardware seed (MUFU.RCP64H, ~20 bits) + two Newton steps -> relative error ~1e-16, no slow path
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// contributions of the triangle (i, p, q), a = x_p - x_i, b = x_q - x_i, to row i: k1 (column p), k2 (column q);
// the diagonal contribution of the stiffness part is -(k1 + k2).  MODE 0: scalar alpha, no mass (c = wsum * alpha).
template <int MODE>
__device__ __forceinline__ void fan_cell(const FanParams& P, double c, double ax, double ay, double bx, double by, double aa,
                                         double bb, double& k1, double& k2, double& kd) {
  const double det = ax * by - ay * bx;
  const double adet = fabs(det);
  const double ridet = fast_rcp(adet);
  if (MODE == 0) {
    const double ab = ax * bx + ay * by;
    const double s = c * ridet;
    k1 = s * (ab - bb);
    k2 = s * (ab - aa);
    kd = -(k1 + k2);
  } else {
    // M = N A N^T / |det| with N = adj(J) = [by -bx; -ay ax];  row 0 of the element matrix: wsum * ghat_b^T M ghat_0
    const double n00 = by, n01 = -bx, n10 = -ay, n11 = ax;
    const double t00 = n00 * P.a00 + n01 * P.a10, t01 = n00 * P.a01 + n01 * P.a11;
    const double t10 = n10 * P.a00 + n11 * P.a10, t11 = n10 * P.a01 + n11 * P.a11;
    const double m00 = t00 * n00 + t01 * n01, m01 = t00 * n10 + t01 * n11;
    const double m10 = t10 * n00 + t11 * n01, m11 = t10 * n10 + t11 * n11;
    const double s = P.wsum * ridet;
    k1 = -s * (m00 + m01);
    k2 = -s * (m10 + m11);
    kd = -(k1 + k2);
    const double gm = P.gamma * adet;
    kd = fma(gm, P.m_diag, kd);
    k1 = fma(gm, P.m_off, k1);
    k2 = fma(gm, P.m_off, k2);
  }
}

__device__ __forceinline__ void prefetch_l2(const void* a) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a)); }

// One row per lane.  nid[]: node ids of the ring in fan order (< 0 = empty slot; a row that is not a single fan has
// nid[0] < 0 and is left alone), pos[]: slot of every ring column inside the row, [v0, v1): the row's value range, posd:
// slot of the diagonal, closed: the fan closes around the node.
// The values of the warp's rows are collected in the warp's shared-memory stage: at offset v0 - wbase when the 32
// rows are consecutive (always without a row list, mostly with the row lists of a partition) -- then the stage is the
// image of one contiguous range of the output and leaves as full 128-byte lines -- else at lane * (W + 1), and every
// lane copies its own row.  A warp that contains a non-fan row (computed by the generic kernel) takes the second way.
template <int W, int MODE>
__device__ __forceinline__ void fan_row(const int32_t (&nid)[W], const int (&pos)[W], int32_t v0, int32_t v1, int posd, bool closed,
                                        int32_t r, bool in_range, int lane, bool listed, double* __restrict__ stage,
                                        const double* __restrict__ node_coords, const FanParams& P, double* __restrict__ values) {
  const bool regular = in_range && nid[0] >= 0;
  const int32_t wbase = __shfl_sync(0xffffffffU, v0, 0);
  const int32_t r_first = __shfl_sync(0xffffffffU, r, 0);
  const bool consecutive = !listed || !__any_sync(0xffffffffU, in_range && r != r_first + lane);
  const bool staged = consecutive && !__any_sync(0xffffffffU, in_range && !regular && v1 > v0);
  double* dst = stage + (staged ? v0 - wbase : lane * (W + 1));
  if (regular) {
    const double2* nc = reinterpret_cast<const double2*>(node_coords);
    const double2 xi = __ldg(nc + r);
    double dx[W], dy[W], dd[W];
    int m = 0;
#pragma unroll
    for (int s = 0; s < W; ++s) {
      const bool valid = nid[s] >= 0;
      const double2 p = __ldg(nc + (valid ? nid[s] : r));
      dx[s] = p.x - xi.x;
      dy[s] = p.y - xi.y;
      dd[s] = dx[s] * dx[s] + dy[s] * dy[s];
      m += valid ? 1 : 0;
    }
    const double c = P.wsum * P.a00;
    const double* old = values + v0;
    double diag = 0.0;
    if (m == W && closed) {
      // full closed ring (every interior vertex of a regular mesh): static indices only
      double k1, k2, kd, carry, first;
      fan_cell<MODE>(P, c, dx[0], dy[0], dx[1], dy[1], dd[0], dd[1], first, carry, kd);
      diag = kd;
#pragma unroll
      for (int s = 1; s < W; ++s) {
        const int u = (s + 1 < W) ? s + 1 : 0;
        fan_cell<MODE>(P, c, dx[s], dy[s], dx[u], dy[u], dd[s], dd[u], k1, k2, kd);
        diag += kd;
        double v = carry + k1;
        if (MODE != 0 && P.beta != 0.0) v = fma(P.beta, old[pos[s]], v);
        dst[pos[s]] = v;
        carry = k2;
      }
      first += carry;
      if (MODE != 0 && P.beta != 0.0) first = fma(P.beta, old[pos[0]], first);
      dst[pos[0]] = first;
    } else {
      double carry = 0.0, first = 0.0;
      double lx = dx[0], ly = dy[0], ld = dd[0];
      int lpos = pos[0];
#pragma unroll
      for (int s = 0; s + 1 < W; ++s) {
        if (s + 1 < m) {  // cell (i, n_s, n_s+1)
          double k1, k2, kd;
          fan_cell<MODE>(P, c, dx[s], dy[s], dx[s + 1], dy[s + 1], dd[s], dd[s + 1], k1, k2, kd);
          diag += kd;
          if (s == 0) {
            first = k1;
          } else {
            double v = carry + k1;
            if (MODE != 0 && P.beta != 0.0) v = fma(P.beta, old[pos[s]], v);
            dst[pos[s]] = v;
          }
          carry = k2;
          lx = dx[s + 1];
          ly = dy[s + 1];
          ld = dd[s + 1];
          lpos = pos[s + 1];
        }
      }
      if (closed) {  // wrap-around cell (i, n_m-1, n_0)
        double k1, k2, kd;
        fan_cell<MODE>(P, c, lx, ly, dx[0], dy[0], ld, dd[0], k1, k2, kd);
        diag += kd;
        carry += k1;
        first += k2;
      }
      if (MODE != 0 && P.beta != 0.0) {
        carry = fma(P.beta, old[lpos], carry);
        first = fma(P.beta, old[pos[0]], first);
      }
      dst[lpos] = carry;
      dst[pos[0]] = first;
    }
    if (MODE != 0 && P.beta != 0.0) diag = fma(P.beta, old[posd], diag);
    dst[posd] = diag;
  }
  __syncwarp();
  if (staged) {
    const unsigned ballot = __ballot_sync(0xffffffffU, in_range);
    if (ballot == 0) return;
    const int total = __shfl_sync(0xffffffffU, v1, 31 - __clz(ballot)) - wbase;
    double* out = values + wbase;
#pragma unroll
    for (int k = 0; k < W + 1; ++k) {  // a row has at most W + 1 stored values
      const int idx = k * 32 + lane;
      if (idx < total) out[idx] = stage[idx];
    }
  } else if (regular) {
    const int len = v1 - v0;
    for (int k = 0; k < len; ++k) values[v0 + k] = dst[k];
  }
}

// The warps of an SM run in phase (all wait for the plan words, all wait for the coordinates, all compute), so the
// DRAM latency of the kernel's two dependent load levels is not hidden by occupancy (ncu: half of all stall samples on
// the first use of either).  Warp 0 of every CTA therefore pulls the lines of the CTA that runs pf_dist rows later
// (about 3/4 of a wave of resident CTAs) into L2 -- plan, row pointers, and the CTA's own coordinates; the ring
// neighbours' coordinates are some other row's own coordinates inside the same window.  Both load levels then hit L2.

// one thread per row: wide plan (ring slot = node id | slot-in-row << 28; rowinfo = slot of the diagonal | closed << 7)
template <int W, int MODE>
__global__ void __launch_bounds__(128, 8) k_assemble_p1_fan(int n_rows, int n_total_rows, const uint32_t* __restrict__ nbr,
                                                         const uint8_t* __restrict__ rowinfo, const double* __restrict__ node_coords,
                                                         const int32_t* __restrict__ outer, const int32_t* __restrict__ row_list, int row0,
                                                         int pf_dist, FanParams P, double* __restrict__ values) {
  extern __shared__ double stage_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = t < n_rows;
  const int32_t r = in_range ? (row_list != nullptr ? row_list[t] : t + row0) : 0;
  if (pf_dist > 0 && row_list == nullptr && warp == 0) {
    const int tp = blockIdx.x * blockDim.x + pf_dist;
    if (tp + 128 <= n_rows) {
      const size_t rp = static_cast<size_t>(tp) + row0;
      for (int L = lane; L < 4 * W + 21; L += 32) {  // 128-byte lines: 4 per ring slot, 1 row info, 4 row pointers, 16 coordinates
        const char* a;
        if (L < 4 * W) {
          a = reinterpret_cast<const char*>(nbr + static_cast<size_t>(L >> 2) * n_total_rows + rp) + (L & 3) * 128;
        } else if (L == 4 * W) {
          a = reinterpret_cast<const char*>(rowinfo + rp);
        } else if (L < 4 * W + 5) {
          a = reinterpret_cast<const char*>(outer + rp) + (L - 4 * W - 1) * 128;
        } else {
          a = reinterpret_cast<const char*>(node_coords + 2 * rp) + (L - 4 * W - 5) * 128;
        }
        prefetch_l2(a);
      }
    }
  }
  int32_t v0 = 0, v1 = 0;
  int32_t nid[W];
  int pos[W];
  int info = 0;
  if (in_range) {
    v0 = __ldg(outer + r);
    v1 = __ldg(outer + r + 1);
    info = __ldg(rowinfo + r);
#pragma unroll
    for (int s = 0; s < W; ++s) {
      const uint32_t u = __ldg(nbr + static_cast<size_t>(s) * n_total_rows + r);
      nid[s] = u == kNil ? -1 : static_cast<int32_t>(u & 0x0fffffffU);
      pos[s] = static_cast<int>(u >> 28);
    }
  } else {
#pragma unroll
    for (int s = 0; s < W; ++s) {
      nid[s] = -1;
      pos[s] = 0;
    }
  }
  fan_row<W, MODE>(nid, pos, v0, v1, info & 0x7f, (info & 0x80) != 0, r, in_range, lane, row_list != nullptr,
                   stage_all + warp * (32 * (W + 2)), node_coords, P, values);
}

// Same kernel on the compact plan (W = 6): ring ids as 16-bit offsets from the row id; the 4-bit slots of the ring
// positions, the diagonal slot and the closed flag packed into one word per row: 16 B of plan per row instead of 25 B.
template <int MODE>
__global__ void __launch_bounds__(128, 8) k_assemble_p1_fan_compact(int n_rows, int n_total_rows, const int16_t* __restrict__ nbr16,
                                                                 const uint32_t* __restrict__ info32, const double* __restrict__ node_coords,
                                                                 const int32_t* __restrict__ outer, const int32_t* __restrict__ row_list,
                                                                 int row0, int pf_dist, FanParams P, double* __restrict__ values) {
  constexpr int W = 6;
  extern __shared__ double stage_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = t < n_rows;
  const int32_t r = in_range ? (row_list != nullptr ? row_list[t] : t + row0) : 0;
  if (pf_dist > 0 && row_list == nullptr && warp == 0) {
    const int tp = blockIdx.x * blockDim.x + pf_dist;
    if (tp + 128 <= n_rows) {
      // 36 lines of 128 B: 2 per ring slot, 4 row info, 4 row pointers, 16 coordinates
      const size_t rp = static_cast<size_t>(tp) + row0;
      const char* a;
      if (lane < 12) {
        a = reinterpret_cast<const char*>(nbr16 + static_cast<size_t>(lane >> 1) * n_total_rows + rp) + (lane & 1) * 128;
      } else if (lane < 16) {
        a = reinterpret_cast<const char*>(info32 + rp) + (lane - 12) * 128;
      } else if (lane < 20) {
        a = reinterpret_cast<const char*>(outer + rp) + (lane - 16) * 128;
      } else {
        a = reinterpret_cast<const char*>(node_coords + 2 * rp) + (lane - 20) * 128;
      }
      prefetch_l2(a);
      if (lane < 4) prefetch_l2(reinterpret_cast<const char*>(node_coords + 2 * rp) + (12 + lane) * 128);
    }
  }
  int32_t v0 = 0, v1 = 0;
  int32_t nid[W];
  int pos[W];
  uint32_t w = 0;
  if (in_range) {
    v0 = __ldg(outer + r);
    v1 = __ldg(outer + r + 1);
    w = __ldg(info32 + r);
#pragma unroll
    for (int s = 0; s < W; ++s) {
      const int d = __ldg(nbr16 + static_cast<size_t>(s) * n_total_rows + r);
      nid[s] = d == 0 ? -1 : r + d;
      pos[s] = static_cast<int>((w >> (4 * s)) & 15U);
    }
  } else {
#pragma unroll
    for (int s = 0; s < W; ++s) {
      nid[s] = -1;
      pos[s] = 0;
    }
  }
  fan_row<W, MODE>(nid, pos, v0, v1, static_cast<int>((w >> 24) & 15U), ((w >> 28) & 1U) != 0, r, in_range, lane, row_list != nullptr,
                   stage_all + warp * (32 * (W + 2)), node_coords, P, values);
}

// plan compaction: one thread per row; *bad is raised if a ring id is further than 32767 from its row
__global__ void k_fan_compact(int64_t n_rows, const uint32_t* __restrict__ nbr, const uint8_t* __restrict__ rowinfo,
                              int16_t* __restrict__ nbr16, uint32_t* __restrict__ info32, int* __restrict__ bad) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int info = rowinfo[r];
  uint32_t w = (static_cast<uint32_t>(info & 0x7f) << 24) | (static_cast<uint32_t>((info >> 7) & 1) << 28);
  for (int s = 0; s < 6; ++s) {
    const uint32_t v = nbr[static_cast<int64_t>(s) * n_rows + r];
    int16_t d = 0;
    if (v != kNil) {
      const int64_t delta = static_cast<int64_t>(v & 0x0fffffffU) - r;
      if (delta < -32767 || delta > 32767 || delta == 0) {
        *bad = 1;
      } else {
        d = static_cast<int16_t>(delta);
      }
      w |= (v >> 28) << (4 * s);
    }
    nbr16[static_cast<int64_t>(s) * n_rows + r] = d;
  }
  info32[r] = w;
}

// ---- load vector on the vertex rings (AssembleVectorLocally + ScalarLoadElementVectorProvider, assembler.h:298-327,
// loc_comp_ellbvp.h:691-746, for FeLagrangeO1Tria with a constant source): entry i = f * lhat * sum over the ring's cells of
// |det J|, lhat = sum_k w_k phi_a(x_k) (the same for a = 0, 1, 2 for every symmetric rule).  One thread per row, the ring is the
// only connectivity read, the result leaves as one coalesced 8-byte store per row: 37 B per row with the compact plan instead
// of a dof-table gather per cell plus FP64 atomics (round 1: 2.74 ms at 1.0e8 triangles for 2.4 GB of compulsory traffic).
__global__ void k_load_fan_fill(int64_t n_rows, int W, const int32_t* __restrict__ adj_ptr, const uint32_t* __restrict__ adj,
                                const uint32_t* __restrict__ cell_nodes, const uint8_t* __restrict__ ring_len, uint32_t* __restrict__ nbr,
                                uint8_t* __restrict__ info) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  uint32_t ring[kMaxFan + 1];
  int len = 0;
  bool closed = false;
  const int32_t it0 = adj_ptr[r];
  const int m = adj_ptr[r + 1] - it0;
  if (ring_len[r] != 0 && ring_len[r] <= W) build_ring(static_cast<int32_t>(r), m, adj, it0, cell_nodes, ring, len, closed);
  for (int s = 0; s < W; ++s) nbr[static_cast<int64_t>(s) * n_rows + r] = s < len ? ring[s] : kNil;
  info[r] = static_cast<uint8_t>(m == 0 ? 3 : (len == 0 ? 2 : (closed ? 1 : 0)));  // 3: no cells, 2: not a single fan (generic kernel)
}

__global__ void k_load_fan_compact(int64_t n_rows, const uint32_t* __restrict__ nbr, int16_t* __restrict__ nbr16, int* __restrict__ bad) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  for (int s = 0; s < 6; ++s) {
    const uint32_t v = nbr[static_cast<int64_t>(s) * n_rows + r];
    int16_t d = 0;
    if (v != kNil) {
      const int64_t delta = static_cast<int64_t>(v) - r;
      if (delta < -32767 || delta > 32767 || delta == 0) *bad = 1;
      else d = static_cast<int16_t>(delta);
    }
    nbr16[static_cast<int64_t>(s) * n_rows + r] = d;
  }
}

__global__ void k_flag_info2(int64_t n, const uint8_t* __restrict__ info, uint8_t* __restrict__ flag) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r < n) flag[r] = info[r] == 2 ? 1 : 0;
}

template <int W, bool COMPACT>
__global__ void __launch_bounds__(128, 8) k_load_p1_fan(int n_rows, const void* __restrict__ nbr_any, const uint8_t* __restrict__ info,
                                                      const double* __restrict__ node_coords, int pf_dist, double c, double beta,
                                                      double* __restrict__ vec) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t* nbr = static_cast<const uint32_t*>(nbr_any);
  const int16_t* nbr16 = static_cast<const int16_t*>(nbr_any);
  if (pf_dist > 0 && warp == 0) {
    const int rp = blockIdx.x * blockDim.x + pf_dist;
    if (rp + 128 <= n_rows) {
      constexpr int per = COMPACT ? 2 : 4;  // 128-byte lines per ring array and CTA
      for (int L = lane; L < per * W + 17; L += 32) {
        const char* a;
        if (L < per * W) {
          a = COMPACT ? reinterpret_cast<const char*>(nbr16 + static_cast<size_t>(L / per) * n_rows + rp) + (L % per) * 128
                      : reinterpret_cast<const char*>(nbr + static_cast<size_t>(L / per) * n_rows + rp) + (L % per) * 128;
        } else if (L == per * W) {
          a = reinterpret_cast<const char*>(info + rp);
        } else {
          a = reinterpret_cast<const char*>(node_coords + 2 * static_cast<size_t>(rp)) + (L - per * W - 1) * 128;
        }
        prefetch_l2(a);
      }
    }
  }
  if (r >= n_rows) return;
  const int inf = __ldg(info + r);
  if (inf == 2) return;  // computed by the generic kernel
  int32_t nid[W];
#pragma unroll
  for (int s = 0; s < W; ++s) {
    if (COMPACT) {
      const int d = __ldg(nbr16 + static_cast<size_t>(s) * n_rows + r);
      nid[s] = d == 0 ? -1 : r + d;
    } else {
      const uint32_t u = __ldg(nbr + static_cast<size_t>(s) * n_rows + r);
      nid[s] = u == kNil ? -1 : static_cast<int32_t>(u);
    }
  }
  const double2* nc = reinterpret_cast<const double2*>(node_coords);
  const double2 xi = __ldg(nc + r);
  double dx[W], dy[W];
  int m = 0;
#pragma unroll
  for (int s = 0; s < W; ++s) {
    const bool valid = nid[s] >= 0;
    const double2 p = __ldg(nc + (valid ? nid[s] : r));
    dx[s] = p.x - xi.x;
    dy[s] = p.y - xi.y;
    m += valid ? 1 : 0;
  }
  // |det J| of the cells (i, n_s, n_s+1) in ring order = ascending position in the fan, the wrap-around cell last
  double sum = 0.0, lx = dx[0], ly = dy[0];
#pragma unroll
  for (int s = 0; s + 1 < W; ++s) {
    if (s + 1 < m) {
      sum += fabs(dx[s] * dy[s + 1] - dy[s] * dx[s + 1]);
      lx = dx[s + 1];
      ly = dy[s + 1];
    }
  }
  if (inf == 1) sum += fabs(lx * dy[0] - ly * dx[0]);
  const double v = c * sum;
  vec[r] = beta == 0.0 ? v : fma(beta, vec[r], v);
}

// ---- the same rings with a source that varies from cell to cell (tabulated per cell or per quadrature point) -----------------
// Entry i = sum over the ring cells K of |det J_K| sum_q w_q phi_a(x_q) f_K(q), a = local index of node i in K: next to the ring
// the plan holds, per ring position, the cell and that local index (cells[s][r] = cell << 2 | a; 24 B per row for rings of six),
// and the kernel reads the cell's record of source values (one 32-byte sector for the default rule).  One thread per row, one
// coalesced store per row, no atomics, fixed order of additions (ring order).  Against the two-pass scheme (assemble.cu), which
// moves every element-vector entry through L2 by a scattered 8-byte store, at 1.0e8 triangles.
__global__ void k_load_fan_cells(int64_t n_rows, int W, const int32_t* __restrict__ adj_ptr, const uint32_t* __restrict__ adj,
                                 const uint32_t* __restrict__ cell_nodes, const uint8_t* __restrict__ info, uint32_t* __restrict__ cells) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  uint32_t ring[kMaxFan + 1], item[kMaxFan + 1];
  int len = 0;
  bool closed = false;
  const int32_t it0 = adj_ptr[r];
  const int m = adj_ptr[r + 1] - it0;
  const bool planned = info[r] < 2 && m >= 1 && m <= W;
  if (planned) build_ring(static_cast<int32_t>(r), m, adj, it0, cell_nodes, ring, len, closed, item);
  for (int s = 0; s < W; ++s)
    cells[static_cast<int64_t>(s) * n_rows + r] = (planned && s < m) ? (((item[s] >> 4) << 2) | (item[s] & 3U)) : kNil;
}

struct LoadSource {
  const double* data;  // [n_cells][stride]
  int stride;          // 1: one value per cell (nq entries of w below are then summed on the host into w[a][0])
  int nq;
  int vec4;            // stride == 4 and a 32-byte aligned table: one 256-bit load per record
  double w[3][4];      // w_q phi_a(x_q) of the rule in use (zero beyond nq)
};

template <int W, bool COMPACT>
__global__ void __launch_bounds__(128, 6) k_load_p1_fan_src(int n_rows, const void* __restrict__ nbr_any, const uint8_t* __restrict__ info,
                                                          const uint32_t* __restrict__ cells, const double* __restrict__ node_coords,
                                                          LoadSource S, double beta, double* __restrict__ vec) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const uint32_t* nbr = static_cast<const uint32_t*>(nbr_any);
  const int16_t* nbr16 = static_cast<const int16_t*>(nbr_any);
  const int inf = __ldg(info + r);
  if (inf == 2) return;  // computed by the generic kernel
  int32_t nid[W];
  uint32_t cw[W];
#pragma unroll
  for (int s = 0; s < W; ++s) {
    if (COMPACT) {
      const int d = __ldg(nbr16 + static_cast<size_t>(s) * n_rows + r);
      nid[s] = d == 0 ? -1 : r + d;
    } else {
      const uint32_t u = __ldg(nbr + static_cast<size_t>(s) * n_rows + r);
      nid[s] = u == kNil ? -1 : static_cast<int32_t>(u);
    }
    cw[s] = __ldg(cells + static_cast<size_t>(s) * n_rows + r);
  }
  const double2* nc = reinterpret_cast<const double2*>(node_coords);
  const double2 xi = __ldg(nc + r);
  double dx[W], dy[W];
#pragma unroll
  for (int s = 0; s < W; ++s) {
    const double2 p = __ldg(nc + (nid[s] >= 0 ? nid[s] : r));
    dx[s] = p.x - xi.x;
    dy[s] = p.y - xi.y;
  }
  double sum = 0.0;
#pragma unroll
  for (int s = 0; s < W; ++s) {
    if (cw[s] != kNil) {  // cell s lies between ring[s] and ring[s + 1] (ring[0] for the cell that closes the ring)
      const int u = (s + 1 < W && nid[s + 1 < W ? s + 1 : 0] >= 0) ? s + 1 : 0;
      const double det = fabs(dx[s] * dy[u] - dy[s] * dx[u]);
      const int a = static_cast<int>(cw[s] & 3U);
      const double* f = S.data + static_cast<size_t>(cw[s] >> 2) * S.stride;
      double e;
      if (S.stride == 1) {
        e = S.w[a][0] * __ldg(f);
      } else if (S.vec4) {  // the cell's record in one 256-bit load (four 8-byte loads at a 32-byte lane stride touch every sector four times)
        double f0, f1, f2, f3;
        asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(f0), "=d"(f1), "=d"(f2), "=d"(f3) : "l"(f));
        e = S.w[a][0] * f0 + (S.nq > 1 ? S.w[a][1] * f1 : 0.0) + (S.nq > 2 ? S.w[a][2] * f2 : 0.0) + (S.nq > 3 ? S.w[a][3] * f3 : 0.0);  // (entries beyond the rule's points are the caller's padding: never multiplied)
      } else {
        e = 0.0;
        for (int q = 0; q < S.nq; ++q) e += S.w[a][q] * __ldg(f + q);
      }
      sum += det * e;
    }
  }
  vec[r] = beta == 0.0 ? sum : fma(beta, vec[r], sum);
}

}  // namespace

// ---- host side ---------------------------------------------------------------------------------------------------------
int p1_fan_prepare(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, lfgpu_pattern* p) {
  if (p->fan_state != 0) return LFGPU_OK;
  p->fan_state = -1;
  if (mesh->n_quad != 0 || mesh->cell_coords != nullptr || p->i_dofs != p->o_dofs || p->n_outer != mesh->n_nodes || p->n_outer >= (1LL << 28) - 1) return LFGPU_OK;  // ids share a word with a 4-bit slot
  cudaStream_t st = ctx->stream;
  int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 256);
  LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_flags, 0, 16, st));
  const unsigned gc = static_cast<unsigned>(cdiv(p->n_cells, 256)), gr = static_cast<unsigned>(cdiv(p->n_outer, 256));
  k_check_nodal<<<gc, 256, 0, st>>>(p->n_cells, p->o_stride, p->o_dofs, p->o_nldof, mesh->cell_nodes, d_flags);
  LFGPU_LAUNCH_CHECK(ctx);
  uint8_t* ring_len = nullptr;
  LFGPU_CUDA_CHECK(ctx, cudaMalloc(&ring_len, p->n_outer));
  k_fan_lengths<<<gr, 256, 0, st>>>(p->n_outer, p->adj_ptr, p->adj, mesh->cell_nodes, ring_len, d_flags + 1);
  LFGPU_LAUNCH_CHECK(ctx);
  int h[2] = {0, 0};
  cudaError_t e = cudaMemcpyAsync(h, d_flags, sizeof(h), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess || h[0] != 0 || h[1] < 2) {
    cudaFree(ring_len);
    LFGPU_CUDA_CHECK(ctx, e);
    return LFGPU_OK;  // not a nodal P1 table: stay with the generic kernels
  }
  // ring width: the kernel is instantiated for 6, 8, 10, 12; longer rings would be irregular rows
  int W = h[1] <= 6 ? 6 : (h[1] <= 8 ? 8 : (h[1] <= 10 ? 10 : 12));
  uint8_t* flag = nullptr;
  int32_t* iota = nullptr;
  int64_t* d_num = nullptr;
  void* tmp = nullptr;
  auto cleanup = [&]() { cudaFree(ring_len); cudaFree(flag); cudaFree(iota); cudaFree(d_num); cudaFree(tmp); };
#define FAN_CHECK(expr)                                                             \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      set_last_error(ctx, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
      cleanup();                                                                    \
      return LFGPU_ERR_CUDA;                                                        \
    }                                                                               \
  } while (0)
  FAN_CHECK(cudaMalloc(&p->fan_nbr, sizeof(uint32_t) * static_cast<size_t>(W) * p->n_outer));
  FAN_CHECK(cudaMalloc(&p->fan_rowinfo, p->n_outer + 4));  // the prefetching kernel reads whole words
  k_fan_fill<<<gr, 256, 0, st>>>(p->n_outer, W, p->adj_ptr, p->adj, mesh->cell_nodes, ring_len, p->fan_nbr, p->fan_rowinfo);
  ctx->launches++;
  FAN_CHECK(cudaMalloc(&flag, p->n_outer));
  k_flag_irregular<<<gr, 256, 0, st>>>(p->n_outer, W, ring_len, p->adj_ptr, flag);
  ctx->launches++;
  FAN_CHECK(cudaMalloc(&iota, sizeof(int32_t) * p->n_outer));
  FAN_CHECK(cudaMalloc(&d_num, sizeof(int64_t)));
  cub::CountingInputIterator<int32_t> count_it(0);
  size_t tb = 0;
  if (and_row_keep(ctx, p->n_outer, flag, p->row_keep) != LFGPU_OK) return LFGPU_ERR_CUDA;  // rows nobody asks for need no generic kernel
  cub::DeviceSelect::Flagged(nullptr, tb, count_it, flag, iota, d_num, p->n_outer, st);
  FAN_CHECK(cudaMalloc(&tmp, tb));
  FAN_CHECK(cub::DeviceSelect::Flagged(tmp, tb, count_it, flag, iota, d_num, p->n_outer, st));
  int64_t n_irr = 0;
  FAN_CHECK(cudaMemcpyAsync(&n_irr, d_num, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  FAN_CHECK(cudaStreamSynchronize(st));
  if (n_irr > 0) {
    FAN_CHECK(cudaMalloc(&p->fan_irregular, sizeof(int32_t) * n_irr));
    FAN_CHECK(cudaMemcpyAsync(p->fan_irregular, iota, sizeof(int32_t) * n_irr, cudaMemcpyDeviceToDevice, st));
    FAN_CHECK(cudaStreamSynchronize(st));
  }
#undef FAN_CHECK
  cleanup();
  p->n_irregular = n_irr;
  p->fan_w = W;
  p->fan_state = 1;
  // compact plan where it applies (LFGPU_FAN_COMPACT=0 keeps the wide one, for A/B measurements)
  const char* env = std::getenv("LFGPU_FAN_COMPACT");
  if (W == 6 && (env == nullptr || env[0] != '0')) {
    int16_t* n16 = nullptr;
    uint32_t* i32 = nullptr;
    cudaError_t ce = cudaMalloc(&n16, sizeof(int16_t) * 6 * static_cast<size_t>(p->n_outer));
    if (ce == cudaSuccess) ce = cudaMalloc(&i32, sizeof(uint32_t) * static_cast<size_t>(p->n_outer));
    int bad = 1;
    if (ce == cudaSuccess) ce = cudaMemsetAsync(d_flags, 0, sizeof(int), st);
    if (ce == cudaSuccess) {
      k_fan_compact<<<gr, 256, 0, st>>>(p->n_outer, p->fan_nbr, p->fan_rowinfo, n16, i32, d_flags);
      ctx->launches++;
      ce = cudaMemcpyAsync(&bad, d_flags, sizeof(int), cudaMemcpyDeviceToHost, st);
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce == cudaSuccess && bad == 0) {
      p->fan_nbr16 = n16;
      p->fan_info32 = i32;
      cudaFree(p->fan_nbr);
      cudaFree(p->fan_rowinfo);
      p->fan_nbr = nullptr;
      p->fan_rowinfo = nullptr;
    } else {
      cudaFree(n16);
      cudaFree(i32);
      (void)cudaGetLastError();  // an allocation failure here only means: stay with the wide plan
    }
  }
  return LFGPU_OK;
}

// launches the fan kernel over all rows (row_list == nullptr, row0 < 0), over the listed rows, or over the contiguous
// range [row0, row0 + n_rows) (row_list == nullptr, row0 >= 0)
int p1_fan_launch(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* p, const double alpha[4], int tensor, double gamma,
                  double wsum, double m_diag, double m_off, double beta, const int32_t* row_list, int64_t n_rows, double* d_values,
                  int64_t row0) {
  FanParams P;
  P.a00 = alpha[0]; P.a01 = tensor ? alpha[1] : 0.0; P.a10 = tensor ? alpha[2] : 0.0; P.a11 = tensor ? alpha[3] : alpha[0];
  P.gamma = gamma;
  P.wsum = wsum;
  P.m_diag = m_diag;
  P.m_off = m_off;
  P.beta = beta;
  const int64_t rows = (row_list != nullptr || row0 >= 0) ? n_rows : p->n_outer;
  if (rows <= 0) return LFGPU_OK;
  const int64_t first_row = (row_list == nullptr && row0 >= 0) ? row0 : 0;
  const int threads = 128;
  const unsigned grid = static_cast<unsigned>(cdiv(rows, threads));
  const int W = p->fan_w;
  const bool simple = !tensor && gamma == 0.0 && beta == 0.0;  // the plain Laplacian: leanest instantiation
  const size_t smem = sizeof(double) * (threads / 32) * 32 * (W + 2);
  // L2 prefetch distance in rows: one wave of resident CTAs (flat optimum between 1/2 and 2 waves on B200, slower
  // below 1/4 and above 4; LFGPU_FAN_PFD = percent of a wave, 0 = off)
  static const int pfd_env = [] { const char* e = std::getenv("LFGPU_FAN_PFD"); return e != nullptr ? std::atoi(e) : 100; }();
  const int ipf = static_cast<int>((static_cast<int64_t>(ctx->sm_count) * 8 * threads * pfd_env / 100) & ~static_cast<int64_t>(127));
  const int irows = static_cast<int>(rows), itotal = static_cast<int>(p->n_outer), ifirst = static_cast<int>(first_row);  // < 2^28 (p1_fan_prepare)
  if (p->fan_nbr16 != nullptr) {
    if (simple)
      k_assemble_p1_fan_compact<0><<<grid, threads, smem, ctx->stream>>>(irows, itotal, p->fan_nbr16, p->fan_info32, mesh->node_coords,
                                                                          p->outer, row_list, ifirst, ipf, P, d_values);
    else
      k_assemble_p1_fan_compact<1><<<grid, threads, smem, ctx->stream>>>(irows, itotal, p->fan_nbr16, p->fan_info32, mesh->node_coords,
                                                                          p->outer, row_list, ifirst, ipf, P, d_values);
    LFGPU_LAUNCH_CHECK(ctx);
    return LFGPU_OK;
  }
#define FAN_LAUNCH(WW)                                                                                                          \
  if (simple)                                                                                                                   \
    k_assemble_p1_fan<WW, 0><<<grid, threads, smem, ctx->stream>>>(irows, itotal, p->fan_nbr, p->fan_rowinfo, mesh->node_coords, \
                                                                   p->outer, row_list, ifirst, ipf, P, d_values);        \
  else                                                                                                                          \
    k_assemble_p1_fan<WW, 1><<<grid, threads, smem, ctx->stream>>>(irows, itotal, p->fan_nbr, p->fan_rowinfo, mesh->node_coords, \
                                                                   p->outer, row_list, ifirst, ipf, P, d_values)
  switch (W) {
    case 6: FAN_LAUNCH(6); break;
    case 8: FAN_LAUNCH(8); break;
    case 10: FAN_LAUNCH(10); break;
    default: FAN_LAUNCH(12); break;
  }
#undef FAN_LAUNCH
  LFGPU_LAUNCH_CHECK(ctx);
  return LFGPU_OK;
}

// Load vector of FeLagrangeO1Tria with a constant source on the vertex rings.  *handled = 1 if the call was served here
// (plan applicable); irregular rows (not a single fan) are returned in the dofmap's lv_irregular list for the generic kernel.
// src_data != nullptr: source tabulated per cell (src_stride 1) or per quadrature point (src_stride >= nq), wtab = [3][nq] w_q phi_a(x_q);
// c is then unused.
int p1_load_fan(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_dofmap* dc, double c, double beta, double* d_vec, int* handled,
                const double* src_data, int src_stride, int nq, const double* wtab) {
  *handled = 0;
  if (src_data != nullptr && (nq < 1 || nq > 4)) return LFGPU_OK;
  lfgpu_dofmap* d = const_cast<lfgpu_dofmap*>(dc);
  if (d->lv_state == 0) {
    d->lv_state = -1;
    if (mesh->n_quad != 0 || mesh->cell_coords != nullptr || d->n_dofs != mesh->n_nodes || d->n_dofs >= (1LL << 31) - 256) return LFGPU_OK;
    int rc = dofmap_gather_plan(ctx, d);
    if (rc != LFGPU_OK) return rc;
    cudaStream_t st = ctx->stream;
    int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 256);
    LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_flags, 0, 16, st));
    const int64_t N = d->n_dofs;
    const unsigned gc = static_cast<unsigned>(cdiv(d->n_cells, 256)), gr = static_cast<unsigned>(cdiv(N, 256));
    k_check_nodal<<<gc, 256, 0, st>>>(d->n_cells, d->stride, d->cell_dofs, d->n_ldof, mesh->cell_nodes, d_flags);
    LFGPU_LAUNCH_CHECK(ctx);
    uint8_t *ring_len = nullptr, *flag = nullptr;
    int32_t* iota = nullptr;
    int64_t* d_num = nullptr;
    void* tmp = nullptr;
    uint32_t* nbr = nullptr;
    uint8_t* info = nullptr;
    int16_t* n16 = nullptr;
    auto cleanup = [&]() { cudaFree(ring_len); cudaFree(flag); cudaFree(iota); cudaFree(d_num); cudaFree(tmp); };
    auto fail = [&]() { cleanup(); cudaFree(nbr); cudaFree(info); cudaFree(n16); };
#define LV_CHECK(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      set_last_error(ctx, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
      fail();                                                                       \
      return LFGPU_ERR_CUDA;                                                        \
    }                                                                               \
  } while (0)
    LV_CHECK(cudaMalloc(&ring_len, N));
    k_fan_lengths<<<gr, 256, 0, st>>>(N, d->g_ptr, d->g_items, mesh->cell_nodes, ring_len, d_flags + 1);
    ctx->launches++;
    int h[2] = {0, 0};
    LV_CHECK(cudaMemcpyAsync(h, d_flags, sizeof(h), cudaMemcpyDeviceToHost, st));
    LV_CHECK(cudaStreamSynchronize(st));
    if (h[0] != 0 || h[1] < 2) {  // not the nodal P1 table
      fail();
      return LFGPU_OK;
    }
    const int W = h[1] <= 6 ? 6 : (h[1] <= 8 ? 8 : (h[1] <= 10 ? 10 : 12));
    LV_CHECK(cudaMalloc(&nbr, sizeof(uint32_t) * (static_cast<size_t>(W) * N + 128)));
    LV_CHECK(cudaMalloc(&info, N + 128));
    k_load_fan_fill<<<gr, 256, 0, st>>>(N, W, d->g_ptr, d->g_items, mesh->cell_nodes, ring_len, nbr, info);
    ctx->launches++;
    LV_CHECK(cudaMalloc(&flag, N));
    k_flag_info2<<<gr, 256, 0, st>>>(N, info, flag);
    ctx->launches++;
    LV_CHECK(cudaMalloc(&iota, sizeof(int32_t) * N));
    LV_CHECK(cudaMalloc(&d_num, sizeof(int64_t)));
    cub::CountingInputIterator<int32_t> count_it(0);
    size_t tb = 0;
    cub::DeviceSelect::Flagged(nullptr, tb, count_it, flag, iota, d_num, N, st);
    LV_CHECK(cudaMalloc(&tmp, tb));
    LV_CHECK(cub::DeviceSelect::Flagged(tmp, tb, count_it, flag, iota, d_num, N, st));
    int64_t n_irr = 0;
    LV_CHECK(cudaMemcpyAsync(&n_irr, d_num, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    LV_CHECK(cudaStreamSynchronize(st));
    if (n_irr > 0) {
      LV_CHECK(cudaMalloc(&d->lv_irregular, sizeof(int32_t) * n_irr));
      LV_CHECK(cudaMemcpyAsync(d->lv_irregular, iota, sizeof(int32_t) * n_irr, cudaMemcpyDeviceToDevice, st));
      LV_CHECK(cudaStreamSynchronize(st));
    }
    d->n_lv_irregular = n_irr;
    if (W == 6) {  // compact ring: 16-bit offsets from the row id
      LV_CHECK(cudaMalloc(&n16, sizeof(int16_t) * (6 * static_cast<size_t>(N) + 256)));
      LV_CHECK(cudaMemsetAsync(d_flags, 0, sizeof(int), st));
      k_load_fan_compact<<<gr, 256, 0, st>>>(N, nbr, n16, d_flags);
      ctx->launches++;
      int bad = 1;
      LV_CHECK(cudaMemcpyAsync(&bad, d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
      LV_CHECK(cudaStreamSynchronize(st));
      if (bad == 0) {
        cudaFree(nbr);
        nbr = nullptr;
      } else {
        cudaFree(n16);
        n16 = nullptr;
      }
    }
#undef LV_CHECK
    cleanup();
    d->lv_w = W;
    d->lv_nbr = nbr;
    d->lv_nbr16 = n16;
    d->lv_info = info;
    d->lv_state = 1;
  }
  if (d->lv_state != 1) return LFGPU_OK;
  const int threads = 128;
  const int n = static_cast<int>(d->n_dofs);
  const unsigned grid = static_cast<unsigned>(cdiv(n, threads));
  const int ipf = static_cast<int>((static_cast<int64_t>(ctx->sm_count) * 8 * threads) & ~static_cast<int64_t>(127));
  if (src_data != nullptr) {
    if (d->lv_cells == nullptr) {  // cell and local index per ring position, built when a tabulated source is first used
      LFGPU_CUDA_CHECK(ctx, cudaMalloc(&d->lv_cells, sizeof(uint32_t) * (static_cast<size_t>(d->lv_w) * n + 128)));
      // the ring ids as 32-bit words are gone if the compact form was adopted; the rings are rebuilt from the adjacency either way
      k_load_fan_cells<<<static_cast<unsigned>(cdiv(n, 256)), 256, 0, ctx->stream>>>(n, d->lv_w, d->g_ptr, d->g_items, mesh->cell_nodes, d->lv_info,
                                                                                    d->lv_cells);
      LFGPU_LAUNCH_CHECK(ctx);
    }
    LoadSource S;
    S.data = src_data;
    S.stride = src_stride;
    S.nq = nq;
    S.vec4 = (src_stride == 4 && (reinterpret_cast<uintptr_t>(src_data) & 31) == 0) ? 1 : 0;
    for (int a = 0; a < 3; ++a)
      for (int q = 0; q < 4; ++q) S.w[a][q] = q < nq ? wtab[a * nq + q] : 0.0;
    if (src_stride == 1) {
      for (int a = 0; a < 3; ++a) {
        double t = 0.0;
        for (int q = 0; q < nq; ++q) t += wtab[a * nq + q];
        S.w[a][0] = t;
      }
    }
    if (d->lv_nbr16 != nullptr) {
      k_load_p1_fan_src<6, true><<<grid, threads, 0, ctx->stream>>>(n, d->lv_nbr16, d->lv_info, d->lv_cells, mesh->node_coords, S, beta, d_vec);
    } else {
      switch (d->lv_w) {
        case 6: k_load_p1_fan_src<6, false><<<grid, threads, 0, ctx->stream>>>(n, d->lv_nbr, d->lv_info, d->lv_cells, mesh->node_coords, S, beta, d_vec); break;
        case 8: k_load_p1_fan_src<8, false><<<grid, threads, 0, ctx->stream>>>(n, d->lv_nbr, d->lv_info, d->lv_cells, mesh->node_coords, S, beta, d_vec); break;
        case 10: k_load_p1_fan_src<10, false><<<grid, threads, 0, ctx->stream>>>(n, d->lv_nbr, d->lv_info, d->lv_cells, mesh->node_coords, S, beta, d_vec); break;
        default: k_load_p1_fan_src<12, false><<<grid, threads, 0, ctx->stream>>>(n, d->lv_nbr, d->lv_info, d->lv_cells, mesh->node_coords, S, beta, d_vec); break;
      }
    }
  } else if (d->lv_nbr16 != nullptr) {
    k_load_p1_fan<6, true><<<grid, threads, 0, ctx->stream>>>(n, d->lv_nbr16, d->lv_info, mesh->node_coords, ipf, c, beta, d_vec);
  } else {
    switch (d->lv_w) {
      case 6: k_load_p1_fan<6, false><<<grid, threads, 0, ctx->stream>>>(n, d->lv_nbr, d->lv_info, mesh->node_coords, ipf, c, beta, d_vec); break;
      case 8: k_load_p1_fan<8, false><<<grid, threads, 0, ctx->stream>>>(n, d->lv_nbr, d->lv_info, mesh->node_coords, ipf, c, beta, d_vec); break;
      case 10: k_load_p1_fan<10, false><<<grid, threads, 0, ctx->stream>>>(n, d->lv_nbr, d->lv_info, mesh->node_coords, ipf, c, beta, d_vec); break;
      default: k_load_p1_fan<12, false><<<grid, threads, 0, ctx->stream>>>(n, d->lv_nbr, d->lv_info, mesh->node_coords, ipf, c, beta, d_vec); break;
    }
  }
  LFGPU_LAUNCH_CHECK(ctx);
  *handled = 1;
  return LFGPU_OK;
}

}  // namespace lfgpu
