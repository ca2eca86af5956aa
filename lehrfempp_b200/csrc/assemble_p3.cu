// P3 (FeLagrangeO3Tria) fast path of the numeric pass: row kernels (product code).  The arithmetic and the index logic live
// in rows_p3_core.h (shared with the host emulation that checks them against the oracle on the CPU); this file holds the
// CUDA side: plan kernels, the three row kernels (vertex rows of 37, edge-dof rows of 16, cell rows of 10 stored values)
// and the launch code.  Layout of the work as in assemble_p2.cu: one thread per row, the 32 consecutive rows of a warp
// staged in shared memory and written as full 128-byte lines, rows that do not fit (boundary, valence != 6) computed by
// the generic gather kernel, L2 prefetch of the plan lines one wave ahead.
#include <algorithm>
#include <cstdlib>

#include <cub/cub.cuh>

#include "lfgpu_internal.cuh"
#include "rows_p3_core.h"

namespace lfgpu {
namespace {

using namespace p3;
constexpr uint32_t kNil = 0xFFFFFFFFu;

// ---- plan construction ------------------------------------------------------------------------------------------------
// dof table == [node ids | 6 edge dofs in [n_nodes, base_int) | base_int + cell] and all cells triangles with ten dofs?
__global__ void k_p3_check(int64_t n_cells, int stride, int64_t n_nodes, int64_t base_int, const int32_t* __restrict__ dofs,
                           const uint8_t* __restrict__ nldof, const uint32_t* __restrict__ cell_nodes, int* __restrict__ bad) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const uint4 v = reinterpret_cast<const uint4*>(cell_nodes)[c];
  const int32_t* d = dofs + c * stride;
  bool ok = v.w == kNil && nldof[c] == 10 && d[0] == static_cast<int32_t>(v.x) && d[1] == static_cast<int32_t>(v.y) &&
            d[2] == static_cast<int32_t>(v.z) && d[9] == base_int + c;
  if (ok) {
    for (int b = 3; b < 9; ++b) ok = ok && d[b] >= n_nodes && d[b] < base_int;
  }
  if (!ok) *bad = 1;
}

__global__ void k_p3_vertex_plan(int64_t n_nodes, int o_stride, int pos_row, const int32_t* __restrict__ adj_ptr,
                                 const uint32_t* __restrict__ adj, const uint32_t* __restrict__ cell_nodes,
                                 const uint8_t* __restrict__ pos, const int32_t* __restrict__ outer, int32_t* __restrict__ nbr,
                                 uint32_t* __restrict__ slots, uint8_t* __restrict__ irregular, int cc) {
  // cc != 0 (mesh with per-cell corner coordinates): nbr holds the (cell, corner) words of rows_p3_core.h instead of node numbers
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_nodes) return;
  const int32_t it0 = adj_ptr[r];
  const int m = adj_ptr[r + 1] - it0;
  int32_t ring[kRing] = {-1, -1, -1, -1, -1, -1};
  uint32_t w[kVertexSlotWords] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t cw[kRing] = {0, 0, 0, 0, 0, 0};
  const bool ok = vertex_plan(r, m, adj + it0, cell_nodes, pos, o_stride, pos_row, outer[r + 1] - outer[r], ring, w, cw);
  if (cc)
    for (int k = 0; k < kRing; ++k) ring[k] = static_cast<int32_t>(cw[k]);
  for (int k = 0; k < kRing; ++k) nbr[static_cast<int64_t>(k) * n_nodes + r] = ok ? ring[k] : -1;
  for (int j = 0; j < kVertexSlotWords; ++j) slots[static_cast<int64_t>(j) * n_nodes + r] = ok ? w[j] : 0U;
  irregular[r] = (!ok && m > 0) ? 1 : 0;
}

// closed rings of 3..8 cells (unstructured meshes): gnbr[k][r] = n_k (-1 beyond the ring / irregular row), gslots[0..12][r]
__global__ void k_p3_vertex_plan_general(int64_t n_nodes, int o_stride, int pos_row, const int32_t* __restrict__ adj_ptr,
                                         const uint32_t* __restrict__ adj, const uint32_t* __restrict__ cell_nodes,
                                         const uint8_t* __restrict__ pos, const int32_t* __restrict__ outer, int32_t* __restrict__ gnbr,
                                         uint32_t* __restrict__ gslots, uint8_t* __restrict__ irregular) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_nodes) return;
  const int32_t it0 = adj_ptr[r];
  const int m = adj_ptr[r + 1] - it0;
  int32_t ring[kMaxRing];
  uint32_t w[kGeneralSlotWords];
  const bool ok = vertex_plan_general(r, m, adj + it0, cell_nodes, pos, o_stride, pos_row, outer[r + 1] - outer[r], ring, w);
  for (int k = 0; k < kMaxRing; ++k) gnbr[static_cast<int64_t>(k) * n_nodes + r] = ok ? ring[k] : -1;
  for (int j = 0; j < kGeneralSlotWords; ++j) gslots[static_cast<int64_t>(j) * n_nodes + r] = ok ? w[j] : 0U;
  irregular[r] = (!ok && m > 0) ? 1 : 0;
}

__global__ void k_p3_edge_plan(int64_t n_nodes, int64_t n_erows, int o_stride, int pos_row, const int32_t* __restrict__ adj_ptr,
                               const uint32_t* __restrict__ adj, const uint32_t* __restrict__ cell_nodes,
                               const uint8_t* __restrict__ pos, const int32_t* __restrict__ outer, int32_t* __restrict__ enb,
                               uint32_t* __restrict__ eslots, uint8_t* __restrict__ irregular, int cc) {
  // cc != 0: enb[0..1] hold the (cell, corner) words of the two cells, enb[2..3] are unused
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_erows) return;
  const int64_t r = n_nodes + e;
  const int32_t it0 = adj_ptr[r];
  const int m = adj_ptr[r + 1] - it0;
  int32_t ids[4] = {-1, -1, -1, -1};
  uint32_t w[kEdgeSlotWords] = {0, 0};
  uint32_t cw[2] = {0, 0};
  const bool ok = edge_plan(m, adj + it0, cell_nodes, pos, o_stride, pos_row, outer[r + 1] - outer[r], ids, w, cw);
  if (cc) {
    ids[0] = static_cast<int32_t>(cw[0]);
    ids[1] = static_cast<int32_t>(cw[1]);
    ids[2] = ids[3] = 0;
  }
  for (int k = 0; k < 4; ++k) enb[static_cast<int64_t>(k) * n_erows + e] = ok ? ids[k] : -1;
  for (int j = 0; j < kEdgeSlotWords; ++j) eslots[static_cast<int64_t>(j) * n_erows + e] = ok ? w[j] : 0U;
  irregular[r] = (!ok && m > 0) ? 1 : 0;
}

// cell rows are always regular if the row has ten stored values
__global__ void k_p3_cell_flags(int64_t n_cells, int64_t base_int, const int32_t* __restrict__ outer, uint8_t* __restrict__ irregular,
                                int* __restrict__ bad) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const int64_t r = base_int + c;
  irregular[r] = 0;
  if (outer[r + 1] - outer[r] != kCellRowLen) *bad = 1;
}

// ---- the kernels --------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void* a) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a)); }

// copy-out shared by the kernels: the warp's stage is the image of the contiguous value range of its 32 rows
// SWZ: the stage is addressed through stage_ix<true> (rows_p3_core.h); `off` = start of the lane's row inside the stage
template <int LEN, bool SWZ = false>
__device__ __forceinline__ void write_rows(bool staged, bool regular, bool in_range, int lane, int32_t v0, int32_t v1, int32_t wbase,
                                           const double* __restrict__ stage, const double* __restrict__ dst, double* __restrict__ values,
                                           double beta, int off = 0) {
  __syncwarp();
  if (staged) {
    const unsigned ballot = __ballot_sync(0xffffffffU, in_range);
    if (ballot == 0) return;
    const int total = __shfl_sync(0xffffffffU, v1, 31 - __clz(ballot)) - wbase;
    double* out = values + wbase;
    if (beta == 0.0) {
#pragma unroll
      for (int k = 0; k < LEN; ++k) {
        const int idx = k * 32 + lane;
        if (idx < total) out[idx] = stage[stage_ix<SWZ>(idx)];
      }
    } else {  // accumulate (assembler.h:84-88)
      for (int idx = lane; idx < total; idx += 32) out[idx] = fma(beta, out[idx], stage[stage_ix<SWZ>(idx)]);
    }
  } else if (regular) {
    if (SWZ) {
      for (int k = 0; k < LEN; ++k) {
        const double v = stage[stage_ix<true>(off + k)];
        values[v0 + k] = beta == 0.0 ? v : fma(beta, values[v0 + k], v);
      }
    } else {
      for (int k = 0; k < LEN; ++k) values[v0 + k] = beta == 0.0 ? dst[k] : fma(beta, values[v0 + k], dst[k]);
    }
  }
}

// edge vectors of a cell from ITS corners: word = cell << 4 | ia << 2 | ib, origin = the third corner (rows_p3_core.h)
struct DevCellVec {
  const double2* cc;    // cell_coords as [n_cells][4] points
  const int32_t* words;
  __device__ __forceinline__ void operator()(int k, double& ax, double& ay, double& bx, double& by) const {
    const uint32_t cw = static_cast<uint32_t>(words[k]);
    const double2* c = cc + 4 * static_cast<size_t>(cw >> 4);
    const int ia = (cw >> 2) & 3, ib = cw & 3;
    const double2 x0 = __ldg(c + (3 - ia - ib)), xa = __ldg(c + ia), xb = __ldg(c + ib);
    ax = xa.x - x0.x; ay = xa.y - x0.y;
    bx = xb.x - x0.x; by = xb.y - x0.y;
  }
};

// MINB = 4 (opt-in, LFGPU_P3_VOCC=4; MODE 1 only): 128 registers instead of 168 -- a fourth CTA per SM for 172 bytes of spills.
// The row-range arguments come LAST in every row kernel: in front they moved `Params P` from a 16-byte to an 8-byte aligned
// offset of the parameter space, and ptxas then spilled 52 bytes in this kernel (MODE 1, 168 registers = the cap of 3 CTAs per
// SM) that the measured kernel did not spill.  With them at the end the layout up to `values` is the measured one.
// CC: node_coords is the mesh's cell_coords array, nbr holds (cell, corner) words
template <int MODE, int MINB = 3, bool CC = false>
__global__ void __launch_bounds__(128, MINB) k_p3_vertex_rows(int n_rows, const int32_t* __restrict__ nbr,
                                                         const uint32_t* __restrict__ slots, const double* __restrict__ node_coords,
                                                         const int32_t* __restrict__ outer, int pf_dist, Params P,
                                                         double* __restrict__ values, int first, int end, double beta) {
  // rows [first, end) of the n_rows vertex rows (the whole range, or one GPU's share of it)
  extern __shared__ double stage_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = first + blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = r < end;
  if (pf_dist > 0 && warp == 0) {
    // lines of the CTA about one wave later: 4 per plan array (6 ring + 9 slot arrays), 4 of row pointers, 16 of coordinates
    const int rp = first + blockIdx.x * blockDim.x + pf_dist;
    if (rp + 128 <= end) {
      for (int L = lane; L < (CC ? 64 : 80); L += 32) {  // CC: the corners are indexed by cell, not by row
        const char* a;
        if (L < 24) a = reinterpret_cast<const char*>(nbr + static_cast<size_t>(L >> 2) * n_rows + rp) + (L & 3) * 128;
        else if (L < 60) a = reinterpret_cast<const char*>(slots + static_cast<size_t>((L - 24) >> 2) * n_rows + rp) + (L & 3) * 128;
        else if (L < 64) a = reinterpret_cast<const char*>(outer + rp) + (L - 60) * 128;
        else a = reinterpret_cast<const char*>(node_coords + 2 * static_cast<size_t>(rp)) + (L - 64) * 128;
        prefetch_l2(a);
      }
    }
  }
  int32_t v0 = 0, v1 = 0;
  int32_t nid[kRing];
  uint32_t w[kVertexSlotWords];
#pragma unroll
  for (int s = 0; s < kRing; ++s) nid[s] = -1;
#pragma unroll
  for (int j = 0; j < kVertexSlotWords; ++j) w[j] = 0U;
  if (in_range) {
    v0 = __ldg(outer + r);
    v1 = __ldg(outer + r + 1);
#pragma unroll
    for (int s = 0; s < kRing; ++s) nid[s] = __ldg(nbr + static_cast<size_t>(s) * n_rows + r);
#pragma unroll
    for (int j = 0; j < kVertexSlotWords; ++j) w[j] = __ldg(slots + static_cast<size_t>(j) * n_rows + r);
  }
  const bool regular = in_range && nid[0] >= 0;
  const int32_t wbase = __shfl_sync(0xffffffffU, v0, 0);
  const bool staged = !__any_sync(0xffffffffU, in_range && !regular && v1 > v0);
  double* stage = stage_all + warp * (32 * (kVertexRowLen + 1));
  double* dst = stage + (staged ? v0 - wbase : lane * (kVertexRowLen + 1));
  if (regular) {
    const double2* nc = reinterpret_cast<const double2*>(node_coords);
    if (CC) {
      vertex_row_cv<MODE>(P, DevCellVec{nc, nid}, w, dst);
    } else {
      const double2 xi = __ldg(nc + r);
      double dx[kRing], dy[kRing];
#pragma unroll
      for (int s = 0; s < kRing; ++s) {
        const double2 p = __ldg(nc + nid[s]);
        dx[s] = p.x - xi.x;
        dy[s] = p.y - xi.y;
      }
      vertex_row<MODE>(P, dx, dy, w, dst);
    }
  }
  write_rows<kVertexRowLen>(staged, regular, in_range, lane, v0, v1, wbase, stage, dst, values, beta);
}

// vertex rows with closed rings of 3..8 cells (rows_p3_core.h: vertex_row_general); rows of different lengths (1 + 6m) share a
// warp, the staged copy-out only needs their value ranges to be consecutive
template <int MODE>
__global__ void __launch_bounds__(128, 2) k_p3_vertex_rows_general(int first, int end, int n_rows, const int32_t* __restrict__ gnbr,
                                                                 const uint32_t* __restrict__ gslots, const double* __restrict__ node_coords,
                                                                 const int32_t* __restrict__ outer, Params P, double* __restrict__ values,
                                                                 double beta) {
  extern __shared__ double stage_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = first + blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = r < end;
  int32_t v0 = 0, v1 = 0;
  int32_t nid[kMaxRing];
  uint32_t w[kGeneralSlotWords];
#pragma unroll
  for (int s = 0; s < kMaxRing; ++s) nid[s] = -1;
#pragma unroll
  for (int j = 0; j < kGeneralSlotWords; ++j) w[j] = 0U;
  if (in_range) {
    v0 = __ldg(outer + r);
    v1 = __ldg(outer + r + 1);
#pragma unroll
    for (int s = 0; s < kMaxRing; ++s) nid[s] = __ldg(gnbr + static_cast<size_t>(s) * n_rows + r);
#pragma unroll
    for (int j = 0; j < kGeneralSlotWords; ++j) w[j] = __ldg(gslots + static_cast<size_t>(j) * n_rows + r);
  }
  const bool regular = in_range && nid[0] >= 0;
  const int32_t wbase = __shfl_sync(0xffffffffU, v0, 0);
  const bool staged = !__any_sync(0xffffffffU, in_range && !regular && v1 > v0);
  double* stage = stage_all + warp * (32 * (kMaxVertexRowLen + 1));
  double* dst = stage + (staged ? v0 - wbase : lane * (kMaxVertexRowLen + 1));
  if (regular) {
    const double2* nc = reinterpret_cast<const double2*>(node_coords);
    const double2 xi = __ldg(nc + r);
    double dx[kMaxRing], dy[kMaxRing];
#pragma unroll
    for (int s = 0; s < kMaxRing; ++s) {
      const double2 q = __ldg(nc + (nid[s] >= 0 ? nid[s] : r));
      dx[s] = q.x - xi.x;
      dy[s] = q.y - xi.y;
    }
    vertex_row_general<MODE>(P, dx, dy, w, dst);
  }
  __syncwarp();
  if (staged) {
    const unsigned ballot = __ballot_sync(0xffffffffU, in_range);
    if (ballot == 0) return;
    const int total = __shfl_sync(0xffffffffU, v1, 31 - __clz(ballot)) - wbase;
    double* out = values + wbase;
#pragma unroll
    for (int k = 0; k < kMaxVertexRowLen; ++k) {
      const int idx = k * 32 + lane;
      if (idx < total) out[idx] = beta == 0.0 ? stage[idx] : fma(beta, out[idx], stage[idx]);
    }
  } else if (regular) {
    for (int k = 0; k < v1 - v0; ++k) values[v0 + k] = beta == 0.0 ? dst[k] : fma(beta, values[v0 + k], dst[k]);
  }
}

template <int MODE, int MINB = 4, bool CC = false>
__global__ void __launch_bounds__(128, MINB) k_p3_edge_rows(int n_erows, int row0, const int32_t* __restrict__ enb,
                                                       const uint32_t* __restrict__ eslots, const double* __restrict__ node_coords,
                                                       const int32_t* __restrict__ outer, int pf_dist, int pfc_dist, Params P,
                                                       double* __restrict__ values, int first, int end, double beta) {
  // edge-dof rows [first, end) of n_erows; row e of them is matrix row row0 + e
  extern __shared__ double stage_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = first + blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = e < end;
  if (pf_dist > 0 && warp == 0) {
    const int ep = first + blockIdx.x * blockDim.x + pf_dist;
    if (ep + 128 <= end && lane < 28) {  // 4 lines per id array, 4 per slot array, 4 of row pointers
      const char* a;
      if (lane < 16) a = reinterpret_cast<const char*>(enb + static_cast<size_t>(lane >> 2) * n_erows + ep) + (lane & 3) * 128;
      else if (lane < 24) a = reinterpret_cast<const char*>(eslots + static_cast<size_t>((lane - 16) >> 2) * n_erows + ep) + (lane & 3) * 128;
      else a = reinterpret_cast<const char*>(outer + row0 + ep) + (lane - 24) * 128;
      prefetch_l2(a);
    }
  }
  // Coordinates of the row pfc_dist rows ahead (about one wave of resident CTAs): every thread reads that row's four ids now and
  // asks L2 for the four coordinate lines at the END of its own work, when the ids have arrived.  ncu (round 2): 28 % of the
  // kernel's stall samples sat on the first use of the gathered coordinates, 9 % on the ids -- edges are numbered column by
  // column, nodes row by row, so a warp's 32 rows gather from 32 different lines per id and the one-wave-ahead prefetch of the
  // PLAN lines (warp 0 above) does nothing for them.  (The round-1 variant read every 4th row's ids only: no gain.)
  int32_t pf_id[4] = {-1, -1, -1, -1};
  if (pfc_dist > 0 && e + pfc_dist < end) {
#pragma unroll
    for (int k = 0; k < (CC ? 2 : 4); ++k) pf_id[k] = __ldg(enb + static_cast<size_t>(k) * n_erows + e + pfc_dist);
  }
  int32_t v0 = 0, v1 = 0;
  int32_t ip = -1, iq = 0, io1 = 0, io2 = 0;
  uint32_t w[kEdgeSlotWords] = {0U, 0U};
  if (in_range) {
    v0 = __ldg(outer + row0 + e);
    v1 = __ldg(outer + row0 + e + 1);
    ip = __ldg(enb + e);
    iq = __ldg(enb + static_cast<size_t>(n_erows) + e);
    if (!CC) {
      io1 = __ldg(enb + 2 * static_cast<size_t>(n_erows) + e);
      io2 = __ldg(enb + 3 * static_cast<size_t>(n_erows) + e);
    }
    w[0] = __ldg(eslots + e);
    w[1] = __ldg(eslots + static_cast<size_t>(n_erows) + e);
  }
  const bool regular = in_range && ip >= 0;
  const int32_t wbase = __shfl_sync(0xffffffffU, v0, 0);
  const bool staged = !__any_sync(0xffffffffU, in_range && !regular && v1 > v0);
  double* stage = stage_all + warp * (32 * kEdgeRowLen);
  const int off = staged ? v0 - wbase : lane * kEdgeRowLen;
  if (regular) {
    const double2* nc = reinterpret_cast<const double2*>(node_coords);
    if (CC) {
      const int32_t cw[2] = {ip, iq};
      const DevCellVec cv{nc, cw};
      double a1x, a1y, b1x, b1y, a2x, a2y, b2x, b2y;
      cv(0, a1x, a1y, b1x, b1y);
      cv(1, a2x, a2y, b2x, b2y);
      edge_row2<MODE, true>(P, a1x, a1y, b1x, b1y, a2x, a2y, b2x, b2y, w, stage, off);
    } else {
      const double2 xp = __ldg(nc + ip), xq = __ldg(nc + iq), x1 = __ldg(nc + io1), x2 = __ldg(nc + io2);
      edge_row<MODE, true>(P, xq.x - xp.x, xq.y - xp.y, x1.x - xp.x, x1.y - xp.y, x2.x - xp.x, x2.y - xp.y, w, stage, off);
    }
  }
#pragma unroll
  for (int k = 0; k < (CC ? 2 : 4); ++k)
    if (pf_id[k] >= 0) prefetch_l2(CC ? node_coords + 8 * static_cast<size_t>(static_cast<uint32_t>(pf_id[k]) >> 4) : node_coords + 2 * static_cast<size_t>(pf_id[k]));
  write_rows<kEdgeRowLen, true>(staged, regular, in_range, lane, v0, v1, wbase, stage, stage, values, beta, off);
}

// slots of the ten list positions of every cell in its own row (list position 9), one nibble each: 8 bytes per cell instead
// of a 12-byte gather out of the cell's 120-byte block of the scatter map (ncu: k_p3_cell_rows read 654 MB for 4.2e6 cells)
__global__ void k_p3_cell_plan(int64_t n_cells, int o_stride, int pos_row, const uint8_t* __restrict__ pos, uint2* __restrict__ cslots) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const uint8_t* pp = pos + (static_cast<size_t>(c) * o_stride + 9) * pos_row;
  uint32_t w0 = 0, w1 = 0;
  for (int b = 0; b < 8; ++b) w0 |= static_cast<uint32_t>(pp[b] & 15U) << (4 * b);
  for (int b = 8; b < 10; ++b) w1 |= static_cast<uint32_t>(pp[b] & 15U) << (4 * (b - 8));
  cslots[c] = make_uint2(w0, w1);
}

// one thread per cell: row 9 of its element matrix (cell_nodes + the compact slot word pair)
// CC: node_coords is the mesh's cell_coords array (the cell's own corners, no gather through cell_nodes)
template <int MODE, bool CC = false>
__global__ void __launch_bounds__(128, 4) k_p3_cell_rows(int row0, const uint32_t* __restrict__ cell_nodes, const uint2* __restrict__ cslots,
                                                       const double* __restrict__ node_coords, const int32_t* __restrict__ outer, int pf_dist,
                                                       Params P, double* __restrict__ values, int first, int end, double beta) {
  // cells [first, end); cell c is matrix row row0 + c
  extern __shared__ double stage_all[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = first + blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = c < end;
  if (pf_dist > 0 && warp == 0) {
    const int cp = first + blockIdx.x * blockDim.x + pf_dist;
    if (cp + 128 <= end && lane < 28) {  // 16 lines of cell corners, 8 of slot words, 4 of row pointers
      const char* a;
      if (lane < 16) a = reinterpret_cast<const char*>(cell_nodes + 4 * static_cast<size_t>(cp)) + lane * 128;
      else if (lane < 24) a = reinterpret_cast<const char*>(cslots + cp) + (lane - 16) * 128;
      else a = reinterpret_cast<const char*>(outer + row0 + cp) + (lane - 24) * 128;
      if (!CC || lane >= 16) prefetch_l2(a);
    }
    if (CC && cp + 128 <= end) {  // 64 lines of corner coordinates (64 bytes per cell)
      prefetch_l2(reinterpret_cast<const char*>(node_coords + 8 * static_cast<size_t>(cp)) + lane * 128);
      prefetch_l2(reinterpret_cast<const char*>(node_coords + 8 * static_cast<size_t>(cp)) + (32 + lane) * 128);
    }
  }
  int32_t v0 = 0, v1 = 0;
  if (in_range) {
    v0 = __ldg(outer + row0 + c);
    v1 = __ldg(outer + row0 + c + 1);
  }
  const int32_t wbase = __shfl_sync(0xffffffffU, v0, 0);
  double* stage = stage_all + warp * (32 * kCellRowLen);
  const int off = v0 - wbase;
  if (in_range) {
    const uint2 sw = __ldg(cslots + c);
    const uint32_t pw[3] = {sw.x, sw.y, 0U};
    const double2* nc = reinterpret_cast<const double2*>(node_coords);
    double2 x0, x1, x2;
    if (CC) {
      x0 = __ldg(nc + 4 * static_cast<size_t>(c));
      x1 = __ldg(nc + 4 * static_cast<size_t>(c) + 1);
      x2 = __ldg(nc + 4 * static_cast<size_t>(c) + 2);
    } else {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(cell_nodes) + c);
      x0 = __ldg(nc + v.x);
      x1 = __ldg(nc + v.y);
      x2 = __ldg(nc + v.z);
    }
    cell_row<MODE, true, true>(P, x1.x - x0.x, x1.y - x0.y, x2.x - x0.x, x2.y - x0.y, pw, stage, off);
  }
  write_rows<kCellRowLen, true>(true, in_range, in_range, lane, v0, v1, wbase, stage, stage, values, beta, off);
}

__global__ void k_count_flags(int64_t n, const uint8_t* __restrict__ flag, int* __restrict__ cnt) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const unsigned b = __ballot_sync(0xffffffffU, i < n && flag[i] != 0);
  if ((threadIdx.x & 31) == 0 && b != 0) atomicAdd(cnt, __popc(b));
}

}  // namespace

// ---- host side -----------------------------------------------------------------------------------------------------------
int p3_rows_prepare(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, lfgpu_pattern* p) {
  if (p->p3_state != 0) return LFGPU_OK;
  p->p3_state = -1;
  const int64_t nn = mesh->n_nodes, base_int = p->n_outer - p->n_cells, ner = base_int - nn;
  if (mesh->n_quad != 0 || p->i_dofs != p->o_dofs || ner <= 0 || p->pos_bytes != 1 || p->pos == nullptr ||
      p->n_outer >= (1LL << 31) - 256 || p->n_cells >= (1LL << 27))
    return LFGPU_OK;
  // cells with their own corner coordinates: the plan carries (cell, corner) words instead of node numbers (rows_p3_core.h);
  // the general-valence vertex plan is not built for them (their other vertex rows go to the generic kernel)
  const int cc = mesh->cell_coords != nullptr ? 1 : 0;
  p->p3_cc = cc != 0;
  cudaStream_t st = ctx->stream;
  int* d_flags = reinterpret_cast<int*>(static_cast<char*>(ctx->d_scratch) + 256);
  LFGPU_CUDA_CHECK(ctx, cudaMemsetAsync(d_flags, 0, 16, st));
  k_p3_check<<<static_cast<unsigned>(cdiv(p->n_cells, 256)), 256, 0, st>>>(p->n_cells, p->o_stride, nn, base_int, p->o_dofs, p->o_nldof,
                                                                            mesh->cell_nodes, d_flags);
  LFGPU_LAUNCH_CHECK(ctx);
  uint8_t* flag = nullptr;
  LFGPU_CUDA_CHECK(ctx, cudaMalloc(&flag, p->n_outer));
  k_p3_cell_flags<<<static_cast<unsigned>(cdiv(p->n_cells, 256)), 256, 0, st>>>(p->n_cells, base_int, p->outer, flag, d_flags);
  ctx->launches++;
  int bad = 1;
  cudaError_t e0 = cudaMemcpyAsync(&bad, d_flags, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e0 == cudaSuccess) e0 = cudaStreamSynchronize(st);
  if (e0 != cudaSuccess || bad != 0) {
    cudaFree(flag);
    LFGPU_CUDA_CHECK(ctx, e0);
    return LFGPU_OK;  // not the dof layout of FeSpaceLagrangeO3 on triangles: stay with the generic kernels
  }
  int32_t* iota = nullptr;
  int64_t* d_num = nullptr;
  void* tmp = nullptr;
  auto cleanup = [&]() { cudaFree(flag); cudaFree(iota); cudaFree(d_num); cudaFree(tmp); };
  auto drop_plan = [&]() {
    cudaFree(p->p3v_nbr); cudaFree(p->p3v_slots); cudaFree(p->p3e_nbr); cudaFree(p->p3e_slots); cudaFree(p->p3_irregular);
    cudaFree(p->p3g_nbr); cudaFree(p->p3g_slots); cudaFree(p->p3c_slots); cudaFree(p->p3e_newid); cudaFree(p->p3e_xy);
    p->p3e_newid = nullptr; p->p3e_xy = nullptr;
    p->p3c_slots = nullptr;
    p->p3v_nbr = nullptr; p->p3v_slots = nullptr; p->p3e_nbr = nullptr; p->p3e_slots = nullptr; p->p3_irregular = nullptr;
    p->p3g_nbr = nullptr; p->p3g_slots = nullptr; p->p3_general = false;
  };
#define P3_CHECK(expr)                                                              \
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
  P3_CHECK(cudaMalloc(&p->p3v_nbr, sizeof(int32_t) * (kRing * static_cast<size_t>(nn) + 128)));
  P3_CHECK(cudaMalloc(&p->p3v_slots, sizeof(uint32_t) * (kVertexSlotWords * static_cast<size_t>(nn) + 128)));
  P3_CHECK(cudaMalloc(&p->p3e_nbr, sizeof(int32_t) * (4 * static_cast<size_t>(ner) + 128)));
  P3_CHECK(cudaMalloc(&p->p3e_slots, sizeof(uint32_t) * (kEdgeSlotWords * static_cast<size_t>(ner) + 128)));
  P3_CHECK(cudaMalloc(&p->p3c_slots, sizeof(uint2) * (static_cast<size_t>(p->n_cells) + 128)));
  k_p3_cell_plan<<<static_cast<unsigned>(cdiv(p->n_cells, 256)), 256, 0, st>>>(p->n_cells, p->o_stride, p->pos_row,
                                                                                static_cast<const uint8_t*>(p->pos), static_cast<uint2*>(p->p3c_slots));
  ctx->launches++;
  k_p3_vertex_plan<<<static_cast<unsigned>(cdiv(nn, 128)), 128, 0, st>>>(nn, p->o_stride, p->pos_row, p->adj_ptr, p->adj, mesh->cell_nodes,
                                                                          static_cast<const uint8_t*>(p->pos), p->outer, p->p3v_nbr,
                                                                          p->p3v_slots, flag, cc);
  ctx->launches++;
  k_p3_edge_plan<<<static_cast<unsigned>(cdiv(ner, 128)), 128, 0, st>>>(nn, ner, p->o_stride, p->pos_row, p->adj_ptr, p->adj, mesh->cell_nodes,
                                                                         static_cast<const uint8_t*>(p->pos), p->outer, p->p3e_nbr,
                                                                         p->p3e_slots, flag, cc);
  ctx->launches++;
  P3_CHECK(cudaGetLastError());
  // unstructured meshes, on request (LFGPU_P3_GENERAL=1; core checked on the CPU, wrapper not yet run on a B200): the plan
  // for closed rings of 3..8 cells replaces the valence-6 plan of the vertex rows
  // automatic (default): the general plan when more than 5 % of the vertex rows miss the valence-6 plan (Gmsh / Delaunay meshes:
  // measured on workload u2, 1.0e6 triangles: 0.297 -> 0.153 ms); LFGPU_P3_GENERAL=1 forces it, =0 never builds it
  static const int general_env = [] { const char* e = std::getenv("LFGPU_P3_GENERAL"); return e == nullptr ? -1 : (e[0] == '1' ? 1 : 0); }();
  if (general_env != 0 && cc == 0) {
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
    P3_CHECK(count_flags(flag, &cnt6));
    if (general_env == 1 || static_cast<int64_t>(cnt6) * 20 > nn) {
      // candidate: plan the vertex rows for closed rings of 3..8 cells; adopted if it takes more than 5 % of the vertex rows off
      // the generic kernel (boundary rows fail both plans, so a small structured mesh keeps the leaner valence-6 kernel)
      uint8_t* flag_g = nullptr;
      P3_CHECK(cudaMalloc(&flag_g, nn));
      cudaError_t eg = cudaMalloc(&p->p3g_nbr, sizeof(int32_t) * (kMaxRing * static_cast<size_t>(nn) + 128));
      if (eg == cudaSuccess) eg = cudaMalloc(&p->p3g_slots, sizeof(uint32_t) * (kGeneralSlotWords * static_cast<size_t>(nn) + 128));
      if (eg == cudaSuccess) {
        k_p3_vertex_plan_general<<<static_cast<unsigned>(cdiv(nn, 128)), 128, 0, st>>>(nn, p->o_stride, p->pos_row, p->adj_ptr, p->adj, mesh->cell_nodes,
                                                                         static_cast<const uint8_t*>(p->pos), p->outer, p->p3g_nbr, p->p3g_slots, flag_g);
        ctx->launches++;
        eg = cudaGetLastError();
      }
      if (eg == cudaSuccess) eg = count_flags(flag_g, &cntg);
      const bool adopt = eg == cudaSuccess && (general_env == 1 || static_cast<int64_t>(cnt6 - cntg) * 20 > nn);
      if (adopt) eg = cudaMemcpyAsync(flag, flag_g, nn, cudaMemcpyDeviceToDevice, st);
      if (eg == cudaSuccess) eg = cudaStreamSynchronize(st);
      cudaFree(flag_g);
      if (!adopt || eg != cudaSuccess) {
        cudaFree(p->p3g_nbr);
        cudaFree(p->p3g_slots);
        p->p3g_nbr = nullptr;
        p->p3g_slots = nullptr;
      }
      P3_CHECK(eg);
      p->p3_general = adopt;
    }
  }
  P3_CHECK(cudaMalloc(&iota, sizeof(int32_t) * p->n_outer));
  P3_CHECK(cudaMalloc(&d_num, sizeof(int64_t)));
  cub::CountingInputIterator<int32_t> count_it(0);
  size_t tb = 0;
  if (and_row_keep(ctx, p->n_outer, flag, p->row_keep) != LFGPU_OK) return LFGPU_ERR_CUDA;  // rows nobody asks for need no generic kernel
  cub::DeviceSelect::Flagged(nullptr, tb, count_it, flag, iota, d_num, p->n_outer, st);
  P3_CHECK(cudaMalloc(&tmp, tb));
  P3_CHECK(cub::DeviceSelect::Flagged(tmp, tb, count_it, flag, iota, d_num, p->n_outer, st));
  int64_t n_irr = 0;
  P3_CHECK(cudaMemcpyAsync(&n_irr, d_num, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  P3_CHECK(cudaStreamSynchronize(st));
  if (n_irr > 0) {
    P3_CHECK(cudaMalloc(&p->p3_irregular, sizeof(int32_t) * n_irr));
    P3_CHECK(cudaMemcpyAsync(p->p3_irregular, iota, sizeof(int32_t) * n_irr, cudaMemcpyDeviceToDevice, st));
    p->p3_irregular_host.resize(static_cast<size_t>(n_irr));  // ascending; lets a row range find its share of the list
    P3_CHECK(cudaMemcpyAsync(p->p3_irregular_host.data(), iota, sizeof(int32_t) * n_irr, cudaMemcpyDeviceToHost, st));
    P3_CHECK(cudaStreamSynchronize(st));
  }
  // the edge rows' own copy of the node positions, in the order the rows use them (plan_dict.cu: edge_node_order; LFGPU_EDGE_ORDER=0
  // keeps the mesh's array): on the builder's numbering the edge-row kernel is 30 % (P2) / 11 % (P3) faster with it
  static const bool order_env = [] { const char* e = std::getenv("LFGPU_EDGE_ORDER"); return e == nullptr || e[0] != '0'; }();
  if (order_env && cc == 0 && n_irr * 2 <= p->n_outer) {
    if (edge_node_order(ctx, nn, ner, p->p3e_nbr, &p->p3e_newid) != LFGPU_OK) {
      cleanup();
      drop_plan();
      return LFGPU_ERR_CUDA;
    }
    if (p->p3e_newid != nullptr) {
      P3_CHECK(cudaMalloc(&p->p3e_xy, sizeof(double) * (2 * static_cast<size_t>(nn) + 32)));
      p->p3e_xy_version = 0;
      p->p3e_xy_mesh = nullptr;
    }
  }
#undef P3_CHECK
  cleanup();
  p->n_p3_irregular = n_irr;
  p->p3_nn = nn;
  if (n_irr * 2 > p->n_outer) {  // mostly irregular rows: the item kernel is the better choice
    drop_plan();
    return LFGPU_OK;
  }
  p->p3_state = 1;
  return LFGPU_OK;
}

// k00 .. km: the reference tensors of FeLagrangeO3Tria for the rule in use, each [10 * 10] row-major (assemble.cu: pack_type)
// rows [r0, r1) of the matrix (the caller has already sent the irregular rows of the range through the generic kernel)
int p3_rows_launch(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* p, const double alpha[4], int tensor, double gamma,
                   const double* k00, const double* k01, const double* k10, const double* k11, const double* km, double* d_values,
                   int64_t r0, int64_t r1, double beta) {
  Params P;
  P.a00 = alpha[0]; P.a01 = tensor ? alpha[1] : 0.0; P.a10 = tensor ? alpha[2] : 0.0; P.a11 = tensor ? alpha[3] : alpha[0];
  P.gamma = gamma;
  const bool simple = !tensor && gamma == 0.0;
  const int rows[3] = {0, 3, 9};
  for (int w = 0; w < 3; ++w)
    for (int b = 0; b < 10; ++b) {
      const int i = rows[w] * 10 + b;
      P.k00[w][b] = k00[i]; P.k01[w][b] = simple ? k01[i] + k10[i] : k01[i]; P.k10[w][b] = k10[i]; P.k11[w][b] = k11[i]; P.km[w][b] = km[i];
    }
  const int threads = 128;
  const int nn = static_cast<int>(p->p3_nn), nc = static_cast<int>(p->n_cells);
  const int base_int = static_cast<int>(p->n_outer - p->n_cells), ner = base_int - nn;
  static const int pfd_env = [] { const char* e = std::getenv("LFGPU_P3_PFD"); return e != nullptr ? std::atoi(e) : 100; }();
  const int ipf_v = static_cast<int>((static_cast<int64_t>(ctx->sm_count) * 3 * threads * pfd_env / 100) & ~static_cast<int64_t>(127));
  const int ipf_e = static_cast<int>((static_cast<int64_t>(ctx->sm_count) * 8 * threads * pfd_env / 100) & ~static_cast<int64_t>(127));  // one wave at 8 CTAs per SM
  // resident CTAs per SM of the edge-dof kernel: 8 (64 registers) measured best on config C4 -- 4 (79 registers, 6 CTAs fit): 0.463 ms,
  // 7: 0.435, 8: 0.423; LFGPU_P3_EOCC=4 / 10 select the other instantiations
  static const int eocc_env = [] { const char* e = std::getenv("LFGPU_P3_EOCC"); return e != nullptr ? std::atoi(e) : 8; }();
  static const int vocc_env = [] { const char* e = std::getenv("LFGPU_P3_VOCC"); return e != nullptr ? std::atoi(e) : 3; }();
  // percent of the plan prefetch distance; measured on config C4's kernel (round 2): 0 -> 0.504 ms, 50 -> 0.466, 100 -> 0.511, 200 -> 0.559
  static const int pfc_env = [] { const char* e = std::getenv("LFGPU_EDGE_PFC"); return e != nullptr ? std::atoi(e) : 50; }();
  const int ipc_e = pfc_env > 0 && ipf_e > 0 ? std::max(128, static_cast<int>((static_cast<int64_t>(ipf_e) * pfc_env / 100) & ~static_cast<int64_t>(127))) : 0;
  const size_t smem_v = sizeof(double) * (threads / 32) * 32 * (kVertexRowLen + 1);
  const size_t smem_e = sizeof(double) * (threads / 32) * 32 * kEdgeRowLen;
  const size_t smem_c = sizeof(double) * (threads / 32) * 32 * kCellRowLen;
  const int ipf_c = static_cast<int>((static_cast<int64_t>(ctx->sm_count) * 12 * threads * pfd_env / 100) & ~static_cast<int64_t>(127));
  // the share of the range in the vertex rows [0, nn), the edge-dof rows [nn, base_int) and the cell rows [base_int, N)
  auto clip = [](int64_t v, int64_t lo, int64_t hi) { return static_cast<int>(std::min<int64_t>(std::max<int64_t>(v, lo), hi)); };
  const int v_first = clip(r0, 0, nn), v_end = clip(r1, 0, nn);
  const int e_first = clip(r0 - nn, 0, ner), e_end = clip(r1 - nn, 0, ner);
  const int c_first = clip(r0 - base_int, 0, nc), c_end = clip(r1 - base_int, 0, nc);
  const size_t smem_g = sizeof(double) * (threads / 32) * 32 * (kMaxVertexRowLen + 1);
  // meshes with per-cell corners: the CC instantiations read mesh->cell_coords (default occupancies only)
#define P3_LAUNCH_CC(MODE)                                                                                                                \
  if (v_end > v_first) {                                                                                                                  \
    k_p3_vertex_rows<MODE, 3, true><<<static_cast<unsigned>(cdiv(v_end - v_first, threads)), threads, smem_v, ctx->stream>>>(             \
        nn, p->p3v_nbr, p->p3v_slots, mesh->cell_coords, p->outer, ipf_v, P, d_values, v_first, v_end, beta);                             \
    LFGPU_LAUNCH_CHECK(ctx);                                                                                                              \
  }                                                                                                                                       \
  if (e_end > e_first) {                                                                                                                  \
    k_p3_edge_rows<MODE, 8, true><<<static_cast<unsigned>(cdiv(e_end - e_first, threads)), threads, smem_e, ctx->stream>>>(               \
        ner, nn, p->p3e_nbr, p->p3e_slots, mesh->cell_coords, p->outer, ipf_e, ipc_e, P, d_values, e_first, e_end, beta);                 \
    LFGPU_LAUNCH_CHECK(ctx);                                                                                                              \
  }                                                                                                                                       \
  if (c_end > c_first) {                                                                                                                  \
    k_p3_cell_rows<MODE, true><<<static_cast<unsigned>(cdiv(c_end - c_first, threads)), threads, smem_c, ctx->stream>>>(                  \
        base_int, mesh->cell_nodes, static_cast<const uint2*>(p->p3c_slots), mesh->cell_coords, p->outer, ipf_c, P, d_values, c_first, c_end, beta); \
    LFGPU_LAUNCH_CHECK(ctx);                                                                                                              \
  }
  const double* exy = mesh->node_coords;
  if (!p->p3_cc && p->p3e_newid != nullptr && e_end > e_first) {  // the edge plan's node numbers are positions in the rows' own copy
    auto* pm = const_cast<lfgpu_pattern*>(p);
    if (pm->p3e_xy_version != mesh->coords_version || pm->p3e_xy_mesh != mesh) {
      const int rc = permute_node_coords(ctx, nn, p->p3e_newid, mesh->node_coords, pm->p3e_xy);
      if (rc != LFGPU_OK) return rc;
      pm->p3e_xy_version = mesh->coords_version;
      pm->p3e_xy_mesh = mesh;
    }
    exy = p->p3e_xy;
  }
#define P3_LAUNCH(MODE)                                                                                                                   \
  if (v_end > v_first && p->p3_general) {                                                                                                 \
    /* 51 200 bytes of stage: above the 48 KB a kernel gets without asking */                                                             \
    LFGPU_CUDA_CHECK(ctx, cudaFuncSetAttribute(k_p3_vertex_rows_general<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,               \
                                               static_cast<int>(smem_g)));                                                                \
    k_p3_vertex_rows_general<MODE><<<static_cast<unsigned>(cdiv(v_end - v_first, threads)), threads, smem_g, ctx->stream>>>(              \
        v_first, v_end, nn, p->p3g_nbr, p->p3g_slots, mesh->node_coords, p->outer, P, d_values, beta);                                          \
    LFGPU_LAUNCH_CHECK(ctx);                                                                                                              \
  } else if (v_end > v_first) {                                                                                                           \
    if (MODE == 1 && vocc_env == 4)                                                                                                       \
      k_p3_vertex_rows<1, 4><<<static_cast<unsigned>(cdiv(v_end - v_first, threads)), threads, smem_v, ctx->stream>>>(                    \
          nn, p->p3v_nbr, p->p3v_slots, mesh->node_coords, p->outer, ipf_v * 4 / 3, P, d_values, v_first, v_end, beta);                         \
    else                                                                                                                                  \
      k_p3_vertex_rows<MODE><<<static_cast<unsigned>(cdiv(v_end - v_first, threads)), threads, smem_v, ctx->stream>>>(                    \
          nn, p->p3v_nbr, p->p3v_slots, mesh->node_coords, p->outer, ipf_v, P, d_values, v_first, v_end, beta);                                 \
    LFGPU_LAUNCH_CHECK(ctx);                                                                                                              \
  }                                                                                                                                       \
  if (e_end > e_first) {                                                                                                                  \
    auto ke = eocc_env == 10 ? k_p3_edge_rows<MODE, 10> : (eocc_env == 4 ? k_p3_edge_rows<MODE, 4> : k_p3_edge_rows<MODE, 8>);             \
    ke<<<static_cast<unsigned>(cdiv(e_end - e_first, threads)), threads, smem_e, ctx->stream>>>(                                          \
        ner, nn, p->p3e_nbr, p->p3e_slots, exy, p->outer, ipf_e, ipc_e, P, d_values, e_first, e_end, beta);                               \
    LFGPU_LAUNCH_CHECK(ctx);                                                                                                              \
  }                                                                                                                                       \
  if (c_end > c_first) {                                                                                                                  \
    k_p3_cell_rows<MODE><<<static_cast<unsigned>(cdiv(c_end - c_first, threads)), threads, smem_c, ctx->stream>>>(                        \
        base_int, mesh->cell_nodes, static_cast<const uint2*>(p->p3c_slots), mesh->node_coords, p->outer, ipf_c, P, d_values, c_first, c_end, beta);               \
    LFGPU_LAUNCH_CHECK(ctx);                                                                                                              \
  }
  if (p->p3_cc) {
    if (mesh->cell_coords == nullptr) LFGPU_FAIL(ctx, LFGPU_ERR_INVALID, "the P3 plan was built for a mesh with cell corners");
    if (simple) {
      P3_LAUNCH_CC(0)
    } else {
      P3_LAUNCH_CC(1)
    }
  } else if (simple) {
    P3_LAUNCH(0)
  } else {
    P3_LAUNCH(1)
  }
#undef P3_LAUNCH
#undef P3_LAUNCH_CC
  return LFGPU_OK;
}

}  // namespace lfgpu
