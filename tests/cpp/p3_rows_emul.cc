// Host emulation of the P3 row kernels (test infrastructure): runs the SAME plan and row functions the CUDA kernels of
// lehrfempp_b200/csrc/assemble_p3.cu call (rows_p3_core.h, compiled here with g++), on the arrays the symbolic pass would
// hold (gather lists, scatter map), so that index logic and arithmetic can be compared with the oracle without a GPU.
#include <algorithm>
#include <cstdint>
#include <limits>
#include <vector>

#include "../../lehrfempp_b200/csrc/rows_p3_core.h"

using namespace lfgpu::p3;

namespace {
bool g_general = false;  // vertex rows through the general-valence functions
const double* g_cell_xy = nullptr;  // [n_cells][4][2]: corners per cell (the kernels' cell_coords mode), else node positions

struct HostCellVec {
  const double* cell_xy;
  const uint32_t* cw;
  void operator()(int k, double& ax, double& ay, double& bx, double& by) const {
    const double* c = cell_xy + 8 * static_cast<size_t>(cw[k] >> 4);
    const int ia = (cw[k] >> 2) & 3, ib = cw[k] & 3, i0 = 3 - ia - ib;
    ax = c[2 * ia] - c[2 * i0]; ay = c[2 * ia + 1] - c[2 * i0 + 1];
    bx = c[2 * ib] - c[2 * i0]; by = c[2 * ib + 1] - c[2 * i0 + 1];
  }
};
template <int MODE>
void run(const Params& P, int64_t n_nodes, int64_t n_cells, const uint32_t* cell_nodes, const double* xy, int stride, int pos_row,
         const std::vector<int64_t>& adj_ptr, const std::vector<uint32_t>& adj, const std::vector<uint8_t>& pos, int64_t n_dofs,
         const int32_t* outer, double* values, uint8_t* regular, int64_t* counts) {
  const int64_t base_int = n_dofs - n_cells;
  for (int64_t r = 0; r < n_dofs; ++r) {
    const int m = static_cast<int>(adj_ptr[r + 1] - adj_ptr[r]);
    const uint32_t* items = adj.data() + adj_ptr[r];
    const int len = outer[r + 1] - outer[r];
    double* dst = values + outer[r];
    regular[r] = 0;
    if (r < n_nodes) {
      if (g_general) {  // closed rings of 3..8 cells
        int32_t ring[kMaxRing];
        uint32_t w[kGeneralSlotWords];
        if (!vertex_plan_general(r, m, items, cell_nodes, pos.data(), stride, pos_row, len, ring, w)) continue;
        double dx[kMaxRing], dy[kMaxRing];
        for (int k = 0; k < kMaxRing; ++k) {
          const int64_t n = ring[k] >= 0 ? ring[k] : r;
          dx[k] = xy[2 * n] - xy[2 * r];
          dy[k] = xy[2 * n + 1] - xy[2 * r + 1];
        }
        vertex_row_general<MODE>(P, dx, dy, w, dst);
      } else if (g_cell_xy != nullptr) {
        int32_t ring[kRing];
        uint32_t w[kVertexSlotWords], cw[kRing];
        if (!vertex_plan(r, m, items, cell_nodes, pos.data(), stride, pos_row, len, ring, w, cw)) continue;
        vertex_row_cv<MODE>(P, HostCellVec{g_cell_xy, cw}, w, dst);
      } else {
        int32_t ring[kRing];
        uint32_t w[kVertexSlotWords];
        if (!vertex_plan(r, m, items, cell_nodes, pos.data(), stride, pos_row, len, ring, w)) continue;
        double dx[kRing], dy[kRing];
        for (int k = 0; k < kRing; ++k) {
          dx[k] = xy[2 * ring[k]] - xy[2 * r];
          dy[k] = xy[2 * ring[k] + 1] - xy[2 * r + 1];
        }
        vertex_row<MODE>(P, dx, dy, w, dst);
      }
      regular[r] = 1;
      counts[0]++;
    } else if (r < base_int) {
      int32_t ids[4];
      uint32_t w[kEdgeSlotWords], cw[2];
      if (!edge_plan(m, items, cell_nodes, pos.data(), stride, pos_row, len, ids, w, cw)) continue;
      if (g_cell_xy != nullptr) {
        double a1x, a1y, b1x, b1y, a2x, a2y, b2x, b2y;
        const HostCellVec cv{g_cell_xy, cw};
        cv(0, a1x, a1y, b1x, b1y);
        cv(1, a2x, a2y, b2x, b2y);
        edge_row2<MODE>(P, a1x, a1y, b1x, b1y, a2x, a2y, b2x, b2y, w, dst);
        regular[r] = 1;
        counts[1]++;
        continue;
      }
      const double px = xy[2 * ids[0]], py = xy[2 * ids[0] + 1];
      edge_row<MODE>(P, xy[2 * ids[1]] - px, xy[2 * ids[1] + 1] - py, xy[2 * ids[2]] - px, xy[2 * ids[2] + 1] - py, xy[2 * ids[3]] - px,
                     xy[2 * ids[3] + 1] - py, w, dst);
      regular[r] = 1;
      counts[1]++;
    } else {
      const int64_t c = r - base_int;
      if (m != 1 || (items[0] >> 4) != static_cast<uint32_t>(c) || (items[0] & 15U) != 9U || len != kCellRowLen) continue;
      const uint32_t* v = cell_nodes + 4 * c;
      const uint8_t* prow = pos.data() + (c * stride + 9) * static_cast<int64_t>(pos_row);
      uint32_t pw[3];
      for (int j = 0; j < 3; ++j)
        pw[j] = static_cast<uint32_t>(prow[4 * j]) | (static_cast<uint32_t>(prow[4 * j + 1]) << 8) |
                (static_cast<uint32_t>(prow[4 * j + 2]) << 16) | (static_cast<uint32_t>(prow[4 * j + 3]) << 24);
      if (g_cell_xy != nullptr) {
        const double* cx = g_cell_xy + 8 * static_cast<size_t>(c);
        cell_row<MODE>(P, cx[2] - cx[0], cx[3] - cx[1], cx[4] - cx[0], cx[5] - cx[1], pw, dst);
      } else {
        const double x0 = xy[2 * v[0]], y0 = xy[2 * v[0] + 1];
        cell_row<MODE>(P, xy[2 * v[1]] - x0, xy[2 * v[1] + 1] - y0, xy[2 * v[2]] - x0, xy[2 * v[2] + 1] - y0, pw, dst);
      }
      regular[r] = 1;
      counts[2]++;
    }
  }
}
}  // namespace

extern "C" void p3_rows_emulate_general(int on) { g_general = on != 0; }
extern "C" void p3_rows_emulate_cell_coords(const double* cell_xy) { g_cell_xy = cell_xy; }

extern "C" int p3_rows_emulate(int64_t n_nodes, int64_t n_cells, const uint32_t* cell_nodes, const double* node_xy, int stride,
                               const int32_t* dofs, int64_t n_dofs, const int32_t* outer, const int32_t* inner, const double* alpha4,
                               int tensor, double gamma, const double* k00, const double* k01, const double* k10, const double* k11,
                               const double* km, double* values, uint8_t* regular, int64_t* counts) {
  // gather lists: items (cell << 4 | list position) per dof, ascending in (cell, position)
  std::vector<int64_t> adj_ptr(n_dofs + 1, 0);
  for (int64_t c = 0; c < n_cells; ++c)
    for (int a = 0; a < 10; ++a) adj_ptr[dofs[c * stride + a] + 1]++;
  for (int64_t r = 0; r < n_dofs; ++r) adj_ptr[r + 1] += adj_ptr[r];
  std::vector<uint32_t> adj(adj_ptr[n_dofs]);
  std::vector<int64_t> fill(adj_ptr.begin(), adj_ptr.end() - 1);
  for (int64_t c = 0; c < n_cells; ++c)
    for (int a = 0; a < 10; ++a) adj[fill[dofs[c * stride + a]]++] = (static_cast<uint32_t>(c) << 4) | static_cast<uint32_t>(a);
  // scatter map (symbolic.cu: k_positions): slot of dof(c, b) inside the row of dof(c, a)
  const int pos_row = (stride + 3) & ~3;
  std::vector<uint8_t> pos(static_cast<size_t>(n_cells) * stride * pos_row, 255);
  for (int64_t c = 0; c < n_cells; ++c)
    for (int a = 0; a < 10; ++a) {
      const int32_t r = dofs[c * stride + a];
      const int32_t* b0 = inner + outer[r];
      const int32_t* b1 = inner + outer[r + 1];
      for (int b = 0; b < 10; ++b) {
        const int32_t* it = std::lower_bound(b0, b1, dofs[c * stride + b]);
        if (it == b1 || *it != dofs[c * stride + b]) return -1;
        pos[(c * stride + a) * static_cast<size_t>(pos_row) + b] = static_cast<uint8_t>(it - b0);
      }
    }
  Params P;
  P.a00 = alpha4[0]; P.a01 = tensor ? alpha4[1] : 0.0; P.a10 = tensor ? alpha4[2] : 0.0; P.a11 = tensor ? alpha4[3] : alpha4[0];
  P.gamma = gamma;
  const bool simple = !tensor && gamma == 0.0;
  const int rows[3] = {0, 3, 9};
  for (int w = 0; w < 3; ++w)
    for (int b = 0; b < 10; ++b) {
      const int i = rows[w] * 10 + b;
      P.k00[w][b] = k00[i]; P.k01[w][b] = simple ? k01[i] + k10[i] : k01[i]; P.k10[w][b] = k10[i]; P.k11[w][b] = k11[i]; P.km[w][b] = km[i];
    }
  const int64_t nnz = outer[n_dofs];
  std::fill(values, values + nnz, std::numeric_limits<double>::quiet_NaN());
  counts[0] = counts[1] = counts[2] = 0;
  if (simple) run<0>(P, n_nodes, n_cells, cell_nodes, node_xy, stride, pos_row, adj_ptr, adj, pos, n_dofs, outer, values, regular, counts);
  else run<1>(P, n_nodes, n_cells, cell_nodes, node_xy, stride, pos_row, adj_ptr, adj, pos, n_dofs, outer, values, regular, counts);
  return 0;
}
