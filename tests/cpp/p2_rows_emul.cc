// Host emulation of the general-valence P2 vertex rows (test infrastructure): the SAME plan and row functions the CUDA kernel
// k_p2_vertex_rows_general calls (lehrfempp_b200/csrc/rows_p2_core.h, compiled here with g++), on the arrays the symbolic
// pass would hold, so that they can be compared with the oracle without a GPU (tests/test_p2_rows_core.py).
#include <algorithm>
#include <cstdint>
#include <limits>
#include <vector>

#include "../../lehrfempp_b200/csrc/rows_p2_core.h"

using namespace lfgpu::p2;

extern "C" int p2_vertex_rows_emulate(int64_t n_nodes, int64_t n_cells, const uint32_t* cell_nodes, const double* xy, int stride,
                                      const int32_t* dofs, int64_t n_dofs, const int32_t* outer, const int32_t* inner,
                                      const double* alpha4, int tensor, double gamma, const double* k00, const double* k01,
                                      const double* k10, const double* k11, const double* km, double* values, uint8_t* regular,
                                      int64_t* ring_histogram /*[9]*/) {
  std::vector<int64_t> adj_ptr(n_dofs + 1, 0);
  for (int64_t c = 0; c < n_cells; ++c)
    for (int a = 0; a < 6; ++a) adj_ptr[dofs[c * stride + a] + 1]++;
  for (int64_t r = 0; r < n_dofs; ++r) adj_ptr[r + 1] += adj_ptr[r];
  std::vector<uint32_t> adj(adj_ptr[n_dofs]);
  std::vector<int64_t> fill(adj_ptr.begin(), adj_ptr.end() - 1);
  for (int64_t c = 0; c < n_cells; ++c)
    for (int a = 0; a < 6; ++a) adj[fill[dofs[c * stride + a]]++] = (static_cast<uint32_t>(c) << 4) | static_cast<uint32_t>(a);
  const int pos_row = (stride + 3) & ~3;
  std::vector<uint8_t> pos(static_cast<size_t>(n_cells) * stride * pos_row, 255);
  for (int64_t c = 0; c < n_cells; ++c)
    for (int a = 0; a < 6; ++a) {
      const int32_t r = dofs[c * stride + a];
      const int32_t* b0 = inner + outer[r];
      const int32_t* b1 = inner + outer[r + 1];
      for (int b = 0; b < 6; ++b) {
        const int32_t* it = std::lower_bound(b0, b1, dofs[c * stride + b]);
        if (it == b1 || *it != dofs[c * stride + b]) return -1;
        pos[(c * stride + a) * static_cast<size_t>(pos_row) + b] = static_cast<uint8_t>(it - b0);
      }
    }
  VertexParams P;
  P.a00 = alpha4[0]; P.a01 = tensor ? alpha4[1] : 0.0; P.a10 = tensor ? alpha4[2] : 0.0; P.a11 = tensor ? alpha4[3] : alpha4[0];
  P.gamma = gamma;
  const bool simple = !tensor && gamma == 0.0;
  for (int b = 0; b < 6; ++b) {
    P.k00[b] = k00[b]; P.k01[b] = simple ? k01[b] + k10[b] : k01[b]; P.k10[b] = k10[b]; P.k11[b] = k11[b]; P.km[b] = km[b];
  }
  std::fill(values, values + outer[n_dofs], std::numeric_limits<double>::quiet_NaN());
  for (int k = 0; k < 9; ++k) ring_histogram[k] = 0;
  for (int64_t r = 0; r < n_dofs; ++r) regular[r] = 0;
  for (int64_t r = 0; r < n_nodes; ++r) {
    const int m = static_cast<int>(adj_ptr[r + 1] - adj_ptr[r]);
    int32_t ring[kMaxRing];
    uint32_t w[kSlotWords];
    if (!vertex_plan_general(r, m, adj.data() + adj_ptr[r], cell_nodes, pos.data(), stride, pos_row, outer[r + 1] - outer[r], ring, w)) continue;
    double dx[kMaxRing], dy[kMaxRing];
    for (int k = 0; k < kMaxRing; ++k) {
      const int64_t n = ring[k] >= 0 ? ring[k] : r;
      dx[k] = xy[2 * n] - xy[2 * r];
      dy[k] = xy[2 * n + 1] - xy[2 * r + 1];
    }
    if (simple) vertex_row_general<0>(P, dx, dy, w, values + outer[r]);
    else vertex_row_general<1>(P, dx, dy, w, values + outer[r]);
    regular[r] = 1;
    ring_histogram[m]++;
  }
  return 0;
}
