// Internal declarations of liblfgpu.so (product code; self-contained, no test infrastructure is included).
#ifndef LFGPU_INTERNAL_CUH
#define LFGPU_INTERNAL_CUH

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/lfgpu.h"

namespace lfgpu {

constexpr int kMaxNsf = 16;  // FeLagrangeO3Quad
constexpr int kMaxNq = 36;         // largest user quadrature rule the tables hold (6x6 Gauss / 33-point triangle rule)
constexpr int kItemThreads = 256;  // block size of the item-parallel kernel = max items of one block of the plan

void set_last_error(const lfgpu_ctx* ctx, const std::string& msg);

}  // namespace lfgpu

struct lfgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int64_t launches = 0;
  std::string last_error;
  void* d_scratch = nullptr;  // small device scratch (flags, counters)
  double* nodal_tab[2] = {nullptr, nullptr};  // per-point tables of node-interpolated coefficients (assemble.cu: resolve_nodal)
  size_t nodal_cap[2] = {0, 0};
  bool geom_check_pending = false;  // a coordinate update queued a degeneracy check whose flag (scratch + 1024) is read at the next synchronize
  struct TableEntry {
    std::vector<double> host;
    double* dev;
  };
  std::vector<TableEntry> table_cache;  // reference-element tables already on the device (assemble.cu)
  // copy streams + events of the host pipeline (hostpipe.cu), created on first use
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  std::vector<cudaEvent_t> pipe_events;
};

struct lfgpu_mesh {
  lfgpu_ctx* ctx = nullptr;
  int64_t n_nodes = 0, n_cells = 0, n_edges = 0, n_tria = 0, n_quad = 0;
  double* node_coords = nullptr;    // [n_nodes][2]
  uint32_t* cell_nodes = nullptr;   // [n_cells][4], NIL padded
  double* cell_coords = nullptr;    // [n_cells][4][2] or nullptr (corners == node positions)
  uint64_t coords_version = 1;      // bumped by every call that rewrites node_coords (the row-kernel plans keep reordered copies)
  // topology (optional, built by lfgpu_mesh_build_topology / tp generators)
  bool has_topology = false;
  uint32_t tp_nx = 0, tp_ny = 0;    // > 0: mesh of the triangle builder, whose explicit edge list is generated on demand
  uint32_t* edge_nodes = nullptr;   // [n_edges][2] (first, second endpoint)
  uint32_t* cell_edges = nullptr;   // [n_cells][4]
  int8_t* cell_edge_ori = nullptr;  // [n_cells][4]
  // cell-interior dof numbering helper: exclusive prefix counts of triangles / quads before each cell is not stored;
  // interior dofs are numbered in cell order across both types (dofhandler.cc:263-281), see dofs.cu
};

struct lfgpu_dofmap {
  lfgpu_ctx* ctx = nullptr;
  int64_t n_cells = 0, n_dofs = 0;
  int stride = 0;
  int32_t* cell_dofs = nullptr;  // [n_cells][stride], -1 padded (int32: the compressed matrix uses int32 indices)
  uint8_t* n_ldof = nullptr;     // [n_cells]
  int max_ldof = 0;
  // uniform layouts built on the device remember their per-entity counts (dofs of node v: v * n_pt + j; interior dofs of
  // edge e: n_nodes * n_pt + e * n_seg + j); -1 for uploaded tables, whose edge dofs are unknown
  int n_pt = -1, n_seg = -1;
  int64_t n_nodes = 0;
  // gather plan of the load vector (dofs.cu: dofmap_gather_plan), built on first use: for dof r the items
  // g_items[g_ptr[r] .. g_ptr[r+1]) = (cell << 4 | local index), ascending in cell index = the order in which
  // AssembleVectorLocally adds to result[r] (assembler.h:322-324)
  int g_state = 0;  // 0 = not built, 1 = ready
  int32_t* g_ptr = nullptr;
  uint32_t* g_items = nullptr;
  int64_t g_n_items = 0;
  // vertex-ring plan of the P1 load vector (assemble_p1.cu: p1_load_fan), built on first use: 0 = not tried, 1 = ready, -1 = n/a
  int lv_state = 0, lv_w = 0;
  uint32_t* lv_nbr = nullptr;      // [lv_w][n_dofs] ring node ids (0xFFFFFFFF = empty), or
  int16_t* lv_nbr16 = nullptr;     // [6][n_dofs] ring ids as offsets from the row id (0 = empty)
  uint8_t* lv_info = nullptr;      // [n_dofs] 0 open fan, 1 closed fan, 2 not a single fan (generic kernel), 3 no cells
  uint32_t* lv_cells = nullptr;    // [lv_w][n_dofs] cell << 2 | local index of the row's node, per ring position (tabulated sources)
  int32_t* lv_irregular = nullptr;
  int64_t n_lv_irregular = 0;
  // two-pass load vector (assemble.cu: k_load_positions): lv_pos [n_cells][lv_ev_stride] = index of (cell, a) in g_items,
  // lv_ev [number of items] = the element-vector entries in item order
  double* lv_ev = nullptr;
  int32_t* lv_pos = nullptr;
  int lv_ev_stride = 0;
};

struct lfgpu_pattern {
  lfgpu_ctx* ctx = nullptr;
  int major = LFGPU_ROW_MAJOR;
  int64_t n_outer = 0, n_inner = 0, nnz = 0, n_cells = 0;
  int32_t* outer = nullptr;  // [n_outer + 1]
  int32_t* inner = nullptr;  // [nnz]
  // gather plan: for outer index r the items adj[adj_ptr[r] .. adj_ptr[r+1]) = (cell << 4 | local outer index a),
  // ascending in cell index (the reference's summation order)
  int32_t* adj_ptr = nullptr;   // [n_outer + 1]
  uint32_t* adj = nullptr;      // [n_items]
  int64_t n_items = 0;
  // scatter map: position of inner dof b inside the outer segment of outer dof a, for every cell:
  // pos[(cell * o_stride + a) * pos_row + b]; uint8 when max_row_len <= 256 else uint16
  int o_stride = 0, i_stride = 0, pos_row = 0, pos_bytes = 1;
  void* pos = nullptr;
  int max_row_len = 0;
  int max_block_nnz = 0;  // max over blocks of 128 consecutive outer indices of their number of stored values
  // item-parallel gather plan: block b of the item kernel owns the outer indices [blk_rows[b], blk_rows[b+1]), chosen so
  // that it has at most 256 items; pos_item = scatter slots re-ordered to item order (streamed, not gathered)
  int64_t n_item_blocks = 0;
  int32_t* blk_rows = nullptr;   // [n_item_blocks + 1]
  void* pos_item = nullptr;      // [n_items][pos_row], same element type as pos
  void* blk_hdr = nullptr;       // int4 [n_item_blocks]: first item, first value, n_items | max rank << 16, number of values
  void* item_sorted = nullptr;   // uint2 [n_items] in thread order: (cell << 4 | a, image offset of the dof | rank << 16)
  uint32_t* item_perm = nullptr; // [n_items] per block: thread t -> local item | local row << 8 | rank-in-row << 16, sorted by (rank, row)
  int max_item_block_nnz = 0;
  double* cell_metric = nullptr; // [n_cells][6] scratch of the numeric pass (assemble.cu: k_cell_metric)
  int max_items = 0;  // max number of cells adjacent to one outer dof
  // rows a later numeric pass has to produce (lfgpu_pattern_restrict_rows; null = all): the row kernels do not send the other
  // rows through the generic kernel when their plan does not cover them (halo rows of a distributed sub-problem)
  uint8_t* row_keep = nullptr;  // [n_outer]
  // dof tables the plan was built from (device copies owned by the pattern)
  int32_t* o_dofs = nullptr;  // [n_cells][o_stride]
  int32_t* i_dofs = nullptr;  // [n_cells][i_stride]
  uint8_t* o_nldof = nullptr;
  uint8_t* i_nldof = nullptr;
  // P1 vertex-fan plan (assemble_p1.cu), built on first use: 0 = not tried, 1 = ready, -1 = not applicable
  int fan_state = 0;
  int fan_w = 0;                     // ring slots per row
  uint32_t* fan_nbr = nullptr;       // [fan_w][n_outer] neighbour ring of every row, slot-major: node id | slot-in-row << 28
  uint8_t* fan_rowinfo = nullptr;    // [n_outer] slot of the diagonal | closed-fan flag << 7
  // compact form of the same plan (rings of <= 6 neighbours whose ids are within +-32767 of the row): fan_nbr is freed
  int16_t* fan_nbr16 = nullptr;      // [6][n_outer] neighbour id - row id, 0 = empty slot
  uint32_t* fan_info32 = nullptr;    // [n_outer] 4 bits slot-in-row per ring position | diagonal slot << 24 | closed << 28
  int32_t* fan_irregular = nullptr;  // rows that are not a single fan (generic kernel)
  int64_t n_irregular = 0;
  // P1 row-kernel plan for hybrid meshes / variable coefficients (assemble_p1h.cu), built on first use
  int p1h_state = 0;
  int p1h_kq = 0, p1h_kt = 0;        // quadrilateral / triangle item slots per row
  uint32_t* p1h_qw = nullptr;        // [4 * kq][n_outer] cell | rot << 28, then three corners node | slot << 28
  uint32_t* p1h_tw = nullptr;        // [3 * kt][n_outer]
  uint8_t* p1h_rowinfo = nullptr;    // [n_outer] slot of the diagonal; 0xFF = generic kernel, 0xFE = no cells
  int32_t* p1h_irregular = nullptr;
  int64_t n_p1h_irregular = 0;
  std::vector<int32_t> p1h_irregular_host;
  // P2 row-kernel plan (assemble_p2.cu), built on first use: 0 = not tried, 1 = ready, -1 = not applicable
  int p2_state = 0;
  int64_t p2_nn = 0;                 // number of vertex rows (= mesh nodes); the edge rows follow
  int32_t* p2v_nbr = nullptr;        // [6][p2_nn] ring of neighbour nodes of every vertex row (-1 in slot 0: not a regular row)
  uint32_t* p2v_slots = nullptr;     // [3][p2_nn] 6 x 5 bits each: slots of the neighbour / spoke-edge / rim-edge columns
  int32_t* p2e_nbr = nullptr;        // [4][n_edges] endpoints p, q and opposite vertices o_1, o_2 of every edge row
  uint32_t* p2e_slots = nullptr;     // [n_edges] 8 x 4 bits: slots of p, q, o_1, o_2, (q,o_1), (o_1,p), (q,o_2), (o_2,p)
  // compact form (assemble_p2.cu "compact plan"): p2v_nbr = uint32 [3][p2_nn] 16-bit differences, p2v_slots = uint4 table of slot
  // triples, p2v_cidx = table index per row (0xFFFF: not planned); p2e_nbr = uint32 [3][n_edges], p2e_slots = uint4 table
  bool p2_compact_v = false, p2_compact_e = false;
  uint16_t* p2v_cidx = nullptr;
  // edge rows: the kernels' own copy of the node positions in the order the edge plan uses them (plan_dict.cu: edge_node_order);
  // p2e_nbr then holds positions in that copy.  Refreshed when the mesh's coords_version moves.
  uint32_t* p2e_newid = nullptr;     // [p2_nn]
  double* p2e_xy = nullptr;          // [p2_nn][2]
  uint64_t p2e_xy_version = 0;
  const void* p2e_xy_mesh = nullptr;
  bool p2_cc = false;                // plan built for a mesh with per-cell corners: p2v_nbr / p2e_nbr hold (cell, corner) words
  bool p2_general = false;           // vertex rows planned for closed rings of 3..8 cells (rows_p2_core.h) instead of exactly 6
  int32_t* p2g_nbr = nullptr;        // [8][p2_nn]
  uint32_t* p2g_slots = nullptr;     // [6][p2_nn]
  int32_t* p2_irregular = nullptr;   // rows left to the generic gather kernel (ascending)
  int64_t n_p2_irregular = 0;
  std::vector<int32_t> p2_irregular_host;  // host copy: a row range looks up its share of the list
  // P3 row-kernel plan (assemble_p3.cu, rows_p3_core.h): vertex rows [0, p3_nn), edge-dof rows, then one row per cell
  int p3_state = 0;
  int64_t p3_nn = 0;
  int32_t* p3v_nbr = nullptr;        // [6][p3_nn] ring of neighbour nodes (-1 in slot 0: not a regular row)
  uint32_t* p3v_slots = nullptr;     // [9][p3_nn] 36 slot bytes per row
  int32_t* p3e_nbr = nullptr;        // [4][n_edge_rows] P, Q, o_1, o_2
  uint32_t* p3e_slots = nullptr;     // [2][n_edge_rows] 16 slot nibbles per row
  void* p3c_slots = nullptr;         // uint2 [n_cells] slots of the ten list positions in the cell's own row, one nibble each
  uint32_t* p3e_newid = nullptr;     // as p2e_newid / p2e_xy for the P3 edge-dof rows
  double* p3e_xy = nullptr;
  uint64_t p3e_xy_version = 0;
  const void* p3e_xy_mesh = nullptr;
  bool p3_cc = false;                // plan built for a mesh with per-cell corners: p3v_nbr / p3e_nbr hold (cell, corner) words
  bool p3_general = false;           // vertex rows planned for closed rings of 3..8 cells instead of exactly 6
  int32_t* p3g_nbr = nullptr;        // [8][p3_nn]
  uint32_t* p3g_slots = nullptr;     // [13][p3_nn]
  int32_t* p3_irregular = nullptr;
  int64_t n_p3_irregular = 0;
  std::vector<int32_t> p3_irregular_host;
  // host pipeline plan (hostpipe.cu): for hp_blocks equal blocks of outer indices, the number of leading node coordinates
  // that must be on the device before block b can be computed (running maximum, so monotone)
  int hp_blocks = 0;
  int64_t hp_row0 = -1, hp_rows = -1;  // the row range the cached plan was built for
  int64_t hp_lo = 0;                   // smallest node index the range refers to
  std::vector<int64_t> hp_need;
  std::vector<int64_t> hp_val;  // [hp_blocks + 1] first stored value of every block
};

namespace lfgpu {

#define LFGPU_CUDA_CHECK(ctx, expr)                                                                     \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      ::lfgpu::set_last_error(ctx, std::string(#expr) + ": " + cudaGetErrorString(_e));                 \
      return LFGPU_ERR_CUDA;                                                                            \
    }                                                                                                   \
  } while (0)

#define LFGPU_FAIL(ctx, code, msg)        \
  do {                                    \
    ::lfgpu::set_last_error(ctx, msg);    \
    return code;                          \
  } while (0)

#define LFGPU_LAUNCH_CHECK(ctx)                                                         \
  do {                                                                                  \
    (ctx)->launches++;                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      ::lfgpu::set_last_error(ctx, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
      return LFGPU_ERR_CUDA;                                                            \
    }                                                                                   \
  } while (0)

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- reference-element tables (host side, fe_tables.cpp) -----------------------------------------------------------
struct FeTable {
  int nsf = 0, nq = 0;
  double w[kMaxNq];
  double qx[kMaxNq], qy[kMaxNq];
  double phi[kMaxNsf * kMaxNq];  // phi[a * nq + k]
  double gx[kMaxNsf * kMaxNq];   // d/dx0
  double gy[kMaxNsf * kMaxNq];   // d/dx1
};
// reference tensors for affine cells with cell-wise constant coefficients:
// khat[i][j][a * nsf + b] = sum_k w_k d_i phi_a(k) d_j phi_b(k),  mhat[a * nsf + b] = sum_k w_k phi_a phi_b,
// lhat[a] = sum_k w_k phi_a
struct FeTensors {
  double k00[kMaxNsf * kMaxNsf], k01[kMaxNsf * kMaxNsf], k10[kMaxNsf * kMaxNsf], k11[kMaxNsf * kMaxNsf];
  double m[kMaxNsf * kMaxNsf];
  double l[kMaxNsf];
};

// FeLagrangeO{1,2,3}Segment at the points of a rule on [0,1] (edges.cu); small enough to travel as a kernel argument
constexpr int kMaxSegNq = 16;
struct SegTable {
  int nsf = 0, nq = 0;
  double w[kMaxSegNq];
  double x[kMaxSegNq];
  double phi[4 * kMaxSegNq];  // phi[a * kMaxSegNq + k]
};
int build_segment_table(int degree, const lfgpu_quad* qr, SegTable* out, std::string* err);

int nsf_of(int degree, int cell_type);
// returns 0 or a negative status; `qr` may be null (default rule 2*degree)
int build_fe_table(int degree, int cell_type, const lfgpu_quad* qr, FeTable* out, std::string* err);
void build_fe_tensors(const FeTable& t, FeTensors* out);
int default_quad_rule(int cell_type, int degree, int capacity, double* points, double* weights);

// flag[r] &= keep[r] on the ctx stream (no-op for keep == nullptr); assemble.cu
int and_row_keep(lfgpu_ctx* ctx, int64_t n, uint8_t* d_flag, const uint8_t* d_keep);
// per-dof gather lists of a dofmap (dofs.cu), cached in the handle
int dofmap_gather_plan(lfgpu_ctx* ctx, const lfgpu_dofmap* d);

// edges numbered? (mesh.cu; lazy for the structured generators)
int ensure_topology(lfgpu_ctx* ctx, lfgpu_mesh* m);

// P1 vertex-fan fast path (assemble_p1.cu)
int p1_fan_prepare(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, lfgpu_pattern* p);
int p1_fan_launch(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* p, const double alpha[4], int tensor, double gamma,
                  double wsum, double m_diag, double m_off, double beta, const int32_t* row_list, int64_t n_rows, double* d_values,
                  int64_t row0 = -1);
// P1 row kernel for quadrilaterals / hybrid meshes / variable coefficients / activity masks (assemble_p1h.cu);
// tt / tq: tables of the rules in use (null = cell type absent from the mesh)
int p1h_prepare(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, lfgpu_pattern* p);
bool p1h_rules_ok(const FeTable* tt, const FeTable* tq);
int p1h_launch(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* p, const FeTable* tt, const FeTable* tq, const lfgpu_coeff* alpha,
               const lfgpu_coeff* gamma, const uint8_t* active, double beta, const int32_t* row_list, int64_t n_rows, int64_t row0,
               double* d_values);
// P1 load vector with a constant source on the vertex rings (assemble_p1.cu)
int p1_load_fan(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_dofmap* d, double c, double beta, double* d_vec, int* handled,
                const double* src_data = nullptr, int src_stride = 0, int nq = 0, const double* wtab = nullptr);
// P2 row kernels (assemble_p2.cu): k00 .. km = reference tensors of FeLagrangeO2Tria, [6 * 6] row-major each
int queue_geometry_check(lfgpu_ctx* ctx, const lfgpu_mesh* mesh);
int edge_node_order(lfgpu_ctx* ctx, int64_t nn, int64_t ne, int32_t* enb, uint32_t** new_id_out);
int permute_node_coords(lfgpu_ctx* ctx, int64_t nn, const uint32_t* new_id, const double* xy, double* xy_perm);
int build_row_dict(lfgpu_ctx* ctx, int n_words, int64_t n, const uint32_t* words, uint16_t* idx, void** dict_out, int* n_dict);
int p2_rows_prepare(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, lfgpu_pattern* p);
int p2_rows_launch(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* p, const double alpha[4], int tensor, double gamma,
                   const double* k00, const double* k01, const double* k10, const double* k11, const double* km, double* d_values,
                   int64_t r0, int64_t r1, double beta = 0.0);
// P3 row kernels (assemble_p3.cu): k00 .. km = reference tensors of FeLagrangeO3Tria, [10 * 10] row-major each
int p3_rows_prepare(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, lfgpu_pattern* p);
int p3_rows_launch(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* p, const double alpha[4], int tensor, double gamma,
                   const double* k00, const double* k01, const double* k10, const double* k11, const double* km, double* d_values,
                   int64_t r0, int64_t r1, double beta = 0.0);
// body of lfgpu_assemble_reaction_diffusion_rows (assemble.cu) with two extras used by the host pipeline (hostpipe.cu)
int assemble_rd_impl(lfgpu_ctx* ctx, const lfgpu_mesh* mesh, const lfgpu_pattern* p, int degree, const lfgpu_quad* qr_tria,
                     const lfgpu_quad* qr_quad, const lfgpu_coeff* alpha, const lfgpu_coeff* gamma, const uint8_t* active, double beta,
                     double* d_values, int algo, const int32_t* d_row_list, int64_t n_rows, int64_t row0, int* fan_query);

}  // namespace lfgpu
#endif
