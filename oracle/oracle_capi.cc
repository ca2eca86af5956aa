// ORACLE (test infrastructure, NOT product code) -- see lfo_base.h header.
// C interface over the CPU restatement so that tests/ (ctypes) and bench.py's cpu_baseline / --impl reference legs can
// drive it.  Nothing in lehrfempp_b200/ links or loads this library.
#include <chrono>
#include <cstdint>
#include <string>

#include "lfo_refinement.h"
#include "lfo_uscalfe.h"

using namespace lfo;

namespace {
thread_local std::string g_err;

struct MeshH {
  std::shared_ptr<mesh::Mesh> mesh;
};
struct DofH {
  std::shared_ptr<mesh::Mesh> mesh;
  std::unique_ptr<assemble::DofHandler> dofh;  // UniformFEDofHandler or DynamicFEDofHandler
  size_type stride = 0;                        // longest cell list (the row length of the exported table)
  void set_stride() {
    stride = 0;
    for (const mesh::Entity* cell : mesh->Entities(0)) stride = std::max(stride, dofh->NumLocalDofs(*cell));
  }
};

double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// assemble/test/assembly_tests.cc:62-81 (TestAssembler), :84-103 (TestVectorAssembler), :323-347 (EdgeDofAssembler)
struct TestAssembler {
  const mesh::Mesh& mesh_;
  bool isActive(const mesh::Entity&) { return true; }
  Mat Eval(const mesh::Entity& cell) {
    const double idx = mesh_.Index(cell);
    Mat m(4, 4);
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) m(i, j) = (i == j) ? idx : idx * -1.0;
    return m;
  }
};
struct TestVectorAssembler {
  const mesh::Mesh& mesh_;
  bool isActive(const mesh::Entity&) { return true; }
  Mat Eval(const mesh::Entity& cell) {
    const double idx = mesh_.Index(cell);
    Mat v(4, 1);
    for (int i = 0; i < 4; ++i) v[i] = idx;
    return v;
  }
};
struct EdgeDofAssembler {
  const mesh::Mesh& mesh_;
  bool isActive(const mesh::Entity&) { return true; }
  Mat Eval(const mesh::Entity& cell) {
    const double idx = mesh_.Index(cell);
    Mat m(8, 8);
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 8; ++j) m(i, j) = idx * -1.0;
    for (int i = 0; i < 8; i += 2)
      for (int j = 0; j < 8; j += 2) m(i, j) = idx;
    for (int i = 0; i < 8; ++i) m(i, i) = idx;
    return m;
  }
};
// assemble/test/assembly_tests.cc:491-560 (BoundaryAssembler: edges with exactly one adjacent cell get [[1,-1],[-1,1]]*?)
}  // namespace

extern "C" {

struct lfo_coeff {
  int kind;            // 0 const scalar, 1 const 2x2 (row-major c[0..3]), 2 builtin scalar fn id=c[0], 3 builtin 2x2 fn,
                       // 4 table (stride values per cell: 1 = per cell, >1 = per quadrature point), 5 callback scalar,
                       // 6 callback 2x2
  double c[4];
  const double* table;
  long stride;
  double (*fn)(double, double);
  void (*fn2)(double, double, double*);
};

const char* lfo_last_error() { return g_err.c_str(); }

#define LFO_TRY try {
#define LFO_CATCH(ret)                 \
  }                                    \
  catch (const std::exception& e) {    \
    g_err = e.what();                  \
    return ret;                        \
  }

void* lfo_mesh_tp_tria(unsigned nx, unsigned ny, double x0, double y0, double x1, double y1) {
  LFO_TRY
  auto m = mesh::utils::TPTriagMeshBuild(nx, ny, x0, y0, x1, y1);
  if (!m) throw LfException("empty mesh");
  return new MeshH{m};
  LFO_CATCH(nullptr)
}
void* lfo_mesh_tp_quad(unsigned nx, unsigned ny, double x0, double y0, double x1, double y1) {
  LFO_TRY
  auto m = mesh::utils::TPQuadMeshBuild(nx, ny, x0, y0, x1, y1);
  if (!m) throw LfException("empty mesh");
  return new MeshH{m};
  LFO_CATCH(nullptr)
}
void* lfo_mesh_hybrid(unsigned n, double jitter, std::uint64_t seed) {
  LFO_TRY
  auto m = mesh::utils::HybridMeshBuild(n, jitter, seed);
  if (!m) throw LfException("empty mesh");
  return new MeshH{m};
  LFO_CATCH(nullptr)
}
// cell_geo: 0 = no geometry supplied (built from node positions), 1 = TriaO1/QuadO1 from cell_coords,
//           2 = Parallelogram from cell_coords (quads only).  n_edges explicit edges (nullable) are registered first.
void* lfo_mesh_from_arrays(std::int64_t n_nodes, const double* xy, std::int64_t n_cells, const std::uint32_t* cell_nodes,
                           const double* cell_coords, const std::uint8_t* cell_geo, std::int64_t n_edges,
                           const std::uint32_t* edge_nodes) {
  LFO_TRY
  mesh::hybrid2d::MeshFactory f;
  for (std::int64_t i = 0; i < n_nodes; ++i) f.AddPoint(xy[2 * i], xy[2 * i + 1]);
  for (std::int64_t e = 0; e < n_edges; ++e) {
    const std::array<size_type, 2> nl{edge_nodes[2 * e], edge_nodes[2 * e + 1]};
    Mat g(2, 2);
    for (int k = 0; k < 2; ++k) {
      g(0, k) = xy[2 * nl[k]];
      g(1, k) = xy[2 * nl[k] + 1];
    }
    f.AddEntity(RefEl::kSegment(), nl, std::make_unique<geometry::SegmentO1>(g));
  }
  for (std::int64_t c = 0; c < n_cells; ++c) {
    const std::uint32_t* cn = cell_nodes + 4 * c;
    const int nv = (cn[3] == kIdxNil) ? 3 : 4;
    const int geo = (cell_geo != nullptr) ? cell_geo[c] : 0;
    mesh::GeometryPtr g;
    if (geo != 0) {
      Mat cc(2, nv);
      for (int k = 0; k < nv; ++k) {
        cc(0, k) = cell_coords[8 * c + 2 * k];
        cc(1, k) = cell_coords[8 * c + 2 * k + 1];
      }
      if (nv == 3) {
        g = std::make_unique<geometry::TriaO1>(cc);
      } else if (geo == 2) {
        g = std::make_unique<geometry::Parallelogram>(cc);
      } else {
        g = std::make_unique<geometry::QuadO1>(cc);
      }
    }
    if (nv == 3) {
      f.AddEntity(RefEl::kTria(), std::span<const size_type>(cn, 3), std::move(g));
    } else {
      f.AddEntity(RefEl::kQuad(), std::span<const size_type>(cn, 4), std::move(g));
    }
  }
  return new MeshH{f.Build()};
  LFO_CATCH(nullptr)
}
// MeshHierarchy::RefineRegular() + getMesh(finest): one regular refinement step with the reference's numbering
void* lfo_mesh_refine_regular(void* h) {
  LFO_TRY
  auto m = refinement::RefineRegular(*static_cast<MeshH*>(h)->mesh);
  if (!m) throw LfException("empty mesh");
  return new MeshH{m};
  LFO_CATCH(nullptr)
}
void lfo_mesh_free(void* h) { delete static_cast<MeshH*>(h); }

void lfo_mesh_counts(void* h, std::int64_t* nn, std::int64_t* ne, std::int64_t* nc, std::int64_t* ntria, std::int64_t* nquad) {
  const auto& m = *static_cast<MeshH*>(h)->mesh;
  *nn = m.NumEntities(2);
  *ne = m.NumEntities(1);
  *nc = m.NumEntities(0);
  *ntria = m.NumEntities(RefEl::kTria());
  *nquad = m.NumEntities(RefEl::kQuad());
}

// all outputs nullable; cell arrays have 4 slots per cell (kIdxNil / 0 padded for triangles)
int lfo_mesh_export(void* h, std::uint8_t* cell_type, std::uint32_t* cell_nodes, double* cell_coords,
                    std::uint32_t* cell_edges, std::int8_t* cell_edge_ori, std::uint32_t* edge_nodes, double* node_coords) {
  LFO_TRY
  const auto& m = *static_cast<MeshH*>(h)->mesh;
  const Mat zero_point(0, 1);
  std::size_t c = 0;
  for (const mesh::Entity* cell : m.Entities(0)) {
    const RefEl r = cell->RefElem();
    const int nv = r.NumNodes();
    if (cell_type) cell_type[c] = static_cast<std::uint8_t>(r.Id());
    const auto nodes = cell->SubEntities(2);
    const auto edges = cell->SubEntities(1);
    const auto ori = cell->RelativeOrientations();
    const Mat corners = cell->Geometry()->Global(r.NodeCoords());
    for (int k = 0; k < 4; ++k) {
      if (cell_nodes) cell_nodes[4 * c + k] = (k < nv) ? m.Index(*nodes[k]) : kIdxNil;
      if (cell_edges) cell_edges[4 * c + k] = (k < nv) ? m.Index(*edges[k]) : kIdxNil;
      if (cell_edge_ori) cell_edge_ori[4 * c + k] = (k < nv) ? static_cast<std::int8_t>(ori[k]) : 0;
      if (cell_coords) {
        cell_coords[8 * c + 2 * k] = (k < nv) ? corners(0, k) : 0.0;
        cell_coords[8 * c + 2 * k + 1] = (k < nv) ? corners(1, k) : 0.0;
      }
    }
    ++c;
  }
  if (edge_nodes) {
    std::size_t e = 0;
    for (const mesh::Entity* edge : m.Entities(1)) {
      const auto ep = edge->SubEntities(1);
      edge_nodes[2 * e] = m.Index(*ep[0]);
      edge_nodes[2 * e + 1] = m.Index(*ep[1]);
      ++e;
    }
  }
  if (node_coords) {
    std::size_t n = 0;
    for (const mesh::Entity* node : m.Entities(2)) {
      const Mat p = node->Geometry()->Global(zero_point);
      node_coords[2 * n] = p(0, 0);
      node_coords[2 * n + 1] = p(1, 0);
      ++n;
    }
  }
  return 0;
  LFO_CATCH(-1)
}

void* lfo_dofh_create(void* mesh_h, unsigned n_pt, unsigned n_seg, unsigned n_tria, unsigned n_quad) {
  LFO_TRY
  auto* mh = static_cast<MeshH*>(mesh_h);
  auto d = new DofH{mh->mesh, nullptr, 0};
  assemble::UniformFEDofHandler::dof_map_t layout;
  if (n_pt) layout[RefEl::kPoint()] = n_pt;
  if (n_seg) layout[RefEl::kSegment()] = n_seg;
  if (n_tria) layout[RefEl::kTria()] = n_tria;
  if (n_quad) layout[RefEl::kQuad()] = n_quad;
  auto u = std::make_unique<assemble::UniformFEDofHandler>(mh->mesh, layout);
  d->stride = u->CellStride();
  d->dofh = std::move(u);
  return d;
  LFO_CATCH(nullptr)
}
// DynamicFEDofHandler(mesh, locdof) with locdof tabulated per entity: n_int_node [n_nodes], n_int_edge [n_edges],
// n_int_cell [n_cells] (any may be null = 0 for that codimension)
void* lfo_dofh_create_dynamic(void* mesh_h, const std::uint32_t* n_int_node, const std::uint32_t* n_int_edge,
                              const std::uint32_t* n_int_cell) {
  LFO_TRY
  auto* mh = static_cast<MeshH*>(mesh_h);
  auto d = new DofH{mh->mesh, nullptr, 0};
  const mesh::Mesh& m = *mh->mesh;
  const std::uint32_t* tab[3] = {n_int_cell, n_int_edge, n_int_node};
  d->dofh = std::make_unique<assemble::DynamicFEDofHandler>(mh->mesh, [&](const mesh::Entity& e) -> size_type {
    const std::uint32_t* t = tab[e.Codim()];
    return t ? t[m.Index(e)] : 0U;
  });
  d->set_stride();
  return d;
  LFO_CATCH(nullptr)
}
void lfo_dofh_free(void* h) { delete static_cast<DofH*>(h); }
std::int64_t lfo_dofh_num_dofs(void* h) { return static_cast<DofH*>(h)->dofh->NumDofs(); }
int lfo_dofh_stride(void* h) { return static_cast<int>(static_cast<DofH*>(h)->stride); }
// cell_dofs [n_cells][stride] (unused slots = -1), n_ldof [n_cells]
int lfo_dofh_export(void* h, std::int64_t* cell_dofs, std::uint8_t* n_ldof) {
  LFO_TRY
  auto* d = static_cast<DofH*>(h);
  const std::size_t stride = d->stride;
  std::size_t c = 0;
  for (const mesh::Entity* cell : d->mesh->Entities(0)) {
    const auto idx = d->dofh->GlobalDofIndices(*cell);
    const size_type n = d->dofh->NumLocalDofs(*cell);
    for (std::size_t k = 0; k < stride; ++k) cell_dofs[c * stride + k] = (k < n) ? idx[k] : -1;
    if (n_ldof) n_ldof[c] = static_cast<std::uint8_t>(n);
    ++c;
  }
  return 0;
  LFO_CATCH(-1)
}
// entity (codim, index) carrying each dof: out_codim[N], out_index[N]
int lfo_dofh_dof_entities(void* h, std::uint8_t* out_codim, std::uint32_t* out_index) {
  LFO_TRY
  auto* d = static_cast<DofH*>(h);
  for (std::int64_t i = 0; i < d->dofh->NumDofs(); ++i) {
    const mesh::Entity& e = d->dofh->Entity(i);
    out_codim[i] = static_cast<std::uint8_t>(e.Codim());
    out_index[i] = d->mesh->Index(e);
  }
  return 0;
  LFO_CATCH(-1)
}

// golden-matrix assemblers of assemble/test/assembly_tests.cc; dense_out is N x N row-major
int lfo_assemble_test_matrix(void* h, int kind, double* dense_out) {
  LFO_TRY
  auto* d = static_cast<DofH*>(h);
  const long N = d->dofh->NumDofs();
  assemble::COOMatrix coo(N, N);
  if (kind == 0) {
    TestAssembler a{*d->mesh};
    assemble::AssembleMatrixLocally(0, *d->dofh, *d->dofh, a, coo);
  } else {
    EdgeDofAssembler a{*d->mesh};
    assemble::AssembleMatrixLocally(0, *d->dofh, *d->dofh, a, coo);
  }
  const auto cm = coo.makeSparse();
  for (long i = 0; i < N * N; ++i) dense_out[i] = 0.0;
  for (long c = 0; c < cm.cols; ++c)
    for (int k = cm.outer[c]; k < cm.outer[c + 1]; ++k) dense_out[cm.inner[k] * N + c] = cm.values[k];
  return 0;
  LFO_CATCH(-1)
}
int lfo_assemble_test_vector(void* h, double* out) {
  LFO_TRY
  auto* d = static_cast<DofH*>(h);
  std::vector<double> v(d->dofh->NumDofs(), 0.0);
  TestVectorAssembler a{*d->mesh};
  assemble::AssembleVectorLocally(0, *d->dofh, a, v);
  std::copy(v.begin(), v.end(), out);
  return 0;
  LFO_CATCH(-1)
}

}  // extern "C"

namespace {
constexpr double kPi = 3.14159265358979323846;
double builtin_scalar(int id, double x, double y) {
  switch (id) {
    case 1: return 1.0 + x * x + y * y;           // uscalfe/test/full_gal_tests.cc:101-103
    case 2: return 1.0 / (1.0 + x * x + y * y);   // uscalfe/test/loc_comp_test.cc:109-111
    case 3: return std::sin(2 * kPi * x) * std::sin(2 * kPi * y);  // uscalfe/test/bvp_fe_tests.cc:33
    case 4: return x * y;
    case 5: return x;
    case 6: return y;
    case 7: return x * x - y * y;                 // uscalfe/test/loc_comp_test.cc:113
    case 8: return x * x + y * y;                 // uscalfe/test/loc_comp_test.cc:106
    case 9: return 1 + x + 2 * y;                 // lagr_fe_tests.cc:826
    case 10: return 3 * x;                        // lagr_fe_tests.cc:828
    case 11: return x * x * x + y * y * y;        // lagr_fe_tests.cc:875
    case 12: return x * y * y;                    // lagr_fe_tests.cc:878
    default: throw LfException("unknown builtin scalar function id");
  }
}
uscalfe::Mat2 builtin_tensor(int id, double x, double y) {
  switch (id) {
    case 100: return uscalfe::Mat2{{{1, x}, {y, x * y}}};  // lagr_fe_tests.cc:818-820
    case 101: return uscalfe::Mat2{{{3.0, 0.0}, {1.0, 2.0}}};
    default: throw LfException("unknown builtin tensor function id");
  }
}

using ScalarMF = std::function<std::vector<double>(const mesh::Entity&, const Mat&)>;
using TensorMF = std::function<std::vector<uscalfe::Mat2>(const mesh::Entity&, const Mat&)>;

ScalarMF make_scalar_mf(const lfo_coeff* c, const mesh::Mesh* mesh) {
  switch (c->kind) {
    case 0: return uscalfe::MeshFunctionConstant<double>(c->c[0]);
    case 2: {
      const int id = static_cast<int>(c->c[0]);
      return uscalfe::MeshFunctionGlobal<double>([id](double x, double y) { return builtin_scalar(id, x, y); });
    }
    case 4: return uscalfe::MeshFunctionTable(mesh, c->table, c->stride);
    case 5: {
      auto fn = c->fn;
      return uscalfe::MeshFunctionGlobal<double>([fn](double x, double y) { return fn(x, y); });
    }
    default: throw LfException("coefficient kind is not scalar");
  }
}
TensorMF make_tensor_mf(const lfo_coeff* c) {
  switch (c->kind) {
    case 1: return uscalfe::MeshFunctionConstant<uscalfe::Mat2>(uscalfe::Mat2{{{c->c[0], c->c[1]}, {c->c[2], c->c[3]}}});
    case 3: {
      const int id = static_cast<int>(c->c[0]);
      return uscalfe::MeshFunctionGlobal<uscalfe::Mat2>([id](double x, double y) { return builtin_tensor(id, x, y); });
    }
    case 6: {
      auto fn = c->fn2;
      return uscalfe::MeshFunctionGlobal<uscalfe::Mat2>([fn](double x, double y) {
        double o[4];
        fn(x, y, o);
        return uscalfe::Mat2{{{o[0], o[1]}, {o[2], o[3]}}};
      });
    }
    default: throw LfException("coefficient kind is not a 2x2 tensor");
  }
}
bool is_tensor(const lfo_coeff* c) { return c->kind == 1 || c->kind == 3 || c->kind == 6; }

std::map<RefEl, quad::QuadRule> make_rules(int qr_tria, int qr_quad) {
  std::map<RefEl, quad::QuadRule> rules;
  if (qr_tria >= 0) rules[RefEl::kTria()] = quad::make_QuadRule(RefEl::kTria(), qr_tria);
  if (qr_quad >= 0) rules[RefEl::kQuad()] = quad::make_QuadRule(RefEl::kQuad(), qr_quad);
  return rules;
}

// provider with an optional activity mask (EntityMatrixProvider::isActive, loc_comp_ellbvp.h:155)
template <class BASE>
struct Masked : BASE {
  using BASE::BASE;
  const std::uint8_t* mask = nullptr;
  const mesh::Mesh* mesh = nullptr;
  bool isActive(const mesh::Entity& cell) override { return mask == nullptr || mask[mesh->Index(cell)] != 0; }
};

// a COO sink that swaps (i, j): its makeSparse() is then the compressed ROW storage (CSR) of the assembled matrix
struct TransposedCOO {
  assemble::COOMatrix& coo;
  void AddToEntry(gdof_idx_t i, gdof_idx_t j, double v) { coo.AddToEntry(j, i, v); }
};

template <class PROVIDER>
assemble::CompressedMatrix* run_matrix(const uscalfe::UniformScalarFESpace& fes, PROVIDER& prov, int transpose,
                                       double* t_assemble, double* t_makesparse, int repeat_accumulate) {
  const auto& dofh = fes.LocGlobMap();
  assemble::COOMatrix coo(dofh.NumDofs(), dofh.NumDofs());
  const double t0 = now();
  for (int r = 0; r < repeat_accumulate; ++r) {
    if (transpose) {
      TransposedCOO tc{coo};
      assemble::AssembleMatrixLocally(0, dofh, dofh, prov, tc);
    } else {
      assemble::AssembleMatrixLocally(0, dofh, dofh, prov, coo);
    }
  }
  const double t1 = now();
  auto* cm = new assemble::CompressedMatrix(coo.makeSparse());
  const double t2 = now();
  if (t_assemble) *t_assemble = t1 - t0;
  if (t_makesparse) *t_makesparse = t2 - t1;
  return cm;
}
}  // namespace

extern "C" {

// Reaction-diffusion Galerkin matrix through the reference call sequence
//   FeSpaceLagrangeO<degree>(mesh) -> ReactionDiffusionElementMatrixProvider(fe_space, alpha, gamma[, rules])
//   -> AssembleMatrixLocally(0, dofh, dofh, provider, COOMatrix) -> makeSparse()
// qr_tria / qr_quad < 0: default rules (degree 2p); if exactly one is >= 0 the other type has NO rule (Eval throws).
// transpose = 0: Eigen column-major arrays of A;  1: compressed-row (CSR) arrays of A.
// repeat_accumulate: number of AssembleMatrixLocally calls into the same COO (assembler.h:84-88 accumulate semantics).
void* lfo_assemble_rd(void* mesh_h, int degree, int qr_tria, int qr_quad, const lfo_coeff* alpha, const lfo_coeff* gamma,
                      const std::uint8_t* active, int transpose, int repeat_accumulate, double* t_assemble,
                      double* t_makesparse) {
  LFO_TRY
  auto* mh = static_cast<MeshH*>(mesh_h);
  auto fes = std::make_shared<uscalfe::UniformScalarFESpace>(mh->mesh, degree);
  const bool custom = qr_tria >= 0 || qr_quad >= 0;
  const auto rules = make_rules(qr_tria, qr_quad);
  if (repeat_accumulate < 1) repeat_accumulate = 1;
  if (alpha->kind == 0 && gamma->kind == 0) {
    // the headline configuration: both coefficients are MeshFunctionConstant<double>, no type erasure
    using P = uscalfe::ReactionDiffusionElementMatrixProvider<uscalfe::MeshFunctionConstant<double>, uscalfe::MeshFunctionConstant<double>>;
    uscalfe::MeshFunctionConstant<double> a(alpha->c[0]), g(gamma->c[0]);
    auto prov = custom ? Masked<P>(fes, a, g, rules) : Masked<P>(fes, a, g);
    prov.mask = active;
    prov.mesh = mh->mesh.get();
    return run_matrix(*fes, prov, transpose, t_assemble, t_makesparse, repeat_accumulate);
  }
  ScalarMF g = make_scalar_mf(gamma, mh->mesh.get());
  if (is_tensor(alpha)) {
    using P = uscalfe::ReactionDiffusionElementMatrixProvider<TensorMF, ScalarMF>;
    TensorMF a = make_tensor_mf(alpha);
    auto prov = custom ? Masked<P>(fes, a, g, rules) : Masked<P>(fes, a, g);
    prov.mask = active;
    prov.mesh = mh->mesh.get();
    return run_matrix(*fes, prov, transpose, t_assemble, t_makesparse, repeat_accumulate);
  }
  using P = uscalfe::ReactionDiffusionElementMatrixProvider<ScalarMF, ScalarMF>;
  ScalarMF a = make_scalar_mf(alpha, mh->mesh.get());
  auto prov = custom ? Masked<P>(fes, a, g, rules) : Masked<P>(fes, a, g);
  prov.mask = active;
  prov.mesh = mh->mesh.get();
  return run_matrix(*fes, prov, transpose, t_assemble, t_makesparse, repeat_accumulate);
  LFO_CATCH(nullptr)
}
// Assemble the P<degree> reaction-diffusion matrix and load vector (constant coefficients alpha, gamma, source f), then
// FixFlaggedSolutionComponents (assemble/fix_dof.h:86-138) with flags/values per dof, then makeSparse.
// rhs (length N) receives the modified right-hand side.  transpose as in lfo_assemble_rd.
void* lfo_assemble_fixed(void* mesh_h, int degree, double alpha, double gamma, double f, const std::uint8_t* fixed,
                         const double* fixed_vals, int transpose, int alt, double* rhs) {
  LFO_TRY
  auto* mh = static_cast<MeshH*>(mesh_h);
  auto fes = std::make_shared<uscalfe::UniformScalarFESpace>(mh->mesh, degree);
  const auto& dofh = fes->LocGlobMap();
  using MF = uscalfe::MeshFunctionConstant<double>;
  uscalfe::ReactionDiffusionElementMatrixProvider<MF, MF> prov(fes, MF(alpha), MF(gamma));
  assemble::COOMatrix coo(dofh.NumDofs(), dofh.NumDofs());
  assemble::AssembleMatrixLocally(0, dofh, dofh, prov, coo);
  uscalfe::ScalarLoadElementVectorProvider<MF> lprov(fes, MF(f));
  std::vector<double> b(dofh.NumDofs(), 0.0);
  assemble::AssembleVectorLocally(0, dofh, lprov, b);
  auto sel = [&](gdof_idx_t i) { return std::make_pair(fixed[i] != 0, fixed_vals[i]); };
  if (alt) {
    assemble::FixFlaggedSolutionCompAlt(sel, coo, b);
  } else {
    assemble::FixFlaggedSolutionComponents(sel, coo, b);
  }
  std::copy(b.begin(), b.end(), rhs);
  if (transpose) {
    assemble::COOMatrix t(coo.rows(), coo.cols());
    for (const auto& tr : coo.triplets()) t.AddToEntry(tr.col, tr.row, tr.value);
    return new assemble::CompressedMatrix(t.makeSparse());
  }
  return new assemble::CompressedMatrix(coo.makeSparse());
  LFO_CATCH(nullptr)
}

// FixFlaggedSolutionComponents on a caller-supplied triplet list (the shape of the reference's own test,
// assemble/test/coomatrix_tests.cc:181-237): n x n COO matrix from AddToEntry(rows[k], cols[k], vals[k]), rhs in/out.
void* lfo_fix_coo(std::int64_t n, std::int64_t n_trip, const std::int32_t* rows, const std::int32_t* cols, const double* vals,
                  const std::uint8_t* fixed, const double* fixed_vals, int alt, double* rhs) {
  LFO_TRY
  assemble::COOMatrix coo(static_cast<size_type>(n), static_cast<size_type>(n));
  for (std::int64_t k = 0; k < n_trip; ++k) coo.AddToEntry(rows[k], cols[k], vals[k]);
  std::vector<double> b(rhs, rhs + n);
  auto sel = [&](gdof_idx_t i) { return std::make_pair(fixed[i] != 0, fixed_vals[i]); };
  if (alt) {
    assemble::FixFlaggedSolutionCompAlt(sel, coo, b);
  } else {
    assemble::FixFlaggedSolutionComponents(sel, coo, b);
  }
  std::copy(b.begin(), b.end(), rhs);
  return new assemble::CompressedMatrix(coo.makeSparse());
  LFO_CATCH(nullptr)
}

// FixSolutionComponentsLse (fix_dof.h:250-280) on a triplet list: prescribed components as (index, value) pairs
void* lfo_fix_coo_lse(std::int64_t n, std::int64_t n_trip, const std::int32_t* rows, const std::int32_t* cols, const double* vals,
                      std::int64_t n_pairs, const std::int64_t* pair_idx, const double* pair_val, double* rhs) {
  LFO_TRY
  assemble::COOMatrix coo(static_cast<size_type>(n), static_cast<size_type>(n));
  for (std::int64_t k = 0; k < n_trip; ++k) coo.AddToEntry(rows[k], cols[k], vals[k]);
  std::vector<double> b(rhs, rhs + n);
  assemble::fixed_components_t fixed;
  for (std::int64_t k = 0; k < n_pairs; ++k) fixed.emplace_back(pair_idx[k], pair_val[k]);
  assemble::FixSolutionComponentsLse(fixed, coo, b);
  std::copy(b.begin(), b.end(), rhs);
  return new assemble::CompressedMatrix(coo.makeSparse());
  LFO_CATCH(nullptr)
}

// ---- edge (codim-1) contributions: SURVEY section 8f row 2 -----------------------------------------------------------
namespace {
// edges with exactly one adjacent cell (mesh/utils: flagEntitiesOnBoundary(mesh, 1) / CountNumSuperEntities(mesh, 1, 1))
std::vector<std::uint8_t> boundary_edge_flags(const mesh::Mesh& m) {
  std::vector<unsigned> cnt(m.NumEntities(1), 0);
  for (const mesh::Entity* cell : m.Entities(0))
    for (const mesh::Entity* e : cell->SubEntities(1)) cnt[m.Index(*e)]++;
  std::vector<std::uint8_t> f(cnt.size());
  for (std::size_t i = 0; i < cnt.size(); ++i) f[i] = cnt[i] == 1 ? 1 : 0;
  return f;
}
struct EdgeMask {
  const std::uint8_t* mask;
  const mesh::Mesh* mesh;
  bool operator()(const mesh::Entity& e) const { return mask == nullptr || mask[mesh->Index(e)] != 0; }
};
quad::QuadRule segment_rule(int degree, int qr_degree) {
  return quad::make_QuadRule(RefEl::kSegment(), qr_degree >= 0 ? static_cast<unsigned>(qr_degree) : 2U * degree);
}
// assemble/test/assembly_tests.cc:491-523 (BoundaryAssembler): boundary edges get idx * [[1, -1], [-1, 1]]
struct BoundaryAssembler {
  const mesh::Mesh& mesh_;
  std::vector<std::uint8_t> bd_;
  explicit BoundaryAssembler(const mesh::Mesh& m) : mesh_(m), bd_(boundary_edge_flags(m)) {}
  bool isActive(const mesh::Entity& edge) { return bd_[mesh_.Index(edge)] != 0; }
  Mat Eval(const mesh::Entity& edge) {
    const double idx = mesh_.Index(edge);
    Mat m(2, 2);
    m(0, 0) = idx;
    m(1, 1) = idx;
    m(0, 1) = -idx;
    m(1, 0) = -idx;
    return m;
  }
};
}  // namespace

int lfo_boundary_edges(void* mesh_h, std::uint8_t* flags) {
  LFO_TRY
  const auto f = boundary_edge_flags(*static_cast<MeshH*>(mesh_h)->mesh);
  std::copy(f.begin(), f.end(), flags);
  return 0;
  LFO_CATCH(-1)
}
// assembly_tests.cc:525-590: AssembleMatrixLocally(1, dofh, BoundaryAssembler) -> dense N x N (row-major)
int lfo_assemble_boundary_test_matrix(void* dofh_h, double* dense_out) {
  LFO_TRY
  auto* d = static_cast<DofH*>(dofh_h);
  const long N = d->dofh->NumDofs();
  assemble::COOMatrix coo(N, N);
  BoundaryAssembler a(*d->mesh);
  assemble::AssembleMatrixLocally(1, *d->dofh, *d->dofh, a, coo);
  const auto cm = coo.makeSparse();
  for (long i = 0; i < N * N; ++i) dense_out[i] = 0.0;
  for (long c = 0; c < cm.cols; ++c)
    for (int k = cm.outer[c]; k < cm.outer[c + 1]; ++k) dense_out[cm.inner[k] * N + c] = cm.values[k];
  return 0;
  LFO_CATCH(-1)
}
// MassEdgeMatrixProvider::Eval / ScalarLoadEdgeVectorProvider::Eval for every edge: out[edge][b * stride + a] (column-major
// blocks) resp. out[edge][a]; qr_degree < 0 = default rule of degree 2p
int lfo_edge_matrices(void* mesh_h, int degree, int qr_degree, const lfo_coeff* eta, double* out, int stride) {
  LFO_TRY
  auto* mh = static_cast<MeshH*>(mesh_h);
  auto fes = std::make_shared<uscalfe::UniformScalarFESpace>(mh->mesh, degree);
  uscalfe::MassEdgeMatrixProvider<ScalarMF, EdgeMask> prov(fes, make_scalar_mf(eta, mh->mesh.get()), segment_rule(degree, qr_degree),
                                                          EdgeMask{nullptr, mh->mesh.get()});
  std::size_t e = 0;
  for (const mesh::Entity* edge : mh->mesh->Entities(1)) {
    const Mat m = prov.Eval(*edge);
    for (long j = 0; j < m.cols(); ++j)
      for (long i = 0; i < m.rows(); ++i) out[e * stride * stride + j * stride + i] = m(i, j);
    ++e;
  }
  return 0;
  LFO_CATCH(-1)
}
int lfo_edge_vectors(void* mesh_h, int degree, int qr_degree, const lfo_coeff* g, double* out, int stride) {
  LFO_TRY
  auto* mh = static_cast<MeshH*>(mesh_h);
  auto fes = std::make_shared<uscalfe::UniformScalarFESpace>(mh->mesh, degree);
  uscalfe::ScalarLoadEdgeVectorProvider<ScalarMF, EdgeMask> prov(fes, make_scalar_mf(g, mh->mesh.get()), segment_rule(degree, qr_degree),
                                                                EdgeMask{nullptr, mh->mesh.get()});
  std::size_t e = 0;
  for (const mesh::Entity* edge : mh->mesh->Entities(1)) {
    const Mat v = prov.Eval(*edge);
    for (long i = 0; i < v.size(); ++i) out[e * stride + i] = v[i];
    ++e;
  }
  return 0;
  LFO_CATCH(-1)
}
// The Galerkin matrix of a second-order BVP with impedance part (uscalfe/test/sec_ord_ell_bvp.h:147-215):
//   AssembleMatrixLocally(0, ..., ReactionDiffusionElementMatrixProvider(alpha, gamma), A);
//   AssembleMatrixLocally(1, ..., MassEdgeMatrixProvider(eta, edge_sel), A);   -> A.makeSparse()
// edge_mask: uint8 per edge (NULL = all edges), transpose as in lfo_assemble_rd.
void* lfo_assemble_rd_edge(void* mesh_h, int degree, const lfo_coeff* alpha, const lfo_coeff* gamma, const lfo_coeff* eta,
                           int qr_degree, const std::uint8_t* edge_mask, int transpose) {
  LFO_TRY
  auto* mh = static_cast<MeshH*>(mesh_h);
  auto fes = std::make_shared<uscalfe::UniformScalarFESpace>(mh->mesh, degree);
  const auto& dofh = fes->LocGlobMap();
  assemble::COOMatrix coo(dofh.NumDofs(), dofh.NumDofs());
  TransposedCOO tcoo{coo};
  ScalarMF g = make_scalar_mf(gamma, mh->mesh.get());
  auto cells = [&](auto& prov) {
    if (transpose) {
      assemble::AssembleMatrixLocally(0, dofh, dofh, prov, tcoo);
    } else {
      assemble::AssembleMatrixLocally(0, dofh, dofh, prov, coo);
    }
  };
  if (is_tensor(alpha)) {
    uscalfe::ReactionDiffusionElementMatrixProvider<TensorMF, ScalarMF> prov(fes, make_tensor_mf(alpha), g);
    cells(prov);
  } else {
    uscalfe::ReactionDiffusionElementMatrixProvider<ScalarMF, ScalarMF> prov(fes, make_scalar_mf(alpha, mh->mesh.get()), g);
    cells(prov);
  }
  uscalfe::MassEdgeMatrixProvider<ScalarMF, EdgeMask> eprov(fes, make_scalar_mf(eta, mh->mesh.get()), segment_rule(degree, qr_degree),
                                                           EdgeMask{edge_mask, mh->mesh.get()});
  if (transpose) {
    assemble::AssembleMatrixLocally(1, dofh, dofh, eprov, tcoo);
  } else {
    assemble::AssembleMatrixLocally(1, dofh, dofh, eprov, coo);
  }
  return new assemble::CompressedMatrix(coo.makeSparse());
  LFO_CATCH(nullptr)
}
// AssembleVectorLocally(1, dofh, ScalarLoadEdgeVectorProvider(g, edge_sel), out): accumulates into out (length N)
int lfo_assemble_edge_load(void* mesh_h, int degree, int qr_degree, const lfo_coeff* g, const std::uint8_t* edge_mask, double* out) {
  LFO_TRY
  auto* mh = static_cast<MeshH*>(mesh_h);
  auto fes = std::make_shared<uscalfe::UniformScalarFESpace>(mh->mesh, degree);
  uscalfe::ScalarLoadEdgeVectorProvider<ScalarMF, EdgeMask> prov(fes, make_scalar_mf(g, mh->mesh.get()), segment_rule(degree, qr_degree),
                                                                EdgeMask{edge_mask, mh->mesh.get()});
  std::span<double> v(out, fes->LocGlobMap().NumDofs());
  assemble::AssembleVectorLocally(1, fes->LocGlobMap(), prov, v);
  return 0;
  LFO_CATCH(-1)
}

void lfo_cm_sizes(void* h, std::int64_t* rows, std::int64_t* cols, std::int64_t* nnz) {
  auto* cm = static_cast<assemble::CompressedMatrix*>(h);
  *rows = cm->rows;
  *cols = cm->cols;
  *nnz = static_cast<std::int64_t>(cm->values.size());
}
void lfo_cm_export(void* h, std::int32_t* outer, std::int32_t* inner, double* values) {
  auto* cm = static_cast<assemble::CompressedMatrix*>(h);
  if (outer) std::copy(cm->outer.begin(), cm->outer.end(), outer);
  if (inner) std::copy(cm->inner.begin(), cm->inner.end(), inner);
  if (values) std::copy(cm->values.begin(), cm->values.end(), values);
}
void lfo_cm_free(void* h) { delete static_cast<assemble::CompressedMatrix*>(h); }

// Load vector: ScalarLoadElementVectorProvider + AssembleVectorLocally; `out` (length N) is accumulated into
int lfo_assemble_load(void* mesh_h, int degree, int qr_tria, int qr_quad, const lfo_coeff* f, const std::uint8_t* active,
                      double* out, double* t_assemble) {
  LFO_TRY
  auto* mh = static_cast<MeshH*>(mesh_h);
  auto fes = std::make_shared<uscalfe::UniformScalarFESpace>(mh->mesh, degree);
  const bool custom = qr_tria >= 0 || qr_quad >= 0;
  using P = uscalfe::ScalarLoadElementVectorProvider<ScalarMF>;
  ScalarMF mf = make_scalar_mf(f, mh->mesh.get());
  auto prov = custom ? Masked<P>(fes, mf, make_rules(qr_tria, qr_quad)) : Masked<P>(fes, mf);
  prov.mask = active;
  prov.mesh = mh->mesh.get();
  std::span<double> v(out, fes->LocGlobMap().NumDofs());
  const double t0 = now();
  assemble::AssembleVectorLocally(0, fes->LocGlobMap(), prov, v);
  if (t_assemble) *t_assemble = now() - t0;
  return 0;
  LFO_CATCH(-1)
}

// per-cell element matrices (column-major nsf x nsf blocks at stride^2 per cell) -- used by the provider KATs
int lfo_element_matrices(void* mesh_h, int degree, int qr_tria, int qr_quad, const lfo_coeff* alpha, const lfo_coeff* gamma,
                         double* out, int stride) {
  LFO_TRY
  auto* mh = static_cast<MeshH*>(mesh_h);
  auto fes = std::make_shared<uscalfe::UniformScalarFESpace>(mh->mesh, degree);
  const bool custom = qr_tria >= 0 || qr_quad >= 0;
  const auto rules = make_rules(qr_tria, qr_quad);
  ScalarMF g = make_scalar_mf(gamma, mh->mesh.get());
  std::size_t c = 0;
  auto run = [&](auto& prov) {
    for (const mesh::Entity* cell : mh->mesh->Entities(0)) {
      const Mat m = prov.Eval(*cell);
      for (long j = 0; j < m.cols(); ++j)
        for (long i = 0; i < m.rows(); ++i) out[c * stride * stride + j * stride + i] = m(i, j);
      ++c;
    }
  };
  if (is_tensor(alpha)) {
    using P = uscalfe::ReactionDiffusionElementMatrixProvider<TensorMF, ScalarMF>;
    TensorMF a = make_tensor_mf(alpha);
    auto prov = custom ? P(fes, a, g, rules) : P(fes, a, g);
    run(prov);
  } else {
    using P = uscalfe::ReactionDiffusionElementMatrixProvider<ScalarMF, ScalarMF>;
    ScalarMF a = make_scalar_mf(alpha, mh->mesh.get());
    auto prov = custom ? P(fes, a, g, rules) : P(fes, a, g);
    run(prov);
  }
  return 0;
  LFO_CATCH(-1)
}

// element matrices of lf::fe::DiffusionElementMatrixProvider (which = 0, coefficient alpha) or MassElementMatrixProvider
// (which = 1, scalar coefficient): out as in lfo_element_matrices
int lfo_fe_element_matrices(void* mesh_h, int degree, int which, const lfo_coeff* coeff, double* out, int stride) {
  LFO_TRY
  auto* mh = static_cast<MeshH*>(mesh_h);
  auto fes = std::make_shared<uscalfe::UniformScalarFESpace>(mh->mesh, degree);
  std::size_t c = 0;
  auto run = [&](auto& prov) {
    for (const mesh::Entity* cell : mh->mesh->Entities(0)) {
      const Mat m = prov.Eval(*cell);
      for (long j = 0; j < m.cols(); ++j)
        for (long i = 0; i < m.rows(); ++i) out[c * stride * stride + j * stride + i] = m(i, j);
      ++c;
    }
  };
  if (which == 1) {
    fe::MassElementMatrixProvider<ScalarMF> prov(fes, make_scalar_mf(coeff, mh->mesh.get()));
    run(prov);
  } else if (is_tensor(coeff)) {
    fe::DiffusionElementMatrixProvider<TensorMF> prov(fes, make_tensor_mf(coeff));
    run(prov);
  } else {
    fe::DiffusionElementMatrixProvider<ScalarMF> prov(fes, make_scalar_mf(coeff, mh->mesh.get()));
    run(prov);
  }
  return 0;
  LFO_CATCH(-1)
}

std::int64_t lfo_fespace_num_dofs(void* mesh_h, int degree) {
  LFO_TRY
  auto* mh = static_cast<MeshH*>(mesh_h);
  uscalfe::UniformScalarFESpace fes(mh->mesh, degree);
  return fes.LocGlobMap().NumDofs();
  LFO_CATCH(-1)
}
// cell dof table of FeSpaceLagrangeO<degree>: cell_dofs [n_cells][stride]; returns stride (or -1)
int lfo_fespace_cell_dofs(void* mesh_h, int degree, std::int64_t* cell_dofs, std::uint8_t* n_ldof) {
  LFO_TRY
  auto* mh = static_cast<MeshH*>(mesh_h);
  uscalfe::UniformScalarFESpace fes(mh->mesh, degree);
  const auto& dofh = fes.LocGlobMap();
  const std::size_t stride = dofh.CellStride();
  if (cell_dofs != nullptr) {
    std::size_t c = 0;
    for (const mesh::Entity* cell : mh->mesh->Entities(0)) {
      const auto idx = dofh.GlobalDofIndices(*cell);
      const size_type n = dofh.NumLocalDofs(*cell);
      for (std::size_t k = 0; k < stride; ++k) cell_dofs[c * stride + k] = (k < n) ? idx[k] : -1;
      if (n_ldof) n_ldof[c] = static_cast<std::uint8_t>(n);
      ++c;
    }
  }
  return static_cast<int>(stride);
  LFO_CATCH(-1)
}

// fe::NodalProjection of a scalar function onto FeSpaceLagrangeO<degree>
int lfo_nodal_projection(void* mesh_h, int degree, const lfo_coeff* u, double* out) {
  LFO_TRY
  auto* mh = static_cast<MeshH*>(mesh_h);
  uscalfe::UniformScalarFESpace fes(mh->mesh, degree);
  ScalarMF mf = make_scalar_mf(u, mh->mesh.get());
  const auto v = uscalfe::NodalProjection(fes, mf);
  std::copy(v.begin(), v.end(), out);
  return 0;
  LFO_CATCH(-1)
}

// quadrature rule: ref_el_id 2 segment / 3 tria / 4 quad.  pts is [dim][n] row-major, returns n (or -1)
int lfo_quad_rule(int ref_el_id, int degree, double* pts, double* wts, int capacity) {
  LFO_TRY
  const RefEl r = ref_el_id == 2 ? RefEl::kSegment() : (ref_el_id == 3 ? RefEl::kTria() : RefEl::kQuad());
  const auto qr = quad::make_QuadRule(r, degree);
  const int n = qr.NumPoints();
  if (pts != nullptr && wts != nullptr) {
    if (n > capacity) throw LfException("capacity too small");
    for (int d = 0; d < static_cast<int>(r.Dimension()); ++d)
      for (int k = 0; k < n; ++k) pts[d * n + k] = qr.Points()(d, k);
    for (int k = 0; k < n; ++k) wts[k] = qr.Weights()[k];
  }
  return n;
  LFO_CATCH(-1)
}

// reference shape functions / gradients of FeLagrangeO<degree>{Tria,Quad} at npts points (pts [2][npts] row-major);
// phi [nsf][npts] row-major, grad [nsf][2*npts] row-major with (2k, 2k+1) = (d/dx0, d/dx1) at point k.  returns nsf
int lfo_eval_fe(int degree, int ref_el_id, int npts, const double* pts, double* phi, double* grad, double* eval_nodes) {
  LFO_TRY
  std::unique_ptr<uscalfe::ScalarReferenceFiniteElement> fe;
  const bool tria = ref_el_id == 3;
  if (degree == 1) fe = tria ? std::unique_ptr<uscalfe::ScalarReferenceFiniteElement>(new uscalfe::FeLagrangeO1Tria) : std::unique_ptr<uscalfe::ScalarReferenceFiniteElement>(new uscalfe::FeLagrangeO1Quad);
  if (degree == 2) fe = tria ? std::unique_ptr<uscalfe::ScalarReferenceFiniteElement>(new uscalfe::FeLagrangeO2Tria) : std::unique_ptr<uscalfe::ScalarReferenceFiniteElement>(new uscalfe::FeLagrangeO2Quad);
  if (degree == 3) fe = tria ? std::unique_ptr<uscalfe::ScalarReferenceFiniteElement>(new uscalfe::FeLagrangeO3Tria) : std::unique_ptr<uscalfe::ScalarReferenceFiniteElement>(new uscalfe::FeLagrangeO3Quad);
  if (!fe) throw LfException("degree must be 1..3");
  const int nsf = fe->NumRefShapeFunctions();
  if (npts > 0) {
    Mat x(2, npts);
    for (int k = 0; k < npts; ++k) {
      x(0, k) = pts[k];
      x(1, k) = pts[npts + k];
    }
    const Mat p = fe->EvalReferenceShapeFunctions(x);
    const Mat g = fe->GradientsReferenceShapeFunctions(x);
    for (int i = 0; i < nsf; ++i) {
      for (int k = 0; k < npts; ++k) phi[i * npts + k] = p(i, k);
      for (int k = 0; k < 2 * npts; ++k) grad[i * 2 * npts + k] = g(i, k);
    }
  }
  if (eval_nodes != nullptr) {
    const Mat n = fe->EvaluationNodes();
    for (int k = 0; k < nsf; ++k) {
      eval_nodes[k] = n(0, k);
      eval_nodes[nsf + k] = n(1, k);
    }
  }
  return nsf;
  LFO_CATCH(-1)
}

// global coordinates of the quadrature points of every cell: out [n_cells][nq_max][2] (unused slots 0)
int lfo_qp_coords(void* mesh_h, int qr_tria, int qr_quad, int nq_max, double* out) {
  LFO_TRY
  auto* mh = static_cast<MeshH*>(mesh_h);
  const auto qt = quad::make_QuadRule(RefEl::kTria(), qr_tria);
  const auto qq = quad::make_QuadRule(RefEl::kQuad(), qr_quad);
  std::size_t c = 0;
  for (const mesh::Entity* cell : mh->mesh->Entities(0)) {
    const auto& qr = cell->RefElem() == RefEl::kTria() ? qt : qq;
    const Mat g = cell->Geometry()->Global(qr.Points());
    for (int k = 0; k < nq_max; ++k) {
      out[(c * nq_max + k) * 2] = k < g.cols() ? g(0, k) : 0.0;
      out[(c * nq_max + k) * 2 + 1] = k < g.cols() ? g(1, k) : 0.0;
    }
    ++c;
  }
  return 0;
  LFO_CATCH(-1)
}

double lfo_builtin_scalar(int id, double x, double y) {
  try {
    return builtin_scalar(id, x, y);
  } catch (...) {
    return std::nan("");
  }
}

}  // extern "C"
