// lf_gpu_shim.hpp -- header-only C++ host shim over the C ABI of liblfgpu.so (include/lfgpu.h).
//
// Keeps the reference call shape
//     lf::assemble::AssembleMatrixLocally(codim, dofh_trial, dofh_test, provider, matrix)   lib/lf/assemble/assembler.h:114-186
//     lf::assemble::AssembleVectorLocally(codim, dofh, provider, vector)                    lib/lf/assemble/assembler.h:298-327
// and adds overloads that are selected by the TARGET type: lfgpu::CsrMatrix / lfgpu::Vector live on the GPU.
//
// The shim is written against the reference's public interfaces only (Mesh::Entities / Index, Entity::RefEl /
// SubEntities / Geometry, Geometry::Global, DofHandler::NumDofs / NumLocalDofs / GlobalDofIndices, QuadRule::Points /
// Weights).  Because the spelling of a few members differs between LehrFEM++ and a stand-in mesh library, those calls
// go through a small ADAPTOR policy; INTEGRATION.md lists the adaptor for the real LehrFEM++ types.  The providers of
// the reference keep their coefficients private (uscalfe/loc_comp_ellbvp.h:178-188), so the shim ships providers with
// the SAME template and constructor signatures plus read-only accessors.
//
// Errors: a negative status of the C ABI becomes lfgpu::Error; LFGPU_ERR_MISSING_RULE is the reference's
// base::LfException case (loc_comp_ellbvp.h:278-287).  There is no CPU fallback.
#ifndef LF_GPU_SHIM_HPP
#define LF_GPU_SHIM_HPP

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "lfgpu.h"

namespace lfgpu {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

class Context {
 public:
  explicit Context(int device = 0) {
    const int rc = lfgpu_ctx_create(device, &ctx_);
    if (rc != 0) throw Error(rc, std::string("lfgpu_ctx_create: ") + lfgpu_last_error(nullptr));
  }
  ~Context() { lfgpu_ctx_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  [[nodiscard]] lfgpu_ctx* get() const { return ctx_; }
  void check(int rc, const char* where) const {
    if (rc != 0) throw Error(rc, std::string(where) + ": " + lfgpu_last_error(ctx_));
  }

 private:
  lfgpu_ctx* ctx_ = nullptr;
};

// ---- coefficient description -------------------------------------------------------------------------------------------
// What the device can evaluate: constants directly; anything else is tabulated by the shim at the quadrature points
// (mesh/utils/mesh_function_global.h:77-88 evaluates the user's functor at Geometry::Global(points) -- so does the shim,
// on the host, once per assembly; the table travels as LFGPU_COEFF_PER_QP).
struct HostCoeff {
  int kind = LFGPU_COEFF_CONST;
  double c[4] = {0, 0, 0, 0};
  std::vector<double> table;  // PER_QP / PER_QP_2X2 values, [n_cells][stride]([4])
  long stride = 0;
};

// Reference-element description handed to the C ABI: quadrature rules per cell type (null = provider default)
struct Rules {
  std::vector<double> pts_tria, w_tria, pts_quad, w_quad;
  bool has_tria = false, has_quad = false;
};

// ---- flattened mesh + dof tables (the shim's "flatten step") -------------------------------------------------------------
struct FlatMesh {
  std::vector<double> node_coords;     // [n_nodes][2]
  std::vector<std::uint32_t> cell_nodes;  // [n_cells][4]
  std::vector<double> cell_coords;     // [n_cells][4][2]
  bool coords_match_nodes = true;
  std::int64_t n_nodes = 0, n_cells = 0;
};

// ADAPTOR requirements (all static):
//   entities(mesh, codim) -> iterable of const ENTITY*;  num_entities(mesh, codim);  index(mesh, entity) -> unsigned
//   is_tria(entity) -> bool;  sub_entities(entity, rel_codim) -> iterable of const ENTITY*
//   corner(entity, k, d) -> double   (= entity.Geometry()->Global(RefEl().NodeCoords())(d, k))
//   num_dofs(dofh), num_local_dofs(dofh, entity), global_dof_indices(dofh, entity) -> indexable, mesh(dofh) -> const MESH&
template <class A, class MESH>
FlatMesh Flatten(const MESH& mesh) {
  FlatMesh f;
  f.n_nodes = A::num_entities(mesh, 2);
  f.n_cells = A::num_entities(mesh, 0);
  f.node_coords.assign(2 * f.n_nodes, 0.0);
  std::vector<char> seen(f.n_nodes, 0);
  f.cell_nodes.assign(4 * f.n_cells, LFGPU_IDX_NIL);
  f.cell_coords.assign(8 * f.n_cells, 0.0);
  for (const auto* cell : A::entities(mesh, 0)) {
    const std::int64_t c = A::index(mesh, *cell);
    const int nv = A::is_tria(*cell) ? 3 : 4;
    int k = 0;
    for (const auto* v : A::sub_entities(*cell, 2)) {
      const std::uint32_t n = A::index(mesh, *v);
      f.cell_nodes[4 * c + k] = n;
      const double x = A::corner(*cell, k, 0), y = A::corner(*cell, k, 1);
      f.cell_coords[8 * c + 2 * k] = x;
      f.cell_coords[8 * c + 2 * k + 1] = y;
      if (!seen[n]) {
        seen[n] = 1;
        f.node_coords[2 * n] = x;
        f.node_coords[2 * n + 1] = y;
      } else if (f.node_coords[2 * n] != x || f.node_coords[2 * n + 1] != y) {
        f.coords_match_nodes = false;  // e.g. refined meshes: child geometries differ from node objects by an ulp
      }
      ++k;
    }
    (void)nv;
  }
  return f;
}

template <class A, class DOFH, class MESH>
void FlattenDofs(const DOFH& dofh, const MESH& mesh, std::vector<std::int64_t>& cell_dofs, std::vector<std::uint8_t>& n_ldof, int& stride) {
  const std::int64_t n_cells = A::num_entities(mesh, 0);
  stride = 0;
  for (const auto* cell : A::entities(mesh, 0)) stride = std::max<int>(stride, A::num_local_dofs(dofh, *cell));
  cell_dofs.assign(static_cast<std::size_t>(n_cells) * stride, -1);
  n_ldof.assign(n_cells, 0);
  for (const auto* cell : A::entities(mesh, 0)) {
    const std::int64_t c = A::index(mesh, *cell);
    const int n = A::num_local_dofs(dofh, *cell);
    const auto idx = A::global_dof_indices(dofh, *cell);
    for (int k = 0; k < n; ++k) cell_dofs[c * stride + k] = idx[k];
    n_ldof[c] = static_cast<std::uint8_t>(n);
  }
}

// ---- Gmsh input ------------------------------------------------------------------------------------------------------------
// lf::io::GmshReader (lib/lf/io/gmsh_reader.h:55-200) with the reference's member names.  The reference identifies an
// entity by `const mesh::Entity&`; here an entity is (codim, index) with the indices reader.mesh() assigns -- the
// numbering lfgpu_gmsh_mesh reproduces on the device (nodes and cells in file order, explicitly listed edges first).
// Errors of the reference (base::LfException: unknown / ambiguous names, unreadable files) become lfgpu::Error.
class GmshReader {
 public:
  using size_type = unsigned int;
  using dim_t = int;  // the reference's default "codim not given" is static_cast<dim_t>(-1)
  explicit GmshReader(const std::string& filename, int dim_world = 2) {
    const int rc = lfgpu_gmsh_read_file(filename.c_str(), dim_world, &g_);
    if (rc != 0) throw Error(rc, std::string("GmshReader: ") + lfgpu_last_error(nullptr));
    lfgpu_gmsh_counts(g_, &n_nodes_, &n_explicit_edges_, &n_cells_, &order_, &n_names_);
  }
  GmshReader(const void* data, std::int64_t n_bytes, int dim_world = 2) {
    const int rc = lfgpu_gmsh_read_memory(data, n_bytes, dim_world, &g_);
    if (rc != 0) throw Error(rc, std::string("GmshReader: ") + lfgpu_last_error(nullptr));
    lfgpu_gmsh_counts(g_, &n_nodes_, &n_explicit_edges_, &n_cells_, &order_, &n_names_);
  }
  ~GmshReader() { lfgpu_gmsh_destroy(g_); }
  GmshReader(const GmshReader&) = delete;
  GmshReader& operator=(const GmshReader&) = delete;

  // reader.mesh(): the mesh on the device (caller owns the handle: lfgpu_mesh_destroy)
  [[nodiscard]] lfgpu_mesh* mesh(const Context& ctx) const {
    lfgpu_mesh* m = nullptr;
    ctx.check(lfgpu_gmsh_mesh(ctx.get(), g_, &m), "lfgpu_gmsh_mesh");
    return m;
  }
  // the arguments of the reference's AddPoint / AddEntity calls, for a caller that builds its own mesh object
  [[nodiscard]] FlatMesh Flat(std::vector<std::uint32_t>* explicit_edges = nullptr) const {
    FlatMesh f;
    f.n_nodes = n_nodes_;
    f.n_cells = n_cells_;
    f.node_coords.assign(2 * n_nodes_, 0.0);
    f.cell_nodes.assign(4 * n_cells_, LFGPU_IDX_NIL);
    std::vector<std::uint32_t> en(2 * n_explicit_edges_);
    lfgpu_gmsh_arrays(g_, f.node_coords.data(), en.data(), f.cell_nodes.data());
    if (explicit_edges != nullptr) *explicit_edges = std::move(en);
    return f;
  }
  [[nodiscard]] std::int64_t NumEntities(dim_t codim) const { return codim == 0 ? n_cells_ : (codim == 1 ? n_explicit_edges_ : n_nodes_); }
  [[nodiscard]] int GeometryOrder() const { return order_; }

  // gmsh_reader.cc:30-99
  [[nodiscard]] size_type PhysicalEntityName2Nr(const std::string& name, dim_t codim = -1) const {
    std::uint32_t nr = 0;
    const int rc = lfgpu_gmsh_physical_name2nr(g_, name.c_str(), codim, &nr);
    if (rc != 0) throw Error(rc, lfgpu_last_error(nullptr));
    return nr;
  }
  [[nodiscard]] std::string PhysicalEntityNr2Name(size_type number, dim_t codim = -1) const {
    char buf[512];
    const int rc = lfgpu_gmsh_physical_nr2name(g_, number, codim, buf, sizeof(buf));
    if (rc < 0) throw Error(rc, lfgpu_last_error(nullptr));
    return buf;
  }
  // gmsh_reader.cc:101-113
  [[nodiscard]] std::vector<std::pair<size_type, std::string>> PhysicalEntities(dim_t codim) const {
    std::vector<std::pair<size_type, std::string>> result;
    for (int i = 0; i < n_names_; ++i) {
      std::uint32_t nr = 0;
      int cd = 0;
      char buf[512];
      lfgpu_gmsh_physical_name(g_, i, &nr, &cd, buf, sizeof(buf));
      if (cd == codim) result.emplace_back(nr, buf);
    }
    return result;
  }
  // gmsh_reader.cc:17-23, 115-118 with the entity given as (codim, index)
  [[nodiscard]] std::vector<size_type> PhysicalEntityNr(dim_t codim, std::int64_t index) const {
    std::uint32_t tmp[32];
    const int n = lfgpu_gmsh_physical_entity_nr(g_, codim, index, 32, tmp);
    if (n < 0) throw Error(n, "PhysicalEntityNr: no such entity");
    std::vector<size_type> out(static_cast<std::size_t>(n));
    if (n > 32) lfgpu_gmsh_physical_entity_nr(g_, codim, index, n, out.data());
    else std::copy(tmp, tmp + n, out.begin());
    return out;
  }
  [[nodiscard]] bool IsPhysicalEntity(dim_t codim, std::int64_t index, size_type physical_entity_nr) const {
    const auto nrs = PhysicalEntityNr(codim, index);
    return std::find(nrs.begin(), nrs.end(), physical_entity_nr) != nrs.end();
  }
  // the same for all n entities of a codimension at once: the selector arrays the device calls take
  [[nodiscard]] std::vector<std::uint8_t> PhysicalEntityFlags(dim_t codim, size_type physical_entity_nr, std::int64_t n) const {
    std::vector<std::uint8_t> flags(static_cast<std::size_t>(n), 0);
    lfgpu_gmsh_physical_flags(g_, codim, physical_entity_nr, n, flags.data());
    return flags;
  }
  [[nodiscard]] const lfgpu_gmsh* get() const { return g_; }

 private:
  lfgpu_gmsh* g_ = nullptr;
  std::int64_t n_nodes_ = 0, n_explicit_edges_ = 0, n_cells_ = 0;
  int order_ = 1, n_names_ = 0;
};

// ---- device-side targets ---------------------------------------------------------------------------------------------------
// Compressed matrix on the GPU, the TMPMATRIX of the GPU overload.  Like COOMatrix it ACCUMULATES: AssembleMatrixLocally
// does not zero it (assembler.h:84-88); setZero() does.  makeSparse-like access: Download() returns Eigen-compatible
// column-major arrays (outer = column pointers) or CSR, as chosen at construction.
class CsrMatrix {
 public:
  CsrMatrix(Context& ctx, int major = LFGPU_COL_MAJOR) : ctx_(ctx), major_(major) {}
  ~CsrMatrix() {
    if (d_values_) lfgpu_free(ctx_.get(), d_values_);
    lfgpu_pattern_destroy(pattern_);
    lfgpu_dofmap_destroy(test_);
    if (trial_ != test_) lfgpu_dofmap_destroy(trial_);
    lfgpu_mesh_destroy(mesh_);
  }
  CsrMatrix(const CsrMatrix&) = delete;
  CsrMatrix& operator=(const CsrMatrix&) = delete;
  // movable, so that the returning forms of AssembleMatrixLocally (assembler.h:209-220, 243-249) can hand it out
  CsrMatrix(CsrMatrix&& o) noexcept
      : empty_(o.empty_), ctx_(o.ctx_), major_(o.major_), mesh_(o.mesh_), test_(o.test_), trial_(o.trial_), pattern_(o.pattern_),
        d_values_(o.d_values_) {
    o.mesh_ = nullptr;
    o.test_ = o.trial_ = nullptr;
    o.pattern_ = nullptr;
    o.d_values_ = nullptr;
  }
  void setZero() {
    if (d_values_) ctx_.check(lfgpu_memset(ctx_.get(), d_values_, 0, 8 * nnz()), "lfgpu_memset");
    empty_ = true;
  }
  [[nodiscard]] std::int64_t rows() const { return pattern_ ? lfgpu_pattern_rows(pattern_) : 0; }
  [[nodiscard]] std::int64_t cols() const { return pattern_ ? lfgpu_pattern_cols(pattern_) : 0; }
  [[nodiscard]] std::int64_t nnz() const { return pattern_ ? lfgpu_pattern_nnz(pattern_) : 0; }
  [[nodiscard]] double* device_values() const { return static_cast<double*>(d_values_); }
  [[nodiscard]] const lfgpu_pattern* pattern() const { return pattern_; }
  void Download(std::vector<std::int32_t>& outer, std::vector<std::int32_t>& inner, std::vector<double>& values) const {
    const std::int64_t n_outer = major_ == LFGPU_ROW_MAJOR ? rows() : cols();
    outer.resize(n_outer + 1);
    inner.resize(nnz());
    values.resize(nnz());
    ctx_.check(lfgpu_pattern_download(ctx_.get(), pattern_, outer.data(), inner.data()), "lfgpu_pattern_download");
    ctx_.check(lfgpu_memcpy_d2h(ctx_.get(), values.data(), d_values_, 8 * nnz()), "lfgpu_memcpy_d2h");
    ctx_.check(lfgpu_ctx_synchronize(ctx_.get()), "lfgpu_ctx_synchronize");
  }

  // internal: (re)build mesh / dof tables / pattern when first used (one-time symbolic pass)
  template <class A, class DOFH>
  void Prepare(const DOFH& trial, const DOFH& test) {
    if (pattern_ != nullptr) return;
    const auto& mesh = A::mesh(trial);
    const FlatMesh f = Flatten<A>(mesh);
    ctx_.check(lfgpu_mesh_upload(ctx_.get(), f.n_nodes, f.node_coords.data(), f.n_cells, f.cell_nodes.data(),
                                 f.coords_match_nodes ? nullptr : f.cell_coords.data(), &mesh_), "lfgpu_mesh_upload");
    std::vector<std::int64_t> dofs;
    std::vector<std::uint8_t> nl;
    int stride = 0;
    FlattenDofs<A>(test, mesh, dofs, nl, stride);
    ctx_.check(lfgpu_dofmap_upload(ctx_.get(), mesh_, A::num_dofs(test), stride, dofs.data(), nl.data(), &test_), "lfgpu_dofmap_upload");
    if (&trial == &test) {
      trial_ = test_;
    } else {
      FlattenDofs<A>(trial, mesh, dofs, nl, stride);
      ctx_.check(lfgpu_dofmap_upload(ctx_.get(), mesh_, A::num_dofs(trial), stride, dofs.data(), nl.data(), &trial_), "lfgpu_dofmap_upload");
    }
    ctx_.check(lfgpu_symbolic(ctx_.get(), mesh_, test_, trial_, major_, &pattern_), "lfgpu_symbolic");
    ctx_.check(lfgpu_malloc(ctx_.get(), 8 * nnz(), &d_values_), "lfgpu_malloc");
    setZero();
  }
  [[nodiscard]] Context& ctx() const { return ctx_; }
  [[nodiscard]] lfgpu_mesh* mesh() const { return mesh_; }
  [[nodiscard]] lfgpu_dofmap* test_dofs() const { return test_; }
  bool empty_ = true;

 private:
  Context& ctx_;
  int major_;
  lfgpu_mesh* mesh_ = nullptr;
  lfgpu_dofmap *test_ = nullptr, *trial_ = nullptr;
  lfgpu_pattern* pattern_ = nullptr;
  void* d_values_ = nullptr;
};

// Dense vector on the GPU, the VECTOR of the GPU overload of AssembleVectorLocally (accumulates, assembler.h:291-293)
class Vector {
 public:
  explicit Vector(Context& ctx) : ctx_(ctx) {}
  ~Vector() {
    if (d_) lfgpu_free(ctx_.get(), d_);
    lfgpu_dofmap_destroy(dofs_);
    lfgpu_mesh_destroy(mesh_);
  }
  Vector(const Vector&) = delete;
  Vector& operator=(const Vector&) = delete;
  Vector(Vector&& o) noexcept : ctx_(o.ctx_), mesh_(o.mesh_), dofs_(o.dofs_), d_(o.d_), n_(o.n_) {  // assembler.h:354-365 returns the vector
    o.mesh_ = nullptr;
    o.dofs_ = nullptr;
    o.d_ = nullptr;
    o.n_ = 0;
  }
  void setZero() {
    if (d_) ctx_.check(lfgpu_memset(ctx_.get(), d_, 0, 8 * n_), "lfgpu_memset");
  }
  [[nodiscard]] std::int64_t size() const { return n_; }
  [[nodiscard]] std::vector<double> Download() const {
    std::vector<double> h(n_);
    ctx_.check(lfgpu_memcpy_d2h(ctx_.get(), h.data(), d_, 8 * n_), "lfgpu_memcpy_d2h");
    ctx_.check(lfgpu_ctx_synchronize(ctx_.get()), "lfgpu_ctx_synchronize");
    return h;
  }
  template <class A, class DOFH>
  void Prepare(const DOFH& dofh) {
    if (dofs_ != nullptr) return;
    const auto& mesh = A::mesh(dofh);
    const FlatMesh f = Flatten<A>(mesh);
    ctx_.check(lfgpu_mesh_upload(ctx_.get(), f.n_nodes, f.node_coords.data(), f.n_cells, f.cell_nodes.data(),
                                 f.coords_match_nodes ? nullptr : f.cell_coords.data(), &mesh_), "lfgpu_mesh_upload");
    std::vector<std::int64_t> dofs;
    std::vector<std::uint8_t> nl;
    int stride = 0;
    FlattenDofs<A>(dofh, mesh, dofs, nl, stride);
    n_ = A::num_dofs(dofh);
    ctx_.check(lfgpu_dofmap_upload(ctx_.get(), mesh_, n_, stride, dofs.data(), nl.data(), &dofs_), "lfgpu_dofmap_upload");
    ctx_.check(lfgpu_malloc(ctx_.get(), 8 * n_, &d_), "lfgpu_malloc");
    setZero();
  }
  [[nodiscard]] Context& ctx() const { return ctx_; }
  [[nodiscard]] lfgpu_mesh* mesh() const { return mesh_; }
  [[nodiscard]] lfgpu_dofmap* dofs() const { return dofs_; }
  [[nodiscard]] double* device() const { return static_cast<double*>(d_); }

 private:
  Context& ctx_;
  lfgpu_mesh* mesh_ = nullptr;
  lfgpu_dofmap* dofs_ = nullptr;
  void* d_ = nullptr;
  std::int64_t n_ = 0;
};

// ---- mesh functions the device understands ---------------------------------------------------------------------------------
// Same names and call operators as lf::mesh::utils::MeshFunctionConstant / MeshFunctionGlobal
// (mesh/utils/mesh_function_constant.h:26-46, mesh_function_global.h:55-98) are NOT required: the shim only needs to
// know how to turn a coefficient object into a HostCoeff.  Specialise CoeffTraits for further mesh-function types.
struct Matrix2 {
  double a[2][2];
};
template <class R>
struct MeshFunctionConstant {
  R value;
};
// Point type handed to functors that take ONE argument, like the reference's MeshFunctionGlobal (mesh_function_global.h:77-88:
// f(Eigen::Vector2d)).  With Eigen available define LFGPU_SHIM_POINT_TYPE=Eigen::Vector2d before including this header and
// the user's lambdas compile unchanged; Vec2 offers the members such lambdas typically use.
struct Vec2 {
  double v[2];
  Vec2(double x, double y) : v{x, y} {}
  double operator[](int i) const { return v[i]; }
  double operator()(int i) const { return v[i]; }
  [[nodiscard]] double x() const { return v[0]; }
  [[nodiscard]] double y() const { return v[1]; }
  [[nodiscard]] double squaredNorm() const { return v[0] * v[0] + v[1] * v[1]; }
  [[nodiscard]] double norm() const { return std::sqrt(squaredNorm()); }
};
#ifndef LFGPU_SHIM_POINT_TYPE
#define LFGPU_SHIM_POINT_TYPE ::lfgpu::Vec2
#endif
template <class F>
struct MeshFunctionGlobal {
  F f;  // R f(double x, double y)  or  R f(point) with point = LFGPU_SHIM_POINT_TYPE;  R = double or Matrix2
  [[nodiscard]] auto operator()(double x, double y) const {
    if constexpr (std::is_invocable_v<const F&, double, double>) {
      return f(x, y);
    } else {
      return f(LFGPU_SHIM_POINT_TYPE(x, y));
    }
  }
};

// A continuous piecewise (bi)linear coefficient given by its values at the mesh nodes, in node-index order: what
// lf::fe::MeshFunctionFE(fe_space_o1, coeff_vector) evaluates (fe/mesh_function_fe.h).  Travels as LFGPU_COEFF_NODAL: 8 bytes per
// node instead of 8 bytes per quadrature point.
struct MeshFunctionNodal {
  std::vector<double> values;  // [n_nodes]
};

template <class MF>
struct CoeffTraits;  // static HostCoeff describe(const MF&, const double* qp_xy /*[n_cells][stride][2]*/, n_cells, stride)

template <>
struct CoeffTraits<MeshFunctionConstant<double>> {
  static constexpr bool needs_points = false;
  static HostCoeff describe(const MeshFunctionConstant<double>& mf, const double*, std::int64_t, long) {
    HostCoeff c;
    c.kind = LFGPU_COEFF_CONST;
    c.c[0] = mf.value;
    return c;
  }
};
template <>
struct CoeffTraits<MeshFunctionConstant<Matrix2>> {
  static constexpr bool needs_points = false;
  static HostCoeff describe(const MeshFunctionConstant<Matrix2>& mf, const double*, std::int64_t, long) {
    HostCoeff c;
    c.kind = LFGPU_COEFF_CONST_2X2;
    c.c[0] = mf.value.a[0][0]; c.c[1] = mf.value.a[0][1]; c.c[2] = mf.value.a[1][0]; c.c[3] = mf.value.a[1][1];
    return c;
  }
};
template <>
struct CoeffTraits<MeshFunctionNodal> {
  static constexpr bool needs_points = false;
  static HostCoeff describe(const MeshFunctionNodal& mf, const double*, std::int64_t, long) {
    HostCoeff c;
    c.kind = LFGPU_COEFF_NODAL;
    c.table = mf.values;
    c.stride = 1;
    return c;
  }
};
template <class F>
struct CoeffTraits<MeshFunctionGlobal<F>> {
  static constexpr bool needs_points = true;
  using R = decltype(std::declval<const MeshFunctionGlobal<F>&>()(0.0, 0.0));
  static HostCoeff describe(const MeshFunctionGlobal<F>& mf, const double* xy, std::int64_t n_cells, long stride) {
    HostCoeff c;
    c.stride = stride;
    if constexpr (std::is_same_v<R, Matrix2>) {
      c.kind = LFGPU_COEFF_PER_QP_2X2;
      c.table.resize(static_cast<std::size_t>(n_cells) * stride * 4);
      for (std::int64_t i = 0; i < n_cells * stride; ++i) {
        const Matrix2 m = mf(xy[2 * i], xy[2 * i + 1]);
        c.table[4 * i] = m.a[0][0]; c.table[4 * i + 1] = m.a[0][1]; c.table[4 * i + 2] = m.a[1][0]; c.table[4 * i + 3] = m.a[1][1];
      }
    } else {
      c.kind = LFGPU_COEFF_PER_QP;
      c.table.resize(static_cast<std::size_t>(n_cells) * stride);
      for (std::int64_t i = 0; i < n_cells * stride; ++i) c.table[i] = mf(xy[2 * i], xy[2 * i + 1]);
    }
    return c;
  }
};

// ---- providers: same template / constructor signatures as the reference, plus accessors -----------------------------------
// lf::uscalfe::ReactionDiffusionElementMatrixProvider<SCALAR, DIFF_COEFF, REACTION_COEFF> (loc_comp_ellbvp.h:85-189):
//   ctor (fe_space, alpha, gamma)            -> default rules of degree 2 * fe->Degree()   (:210-231)
//   ctor (fe_space, alpha, gamma, qr_map)    -> user rules per RefEl                        (:234-263)
// FE_SPACE must offer Degree() (1..3 = FeSpaceLagrangeO1/O2/O3) and LocGlobMap().
template <class SCALAR, class DIFF_COEFF, class REACTION_COEFF>
class ReactionDiffusionElementMatrixProvider {
 public:
  static_assert(std::is_same_v<SCALAR, double>, "the GPU path computes in double (the reference's default scalar)");
  template <class FE_SPACE>
  ReactionDiffusionElementMatrixProvider(std::shared_ptr<const FE_SPACE> fe_space, DIFF_COEFF alpha, REACTION_COEFF gamma)
      : alpha_(std::move(alpha)), gamma_(std::move(gamma)), degree_(fe_space->Degree()) {}
  template <class FE_SPACE, class QR_MAP>
  ReactionDiffusionElementMatrixProvider(std::shared_ptr<const FE_SPACE> fe_space, DIFF_COEFF alpha, REACTION_COEFF gamma,
                                         const QR_MAP& qr_collection)
      : alpha_(std::move(alpha)), gamma_(std::move(gamma)), degree_(fe_space->Degree()) {
    custom_rules_ = true;
    for (const auto& kv : qr_collection) AddRule(kv.first.Id(), kv.second);
  }
  ReactionDiffusionElementMatrixProvider(const ReactionDiffusionElementMatrixProvider&) = delete;
  ReactionDiffusionElementMatrixProvider(ReactionDiffusionElementMatrixProvider&&) noexcept = default;
  using alpha_type = DIFF_COEFF;
  using gamma_type = REACTION_COEFF;
  // loc_comp_ellbvp.h:155: "all cells are active"; a derived provider that declares its own isActive(cell) is honoured by the GPU
  // overload (assembler.h:127 skips inactive cells): the predicate is evaluated on the host, the mask travels to the device
  template <class CELL>
  bool isActive(const CELL& /*cell*/) { return true; }
  [[nodiscard]] const DIFF_COEFF& Alpha() const { return alpha_; }
  [[nodiscard]] const REACTION_COEFF& Gamma() const { return gamma_; }
  [[nodiscard]] int Degree() const { return degree_; }
  [[nodiscard]] const Rules& QuadRules() const { return rules_; }
  [[nodiscard]] bool HasCustomRules() const { return custom_rules_; }

 private:
  template <class QR>
  void AddRule(unsigned ref_el_id, const QR& qr) {
    const int n = static_cast<int>(qr.NumPoints());
    auto& pts = ref_el_id == 3 ? rules_.pts_tria : rules_.pts_quad;
    auto& w = ref_el_id == 3 ? rules_.w_tria : rules_.w_quad;
    pts.resize(2 * n);
    w.resize(n);
    for (int k = 0; k < n; ++k) {
      pts[k] = qr.Points()(0, k);
      pts[n + k] = qr.Points()(1, k);
      w[k] = qr.Weights()[k];
    }
    (ref_el_id == 3 ? rules_.has_tria : rules_.has_quad) = true;
  }
  DIFF_COEFF alpha_;
  REACTION_COEFF gamma_;
  int degree_;
  Rules rules_;
  bool custom_rules_ = false;
};

// lf::uscalfe::ScalarLoadElementVectorProvider<SCALAR, MESH_FUNCTION> (loc_comp_ellbvp.h:562-624, ctors :638-686)
template <class SCALAR, class MESH_FUNCTION>
class ScalarLoadElementVectorProvider {
 public:
  static_assert(std::is_same_v<SCALAR, double>, "the GPU path computes in double");
  template <class FE_SPACE>
  ScalarLoadElementVectorProvider(std::shared_ptr<const FE_SPACE> fe_space, MESH_FUNCTION f) : f_(std::move(f)), degree_(fe_space->Degree()) {}
  // loc_comp_ellbvp.h:660-686: user rules per reference element
  template <class FE_SPACE, class QR_MAP>
  ScalarLoadElementVectorProvider(std::shared_ptr<const FE_SPACE> fe_space, MESH_FUNCTION f, const QR_MAP& qr_collection)
      : f_(std::move(f)), degree_(fe_space->Degree()) {
    custom_rules_ = true;
    for (const auto& kv : qr_collection) {
      const auto& qr = kv.second;
      const int n = static_cast<int>(qr.NumPoints());
      const bool tria = kv.first.Id() == 3;
      auto& pts = tria ? rules_.pts_tria : rules_.pts_quad;
      auto& w = tria ? rules_.w_tria : rules_.w_quad;
      pts.resize(2 * n);
      w.resize(n);
      for (int k = 0; k < n; ++k) {
        pts[k] = qr.Points()(0, k);
        pts[n + k] = qr.Points()(1, k);
        w[k] = qr.Weights()[k];
      }
      (tria ? rules_.has_tria : rules_.has_quad) = true;
    }
  }
  using function_type = MESH_FUNCTION;
  template <class CELL>
  bool isActive(const CELL& /*cell*/) { return true; }  // loc_comp_ellbvp.h:608
  [[nodiscard]] const MESH_FUNCTION& F() const { return f_; }
  [[nodiscard]] int Degree() const { return degree_; }
  [[nodiscard]] const Rules& QuadRules() const { return rules_; }
  [[nodiscard]] bool HasCustomRules() const { return custom_rules_; }

 private:
  MESH_FUNCTION f_;
  int degree_;
  Rules rules_;
  bool custom_rules_ = false;
};

namespace detail {
struct DeviceCoeff {
  lfgpu_coeff c{};
  void* d_table = nullptr;
  lfgpu_ctx* ctx = nullptr;
  ~DeviceCoeff() {
    if (d_table) lfgpu_free(ctx, d_table);
  }
};
inline void to_device(Context& ctx, const HostCoeff& h, DeviceCoeff& d) {
  d.ctx = ctx.get();
  d.c.kind = h.kind;
  for (int i = 0; i < 4; ++i) d.c.c[i] = h.c[i];
  d.c.stride = h.stride;
  d.c.data = nullptr;
  if (!h.table.empty()) {
    ctx.check(lfgpu_malloc(ctx.get(), 8 * static_cast<std::int64_t>(h.table.size()), &d.d_table), "lfgpu_malloc");
    ctx.check(lfgpu_memcpy_h2d(ctx.get(), d.d_table, h.table.data(), 8 * static_cast<std::int64_t>(h.table.size())), "lfgpu_memcpy_h2d");
    d.c.data = static_cast<const double*>(d.d_table);
  }
}
// global coordinates of all quadrature points, computed on the device (Geometry::Global), for host-evaluated functors
inline std::vector<double> qp_coords(Context& ctx, lfgpu_mesh* mesh, std::int64_t n_cells, int degree, const lfgpu_quad* qt,
                                     const lfgpu_quad* qq, int stride) {
  std::vector<double> xy(static_cast<std::size_t>(n_cells) * stride * 2);
  void* d = nullptr;
  ctx.check(lfgpu_malloc(ctx.get(), 8 * static_cast<std::int64_t>(xy.size()), &d), "lfgpu_malloc");
  ctx.check(lfgpu_qp_coords(ctx.get(), mesh, degree, qt, qq, stride, static_cast<double*>(d)), "lfgpu_qp_coords");
  ctx.check(lfgpu_memcpy_d2h(ctx.get(), xy.data(), d, 8 * static_cast<std::int64_t>(xy.size())), "lfgpu_memcpy_d2h");
  ctx.check(lfgpu_ctx_synchronize(ctx.get()), "lfgpu_ctx_synchronize");
  lfgpu_free(ctx.get(), d);
  return xy;
}
// provider.isActive(cell) for every cell (assembler.h:127), evaluated on the host like every user predicate; null if all active
struct DeviceMask {
  void* d = nullptr;
  lfgpu_ctx* ctx = nullptr;
  ~DeviceMask() {
    if (d) lfgpu_free(ctx, d);
  }
  [[nodiscard]] const std::uint8_t* get() const { return static_cast<const std::uint8_t*>(d); }
};
template <class A, class MESH, class PROVIDER>
std::vector<std::uint8_t> activity_flags(const MESH& mesh, PROVIDER& prov, bool* all_active) {
  std::vector<std::uint8_t> act(static_cast<std::size_t>(A::num_entities(mesh, 0)), 1);
  *all_active = true;
  for (const auto* cell : A::entities(mesh, 0)) {
    if (!prov.isActive(*cell)) {
      act[A::index(mesh, *cell)] = 0;
      *all_active = false;
    }
  }
  return act;
}
template <class A, class MESH, class PROVIDER>
void activity_mask(Context& ctx, const MESH& mesh, PROVIDER& prov, DeviceMask& m) {
  bool all = true;
  const auto act = activity_flags<A>(mesh, prov, &all);
  if (all) return;
  m.ctx = ctx.get();
  ctx.check(lfgpu_malloc(ctx.get(), static_cast<std::int64_t>(act.size()), &m.d), "lfgpu_malloc");
  ctx.check(lfgpu_memcpy_h2d(ctx.get(), m.d, act.data(), static_cast<std::int64_t>(act.size())), "lfgpu_memcpy_h2d");
  ctx.check(lfgpu_ctx_synchronize(ctx.get()), "lfgpu_ctx_synchronize");  // `act` goes out of scope
}
}  // namespace detail

inline constexpr int kMaxPoints = 36;  // table stride for host-evaluated coefficients (largest rule the library holds)

// ---- the overloads ------------------------------------------------------------------------------------------------------------
// GPU overload of lf::assemble::AssembleMatrixLocally (assembler.h:114-186): same arguments, TMPMATRIX = lfgpu::CsrMatrix.
// PROVIDER: ReactionDiffusionElementMatrixProvider<double, ALPHA, GAMMA> or a class derived from it (its own isActive is used).
template <class A, class DOFH, class PROVIDER, class ALPHA = typename PROVIDER::alpha_type, class GAMMA = typename PROVIDER::gamma_type>
void AssembleMatrixLocally(unsigned codim, const DOFH& dof_handler_trial, const DOFH& dof_handler_test, PROVIDER& emp, CsrMatrix& matrix) {
  if (codim != 0) throw Error(LFGPU_ERR_UNSUPPORTED, "the GPU overload assembles cell (codim 0) contributions");
  if (&A::mesh(dof_handler_trial) != &A::mesh(dof_handler_test))
    throw Error(LFGPU_ERR_INVALID, "Trial and test space must be defined on the same mesh");  // assembler.h:121-122
  Context& ctx = matrix.ctx();
  matrix.template Prepare<A>(dof_handler_trial, dof_handler_test);
  const Rules& r = emp.QuadRules();
  lfgpu_quad qt{static_cast<int>(r.w_tria.size()), r.pts_tria.data(), r.w_tria.data()};
  lfgpu_quad qq{static_cast<int>(r.w_quad.size()), r.pts_quad.data(), r.w_quad.data()};
  const lfgpu_quad* pqt = (emp.HasCustomRules() && r.has_tria) ? &qt : nullptr;
  const lfgpu_quad* pqq = (emp.HasCustomRules() && r.has_quad) ? &qq : nullptr;
  std::int64_t n_cells = 0;
  lfgpu_mesh_counts(matrix.mesh(), nullptr, nullptr, &n_cells, nullptr, nullptr);
  std::vector<double> xy;
  int stride = 0;
  if (CoeffTraits<ALPHA>::needs_points || CoeffTraits<GAMMA>::needs_points) {
    stride = kMaxPoints;
    xy = detail::qp_coords(ctx, matrix.mesh(), n_cells, emp.Degree(), pqt, pqq, stride);
  }
  detail::DeviceCoeff da, dg;
  detail::to_device(ctx, CoeffTraits<ALPHA>::describe(emp.Alpha(), xy.data(), n_cells, stride), da);
  detail::to_device(ctx, CoeffTraits<GAMMA>::describe(emp.Gamma(), xy.data(), n_cells, stride), dg);
  detail::DeviceMask active;
  detail::activity_mask<A>(ctx, A::mesh(dof_handler_trial), emp, active);
  // accumulate like the reference (assembler.h:84-88); a freshly zeroed matrix is simply overwritten
  const double beta = matrix.empty_ ? 0.0 : 1.0;
  ctx.check(lfgpu_assemble_reaction_diffusion(ctx.get(), matrix.mesh(), matrix.pattern(), emp.Degree(), pqt, pqq, &da.c, &dg.c,
                                              active.get(), beta, matrix.device_values(), LFGPU_ALGO_AUTO),
            "lfgpu_assemble_reaction_diffusion");
  ctx.check(lfgpu_ctx_synchronize(ctx.get()), "lfgpu_ctx_synchronize");
  matrix.empty_ = false;
}
// GPU overload of lf::assemble::AssembleVectorLocally (assembler.h:298-327): VECTOR = lfgpu::Vector
template <class A, class DOFH, class PROVIDER, class F = typename PROVIDER::function_type>
void AssembleVectorLocally(unsigned codim, const DOFH& dof_handler, PROVIDER& evp, Vector& v) {
  if (codim != 0) throw Error(LFGPU_ERR_UNSUPPORTED, "the GPU overload assembles cell (codim 0) contributions");
  Context& ctx = v.ctx();
  v.template Prepare<A>(dof_handler);
  std::int64_t n_cells = 0;
  lfgpu_mesh_counts(v.mesh(), nullptr, nullptr, &n_cells, nullptr, nullptr);
  const Rules& r = evp.QuadRules();
  lfgpu_quad qt{static_cast<int>(r.w_tria.size()), r.pts_tria.data(), r.w_tria.data()};
  lfgpu_quad qq{static_cast<int>(r.w_quad.size()), r.pts_quad.data(), r.w_quad.data()};
  const lfgpu_quad* pqt = (evp.HasCustomRules() && r.has_tria) ? &qt : nullptr;
  const lfgpu_quad* pqq = (evp.HasCustomRules() && r.has_quad) ? &qq : nullptr;
  std::vector<double> xy;
  int stride = 0;
  if (CoeffTraits<F>::needs_points) {
    stride = 36;
    xy = detail::qp_coords(ctx, v.mesh(), n_cells, evp.Degree(), pqt, pqq, stride);
  }
  detail::DeviceCoeff df;
  detail::to_device(ctx, CoeffTraits<F>::describe(evp.F(), xy.data(), n_cells, stride), df);
  detail::DeviceMask active;
  detail::activity_mask<A>(ctx, A::mesh(dof_handler), evp, active);
  ctx.check(lfgpu_assemble_load(ctx.get(), v.mesh(), v.dofs(), evp.Degree(), pqt, pqq, &df.c, active.get(), 1.0, v.device(), LFGPU_ALGO_AUTO),
            "lfgpu_assemble_load");
  ctx.check(lfgpu_ctx_synchronize(ctx.get()), "lfgpu_ctx_synchronize");
}

// lf::fe::DiffusionElementMatrixProvider / MassElementMatrixProvider (lib/lf/fe/loc_comp_ellbvp.h:76-226, 256-384), the
// providers of the generic lf::fe module, for the Lagrange spaces the device path knows: same constructor shape
// (fe_space, coefficient), rule of degree 2 * Degree() (:186, :365) = the uscalfe provider's default.  They are the
// reaction-diffusion provider with the other coefficient identically zero -- bitwise so in the reference's arithmetic
// (w * (x + 0) = w * x; checked on the oracle, tests/test_oracle_fe_providers.py) -- hence take every overload above.
namespace fe {
template <class SCALAR, class DIFF_COEFF>
class DiffusionElementMatrixProvider : public ReactionDiffusionElementMatrixProvider<SCALAR, DIFF_COEFF, MeshFunctionConstant<double>> {
 public:
  template <class FE_SPACE>
  DiffusionElementMatrixProvider(std::shared_ptr<const FE_SPACE> fe_space, DIFF_COEFF alpha)
      : ReactionDiffusionElementMatrixProvider<SCALAR, DIFF_COEFF, MeshFunctionConstant<double>>(std::move(fe_space), std::move(alpha),
                                                                                                   MeshFunctionConstant<double>{0.0}) {}
};
template <class SCALAR, class REACTION_COEFF>
class MassElementMatrixProvider : public ReactionDiffusionElementMatrixProvider<SCALAR, MeshFunctionConstant<double>, REACTION_COEFF> {
 public:
  template <class FE_SPACE>
  MassElementMatrixProvider(std::shared_ptr<const FE_SPACE> fe_space, REACTION_COEFF gamma)
      : ReactionDiffusionElementMatrixProvider<SCALAR, MeshFunctionConstant<double>, REACTION_COEFF>(
            std::move(fe_space), MeshFunctionConstant<double>{0.0}, std::move(gamma)) {}
};
// lf::fe::ScalarLoadElementVectorProvider (:569-700) has the uscalfe provider's form and rule
template <class SCALAR, class MESH_FUNCTION>
using ScalarLoadElementVectorProvider = ::lfgpu::ScalarLoadElementVectorProvider<SCALAR, MESH_FUNCTION>;
}  // namespace fe

// The returning forms of the reference (assembler.h:209-220 two handlers, :243-249 one handler, :354-365 vector): the
// reference builds TMPMATRIX{rows, cols} / VECTOR(size) itself; a device target needs its context, the one extra argument.
template <class A, class DOFH, class PROVIDER>
CsrMatrix AssembleMatrixLocally(Context& ctx, unsigned codim, const DOFH& dof_handler_trial, const DOFH& dof_handler_test,
                                PROVIDER& entity_matrix_provider, int major = LFGPU_COL_MAJOR) {
  CsrMatrix matrix(ctx, major);
  AssembleMatrixLocally<A>(codim, dof_handler_trial, dof_handler_test, entity_matrix_provider, matrix);
  return matrix;
}
template <class A, class DOFH, class PROVIDER>
CsrMatrix AssembleMatrixLocally(Context& ctx, unsigned codim, const DOFH& dof_handler, PROVIDER& entity_matrix_provider,
                                int major = LFGPU_COL_MAJOR) {
  return AssembleMatrixLocally<A>(ctx, codim, dof_handler, dof_handler, entity_matrix_provider, major);
}
template <class A, class DOFH, class PROVIDER>
Vector AssembleVectorLocally(Context& ctx, unsigned codim, const DOFH& dof_handler, PROVIDER& entity_vector_provider) {
  Vector v(ctx);
  AssembleVectorLocally<A>(codim, dof_handler, entity_vector_provider, v);
  return v;
}

// ---- several GPUs (include/lfgpu.h: lfgpu_multi_*) ------------------------------------------------------------------------------
// TMPMATRIX for a device LIST: the GPU overload below cuts the problem into one sub-problem per device (Morton cell ranges, every
// device owns the rows of its cells, one-cell halo), assembles on all devices at once and leaves the rows where they were computed.
// Accumulates like COOMatrix (assembler.h:84-88).  Part(k) hands out device k's rows with GLOBAL indices; Gather() concatenates
// the parts into one compressed matrix in the reference's row order (for checks and for callers that want makeSparse()'s result).
class MultiCsrMatrix {
 public:
  explicit MultiCsrMatrix(const std::vector<int>& device_ids, int major = LFGPU_COL_MAJOR) : major_(major) {
    const int rc = lfgpu_multi_create(device_ids.data(), static_cast<int>(device_ids.size()), &m_);
    if (rc != 0) throw Error(rc, std::string("lfgpu_multi_create: ") + lfgpu_last_error(nullptr));
  }
  ~MultiCsrMatrix() { lfgpu_multi_destroy(m_); }
  MultiCsrMatrix(const MultiCsrMatrix&) = delete;
  MultiCsrMatrix& operator=(const MultiCsrMatrix&) = delete;
  void check(int rc, const char* where) const {
    if (rc != 0) throw Error(rc, std::string(where) + ": " + lfgpu_multi_last_error(m_));
  }
  [[nodiscard]] int NumDevices() const { return lfgpu_multi_num_devices(m_); }
  [[nodiscard]] lfgpu_multi* get() const { return m_; }
  void setZero() {
    if (ready_) check(lfgpu_multi_set_zero(m_), "lfgpu_multi_set_zero");
    empty_ = true;
  }
  struct Block {
    std::vector<std::int64_t> rows, row_ptr;  // global outer indices (ascending), offsets into cols / values
    std::vector<std::int32_t> cols;           // global inner indices, ascending inside a row
    std::vector<double> values;
  };
  [[nodiscard]] Block Part(int k) const {
    Block b;
    std::int64_t nr = 0, nnz = 0;
    check(lfgpu_multi_part_sizes(m_, k, &nr, &nnz, nullptr, nullptr, nullptr), "lfgpu_multi_part_sizes");
    b.rows.resize(nr);
    b.row_ptr.resize(nr + 1);
    b.cols.resize(nnz);
    b.values.resize(nnz);
    check(lfgpu_multi_part_download(m_, k, b.rows.data(), b.row_ptr.data(), b.cols.data(), b.values.data()), "lfgpu_multi_part_download");
    return b;
  }
  // all parts as ONE compressed matrix with the reference's index arrays (outer [n + 1], inner, values)
  void Gather(std::vector<std::int64_t>& outer, std::vector<std::int32_t>& inner, std::vector<double>& values) const {
    outer.assign(n_outer_ + 1, 0);
    std::vector<Block> blocks;
    for (int k = 0; k < NumDevices(); ++k) blocks.push_back(Part(k));
    for (const auto& b : blocks)
      for (std::size_t i = 0; i < b.rows.size(); ++i) outer[b.rows[i] + 1] = b.row_ptr[i + 1] - b.row_ptr[i];
    for (std::int64_t r = 0; r < n_outer_; ++r) outer[r + 1] += outer[r];
    inner.resize(outer[n_outer_]);
    values.resize(outer[n_outer_]);
    for (const auto& b : blocks)
      for (std::size_t i = 0; i < b.rows.size(); ++i) {
        std::copy(b.cols.begin() + b.row_ptr[i], b.cols.begin() + b.row_ptr[i + 1], inner.begin() + outer[b.rows[i]]);
        std::copy(b.values.begin() + b.row_ptr[i], b.values.begin() + b.row_ptr[i + 1], values.begin() + outer[b.rows[i]]);
      }
  }
  template <class A, class DOFH>
  void Prepare(const DOFH& trial, const DOFH& test) {
    if (ready_) return;
    if (&trial != &test) throw Error(LFGPU_ERR_UNSUPPORTED, "the multi-device matrix needs trial space == test space");
    const auto& mesh = A::mesh(test);
    flat_ = Flatten<A>(mesh);
    std::vector<std::int64_t> dofs;
    std::vector<std::uint8_t> nl;
    int stride = 0;
    FlattenDofs<A>(test, mesh, dofs, nl, stride);
    n_outer_ = A::num_dofs(test);
    check(lfgpu_multi_setup(m_, flat_.n_nodes, flat_.node_coords.data(), flat_.n_cells, flat_.cell_nodes.data(),
                            flat_.coords_match_nodes ? nullptr : flat_.cell_coords.data(), n_outer_, stride, dofs.data(), nl.data(), major_),
          "lfgpu_multi_setup");
    ready_ = true;
    empty_ = true;
  }
  [[nodiscard]] const FlatMesh& flat() const { return flat_; }
  bool empty_ = true;

 private:
  lfgpu_multi* m_ = nullptr;
  int major_;
  bool ready_ = false;
  std::int64_t n_outer_ = 0;
  FlatMesh flat_;
};

// GPU overload of AssembleMatrixLocally for a device list.  Coefficient functors are tabulated on the host at the quadrature points
// of ALL cells (on device 0, which sees the whole mesh for the time of that call); every device then receives its cells' entries.
template <class A, class DOFH, class PROVIDER, class ALPHA = typename PROVIDER::alpha_type, class GAMMA = typename PROVIDER::gamma_type>
void AssembleMatrixLocally(unsigned codim, const DOFH& dof_handler_trial, const DOFH& dof_handler_test, PROVIDER& emp, MultiCsrMatrix& matrix) {
  if (codim != 0) throw Error(LFGPU_ERR_UNSUPPORTED, "the GPU overload assembles cell (codim 0) contributions");
  matrix.template Prepare<A>(dof_handler_trial, dof_handler_test);
  const Rules& r = emp.QuadRules();
  lfgpu_quad qt{static_cast<int>(r.w_tria.size()), r.pts_tria.data(), r.w_tria.data()};
  lfgpu_quad qq{static_cast<int>(r.w_quad.size()), r.pts_quad.data(), r.w_quad.data()};
  const lfgpu_quad* pqt = (emp.HasCustomRules() && r.has_tria) ? &qt : nullptr;
  const lfgpu_quad* pqq = (emp.HasCustomRules() && r.has_quad) ? &qq : nullptr;
  bool all_active = true;
  detail::activity_flags<A>(A::mesh(dof_handler_trial), emp, &all_active);
  if (!all_active) throw Error(LFGPU_ERR_UNSUPPORTED, "the multi-device overload assembles all cells (isActive must be true)");
  const FlatMesh& f = matrix.flat();
  std::vector<double> xy;
  int stride = 0;
  if (CoeffTraits<ALPHA>::needs_points || CoeffTraits<GAMMA>::needs_points) {
    stride = kMaxPoints;
    lfgpu_ctx* c0 = lfgpu_multi_ctx(matrix.get(), 0);
    lfgpu_mesh* whole = nullptr;
    int rc = lfgpu_mesh_upload(c0, f.n_nodes, f.node_coords.data(), f.n_cells, f.cell_nodes.data(),
                               f.coords_match_nodes ? nullptr : f.cell_coords.data(), &whole);
    if (rc != 0) throw Error(rc, std::string("lfgpu_mesh_upload: ") + lfgpu_last_error(c0));
    xy.resize(static_cast<std::size_t>(f.n_cells) * stride * 2);
    void* d = nullptr;
    rc = lfgpu_malloc(c0, 8 * static_cast<std::int64_t>(xy.size()), &d);
    if (rc == 0) rc = lfgpu_qp_coords(c0, whole, emp.Degree(), pqt, pqq, stride, static_cast<double*>(d));
    if (rc == 0) rc = lfgpu_memcpy_d2h(c0, xy.data(), d, 8 * static_cast<std::int64_t>(xy.size()));
    if (rc == 0) rc = lfgpu_ctx_synchronize(c0);
    if (d) lfgpu_free(c0, d);
    lfgpu_mesh_destroy(whole);
    if (rc != 0) throw Error(rc, std::string("quadrature points: ") + lfgpu_last_error(c0));
  }
  const HostCoeff ha = CoeffTraits<ALPHA>::describe(emp.Alpha(), xy.data(), f.n_cells, stride);
  const HostCoeff hg = CoeffTraits<GAMMA>::describe(emp.Gamma(), xy.data(), f.n_cells, stride);
  auto as_c = [](const HostCoeff& h) {
    lfgpu_coeff c{};
    c.kind = h.kind;
    for (int i = 0; i < 4; ++i) c.c[i] = h.c[i];
    c.stride = h.stride;
    c.data = h.table.empty() ? nullptr : h.table.data();  // HOST tables over the whole mesh (lfgpu_multi_assemble_reaction_diffusion)
    return c;
  };
  const lfgpu_coeff ca = as_c(ha), cg = as_c(hg);
  matrix.check(lfgpu_multi_assemble_reaction_diffusion(matrix.get(), emp.Degree(), pqt, pqq, &ca, &cg, matrix.empty_ ? 0 : 1),
               "lfgpu_multi_assemble_reaction_diffusion");
  matrix.empty_ = false;
}

// ---- edge (codim-1) providers ------------------------------------------------------------------------------------------------
// lf::uscalfe::MassEdgeMatrixProvider<SCALAR, COEFF, EDGESELECTOR> (loc_comp_ellbvp.h:367-447) and
// lf::uscalfe::ScalarLoadEdgeVectorProvider<SCALAR, FUNCTOR, EDGESELECTOR> (:784-850): same constructor shape
// (fe_space, coefficient, edge_selector); the selector is called on the host for every edge, exactly like isActive().
struct PredicateTrue {
  template <class... T>
  bool operator()(T&&...) const { return true; }
};
template <class SCALAR, class COEFF, class EDGESELECTOR = PredicateTrue>
class MassEdgeMatrixProvider {
 public:
  static_assert(std::is_same_v<SCALAR, double>, "the GPU path computes in double");
  template <class FE_SPACE>
  MassEdgeMatrixProvider(std::shared_ptr<const FE_SPACE> fe_space, COEFF gamma, EDGESELECTOR edge_selector = EDGESELECTOR{})
      : gamma_(std::move(gamma)), edge_sel_(std::move(edge_selector)), degree_(fe_space->Degree()) {}
  template <class EDGE>
  bool isActive(const EDGE& edge) { return edge_sel_(edge); }
  [[nodiscard]] const COEFF& Coeff() const { return gamma_; }
  [[nodiscard]] int Degree() const { return degree_; }

 private:
  COEFF gamma_;
  EDGESELECTOR edge_sel_;
  int degree_;
};
template <class SCALAR, class FUNCTOR, class EDGESELECTOR = PredicateTrue>
class ScalarLoadEdgeVectorProvider {
 public:
  static_assert(std::is_same_v<SCALAR, double>, "the GPU path computes in double");
  template <class FE_SPACE>
  ScalarLoadEdgeVectorProvider(std::shared_ptr<const FE_SPACE> fe_space, FUNCTOR g, EDGESELECTOR edge_sel = EDGESELECTOR{})
      : g_(std::move(g)), edge_sel_(std::move(edge_sel)), degree_(fe_space->Degree()) {}
  template <class EDGE>
  bool isActive(const EDGE& edge) { return edge_sel_(edge); }
  [[nodiscard]] const FUNCTOR& Coeff() const { return g_; }
  [[nodiscard]] int Degree() const { return degree_; }

 private:
  FUNCTOR g_;
  EDGESELECTOR edge_sel_;
  int degree_;
};

namespace detail {
// the active edges of a provider as flat arrays + the coefficient tabulated at their quadrature points
struct SegmentList {
  std::vector<double> xy;          // [n][4]
  std::vector<std::int32_t> dofs;  // [n][degree + 1]
  std::int64_t n = 0;
  void *d_xy = nullptr, *d_dofs = nullptr;
  lfgpu_ctx* ctx = nullptr;
  ~SegmentList() {
    if (d_xy) lfgpu_free(ctx, d_xy);
    if (d_dofs) lfgpu_free(ctx, d_dofs);
  }
};
template <class A, class DOFH, class PROVIDER>
void collect_segments(Context& ctx, const DOFH& dofh, PROVIDER& prov, SegmentList& s, DeviceCoeff& dc) {
  const auto& mesh = A::mesh(dofh);
  const int nsf = prov.Degree() + 1;
  for (const auto* edge : A::entities(mesh, 1)) {
    if (!prov.isActive(*edge)) continue;
    if (A::num_local_dofs(dofh, *edge) != nsf) throw Error(LFGPU_ERR_INVALID, "edge dofs do not match the Lagrange degree of the provider");
    for (int k = 0; k < 2; ++k)
      for (int d = 0; d < 2; ++d) s.xy.push_back(A::corner(*edge, k, d));
    for (const auto g : A::global_dof_indices(dofh, *edge)) s.dofs.push_back(static_cast<std::int32_t>(g));
    ++s.n;
  }
  s.ctx = ctx.get();
  if (s.n == 0) return;
  ctx.check(lfgpu_malloc(ctx.get(), 8 * static_cast<std::int64_t>(s.xy.size()), &s.d_xy), "lfgpu_malloc");
  ctx.check(lfgpu_malloc(ctx.get(), 4 * static_cast<std::int64_t>(s.dofs.size()), &s.d_dofs), "lfgpu_malloc");
  ctx.check(lfgpu_memcpy_h2d(ctx.get(), s.d_xy, s.xy.data(), 8 * static_cast<std::int64_t>(s.xy.size())), "lfgpu_memcpy_h2d");
  ctx.check(lfgpu_memcpy_h2d(ctx.get(), s.d_dofs, s.dofs.data(), 4 * static_cast<std::int64_t>(s.dofs.size())), "lfgpu_memcpy_h2d");
  // coefficient at the quadrature points of the default rule make_QuadRule(kSegment, 2p): SegmentO1::Global
  using MF = std::decay_t<decltype(prov.Coeff())>;
  std::vector<double> qxy;
  int stride = 0;
  if (CoeffTraits<MF>::needs_points) {
    double pts[16], w[16];
    const int nq = lfgpu_default_quad_rule(2, 2 * prov.Degree(), 16, pts, w);
    if (nq < 0) throw Error(nq, "no default segment rule");
    stride = nq;
    qxy.resize(static_cast<std::size_t>(s.n) * nq * 2);
    for (std::int64_t e = 0; e < s.n; ++e)
      for (int k = 0; k < nq; ++k)
        for (int d = 0; d < 2; ++d) qxy[(e * nq + k) * 2 + d] = s.xy[4 * e + 2 + d] * pts[k] + s.xy[4 * e + d] * (1 - pts[k]);
  }
  to_device(ctx, CoeffTraits<MF>::describe(prov.Coeff(), qxy.data(), s.n, stride), dc);
}
}  // namespace detail

// AssembleMatrixLocally(1, dofh, dofh, MassEdgeMatrixProvider, A): adds to the matrix that holds the cell terms
template <class A, class DOFH, class COEFF, class SEL>
void AssembleMatrixLocally(unsigned codim, const DOFH& dof_handler_trial, const DOFH& dof_handler_test,
                           MassEdgeMatrixProvider<double, COEFF, SEL>& emp, CsrMatrix& matrix) {
  if (codim != 1) throw Error(LFGPU_ERR_UNSUPPORTED, "MassEdgeMatrixProvider assembles edge (codim 1) contributions");
  if (&dof_handler_trial != &dof_handler_test) throw Error(LFGPU_ERR_UNSUPPORTED, "edge mass terms need trial space == test space");
  Context& ctx = matrix.ctx();
  matrix.template Prepare<A>(dof_handler_trial, dof_handler_test);
  detail::SegmentList s;
  detail::DeviceCoeff dc;
  detail::collect_segments<A>(ctx, dof_handler_test, emp, s, dc);
  if (s.n == 0) return;
  ctx.check(lfgpu_assemble_segment_mass(ctx.get(), matrix.pattern(), emp.Degree(), nullptr, s.n, static_cast<const double*>(s.d_xy),
                                        static_cast<const std::int32_t*>(s.d_dofs), &dc.c, matrix.device_values()),
            "lfgpu_assemble_segment_mass");
  matrix.empty_ = false;
}
// AssembleVectorLocally(1, dofh, ScalarLoadEdgeVectorProvider, vec)
template <class A, class DOFH, class F, class SEL>
void AssembleVectorLocally(unsigned codim, const DOFH& dof_handler, ScalarLoadEdgeVectorProvider<double, F, SEL>& evp, Vector& v) {
  if (codim != 1) throw Error(LFGPU_ERR_UNSUPPORTED, "ScalarLoadEdgeVectorProvider assembles edge (codim 1) contributions");
  Context& ctx = v.ctx();
  v.template Prepare<A>(dof_handler);
  detail::SegmentList s;
  detail::DeviceCoeff dc;
  detail::collect_segments<A>(ctx, dof_handler, evp, s, dc);
  if (s.n == 0) return;
  ctx.check(lfgpu_assemble_segment_load(ctx.get(), evp.Degree(), nullptr, s.n, static_cast<const double*>(s.d_xy),
                                        static_cast<const std::int32_t*>(s.d_dofs), &dc.c, v.size(), v.device()),
            "lfgpu_assemble_segment_load");
}

// GPU overloads of lf::assemble::FixFlaggedSolutionComponents / FixFlaggedSolutionCompAlt (fix_dof.h:86-138,181-218):
// same SELECTOR contract (gdof index -> std::pair<bool, double>), evaluated once per dof on the host; the matrix keeps
// the pattern of the symbolic pass (erased entries become explicit zeros), the vector is edited in place.
namespace detail {
template <class SELECTOR>
void fix_components(SELECTOR&& selectvals, CsrMatrix& A, Vector& b, bool rows_only) {
  Context& ctx = A.ctx();
  const std::int64_t n = A.cols();
  if (A.rows() != n) throw Error(LFGPU_ERR_INVALID, "Matrix must be square!");
  if (b.size() != n) throw Error(LFGPU_ERR_INVALID, "Mismatch of matrix and right-hand-side size");
  std::vector<std::uint8_t> flags(n);
  std::vector<double> vals(n, 0.0);
  for (std::int64_t k = 0; k < n; ++k) {
    const auto selval{selectvals(k)};
    flags[k] = selval.first ? 1 : 0;
    if (selval.first) vals[k] = selval.second;
  }
  void *d_flags = nullptr, *d_vals = nullptr;
  ctx.check(lfgpu_malloc(ctx.get(), n, &d_flags), "lfgpu_malloc");
  ctx.check(lfgpu_malloc(ctx.get(), 8 * n, &d_vals), "lfgpu_malloc");
  int rc = lfgpu_memcpy_h2d(ctx.get(), d_flags, flags.data(), n);
  if (rc == LFGPU_OK) rc = lfgpu_memcpy_h2d(ctx.get(), d_vals, vals.data(), 8 * n);
  if (rc == LFGPU_OK) {
    auto* fn = rows_only ? lfgpu_fix_flagged_solution_comp_alt : lfgpu_fix_flagged_solution_components;
    rc = fn(ctx.get(), A.pattern(), A.device_values(), b.device(), static_cast<const std::uint8_t*>(d_flags),
            static_cast<const double*>(d_vals), nullptr, nullptr, nullptr, nullptr);
  }
  if (rc == LFGPU_OK) rc = lfgpu_ctx_synchronize(ctx.get());
  lfgpu_free(ctx.get(), d_flags);
  lfgpu_free(ctx.get(), d_vals);
  ctx.check(rc, "lfgpu_fix_flagged_solution_components");
}
}  // namespace detail
template <class SCALAR = double, class SELECTOR>
void FixFlaggedSolutionComponents(SELECTOR&& selectvals, CsrMatrix& A, Vector& b) {
  detail::fix_components(selectvals, A, b, false);
}
template <class SCALAR = double, class SELECTOR>
void FixFlaggedSolutionCompAlt(SELECTOR&& selectvals, CsrMatrix& A, Vector& b) {
  detail::fix_components(selectvals, A, b, true);
}

// lf::assemble::FixSolutionComponentsLse (fix_dof.h:250-280): the prescribed components come as (index, value) pairs --
// values of a repeated index ADD UP (:268) -- and are imposed through the row-only variant (:272-279).
template <class SCALAR = double>
using fixed_components_t = std::vector<std::pair<std::int64_t, SCALAR>>;  // fix_dof.h:222-223
template <class SCALAR = double>
void FixSolutionComponentsLse(const fixed_components_t<SCALAR>& fixed_components, CsrMatrix& A, Vector& b) {
  const std::int64_t N = A.cols();
  if (A.rows() != N) throw Error(LFGPU_ERR_INVALID, "Matrix must be square!");
  if (b.size() != N) throw Error(LFGPU_ERR_INVALID, "Mismatch of matrix and right-hand-side size");
  std::vector<double> fixed_vec(static_cast<std::size_t>(N), 0.0);
  std::vector<char> fixed_comp_flags(static_cast<std::size_t>(N), 0);
  for (const auto& idx_val_pair : fixed_components) {
    if (idx_val_pair.first < 0 || idx_val_pair.first >= N) throw Error(LFGPU_ERR_INVALID, "Index " + std::to_string(idx_val_pair.first) + " >= N");
    fixed_vec[static_cast<std::size_t>(idx_val_pair.first)] += idx_val_pair.second;
    fixed_comp_flags[static_cast<std::size_t>(idx_val_pair.first)] = 1;
  }
  FixFlaggedSolutionCompAlt<SCALAR>(
      [&](std::int64_t i) { return std::make_pair(fixed_comp_flags[static_cast<std::size_t>(i)] != 0, fixed_vec[static_cast<std::size_t>(i)]); }, A, b);
}

// Conjugate gradients on the device for the (symmetric positive definite) system held by A and b -- the step the
// reference does with an Eigen solver on makeSparse() (homDir_linfe_demo.cc:166-175).  Returns the solution on the host.
inline std::vector<double> SolveCG(CsrMatrix& A, Vector& b, double rel_tol = 1e-10, int max_iter = 10000, int* iterations = nullptr,
                                   double* rel_residual = nullptr) {
  Context& ctx = A.ctx();
  const std::int64_t n = A.rows();
  if (b.size() != n) throw Error(LFGPU_ERR_INVALID, "Mismatch of matrix and right-hand-side size");
  void* d_x = nullptr;
  ctx.check(lfgpu_malloc(ctx.get(), 8 * n, &d_x), "lfgpu_malloc");
  int rc = lfgpu_memset(ctx.get(), d_x, 0, 8 * n);
  if (rc == LFGPU_OK)
    rc = lfgpu_cg_solve(ctx.get(), A.pattern(), A.device_values(), b.device(), static_cast<double*>(d_x), rel_tol, max_iter, 1, iterations,
                        rel_residual);
  std::vector<double> x(n);
  if (rc == LFGPU_OK) rc = lfgpu_memcpy_d2h(ctx.get(), x.data(), d_x, 8 * n);
  if (rc == LFGPU_OK) rc = lfgpu_ctx_synchronize(ctx.get());
  lfgpu_free(ctx.get(), d_x);
  ctx.check(rc, "lfgpu_cg_solve");
  return x;
}

}  // namespace lfgpu
#endif
