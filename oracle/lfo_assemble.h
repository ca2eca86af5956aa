// ORACLE (test infrastructure, NOT product code) -- see lfo_base.h header.
// lib/lf/assemble: dofhandler.h:112-228,260-503, dofhandler.cc:86-338, coomatrix.h:52-225, assembler.h:114-186,298-327
#ifndef LFO_ASSEMBLE_H
#define LFO_ASSEMBLE_H

#include <algorithm>
#include <map>

#include "lfo_mesh.h"

namespace lfo::assemble {

// lib/lf/assemble/dofhandler.h:112-228
class DofHandler {
 public:
  virtual ~DofHandler() = default;
  [[nodiscard]] virtual size_type NumDofs() const = 0;
  [[nodiscard]] virtual size_type NumLocalDofs(const mesh::Entity& entity) const = 0;
  [[nodiscard]] virtual size_type NumInteriorDofs(const mesh::Entity& entity) const = 0;
  [[nodiscard]] virtual std::span<const gdof_idx_t> GlobalDofIndices(const mesh::Entity& entity) const = 0;
  [[nodiscard]] virtual std::span<const gdof_idx_t> InteriorGlobalDofIndices(const mesh::Entity& entity) const = 0;
  [[nodiscard]] virtual const mesh::Entity& Entity(gdof_idx_t dofnum) const = 0;
  [[nodiscard]] virtual std::shared_ptr<const mesh::Mesh> Mesh() const = 0;
};

// lib/lf/assemble/dofhandler.h:260-503, dofhandler.cc:86-338
class UniformFEDofHandler final : public DofHandler {
 public:
  using dof_map_t = std::map<RefEl, size_type>;
  UniformFEDofHandler(std::shared_ptr<const mesh::Mesh> mesh, const dof_map_t& dofmap, bool check_edge_orientation = true)
      : mesh_(std::move(mesh)), check_edge_orientation_(check_edge_orientation) {
    auto get = [&](RefEl r) {
      auto it = dofmap.find(r);
      return it == dofmap.end() ? 0U : it->second;
    };
    num_loc_dof_point_ = get(RefEl::kPoint());
    num_loc_dof_segment_ = get(RefEl::kSegment());
    num_loc_dof_tria_ = get(RefEl::kTria());
    num_loc_dof_quad_ = get(RefEl::kQuad());
    // dofhandler.cc:130-138
    num_dofs_[kNodeOrd] = num_loc_dof_point_;
    num_dofs_[kEdgeOrd] = 2 * num_loc_dof_point_ + num_loc_dof_segment_;
    num_dofs_tria_ = 3 * num_loc_dof_point_ + 3 * num_loc_dof_segment_ + num_loc_dof_tria_;
    num_dofs_quad_ = 4 * num_loc_dof_point_ + 4 * num_loc_dof_segment_ + num_loc_dof_quad_;
    num_dofs_[kCellOrd] = std::max(num_dofs_tria_, num_dofs_quad_);
    initIndexArrays();
  }

  [[nodiscard]] size_type NumDofs() const override { return num_dof_; }
  [[nodiscard]] size_type NumLocalDofs(const mesh::Entity& e) const override { return NumCoveredDofs(e.RefElem()); }
  [[nodiscard]] size_type NumInteriorDofs(const mesh::Entity& e) const override { return NumInterior(e.RefElem()); }
  [[nodiscard]] std::span<const gdof_idx_t> GlobalDofIndices(const mesh::Entity& e) const override {
    return GlobalDofIndices(e.RefElem(), mesh_->Index(e));
  }
  // dofhandler.cc:286-300
  [[nodiscard]] std::span<const gdof_idx_t> GlobalDofIndices(RefEl ref_el, glb_idx_t entity_index) const {
    const dim_t codim = 2 - ref_el.Dimension();
    const size_type no_covered = NumCoveredDofs(ref_el);
    const gdof_idx_t* begin = dofs_[codim].data() + (static_cast<std::size_t>(num_dofs_[codim]) * entity_index);
    return {begin, begin + no_covered};
  }
  [[nodiscard]] std::span<const gdof_idx_t> InteriorGlobalDofIndices(const mesh::Entity& e) const override {
    const RefEl ref_el = e.RefElem();
    const dim_t codim = 2 - ref_el.Dimension();
    const size_type no_covered = NumCoveredDofs(ref_el), no_loc = NumInterior(ref_el);
    const gdof_idx_t* begin = dofs_[codim].data() + (static_cast<std::size_t>(num_dofs_[codim]) * mesh_->Index(e));
    return {begin + (no_covered - no_loc), begin + no_covered};
  }
  [[nodiscard]] const mesh::Entity& Entity(gdof_idx_t dofnum) const override { return *dof_entities_[dofnum]; }
  [[nodiscard]] std::shared_ptr<const mesh::Mesh> Mesh() const override { return mesh_; }
  [[nodiscard]] size_type CellStride() const { return num_dofs_[kCellOrd]; }
  [[nodiscard]] const std::vector<gdof_idx_t>& CellDofArray() const { return dofs_[kCellOrd]; }

 private:
  static constexpr int kNodeOrd = 2, kEdgeOrd = 1, kCellOrd = 0;
  [[nodiscard]] size_type NumCoveredDofs(RefEl r) const {
    switch (r.Id()) {
      case 1: return num_dofs_[kNodeOrd];
      case 2: return num_dofs_[kEdgeOrd];
      case 3: return num_dofs_tria_;
      default: return num_dofs_quad_;
    }
  }
  [[nodiscard]] size_type NumInterior(RefEl r) const {
    switch (r.Id()) {
      case 1: return num_loc_dof_point_;
      case 2: return num_loc_dof_segment_;
      case 3: return num_loc_dof_tria_;
      default: return num_loc_dof_quad_;
    }
  }
  // dofhandler.cc:141-284
  void initIndexArrays() {
    gdof_idx_t dof_idx = 0;
    // Step I: nodes in index order
    const size_type no_nodes = mesh_->NumEntities(2);
    dofs_[kNodeOrd].resize(static_cast<std::size_t>(no_nodes) * num_dofs_[kNodeOrd]);
    for (glb_idx_t node_idx = 0; node_idx < no_nodes; node_idx++) {
      const mesh::Entity* node_p = mesh_->EntityByIndex(2, node_idx);
      std::size_t off = static_cast<std::size_t>(node_idx) * num_dofs_[kNodeOrd];
      for (unsigned j = 0; j < num_loc_dof_point_; j++) {
        dofs_[kNodeOrd][off++] = dof_idx;
        dof_entities_.push_back(node_p);
        dof_idx++;
      }
    }
    // Step II: edges in index order
    const size_type no_edges = mesh_->NumEntities(1);
    dofs_[kEdgeOrd].resize(static_cast<std::size_t>(no_edges) * num_dofs_[kEdgeOrd]);
    for (glb_idx_t edge_idx = 0; edge_idx < no_edges; edge_idx++) {
      const mesh::Entity* edge_p = mesh_->EntityByIndex(1, edge_idx);
      std::size_t off = static_cast<std::size_t>(edge_idx) * num_dofs_[kEdgeOrd];
      for (const mesh::Entity* endpoint : edge_p->SubEntities(1)) {
        const glb_idx_t ep_idx = mesh_->Index(*endpoint);
        std::size_t ep_off = static_cast<std::size_t>(ep_idx) * num_dofs_[kNodeOrd];
        for (unsigned j = 0; j < num_dofs_[kNodeOrd]; j++) dofs_[kEdgeOrd][off++] = dofs_[kNodeOrd][ep_off++];
      }
      for (unsigned j = 0; j < num_loc_dof_segment_; j++) {
        dofs_[kEdgeOrd][off++] = dof_idx;
        dof_entities_.push_back(edge_p);
        dof_idx++;
      }
    }
    // Step III: cells in index order
    const size_type no_cells = mesh_->NumEntities(0);
    dofs_[kCellOrd].resize(static_cast<std::size_t>(no_cells) * num_dofs_[kCellOrd]);
    const size_type no_int_dof_edge = num_loc_dof_segment_;
    const size_type num_ext_dof_edge = num_dofs_[kEdgeOrd] - no_int_dof_edge;
    for (glb_idx_t cell_idx = 0; cell_idx < no_cells; cell_idx++) {
      const mesh::Entity* cell_p = mesh_->EntityByIndex(0, cell_idx);
      std::size_t off = static_cast<std::size_t>(cell_idx) * num_dofs_[kCellOrd];
      for (const mesh::Entity* vertex : cell_p->SubEntities(2)) {
        const glb_idx_t vt_idx = mesh_->Index(*vertex);
        std::size_t vt_off = static_cast<std::size_t>(vt_idx) * num_dofs_[kNodeOrd];
        for (unsigned j = 0; j < num_dofs_[kNodeOrd]; j++) dofs_[kCellOrd][off++] = dofs_[kNodeOrd][vt_off++];
      }
      const auto edge_orientations = cell_p->RelativeOrientations();
      const auto edges = cell_p->SubEntities(1);
      const size_type no_edges_cell = cell_p->RefElem().NumSubEntities(1);
      for (size_type ed_sub_idx = 0; ed_sub_idx < no_edges_cell; ed_sub_idx++) {
        const glb_idx_t edge_idx = mesh_->Index(*edges[ed_sub_idx]);
        const std::size_t edge_int_off = static_cast<std::size_t>(edge_idx) * num_dofs_[kEdgeOrd] + num_ext_dof_edge;
        if (!check_edge_orientation_ || edge_orientations[ed_sub_idx] == mesh::Orientation::positive) {
          for (size_type j = 0; j < no_int_dof_edge; j++) dofs_[kCellOrd][off++] = dofs_[kEdgeOrd][edge_int_off + j];
        } else {
          // dofhandler.cc:252-258: reversed numbering of the edge-interior dofs
          for (int j = static_cast<int>(no_int_dof_edge) - 1; j >= 0; j--) dofs_[kCellOrd][off++] = dofs_[kEdgeOrd][edge_int_off + j];
        }
      }
      const size_type num_int = (cell_p->RefElem() == RefEl::kTria()) ? num_loc_dof_tria_ : num_loc_dof_quad_;
      for (unsigned j = 0; j < num_int; j++) {
        dofs_[kCellOrd][off++] = dof_idx;
        dof_entities_.push_back(cell_p);
        dof_idx++;
      }
    }
    num_dof_ = static_cast<size_type>(dof_idx);
  }

  std::shared_ptr<const mesh::Mesh> mesh_;
  size_type num_dof_ = 0;
  std::array<size_type, 3> num_dofs_{};
  size_type num_dofs_tria_ = 0, num_dofs_quad_ = 0;
  size_type num_loc_dof_point_ = 0, num_loc_dof_segment_ = 0, num_loc_dof_tria_ = 0, num_loc_dof_quad_ = 0;
  bool check_edge_orientation_;
  std::array<std::vector<gdof_idx_t>, 3> dofs_;
  std::vector<const mesh::Entity*> dof_entities_;
};

// lib/lf/assemble/dofhandler.h:514-789 (constructor :566-718), dofhandler.cc:340-404 (accessors).
// Variable number of interior dofs per entity (hp-FEM).  LOCALDOFINFO: const mesh::Entity& -> size_type.
// Numbering: nodes by index, then edges by index, then cells by index, each entity taking as many consecutive
// indices as it has interior dofs.  The list of an edge = dofs of endpoint 0, endpoint 1, own; the list of a cell =
// vertex dofs in local vertex order, then per local edge its interior dofs (reversed for a negative relative
// orientation, :669-688), then own.  Lists are stored back to back with an offset array per codimension.
class DynamicFEDofHandler final : public DofHandler {
 public:
  template <typename LOCALDOFINFO>
  DynamicFEDofHandler(std::shared_ptr<const mesh::Mesh> mesh, LOCALDOFINFO&& locdof) : mesh_(std::move(mesh)) {
    LFO_VERIFY(mesh_->DimMesh() == 2, "Can handle 2D meshes only");
    gdof_idx_t next = 0;
    // nodes (:577-601)
    const size_type nn = mesh_->NumEntities(2);
    n_int_[2].assign(nn, 0);
    off_[2].assign(nn + 1, 0);
    for (glb_idx_t i = 0; i < nn; ++i) {
      const mesh::Entity* node = mesh_->EntityByIndex(2, i);
      off_[2][i] = static_cast<size_type>(next);
      const size_type k = locdof(*node);
      n_int_[2][i] = k;
      for (size_type j = 0; j < k; ++j) {
        list_[2].push_back(next++);
        owner_.push_back(node);
      }
    }
    off_[2][nn] = static_cast<size_type>(next);
    // edges (:603-639)
    const size_type ne = mesh_->NumEntities(1);
    n_int_[1].assign(ne, 0);
    off_[1].assign(ne + 1, 0);
    for (glb_idx_t i = 0; i < ne; ++i) {
      const mesh::Entity* edge = mesh_->EntityByIndex(1, i);
      off_[1][i] = static_cast<size_type>(list_[1].size());
      const size_type k = locdof(*edge);
      n_int_[1][i] = k;
      for (const mesh::Entity* ep : edge->SubEntities(1)) append_node_dofs(mesh_->Index(*ep), list_[1]);
      for (size_type j = 0; j < k; ++j) {
        list_[1].push_back(next++);
        owner_.push_back(edge);
      }
    }
    off_[1][ne] = static_cast<size_type>(list_[1].size());
    // cells (:641-713)
    const size_type nc = mesh_->NumEntities(0);
    n_int_[0].assign(nc, 0);
    off_[0].assign(nc + 1, 0);
    for (glb_idx_t i = 0; i < nc; ++i) {
      const mesh::Entity* cell = mesh_->EntityByIndex(0, i);
      off_[0][i] = static_cast<size_type>(list_[0].size());
      const size_type k = locdof(*cell);
      n_int_[0][i] = k;
      for (const mesh::Entity* v : cell->SubEntities(2)) append_node_dofs(mesh_->Index(*v), list_[0]);
      const auto ori = cell->RelativeOrientations();
      const auto edges = cell->SubEntities(1);
      const size_type n_loc_edges = cell->RefElem().NumSubEntities(1);
      for (size_type l = 0; l < n_loc_edges; ++l) {
        const glb_idx_t e = mesh_->Index(*edges[l]);
        const size_type ke = n_int_[1][e];
        const size_type first = off_[1][e + 1] - ke;  // interior dofs sit at the end of the edge's list
        if (ori[l] == mesh::Orientation::positive) {
          for (size_type j = 0; j < ke; ++j) list_[0].push_back(list_[1][first + j]);
        } else {
          for (size_type j = ke; j-- > 0;) list_[0].push_back(list_[1][first + j]);
        }
      }
      for (size_type j = 0; j < k; ++j) {
        list_[0].push_back(next++);
        owner_.push_back(cell);
      }
    }
    off_[0][nc] = static_cast<size_type>(list_[0].size());
    num_dof_ = static_cast<size_type>(next);
  }

  [[nodiscard]] size_type NumDofs() const override { return num_dof_; }
  // dofhandler.cc:340-359
  [[nodiscard]] size_type NumLocalDofs(const mesh::Entity& e) const override {
    const dim_t cd = e.Codim();
    const glb_idx_t i = mesh_->Index(e);
    return off_[cd][i + 1] - off_[cd][i];
  }
  [[nodiscard]] size_type NumInteriorDofs(const mesh::Entity& e) const override { return n_int_[e.Codim()][mesh_->Index(e)]; }
  // dofhandler.cc:361-384
  [[nodiscard]] std::span<const gdof_idx_t> GlobalDofIndices(const mesh::Entity& e) const override {
    const dim_t cd = e.Codim();
    const glb_idx_t i = mesh_->Index(e);
    const gdof_idx_t* b = list_[cd].data();
    return {b + off_[cd][i], b + off_[cd][i + 1]};
  }
  // dofhandler.cc:386-404
  [[nodiscard]] std::span<const gdof_idx_t> InteriorGlobalDofIndices(const mesh::Entity& e) const override {
    const dim_t cd = e.Codim();
    const glb_idx_t i = mesh_->Index(e);
    const gdof_idx_t* b = list_[cd].data();
    return {b + (off_[cd][i + 1] - n_int_[cd][i]), b + off_[cd][i + 1]};
  }
  [[nodiscard]] const mesh::Entity& Entity(gdof_idx_t dofnum) const override {
    LFO_VERIFY(dofnum >= 0 && static_cast<std::size_t>(dofnum) < owner_.size(), "Illegal dof index");
    return *owner_[dofnum];
  }
  [[nodiscard]] std::shared_ptr<const mesh::Mesh> Mesh() const override { return mesh_; }

 private:
  void append_node_dofs(glb_idx_t node, std::vector<gdof_idx_t>& dst) const {
    for (size_type j = 0; j < n_int_[2][node]; ++j) dst.push_back(list_[2][off_[2][node] + j]);
  }
  std::shared_ptr<const mesh::Mesh> mesh_;
  size_type num_dof_ = 0;
  std::vector<const mesh::Entity*> owner_;
  std::array<std::vector<size_type>, 3> n_int_, off_;  // index = codimension
  std::array<std::vector<gdof_idx_t>, 3> list_;
};

// Eigen::Triplet<double> : int row, int col, double value (16 bytes)
struct Triplet {
  int row, col;
  double value;
};

// Compressed column-major matrix with int32 indices == Eigen::SparseMatrix<double> (ColMajor, StorageIndex=int)
struct CompressedMatrix {
  long rows = 0, cols = 0;
  std::vector<int> outer;  // cols + 1
  std::vector<int> inner;  // nnz, ascending within each column
  std::vector<double> values;
};

// lib/lf/assemble/coomatrix.h:52-225
class COOMatrix {
 public:
  COOMatrix(size_type num_rows, size_type num_cols) : rows_(num_rows), cols_(num_cols) {}
  [[nodiscard]] gdof_idx_t rows() const { return rows_; }
  [[nodiscard]] gdof_idx_t cols() const { return cols_; }
  // coomatrix.h:87-91
  void AddToEntry(gdof_idx_t i, gdof_idx_t j, double increment) {
    rows_ = (i + 1 > rows_) ? i + 1 : rows_;
    cols_ = (j + 1 > cols_) ? j + 1 : cols_;
    triplets_.push_back(Triplet{static_cast<int>(i), static_cast<int>(j), increment});
  }
  void setZero() { triplets_.clear(); }
  // coomatrix.h:108-115: remove every triplet for which pred(row, col) holds
  template <typename PREDICATE>
  void setZero(PREDICATE&& pred) {
    auto new_last = std::remove_if(triplets_.begin(), triplets_.end(), [&pred](Triplet& t) { return pred(t.row, t.col); });
    triplets_.erase(new_last, triplets_.end());
  }
  // coomatrix.h:250-262: resvec += alpha * A * vec, triplet by triplet
  template <typename VECTOR, typename RESULTVECTOR>
  void MatVecMult(double alpha, const VECTOR& vec, RESULTVECTOR& resvec) const {
    for (const Triplet& t : triplets_) resvec[t.row] += t.value * (alpha * vec[t.col]);
  }
  [[nodiscard]] const std::vector<Triplet>& triplets() const { return triplets_; }

  // coomatrix.h:172-180 -> Eigen 3.4.0 SparseMatrix::setFromTriplets (set_from_triplets in SparseMatrix.h):
  //  pass 1 count entries per row; pass 2 fill a ROW-major temporary in triplet order (insertBackUncompressed);
  //  pass 3 collapseDuplicates: within each row keep the FIRST occurrence of a column and add later duplicates
  //  to it (sum in insertion order), explicit zeros are kept; pass 4 transposed copy into the column-major result,
  //  which orders the inner (row) indices of every column ascending.
  [[nodiscard]] CompressedMatrix makeSparse() const {
    LFO_VERIFY(rows_ > 0 && cols_ > 0, "matrix has zero rows or columns, this is probably an error.");
    const long R = rows_, C = cols_;
    std::vector<int> wi(R, 0);
    for (const auto& t : triplets_) wi[t.row]++;
    std::vector<long> rstart(R + 1, 0);
    for (long r = 0; r < R; ++r) rstart[r + 1] = rstart[r] + wi[r];
    std::vector<int> tcol(triplets_.size());
    std::vector<double> tval(triplets_.size());
    std::vector<long> fill(rstart.begin(), rstart.end() - 1);
    for (const auto& t : triplets_) {
      const long p = fill[t.row]++;
      tcol[p] = t.col;
      tval[p] = t.value;
    }
    // collapseDuplicates
    std::vector<long> mark(C, -1);
    std::vector<long> rend(R);
    long count = 0;
    std::vector<long> new_rstart(R + 1, 0);
    for (long r = 0; r < R; ++r) {
      const long start = count;
      for (long k = rstart[r]; k < rstart[r + 1]; ++k) {
        const int c = tcol[k];
        if (mark[c] >= start) {
          tval[mark[c]] += tval[k];
        } else {
          tval[count] = tval[k];
          tcol[count] = c;
          mark[c] = count;
          ++count;
        }
      }
      new_rstart[r] = start;
    }
    new_rstart[R] = count;
    // transposed copy (row-major temp -> column-major result)
    CompressedMatrix m;
    m.rows = R;
    m.cols = C;
    m.outer.assign(C + 1, 0);
    for (long k = 0; k < count; ++k) m.outer[tcol[k] + 1]++;
    for (long c = 0; c < C; ++c) m.outer[c + 1] += m.outer[c];
    m.inner.resize(count);
    m.values.resize(count);
    std::vector<int> pos(m.outer.begin(), m.outer.end() - 1);
    for (long r = 0; r < R; ++r) {
      for (long k = new_rstart[r]; k < new_rstart[r + 1]; ++k) {
        const int p = pos[tcol[k]]++;
        m.inner[p] = static_cast<int>(r);
        m.values[p] = tval[k];
      }
    }
    return m;
  }

 private:
  gdof_idx_t rows_, cols_;
  std::vector<Triplet> triplets_;
};

// lib/lf/assemble/fix_dof.h:86-138: enforce prescribed solution components on the COO matrix and the right-hand side.
// SELECTOR: idx -> std::pair<bool, double>
template <typename SELECTOR, typename RHSVECTOR>
void FixFlaggedSolutionComponents(SELECTOR&& selectvals, COOMatrix& A, RHSVECTOR& b) {
  const gdof_idx_t N = A.cols();
  LFO_VERIFY(A.rows() == N, "Matrix must be square!");
  {
    std::vector<double> tmp_vec(N);
    for (gdof_idx_t k = 0; k < N; ++k) {
      const auto selval{selectvals(k)};
      tmp_vec[k] = selval.first ? selval.second : 0.0;
    }
    A.MatVecMult(-1.0, tmp_vec, b);
  }
  for (gdof_idx_t k = 0; k < N; ++k) {
    const auto selval{selectvals(k)};
    if (selval.first) b[k] = selval.second;
  }
  A.setZero([&selectvals](gdof_idx_t i, gdof_idx_t j) { return selectvals(i).first || selectvals(j).first; });
  for (gdof_idx_t dofnum = 0; dofnum < N; ++dofnum) {
    if (selectvals(dofnum).first) A.AddToEntry(dofnum, dofnum, 1.0);
  }
}

// lib/lf/assemble/fix_dof.h:181-218: the non-symmetric variant -- only the ROWS of fixed components are replaced by
// unit rows, the right-hand side just receives the prescribed values.
template <typename SELECTOR, typename RHSVECTOR>
void FixFlaggedSolutionCompAlt(SELECTOR&& selectvals, COOMatrix& A, RHSVECTOR& b) {
  const gdof_idx_t N = A.cols();
  LFO_VERIFY(A.rows() == N, "Matrix must be square!");
  for (gdof_idx_t k = 0; k < N; ++k) {
    const auto selval{selectvals(k)};
    if (selval.first) b[k] = selval.second;
  }
  A.setZero([&selectvals](gdof_idx_t i, gdof_idx_t /*j*/) { return selectvals(i).first; });
  for (gdof_idx_t dofnum = 0; dofnum < N; ++dofnum) {
    if (selectvals(dofnum).first) A.AddToEntry(dofnum, dofnum, 1.0);
  }
}

// lib/lf/assemble/fix_dof.h:250-280: prescribed components as (index, value) pairs; values of a repeated index add up
// (:268); imposed through the row-only variant (:272-279).
using fixed_components_t = std::vector<std::pair<gdof_idx_t, double>>;
template <typename RHSVECTOR>
void FixSolutionComponentsLse(const fixed_components_t& fixed_components, COOMatrix& A, RHSVECTOR& b) {
  const gdof_idx_t N = A.cols();
  LFO_VERIFY(A.rows() == N, "Matrix must be square!");
  std::vector<double> fixed_vec(N, 0.0);
  std::vector<bool> fixed_comp_flags(N, false);
  for (const auto& idx_val_pair : fixed_components) {
    LFO_VERIFY(idx_val_pair.first < N, "Index >= N");
    fixed_vec[idx_val_pair.first] += idx_val_pair.second;
    fixed_comp_flags[idx_val_pair.first] = true;
  }
  FixFlaggedSolutionCompAlt(
      [&fixed_comp_flags, &fixed_vec](gdof_idx_t i) -> std::pair<bool, double> { return std::make_pair(fixed_comp_flags[i], fixed_vec[i]); }, A,
      b);
}

// lib/lf/assemble/assembler.h:114-186
template <typename TMPMATRIX, typename ENTITY_MATRIX_PROVIDER>
void AssembleMatrixLocally(dim_t codim, const DofHandler& dof_handler_trial, const DofHandler& dof_handler_test,
                           ENTITY_MATRIX_PROVIDER& entity_matrix_provider, TMPMATRIX& matrix) {
  auto mesh = dof_handler_trial.Mesh();
  LFO_VERIFY(mesh == dof_handler_test.Mesh(), "Trial and test space must be defined on the same mesh");
  for (const mesh::Entity* entity : mesh->Entities(codim)) {
    if (entity_matrix_provider.isActive(*entity)) {
      const size_type nrows_loc = dof_handler_test.NumLocalDofs(*entity);
      const size_type ncols_loc = dof_handler_trial.NumLocalDofs(*entity);
      std::span<const gdof_idx_t> row_idx(dof_handler_test.GlobalDofIndices(*entity));
      std::span<const gdof_idx_t> col_idx(dof_handler_trial.GlobalDofIndices(*entity));
      const auto elem_mat{entity_matrix_provider.Eval(*entity)};
      for (size_type i = 0; i < nrows_loc; i++) {
        for (size_type j = 0; j < ncols_loc; j++) matrix.AddToEntry(row_idx[i], col_idx[j], elem_mat(i, j));
      }
    }
  }
}

// lib/lf/assemble/assembler.h:298-327
template <typename VECTOR, typename ENTITY_VECTOR_PROVIDER>
void AssembleVectorLocally(dim_t codim, const DofHandler& dof_handler, ENTITY_VECTOR_PROVIDER& entity_vector_provider,
                           VECTOR& resultvector) {
  auto mesh = dof_handler.Mesh();
  for (const mesh::Entity* entity : mesh->Entities(codim)) {
    if (entity_vector_provider.isActive(*entity)) {
      const size_type veclen = dof_handler.NumLocalDofs(*entity);
      const std::span<const gdof_idx_t> dof_idx(dof_handler.GlobalDofIndices(*entity));
      const auto elem_vec{entity_vector_provider.Eval(*entity)};
      for (size_type i = 0; i < veclen; i++) resultvector[dof_idx[i]] += elem_vec[i];
    }
  }
}

}  // namespace lfo::assemble
#endif
