// ORACLE (test infrastructure, NOT product code) -- see lfo_base.h header.
// Mesh topology exactly as the reference builds it:
//   lib/lf/mesh/entity.h:16-60, mesh_interface.h, hybrid2d/mesh.cc:84-104,178-810, hybrid2d/mesh_factory.cc:30-121,
//   hybrid2d/triangle.cc:16-78, hybrid2d/quad.cc:16-81, utils/tp_triag_mesh_builder.cc:18-178,
//   utils/tp_quad_mesh_builder.cc:19-95
#ifndef LFO_MESH_H
#define LFO_MESH_H

#include <array>
#include <map>
#include <utility>

#include "lfo_geometry.h"

namespace lfo::mesh {

using GeometryPtr = std::unique_ptr<geometry::Geometry>;

enum class Orientation : int { positive = 1, negative = -1 };

// lib/lf/mesh/entity.h:16-60
class Entity {
 public:
  virtual ~Entity() = default;
  [[nodiscard]] virtual unsigned Codim() const = 0;
  [[nodiscard]] virtual std::span<const Entity* const> SubEntities(unsigned rel_codim) const = 0;
  [[nodiscard]] virtual std::span<const Orientation> RelativeOrientations() const = 0;
  [[nodiscard]] virtual const geometry::Geometry* Geometry() const = 0;
  [[nodiscard]] virtual RefEl RefElem() const = 0;
};

namespace hybrid2d {

class Point final : public Entity {
 public:
  Point(size_type index, GeometryPtr&& geo) : index_(index), geometry_(std::move(geo)), this_(this) {}
  Point(Point&& o) noexcept : index_(o.index_), geometry_(std::move(o.geometry_)), this_(this) {}
  [[nodiscard]] unsigned Codim() const override { return 2; }
  [[nodiscard]] std::span<const Entity* const> SubEntities(unsigned) const override { return {&this_, 1}; }
  [[nodiscard]] std::span<const Orientation> RelativeOrientations() const override { return {}; }
  [[nodiscard]] const geometry::Geometry* Geometry() const override { return geometry_.get(); }
  [[nodiscard]] RefEl RefElem() const override { return RefEl::kPoint(); }
  [[nodiscard]] size_type index() const { return index_; }

 private:
  size_type index_;
  GeometryPtr geometry_;
  const Entity* this_;
};

class Segment final : public Entity {
 public:
  Segment(size_type index, GeometryPtr&& geo, const Point* p0, const Point* p1)
      : index_(index), geometry_(std::move(geo)), nodes_({p0, p1}), this_(this) {}
  Segment(Segment&& o) noexcept : index_(o.index_), geometry_(std::move(o.geometry_)), nodes_(o.nodes_), this_(this) {}
  [[nodiscard]] unsigned Codim() const override { return 1; }
  [[nodiscard]] std::span<const Entity* const> SubEntities(unsigned rel_codim) const override {
    if (rel_codim == 1) return {reinterpret_cast<const Entity* const*>(nodes_.data()), 2};
    return {&this_, 1};
  }
  [[nodiscard]] std::span<const Orientation> RelativeOrientations() const override { return kEndpointOri; }
  [[nodiscard]] const geometry::Geometry* Geometry() const override { return geometry_.get(); }
  [[nodiscard]] RefEl RefElem() const override { return RefEl::kSegment(); }
  [[nodiscard]] size_type index() const { return index_; }
  [[nodiscard]] const Point* node(int k) const { return static_cast<const Point*>(nodes_[k]); }

 private:
  static constexpr std::array<Orientation, 2> kEndpointOri{Orientation::negative, Orientation::positive};
  size_type index_;
  GeometryPtr geometry_;
  std::array<const Entity*, 2> nodes_;
  const Entity* this_;
};

// hybrid2d/triangle.cc:16-78 and hybrid2d/quad.cc:16-81 -- NV = 3 or 4
template <int NV>
class Cell final : public Entity {
 public:
  Cell(size_type index, GeometryPtr&& geo, const std::array<const Point*, NV>& corners,
       const std::array<const Segment*, NV>& edges)
      : index_(index), geometry_(std::move(geo)), this_(this) {
    for (int i = 0; i < NV; ++i) {
      LFO_VERIFY(corners[i] != nullptr, "Invalid pointer to corner");
      LFO_VERIFY(edges[i] != nullptr, "Invalid pointer to edge");
      nodes_[i] = corners[i];
      edges_[i] = edges[i];
    }
    const RefEl ref_el = RefElem();
    for (int e = 0; e < NV; ++e) {
      const Segment* ed = static_cast<const Segment*>(edges_[e]);
      const Entity* p0 = nodes_[ref_el.EdgeEndpoint(e, 0)];
      const Entity* p1 = nodes_[ref_el.EdgeEndpoint(e, 1)];
      LFO_VERIFY(ed->node(0) == p0 || ed->node(0) == p1, "Node 0 of edge not a cell node");
      LFO_VERIFY(ed->node(1) == p0 || ed->node(1) == p1, "Node 1 of edge not a cell node");
      // triangle.cc:70-77: edge i has positive orientation iff its first node agrees with vertex i
      edge_ori_[e] = (ed->node(0) == nodes_[e]) ? Orientation::positive : Orientation::negative;
    }
  }
  Cell(Cell&& o) noexcept
      : index_(o.index_), geometry_(std::move(o.geometry_)), nodes_(o.nodes_), edges_(o.edges_),
        edge_ori_(o.edge_ori_), this_(this) {}
  [[nodiscard]] unsigned Codim() const override { return 0; }
  [[nodiscard]] std::span<const Entity* const> SubEntities(unsigned rel_codim) const override {
    if (rel_codim == 2) return {nodes_.data(), NV};
    if (rel_codim == 1) return {edges_.data(), NV};
    return {&this_, 1};
  }
  [[nodiscard]] std::span<const Orientation> RelativeOrientations() const override { return edge_ori_; }
  [[nodiscard]] const geometry::Geometry* Geometry() const override { return geometry_.get(); }
  [[nodiscard]] RefEl RefElem() const override { return NV == 3 ? RefEl::kTria() : RefEl::kQuad(); }
  [[nodiscard]] size_type index() const { return index_; }

 private:
  size_type index_;
  GeometryPtr geometry_;
  std::array<const Entity*, NV> nodes_{};
  std::array<const Entity*, NV> edges_{};
  std::array<Orientation, NV> edge_ori_{};
  const Entity* this_;
};
using Triangle = Cell<3>;
using Quadrilateral = Cell<4>;

// hybrid2d/mesh.cc:128-165
class EndpointIndexPair {
 public:
  EndpointIndexPair(size_type p0, size_type p1) : p0_(p0), p1_(p1) {
    LFO_VERIFY(p0 != p1, "No loops allowed");
    if (p1 > p0) {
      cmp_p0_ = p0;
      cmp_p1_ = p1;
    } else {
      cmp_p0_ = p1;
      cmp_p1_ = p0;
    }
  }
  [[nodiscard]] size_type first_node() const { return p0_; }
  [[nodiscard]] size_type second_node() const { return p1_; }
  friend bool operator<(const EndpointIndexPair& a, const EndpointIndexPair& b) {
    return (a.cmp_p0_ == b.cmp_p0_) ? (a.cmp_p1_ < b.cmp_p1_) : (a.cmp_p0_ < b.cmp_p0_);
  }
  friend bool coincide(const EndpointIndexPair& a, const EndpointIndexPair& b) {
    return a.p0_ == b.p0_ && a.p1_ == b.p1_;
  }

 private:
  size_type p0_, p1_, cmp_p0_, cmp_p1_;
};

class Mesh {
 public:
  using NodeCoordList = std::vector<GeometryPtr>;
  using EdgeList = std::vector<std::pair<std::array<size_type, 2>, GeometryPtr>>;
  using CellList = std::vector<std::pair<std::array<size_type, 4>, GeometryPtr>>;

  // hybrid2d/mesh.cc:178-810
  Mesh(NodeCoordList nodes, EdgeList edges, CellList cells, bool check_completeness = true) {
    struct AdjCellInfo {
      size_type cell_idx, edge_idx;
    };
    struct EdgeData {
      GeometryPtr geo_uptr;
      std::vector<AdjCellInfo> adj_cells_list;
      glb_idx_t edge_global_index = kIdxNil;
      bool reversed = false;
    };
    using EdgeMap = std::map<EndpointIndexPair, EdgeData>;
    const Mat zero_point(0, 1);
    const size_type no_of_nodes = static_cast<size_type>(nodes.size());

    // STEP I (mesh.cc:231-274): register the supplied edges; index = position in `edges`
    EdgeMap edge_map;
    glb_idx_t edge_index = 0;
    for (auto& e : edges) {
      const std::array<size_type, 2> end_nodes(e.first);
      const EndpointIndexPair key(end_nodes[0], end_nodes[1]);
      LFO_VERIFY(end_nodes[0] < no_of_nodes && end_nodes[1] < no_of_nodes, "Illegal edge node numbers");
      LFO_VERIFY(e.second != nullptr, "Edge: missing geometry!");
      for (int j = 0; j < 2; ++j) {
        if (nodes[end_nodes[j]] == nullptr) nodes[end_nodes[j]] = e.second->SubGeometry(1, j);
      }
      EdgeData ed;
      ed.geo_uptr = std::move(e.second);
      ed.edge_global_index = edge_index;
      const auto st = edge_map.insert(std::make_pair(key, std::move(ed)));
      LFO_VERIFY(st.second, "Duplicate edge");
      edge_index++;
    }

    // STEP II (mesh.cc:299-446): scan the cells, create missing edges, record adjacency
    size_type cell_index = 0, no_of_trilaterals = 0, no_of_quadrilaterals = 0;
    for (const auto& c : cells) {
      const std::array<size_type, 4>& cell_node_list(c.first);
      const GeometryPtr& cell_geometry(c.second);
      size_type no_of_vertices;
      if (cell_node_list[3] == kIdxNil) {
        no_of_vertices = 3;
        no_of_trilaterals++;
      } else {
        no_of_vertices = 4;
        no_of_quadrilaterals++;
      }
      const RefEl ref_el = (no_of_vertices == 3) ? RefEl::kTria() : RefEl::kQuad();
      for (unsigned l = 0; l < no_of_vertices; l++) LFO_VERIFY(cell_node_list[l] < no_of_nodes, "invalid node index");
      if (cell_geometry != nullptr) {
        for (unsigned j = 0; j < no_of_vertices; ++j) {
          if (nodes[cell_node_list[j]] == nullptr) nodes[cell_node_list[j]] = cell_geometry->SubGeometry(2, j);
        }
      }
      for (unsigned j = 0; j < ref_el.NumSubEntities(1); j++) {
        const size_type p0l = ref_el.EdgeEndpoint(j, 0), p1l = ref_el.EdgeEndpoint(j, 1);
        const EndpointIndexPair key(cell_node_list[p0l], cell_node_list[p1l]);
        const AdjCellInfo info{cell_index, j};
        auto it = edge_map.find(key);
        if (it == edge_map.end()) {
          EdgeData ed;
          if (cell_geometry) ed.geo_uptr = cell_geometry->SubGeometry(1, j);
          ed.adj_cells_list.push_back(info);
          const auto st = edge_map.insert(std::make_pair(key, std::move(ed)));
          LFO_VERIFY(st.second, "Duplicate not found earlier!");
        } else {
          it->second.adj_cells_list.push_back(info);
          if (it->second.geo_uptr == nullptr && cell_geometry) {
            it->second.geo_uptr = cell_geometry->SubGeometry(1, j);
            if (!coincide(key, it->first)) it->second.reversed = true;  // mesh.cc:413-428
          }
        }
      }
      cell_index++;
    }

    // nodes (mesh.cc:484-500)
    points_.reserve(no_of_nodes);
    size_type node_index = 0;
    for (GeometryPtr& g : nodes) {
      LFO_VERIFY(g != nullptr, "Missing geometry for node");
      points_.emplace_back(node_index, std::move(g));
      node_index++;
    }

    // edges in map (sorted-key) order (mesh.cc:502-580)
    const size_type no_of_edges = static_cast<size_type>(edge_map.size());
    segments_.reserve(no_of_edges);
    std::vector<bool> node_has_super;
    if (check_completeness) node_has_super.resize(no_of_nodes, false);
    for (auto& edge : edge_map) {
      size_type p0 = edge.first.first_node(), p1 = edge.first.second_node();
      if (edge.second.reversed) std::swap(p0, p1);
      const Point* p0_ptr = &points_[p0];
      const Point* p1_ptr = &points_[p1];
      GeometryPtr geo(std::move(edge.second.geo_uptr));
      if (!geo) {
        Mat cc(2, 2);
        const Mat a = p0_ptr->Geometry()->Global(zero_point), b = p1_ptr->Geometry()->Global(zero_point);
        cc(0, 0) = a(0, 0); cc(1, 0) = a(1, 0); cc(0, 1) = b(0, 0); cc(1, 1) = b(1, 0);
        geo = std::make_unique<geometry::SegmentO1>(cc);
      }
      if (edge.second.edge_global_index == kIdxNil) {
        edge.second.edge_global_index = edge_index;  // mesh.cc:550-556: new edges numbered in map order
        edge_index++;
      }
      if (check_completeness) {
        node_has_super[p0] = true;
        node_has_super[p1] = true;
        LFO_VERIFY(!edge.second.adj_cells_list.empty(), "Mesh is incomplete: edge does not belong to a cell");
      }
      segments_.emplace_back(edge.second.edge_global_index, std::move(geo), p0_ptr, p1_ptr);
    }
    LFO_VERIFY(edge_index == no_of_edges, "Edge index mismatch");
    if (check_completeness) {
      for (size_type i = 0; i < no_of_nodes; ++i) LFO_VERIFY(node_has_super[i], "Mesh is incomplete: isolated node");
    }

    // cells (mesh.cc:582-760)
    const size_type no_of_cells = static_cast<size_type>(cells.size());
    std::vector<std::array<size_type, 4>> edge_indices(no_of_cells);
    size_type edge_array_position = 0;
    for (const auto& edge : edge_map) {
      for (const auto& adj : edge.second.adj_cells_list) edge_indices[adj.cell_idx][adj.edge_idx] = edge_array_position;
      edge_array_position++;
    }
    trias_.reserve(no_of_trilaterals);
    quads_.reserve(no_of_quadrilaterals);
    cell_index = 0;
    for (auto& c : cells) {
      const std::array<size_type, 4>& cn(c.first);
      const std::array<size_type, 4>& ce(edge_indices[cell_index]);
      const size_type nv = (cn[3] == kIdxNil) ? 3 : 4;
      for (unsigned l = 0; l < nv; l++) LFO_VERIFY(ce[l] < no_of_edges, "invalid edge index");
      GeometryPtr geo(std::move(c.second));
      if (nv == 3) {
        if (!geo) {
          Mat cc(2, 3);
          for (int k = 0; k < 3; ++k) {
            const Mat p = points_[cn[k]].Geometry()->Global(zero_point);
            cc(0, k) = p(0, 0);
            cc(1, k) = p(1, 0);
          }
          geo = std::make_unique<geometry::TriaO1>(cc);
        }
        trias_.emplace_back(cell_index, std::move(geo),
                            std::array<const Point*, 3>{&points_[cn[0]], &points_[cn[1]], &points_[cn[2]]},
                            std::array<const Segment*, 3>{&segments_[ce[0]], &segments_[ce[1]], &segments_[ce[2]]});
      } else {
        if (!geo) {
          Mat cc(2, 4);
          for (int k = 0; k < 4; ++k) {
            const Mat p = points_[cn[k]].Geometry()->Global(zero_point);
            cc(0, k) = p(0, 0);
            cc(1, k) = p(1, 0);
          }
          geo = std::make_unique<geometry::QuadO1>(cc);
        }
        quads_.emplace_back(
            cell_index, std::move(geo),
            std::array<const Point*, 4>{&points_[cn[0]], &points_[cn[1]], &points_[cn[2]], &points_[cn[3]]},
            std::array<const Segment*, 4>{&segments_[ce[0]], &segments_[ce[1]], &segments_[ce[2]], &segments_[ce[3]]});
      }
      cell_index++;
    }

    // entity pointer arrays ordered by index (mesh.cc:762-806)
    entity_pointers_[0].assign(trias_.size() + quads_.size(), nullptr);
    for (auto& t : trias_) {
      LFO_VERIFY(entity_pointers_[0][t.index()] == nullptr, "Cell index occurs twice!");
      entity_pointers_[0][t.index()] = &t;
    }
    for (auto& q : quads_) {
      LFO_VERIFY(entity_pointers_[0][q.index()] == nullptr, "Cell index occurs twice!");
      entity_pointers_[0][q.index()] = &q;
    }
    entity_pointers_[1].assign(segments_.size(), nullptr);
    entity_pointers_[2].assign(points_.size(), nullptr);
    for (auto& p : points_) entity_pointers_[2][p.index()] = &p;
    for (auto& s : segments_) entity_pointers_[1][s.index()] = &s;
  }

  [[nodiscard]] unsigned DimMesh() const { return 2; }
  [[nodiscard]] std::span<const Entity* const> Entities(unsigned codim) const {
    return {entity_pointers_[codim].data(), entity_pointers_[codim].size()};
  }
  [[nodiscard]] size_type NumEntities(unsigned codim) const {
    return static_cast<size_type>(entity_pointers_[codim].size());
  }
  [[nodiscard]] size_type NumEntities(RefEl r) const {
    if (r == RefEl::kTria()) return static_cast<size_type>(trias_.size());
    if (r == RefEl::kQuad()) return static_cast<size_type>(quads_.size());
    return NumEntities(2 - r.Dimension());
  }
  // hybrid2d/mesh.cc:84-104: dispatch on codim, then dynamic_cast on the concrete cell type
  [[nodiscard]] size_type Index(const Entity& e) const {
    switch (e.Codim()) {
      case 0: {
        if (e.RefElem() == RefEl::kTria()) return dynamic_cast<const Triangle&>(e).index();
        if (e.RefElem() == RefEl::kQuad()) return dynamic_cast<const Quadrilateral&>(e).index();
        LFO_VERIFY(false, "Illegal cell type");
        return kIdxNil;
      }
      case 1:
        return dynamic_cast<const Segment&>(e).index();
      case 2:
        return dynamic_cast<const Point&>(e).index();
      default:
        LFO_VERIFY(false, "Illegal codim");
        return kIdxNil;
    }
  }
  [[nodiscard]] const Entity* EntityByIndex(unsigned codim, glb_idx_t index) const {
    return entity_pointers_[codim][index];
  }

 private:
  std::vector<Point> points_;
  std::vector<Segment> segments_;
  std::vector<Triangle> trias_;
  std::vector<Quadrilateral> quads_;
  std::array<std::vector<const Entity*>, 3> entity_pointers_;
};

// hybrid2d/mesh_factory.cc:30-121
class MeshFactory {
 public:
  size_type AddPoint(double x, double y) {
    nodes_.emplace_back(std::make_unique<geometry::Point>(x, y));
    return static_cast<size_type>(nodes_.size() - 1);
  }
  size_type AddEntity(RefEl ref_el, std::span<const size_type> nodes, GeometryPtr&& geometry) {
    if (ref_el == RefEl::kSegment()) {
      LFO_VERIFY(nodes.size() == 2, "ref_el = segment needs 2 nodes");
      std::array<size_type, 2> ns{nodes[0], nodes[1]};
      for (auto n : ns) LFO_VERIFY(n < nodes_.size(), "node must be inserted with AddPoint() first");
      edges_.emplace_back(ns, std::move(geometry));
      return static_cast<size_type>(edges_.size() - 1);
    }
    LFO_VERIFY(nodes.size() == ref_el.NumNodes(), "wrong number of nodes for cell");
    std::array<size_type, 4> ns{};
    unsigned char count = 0;
    for (auto n : nodes) {
      LFO_VERIFY(n < nodes_.size(), "node must be inserted with AddPoint() first");
      ns[count++] = n;
    }
    if (count == 3) ns[3] = kIdxNil;  // mesh_factory.cc:97-100
    elements_.emplace_back(ns, std::move(geometry));
    return static_cast<size_type>(elements_.size() - 1);
  }
  std::shared_ptr<Mesh> Build() {
    auto m = std::make_shared<Mesh>(std::move(nodes_), std::move(edges_), std::move(elements_), true);
    nodes_.clear();
    edges_.clear();
    elements_.clear();
    return m;
  }

 private:
  Mesh::NodeCoordList nodes_;
  Mesh::EdgeList edges_;
  Mesh::CellList elements_;
};

}  // namespace hybrid2d

using Mesh = hybrid2d::Mesh;

namespace utils {

// utils/tp_triag_mesh_builder.cc:18-178 (all edges supplied explicitly, two triangles per square)
inline std::shared_ptr<Mesh> TPTriagMeshBuild(size_type nx, size_type ny, double blx, double bly, double trx, double try_) {
  hybrid2d::MeshFactory factory;
  const unsigned no_of_cells = 2 * nx * ny;
  if (no_of_cells == 0) return nullptr;
  const double x_size = trx - blx, y_size = try_ - bly;
  if (x_size <= 0.0 || y_size <= 0.0) return nullptr;
  const double hx = x_size / nx, hy = y_size / ny;
  auto VertexIndex = [nx](size_type i, size_type j) { return i + j * (nx + 1); };
  std::vector<size_type> v_idx((nx + 1) * (ny + 1));
  int node_cnt = 0;
  for (size_type j = 0; j <= ny; ++j) {
    for (size_type i = 0; i <= nx; ++i, ++node_cnt) v_idx[node_cnt] = factory.AddPoint(blx + i * hx, bly + j * hy);
  }
  auto seg = [&](double ax, double ay, double bx, double by) {
    Mat g(2, 2);
    g(0, 0) = ax; g(0, 1) = bx; g(1, 0) = ay; g(1, 1) = by;
    return std::make_unique<geometry::SegmentO1>(g);
  };
  for (size_type i = 0; i < nx; ++i) {  // horizontal
    for (size_type j = 0; j <= ny; ++j) {
      const std::array<size_type, 2> nl{v_idx[VertexIndex(i, j)], v_idx[VertexIndex(i + 1, j)]};
      factory.AddEntity(RefEl::kSegment(), nl, seg(blx + i * hx, bly + j * hy, blx + (i + 1) * hx, bly + j * hy));
    }
  }
  for (size_type i = 0; i <= nx; ++i) {  // vertical
    for (size_type j = 0; j < ny; ++j) {
      const std::array<size_type, 2> nl{v_idx[VertexIndex(i, j)], v_idx[VertexIndex(i, j + 1)]};
      factory.AddEntity(RefEl::kSegment(), nl, seg(blx + i * hx, bly + j * hy, blx + i * hx, bly + (j + 1) * hy));
    }
  }
  for (size_type i = 0; i < nx; ++i) {  // diagonal
    for (size_type j = 0; j < ny; ++j) {
      const std::array<size_type, 2> nl{v_idx[VertexIndex(i, j)], v_idx[VertexIndex(i + 1, j + 1)]};
      factory.AddEntity(RefEl::kSegment(), nl, seg(blx + i * hx, bly + j * hy, blx + (i + 1) * hx, bly + (j + 1) * hy));
    }
  }
  for (size_type i = 0; i < nx; ++i) {
    for (size_type j = 0; j < ny; ++j) {
      {  // triangle above the diagonal
        const std::array<size_type, 3> vl{v_idx[VertexIndex(i, j)], v_idx[VertexIndex(i + 1, j + 1)], v_idx[VertexIndex(i, j + 1)]};
        Mat g(2, 3);
        g(0, 0) = blx + i * hx; g(0, 1) = blx + (i + 1) * hx; g(0, 2) = blx + i * hx;
        g(1, 0) = bly + j * hy; g(1, 1) = bly + (j + 1) * hy; g(1, 2) = bly + (j + 1) * hy;
        factory.AddEntity(RefEl::kTria(), vl, std::make_unique<geometry::TriaO1>(g));
      }
      {  // triangle below the diagonal
        const std::array<size_type, 3> vl{v_idx[VertexIndex(i, j)], v_idx[VertexIndex(i + 1, j)], v_idx[VertexIndex(i + 1, j + 1)]};
        Mat g(2, 3);
        g(0, 0) = blx + i * hx; g(0, 1) = blx + (i + 1) * hx; g(0, 2) = blx + (i + 1) * hx;
        g(1, 0) = bly + j * hy; g(1, 1) = bly + j * hy; g(1, 2) = bly + (j + 1) * hy;
        factory.AddEntity(RefEl::kTria(), vl, std::make_unique<geometry::TriaO1>(g));
      }
    }
  }
  return factory.Build();
}

// utils/tp_quad_mesh_builder.cc:19-95 (no explicit edges)
inline std::shared_ptr<Mesh> TPQuadMeshBuild(size_type nx, size_type ny, double blx, double bly, double trx, double try_) {
  hybrid2d::MeshFactory factory;
  if (nx * ny == 0) return nullptr;
  const double x_size = trx - blx, y_size = try_ - bly;
  if (x_size <= 0.0 || y_size <= 0.0) return nullptr;
  const double hx = x_size / nx, hy = y_size / ny;
  auto VertexIndex = [nx](size_type i, size_type j) { return i + j * (nx + 1); };
  std::vector<size_type> v_idx((nx + 1) * (ny + 1));
  int node_cnt = 0;
  for (size_type j = 0; j <= ny; ++j) {
    for (size_type i = 0; i <= nx; ++i, ++node_cnt) v_idx[node_cnt] = factory.AddPoint(blx + i * hx, bly + j * hy);
  }
  for (size_type i = 0; i < nx; ++i) {
    for (size_type j = 0; j < ny; ++j) {
      const std::array<size_type, 4> vl{v_idx[VertexIndex(i, j)], v_idx[VertexIndex(i + 1, j)],
                                        v_idx[VertexIndex(i + 1, j + 1)], v_idx[VertexIndex(i, j + 1)]};
      Mat g(2, 4);
      g(0, 0) = blx + i * hx; g(0, 1) = blx + (i + 1) * hx; g(0, 2) = blx + (i + 1) * hx; g(0, 3) = blx + i * hx;
      g(1, 0) = bly + j * hy; g(1, 1) = bly + j * hy; g(1, 2) = bly + (j + 1) * hy; g(1, 3) = bly + (j + 1) * hy;
      factory.AddEntity(RefEl::kQuad(), vl, std::make_unique<geometry::QuadO1>(g));
    }
  }
  return factory.Build();
}

// Synthetic hybrid mesh of config C2 (the reference has no such generator -- the SPEC is ours,
// stated in DESIGN.md, and the product implements it independently on the device):
//   n x n squares on [0,1]^2, nodes i + j(n+1) at (i h, j h) with h = 1/n; interior nodes are displaced by
//   jitter*h*(2u-1) per coordinate, u = (splitmix64(seed + 2*node + d) >> 11) * 2^-53;
//   squares visited i outer / j inner; square (i,j) is one QuadO1 if (i+j) even, else the two triangles of
//   tp_triag_mesh_builder.cc:144-176 ("upper" first); no explicit edges, no explicit geometries.
inline std::uint64_t SplitMix64(std::uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
inline std::shared_ptr<Mesh> HybridMeshBuild(size_type n, double jitter, std::uint64_t seed) {
  hybrid2d::MeshFactory factory;
  if (n == 0) return nullptr;
  const double h = 1.0 / n;
  auto VertexIndex = [n](size_type i, size_type j) { return i + j * (n + 1); };
  for (size_type j = 0; j <= n; ++j) {
    for (size_type i = 0; i <= n; ++i) {
      double x = i * h, y = j * h;
      if (i > 0 && i < n && j > 0 && j < n) {
        const std::uint64_t node = VertexIndex(i, j);
        const double u0 = static_cast<double>(SplitMix64(seed + 2 * node) >> 11) * 0x1.0p-53;
        const double u1 = static_cast<double>(SplitMix64(seed + 2 * node + 1) >> 11) * 0x1.0p-53;
        x += jitter * h * (2.0 * u0 - 1.0);
        y += jitter * h * (2.0 * u1 - 1.0);
      }
      factory.AddPoint(x, y);
    }
  }
  for (size_type i = 0; i < n; ++i) {
    for (size_type j = 0; j < n; ++j) {
      if ((i + j) % 2 == 0) {
        const std::array<size_type, 4> vl{VertexIndex(i, j), VertexIndex(i + 1, j), VertexIndex(i + 1, j + 1), VertexIndex(i, j + 1)};
        factory.AddEntity(RefEl::kQuad(), vl, nullptr);
      } else {
        const std::array<size_type, 3> up{VertexIndex(i, j), VertexIndex(i + 1, j + 1), VertexIndex(i, j + 1)};
        const std::array<size_type, 3> lo{VertexIndex(i, j), VertexIndex(i + 1, j), VertexIndex(i + 1, j + 1)};
        factory.AddEntity(RefEl::kTria(), up, nullptr);
        factory.AddEntity(RefEl::kTria(), lo, nullptr);
      }
    }
  }
  return factory.Build();
}

}  // namespace utils
}  // namespace lfo::mesh
#endif
