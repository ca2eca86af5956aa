// ORACLE (test infrastructure, NOT product code) -- see lfo_base.h header.
// Regular refinement of a hybrid 2D mesh with the reference's numbering: the part of
//   lib/lf/refinement/mesh_hierarchy.cc:368-1262  (MeshHierarchy::PerformRefinement)
// that MeshHierarchy::RefineRegular(rp_regular) (:72-114) exercises -- every edge is split (rp_split), every cell is
// refined regularly, no anchors -- together with the child geometries of
//   lib/lf/refinement/hybrid2d_refinement_pattern.cc:281-983 (ChildPolygons on the lattice with constant 6)
//   lib/lf/geometry/{segment_o1,tria_o1,quad_o1}.cc ChildGeometry = parent Global(lattice point / 6).
//
// Numbering produced (the order of the MeshFactory calls of the reference):
//   nodes : copies of the parent nodes (same index), then one midpoint per parent edge in edge order, then -- inside the
//           cell loop -- the centre of every quadrilateral in cell order                       (:393-412, :474-476, :1099)
//   edges : all supplied explicitly: (p0, mid), (mid, p1) per parent edge in edge order (:490-507), then per parent cell
//           its interior edges: triangle (m0,m2), (m0,m1), (m2,m1) (:881-890); quad (m_k, centre), k = 0..3 (:1126-1137)
//   cells : per parent cell four children: triangle (v0,m0,m2), (v1,m0,m1), (v2,m2,m1), (m0,m1,m2) (:862-879);
//           quad (v0,m0,c,m3), (v1,m1,c,m0), (v2,m1,c,m2), (v3,m2,c,m3) (:1103-1122)
//   where m_j is the midpoint of local edge j.  Every child carries its own geometry (corner coordinates from the PARENT
//   cell's map), which is why the assembler must read cell corners from the cell geometry.
//
// PINNING: the reference's refinement tests check relations (father/child consistency), not literal index tables, so
// this restatement is checked by invariants only (tests/test_oracle_refinement.py): "parity unpinned" for numbering.
#ifndef LFO_REFINEMENT_H
#define LFO_REFINEMENT_H

#include "lfo_mesh.h"

namespace lfo::refinement {

inline std::shared_ptr<mesh::Mesh> RefineRegular(const mesh::Mesh& parent) {
  mesh::hybrid2d::MeshFactory f;
  const double h_lattice = 1.0 / 6.0;  // RefinementPattern::lattice_const_ = 6 (geometry/refinement_pattern.h:48)
  const Mat zero_point(0, 1);
  auto lattice = [&](std::initializer_list<std::array<int, 2>> pts) {  // reference coordinates h * lattice point, 2 x n
    Mat m(2, static_cast<long>(pts.size()));
    long k = 0;
    for (const auto& p : pts) {
      m(0, k) = h_lattice * static_cast<double>(p[0]);
      m(1, k) = h_lattice * static_cast<double>(p[1]);
      ++k;
    }
    return m;
  };
  // ---- nodes: rp_copy (:393-412)
  for (const mesh::Entity* node : parent.Entities(2)) {
    const Mat x = node->Geometry()->Global(zero_point);
    f.AddPoint(x(0, 0), x(1, 0));
  }
  // ---- edges: rp_split (:423-512)
  std::vector<size_type> midpoint(parent.NumEntities(1));
  for (const mesh::Entity* edge : parent.Entities(1)) {
    const size_type e = parent.Index(*edge);
    const auto ends = edge->SubEntities(1);
    const size_type p0 = parent.Index(*ends[0]), p1 = parent.Index(*ends[1]);
    const geometry::Geometry* g = edge->Geometry();
    Mat t(1, 1);
    t(0, 0) = h_lattice * 3.0;
    const Mat mid = g->Global(t);
    midpoint[e] = f.AddPoint(mid(0, 0), mid(1, 0));
    Mat a(1, 2), b(1, 2);
    a(0, 0) = h_lattice * 0.0;
    a(0, 1) = h_lattice * 3.0;
    b(0, 0) = h_lattice * 3.0;
    b(0, 1) = h_lattice * 6.0;
    const std::array<size_type, 2> n0{p0, midpoint[e]}, n1{midpoint[e], p1};
    f.AddEntity(RefEl::kSegment(), n0, std::make_unique<geometry::SegmentO1>(g->Global(a)));
    f.AddEntity(RefEl::kSegment(), n1, std::make_unique<geometry::SegmentO1>(g->Global(b)));
  }
  // ---- cells: rp_regular (:525-1262)
  for (const mesh::Entity* cell : parent.Entities(0)) {
    const geometry::Geometry* g = cell->Geometry();
    const auto nodes = cell->SubEntities(2);
    const auto edges = cell->SubEntities(1);
    if (cell->RefElem() == RefEl::kTria()) {
      const size_type v[3] = {parent.Index(*nodes[0]), parent.Index(*nodes[1]), parent.Index(*nodes[2])};
      const size_type m[3] = {midpoint[parent.Index(*edges[0])], midpoint[parent.Index(*edges[1])], midpoint[parent.Index(*edges[2])]};
      // lattice: vertices (0,0) (6,0) (0,6); edge midpoints (3,0) (3,3) (0,3)
      const std::array<int, 2> V[3] = {{0, 0}, {6, 0}, {0, 6}}, M[3] = {{3, 0}, {3, 3}, {0, 3}};
      // interior edges first (:1171-1188), then the children (:1190-1230)
      const std::array<size_type, 2> en[3] = {{m[0], m[2]}, {m[0], m[1]}, {m[2], m[1]}};
      const Mat eg[3] = {g->Global(lattice({M[0], M[2]})), g->Global(lattice({M[0], M[1]})), g->Global(lattice({M[2], M[1]}))};
      for (int k = 0; k < 3; ++k) f.AddEntity(RefEl::kSegment(), en[k], std::make_unique<geometry::SegmentO1>(eg[k]));
      const std::array<size_type, 3> cn[4] = {{v[0], m[0], m[2]}, {v[1], m[0], m[1]}, {v[2], m[2], m[1]}, {m[0], m[1], m[2]}};
      const Mat cg[4] = {g->Global(lattice({V[0], M[0], M[2]})), g->Global(lattice({V[1], M[0], M[1]})),
                         g->Global(lattice({V[2], M[2], M[1]})), g->Global(lattice({M[0], M[1], M[2]}))};
      for (int k = 0; k < 4; ++k) f.AddEntity(RefEl::kTria(), cn[k], std::make_unique<geometry::TriaO1>(cg[k]));
    } else {
      const size_type v[4] = {parent.Index(*nodes[0]), parent.Index(*nodes[1]), parent.Index(*nodes[2]), parent.Index(*nodes[3])};
      const size_type m[4] = {midpoint[parent.Index(*edges[0])], midpoint[parent.Index(*edges[1])], midpoint[parent.Index(*edges[2])],
                              midpoint[parent.Index(*edges[3])]};
      const std::array<int, 2> V[4] = {{0, 0}, {6, 0}, {6, 6}, {0, 6}}, M[4] = {{3, 0}, {6, 3}, {3, 6}, {0, 3}}, C = {3, 3};
      const Mat cpt = g->Global(lattice({C}));
      const size_type c = f.AddPoint(cpt(0, 0), cpt(1, 0));  // :1090-1100
      for (int k = 0; k < 4; ++k) {
        const std::array<size_type, 2> en{m[k], c};
        f.AddEntity(RefEl::kSegment(), en, std::make_unique<geometry::SegmentO1>(g->Global(lattice({M[k], C}))));
      }
      const std::array<size_type, 4> cn[4] = {{v[0], m[0], c, m[3]}, {v[1], m[1], c, m[0]}, {v[2], m[1], c, m[2]}, {v[3], m[2], c, m[3]}};
      const Mat cg[4] = {g->Global(lattice({V[0], M[0], C, M[3]})), g->Global(lattice({V[1], M[1], C, M[0]})),
                         g->Global(lattice({V[2], M[1], C, M[2]})), g->Global(lattice({V[3], M[2], C, M[3]}))};
      for (int k = 0; k < 4; ++k) f.AddEntity(RefEl::kQuad(), cn[k], std::make_unique<geometry::QuadO1>(cg[k]));
    }
  }
  return f.Build();
}

}  // namespace lfo::refinement
#endif
