// The reference's reader tests (lib/lf/io/test/gmsh_reader_tests.cc: checkTwoElementMesh :23-140, checkPieceOfCake :232-300,
// readLectureDemoMesh :184-188) written against the shim's lfgpu::GmshReader (include/lf_gpu_shim.hpp).  Host only: the
// reader is host code behind the C ABI, so this runs without a GPU.  Entities are (codim, index) instead of Entity
// references; geometric look-ups ("the node at the origin", "the diagonal edge") use the arrays the reader hands out.
#include <cmath>
#include <cstdio>
#include <string>

#include "../../include/lf_gpu_shim.hpp"

static int failures = 0;
#define EXPECT(cond)                                                      \
  do {                                                                    \
    if (!(cond)) {                                                        \
      std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);         \
      ++failures;                                                         \
    }                                                                     \
  } while (0)
#define EXPECT_THROW(expr)                                                \
  do {                                                                    \
    bool thrown = false;                                                  \
    try {                                                                 \
      (void)(expr);                                                       \
    } catch (const lfgpu::Error&) {                                       \
      thrown = true;                                                      \
    }                                                                     \
    if (!thrown) {                                                        \
      std::printf("FAIL %s:%d: %s did not throw\n", __FILE__, __LINE__, #expr); \
      ++failures;                                                         \
    }                                                                     \
  } while (0)

using pe_t = std::pair<unsigned, std::string>;
template <class V, class T>
static bool contains(const V& v, const T& x) { return std::find(v.begin(), v.end(), x) != v.end(); }

static void checkTwoElementMesh(const lfgpu::GmshReader& reader) {
  std::vector<std::uint32_t> edges;
  const lfgpu::FlatMesh m = reader.Flat(&edges);
  EXPECT(reader.NumEntities(0) == 2);
  EXPECT(reader.NumEntities(2) == 5);
  EXPECT(edges.size() == 2);  // one explicitly listed edge; the mesh has 6 (numbered on the device)
  // codim = 2
  std::int64_t origin = -1;
  for (std::int64_t i = 0; i < m.n_nodes; ++i)
    if (m.node_coords[2 * i] * m.node_coords[2 * i] + m.node_coords[2 * i + 1] * m.node_coords[2 * i + 1] < 1e-10) origin = i;
  EXPECT(origin >= 0);
  EXPECT((reader.PhysicalEntityNr(2, origin) == std::vector<unsigned>{1, 2}));
  EXPECT(reader.PhysicalEntityNr2Name(1, 2) == "physicalEntity1");
  EXPECT_THROW(reader.PhysicalEntityNr2Name(1));
  EXPECT(reader.PhysicalEntityNr2Name(2, 2) == "physicalEntity2");
  EXPECT(reader.PhysicalEntityNr2Name(2) == "physicalEntity2");
  EXPECT(reader.PhysicalEntityName2Nr("physicalEntity1", 2) == 1);
  EXPECT(reader.PhysicalEntityName2Nr("physicalEntity2", 2) == 2);
  EXPECT_THROW(reader.PhysicalEntityName2Nr("physicalEntity1"));
  EXPECT(reader.PhysicalEntityName2Nr("physicalEntity2") == 2);
  EXPECT_THROW(reader.PhysicalEntityNr2Name(100));
  EXPECT_THROW(reader.PhysicalEntityName2Nr("gugus"));
  const auto pe2 = reader.PhysicalEntities(2);
  EXPECT(pe2.size() == 2);
  EXPECT(contains(pe2, pe_t{1, "physicalEntity1"}));
  EXPECT(contains(pe2, pe_t{2, "physicalEntity2"}));
  for (std::int64_t i = 0; i < m.n_nodes; ++i)
    if (i != origin) EXPECT(reader.PhysicalEntityNr(2, i).empty());
  // codim = 1: the explicitly listed edge is the diagonal
  const double dx = m.node_coords[2 * edges[1]] - m.node_coords[2 * edges[0]], dy = m.node_coords[2 * edges[1] + 1] - m.node_coords[2 * edges[0] + 1];
  EXPECT(std::sqrt(dx * dx + dy * dy) > 1.1);
  const unsigned diagonal_nr = reader.PhysicalEntityName2Nr("diagonal");
  EXPECT(reader.PhysicalEntityNr2Name(diagonal_nr) == "diagonal");
  EXPECT((reader.PhysicalEntityNr(1, 0) == std::vector<unsigned>{diagonal_nr}));
  const auto pe1 = reader.PhysicalEntities(1);
  EXPECT(pe1.size() == 1 && pe1[0].first == 4 && pe1[0].second == "diagonal");
  for (std::int64_t e = 1; e < 6; ++e) EXPECT(reader.PhysicalEntityNr(1, e).empty());
  // codim = 0
  std::int64_t square = -1, triangle = -1;
  for (std::int64_t c = 0; c < m.n_cells; ++c) (m.cell_nodes[4 * c + 3] == LFGPU_IDX_NIL ? triangle : square) = c;
  EXPECT(square >= 0 && triangle >= 0);
  const unsigned square_nr = reader.PhysicalEntityName2Nr("square");
  EXPECT(reader.PhysicalEntityNr2Name(square_nr) == "square");
  EXPECT((reader.PhysicalEntityNr(0, square) == std::vector<unsigned>{square_nr}));
  EXPECT(reader.PhysicalEntityName2Nr("physicalEntity1", 0) == 1);
  EXPECT(reader.PhysicalEntityName2Nr("physicalEntity3") == 3);
  EXPECT(reader.PhysicalEntityNr2Name(1, 0) == "physicalEntity1");
  EXPECT(reader.PhysicalEntityNr2Name(3) == "physicalEntity3");
  EXPECT(reader.PhysicalEntityNr2Name(3, 0) == "physicalEntity3");
  EXPECT_THROW(reader.PhysicalEntityNr2Name(3, 1));
  EXPECT((reader.PhysicalEntityNr(0, triangle) == std::vector<unsigned>{1, 3}));
  const auto pe0 = reader.PhysicalEntities(0);
  EXPECT(pe0.size() == 3);
  EXPECT(contains(pe0, pe_t{1, "physicalEntity1"}));
  EXPECT(contains(pe0, pe_t{3, "physicalEntity3"}));
  EXPECT(contains(pe0, pe_t{5, "square"}));
}

static void checkPieceOfCake(const lfgpu::GmshReader& reader) {
  std::vector<std::uint32_t> edges;
  const lfgpu::FlatMesh m = reader.Flat(&edges);
  EXPECT(reader.NumEntities(0) == 2);
  EXPECT(reader.NumEntities(2) == 4);
  std::int64_t origin = -1;
  for (std::int64_t i = 0; i < m.n_nodes; ++i)
    if (std::hypot(m.node_coords[2 * i], m.node_coords[2 * i + 1]) < 1e-5) origin = i;
  EXPECT(origin >= 0);
  EXPECT((reader.PhysicalEntityNr(2, origin) == std::vector<unsigned>{1}));
  EXPECT(reader.IsPhysicalEntity(2, origin, 1));
  for (std::size_t e = 0; e < edges.size() / 2; ++e) {
    const double r0 = std::hypot(m.node_coords[2 * edges[2 * e]], m.node_coords[2 * edges[2 * e] + 1]);
    const double r1 = std::hypot(m.node_coords[2 * edges[2 * e + 1]], m.node_coords[2 * edges[2 * e + 1] + 1]);
    if (std::abs(r0 - 1) < 1e-6 && std::abs(r1 - 1) < 1e-6) {
      EXPECT((reader.PhysicalEntityNr(1, static_cast<std::int64_t>(e)) == std::vector<unsigned>{2}));
      EXPECT(reader.IsPhysicalEntity(1, static_cast<std::int64_t>(e), 2));
    } else {
      EXPECT(!reader.IsPhysicalEntity(1, static_cast<std::int64_t>(e), 2));
    }
  }
  const auto arc = reader.PhysicalEntityFlags(1, 2, 5);
  EXPECT(arc[0] + arc[1] + arc[2] + arc[3] + arc[4] == 2);
  for (std::int64_t c = 0; c < 2; ++c) {
    EXPECT((reader.PhysicalEntityNr(0, c) == std::vector<unsigned>{3}));
    EXPECT(reader.IsPhysicalEntity(0, c, 3));
  }
  EXPECT(reader.PhysicalEntityName2Nr("origin") == 1);
  EXPECT(reader.PhysicalEntityName2Nr("arc") == 2);
  EXPECT(reader.PhysicalEntityNr2Name(1) == "origin");
  EXPECT(reader.PhysicalEntityNr2Name(1, 2) == "origin");
  EXPECT(reader.PhysicalEntityNr2Name(2) == "arc");
  EXPECT(reader.PhysicalEntityNr2Name(2, 1) == "arc");
  EXPECT_THROW(reader.PhysicalEntityNr2Name(3));
  EXPECT_THROW(reader.PhysicalEntityNr2Name(3, 1));
  EXPECT(reader.PhysicalEntities(0).empty());
  EXPECT(reader.PhysicalEntities(1).size() == 1 && reader.PhysicalEntities(1)[0] == (pe_t{2, "arc"}));
  EXPECT(reader.PhysicalEntities(2).size() == 1 && reader.PhysicalEntities(2)[0] == (pe_t{1, "origin"}));
}

int main(int argc, char** argv) {
  const std::string dir = argc > 1 ? argv[1] : "tests/golden/msh";
  for (const char* name : {"two_element_hybrid_2d.msh", "two_element_hybrid_2d_binary.msh", "two_element_hybrid_2d_v4.msh",
                           "two_element_hybrid_2d_v4_binary.msh", "two_element_hybrid_2d_second_order.msh",
                           "two_element_hybrid_2d_second_order_v4.msh"}) {
    lfgpu::GmshReader reader(dir + "/" + name);
    checkTwoElementMesh(reader);
    EXPECT(reader.GeometryOrder() == (std::string(name).find("second_order") != std::string::npos ? 2 : 1));
  }
  { lfgpu::GmshReader reader(dir + "/lecturedemomesh.msh"); EXPECT(reader.NumEntities(0) == 5); }  // trailing blank at the end of a line
  { lfgpu::GmshReader reader(dir + "/piece_of_cake.msh"); checkPieceOfCake(reader); }
  for (const char* name : {"curved_square_quads_2nd_order.msh", "curved_square_trias_2nd_order.msh", "curved_square_quads_2nd_order_v4.msh",
                           "curved_square_trias_2nd_order_v4.msh"}) {
    lfgpu::GmshReader reader(dir + "/" + name);  // curvedSquareTests: the files can be read
    EXPECT(reader.NumEntities(0) > 0);
  }
  EXPECT_THROW(lfgpu::GmshReader(dir + "/does_not_exist.msh"));
  std::printf(failures == 0 ? "GMSH_SHIM_TEST_OK\n" : "GMSH_SHIM_TEST_FAILED\n");
  return failures == 0 ? 0 : 1;
}
