// Gmsh input (product code, host side): stands in for lf::io::GmshReader on the way INTO the device mesh.
//
// Reference (lib/lf/io/): ReadGmshFile gmsh_reader.cc:629-697, the MSH 2.2 / 4.1 file structures gmsh_file_v2.h:27-268 and
// gmsh_file_v4.h:29-540, and GmshReader::InitGmshFile gmsh_reader.cc:121-340 (2.2), :343-627 (4.1).
//
// What matters for the assembly path is the NUMBERING reader.mesh() ends up with, because dof numbers follow entity
// indices: nodes = the main nodes (vertices of elements) in file order, explicitly listed edges in file order ahead of all
// other edges, cells in file order, with consecutive repetitions of one element merged (they only add physical numbers).
// This file produces exactly those arrays from the bytes of the file in ONE pass over a flat buffer (no per-element heap
// objects); lfgpu_gmsh_mesh hands them to lfgpu_mesh_upload + lfgpu_mesh_build_topology, which numbers the remaining
// edges on the device the way hybrid2d::MeshFactory::Build does.
#include <algorithm>
#include <array>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "lfgpu_internal.cuh"

namespace {

struct ParseError {
  std::string msg;
};
// message for the caller: file content quoted in it is cut to 40 characters and made printable
[[noreturn]] void fail(const std::string& m) { throw ParseError{m}; }
std::string shown(const std::string& t) {
  std::string s = t.substr(0, 40);
  for (char& ch : s)
    if (!std::isprint(static_cast<unsigned char>(ch))) ch = '?';
  return s;
}

// number of nodes and dimension of the Gmsh element types (gmsh_file_v2.h:33-101)
bool element_info(int type, int* n_nodes, int* dim) {
  static const struct { int type, nn, dim; } kTab[] = {
      {1, 2, 1},   {2, 3, 2},   {3, 4, 2},   {4, 4, 3},   {5, 8, 3},   {6, 6, 3},   {7, 5, 3},   {8, 3, 1},   {9, 6, 2},
      {10, 9, 2},  {11, 10, 3}, {12, 27, 3}, {13, 18, 3}, {14, 14, 3}, {15, 1, 0},  {16, 8, 2},  {17, 20, 3}, {18, 15, 3},
      {19, 13, 3}, {20, 9, 2},  {21, 10, 2}, {22, 12, 2}, {23, 15, 2}, {24, 15, 2}, {25, 21, 2}, {26, 4, 1},  {27, 5, 1},
      {28, 6, 1},  {29, 20, 3}, {30, 35, 3}, {31, 56, 3}, {92, 64, 3}, {93, 125, 3}};
  for (const auto& e : kTab) {
    if (e.type == type) {
      *n_nodes = e.nn;
      *dim = e.dim;
      return true;
    }
  }
  return false;
}
// element types the reader turns into mesh entities (gmsh_reader.cc:259-292): main nodes, geometry order; 0 = unsupported
int main_nodes_of(int type, int* order) {
  switch (type) {
    case 1: *order = 1; return 2;
    case 8: *order = 2; return 2;
    case 2: *order = 1; return 3;
    case 9: *order = 2; return 3;
    case 3: *order = 1; return 4;
    case 16: case 10: *order = 2; return 4;
    default: return 0;
  }
}

// cursor over the file bytes: whitespace-separated text tokens and raw little/big endian binary fields
class Cursor {
 public:
  Cursor(const char* d, size_t n) : d_(d), n_(n) {}
  void skip_ws() {
    while (p_ < n_ && std::isspace(static_cast<unsigned char>(d_[p_]))) ++p_;
  }
  bool at_end() {
    skip_ws();
    return p_ >= n_;
  }
  std::string token() {
    skip_ws();
    const size_t b = p_;
    while (p_ < n_ && !std::isspace(static_cast<unsigned char>(d_[p_]))) ++p_;
    if (p_ == b) fail("unexpected end of file");
    return std::string(d_ + b, p_ - b);
  }
  long long integer() {
    const std::string t = token();
    char* end = nullptr;
    const long long v = std::strtoll(t.c_str(), &end, 10);
    if (end == t.c_str() || *end != '\0') fail("expected an integer, found '" + shown(t) + "'");
    return v;
  }
  double real() {
    const std::string t = token();
    char* end = nullptr;
    const double v = std::strtod(t.c_str(), &end);
    if (end == t.c_str() || *end != '\0') fail("expected a number, found '" + shown(t) + "'");
    return v;
  }
  void expect(const char* word) {
    const std::string t = token();
    if (t != word) fail(std::string("expected ") + word + ", found '" + shown(t) + "'");
  }
  std::string quoted() {
    skip_ws();
    if (p_ >= n_ || d_[p_] != '"') fail("expected a quoted string");
    const size_t b = ++p_;
    while (p_ < n_ && d_[p_] != '"') ++p_;
    if (p_ >= n_) fail("unterminated string");
    return std::string(d_ + b, p_++ - b);
  }
  void eol() {  // the single line break in front of a binary payload
    if (p_ + 1 < n_ && d_[p_] == '\r' && d_[p_ + 1] == '\n') p_ += 2;
    else if (p_ < n_ && d_[p_] == '\n') p_ += 1;
    else fail("expected end of line before binary data");
  }
  template <typename T>
  T raw(bool swap) {
    if (p_ + sizeof(T) > n_) fail("binary payload truncated");
    unsigned char b[sizeof(T)];
    std::memcpy(b, d_ + p_, sizeof(T));
    p_ += sizeof(T);
    if (swap) std::reverse(b, b + sizeof(T));
    T v;
    std::memcpy(&v, b, sizeof(T));
    return v;
  }
  // bytes left: an upper bound for any count the rest of the file can honour (corrupt counts must not drive allocations)
  size_t remaining() const { return n_ - p_; }
  void skip_section(const std::string& name) {
    const std::string end = "$End" + name.substr(1);
    const char* b = d_ + p_;
    const char* e = d_ + n_;
    const char* q = std::search(b, e, end.begin(), end.end());
    if (q == e) fail("section " + shown(name) + " is not closed");
    p_ = static_cast<size_t>(q - d_) + end.size();
  }

 private:
  const char* d_;
  size_t n_;
  size_t p_ = 0;
};

struct PhysicalName {
  int dim;
  uint32_t nr;
  std::string name;
};

}  // namespace

// the flattened result: what InitGmshFile hands to the MeshFactory + the physical-entity tables
struct lfgpu_gmsh {
  int dim_world = 2;
  int order = 1;
  std::vector<double> node_xy;          // [n_nodes][2]
  std::vector<uint32_t> edge_nodes;     // [n_explicit][2]
  std::vector<uint32_t> cell_nodes;     // [n_cells][4]
  // physical numbers per entity: ent_phys[codim][entity index] (short lists; the node table grows on demand)
  std::vector<std::vector<uint32_t>> ent_phys[3];
  std::vector<PhysicalName> names;
};

namespace {

// shared by both format versions: collects points / entities in AddPoint / AddEntity order
class Builder {
 public:
  explicit Builder(lfgpu_gmsh* g) : g_(g) {}
  std::unordered_map<uint64_t, uint32_t> gi2mi;                 // gmsh node tag -> mesh node index (main nodes only)

  void add_point(uint64_t tag, double x, double y, double z) {
    if (g_->dim_world == 2 && z != 0.0) fail("In a 2D GmshMesh, the z-coordinate of every node must be zero");
    gi2mi[tag] = static_cast<uint32_t>(g_->node_xy.size() / 2);
    g_->node_xy.push_back(x);
    g_->node_xy.push_back(y);
  }
  // one element of the file; `same_as_previous` = consecutive repetition (only its physical numbers are recorded)
  void element(int type, const uint64_t* nodes, const uint32_t* phys, size_t n_phys, bool same_as_previous) {
    if (same_as_previous) {
      if (last_ != nullptr) last_->insert(last_->end(), phys, phys + n_phys);
      return;
    }
    if (type == 15) {
      auto it = gi2mi.find(nodes[0]);
      if (it == gi2mi.end()) {  // auxiliary node: not part of the mesh (gmsh_reader.cc:235-245)
        last_ = nullptr;
        return;
      }
      auto& lists = g_->ent_phys[2];
      if (lists.size() <= it->second) lists.resize(it->second + 1);
      lists[it->second].insert(lists[it->second].end(), phys, phys + n_phys);
      last_ = &lists.back();  // a repetition is appended to mi2gi[dim_mesh].back() (:224-229): the highest index so far
      return;
    }
    int order = 1;
    const int nm = main_nodes_of(type, &order);
    if (nm == 0) fail("Gmsh element type " + std::to_string(type) + " not (yet) supported by GmshReader.");
    g_->order = std::max(g_->order, order);
    uint32_t mi[4] = {LFGPU_IDX_NIL, LFGPU_IDX_NIL, LFGPU_IDX_NIL, LFGPU_IDX_NIL};
    for (int i = 0; i < nm; ++i) {
      auto it = gi2mi.find(nodes[i]);
      if (it == gi2mi.end()) fail("element refers to node " + std::to_string(nodes[i]) + " which the file does not define");
      mi[i] = it->second;
    }
    const int codim = (nm == 2) ? 1 : 0;
    if (codim == 1) g_->edge_nodes.insert(g_->edge_nodes.end(), mi, mi + 2);
    else g_->cell_nodes.insert(g_->cell_nodes.end(), mi, mi + 4);
    g_->ent_phys[codim].emplace_back(phys, phys + n_phys);
    last_ = &g_->ent_phys[codim].back();
  }

 private:
  lfgpu_gmsh* g_;
  std::vector<uint32_t>* last_ = nullptr;
};

void read_physical_names(Cursor& c, lfgpu_gmsh* g) {
  const long long n = c.integer();
  for (long long i = 0; i < n; ++i) {
    PhysicalName p;
    p.dim = static_cast<int>(c.integer());
    p.nr = static_cast<uint32_t>(c.integer());
    p.name = c.quoted();
    g->names.push_back(std::move(p));
  }
  c.expect("$EndPhysicalNames");
}

// ---- MSH 2.2 (gmsh_file_v2.cc:425-735 + gmsh_reader.cc:121-340) -------------------------------------------------------
struct V2Element {
  int type;
  uint32_t physical;
  size_t first_node;  // into the flat node-number array
};

void parse_v2(Cursor& c, bool binary, bool swap, lfgpu_gmsh* g) {
  std::vector<std::pair<uint64_t, std::array<double, 3>>> nodes;
  std::vector<V2Element> elems;
  std::vector<uint64_t> elem_nodes;
  while (!c.at_end()) {
    const std::string sec = c.token();
    if (sec == "$PhysicalNames") {
      read_physical_names(c, g);
    } else if (sec == "$Nodes") {
      const long long n = c.integer();
      if (n < 0) fail("negative number of nodes");
      nodes.reserve(std::min(static_cast<size_t>(n), c.remaining()));
      if (binary) c.eol();
      for (long long i = 0; i < n; ++i) {
        uint64_t tag;
        std::array<double, 3> p;
        if (binary) {
          tag = static_cast<uint32_t>(c.raw<int32_t>(swap));
          for (double& v : p) v = c.raw<double>(swap);
        } else {
          tag = static_cast<uint64_t>(c.integer());
          for (double& v : p) v = c.real();
        }
        nodes.emplace_back(tag, p);
      }
      c.expect("$EndNodes");
    } else if (sec == "$Elements") {
      const long long n = c.integer();
      if (n < 0) fail("negative number of elements");
      elems.reserve(std::min(static_cast<size_t>(n), c.remaining()));
      auto push = [&](int type, const std::vector<long long>& tags, int nn, auto&& next_node) {
        if (tags.size() < 2) fail("element with fewer than two tags");
        elems.push_back({type, static_cast<uint32_t>(tags[0]), elem_nodes.size()});
        for (int k = 0; k < nn; ++k) elem_nodes.push_back(static_cast<uint64_t>(next_node()));
      };
      if (binary) {
        c.eol();
        long long done = 0;
        while (done < n) {
          const int type = c.raw<int32_t>(swap), count = c.raw<int32_t>(swap), ntags = c.raw<int32_t>(swap);
          int nn = 0, dim = 0;
          if (!element_info(type, &nn, &dim) || count < 0 || ntags < 0 || static_cast<size_t>(ntags) > c.remaining())
            fail("unknown element type " + std::to_string(type) + " or corrupt block header");
          for (int e = 0; e < count; ++e) {
            c.raw<int32_t>(swap);  // element number
            std::vector<long long> tags(static_cast<size_t>(ntags));
            for (auto& t : tags) t = c.raw<int32_t>(swap);
            push(type, tags, nn, [&] { return static_cast<long long>(c.raw<int32_t>(swap)); });
          }
          done += count;
        }
      } else {
        for (long long i = 0; i < n; ++i) {
          c.integer();  // element number
          const int type = static_cast<int>(c.integer());
          const long long ntags = c.integer();
          int nn = 0, dim = 0;
          if (!element_info(type, &nn, &dim) || ntags < 0 || static_cast<size_t>(ntags) > c.remaining())
            fail("unknown element type " + std::to_string(type) + " or corrupt tag count");
          std::vector<long long> tags(static_cast<size_t>(ntags));
          for (auto& t : tags) t = c.integer();
          push(type, tags, nn, [&] { return c.integer(); });
        }
      }
      c.expect("$EndElements");
    } else if (!sec.empty() && sec[0] == '$') {
      c.skip_section(sec);  // $Periodic is parsed and ignored by GmshReader (gmsh_reader.cc:333-339); comment sections
    } else {
      fail("Could not parse file: unexpected '" + shown(sec) + "'");
    }
  }
  // gmsh_reader.cc:131-175: the main nodes are the vertices of EVERY non-point element
  Builder b(g);
  std::unordered_map<uint64_t, char> is_main;
  long long n_top = 0;
  for (const V2Element& e : elems) {
    int nn = 0, dim = 0;
    element_info(e.type, &nn, &dim);
    if (dim > 2) fail("mesh_factory->DimMesh() = 2, but msh-file contains entities with dimension " + std::to_string(dim));
    n_top += (dim == 2);
    if (e.type == 15) continue;
    int order = 1;
    const int nm = main_nodes_of(e.type, &order);
    if (nm == 0) fail("Gmsh element type " + std::to_string(e.type) + " not (yet) supported by GmshReader.");
    for (int k = 0; k < nm; ++k) is_main[elem_nodes[e.first_node + k]] = 1;
  }
  if (n_top == 0) fail("MshFile contains no elements with dimension 2");
  for (const auto& nd : nodes) {
    if (is_main.count(nd.first)) b.add_point(nd.first, nd.second[0], nd.second[1], nd.second[2]);
  }
  // gmsh_reader.cc:215-300
  const V2Element* run = nullptr;
  for (const V2Element& e : elems) {
    int nn = 0, dim = 0;
    element_info(e.type, &nn, &dim);
    bool same = false;
    if (run != nullptr && run->type == e.type) {
      same = std::equal(elem_nodes.begin() + static_cast<long>(e.first_node), elem_nodes.begin() + static_cast<long>(e.first_node) + nn,
                        elem_nodes.begin() + static_cast<long>(run->first_node));
    }
    if (!same) run = &e;
    b.element(e.type, elem_nodes.data() + e.first_node, &e.physical, 1, same);
  }
}

// ---- MSH 4.1 (gmsh_file_v4_text.cc, gmsh_file_v4_binary.cc + gmsh_reader.cc:343-627) --------------------------------------
void parse_v4(Cursor& c, bool binary, bool swap, lfgpu_gmsh* g) {
  auto rd_int = [&]() -> long long { return binary ? c.raw<int32_t>(swap) : c.integer(); };
  auto rd_size = [&]() -> uint64_t { return binary ? c.raw<uint64_t>(swap) : static_cast<uint64_t>(c.integer()); };
  auto rd_real = [&]() -> double { return binary ? c.raw<double>(swap) : c.real(); };
  using PhysMap = std::unordered_map<long long, std::vector<uint32_t>>;
  PhysMap entities[4], part_entities[4];
  uint64_t num_partitions = 0;
  struct NodeRec { uint64_t tag; double x, y, z; };
  std::vector<NodeRec> nodes;
  struct Block { int dim; long long entity_tag; int type; size_t first, count; };
  std::vector<Block> blocks;
  std::vector<uint64_t> elem_nodes;  // per element: its node tags (element tags are not needed)

  auto read_entity = [&](int dim, bool partitioned, PhysMap* dst) {
    const long long tag = rd_int();
    if (partitioned) {
      rd_int();  // parent dimension
      rd_int();  // parent tag
      for (uint64_t k = rd_size(); k > 0; --k) rd_int();  // partitions
    }
    for (int k = 0; k < (dim == 0 ? 3 : 6); ++k) rd_real();
    std::vector<uint32_t>& phys = dst[dim][tag];
    for (uint64_t k = rd_size(); k > 0; --k) phys.push_back(static_cast<uint32_t>(rd_int()));
    if (dim > 0) {
      for (uint64_t k = rd_size(); k > 0; --k) rd_int();  // bounding entities
    }
  };

  while (!c.at_end()) {
    const std::string sec = c.token();
    if (sec == "$PhysicalNames") {
      read_physical_names(c, g);
    } else if (sec == "$Entities") {
      if (binary) c.eol();
      uint64_t cnt[4];
      for (auto& v : cnt) v = rd_size();
      for (int dim = 0; dim < 4; ++dim)
        for (uint64_t i = 0; i < cnt[dim]; ++i) read_entity(dim, false, entities);
      c.expect("$EndEntities");
    } else if (sec == "$PartitionedEntities") {
      if (binary) c.eol();
      num_partitions = rd_size();
      for (uint64_t k = rd_size(); k > 0; --k) {
        rd_int();
        rd_int();
      }
      uint64_t cnt[4];
      for (auto& v : cnt) v = rd_size();
      for (int dim = 0; dim < 4; ++dim)
        for (uint64_t i = 0; i < cnt[dim]; ++i) read_entity(dim, true, part_entities);
      c.expect("$EndPartitionedEntities");
    } else if (sec == "$Nodes") {
      if (binary) c.eol();
      const uint64_t nblocks = rd_size();
      const uint64_t total = rd_size();
      rd_size();  // min tag
      rd_size();  // max tag
      nodes.reserve(static_cast<size_t>(std::min<uint64_t>(total, c.remaining())));
      for (uint64_t bl = 0; bl < nblocks; ++bl) {
        const int dim = static_cast<int>(rd_int());
        rd_int();  // entity tag
        const bool parametric = rd_int() != 0;
        const uint64_t n = rd_size();
        const size_t first = nodes.size();
        for (uint64_t k = 0; k < n; ++k) nodes.push_back({rd_size(), 0.0, 0.0, 0.0});
        for (uint64_t k = 0; k < n; ++k) {
          NodeRec& r = nodes[first + k];
          r.x = rd_real();
          r.y = rd_real();
          r.z = rd_real();
          if (parametric)
            for (int q = 0; q < dim; ++q) rd_real();
        }
      }
      c.expect("$EndNodes");
    } else if (sec == "$Elements") {
      if (binary) c.eol();
      const uint64_t nblocks = rd_size();
      rd_size();
      rd_size();
      rd_size();
      for (uint64_t bl = 0; bl < nblocks; ++bl) {
        Block b;
        b.dim = static_cast<int>(rd_int());
        b.entity_tag = rd_int();
        b.type = static_cast<int>(rd_int());
        b.count = static_cast<size_t>(rd_size());
        b.first = elem_nodes.size();
        int nn = 0, dim = 0;
        if (!element_info(b.type, &nn, &dim)) fail("unknown element type " + std::to_string(b.type));
        if (dim != b.dim) fail("error in GmshFile: Mismatch between entity block type and dimension");
        for (size_t e = 0; e < b.count; ++e) {
          rd_size();  // element tag
          for (int k = 0; k < nn; ++k) elem_nodes.push_back(rd_size());
        }
        blocks.push_back(b);
      }
      c.expect("$EndElements");
    } else if (!sec.empty() && sec[0] == '$') {
      c.skip_section(sec);  // $Periodic, $GhostElements, ... are not used by GmshReader
    } else {
      fail("Could not parse file: unexpected '" + shown(sec) + "'");
    }
  }
  // gmsh_reader.cc:367-402: the main nodes are the vertices of the elements of dimension dim_mesh ONLY
  Builder b(g);
  std::unordered_map<uint64_t, char> is_main;
  size_t n_top = 0;
  for (const Block& bl : blocks) {
    if (bl.dim > 2) fail("mesh_factory->DimMesh() = 2, but msh-file contains entities with dimension " + std::to_string(bl.dim));
    if (bl.dim != 2) continue;
    n_top += bl.count;
    int nn = 0, dim = 0, order = 1;
    element_info(bl.type, &nn, &dim);
    const int nm = main_nodes_of(bl.type, &order);
    if (nm == 0) fail("Gmsh element type " + std::to_string(bl.type) + " not (yet) supported by GmshReader.");
    for (size_t e = 0; e < bl.count; ++e)
      for (int k = 0; k < nm; ++k) is_main[elem_nodes[bl.first + e * static_cast<size_t>(nn) + static_cast<size_t>(k)]] = 1;
  }
  if (n_top == 0) fail("MshFile contains no elements with dimension 2");
  for (const NodeRec& r : nodes) {
    if (is_main.count(r.tag)) b.add_point(r.tag, r.x, r.y, r.z);
  }
  // gmsh_reader.cc:457-540 (entities) and :546-604 (physical tags of the block's gmsh entity)
  PhysMap* ent = (num_partitions != 0) ? part_entities : entities;
  for (const Block& bl : blocks) {
    int nn = 0, dim = 0;
    element_info(bl.type, &nn, &dim);
    static const std::vector<uint32_t> kNone;
    auto it = ent[bl.dim].find(bl.entity_tag);
    const std::vector<uint32_t>& phys = (it == ent[bl.dim].end()) ? kNone : it->second;
    const uint64_t* run = nullptr;
    for (size_t e = 0; e < bl.count; ++e) {
      const uint64_t* en = elem_nodes.data() + bl.first + e * static_cast<size_t>(nn);
      const bool same = run != nullptr && std::equal(en, en + nn, run);
      if (!same) run = en;
      b.element(bl.type, en, phys.data(), phys.size(), same);
    }
  }
}

int parse_bytes(const char* data, size_t n, int dim_world, lfgpu_gmsh** out) {
  if (out == nullptr || data == nullptr) return LFGPU_ERR_INVALID;
  *out = nullptr;
  if (dim_world != 2) {
    lfgpu::set_last_error(nullptr, "lfgpu reads planar meshes only (dim_world = 2)");
    return LFGPU_ERR_UNSUPPORTED;
  }
  auto g = new lfgpu_gmsh;
  g->dim_world = dim_world;
  try {
    // header (gmsh_reader.cc:651-683): "$MeshFormat version is_binary sizeof(size_t)" [+ the int 1 in the file's byte order]
    Cursor c(data, n);
    c.expect("$MeshFormat");
    const std::string version = c.token();
    const long long is_binary = c.integer();
    const long long size_t_size = c.integer();
    bool swap = false;
    if (is_binary == 1) {
      c.eol();
      swap = c.raw<int32_t>(false) != 1;
    } else if (is_binary != 0) {
      fail("Could not read header");
    }
    c.expect("$EndMeshFormat");
    if (size_t_size != 8) fail("Size of std::size_t must be 8.");
    if (version == "4.1") parse_v4(c, is_binary == 1, swap, g);
    else if (version == "2.2") parse_v2(c, is_binary == 1, swap, g);
    else fail("GmshFiles with Version " + shown(version) + " are not yet supported.");
  } catch (const ParseError& e) {
    lfgpu::set_last_error(nullptr, "gmsh: " + e.msg);
    delete g;
    return LFGPU_ERR_INVALID;
  } catch (const std::exception& e) {
    lfgpu::set_last_error(nullptr, std::string("gmsh: ") + e.what());
    delete g;
    return LFGPU_ERR_INVALID;
  }
  *out = g;
  return LFGPU_OK;
}

const std::vector<uint32_t>* phys_of(const lfgpu_gmsh* g, int codim, int64_t index) {
  if (g == nullptr || codim < 0 || codim > 2 || index < 0) return nullptr;
  const auto& lists = g->ent_phys[codim];
  static const std::vector<uint32_t> kNone;
  return static_cast<size_t>(index) < lists.size() ? &lists[static_cast<size_t>(index)] : &kNone;
}

}  // namespace

extern "C" {

int lfgpu_gmsh_read_memory(const void* data, int64_t n_bytes, int dim_world, lfgpu_gmsh** out) {
  if (n_bytes < 0) return LFGPU_ERR_INVALID;
  return parse_bytes(static_cast<const char*>(data), static_cast<size_t>(n_bytes), dim_world, out);
}

int lfgpu_gmsh_read_file(const char* filename, int dim_world, lfgpu_gmsh** out) {
  if (filename == nullptr || out == nullptr) return LFGPU_ERR_INVALID;
  *out = nullptr;
  std::ifstream in(filename, std::ios::in | std::ios::binary);
  if (!in) {
    lfgpu::set_last_error(nullptr, std::string("Could not open file ") + filename);  // gmsh_reader.cc:633-637
    return LFGPU_ERR_INVALID;
  }
  std::string bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  return parse_bytes(bytes.data(), bytes.size(), dim_world, out);
}

void lfgpu_gmsh_destroy(lfgpu_gmsh* g) { delete g; }

int lfgpu_gmsh_counts(const lfgpu_gmsh* g, int64_t* n_nodes, int64_t* n_explicit_edges, int64_t* n_cells, int* geometry_order,
                      int* n_physical_names) {
  if (g == nullptr) return LFGPU_ERR_INVALID;
  if (n_nodes) *n_nodes = static_cast<int64_t>(g->node_xy.size() / 2);
  if (n_explicit_edges) *n_explicit_edges = static_cast<int64_t>(g->edge_nodes.size() / 2);
  if (n_cells) *n_cells = static_cast<int64_t>(g->cell_nodes.size() / 4);
  if (geometry_order) *geometry_order = g->order;
  if (n_physical_names) *n_physical_names = static_cast<int>(g->names.size());
  return LFGPU_OK;
}

int lfgpu_gmsh_arrays(const lfgpu_gmsh* g, double* node_coords, uint32_t* edge_nodes, uint32_t* cell_nodes) {
  if (g == nullptr) return LFGPU_ERR_INVALID;
  if (node_coords) std::copy(g->node_xy.begin(), g->node_xy.end(), node_coords);
  if (edge_nodes) std::copy(g->edge_nodes.begin(), g->edge_nodes.end(), edge_nodes);
  if (cell_nodes) std::copy(g->cell_nodes.begin(), g->cell_nodes.end(), cell_nodes);
  return LFGPU_OK;
}

int lfgpu_gmsh_physical_entity_nr(const lfgpu_gmsh* g, int codim, int64_t index, int capacity, uint32_t* out) {
  const std::vector<uint32_t>* p = phys_of(g, codim, index);
  if (p == nullptr) return LFGPU_ERR_INVALID;
  for (int k = 0; k < capacity && k < static_cast<int>(p->size()); ++k) out[k] = (*p)[static_cast<size_t>(k)];
  return static_cast<int>(p->size());
}

int lfgpu_gmsh_physical_flags(const lfgpu_gmsh* g, int codim, uint32_t nr, int64_t n, uint8_t* flags) {
  if (g == nullptr || codim < 0 || codim > 2 || n < 0 || (n > 0 && flags == nullptr)) return LFGPU_ERR_INVALID;
  for (int64_t i = 0; i < n; ++i) {
    const std::vector<uint32_t>* p = phys_of(g, codim, i);
    flags[i] = std::find(p->begin(), p->end(), nr) != p->end() ? 1 : 0;
  }
  return LFGPU_OK;
}

int lfgpu_gmsh_physical_name(const lfgpu_gmsh* g, int i, uint32_t* nr, int* codim, char* buf, int capacity) {
  if (g == nullptr || i < 0 || i >= static_cast<int>(g->names.size())) return LFGPU_ERR_INVALID;
  const PhysicalName& p = g->names[static_cast<size_t>(i)];
  if (nr) *nr = p.nr;
  if (codim) *codim = 2 - p.dim;
  if (buf != nullptr && capacity > 0) {
    std::strncpy(buf, p.name.c_str(), static_cast<size_t>(capacity) - 1);
    buf[capacity - 1] = '\0';
  }
  return static_cast<int>(p.name.size());
}

int lfgpu_gmsh_physical_name2nr(const lfgpu_gmsh* g, const char* name, int codim, uint32_t* nr) {
  if (g == nullptr || name == nullptr || nr == nullptr) return LFGPU_ERR_INVALID;
  int hits = 0, found = -1;
  for (size_t i = 0; i < g->names.size(); ++i) {
    if (g->names[i].name != name) continue;
    ++hits;
    if (codim < 0 ? found < 0 : (2 - g->names[i].dim == codim && found < 0)) found = static_cast<int>(i);
  }
  if (hits == 0) {
    lfgpu::set_last_error(nullptr, "No Physical Entity with this name found.");
    return LFGPU_ERR_INVALID;
  }
  if (codim < 0 && hits > 1) {
    lfgpu::set_last_error(nullptr, std::string("There are multiple physical entities with the name ") + name +
                                       ", please specify also the codimension.");
    return LFGPU_ERR_INVALID;
  }
  if (found < 0) {
    lfgpu::set_last_error(nullptr, std::string("Physical Entity with name='") + name + "' and codimension=" + std::to_string(codim) +
                                       "' not found.");
    return LFGPU_ERR_INVALID;
  }
  *nr = g->names[static_cast<size_t>(found)].nr;
  return LFGPU_OK;
}

int lfgpu_gmsh_physical_nr2name(const lfgpu_gmsh* g, uint32_t nr, int codim, char* buf, int capacity) {
  if (g == nullptr || buf == nullptr || capacity < 1) return LFGPU_ERR_INVALID;
  int hits = 0, found = -1;
  for (size_t i = 0; i < g->names.size(); ++i) {
    if (g->names[i].nr != nr) continue;
    ++hits;
    if (codim < 0 ? found < 0 : (2 - g->names[i].dim == codim && found < 0)) found = static_cast<int>(i);
  }
  if (hits == 0) {
    lfgpu::set_last_error(nullptr, "Physical entity with number " + std::to_string(nr) + " not found.");
    return LFGPU_ERR_INVALID;
  }
  if (codim < 0 && hits > 1) {
    lfgpu::set_last_error(nullptr, "There are multiple physical entities with the Number " + std::to_string(nr) +
                                       ", please specify also the codimension");
    return LFGPU_ERR_INVALID;
  }
  if (found < 0) {
    lfgpu::set_last_error(nullptr, "Physical entity with number=" + std::to_string(nr) + ", codim=" + std::to_string(codim) + " not found.");
    return LFGPU_ERR_INVALID;
  }
  const std::string& s = g->names[static_cast<size_t>(found)].name;
  std::strncpy(buf, s.c_str(), static_cast<size_t>(capacity) - 1);
  buf[capacity - 1] = '\0';
  return static_cast<int>(s.size());
}

int lfgpu_gmsh_mesh(lfgpu_ctx* ctx, const lfgpu_gmsh* g, lfgpu_mesh** out) {
  if (ctx == nullptr || g == nullptr || out == nullptr) return LFGPU_ERR_INVALID;
  *out = nullptr;
  if (g->order != 1)
    LFGPU_FAIL(ctx, LFGPU_ERR_UNSUPPORTED,
               "the file holds second-order elements (TriaO2 / QuadO2 / SegmentO2 geometries); the device path computes on "
               "TriaO1 / QuadO1 cells only");
  lfgpu_mesh* m = nullptr;
  int rc = lfgpu_mesh_upload(ctx, static_cast<int64_t>(g->node_xy.size() / 2), g->node_xy.data(),
                             static_cast<int64_t>(g->cell_nodes.size() / 4), g->cell_nodes.data(), nullptr, &m);
  if (rc != LFGPU_OK) return rc;
  rc = lfgpu_mesh_build_topology(ctx, m, static_cast<int64_t>(g->edge_nodes.size() / 2),
                                 g->edge_nodes.empty() ? nullptr : g->edge_nodes.data(), nullptr);
  if (rc != LFGPU_OK) {
    lfgpu_mesh_destroy(m);
    return rc;
  }
  *out = m;
  return LFGPU_OK;
}

}  // extern "C"
