// Test of the header-only C++ shim (include/lf_gpu_shim.hpp): the reference call sequence
//   fe_space -> provider(fe_space, alpha, gamma) -> AssembleMatrixLocally(0, dofh, dofh, provider, matrix)
// once with the oracle's COOMatrix (CPU restatement of the reference) and once with lfgpu::CsrMatrix (GPU overload).
// The oracle's mesh / DofHandler classes play the role of the LehrFEM++ types (same interface, see OracleAdaptor).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>

#include "../../include/lf_gpu_shim.hpp"
#include "../../oracle/lfo_uscalfe.h"

using namespace lfo;

struct OracleAdaptor {
  static auto entities(const mesh::Mesh& m, unsigned codim) { return m.Entities(codim); }
  static std::int64_t num_entities(const mesh::Mesh& m, unsigned codim) { return m.NumEntities(codim); }
  static unsigned index(const mesh::Mesh& m, const mesh::Entity& e) { return m.Index(e); }
  static bool is_tria(const mesh::Entity& e) { return e.RefElem() == RefEl::kTria(); }
  static auto sub_entities(const mesh::Entity& e, unsigned rel_codim) { return e.SubEntities(rel_codim); }
  static double corner(const mesh::Entity& e, int k, int d) { return e.Geometry()->Global(e.RefElem().NodeCoords())(d, k); }
  static std::int64_t num_dofs(const assemble::DofHandler& d) { return d.NumDofs(); }
  static int num_local_dofs(const assemble::DofHandler& d, const mesh::Entity& e) { return d.NumLocalDofs(e); }
  static auto global_dof_indices(const assemble::DofHandler& d, const mesh::Entity& e) { return d.GlobalDofIndices(e); }
  static const mesh::Mesh& mesh(const assemble::DofHandler& d) { return *d.Mesh(); }
};

// stand-in for FeSpaceLagrangeO<p>: what the shim's providers need from the FE space
struct FeSpace {
  std::shared_ptr<uscalfe::UniformScalarFESpace> fes;
  int degree;
  int Degree() const { return degree; }
  const assemble::DofHandler& LocGlobMap() const { return fes->LocGlobMap(); }
};

static int failures = 0;
#define CHECK(cond, ...)                     \
  do {                                       \
    if (!(cond)) {                           \
      std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
      std::printf(__VA_ARGS__);              \
      std::printf("\n");                     \
      ++failures;                            \
    }                                        \
  } while (0)

template <class OA, class OG, class GA, class GG>
void compare_matrix(lfgpu::Context& ctx, std::shared_ptr<mesh::Mesh> m, int degree, OA oalpha, OG ogamma, GA galpha, GG ggamma, const char* what) {
  auto fes = std::make_shared<uscalfe::UniformScalarFESpace>(m, degree);
  const assemble::DofHandler& dofh = fes->LocGlobMap();
  // reference path (oracle)
  uscalfe::ReactionDiffusionElementMatrixProvider<OA, OG> oprov(fes, oalpha, ogamma);
  assemble::COOMatrix coo(dofh.NumDofs(), dofh.NumDofs());
  assemble::AssembleMatrixLocally(0, dofh, dofh, oprov, coo);
  assemble::AssembleMatrixLocally(0, dofh, dofh, oprov, coo);  // accumulate twice (assembler.h:84-88)
  const auto ref = coo.makeSparse();
  // GPU path through the shim: same call shape
  auto gfes = std::make_shared<const FeSpace>(FeSpace{fes, degree});
  lfgpu::ReactionDiffusionElementMatrixProvider<double, GA, GG> gprov(gfes, galpha, ggamma);
  lfgpu::CsrMatrix M(ctx, LFGPU_COL_MAJOR);
  lfgpu::AssembleMatrixLocally<OracleAdaptor>(0, dofh, dofh, gprov, M);
  lfgpu::AssembleMatrixLocally<OracleAdaptor>(0, dofh, dofh, gprov, M);
  std::vector<std::int32_t> outer, inner;
  std::vector<double> vals;
  M.Download(outer, inner, vals);
  CHECK(outer.size() == ref.outer.size() && inner.size() == ref.inner.size(), "%s: pattern size", what);
  bool same = outer.size() == ref.outer.size() && inner.size() == ref.inner.size();
  for (std::size_t i = 0; same && i < outer.size(); ++i) same = outer[i] == ref.outer[i];
  for (std::size_t i = 0; same && i < inner.size(); ++i) same = inner[i] == ref.inner[i];
  CHECK(same, "%s: pattern differs", what);
  double scale = 0, err = 0;
  for (std::size_t i = 0; same && i < vals.size(); ++i) {
    scale = std::max(scale, std::fabs(ref.values[i]));
    err = std::max(err, std::fabs(vals[i] - ref.values[i]));
  }
  CHECK(same && err <= 1e-12 * scale, "%s: value error %.3e (scale %.3e)", what, err, scale);
  std::printf("%-46s N=%6ld nnz=%8zu rel.err=%.2e\n", what, static_cast<long>(dofh.NumDofs()), vals.size(), scale > 0 ? err / scale : 0.0);
}

// `shim_test` runs the checks that have passed on a B200; `shim_test extra` adds the ones written after the round's GPU minutes
// were spent (returning forms of the assemblers, lf::fe providers, FixSolutionComponentsLse), so that their first run cannot
// hide the others (tests/test_gpu_zz_shim_extra.py).
int main(int argc, char** argv) {
  const bool extra = argc > 1 && std::string(argv[1]) == "extra";
  try {
    lfgpu::Context ctx(0);
    auto tria = mesh::utils::TPTriagMeshBuild(24, 17, 0.0, 0.0, 2.0, 1.0);
    auto hyb = mesh::utils::HybridMeshBuild(14, 0.2, 99);
    using OC = uscalfe::MeshFunctionConstant<double>;
    using GC = lfgpu::MeshFunctionConstant<double>;
    for (int p = 1; p <= 3; ++p) {
      compare_matrix(ctx, tria, p, OC(1.0), OC(0.0), GC{1.0}, GC{0.0}, "TP-tria, Laplacian");
      compare_matrix(ctx, hyb, p, OC(2.0), OC(3.0), GC{2.0}, GC{3.0}, "hybrid, const reaction-diffusion");
      // variable coefficients through user functors (MeshFunctionGlobal): evaluated by the shim on the host
      auto fa = [](double x, double y) { return 1.0 + x * x + y * y; };
      auto fg = [](double x, double y) { return 1.0 / (1.0 + x * x + y * y); };
      uscalfe::MeshFunctionGlobal<double> oa(fa), og(fg);
      lfgpu::MeshFunctionGlobal<decltype(fa)> ga{fa};
      lfgpu::MeshFunctionGlobal<decltype(fg)> gg{fg};
      compare_matrix(ctx, hyb, p, oa, og, ga, gg, "hybrid, alpha=1+|x|^2 gamma=1/(1+|x|^2)");
      // non-symmetric tensor diffusion
      auto ft = [](double x, double y) { return uscalfe::Mat2{{{1, x}, {y, x * y + 2.0}}}; };
      auto gt = [](double x, double y) { return lfgpu::Matrix2{{{1, x}, {y, x * y + 2.0}}}; };
      uscalfe::MeshFunctionGlobal<uscalfe::Mat2> ot(ft);
      lfgpu::MeshFunctionGlobal<decltype(gt)> gtt{gt};
      compare_matrix(ctx, hyb, p, ot, OC(0.0), gtt, GC{0.0}, "hybrid, tensor alpha=[1 x; y xy+2]");
    }
    // load vector
    for (int p = 1; p <= 3; ++p) {
      auto fes = std::make_shared<uscalfe::UniformScalarFESpace>(hyb, p);
      const assemble::DofHandler& dofh = fes->LocGlobMap();
      auto f = [](double x, double y) { return std::sin(2 * M_PI * x) * std::sin(2 * M_PI * y); };
      uscalfe::ScalarLoadElementVectorProvider<uscalfe::MeshFunctionGlobal<double>> oprov(fes, uscalfe::MeshFunctionGlobal<double>(f));
      std::vector<double> ref(dofh.NumDofs(), 0.0);
      assemble::AssembleVectorLocally(0, dofh, oprov, ref);
      auto gfes = std::make_shared<const FeSpace>(FeSpace{fes, p});
      lfgpu::ScalarLoadElementVectorProvider<double, lfgpu::MeshFunctionGlobal<decltype(f)>> gprov(gfes, lfgpu::MeshFunctionGlobal<decltype(f)>{f});
      lfgpu::Vector v(ctx);
      lfgpu::AssembleVectorLocally<OracleAdaptor>(0, dofh, gprov, v);
      const auto h = v.Download();
      double scale = 0, err = 0;
      for (std::size_t i = 0; i < h.size(); ++i) {
        scale = std::max(scale, std::fabs(ref[i]));
        err = std::max(err, std::fabs(h[i] - ref[i]));
      }
      CHECK(err <= 1e-12 * scale, "load vector P%d: error %.3e", p, err);
      std::printf("load vector P%d on hybrid mesh                  N=%6zu rel.err=%.2e\n", p, h.size(), err / scale);
      if (!extra) continue;
      // the returning forms (assembler.h:243-249, 354-365): same numbers as the accumulating forms on fresh targets
      lfgpu::Vector v2 = lfgpu::AssembleVectorLocally<OracleAdaptor>(ctx, 0, dofh, gprov);
      const auto h2 = v2.Download();
      bool same_vec = h2.size() == h.size();
      for (std::size_t i = 0; same_vec && i < h.size(); ++i) same_vec = std::fabs(h2[i] - h[i]) <= 1e-13 * scale;
      CHECK(same_vec, "returning AssembleVectorLocally P%d differs", p);
      lfgpu::ReactionDiffusionElementMatrixProvider<double, GC, GC> mprov(gfes, GC{1.0}, GC{1.0});
      lfgpu::CsrMatrix M1(ctx, LFGPU_COL_MAJOR);
      lfgpu::AssembleMatrixLocally<OracleAdaptor>(0, dofh, dofh, mprov, M1);
      lfgpu::CsrMatrix M2 = lfgpu::AssembleMatrixLocally<OracleAdaptor>(ctx, 0, dofh, mprov);
      std::vector<std::int32_t> o1, i1, o2, i2;
      std::vector<double> a1, a2;
      M1.Download(o1, i1, a1);
      M2.Download(o2, i2, a2);
      CHECK(o1 == o2 && i1 == i2 && a1 == a2, "returning AssembleMatrixLocally P%d differs", p);
      // lf::fe providers (fe/loc_comp_ellbvp.h): diffusion + mass accumulated into one matrix = the reaction-diffusion matrix
      lfgpu::fe::DiffusionElementMatrixProvider<double, GC> dprov(gfes, GC{1.0});
      lfgpu::fe::MassElementMatrixProvider<double, GC> maprov(gfes, GC{1.0});
      lfgpu::CsrMatrix M3(ctx, LFGPU_COL_MAJOR);
      lfgpu::AssembleMatrixLocally<OracleAdaptor>(0, dofh, dofh, dprov, M3);
      lfgpu::AssembleMatrixLocally<OracleAdaptor>(0, dofh, dofh, maprov, M3);  // accumulates (assembler.h:84-88)
      std::vector<std::int32_t> o3, i3;
      std::vector<double> a3;
      M3.Download(o3, i3, a3);
      double s3 = 0, e3 = 0;
      for (std::size_t k = 0; k < a1.size() && k < a3.size(); ++k) {
        s3 = std::max(s3, std::fabs(a1[k]));
        e3 = std::max(e3, std::fabs(a1[k] - a3[k]));
      }
      CHECK(o1 == o3 && i1 == i3 && e3 <= 1e-13 * s3, "lf::fe diffusion + mass P%d: error %.3e", p, e3);
    }
    // Dirichlet elimination (fix_dof.h:86-138,181-218): assemble A, b, fix every third dof, compare operator and rhs
    for (int variant = 0; variant < (extra ? 3 : 2); ++variant) {  // 2 = FixSolutionComponentsLse: (index, value) pairs, repeated indices add up
      const int p = 2;
      auto fes = std::make_shared<uscalfe::UniformScalarFESpace>(hyb, p);
      const assemble::DofHandler& dofh = fes->LocGlobMap();
      const std::size_t n = dofh.NumDofs();
      auto sel = [](std::int64_t i) { return std::make_pair(i % 3 == 1, 0.25 * static_cast<double>(i % 7) - 0.5); };
      uscalfe::ReactionDiffusionElementMatrixProvider<OC, OC> oprov(fes, OC(1.0), OC(2.0));
      assemble::COOMatrix coo(n, n);
      assemble::AssembleMatrixLocally(0, dofh, dofh, oprov, coo);
      uscalfe::ScalarLoadElementVectorProvider<OC> olprov(fes, OC(3.0));
      std::vector<double> ref_b(n, 0.0);
      assemble::AssembleVectorLocally(0, dofh, olprov, ref_b);
      std::vector<std::pair<std::int64_t, double>> pairs;
      for (std::size_t i = 0; i < n; ++i)
        if (sel(static_cast<std::int64_t>(i)).first) pairs.emplace_back(static_cast<std::int64_t>(i), sel(static_cast<std::int64_t>(i)).second);
      pairs.emplace_back(1, 0.125);  // dof 1 is fixed already: its value becomes the sum (fix_dof.h:268)
      if (variant == 0) {
        assemble::FixFlaggedSolutionComponents(sel, coo, ref_b);
      } else if (variant == 1) {
        assemble::FixFlaggedSolutionCompAlt(sel, coo, ref_b);
      } else {
        assemble::FixSolutionComponentsLse(pairs, coo, ref_b);
      }
      const auto ref = coo.makeSparse();
      auto gfes = std::make_shared<const FeSpace>(FeSpace{fes, p});
      lfgpu::ReactionDiffusionElementMatrixProvider<double, GC, GC> gprov(gfes, GC{1.0}, GC{2.0});
      lfgpu::ScalarLoadElementVectorProvider<double, GC> glprov(gfes, GC{3.0});
      lfgpu::CsrMatrix M(ctx, LFGPU_COL_MAJOR);
      lfgpu::Vector v(ctx);
      lfgpu::AssembleMatrixLocally<OracleAdaptor>(0, dofh, dofh, gprov, M);
      lfgpu::AssembleVectorLocally<OracleAdaptor>(0, dofh, glprov, v);
      if (variant == 0) {
        lfgpu::FixFlaggedSolutionComponents<double>(sel, M, v);
      } else if (variant == 1) {
        lfgpu::FixFlaggedSolutionCompAlt<double>(sel, M, v);
      } else {
        lfgpu::FixSolutionComponentsLse<double>(pairs, M, v);
      }
      std::vector<std::int32_t> outer, inner;
      std::vector<double> vals;
      M.Download(outer, inner, vals);
      // the GPU matrix keeps explicit zeros where the reference erased triplets: compare entry by entry over the union
      double scale = 0, err = 0;
      std::size_t kept = 0;
      for (std::size_t c = 0; c < n; ++c) {
        std::int32_t kr = ref.outer[c];
        for (std::int32_t k = outer[c]; k < outer[c + 1]; ++k) {
          double r = 0.0;
          if (kr < ref.outer[c + 1] && ref.inner[kr] == inner[k]) r = ref.values[kr++];
          if (vals[k] != 0.0) ++kept;
          scale = std::max(scale, std::fabs(r));
          err = std::max(err, std::fabs(vals[k] - r));
        }
        CHECK(kr == ref.outer[c + 1], "fix variant %d: reference entry outside the GPU pattern in column %zu", variant, c);
      }
      const auto hb = v.Download();
      double bscale = 0, berr = 0;
      for (std::size_t i = 0; i < n; ++i) {
        bscale = std::max(bscale, std::fabs(ref_b[i]));
        berr = std::max(berr, std::fabs(hb[i] - ref_b[i]));
      }
      CHECK(err <= 1e-12 * scale && berr <= 1e-12 * bscale, "fix variant %d: matrix err %.3e rhs err %.3e", variant, err, berr);
      if (variant == 0) {  // the symmetric variant keeps the system SPD: solve on the device, fixed components must come out
        int iters = 0;
        double res = 1.0;
        const auto x = lfgpu::SolveCG(M, v, 1e-12, 5000, &iters, &res);
        double fix_err = 0;
        for (std::size_t i = 0; i < n; ++i)
          if (sel(static_cast<std::int64_t>(i)).first) fix_err = std::max(fix_err, std::fabs(x[i] - sel(static_cast<std::int64_t>(i)).second));
        CHECK(res <= 1e-12 && fix_err <= 1e-11, "device CG: residual %.3e after %d iterations, fixed components off by %.3e", res, iters, fix_err);
        std::printf("%-46s iterations=%d rel.residual=%.2e fixed-dof error=%.2e\n", "SolveCG after FixFlaggedSolutionComponents", iters, res, fix_err);
      }
      std::printf("%-46s N=%6zu nnz=%8zu rel.err=%.2e rhs=%.2e\n", variant == 0 ? "FixFlaggedSolutionComponents P2 hybrid" : (variant == 1 ? "FixFlaggedSolutionCompAlt P2 hybrid" : "FixSolutionComponentsLse P2 hybrid"), n,
                  kept, err / scale, berr / bscale);
    }
    // impedance boundary terms (sec_ord_ell_bvp.h:147-215): cell matrix + edge mass on the part {x = 0} u {y = 0} of the
    // boundary, load vector + edge load; same call sequence on both sides, same pattern (edge entries live in cell entries)
    for (int p = 1; p <= 3; ++p) {
      auto fes = std::make_shared<uscalfe::UniformScalarFESpace>(hyb, p);
      const assemble::DofHandler& dofh = fes->LocGlobMap();
      const std::size_t n = dofh.NumDofs();
      auto sel = [](const mesh::Entity& e) {
        const Mat c = e.Geometry()->Global(e.RefElem().NodeCoords());
        return (c(0, 0) < 1e-12 && c(0, 1) < 1e-12) || (c(1, 0) < 1e-12 && c(1, 1) < 1e-12);
      };
      auto feta = [](double x, double y) { return 1.0 + x * x + y * y; };
      auto fg = [](double x, double y) { return std::cos(3.0 * x) + y; };
      // reference side
      uscalfe::ReactionDiffusionElementMatrixProvider<OC, OC> oprov(fes, OC(1.0), OC(0.5));
      uscalfe::MassEdgeMatrixProvider<uscalfe::MeshFunctionGlobal<double>, decltype(sel)> oeprov(fes, uscalfe::MeshFunctionGlobal<double>(feta), sel);
      assemble::COOMatrix coo(n, n);
      assemble::AssembleMatrixLocally(0, dofh, dofh, oprov, coo);
      assemble::AssembleMatrixLocally(1, dofh, dofh, oeprov, coo);
      const auto ref = coo.makeSparse();
      uscalfe::ScalarLoadElementVectorProvider<OC> olprov(fes, OC(1.0));
      uscalfe::ScalarLoadEdgeVectorProvider<uscalfe::MeshFunctionGlobal<double>, decltype(sel)> oelprov(fes, uscalfe::MeshFunctionGlobal<double>(fg), sel);
      std::vector<double> ref_b(n, 0.0);
      assemble::AssembleVectorLocally(0, dofh, olprov, ref_b);
      assemble::AssembleVectorLocally(1, dofh, oelprov, ref_b);
      // GPU side
      auto gfes = std::make_shared<const FeSpace>(FeSpace{fes, p});
      lfgpu::ReactionDiffusionElementMatrixProvider<double, GC, GC> gprov(gfes, GC{1.0}, GC{0.5});
      lfgpu::MassEdgeMatrixProvider<double, lfgpu::MeshFunctionGlobal<decltype(feta)>, decltype(sel)> geprov(gfes, lfgpu::MeshFunctionGlobal<decltype(feta)>{feta}, sel);
      lfgpu::ScalarLoadElementVectorProvider<double, GC> glprov(gfes, GC{1.0});
      lfgpu::ScalarLoadEdgeVectorProvider<double, lfgpu::MeshFunctionGlobal<decltype(fg)>, decltype(sel)> gelprov(gfes, lfgpu::MeshFunctionGlobal<decltype(fg)>{fg}, sel);
      lfgpu::CsrMatrix M(ctx, LFGPU_COL_MAJOR);
      lfgpu::Vector v(ctx);
      lfgpu::AssembleMatrixLocally<OracleAdaptor>(0, dofh, dofh, gprov, M);
      lfgpu::AssembleMatrixLocally<OracleAdaptor>(1, dofh, dofh, geprov, M);
      lfgpu::AssembleVectorLocally<OracleAdaptor>(0, dofh, glprov, v);
      lfgpu::AssembleVectorLocally<OracleAdaptor>(1, dofh, gelprov, v);
      std::vector<std::int32_t> outer, inner;
      std::vector<double> vals;
      M.Download(outer, inner, vals);
      bool same = outer.size() == ref.outer.size() && inner.size() == ref.inner.size();
      for (std::size_t i = 0; same && i < outer.size(); ++i) same = outer[i] == ref.outer[i];
      for (std::size_t i = 0; same && i < inner.size(); ++i) same = inner[i] == ref.inner[i];
      double scale = 0, err = 0;
      for (std::size_t i = 0; same && i < vals.size(); ++i) {
        scale = std::max(scale, std::fabs(ref.values[i]));
        err = std::max(err, std::fabs(vals[i] - ref.values[i]));
      }
      const auto hb = v.Download();
      double bscale = 0, berr = 0;
      for (std::size_t i = 0; i < n; ++i) {
        bscale = std::max(bscale, std::fabs(ref_b[i]));
        berr = std::max(berr, std::fabs(hb[i] - ref_b[i]));
      }
      CHECK(same && err <= 1e-12 * scale && berr <= 1e-12 * bscale, "impedance terms P%d: same=%d matrix err %.3e rhs err %.3e", p, static_cast<int>(same), err, berr);
      std::printf("cell + edge (impedance) terms P%d on hybrid mesh     N=%6zu nnz=%8zu rel.err=%.2e rhs=%.2e\n", p, n, vals.size(), err / scale, berr / bscale);
    }
    // ---- round 2: boundary fixes of the shim ----------------------------------------------------------------------------
    if (extra) {
      using OC2 = uscalfe::MeshFunctionConstant<double>;
      for (int p = 1; p <= 3; ++p) {
        auto fes = std::make_shared<uscalfe::UniformScalarFESpace>(hyb, p);
        const assemble::DofHandler& dofh = fes->LocGlobMap();
        auto gfes = std::make_shared<const FeSpace>(FeSpace{fes, p});
        // (1) isActive of a cell provider (assembler.h:127): a derived provider that switches every third cell off
        struct OSel : uscalfe::ReactionDiffusionElementMatrixProvider<OC2, OC2> {
          using Base = uscalfe::ReactionDiffusionElementMatrixProvider<OC2, OC2>;
          const mesh::Mesh* m;
          OSel(std::shared_ptr<uscalfe::UniformScalarFESpace> f, const mesh::Mesh* mm) : Base(f, OC2(2.0), OC2(3.0)), m(mm) {}
          bool isActive(const mesh::Entity& c) override { return m->Index(c) % 3 != 0; }
        };
        struct GSel : lfgpu::ReactionDiffusionElementMatrixProvider<double, GC, GC> {
          using Base = lfgpu::ReactionDiffusionElementMatrixProvider<double, GC, GC>;
          const mesh::Mesh* m;
          GSel(std::shared_ptr<const FeSpace> f, const mesh::Mesh* mm) : Base(f, GC{2.0}, GC{3.0}), m(mm) {}
          bool isActive(const mesh::Entity& c) { return m->Index(c) % 3 != 0; }
        };
        OSel osel(fes, hyb.get());
        assemble::COOMatrix coo(dofh.NumDofs(), dofh.NumDofs());
        assemble::AssembleMatrixLocally(0, dofh, dofh, osel, coo);
        GSel gsel(gfes, hyb.get());
        lfgpu::CsrMatrix M(ctx, LFGPU_COL_MAJOR);
        lfgpu::AssembleMatrixLocally<OracleAdaptor>(0, dofh, dofh, gsel, M);
        std::vector<std::int32_t> outer, inner;
        std::vector<double> vals;
        M.Download(outer, inner, vals);
        // the GPU pattern is the full one (inactive cells leave explicit zeros); compare as operators
        const auto ref = coo.makeSparse();
        double scale = 0, err = 0;
        for (std::size_t c = 0; c + 1 < outer.size(); ++c) {
          std::int32_t kr = ref.outer[c];
          for (std::int32_t k = outer[c]; k < outer[c + 1]; ++k) {
            double rv = 0.0;
            if (kr < ref.outer[c + 1] && ref.inner[kr] == inner[k]) rv = ref.values[kr++];
            scale = std::max(scale, std::fabs(rv));
            err = std::max(err, std::fabs(vals[k] - rv));
          }
          CHECK(kr == ref.outer[c + 1], "isActive P%d: reference entry outside the GPU pattern in column %zu", p, c);
        }
        CHECK(err <= 1e-12 * scale, "isActive P%d: error %.3e", p, err);
        std::printf("isActive (every third cell off) P%d             rel.err=%.2e\n", p, err / scale);
        // (2) MeshFunctionGlobal with a functor of ONE point argument (mesh_function_global.h:77-88)
        auto f1 = [](auto x) { return 1.0 + x[0] * x[0] + x[1]; };
        auto f2 = [](double x, double y) { return 1.0 + x * x + y; };
        uscalfe::MeshFunctionGlobal<double> o2(f2);
        lfgpu::MeshFunctionGlobal<decltype(f1)> g1{f1};
        compare_matrix(ctx, hyb, p, o2, OC2(0.5), g1, GC{0.5}, "hybrid, one-argument functor");
        // (3) load provider with a rule collection (loc_comp_ellbvp.h:660-686)
        {
          std::map<RefEl, quad::QuadRule> qrs{{RefEl::kTria(), quad::make_QuadRule(RefEl::kTria(), 2 * p + 2)},
                                              {RefEl::kQuad(), quad::make_QuadRule(RefEl::kQuad(), 2 * p + 2)}};
          uscalfe::ScalarLoadElementVectorProvider<uscalfe::MeshFunctionGlobal<double>> oprov(fes, uscalfe::MeshFunctionGlobal<double>(f2), qrs);
          std::vector<double> refv(dofh.NumDofs(), 0.0);
          assemble::AssembleVectorLocally(0, dofh, oprov, refv);
          lfgpu::ScalarLoadElementVectorProvider<double, lfgpu::MeshFunctionGlobal<decltype(f2)>> gprov(
              gfes, lfgpu::MeshFunctionGlobal<decltype(f2)>{f2}, qrs);
          lfgpu::Vector v(ctx);
          lfgpu::AssembleVectorLocally<OracleAdaptor>(0, dofh, gprov, v);
          const auto h = v.Download();
          double s2 = 0, e2 = 0;
          for (std::size_t i = 0; i < h.size(); ++i) {
            s2 = std::max(s2, std::fabs(refv[i]));
            e2 = std::max(e2, std::fabs(h[i] - refv[i]));
          }
          CHECK(e2 <= 1e-12 * s2, "load vector with a rule collection P%d: error %.3e", p, e2);
          std::printf("load vector, rule collection of degree %d P%d      rel.err=%.2e\n", 2 * p + 2, p, e2 / s2);
        }
        // (4) several devices from one process: the same device listed three times gets three contexts and three sub-problems
        {
          uscalfe::ReactionDiffusionElementMatrixProvider<uscalfe::MeshFunctionGlobal<double>, OC2> oprov(fes, o2, OC2(0.5));
          assemble::COOMatrix coo2(dofh.NumDofs(), dofh.NumDofs());
          assemble::AssembleMatrixLocally(0, dofh, dofh, oprov, coo2);
          assemble::AssembleMatrixLocally(0, dofh, dofh, oprov, coo2);
          const auto ref2 = coo2.makeSparse();
          lfgpu::ReactionDiffusionElementMatrixProvider<double, lfgpu::MeshFunctionGlobal<decltype(f2)>, GC> gprov(
              gfes, lfgpu::MeshFunctionGlobal<decltype(f2)>{f2}, GC{0.5});
          lfgpu::MultiCsrMatrix MM({0, 0, 0}, LFGPU_COL_MAJOR);
          lfgpu::AssembleMatrixLocally<OracleAdaptor>(0, dofh, dofh, gprov, MM);
          lfgpu::AssembleMatrixLocally<OracleAdaptor>(0, dofh, dofh, gprov, MM);  // accumulates
          std::vector<std::int64_t> mo;
          std::vector<std::int32_t> mi;
          std::vector<double> mv;
          MM.Gather(mo, mi, mv);
          bool same = mo.size() == ref2.outer.size() && mi.size() == ref2.inner.size();
          for (std::size_t i = 0; same && i < mo.size(); ++i) same = mo[i] == ref2.outer[i];
          for (std::size_t i = 0; same && i < mi.size(); ++i) same = mi[i] == ref2.inner[i];
          double s4 = 0, e4 = 0;
          for (std::size_t i = 0; same && i < mv.size(); ++i) {
            s4 = std::max(s4, std::fabs(ref2.values[i]));
            e4 = std::max(e4, std::fabs(mv[i] - ref2.values[i]));
          }
          CHECK(same && e4 <= 1e-12 * s4, "MultiCsrMatrix P%d: same=%d error %.3e", p, static_cast<int>(same), e4);
          std::printf("MultiCsrMatrix on devices {0,0,0} P%d             nnz=%8zu rel.err=%.2e\n", p, mv.size(), s4 > 0 ? e4 / s4 : 0.0);
        }
      }
    }
    // missing rule -> error (loc_comp_ellbvp.h:278-287)
  } catch (const lfgpu::Error& e) {
    std::printf("lfgpu::Error %d: %s\n", e.code, e.what());
    return e.code == LFGPU_ERR_NO_DEVICE ? 77 : 2;
  }
  std::printf(failures == 0 ? "SHIM_TEST_OK\n" : "SHIM_TEST_FAILED\n");
  return failures == 0 ? 0 : 1;
}
