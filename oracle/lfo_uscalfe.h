// ORACLE (test infrastructure, NOT product code) -- see lfo_base.h header.
// lib/lf/uscalfe: lagr_fe.h:56-1480, uniform_scalar_fe_space.h:50-342, fe_space_lagrange_o{1,2,3}.h,
// precomputed_scalar_reference_finite_element.h:44-185, loc_comp_ellbvp.h:85-339,562-746;
// lib/lf/mesh/utils/mesh_function_{constant,global}.h; lib/lf/fe/fe_tools.h:198-258 (NodalProjection)
#ifndef LFO_USCALFE_H
#define LFO_USCALFE_H

#include <functional>

#include "lfo_assemble.h"
#include "lfo_quad.h"

namespace lfo::uscalfe {

// lib/lf/fe/scalar_reference_finite_element.h:82-371 (members used on the path)
class ScalarReferenceFiniteElement {
 public:
  virtual ~ScalarReferenceFiniteElement() = default;
  [[nodiscard]] virtual RefEl RefElem() const = 0;
  [[nodiscard]] virtual unsigned Degree() const = 0;
  [[nodiscard]] virtual size_type NumRefShapeFunctions() const = 0;
  [[nodiscard]] virtual size_type NumRefShapeFunctions(dim_t codim) const = 0;  // interior dofs per sub-entity
  [[nodiscard]] virtual Mat EvalReferenceShapeFunctions(const Mat& refcoords) const = 0;       // nsf x n
  [[nodiscard]] virtual Mat GradientsReferenceShapeFunctions(const Mat& refcoords) const = 0;  // nsf x (dim*n)
  [[nodiscard]] virtual Mat EvaluationNodes() const = 0;
};

// helper: gradient storage trick of lagr_fe.h:241-251 -- an (nsf x 2n) column-major matrix viewed as (2 nsf x n):
// temp(i, k) = d/dx0 of sf i at point k -> result(i, 2k); temp(i + nsf, k) = d/dx1 -> result(i, 2k + 1)
struct GradView {
  Mat& m;
  long nsf;
  double& dx0(long i, long k) { return m(i, 2 * k); }
  double& dx1(long i, long k) { return m(i, 2 * k + 1); }
};

// ---- order 1 ----------------------------------------------------------------------------------------------------
// lagr_fe.h:56-145
class FeLagrangeO1Tria final : public ScalarReferenceFiniteElement {
 public:
  [[nodiscard]] RefEl RefElem() const override { return RefEl::kTria(); }
  [[nodiscard]] unsigned Degree() const override { return 1; }
  [[nodiscard]] size_type NumRefShapeFunctions() const override { return 3; }
  [[nodiscard]] size_type NumRefShapeFunctions(dim_t codim) const override { return codim == 2 ? 1 : 0; }
  [[nodiscard]] Mat EvalReferenceShapeFunctions(const Mat& x) const override {
    Mat r(3, x.cols());
    for (long k = 0; k < x.cols(); ++k) {
      r(0, k) = 1.0 - x(0, k) - x(1, k);
      r(1, k) = x(0, k);
      r(2, k) = x(1, k);
    }
    return r;
  }
  [[nodiscard]] Mat GradientsReferenceShapeFunctions(const Mat& x) const override {
    Mat r(3, 2 * x.cols());
    for (long k = 0; k < x.cols(); ++k) {
      r(0, 2 * k) = -1; r(0, 2 * k + 1) = -1;
      r(1, 2 * k) = 1;  r(1, 2 * k + 1) = 0;
      r(2, 2 * k) = 0;  r(2, 2 * k + 1) = 1;
    }
    return r;
  }
  [[nodiscard]] Mat EvaluationNodes() const override { return RefEl::kTria().NodeCoords(); }
};

// lagr_fe.h:164-263
class FeLagrangeO1Quad final : public ScalarReferenceFiniteElement {
 public:
  [[nodiscard]] RefEl RefElem() const override { return RefEl::kQuad(); }
  [[nodiscard]] unsigned Degree() const override { return 1; }
  [[nodiscard]] size_type NumRefShapeFunctions() const override { return 4; }
  [[nodiscard]] size_type NumRefShapeFunctions(dim_t codim) const override { return codim == 2 ? 1 : 0; }
  [[nodiscard]] Mat EvalReferenceShapeFunctions(const Mat& x) const override {
    Mat r(4, x.cols());
    for (long k = 0; k < x.cols(); ++k) {
      const double x0 = x(0, k), x1 = x(1, k);
      r(0, k) = (1 - x0) * (1 - x1);
      r(1, k) = x0 * (1 - x1);
      r(2, k) = x0 * x1;
      r(3, k) = (1 - x0) * x1;
    }
    return r;
  }
  [[nodiscard]] Mat GradientsReferenceShapeFunctions(const Mat& x) const override {
    Mat r(4, 2 * x.cols());
    GradView t{r, 4};
    for (long k = 0; k < x.cols(); ++k) {
      const double x0 = x(0, k), x1 = x(1, k);
      t.dx0(0, k) = x1 - 1.0; t.dx0(1, k) = 1.0 - x1; t.dx0(2, k) = x1;  t.dx0(3, k) = -x1;
      t.dx1(0, k) = x0 - 1.0; t.dx1(1, k) = -x0;      t.dx1(2, k) = x0;  t.dx1(3, k) = 1.0 - x0;
    }
    return r;
  }
  [[nodiscard]] Mat EvaluationNodes() const override { return RefEl::kQuad().NodeCoords(); }
};

// lagr_fe.h:280-380
class FeLagrangeO1Segment final : public ScalarReferenceFiniteElement {
 public:
  [[nodiscard]] RefEl RefElem() const override { return RefEl::kSegment(); }
  [[nodiscard]] unsigned Degree() const override { return 1; }
  [[nodiscard]] size_type NumRefShapeFunctions() const override { return 2; }
  [[nodiscard]] size_type NumRefShapeFunctions(dim_t codim) const override { return codim == 1 ? 1 : 0; }
  [[nodiscard]] Mat EvalReferenceShapeFunctions(const Mat& x) const override {
    Mat r(2, x.cols());
    for (long k = 0; k < x.cols(); ++k) {
      r(0, k) = 1.0 - x(0, k);
      r(1, k) = x(0, k);
    }
    return r;
  }
  [[nodiscard]] Mat GradientsReferenceShapeFunctions(const Mat& x) const override {
    Mat r(2, x.cols());
    for (long k = 0; k < x.cols(); ++k) {
      r(0, k) = -1;
      r(1, k) = 1;
    }
    return r;
  }
  [[nodiscard]] Mat EvaluationNodes() const override { return RefEl::kSegment().NodeCoords(); }
};

// ---- order 2 ----------------------------------------------------------------------------------------------------
// lagr_fe.h:562-686
class FeLagrangeO2Segment final : public ScalarReferenceFiniteElement {
 public:
  [[nodiscard]] RefEl RefElem() const override { return RefEl::kSegment(); }
  [[nodiscard]] unsigned Degree() const override { return 2; }
  [[nodiscard]] size_type NumRefShapeFunctions() const override { return 3; }
  [[nodiscard]] size_type NumRefShapeFunctions(dim_t) const override { return 1; }
  [[nodiscard]] Mat EvalReferenceShapeFunctions(const Mat& xx) const override {
    Mat r(3, xx.cols());
    for (long k = 0; k < xx.cols(); ++k) {
      const double x = xx(0, k);
      r(0, k) = 2.0 * (1.0 - x) * (0.5 - x);
      r(1, k) = 2.0 * x * (x - 0.5);
      r(2, k) = 4.0 * (1.0 - x) * x;
    }
    return r;
  }
  [[nodiscard]] Mat GradientsReferenceShapeFunctions(const Mat& xx) const override {
    Mat r(3, xx.cols());
    for (long k = 0; k < xx.cols(); ++k) {
      const double x = xx(0, k);
      r(0, k) = 4.0 * x - 3.0;
      r(1, k) = 4.0 * x - 1.0;
      r(2, k) = 4.0 - 8.0 * x;
    }
    return r;
  }
  [[nodiscard]] Mat EvaluationNodes() const override {
    Mat n(1, 3);
    n(0, 0) = 0.0; n(0, 1) = 1.0; n(0, 2) = 0.5;
    return n;
  }
};

// lagr_fe.h:399-546
class FeLagrangeO2Tria final : public ScalarReferenceFiniteElement {
 public:
  [[nodiscard]] RefEl RefElem() const override { return RefEl::kTria(); }
  [[nodiscard]] unsigned Degree() const override { return 2; }
  [[nodiscard]] size_type NumRefShapeFunctions() const override { return 6; }
  [[nodiscard]] size_type NumRefShapeFunctions(dim_t codim) const override { return codim == 0 ? 0 : 1; }
  [[nodiscard]] Mat EvalReferenceShapeFunctions(const Mat& x) const override {
    Mat r(6, x.cols());
    for (long k = 0; k < x.cols(); ++k) {
      const double x0 = x(0, k), x1 = x(1, k);
      r(0, k) = 2.0 * (1 - x0 - x1) * (0.5 - x0 - x1);
      r(1, k) = 2.0 * x0 * (x0 - 0.5);
      r(2, k) = 2.0 * x1 * (x1 - 0.5);
      r(3, k) = 4.0 * (1 - x0 - x1) * x0;
      r(4, k) = 4.0 * x0 * x1;
      r(5, k) = 4.0 * (1 - x0 - x1) * x1;
    }
    return r;
  }
  [[nodiscard]] Mat GradientsReferenceShapeFunctions(const Mat& x) const override {
    Mat r(6, 2 * x.cols());
    GradView t{r, 6};
    for (long k = 0; k < x.cols(); ++k) {
      const double x0 = x(0, k), x1 = x(1, k);
      t.dx0(0, k) = -3.0 + 4.0 * x0 + 4.0 * x1;
      t.dx0(1, k) = 4.0 * x0 - 1.0;
      t.dx0(2, k) = 0.0;
      t.dx0(3, k) = 4 - 0 - 8.0 * x0 - 4.0 * x1;
      t.dx0(4, k) = 4.0 * x1;
      t.dx0(5, k) = -4.0 * x1;
      t.dx1(0, k) = -3.0 + 4.0 * x0 + 4.0 * x1;
      t.dx1(1, k) = 0.0;
      t.dx1(2, k) = 4.0 * x1 - 1.0;
      t.dx1(3, k) = -4 * x0;
      t.dx1(4, k) = 4.0 * x0;
      t.dx1(5, k) = 4.0 - 8.0 * x1 - 4.0 * x0;
    }
    return r;
  }
  [[nodiscard]] Mat EvaluationNodes() const override {
    static const double n[2][6] = {{0.0, 1.0, 0.0, 0.5, 0.5, 0.0}, {0.0, 0.0, 1.0, 0.0, 0.5, 0.5}};
    Mat m(2, 6);
    for (int k = 0; k < 6; ++k) {
      m(0, k) = n[0][k];
      m(1, k) = n[1][k];
    }
    return m;
  }
};

// Tensor-product quad elements: lagr_fe.h:712-925 (O2, map :913-924) and :1278-1480 (O3, map :1461-1479)
template <class SEGMENT_FE, int NSF>
class FeLagrangeTPQuad : public ScalarReferenceFiniteElement {
 public:
  explicit FeLagrangeTPQuad(const int (*map)[2]) : map_(map) {}
  [[nodiscard]] RefEl RefElem() const override { return RefEl::kQuad(); }
  [[nodiscard]] size_type NumRefShapeFunctions() const override { return NSF; }
  [[nodiscard]] Mat EvalReferenceShapeFunctions(const Mat& x) const override {
    const long n = x.cols();
    Mat r0(1, n), r1(1, n);
    for (long k = 0; k < n; ++k) {
      r0(0, k) = x(0, k);
      r1(0, k) = x(1, k);
    }
    const Mat s0 = seg_.EvalReferenceShapeFunctions(r0), s1 = seg_.EvalReferenceShapeFunctions(r1);
    Mat r(NSF, n);
    for (int i = 0; i < NSF; ++i) {
      for (long k = 0; k < n; ++k) r(i, k) = s0(map_[i][0], k) * s1(map_[i][1], k);
    }
    return r;
  }
  [[nodiscard]] Mat GradientsReferenceShapeFunctions(const Mat& x) const override {
    const long n = x.cols();
    Mat r0(1, n), r1(1, n);
    for (long k = 0; k < n; ++k) {
      r0(0, k) = x(0, k);
      r1(0, k) = x(1, k);
    }
    const Mat s0 = seg_.EvalReferenceShapeFunctions(r0), s1 = seg_.EvalReferenceShapeFunctions(r1);
    const Mat g0 = seg_.GradientsReferenceShapeFunctions(r0), g1 = seg_.GradientsReferenceShapeFunctions(r1);
    Mat r(NSF, 2 * n);
    GradView t{r, NSF};
    for (int i = 0; i < NSF; ++i) {
      for (long k = 0; k < n; ++k) {
        t.dx0(i, k) = g0(map_[i][0], k) * s1(map_[i][1], k);
        t.dx1(i, k) = g1(map_[i][1], k) * s0(map_[i][0], k);
      }
    }
    return r;
  }

 private:
  SEGMENT_FE seg_;
  const int (*map_)[2];
};

inline constexpr int kO2QuadMap[9][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}, {2, 0}, {1, 2}, {2, 1}, {0, 2}, {2, 2}};
class FeLagrangeO2Quad final : public FeLagrangeTPQuad<FeLagrangeO2Segment, 9> {
 public:
  FeLagrangeO2Quad() : FeLagrangeTPQuad(kO2QuadMap) {}
  [[nodiscard]] unsigned Degree() const override { return 2; }
  [[nodiscard]] size_type NumRefShapeFunctions(dim_t) const override { return 1; }
  using FeLagrangeTPQuad::NumRefShapeFunctions;
  [[nodiscard]] Mat EvaluationNodes() const override {
    static const double n[2][9] = {{0.0, 1.0, 1.0, 0.0, 0.5, 1.0, 0.5, 0.0, 0.5}, {0.0, 0.0, 1.0, 1.0, 0.0, 0.5, 1.0, 0.5, 0.5}};
    Mat m(2, 9);
    for (int k = 0; k < 9; ++k) {
      m(0, k) = n[0][k];
      m(1, k) = n[1][k];
    }
    return m;
  }
};

// ---- order 3 ----------------------------------------------------------------------------------------------------
// lagr_fe.h:1154-1252
class FeLagrangeO3Segment final : public ScalarReferenceFiniteElement {
 public:
  [[nodiscard]] RefEl RefElem() const override { return RefEl::kSegment(); }
  [[nodiscard]] unsigned Degree() const override { return 3; }
  [[nodiscard]] size_type NumRefShapeFunctions() const override { return 4; }
  [[nodiscard]] size_type NumRefShapeFunctions(dim_t codim) const override { return codim == 0 ? 2 : 1; }
  [[nodiscard]] Mat EvalReferenceShapeFunctions(const Mat& xx) const override {
    Mat r(4, xx.cols());
    for (long k = 0; k < xx.cols(); ++k) {
      const double x = xx(0, k);
      r(0, k) = 4.5 * (1.0 - x) * (1.0 / 3.0 - x) * (2.0 / 3.0 - x);
      r(1, k) = 4.5 * x * (x - 1.0 / 3.0) * (x - 2.0 / 3.0);
      r(2, k) = 13.5 * x * (1 - x) * (2.0 / 3.0 - x);
      r(3, k) = 13.5 * x * (1 - x) * (x - 1.0 / 3.0);
    }
    return r;
  }
  [[nodiscard]] Mat GradientsReferenceShapeFunctions(const Mat& xx) const override {
    Mat r(4, xx.cols());
    for (long k = 0; k < xx.cols(); ++k) {
      const double x = xx(0, k);
      r(0, k) = -13.5 * x * x + 18.0 * x - 5.5;
      r(1, k) = 13.5 * x * x - 9.0 * x + 1.0;
      r(2, k) = 40.5 * x * x - 45 * x + 9;
      r(3, k) = -40.5 * x * x + 36 * x - 4.5;
    }
    return r;
  }
  [[nodiscard]] Mat EvaluationNodes() const override {
    Mat n(1, 4);
    n(0, 0) = 0.0; n(0, 1) = 1.0; n(0, 2) = 1.0 / 3.0; n(0, 3) = 2.0 / 3.0;
    return n;
  }
};

// lagr_fe.h:944-1140
class FeLagrangeO3Tria final : public ScalarReferenceFiniteElement {
 public:
  [[nodiscard]] RefEl RefElem() const override { return RefEl::kTria(); }
  [[nodiscard]] unsigned Degree() const override { return 3; }
  [[nodiscard]] size_type NumRefShapeFunctions() const override { return 10; }
  [[nodiscard]] size_type NumRefShapeFunctions(dim_t codim) const override { return codim == 1 ? 2 : 1; }
  [[nodiscard]] Mat EvalReferenceShapeFunctions(const Mat& x) const override {
    Mat r(10, x.cols());
    for (long k = 0; k < x.cols(); ++k) {
      const double lambda0 = 1 - x(0, k) - x(1, k), lambda1 = x(0, k), lambda2 = x(1, k);
      r(0, k) = 4.5 * lambda0 * (lambda0 - 1 / 3.0) * (lambda0 - 2 / 3.0);
      r(1, k) = 4.5 * lambda1 * (lambda1 - 1 / 3.0) * (lambda1 - 2 / 3.0);
      r(2, k) = 4.5 * lambda2 * (lambda2 - 1 / 3.0) * (lambda2 - 2 / 3.0);
      r(3, k) = 13.5 * lambda0 * lambda1 * (lambda0 - 1 / 3.0);
      r(4, k) = 13.5 * lambda0 * lambda1 * (lambda1 - 1 / 3.0);
      r(5, k) = 13.5 * lambda1 * lambda2 * (lambda1 - 1 / 3.0);
      r(6, k) = 13.5 * lambda1 * lambda2 * (lambda2 - 1 / 3.0);
      r(7, k) = 13.5 * lambda2 * lambda0 * (lambda2 - 1 / 3.0);
      r(8, k) = 13.5 * lambda2 * lambda0 * (lambda0 - 1 / 3.0);
      r(9, k) = 27.0 * lambda0 * lambda1 * lambda2;
    }
    return r;
  }
  [[nodiscard]] Mat GradientsReferenceShapeFunctions(const Mat& x) const override {
    Mat r(10, 2 * x.cols());
    GradView t{r, 10};
    for (long k = 0; k < x.cols(); ++k) {
      const double l0 = 1 - x(0, k) - x(1, k), l1 = x(0, k), l2 = x(1, k);
      t.dx0(0, k) = -4.5 * ((l0 - 1 / 3.0) * (l0 - 2 / 3.0) + l0 * (l0 - 2 / 3.0) + l0 * (l0 - 1 / 3.0));
      t.dx1(0, k) = -4.5 * ((l0 - 1 / 3.0) * (l0 - 2 / 3.0) + l0 * (l0 - 2 / 3.0) + l0 * (l0 - 1 / 3.0));
      t.dx0(1, k) = 4.5 * ((l1 - 1 / 3.0) * (l1 - 2 / 3.0) + l1 * (l1 - 2 / 3.0) + l1 * (l1 - 1 / 3.0));
      t.dx1(1, k) = 0.0;
      t.dx0(2, k) = 0.0;
      t.dx1(2, k) = 4.5 * ((l2 - 1 / 3.0) * (l2 - 2 / 3.0) + l2 * (l2 - 2 / 3.0) + l2 * (l2 - 1 / 3.0));
      t.dx0(3, k) = 13.5 * (-l1 * (l0 - 1 / 3.0) + l0 * (l0 - 1 / 3.0) - l0 * l1);
      t.dx1(3, k) = -13.5 * (l1 * (l0 - 1 / 3.0) + l0 * l1);
      t.dx0(4, k) = 13.5 * (-l1 * (l1 - 1 / 3.0) + l0 * (l1 - 1 / 3.0) + l0 * l1);
      t.dx1(4, k) = -13.5 * l1 * (l1 - 1 / 3.0);
      t.dx0(5, k) = 13.5 * (l2 * (l1 - 1 / 3.0) + l1 * l2);
      t.dx1(5, k) = 13.5 * (l1 * (l1 - 1 / 3.0));
      t.dx0(6, k) = 13.5 * (l2 * (l2 - 1 / 3.0));
      t.dx1(6, k) = 13.5 * (l1 * (l2 - 1 / 3.0) + l1 * l2);
      t.dx0(7, k) = -13.5 * l2 * (l2 - 1 / 3.0);
      t.dx1(7, k) = 13.5 * (l0 * (l2 - 1 / 3.0) - l2 * (l2 - 1 / 3.0) + l0 * l2);
      t.dx0(8, k) = -13.5 * (l2 * (l0 - 1 / 3.0) + l2 * l0);
      t.dx1(8, k) = 13.5 * (l0 * (l0 - 1 / 3.0) - l2 * (l0 - 1 / 3.0) - l2 * l0);
      t.dx0(9, k) = 27.0 * (-l1 * l2 + l0 * l2);
      t.dx1(9, k) = 27.0 * (-l1 * l2 + l0 * l1);
    }
    return r;
  }
  [[nodiscard]] Mat EvaluationNodes() const override {
    static const double e[2][6] = {{1.0, 2.0, 2.0, 1.0, 0.0, 0.0}, {0.0, 0.0, 1.0, 2.0, 2.0, 1.0}};
    Mat m(2, 10);
    m(0, 0) = 0; m(1, 0) = 0; m(0, 1) = 1; m(1, 1) = 0; m(0, 2) = 0; m(1, 2) = 1;
    for (int k = 0; k < 6; ++k) {
      m(0, 3 + k) = e[0][k] * (1.0 / 3.0);
      m(1, 3 + k) = e[1][k] * (1.0 / 3.0);
    }
    m(0, 9) = 1.0 / 3.0;
    m(1, 9) = 1.0 / 3.0;
    return m;
  }
};

inline constexpr int kO3QuadMap[16][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}, {2, 0}, {3, 0}, {1, 2}, {1, 3},
                                          {3, 1}, {2, 1}, {0, 3}, {0, 2}, {2, 2}, {3, 2}, {3, 3}, {2, 3}};
class FeLagrangeO3Quad final : public FeLagrangeTPQuad<FeLagrangeO3Segment, 16> {
 public:
  FeLagrangeO3Quad() : FeLagrangeTPQuad(kO3QuadMap) {}
  [[nodiscard]] unsigned Degree() const override { return 3; }
  [[nodiscard]] size_type NumRefShapeFunctions(dim_t codim) const override { return codim == 0 ? 4 : (codim == 1 ? 2 : 1); }
  using FeLagrangeTPQuad::NumRefShapeFunctions;
  [[nodiscard]] Mat EvaluationNodes() const override {
    static const double mid[2][8] = {{1.0, 2.0, 3.0, 3.0, 2.0, 1.0, 0.0, 0.0}, {0.0, 0.0, 1.0, 2.0, 3.0, 3.0, 2.0, 1.0}};
    static const double in[2][4] = {{1.0, 2.0, 2.0, 1.0}, {1.0, 1.0, 2.0, 2.0}};
    Mat m(2, 16);
    const Mat v = RefEl::kQuad().NodeCoords();
    for (int k = 0; k < 4; ++k) {
      m(0, k) = v(0, k);
      m(1, k) = v(1, k);
    }
    for (int k = 0; k < 8; ++k) {
      m(0, 4 + k) = mid[0][k] * (1.0 / 3.0);
      m(1, 4 + k) = mid[1][k] * (1.0 / 3.0);
    }
    for (int k = 0; k < 4; ++k) {
      m(0, 12 + k) = in[0][k] * (1.0 / 3.0);
      m(1, 12 + k) = in[1][k] * (1.0 / 3.0);
    }
    return m;
  }
};

// lib/lf/uscalfe/uniform_scalar_fe_space.h:50-181, InitDofHandler :241-342; fe_space_lagrange_o{1,2,3}.h
class UniformScalarFESpace {
 public:
  UniformScalarFESpace(std::shared_ptr<const mesh::Mesh> mesh, unsigned degree) : mesh_(std::move(mesh)) {
    switch (degree) {
      case 1:
        tria_ = std::make_shared<FeLagrangeO1Tria>();
        quad_ = std::make_shared<FeLagrangeO1Quad>();
        seg_ = std::make_shared<FeLagrangeO1Segment>();
        break;
      case 2:
        tria_ = std::make_shared<FeLagrangeO2Tria>();
        quad_ = std::make_shared<FeLagrangeO2Quad>();
        seg_ = std::make_shared<FeLagrangeO2Segment>();
        break;
      case 3:
        tria_ = std::make_shared<FeLagrangeO3Tria>();
        quad_ = std::make_shared<FeLagrangeO3Quad>();
        seg_ = std::make_shared<FeLagrangeO3Segment>();
        break;
      default:
        LFO_VERIFY(false, "FeSpaceLagrangeO{1,2,3} only");
    }
    // uniform_scalar_fe_space.h:250-341: interior dof counts per entity type must agree between tria/quad/segment
    const size_type n_pt = tria_->NumRefShapeFunctions(2);
    const size_type n_seg = tria_->NumRefShapeFunctions(1);
    LFO_VERIFY(quad_->NumRefShapeFunctions(2) == n_pt && quad_->NumRefShapeFunctions(1) == n_seg, "dof layout mismatch");
    assemble::UniformFEDofHandler::dof_map_t layout{{RefEl::kPoint(), n_pt},
                                                    {RefEl::kSegment(), n_seg},
                                                    {RefEl::kTria(), tria_->NumRefShapeFunctions(0)},
                                                    {RefEl::kQuad(), quad_->NumRefShapeFunctions(0)}};
    dofh_ = std::make_unique<assemble::UniformFEDofHandler>(mesh_, layout);
  }
  [[nodiscard]] std::shared_ptr<const mesh::Mesh> Mesh() const { return mesh_; }
  [[nodiscard]] const assemble::UniformFEDofHandler& LocGlobMap() const { return *dofh_; }
  [[nodiscard]] const ScalarReferenceFiniteElement* ShapeFunctionLayout(RefEl r) const {
    if (r == RefEl::kTria()) return tria_.get();
    if (r == RefEl::kQuad()) return quad_.get();
    if (r == RefEl::kSegment()) return seg_.get();
    return nullptr;
  }

 private:
  std::shared_ptr<const mesh::Mesh> mesh_;
  std::shared_ptr<const ScalarReferenceFiniteElement> tria_, quad_, seg_;
  std::unique_ptr<assemble::UniformFEDofHandler> dofh_;
};

// lib/lf/uscalfe/precomputed_scalar_reference_finite_element.h:44-185
class PrecomputedScalarReferenceFiniteElement {
 public:
  PrecomputedScalarReferenceFiniteElement() = default;
  PrecomputedScalarReferenceFiniteElement(const ScalarReferenceFiniteElement* fe, quad::QuadRule qr)
      : fe_(fe), qr_(std::move(qr)), shap_fun_(fe->EvalReferenceShapeFunctions(qr_.Points())),
        grad_shape_fun_(fe->GradientsReferenceShapeFunctions(qr_.Points())) {}
  [[nodiscard]] bool isInitialized() const { return fe_ != nullptr; }
  [[nodiscard]] const quad::QuadRule& Qr() const { return qr_; }
  [[nodiscard]] size_type NumRefShapeFunctions() const { return fe_->NumRefShapeFunctions(); }
  [[nodiscard]] const Mat& PrecompReferenceShapeFunctions() const { return shap_fun_; }
  [[nodiscard]] const Mat& PrecompGradientsReferenceShapeFunctions() const { return grad_shape_fun_; }

 private:
  const ScalarReferenceFiniteElement* fe_ = nullptr;
  quad::QuadRule qr_;
  Mat shap_fun_, grad_shape_fun_;
};

// ---- mesh functions ---------------------------------------------------------------------------------------------
struct Mat2 {  // Eigen::Matrix2d stand-in, row-major a[r][c]
  double a[2][2];
};

// lib/lf/mesh/utils/mesh_function_constant.h:26-46
template <class R>
class MeshFunctionConstant {
 public:
  explicit MeshFunctionConstant(R value) : value_(value) {}
  std::vector<R> operator()(const mesh::Entity&, const Mat& local) const { return std::vector<R>(local.cols(), value_); }

 private:
  R value_;
};

// lib/lf/mesh/utils/mesh_function_global.h:55-98 -- F: (x, y) -> R
template <class R>
class MeshFunctionGlobal {
 public:
  explicit MeshFunctionGlobal(std::function<R(double, double)> f) : f_(std::move(f)) {}
  std::vector<R> operator()(const mesh::Entity& e, const Mat& local) const {
    std::vector<R> result;
    result.reserve(local.cols());
    const Mat global_points = e.Geometry()->Global(local);
    for (long i = 0; i < local.cols(); ++i) result.push_back(f_(global_points(0, i), global_points(1, i)));
    return result;
  }

 private:
  std::function<R(double, double)> f_;
};

// A mesh function looked up from a per-cell / per-quadrature-point table (what a user of the reference would write
// to feed tabulated coefficients); `stride` values per cell, R = double.
class MeshFunctionTable {
 public:
  MeshFunctionTable(const mesh::Mesh* mesh, const double* table, long stride) : mesh_(mesh), table_(table), stride_(stride) {}
  std::vector<double> operator()(const mesh::Entity& e, const Mat& local) const {
    const std::size_t c = mesh_->Index(e);
    std::vector<double> r(local.cols());
    for (long i = 0; i < local.cols(); ++i) r[i] = table_[c * stride_ + (stride_ == 1 ? 0 : i)];
    return r;
  }

 private:
  const mesh::Mesh* mesh_;
  const double* table_;
  long stride_;
};

// ---- providers --------------------------------------------------------------------------------------------------
namespace detail {
// alphaval[k] * trf_grad for scalar and 2x2 coefficients (loc_comp_ellbvp.h:332)
inline void ApplyCoeff(double a, const Mat& g, Mat& out) {
  for (long j = 0; j < g.cols(); ++j) {
    out(0, j) = a * g(0, j);
    out(1, j) = a * g(1, j);
  }
}
inline void ApplyCoeff(const Mat2& a, const Mat& g, Mat& out) {
  for (long j = 0; j < g.cols(); ++j) {
    out(0, j) = a.a[0][0] * g(0, j) + a.a[0][1] * g(1, j);
    out(1, j) = a.a[1][0] * g(0, j) + a.a[1][1] * g(1, j);
  }
}
}  // namespace detail

// lib/lf/uscalfe/loc_comp_ellbvp.h:85-339
template <class DIFF_COEFF, class REACTION_COEFF>
class ReactionDiffusionElementMatrixProvider {
 public:
  using ElemMat = Mat;
  using quad_rule_collection_t = std::map<RefEl, quad::QuadRule>;
  // :210-231 default rules of degree 2 * fe->Degree()
  ReactionDiffusionElementMatrixProvider(std::shared_ptr<const UniformScalarFESpace> fe_space, DIFF_COEFF alpha,
                                         REACTION_COEFF gamma)
      : alpha_(std::move(alpha)), gamma_(std::move(gamma)) {
    for (auto ref_el : {RefEl::kTria(), RefEl::kQuad()}) {
      auto fe = fe_space->ShapeFunctionLayout(ref_el);
      if (fe != nullptr) {
        fe_precomp_[ref_el.Id()] = PrecomputedScalarReferenceFiniteElement(fe, quad::make_QuadRule(ref_el, 2 * fe->Degree()));
      }
    }
  }
  // :234-263 user-supplied rules
  ReactionDiffusionElementMatrixProvider(std::shared_ptr<const UniformScalarFESpace> fe_space, DIFF_COEFF alpha,
                                         REACTION_COEFF gamma, const quad_rule_collection_t& qr_collection)
      : alpha_(std::move(alpha)), gamma_(std::move(gamma)) {
    for (auto ref_el : {RefEl::kTria(), RefEl::kQuad()}) {
      auto fe = fe_space->ShapeFunctionLayout(ref_el);
      if (fe != nullptr) {
        auto it = qr_collection.find(ref_el);
        if (it != qr_collection.end()) {
          LFO_VERIFY(it->second.RefElem() == ref_el, "qr.RefEl() mismatch");
          fe_precomp_[ref_el.Id()] = PrecomputedScalarReferenceFiniteElement(fe, it->second);
        }
      }
    }
  }
  virtual ~ReactionDiffusionElementMatrixProvider() = default;
  virtual bool isActive(const mesh::Entity& /*cell*/) { return true; }

  // :266-339
  ElemMat Eval(const mesh::Entity& cell) {
    const RefEl ref_el{cell.RefElem()};
    const PrecomputedScalarReferenceFiniteElement& pfe = fe_precomp_[ref_el.Id()];
    if (!pfe.isInitialized()) {
      throw LfException("No local shape function information or no quadrature rule for reference element type");
    }
    const geometry::Geometry* geo_ptr = cell.Geometry();
    const Mat determinants(geo_ptr->IntegrationElement(pfe.Qr().Points()));
    const Mat JinvT(geo_ptr->JacobianInverseGramian(pfe.Qr().Points()));
    auto alphaval = alpha_(cell, pfe.Qr().Points());
    auto gammaval = gamma_(cell, pfe.Qr().Points());
    const long nsf = pfe.NumRefShapeFunctions();
    ElemMat mat(nsf, nsf);
    mat.setZero();
    const Mat& gradhat = pfe.PrecompGradientsReferenceShapeFunctions();
    const Mat& phi = pfe.PrecompReferenceShapeFunctions();
    for (size_type k = 0; k < pfe.Qr().NumPoints(); ++k) {
      const double w = pfe.Qr().Weights()[k] * determinants[k];
      // trf_grad = JinvT.block(0,2k,2,2) * gradhat.block(0,2k,nsf,2)^T   (2 x nsf)
      Mat trf_grad(2, nsf);
      for (long a = 0; a < nsf; ++a) {
        trf_grad(0, a) = JinvT(0, 2 * k) * gradhat(a, 2 * k) + JinvT(0, 2 * k + 1) * gradhat(a, 2 * k + 1);
        trf_grad(1, a) = JinvT(1, 2 * k) * gradhat(a, 2 * k) + JinvT(1, 2 * k + 1) * gradhat(a, 2 * k + 1);
      }
      Mat alpha_trf_grad(2, nsf);
      detail::ApplyCoeff(alphaval[k], trf_grad, alpha_trf_grad);
      // mat += w * (trf_grad^H * alpha_trf_grad + (gamma * phi_k) * phi_k^H)
      for (long b = 0; b < nsf; ++b) {
        for (long a = 0; a < nsf; ++a) {
          const double stiff = trf_grad(0, a) * alpha_trf_grad(0, b) + trf_grad(1, a) * alpha_trf_grad(1, b);
          const double mass = (gammaval[k] * phi(a, k)) * phi(b, k);
          mat(a, b) += w * (stiff + mass);
        }
      }
    }
    return mat;
  }

 private:
  DIFF_COEFF alpha_;
  REACTION_COEFF gamma_;
  std::array<PrecomputedScalarReferenceFiniteElement, 5> fe_precomp_;
};

// lib/lf/uscalfe/loc_comp_ellbvp.h:562-746
template <class MESH_FUNCTION>
class ScalarLoadElementVectorProvider {
 public:
  using ElemVec = Mat;
  using quad_rule_collection_t = std::map<RefEl, quad::QuadRule>;
  ScalarLoadElementVectorProvider(std::shared_ptr<const UniformScalarFESpace> fe_space, MESH_FUNCTION f) : f_(std::move(f)) {
    for (auto ref_el : {RefEl::kTria(), RefEl::kQuad()}) {
      auto fe = fe_space->ShapeFunctionLayout(ref_el);
      if (fe != nullptr) {
        fe_precomp_[ref_el.Id()] = PrecomputedScalarReferenceFiniteElement(fe, quad::make_QuadRule(ref_el, 2 * fe->Degree()));
      }
    }
  }
  ScalarLoadElementVectorProvider(std::shared_ptr<const UniformScalarFESpace> fe_space, MESH_FUNCTION f,
                                  const quad_rule_collection_t& qr_collection)
      : f_(std::move(f)) {
    for (auto ref_el : {RefEl::kTria(), RefEl::kQuad()}) {
      auto fe = fe_space->ShapeFunctionLayout(ref_el);
      if (fe != nullptr) {
        auto it = qr_collection.find(ref_el);
        LFO_VERIFY(it != qr_collection.end(), "Quadrature rule missing");  // :680-684
        fe_precomp_[ref_el.Id()] = PrecomputedScalarReferenceFiniteElement(fe, it->second);
      }
    }
  }
  virtual ~ScalarLoadElementVectorProvider() = default;
  virtual bool isActive(const mesh::Entity& /*cell*/) { return true; }

  // :691-746
  ElemVec Eval(const mesh::Entity& cell) {
    const RefEl ref_el{cell.RefElem()};
    auto& pfe = fe_precomp_[ref_el.Id()];
    LFO_VERIFY(pfe.isInitialized(), "No local shape function information for entity type");
    const geometry::Geometry* geo_ptr = cell.Geometry();
    const Mat determinants(geo_ptr->IntegrationElement(pfe.Qr().Points()));
    const long nsf = pfe.NumRefShapeFunctions();
    ElemVec vec(nsf, 1);
    vec.setZero();
    auto fval = f_(cell, pfe.Qr().Points());
    const Mat& phi = pfe.PrecompReferenceShapeFunctions();
    for (long k = 0; k < determinants.size(); ++k) {
      const double s = pfe.Qr().Weights()[k] * determinants[k] * fval[k];
      for (long a = 0; a < nsf; ++a) vec[a] += s * phi(a, k);
    }
    return vec;
  }

 private:
  MESH_FUNCTION f_;
  std::array<PrecomputedScalarReferenceFiniteElement, 5> fe_precomp_;
};

// lib/lf/uscalfe/loc_comp_ellbvp.h:367-529 -- edge (codim-1) mass matrix, for impedance / Robin boundary terms.
// EDGESELECTOR: const mesh::Entity& -> bool
template <class COEFF, class EDGESELECTOR>
class MassEdgeMatrixProvider {
 public:
  using ElemMat = Mat;
  MassEdgeMatrixProvider(std::shared_ptr<const UniformScalarFESpace> fe_space, COEFF gamma, EDGESELECTOR edge_selector)
      : gamma_(std::move(gamma)), edge_sel_(std::move(edge_selector)) {
    auto fe = fe_space->ShapeFunctionLayout(RefEl::kSegment());
    LFO_VERIFY(fe != nullptr, "No shape functions specified for edges");
    fe_precomp_ = PrecomputedScalarReferenceFiniteElement(fe, quad::make_QuadRule(RefEl::kSegment(), 2 * fe->Degree()));
  }
  MassEdgeMatrixProvider(std::shared_ptr<const UniformScalarFESpace> fe_space, COEFF gamma, quad::QuadRule quadrule,
                         EDGESELECTOR edge_selector)
      : gamma_(std::move(gamma)), edge_sel_(std::move(edge_selector)) {
    auto fe = fe_space->ShapeFunctionLayout(RefEl::kSegment());
    LFO_VERIFY(fe != nullptr, "No shape functions specified for edges");
    LFO_VERIFY(quadrule.RefElem() == RefEl::kSegment(), "Quadrature rule not meant for EDGE entities!");
    fe_precomp_ = PrecomputedScalarReferenceFiniteElement(fe, std::move(quadrule));
  }
  virtual ~MassEdgeMatrixProvider() = default;
  bool isActive(const mesh::Entity& edge) {
    LFO_VERIFY(edge.RefElem() == RefEl::kSegment(), "Wrong type for an edge");
    return edge_sel_(edge);
  }
  // :491-529
  ElemMat Eval(const mesh::Entity& edge) {
    LFO_VERIFY(edge.RefElem() == RefEl::kSegment(), "Edge must be of segment type");
    const geometry::Geometry* geo_ptr = edge.Geometry();
    LFO_VERIFY(geo_ptr != nullptr, "Invalid geometry!");
    const Mat determinants(geo_ptr->IntegrationElement(fe_precomp_.Qr().Points()));
    const long nsf = fe_precomp_.NumRefShapeFunctions();
    ElemMat mat(nsf, nsf);
    mat.setZero();
    auto gammaval = gamma_(edge, fe_precomp_.Qr().Points());
    const Mat& phi = fe_precomp_.PrecompReferenceShapeFunctions();
    for (long k = 0; k < determinants.size(); ++k) {
      const double w = (fe_precomp_.Qr().Weights()[k] * determinants[k]) * gammaval[k];
      for (long b = 0; b < nsf; ++b)
        for (long a = 0; a < nsf; ++a) mat(a, b) += (phi(a, k) * phi(b, k)) * w;
    }
    return mat;
  }

 private:
  COEFF gamma_;
  EDGESELECTOR edge_sel_;
  PrecomputedScalarReferenceFiniteElement fe_precomp_;
};

// lib/lf/uscalfe/loc_comp_ellbvp.h:784-921 -- edge load vector (Neumann / impedance data)
template <class FUNCTOR, class EDGESELECTOR>
class ScalarLoadEdgeVectorProvider {
 public:
  using ElemVec = Mat;
  ScalarLoadEdgeVectorProvider(std::shared_ptr<const UniformScalarFESpace> fe_space, FUNCTOR g, EDGESELECTOR edge_sel)
      : g_(std::move(g)), edge_sel_(std::move(edge_sel)) {
    auto fe = fe_space->ShapeFunctionLayout(RefEl::kSegment());
    LFO_VERIFY(fe != nullptr, "No shape functions specified for edges");
    pfe_ = PrecomputedScalarReferenceFiniteElement(fe, quad::make_QuadRule(RefEl::kSegment(), 2 * fe->Degree()));
  }
  ScalarLoadEdgeVectorProvider(std::shared_ptr<const UniformScalarFESpace> fe_space, FUNCTOR g, quad::QuadRule quadrule,
                               EDGESELECTOR edge_sel)
      : g_(std::move(g)), edge_sel_(std::move(edge_sel)) {
    auto fe = fe_space->ShapeFunctionLayout(RefEl::kSegment());
    LFO_VERIFY(fe != nullptr, "No shape functions specified for edges");
    LFO_VERIFY(quadrule.RefElem() == RefEl::kSegment(), "Quadrature rule not meant for EDGE entities!");
    pfe_ = PrecomputedScalarReferenceFiniteElement(fe, std::move(quadrule));
  }
  virtual ~ScalarLoadEdgeVectorProvider() = default;
  virtual bool isActive(const mesh::Entity& edge) { return edge_sel_(edge); }
  // :886-921
  ElemVec Eval(const mesh::Entity& edge) {
    LFO_VERIFY(edge.RefElem() == RefEl::kSegment(), "Edge must be of segment type");
    const geometry::Geometry* geo_ptr = edge.Geometry();
    LFO_VERIFY(geo_ptr != nullptr, "Invalid geometry!");
    const Mat determinants(geo_ptr->IntegrationElement(pfe_.Qr().Points()));
    const long nsf = pfe_.NumRefShapeFunctions();
    ElemVec vec(nsf, 1);
    vec.setZero();
    auto g_vals = g_(edge, pfe_.Qr().Points());
    const Mat& phi = pfe_.PrecompReferenceShapeFunctions();
    for (long k = 0; k < determinants.size(); ++k) {
      const double w = (pfe_.Qr().Weights()[k] * determinants[k]) * g_vals[k];
      for (long a = 0; a < nsf; ++a) vec[a] += phi(a, k) * w;
    }
    return vec;
  }

 private:
  FUNCTOR g_;
  EDGESELECTOR edge_sel_;
  PrecomputedScalarReferenceFiniteElement pfe_;
};

// lib/lf/fe/fe_tools.h:198-258 (NodalValuesToDofs is the identity for the Lagrange elements here)
template <class MF>
std::vector<double> NodalProjection(const UniformScalarFESpace& fe_space, const MF& u) {
  const mesh::Mesh& mesh = *fe_space.Mesh();
  const assemble::DofHandler& dofh = fe_space.LocGlobMap();
  std::vector<double> glob(dofh.NumDofs(), 0.0);
  for (const mesh::Entity* cell : mesh.Entities(0)) {
    const auto* rsf = fe_space.ShapeFunctionLayout(cell->RefElem());
    const Mat ref_nodes(rsf->EvaluationNodes());
    auto uval = u(*cell, ref_nodes);
    const size_type n = dofh.NumLocalDofs(*cell);
    const auto idx = dofh.GlobalDofIndices(*cell);
    for (size_type j = 0; j < n; ++j) glob[idx[j]] = uval[j];
  }
  return glob;
}

}  // namespace lfo::uscalfe

// lib/lf/fe/loc_comp_ellbvp.h: the providers of the generic lf::fe module, restated for the spaces this oracle has (the
// Lagrange spaces; in the reference FeSpaceLagrangeO<p> IS-A lf::fe::ScalarFESpace).  They differ from the uscalfe provider
// in form only: one term each, shape functions evaluated per call instead of precomputed, rule of degree 2 * Degree() from a
// QuadRuleCache -- the same rule the uscalfe provider uses by default.
namespace lfo::fe {
using uscalfe::UniformScalarFESpace;

// fe/loc_comp_ellbvp.h:76-226
template <class DIFF_COEFF>
class DiffusionElementMatrixProvider {
 public:
  using ElemMat = Mat;
  DiffusionElementMatrixProvider(std::shared_ptr<const UniformScalarFESpace> fe_space, DIFF_COEFF alpha)
      : alpha_(std::move(alpha)), fe_space_(std::move(fe_space)) {}
  bool isActive(const mesh::Entity& /*cell*/) const { return true; }
  // :170-226
  ElemMat Eval(const mesh::Entity& cell) {
    const geometry::Geometry* geo_ptr = cell.Geometry();
    const auto sfl = fe_space_->ShapeFunctionLayout(cell.RefElem());
    const quad::QuadRule qr = quad::make_QuadRule(cell.RefElem(), 2 * sfl->Degree());
    const Mat determinants(geo_ptr->IntegrationElement(qr.Points()));
    const Mat JinvT(geo_ptr->JacobianInverseGramian(qr.Points()));
    auto alphaval = alpha_(cell, qr.Points());
    const long nsf = sfl->NumRefShapeFunctions();
    ElemMat mat(nsf, nsf);
    mat.setZero();
    const Mat grsf = sfl->GradientsReferenceShapeFunctions(qr.Points());
    for (size_type k = 0; k < qr.NumPoints(); ++k) {
      const double w = qr.Weights()[k] * determinants[k];
      Mat trf_grad(2, nsf);
      for (long a = 0; a < nsf; ++a) {
        trf_grad(0, a) = JinvT(0, 2 * k) * grsf(a, 2 * k) + JinvT(0, 2 * k + 1) * grsf(a, 2 * k + 1);
        trf_grad(1, a) = JinvT(1, 2 * k) * grsf(a, 2 * k) + JinvT(1, 2 * k + 1) * grsf(a, 2 * k + 1);
      }
      Mat alpha_trf_grad(2, nsf);
      uscalfe::detail::ApplyCoeff(alphaval[k], trf_grad, alpha_trf_grad);
      // mat += w * trf_grad^H * (alpha * trf_grad)
      for (long b = 0; b < nsf; ++b)
        for (long a = 0; a < nsf; ++a) mat(a, b) += w * (trf_grad(0, a) * alpha_trf_grad(0, b) + trf_grad(1, a) * alpha_trf_grad(1, b));
    }
    return mat;
  }

 private:
  DIFF_COEFF alpha_;
  std::shared_ptr<const UniformScalarFESpace> fe_space_;
};

// fe/loc_comp_ellbvp.h:256-384
template <class REACTION_COEFF>
class MassElementMatrixProvider {
 public:
  using ElemMat = Mat;
  MassElementMatrixProvider(std::shared_ptr<const UniformScalarFESpace> fe_space, REACTION_COEFF gamma)
      : gamma_(std::move(gamma)), fe_space_(std::move(fe_space)) {}
  bool isActive(const mesh::Entity& /*cell*/) const { return true; }
  // :351-384
  ElemMat Eval(const mesh::Entity& cell) {
    const geometry::Geometry* geo_ptr = cell.Geometry();
    const auto sfl = fe_space_->ShapeFunctionLayout(cell.RefElem());
    const quad::QuadRule qr = quad::make_QuadRule(cell.RefElem(), 2 * sfl->Degree());
    const Mat determinants(geo_ptr->IntegrationElement(qr.Points()));
    auto gammaval = gamma_(cell, qr.Points());
    const long nsf = sfl->NumRefShapeFunctions();
    ElemMat mat(nsf, nsf);
    mat.setZero();
    const Mat rsf = sfl->EvalReferenceShapeFunctions(qr.Points());
    for (size_type k = 0; k < qr.NumPoints(); ++k) {
      const double w = qr.Weights()[k] * determinants[k];
      for (long b = 0; b < nsf; ++b)
        for (long a = 0; a < nsf; ++a) mat(a, b) += w * ((gammaval[k] * rsf(a, k)) * rsf(b, k));
    }
    return mat;
  }

 private:
  REACTION_COEFF gamma_;
  std::shared_ptr<const UniformScalarFESpace> fe_space_;
};

}  // namespace lfo::fe
#endif
