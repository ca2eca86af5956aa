// ORACLE (test infrastructure, NOT product code) -- see lfo_base.h header.
// Geometry objects of the hot path: lib/lf/geometry/{geometry_interface.h,tria_o1.cc,quad_o1.cc,point.cc,segment_o1.cc}
#ifndef LFO_GEOMETRY_H
#define LFO_GEOMETRY_H

#include "lfo_base.h"

namespace lfo::geometry {

// lib/lf/geometry/geometry_interface.h:21-227 (only the members the assembly path calls)
class Geometry {
 public:
  virtual ~Geometry() = default;
  [[nodiscard]] virtual dim_t DimLocal() const = 0;
  [[nodiscard]] virtual dim_t DimGlobal() const = 0;
  [[nodiscard]] virtual RefEl RefElem() const = 0;
  [[nodiscard]] virtual Mat Global(const Mat& local) const = 0;
  [[nodiscard]] virtual Mat Jacobian(const Mat& local) const = 0;
  [[nodiscard]] virtual Mat JacobianInverseGramian(const Mat& local) const = 0;
  [[nodiscard]] virtual Mat IntegrationElement(const Mat& local) const = 0;  // returned as n x 1
  [[nodiscard]] virtual std::unique_ptr<Geometry> SubGeometry(dim_t codim, dim_t i) const = 0;
};

// Eigen 3.4 fixed-size 2x2 semantics (App. A.6 of SURVEY.md): determinant = ad - bc, inverse = adjugate * (1/det)
inline double Det2(double a, double b, double c, double d) { return a * d - b * c; }
// inverse of the TRANSPOSE of J = [a b; c d] written into out(0..1, col..col+1)
inline void InvTranspose2(double a, double b, double c, double d, Mat& out, long col) {
  // J^T = [a c; b d]; det(J^T) = a*d - c*b ; inverse = 1/det * [d -c; -b a]
  const double det = a * d - c * b;
  const double invdet = 1.0 / det;
  out(0, col) = d * invdet;
  out(0, col + 1) = -c * invdet;
  out(1, col) = -b * invdet;
  out(1, col + 1) = a * invdet;
}

// lib/lf/geometry/point.cc
class Point final : public Geometry {
 public:
  explicit Point(double x, double y) : x_(x), y_(y) {}
  [[nodiscard]] dim_t DimLocal() const override { return 0; }
  [[nodiscard]] dim_t DimGlobal() const override { return 2; }
  [[nodiscard]] RefEl RefElem() const override { return RefEl::kPoint(); }
  [[nodiscard]] Mat Global(const Mat& local) const override {
    Mat r(2, local.cols() > 0 ? local.cols() : 1);
    for (long i = 0; i < r.cols(); ++i) {
      r(0, i) = x_;
      r(1, i) = y_;
    }
    return r;
  }
  [[nodiscard]] Mat Jacobian(const Mat&) const override { return Mat(2, 0); }
  [[nodiscard]] Mat JacobianInverseGramian(const Mat&) const override { return Mat(2, 0); }
  [[nodiscard]] Mat IntegrationElement(const Mat& local) const override {
    Mat r(local.cols(), 1);
    for (long i = 0; i < r.size(); ++i) r[i] = 1.0;
    return r;
  }
  [[nodiscard]] std::unique_ptr<Geometry> SubGeometry(dim_t, dim_t) const override {
    return std::make_unique<Point>(x_, y_);
  }
  [[nodiscard]] double x() const { return x_; }
  [[nodiscard]] double y() const { return y_; }

 private:
  double x_, y_;
};

// lib/lf/geometry/segment_o1.cc (straight edge; only what mesh construction needs)
class SegmentO1 final : public Geometry {
 public:
  explicit SegmentO1(const Mat& coords) : coords_(coords) {}
  [[nodiscard]] dim_t DimLocal() const override { return 1; }
  [[nodiscard]] dim_t DimGlobal() const override { return 2; }
  [[nodiscard]] RefEl RefElem() const override { return RefEl::kSegment(); }
  [[nodiscard]] Mat Global(const Mat& local) const override {
    Mat r(2, local.cols());
    for (long i = 0; i < local.cols(); ++i) {
      for (int d = 0; d < 2; ++d) r(d, i) = coords_(d, 1) * local(0, i) + coords_(d, 0) * (1 - local(0, i));
    }
    return r;
  }
  [[nodiscard]] Mat Jacobian(const Mat& local) const override {
    Mat r(2, local.cols());
    for (long i = 0; i < local.cols(); ++i) {
      for (int d = 0; d < 2; ++d) r(d, i) = coords_(d, 1) - coords_(d, 0);
    }
    return r;
  }
  [[nodiscard]] Mat JacobianInverseGramian(const Mat& local) const override {
    Mat r(2, local.cols());
    const double dx = coords_(0, 1) - coords_(0, 0), dy = coords_(1, 1) - coords_(1, 0);
    const double n2 = dx * dx + dy * dy;
    for (long i = 0; i < local.cols(); ++i) {
      r(0, i) = dx / n2;
      r(1, i) = dy / n2;
    }
    return r;
  }
  [[nodiscard]] Mat IntegrationElement(const Mat& local) const override {
    Mat r(local.cols(), 1);
    const double dx = coords_(0, 1) - coords_(0, 0), dy = coords_(1, 1) - coords_(1, 0);
    for (long i = 0; i < r.size(); ++i) r[i] = std::sqrt(dx * dx + dy * dy);
    return r;
  }
  [[nodiscard]] std::unique_ptr<Geometry> SubGeometry(dim_t codim, dim_t i) const override {
    if (codim == 0) return std::make_unique<SegmentO1>(coords_);
    return std::make_unique<Point>(coords_(0, i), coords_(1, i));
  }

 private:
  Mat coords_;
};

// lib/lf/geometry/tria_o1.cc:10-48
inline void assertNonDegenerateTriangle(const Mat& c, double tol = 1.0e-8) {
  auto sq = [&](int a, int b) {
    const double dx = c(0, a) - c(0, b), dy = c(1, a) - c(1, b);
    return dx * dx + dy * dy;
  };
  const double e0 = sq(1, 0), e1 = sq(2, 1), e2 = sq(0, 2);
  const double circum = e0 + e1 + e2;
  LFO_VERIFY(e0 > tol * circum, "Collapsed edge 0");
  LFO_VERIFY(e1 > tol * circum, "Collapsed edge 1");
  LFO_VERIFY(e2 > tol * circum, "Collapsed edge 2");
  const double area = std::fabs((c(0, 1) - c(0, 0)) * (c(1, 2) - c(1, 0)) - (c(1, 1) - c(1, 0)) * (c(0, 2) - c(0, 0)));
  LFO_VERIFY(area > tol * circum, "Degenerate 2D triangle");
}

// lib/lf/geometry/tria_o1.cc:50-74, tria_o1.h:35-46 : affine triangle, all metric data constant and precomputed
class TriaO1 final : public Geometry {
 public:
  explicit TriaO1(const Mat& coords) : coords_(coords), jacobian_(2, 2), jinvt_(2, 2) {
    assertNonDegenerateTriangle(coords_);
    // jacobian_ << c1 - c0, c2 - c0  (tria_o1.cc:57)
    for (int d = 0; d < 2; ++d) {
      jacobian_(d, 0) = coords_(d, 1) - coords_(d, 0);
      jacobian_(d, 1) = coords_(d, 2) - coords_(d, 0);
    }
    // jacobian_.transpose().inverse(), std::abs(jacobian_.determinant())  (tria_o1.cc:60-61)
    InvTranspose2(jacobian_(0, 0), jacobian_(0, 1), jacobian_(1, 0), jacobian_(1, 1), jinvt_, 0);
    integration_element_ = std::abs(Det2(jacobian_(0, 0), jacobian_(0, 1), jacobian_(1, 0), jacobian_(1, 1)));
  }
  [[nodiscard]] dim_t DimLocal() const override { return 2; }
  [[nodiscard]] dim_t DimGlobal() const override { return 2; }
  [[nodiscard]] RefEl RefElem() const override { return RefEl::kTria(); }
  // tria_o1.cc:70-74
  [[nodiscard]] Mat Global(const Mat& local) const override {
    Mat r(2, local.cols());
    for (long i = 0; i < local.cols(); ++i) {
      const double l0 = 1 - local(0, i) - local(1, i);
      for (int d = 0; d < 2; ++d) r(d, i) = coords_(d, 0) * l0 + coords_(d, 1) * local(0, i) + coords_(d, 2) * local(1, i);
    }
    return r;
  }
  // tria_o1.h:35-46: replicate the constant matrices
  [[nodiscard]] Mat Jacobian(const Mat& local) const override { return Replicate(jacobian_, local.cols()); }
  [[nodiscard]] Mat JacobianInverseGramian(const Mat& local) const override { return Replicate(jinvt_, local.cols()); }
  [[nodiscard]] Mat IntegrationElement(const Mat& local) const override {
    Mat r(local.cols(), 1);
    for (long i = 0; i < r.size(); ++i) r[i] = integration_element_;
    return r;
  }
  [[nodiscard]] std::unique_ptr<Geometry> SubGeometry(dim_t codim, dim_t i) const override {
    if (codim == 0) return std::make_unique<TriaO1>(coords_);
    if (codim == 1) {
      Mat c(2, 2);
      for (int d = 0; d < 2; ++d) {
        c(d, 0) = coords_(d, RefEl::kTria().EdgeEndpoint(i, 0));
        c(d, 1) = coords_(d, RefEl::kTria().EdgeEndpoint(i, 1));
      }
      return std::make_unique<SegmentO1>(c);
    }
    return std::make_unique<Point>(coords_(0, i), coords_(1, i));
  }

 private:
  static Mat Replicate(const Mat& m, long n) {
    Mat r(2, 2 * n);
    for (long k = 0; k < n; ++k) {
      r(0, 2 * k) = m(0, 0); r(1, 2 * k) = m(1, 0); r(0, 2 * k + 1) = m(0, 1); r(1, 2 * k + 1) = m(1, 1);
    }
    return r;
  }
  Mat coords_;
  Mat jacobian_;
  Mat jinvt_;
  double integration_element_ = 0;
};

// lib/lf/geometry/quad_o1.cc:14-59 (area/edge sanity check, simplified to the 2D branch)
inline void assertNonDegenerateQuad(const Mat& c, double tol = 1.0e-8) {
  auto sq = [&](int a, int b) {
    const double dx = c(0, a) - c(0, b), dy = c(1, a) - c(1, b);
    return dx * dx + dy * dy;
  };
  const double e0 = sq(1, 0), e1 = sq(2, 1), e2 = sq(3, 2), e3 = sq(0, 3);
  const double circum = e0 + e1 + e2 + e3;
  LFO_VERIFY(e0 > tol * circum, "Collapsed edge 0");
  LFO_VERIFY(e1 > tol * circum, "Collapsed edge 1");
  LFO_VERIFY(e2 > tol * circum, "Collapsed edge 2");
  LFO_VERIFY(e3 > tol * circum, "Collapsed edge 3");
  const double ar1 = ((c(0, 1) - c(0, 0)) * (c(1, 2) - c(1, 0)) - (c(1, 1) - c(1, 0)) * (c(0, 2) - c(0, 0)));
  const double ar2 = ((c(0, 3) - c(0, 0)) * (c(1, 2) - c(1, 0)) - (c(1, 3) - c(1, 0)) * (c(0, 2) - c(0, 0)));
  const double area = std::fabs(ar1) + std::fabs(ar2);
  LFO_VERIFY(area > tol * circum, "Degenerate 2D quad");
}

// lib/lf/geometry/quad_o1.cc:61-158 : bilinear quadrilateral, metric data recomputed per evaluation point
class QuadO1 final : public Geometry {
 public:
  explicit QuadO1(const Mat& coords) : coords_(coords) { assertNonDegenerateQuad(coords_); }
  [[nodiscard]] dim_t DimLocal() const override { return 2; }
  [[nodiscard]] dim_t DimGlobal() const override { return 2; }
  [[nodiscard]] RefEl RefElem() const override { return RefEl::kQuad(); }
  // quad_o1.cc:68-83
  [[nodiscard]] Mat Global(const Mat& local) const override {
    Mat r(2, local.cols());
    for (long i = 0; i < local.cols(); ++i) {
      const double x0 = local(0, i), x1 = local(1, i);
      for (int d = 0; d < 2; ++d) {
        r(d, i) = coords_(d, 0) * ((1 - x0) * (1 - x1)) + coords_(d, 1) * (x0 * (1 - x1)) + coords_(d, 2) * (x0 * x1) +
                  coords_(d, 3) * ((1 - x0) * x1);
      }
    }
    return r;
  }
  // quad_o1.cc:114-117
  void JacobianAt(double x0, double x1, double J[4]) const {  // J = [J00 J01; J10 J11] row-major
    for (int d = 0; d < 2; ++d) {
      J[2 * d + 0] = (coords_(d, 1) - coords_(d, 0)) * (1 - x1) + (coords_(d, 2) - coords_(d, 3)) * x1;
      J[2 * d + 1] = (coords_(d, 3) - coords_(d, 0)) * (1 - x0) + (coords_(d, 2) - coords_(d, 1)) * x0;
    }
  }
  [[nodiscard]] Mat Jacobian(const Mat& local) const override {
    Mat r(2, 2 * local.cols());
    for (long i = 0; i < local.cols(); ++i) {
      double J[4];
      JacobianAt(local(0, i), local(1, i), J);
      r(0, 2 * i) = J[0]; r(1, 2 * i) = J[2]; r(0, 2 * i + 1) = J[1]; r(1, 2 * i + 1) = J[3];
    }
    return r;
  }
  // quad_o1.cc:123
  [[nodiscard]] Mat JacobianInverseGramian(const Mat& local) const override {
    Mat r(2, 2 * local.cols());
    for (long i = 0; i < local.cols(); ++i) {
      double J[4];
      JacobianAt(local(0, i), local(1, i), J);
      InvTranspose2(J[0], J[1], J[2], J[3], r, 2 * i);
    }
    return r;
  }
  // quad_o1.cc:150
  [[nodiscard]] Mat IntegrationElement(const Mat& local) const override {
    Mat r(local.cols(), 1);
    for (long i = 0; i < local.cols(); ++i) {
      double J[4];
      JacobianAt(local(0, i), local(1, i), J);
      r[i] = std::abs(Det2(J[0], J[1], J[2], J[3]));
    }
    return r;
  }
  [[nodiscard]] std::unique_ptr<Geometry> SubGeometry(dim_t codim, dim_t i) const override {
    if (codim == 0) return std::make_unique<QuadO1>(coords_);
    if (codim == 1) {
      Mat c(2, 2);
      for (int d = 0; d < 2; ++d) {
        c(d, 0) = coords_(d, RefEl::kQuad().EdgeEndpoint(i, 0));
        c(d, 1) = coords_(d, RefEl::kQuad().EdgeEndpoint(i, 1));
      }
      return std::make_unique<SegmentO1>(c);
    }
    return std::make_unique<Point>(coords_(0, i), coords_(1, i));
  }

 private:
  Mat coords_;
};

// lib/lf/geometry/quad_o1.cc:255-330 : affine quadrilateral, constant metric data from corners 0,1,3
class Parallelogram final : public Geometry {
 public:
  explicit Parallelogram(const Mat& coords) : coords_(coords), jacobian_(2, 2), jinvt_(2, 2) {
    assertNonDegenerateQuad(coords_);
    for (int d = 0; d < 2; ++d) {
      jacobian_(d, 0) = coords_(d, 1) - coords_(d, 0);
      jacobian_(d, 1) = coords_(d, 3) - coords_(d, 0);
    }
    InvTranspose2(jacobian_(0, 0), jacobian_(0, 1), jacobian_(1, 0), jacobian_(1, 1), jinvt_, 0);
    integration_element_ = std::abs(Det2(jacobian_(0, 0), jacobian_(0, 1), jacobian_(1, 0), jacobian_(1, 1)));
  }
  [[nodiscard]] dim_t DimLocal() const override { return 2; }
  [[nodiscard]] dim_t DimGlobal() const override { return 2; }
  [[nodiscard]] RefEl RefElem() const override { return RefEl::kQuad(); }
  [[nodiscard]] Mat Global(const Mat& local) const override {
    Mat r(2, local.cols());
    for (long i = 0; i < local.cols(); ++i) {
      const double l0 = 1 - local(0, i) - local(1, i);
      for (int d = 0; d < 2; ++d) r(d, i) = coords_(d, 0) * l0 + coords_(d, 1) * local(0, i) + coords_(d, 3) * local(1, i);
    }
    return r;
  }
  [[nodiscard]] Mat Jacobian(const Mat& local) const override { return Rep(jacobian_, local.cols()); }
  [[nodiscard]] Mat JacobianInverseGramian(const Mat& local) const override { return Rep(jinvt_, local.cols()); }
  [[nodiscard]] Mat IntegrationElement(const Mat& local) const override {
    Mat r(local.cols(), 1);
    for (long i = 0; i < r.size(); ++i) r[i] = integration_element_;
    return r;
  }
  [[nodiscard]] std::unique_ptr<Geometry> SubGeometry(dim_t codim, dim_t i) const override {
    if (codim == 0) return std::make_unique<Parallelogram>(coords_);
    if (codim == 1) {
      Mat c(2, 2);
      for (int d = 0; d < 2; ++d) {
        c(d, 0) = coords_(d, RefEl::kQuad().EdgeEndpoint(i, 0));
        c(d, 1) = coords_(d, RefEl::kQuad().EdgeEndpoint(i, 1));
      }
      return std::make_unique<SegmentO1>(c);
    }
    return std::make_unique<Point>(coords_(0, i), coords_(1, i));
  }

 private:
  static Mat Rep(const Mat& m, long n) {
    Mat r(2, 2 * n);
    for (long k = 0; k < n; ++k) {
      r(0, 2 * k) = m(0, 0); r(1, 2 * k) = m(1, 0); r(0, 2 * k + 1) = m(0, 1); r(1, 2 * k + 1) = m(1, 1);
    }
    return r;
  }
  Mat coords_, jacobian_, jinvt_;
  double integration_element_ = 0;
};

}  // namespace lfo::geometry
#endif
