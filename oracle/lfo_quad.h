// ORACLE (test infrastructure, NOT product code) -- see lfo_base.h header.
// Quadrature rules: lib/lf/quad/quad_rule.h, make_quad_rule.cc:21-157, quad_rules_tria.cc, gauss_quadrature.cc:16-67
#ifndef LFO_QUAD_H
#define LFO_QUAD_H

#include "lfo_base.h"

namespace lfo::quad {

// lib/lf/quad/quad_rule.h
class QuadRule {
 public:
  QuadRule() : ref_el_(RefEl::kPoint()), degree_(0) {}
  QuadRule(RefEl ref_el, Mat points, Mat weights, unsigned degree)
      : ref_el_(ref_el), degree_(degree), points_(std::move(points)), weights_(std::move(weights)) {}
  [[nodiscard]] RefEl RefElem() const { return ref_el_; }
  [[nodiscard]] unsigned Degree() const { return degree_; }
  [[nodiscard]] const Mat& Points() const { return points_; }    // dim x n
  [[nodiscard]] const Mat& Weights() const { return weights_; }  // n x 1
  [[nodiscard]] size_type NumPoints() const { return static_cast<size_type>(weights_.size()); }

 private:
  RefEl ref_el_;
  unsigned degree_;
  Mat points_;
  Mat weights_;
};

// Gauss-Legendre on [0,1]: lib/lf/quad/gauss_quadrature.cc:16-67.
// The reference runs this Newton iteration in 57-bit boost::multiprecision and rounds to double; here it runs in
// x87 long double (64-bit mantissa).  Both round a value that is accurate to far better than a double ulp, so the
// doubles agree except for possible 1-ulp differences at rounding ties (documented in DESIGN.md: "GL nodes unpinned
// at the last bit"; the value tolerance of the path is 1e-12).
inline void GaussLegendre(unsigned num_points, Mat& points, Mat& weights) {
  LFO_VERIFY(num_points > 0, "num_points must be positive.");
  using scalar_t = long double;
  const scalar_t kPi = 3.14159265358979323846264338327950288L;
  points = Mat(num_points, 1);
  weights = Mat(num_points, 1);
  const unsigned m = (num_points + 1) / 2;
  for (unsigned i = 0; i < m; ++i) {
    scalar_t z = cosl(kPi * (i + 0.75L) / (num_points + 0.5L));
    scalar_t z1, pp;
    do {
      scalar_t p1 = 1.0L, p2 = 0.0L, p3;
      for (unsigned j = 0; j < num_points; ++j) {
        p3 = p2;
        p2 = p1;
        p1 = ((2.0L * j + 1.0L) * z * p2 - j * p3) / (j + 1.0L);
      }
      pp = num_points * (z * p1 - p2) / (z * z - 1.0L);
      z1 = z;
      z = z1 - p1 / pp;
    } while (fabsl(z - z1) > 1e-17L);
    points[i] = static_cast<double>(0.5L * (1 - z));
    points[num_points - 1 - i] = static_cast<double>(0.5L * (1 + z));
    weights[i] = static_cast<double>(1.0L / ((1.0L - z * z) * pp * pp));
    weights[num_points - 1 - i] = weights[i];
  }
}

namespace detail {
struct TriaRule {
  int degree;
  int npts;
  const double (*data)[3];
};
#define LFO_TRIA_RULE(DEG, N, ...) static const double kTriaRuleData##DEG[N][3] = {__VA_ARGS__};
#include "quad_tria_tables.inc"
#undef LFO_TRIA_RULE
inline const TriaRule* FindTriaRule(unsigned degree) {
  static const TriaRule rules[] = {
      {1, 1, kTriaRuleData1},   {2, 3, kTriaRuleData2},   {4, 6, kTriaRuleData4},    {5, 7, kTriaRuleData5},
      {6, 12, kTriaRuleData6},  {7, 15, kTriaRuleData7},  {8, 16, kTriaRuleData8},   {9, 19, kTriaRuleData9},
      {10, 25, kTriaRuleData10}, {11, 28, kTriaRuleData11}, {12, 33, kTriaRuleData12}};
  for (const auto& r : rules) {
    if (r.degree == static_cast<int>(degree)) return &r;
  }
  return nullptr;
}
}  // namespace detail

// lib/lf/quad/make_quad_rule.cc:21-157
inline QuadRule make_QuadRule(RefEl ref_el, unsigned degree) {
  if (ref_el == RefEl::kSegment()) {
    const unsigned n = degree / 2 + 1;
    Mat p, w;
    GaussLegendre(n, p, w);
    Mat pts(1, n);
    for (unsigned i = 0; i < n; ++i) pts(0, i) = p[i];
    return QuadRule(RefEl::kSegment(), std::move(pts), std::move(w), 2 * n - 1);
  }
  if (ref_el == RefEl::kQuad()) {
    // make_quad_rule.cc:28-37: points2d.row(0) = kron(points1d^T, ones(1,n)); row(1) = points1d^T replicated;
    // weights = kron(w, w)  => point index i*n + j has x0 = p_i, x1 = p_j, weight w_i * w_j
    const unsigned n = degree / 2 + 1;
    Mat p, w;
    GaussLegendre(n, p, w);
    Mat pts(2, n * n), wts(n * n, 1);
    for (unsigned i = 0; i < n; ++i) {
      for (unsigned j = 0; j < n; ++j) {
        pts(0, i * n + j) = p[i];
        pts(1, i * n + j) = p[j];
        wts[i * n + j] = w[i] * w[j];
      }
    }
    return QuadRule(RefEl::kQuad(), std::move(pts), std::move(wts), 2 * n - 1);
  }
  if (ref_el == RefEl::kTria()) {
    // make_quad_rule.cc:44-46: degree 3 silently uses the degree-4 rule
    unsigned d = (degree == 3) ? 4 : degree;
    if (d == 0) d = 1;
    const detail::TriaRule* r = detail::FindTriaRule(d);
    LFO_VERIFY(r != nullptr, "oracle: triangle rule of this degree not tabulated (tables cover 1..12)");
    Mat pts(2, r->npts), wts(r->npts, 1);
    for (int k = 0; k < r->npts; ++k) {
      pts(0, k) = r->data[k][0];
      pts(1, k) = r->data[k][1];
      wts[k] = r->data[k][2];
    }
    return QuadRule(RefEl::kTria(), std::move(pts), std::move(wts), d);
  }
  LFO_VERIFY(false, "No quadrature rule for this reference element");
  return QuadRule();
}

}  // namespace lfo::quad
#endif
