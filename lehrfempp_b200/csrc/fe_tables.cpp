// Host-side tabulation of reference-element data for the numeric pass (product code).
//
// Stands in for (reference, paths relative to lib/lf/):
//   uscalfe/lagr_fe.h:56-1480            FeLagrangeO{1,2,3}{Tria,Quad}: shape functions, gradients, local ordering
//   uscalfe/precomputed_scalar_reference_finite_element.h:72-78   tabulation at the rule's points
//   quad/make_quad_rule.cc:21-157, quad/quad_rules_tria.cc, quad/gauss_quadrature.cc:16-67   default rules
//
// Unlike the reference (one hand-expanded formula per shape function) the basis is generated from the lattice of
// Lagrange nodes: on the triangle  phi_(l0,l1,l2) = prod_d prod_{m<l_d} (p*lambda_d - m)/(m+1)  and on the square the
// tensor product of 1D Lagrange polynomials on the nodes m/p.  The LOCAL ORDER of the functions is the reference's:
// vertices, then for each local edge j (from vertex j to vertex j+1) its interior nodes along that direction, then the
// cell-interior nodes (lagr_fe.h:1029-1040 for the cubic triangle, index maps :913-924 and :1461-1479 for the squares).
#include <cmath>
#include <cstring>

#include "lfgpu_internal.cuh"

namespace lfgpu {
namespace {

// lattice coordinates (i, j) of the node of local shape function k; triangle: barycentric (p-i-j, i, j)/p
const int kTriaLattice[3][10][2] = {
    {{0, 0}, {1, 0}, {0, 1}},
    {{0, 0}, {2, 0}, {0, 2}, {1, 0}, {1, 1}, {0, 1}},
    {{0, 0}, {3, 0}, {0, 3}, {1, 0}, {2, 0}, {2, 1}, {1, 2}, {0, 2}, {0, 1}, {1, 1}}};
const int kQuadLattice[3][16][2] = {
    {{0, 0}, {1, 0}, {1, 1}, {0, 1}},
    {{0, 0}, {2, 0}, {2, 2}, {0, 2}, {1, 0}, {2, 1}, {1, 2}, {0, 1}, {1, 1}},
    {{0, 0}, {3, 0}, {3, 3}, {0, 3}, {1, 0}, {2, 0}, {3, 1}, {3, 2}, {2, 3}, {1, 3}, {0, 2}, {0, 1}, {1, 1}, {2, 1}, {2, 2}, {1, 2}}};

// f(l, t) = prod_{m<l} (p t - m)/(m+1) and its derivative in t
void bary_factor(int p, int l, double t, double* f, double* df) {
  double val = 1.0, der = 0.0;
  for (int m = 0; m < l; ++m) {
    const double term = (p * t - m) / (m + 1.0);
    const double dterm = p / (m + 1.0);
    der = der * term + val * dterm;
    val *= term;
  }
  *f = val;
  *df = der;
}

// 1D Lagrange polynomial L_m on nodes n/p, n = 0..p, and derivative
void lagrange_1d(int p, int m, double t, double* f, double* df) {
  double val = 1.0, der = 0.0;
  for (int n = 0; n <= p; ++n) {
    if (n == m) continue;
    const double term = (p * t - n) / static_cast<double>(m - n);
    const double dterm = p / static_cast<double>(m - n);
    der = der * term + val * dterm;
    val *= term;
  }
  *f = val;
  *df = der;
}

struct TriaRule {
  int degree, npts;
  const double (*data)[3];
};
#define LFO_TRIA_RULE(DEG, N, ...) const double kTriaRuleData##DEG[N][3] = {__VA_ARGS__};
#include "quad_tria_tables.inc"  // numeric data of quad/quad_rules_tria.cc, degrees 1..12
#undef LFO_TRIA_RULE
const TriaRule kTriaRules[] = {{1, 1, kTriaRuleData1},    {2, 3, kTriaRuleData2},    {4, 6, kTriaRuleData4},
                               {5, 7, kTriaRuleData5},    {6, 12, kTriaRuleData6},   {7, 15, kTriaRuleData7},
                               {8, 16, kTriaRuleData8},   {9, 19, kTriaRuleData9},   {10, 25, kTriaRuleData10},
                               {11, 28, kTriaRuleData11}, {12, 33, kTriaRuleData12}};

// Gauss-Legendre nodes/weights on [0,1], ascending.  Roots of P_n by Newton from the Chebyshev guess in long double
// (same limit as the reference's 57-bit iteration; agreement to the last bit is not guaranteed).
void gauss_legendre_01(int n, double* x, double* w) {
  const long double pi = 3.14159265358979323846264338327950288L;
  for (int i = 0; i < (n + 1) / 2; ++i) {
    long double z = cosl(pi * (i + 0.75L) / (n + 0.5L));
    long double dp = 1.0L;
    for (int it = 0; it < 100; ++it) {
      long double p0 = 1.0L, p1 = z;
      for (int k = 2; k <= n; ++k) {
        const long double pk = ((2 * k - 1) * z * p1 - (k - 1) * p0) / k;
        p0 = p1;
        p1 = pk;
      }
      if (n == 0) p1 = 1.0L;
      // derivative: P_n'(z) = n (z P_n - P_{n-1}) / (z^2 - 1)
      dp = n * (z * p1 - p0) / (z * z - 1.0L);
      const long double dz = p1 / dp;
      z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    // recompute derivative at the converged root for the weight
    {
      long double p0 = 1.0L, p1 = z;
      for (int k = 2; k <= n; ++k) {
        const long double pk = ((2 * k - 1) * z * p1 - (k - 1) * p0) / k;
        p0 = p1;
        p1 = pk;
      }
      dp = n * (z * p1 - p0) / (z * z - 1.0L);
    }
    const long double wt = 1.0L / ((1.0L - z * z) * dp * dp);
    x[i] = static_cast<double>(0.5L * (1.0L - z));
    x[n - 1 - i] = static_cast<double>(0.5L * (1.0L + z));
    w[i] = static_cast<double>(wt);
    w[n - 1 - i] = w[i];
  }
}

}  // namespace

int nsf_of(int degree, int cell_type) {
  static const int t[3] = {3, 6, 10}, q[3] = {4, 9, 16};
  if (degree < 1 || degree > 3) return -1;
  return cell_type == 3 ? t[degree - 1] : q[degree - 1];
}

int default_quad_rule(int cell_type, int degree, int capacity, double* points, double* weights) {
  if (cell_type == 3) {
    int d = degree == 3 ? 4 : degree;  // make_quad_rule.cc:44-46
    if (d == 0) d = 1;
    for (const auto& r : kTriaRules) {
      if (r.degree == d) {
        if (points != nullptr && weights != nullptr) {
          if (r.npts > capacity) return LFGPU_ERR_INVALID;
          for (int k = 0; k < r.npts; ++k) {
            points[k] = r.data[k][0];
            points[r.npts + k] = r.data[k][1];
            weights[k] = r.data[k][2];
          }
        }
        return r.npts;
      }
    }
    return LFGPU_ERR_MISSING_RULE;
  }
  if (cell_type == 4) {
    const int n = degree / 2 + 1;  // make_quad_rule.cc:29
    if (points != nullptr && weights != nullptr) {
      if (n * n > capacity || n > 64) return LFGPU_ERR_INVALID;
      double x[64], w[64];
      gauss_legendre_01(n, x, w);
      for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {  // point i*n + j = (x_i, x_j), weight w_i w_j (make_quad_rule.cc:31-37)
          points[i * n + j] = x[i];
          points[n * n + i * n + j] = x[j];
          weights[i * n + j] = w[i] * w[j];
        }
      }
    }
    return n * n;
  }
  if (cell_type == 2) {  // segment: make_quad_rule.cc:22-27
    const int n = degree / 2 + 1;
    if (points != nullptr && weights != nullptr) {
      if (n > capacity || n > 64) return LFGPU_ERR_INVALID;
      gauss_legendre_01(n, points, weights);
    }
    return n;
  }
  return LFGPU_ERR_INVALID;
}

int build_fe_table(int degree, int cell_type, const lfgpu_quad* qr, FeTable* out, std::string* err) {
  const int nsf = nsf_of(degree, cell_type);
  if (nsf < 0 || (cell_type != 3 && cell_type != 4)) {
    if (err) *err = "degree must be 1..3 and cell_type 3 (tria) or 4 (quad)";
    return LFGPU_ERR_INVALID;
  }
  *out = FeTable{};
  out->nsf = nsf;
  double pts[2 * kMaxNq], wts[kMaxNq];
  int nq;
  if (qr != nullptr) {
    nq = qr->n;
    if (nq < 1 || nq > kMaxNq || qr->points == nullptr || qr->weights == nullptr) {
      if (err) *err = "quadrature rule must have 1.." + std::to_string(kMaxNq) + " points";
      return LFGPU_ERR_INVALID;
    }
    std::memcpy(pts, qr->points, sizeof(double) * 2 * nq);
    std::memcpy(wts, qr->weights, sizeof(double) * nq);
  } else {
    nq = default_quad_rule(cell_type, 2 * degree, kMaxNq, pts, wts);  // loc_comp_ellbvp.h:227-228
    if (nq < 0) {
      if (err) *err = "no default quadrature rule";
      return nq;
    }
  }
  out->nq = nq;
  const int p = degree;
  for (int k = 0; k < nq; ++k) {
    const double x0 = pts[k], x1 = pts[nq + k];
    out->w[k] = wts[k];
    out->qx[k] = x0;
    out->qy[k] = x1;
    for (int a = 0; a < nsf; ++a) {
      double v, d0, d1;
      if (cell_type == 3) {
        const int l1 = kTriaLattice[p - 1][a][0], l2 = kTriaLattice[p - 1][a][1], l0 = p - l1 - l2;
        double f0, g0, f1, g1, f2, g2;
        bary_factor(p, l0, 1.0 - x0 - x1, &f0, &g0);
        bary_factor(p, l1, x0, &f1, &g1);
        bary_factor(p, l2, x1, &f2, &g2);
        v = f0 * f1 * f2;
        d0 = -g0 * f1 * f2 + f0 * g1 * f2;
        d1 = -g0 * f1 * f2 + f0 * f1 * g2;
      } else {
        const int ix = kQuadLattice[p - 1][a][0], iy = kQuadLattice[p - 1][a][1];
        double fx, gx, fy, gy;
        lagrange_1d(p, ix, x0, &fx, &gx);
        lagrange_1d(p, iy, x1, &fy, &gy);
        v = fx * fy;
        d0 = gx * fy;
        d1 = fx * gy;
      }
      out->phi[a * nq + k] = v;
      out->gx[a * nq + k] = d0;
      out->gy[a * nq + k] = d1;
    }
  }
  return 0;
}

// FeLagrangeO{1,2,3}Segment (uscalfe/lagr_fe.h:243-305, 451-528, 839-931) tabulated at a rule on [0,1]: local shape
// functions 0, 1 belong to the endpoints, 2.. to the interior nodes in ascending position.  qr: points[n] (only row 0 is
// read), NULL = make_QuadRule(kSegment, 2 * degree) = Gauss-Legendre with degree + 1 points (make_quad_rule.cc:22-27)
int build_segment_table(int degree, const lfgpu_quad* qr, SegTable* out, std::string* err) {
  if (degree < 1 || degree > 3) {
    if (err) *err = "degree must be 1, 2 or 3";
    return LFGPU_ERR_INVALID;
  }
  *out = SegTable{};
  const int p = degree;
  out->nsf = p + 1;
  if (qr != nullptr) {
    if (qr->n < 1 || qr->n > kMaxSegNq || qr->points == nullptr || qr->weights == nullptr) {
      if (err) *err = "segment quadrature rule must have 1.." + std::to_string(kMaxSegNq) + " points";
      return LFGPU_ERR_INVALID;
    }
    out->nq = qr->n;
    std::memcpy(out->x, qr->points, sizeof(double) * qr->n);
    std::memcpy(out->w, qr->weights, sizeof(double) * qr->n);
  } else {
    out->nq = (2 * p) / 2 + 1;
    gauss_legendre_01(out->nq, out->x, out->w);
  }
  for (int a = 0; a <= p; ++a) {
    const int m = a == 0 ? 0 : (a == 1 ? p : a - 1);  // lattice node of local shape function a
    for (int k = 0; k < out->nq; ++k) {
      double f, df;
      lagrange_1d(p, m, out->x[k], &f, &df);
      out->phi[a * kMaxSegNq + k] = f;
    }
  }
  return LFGPU_OK;
}

void build_fe_tensors(const FeTable& t, FeTensors* out) {
  std::memset(out, 0, sizeof(FeTensors));
  const int nsf = t.nsf, nq = t.nq;
  for (int a = 0; a < nsf; ++a) {
    for (int b = 0; b < nsf; ++b) {
      double k00 = 0, k01 = 0, k10 = 0, k11 = 0, m = 0;
      for (int k = 0; k < nq; ++k) {
        const double w = t.w[k];
        k00 += w * t.gx[a * nq + k] * t.gx[b * nq + k];
        k01 += w * t.gx[a * nq + k] * t.gy[b * nq + k];
        k10 += w * t.gy[a * nq + k] * t.gx[b * nq + k];
        k11 += w * t.gy[a * nq + k] * t.gy[b * nq + k];
        m += w * t.phi[a * nq + k] * t.phi[b * nq + k];
      }
      out->k00[a * nsf + b] = k00;
      out->k01[a * nsf + b] = k01;
      out->k10[a * nsf + b] = k10;
      out->k11[a * nsf + b] = k11;
      out->m[a * nsf + b] = m;
    }
    double l = 0;
    for (int k = 0; k < nq; ++k) l += t.w[k] * t.phi[a * nq + k];
    out->l[a] = l;
  }
}

}  // namespace lfgpu

extern "C" int lfgpu_fe_tabulate(int degree, int cell_type, const lfgpu_quad* qr, double* phi, double* grad) {
  lfgpu::FeTable t;
  std::string err;
  const int rc = lfgpu::build_fe_table(degree, cell_type, qr, &t, &err);
  if (rc < 0) {
    lfgpu::set_last_error(nullptr, err);
    return rc;
  }
  for (int a = 0; a < t.nsf; ++a) {
    for (int k = 0; k < t.nq; ++k) {
      if (phi) phi[a * t.nq + k] = t.phi[a * t.nq + k];
      if (grad) {
        grad[a * 2 * t.nq + 2 * k] = t.gx[a * t.nq + k];
        grad[a * 2 * t.nq + 2 * k + 1] = t.gy[a * t.nq + k];
      }
    }
  }
  return t.nsf;
}

extern "C" int lfgpu_default_quad_rule(int cell_type, int degree, int capacity, double* points, double* weights) {
  return lfgpu::default_quad_rule(cell_type, degree, capacity, points, weights);
}
