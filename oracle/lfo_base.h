// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the LehrFEM++ assembly hot path, written from the reference's behaviour; no Eigen/Boost.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use anything
// under oracle/.  Every function cites the reference file:line it follows (paths relative to /root/reference).
//
// Parity status: pinned against the reference's own literal goldens (tests/golden/, extracted by
// oracle/tools/extract_reference_data.py) -- see tests/test_oracle_goldens.py.  The reference itself cannot be
// compiled in this image (needs Eigen 3.4 / Boost 1.86 / GTest via Hunter, no network), so oracle/_ref does not exist.
#ifndef LFO_BASE_H
#define LFO_BASE_H

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <span>
#include <stdexcept>
#include <string>
#include <vector>

namespace lfo {

// lib/lf/base/types.h:20-36
using size_type = unsigned int;
using glb_idx_t = unsigned int;
using sub_idx_t = unsigned int;
using dim_t = unsigned int;
constexpr unsigned int kIdxNil = static_cast<unsigned int>(-1);
// lib/lf/assemble/assembly_types.h:22  (Eigen::Index)
using gdof_idx_t = std::int64_t;

// lib/lf/base/lf_exception.h
struct LfException : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// lib/lf/base/lf_assert.h:45-53 -- LF_VERIFY_MSG aborts in the reference; the oracle throws so that tests can see it
#define LFO_VERIFY(expr, msg)                                                              \
  do {                                                                                     \
    if (!(expr)) throw ::lfo::LfException(std::string("LF_VERIFY failed: ") + #expr + ": " + (msg)); \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------
// Minimal stand-in for Eigen::MatrixXd: column-major, heap allocated on every construction (this keeps the
// per-cell allocation cost shape of the reference's Eval(), which returns Eigen::MatrixXd temporaries).
// ---------------------------------------------------------------------------------------------------------------
class Mat {
 public:
  Mat() = default;
  Mat(long r, long c) : r_(r), c_(c), d_(r * c > 0 ? new double[r * c] : nullptr) {}
  Mat(const Mat& o) : r_(o.r_), c_(o.c_), d_(o.size() > 0 ? new double[o.size()] : nullptr) {
    if (size() > 0) std::memcpy(d_.get(), o.d_.get(), sizeof(double) * size());
  }
  Mat(Mat&&) noexcept = default;
  Mat& operator=(const Mat& o) {
    if (this != &o) {
      Mat t(o);
      *this = std::move(t);
    }
    return *this;
  }
  Mat& operator=(Mat&&) noexcept = default;
  [[nodiscard]] long rows() const { return r_; }
  [[nodiscard]] long cols() const { return c_; }
  [[nodiscard]] long size() const { return r_ * c_; }
  double& operator()(long i, long j) { return d_[i + j * r_]; }
  const double& operator()(long i, long j) const { return d_[i + j * r_]; }
  double& operator[](long i) { return d_[i]; }
  const double& operator[](long i) const { return d_[i]; }
  [[nodiscard]] double* data() { return d_.get(); }
  [[nodiscard]] const double* data() const { return d_.get(); }
  void setZero() {
    for (long i = 0; i < size(); ++i) d_[i] = 0.0;
  }
  static Mat Zero(long r, long c) {
    Mat m(r, c);
    m.setZero();
    return m;
  }

 private:
  long r_ = 0, c_ = 0;
  std::unique_ptr<double[]> d_;
};

// ---------------------------------------------------------------------------------------------------------------
// RefEl: lib/lf/base/ref_el.h:31-40 (ids), :127-129 (edge -> endpoint tables), ref_el.cc:10-14 (node coordinates)
// ---------------------------------------------------------------------------------------------------------------
enum class RefElType : unsigned char { kPoint = 1, kSegment = 2, kTria = 3, kQuad = 4 };

class RefEl {
 public:
  constexpr RefEl(RefElType t) : type_(t) {}  // NOLINT
  static constexpr RefEl kPoint() { return RefEl(RefElType::kPoint); }
  static constexpr RefEl kSegment() { return RefEl(RefElType::kSegment); }
  static constexpr RefEl kTria() { return RefEl(RefElType::kTria); }
  static constexpr RefEl kQuad() { return RefEl(RefElType::kQuad); }
  [[nodiscard]] constexpr unsigned Id() const { return static_cast<unsigned>(type_); }
  [[nodiscard]] constexpr dim_t Dimension() const {
    return type_ == RefElType::kPoint ? 0 : (type_ == RefElType::kSegment ? 1 : 2);
  }
  [[nodiscard]] constexpr size_type NumNodes() const {
    return type_ == RefElType::kPoint ? 1 : (type_ == RefElType::kSegment ? 2 : (type_ == RefElType::kTria ? 3 : 4));
  }
  [[nodiscard]] constexpr size_type NumSubEntities(dim_t codim) const {
    if (codim == 0) return 1;
    if (type_ == RefElType::kSegment) return 2;
    if (type_ == RefElType::kTria) return 3;
    if (type_ == RefElType::kQuad) return 4;
    return 0;
  }
  // endpoint `sub_sub` of edge `sub` of a cell (ref_el.h:127-129): edge j joins local vertices (j, j+1 mod nv)
  [[nodiscard]] constexpr sub_idx_t EdgeEndpoint(sub_idx_t edge, sub_idx_t endpoint) const {
    const sub_idx_t nv = NumNodes();
    return (edge + endpoint) % nv;
  }
  // 2 x NumNodes reference node coordinates (ref_el.cc:10-14)
  [[nodiscard]] Mat NodeCoords() const {
    if (type_ == RefElType::kTria) {
      Mat m(2, 3);
      m(0, 0) = 0; m(1, 0) = 0; m(0, 1) = 1; m(1, 1) = 0; m(0, 2) = 0; m(1, 2) = 1;
      return m;
    }
    if (type_ == RefElType::kQuad) {
      Mat m(2, 4);
      m(0, 0) = 0; m(1, 0) = 0; m(0, 1) = 1; m(1, 1) = 0; m(0, 2) = 1; m(1, 2) = 1; m(0, 3) = 0; m(1, 3) = 1;
      return m;
    }
    if (type_ == RefElType::kSegment) {
      Mat m(1, 2);
      m(0, 0) = 0; m(0, 1) = 1;
      return m;
    }
    return Mat(0, 1);
  }
  friend constexpr bool operator==(RefEl a, RefEl b) { return a.type_ == b.type_; }
  friend constexpr bool operator!=(RefEl a, RefEl b) { return a.type_ != b.type_; }
  friend constexpr bool operator<(RefEl a, RefEl b) { return a.Id() < b.Id(); }

 private:
  RefElType type_;
};

}  // namespace lfo
#endif
