/*
 * tloam_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT). See tloam_oracle.h.
 *
 * PARITY UNPINNED (no reference tests / golden vectors exist for this path; the reference cannot be
 * built here).  Every function cites the reference lines it restates.  "ref:" paths are relative to
 * /root/reference; "sophus:" = include/third_party/sophus.
 *
 * Third-party behaviour restated from published sources (not present under /root/reference):
 *   - Ceres Solver 2.0 (README.md:111): trust_region_minimizer.cc, dogleg_strategy.cc,
 *     trust_region_step_evaluator.cc, corrector.cc, loss_function.cc (CauchyLoss), residual_block.cc,
 *     dense_qr_solver.cc -- call sites ref: registration.cpp:970-974, 1036-1047.
 *   - Open3D 0.12 KDTreeFlann::SearchHybrid (nanoflann knnSearch + truncation at d2 < r^2)
 *     -- call sites ref: registration.cpp:272,444,535,588,731.
 *   - Eigen 3: Quaternion(Matrix3) (Shoemake), Quaternion::toRotationMatrix, _transformVector,
 *     SelfAdjointEigenSolver (restated as a cyclic Jacobi iteration, same result to ~1 ulp).
 */
#include "tloam_oracle.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <numeric>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using std::size_t;
constexpr double kSophusEps = 1e-10;  // sophus: common.hpp:94
constexpr double kPi = 3.14159265358979323846;

struct V3 {
  double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }

// ---------------------------------------------------------------------------------------------
// Lie group math, mirroring Sophus (quaternion + translation representation).
// ---------------------------------------------------------------------------------------------
struct SE3 {
  double w, x, y, z;  // unit quaternion
  V3 t;
};

// Eigen Quaternion::toRotationMatrix(), row-major 3x3 out.
void quat_to_R(const SE3& T, double R[9]) {
  const double tx = 2 * T.x, ty = 2 * T.y, tz = 2 * T.z;
  const double twx = tx * T.w, twy = ty * T.w, twz = tz * T.w;
  const double txx = tx * T.x, txy = ty * T.x, txz = tz * T.x;
  const double tyy = ty * T.y, tyz = tz * T.y, tzz = tz * T.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// Eigen Quaternion::_transformVector (what Sophus' SO3 * point calls, sophus: so3.hpp:358-362).
inline V3 quat_rotate(const SE3& T, V3 v) {
  V3 q{T.x, T.y, T.z};
  V3 uv = cross(q, v);
  uv = uv + uv;
  return v + T.w * uv + cross(q, uv);
}
inline V3 se3_act(const SE3& T, V3 p) { return quat_rotate(T, p) + T.t; }  // sophus: se3.hpp:321-324

// sophus: so3.hpp:583-619 (expAndTheta) + se3.hpp:761-783 (exp).
SE3 se3_exp(const double a[6]) {
  const V3 ups{a[0], a[1], a[2]};
  const V3 om{a[3], a[4], a[5]};
  const double theta_sq = dot(om, om);
  double theta, imag, real;
  if (theta_sq < kSophusEps * kSophusEps) {
    theta = 0.0;
    const double theta_po4 = theta_sq * theta_sq;
    imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    theta = std::sqrt(theta_sq);
    const double half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  SE3 T;
  T.w = real; T.x = imag * om.x; T.y = imag * om.y; T.z = imag * om.z;
  // V = left Jacobian (se3.hpp:772-780); Omega = hat(omega)
  double V[9];
  if (theta < kSophusEps) {
    quat_to_R(T, V);  // "V = so3.matrix()"
  } else {
    const double O[9] = {0, -om.z, om.y, om.z, 0, -om.x, -om.y, om.x, 0};
    double O2[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += O[i * 3 + k] * O[k * 3 + j];
        O2[i * 3 + j] = s;
      }
    const double c1 = (1.0 - std::cos(theta)) / theta_sq;
    const double c2 = (theta - std::sin(theta)) / (theta_sq * theta);
    for (int i = 0; i < 9; ++i) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * O[i] + c2 * O2[i];
  }
  T.t = {V[0] * ups.x + V[1] * ups.y + V[2] * ups.z, V[3] * ups.x + V[4] * ups.y + V[5] * ups.z,
         V[6] * ups.x + V[7] * ups.y + V[8] * ups.z};
  return T;
}

// sophus: so3.hpp:247-290 (logAndTheta) + se3.hpp:223-257 (log).
void se3_log(const SE3& T, double out[6]) {
  const double squared_n = T.x * T.x + T.y * T.y + T.z * T.z;
  const double w = T.w;
  double two_atan_nbyw_by_n, theta;
  if (squared_n < kSophusEps * kSophusEps) {
    const double squared_w = w * w;
    two_atan_nbyw_by_n = 2.0 / w - (2.0 / 3.0) * squared_n / (w * squared_w);
    theta = 2.0 * squared_n / w;
  } else {
    const double n = std::sqrt(squared_n);
    if (std::fabs(w) < kSophusEps) {
      two_atan_nbyw_by_n = (w > 0.0 ? kPi : -kPi) / n;
    } else {
      two_atan_nbyw_by_n = 2.0 * std::atan(n / w) / n;
    }
    theta = two_atan_nbyw_by_n * n;
  }
  const V3 om{two_atan_nbyw_by_n * T.x, two_atan_nbyw_by_n * T.y, two_atan_nbyw_by_n * T.z};
  const double O[9] = {0, -om.z, om.y, om.z, 0, -om.x, -om.y, om.x, 0};
  double O2[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += O[i * 3 + k] * O[k * 3 + j];
      O2[i * 3 + j] = s;
    }
  double Vinv[9];
  if (std::fabs(theta) < kSophusEps) {
    for (int i = 0; i < 9; ++i) Vinv[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * O[i] + (1.0 / 12.0) * O2[i];
  } else {
    const double half = 0.5 * theta;
    const double c = (1.0 - theta * std::cos(half) / (2.0 * std::sin(half))) / (theta * theta);
    for (int i = 0; i < 9; ++i) Vinv[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * O[i] + c * O2[i];
  }
  out[0] = Vinv[0] * T.t.x + Vinv[1] * T.t.y + Vinv[2] * T.t.z;
  out[1] = Vinv[3] * T.t.x + Vinv[4] * T.t.y + Vinv[5] * T.t.z;
  out[2] = Vinv[6] * T.t.x + Vinv[7] * T.t.y + Vinv[8] * T.t.z;
  out[3] = om.x; out[4] = om.y; out[5] = om.z;
}

// sophus: so3.hpp:325-340 (quaternion product) + SO3(quaternion) ctor normalisation (so3.hpp:297-303,
// 480-487) + se3.hpp:304-312 (translation + R * other.translation).
SE3 se3_mul(const SE3& a, const SE3& b) {
  SE3 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  const double len = std::sqrt(r.w * r.w + r.x * r.x + r.y * r.y + r.z * r.z);
  r.w /= len; r.x /= len; r.y /= len; r.z /= len;
  r.t = a.t + quat_rotate(a, b.t);
  return r;
}

// Sophus::SE3d(Matrix4) (sophus: se3.hpp:497-503) -> SO3(R) (so3.hpp:469-474) -> Eigen Quaternion(R)
// (Shoemake's algorithm, Eigen/src/Geometry/Quaternion.h). T is 4x4 COLUMN-major (Eigen::Isometry3d).
SE3 se3_from_matrix(const double T[16]) {
  auto m = [&](int r, int c) { return T[c * 4 + r]; };
  SE3 q;
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m(2, 1) - m(1, 2)) * t;
    q.y = (m(0, 2) - m(2, 0)) * t;
    q.z = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m(k, j) - m(j, k)) * t;
    v[j] = (m(j, i) + m(i, j)) * t;
    v[k] = (m(k, i) + m(i, k)) * t;
    q.x = v[0]; q.y = v[1]; q.z = v[2];
  }
  q.t = {m(0, 3), m(1, 3), m(2, 3)};
  return q;
}

void se3_to_matrix(const SE3& S, double T[16]) {
  double R[9];
  quat_to_R(S, R);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T[c * 4 + r] = R[r * 3 + c];
    T[r * 4 + 3] = 0.0;  // bottom row (row 3 of columns 0..2)
  }
  T[12] = S.t.x; T[13] = S.t.y; T[14] = S.t.z; T[15] = 1.0;
}

// PoseSE3Parameterization::Plus, ref: registration.cpp:162-173: x+ = log(exp(delta) * exp(x)).
void se3_plus(const double x[6], const double delta[6], double out[6]) {
  se3_log(se3_mul(se3_exp(delta), se3_exp(x)), out);
}

// ---------------------------------------------------------------------------------------------
// 3x3 symmetric eigen-decomposition (stands in for Eigen::SelfAdjointEigenSolver<Matrix3d>::compute,
// ref: registration.cpp:476-479): cyclic Jacobi to machine precision, eigenvalues ascending,
// eigenvectors as columns (col-major 3x3 out).
// ---------------------------------------------------------------------------------------------
void sym_eig3(const double cov[9], double eig[3], double vec[9]) {
  double a[3][3], v[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      a[i][j] = cov[i * 3 + j];
      v[i][j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A * G
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- G^T * A
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int order[3] = {0, 1, 2};
  std::sort(order, order + 3, [&](int i, int j) { return a[i][i] < a[j][j]; });
  for (int c = 0; c < 3; ++c) {
    eig[c] = a[order[c]][order[c]];
    for (int r = 0; r < 3; ++r) vec[c * 3 + r] = v[r][order[c]];
  }
}

// ---------------------------------------------------------------------------------------------
// Exact KD-tree (stands in for open3d::geometry::KDTreeFlann = nanoflann, leaf size 15).
// SearchHybrid(query, radius, max_nn): k nearest, ascending squared distance, truncated to d2 < r^2.
// Ties broken by smaller point index (measure-zero for continuous data).
// ---------------------------------------------------------------------------------------------
struct KDTree {
  struct Node {
    int lo, hi;         // point range [lo,hi) in perm (leaf)
    int left, right;    // children (-1 for leaf)
    int dim;
    double split_lo, split_hi;  // max of left side / min of right side along dim
  };
  const double* pts = nullptr;  // AoS xyz
  size_t n = 0;
  std::vector<int> perm;
  std::vector<Node> nodes;
  double bb_lo[3], bb_hi[3];
  static constexpr int kLeaf = 15;

  void build(const double* p, size_t count) {
    pts = p; n = count;
    perm.resize(n);
    std::iota(perm.begin(), perm.end(), 0);
    nodes.clear();
    nodes.reserve(2 * (n / kLeaf + 2));
    for (int d = 0; d < 3; ++d) { bb_lo[d] = std::numeric_limits<double>::infinity(); bb_hi[d] = -bb_lo[d]; }
    for (size_t i = 0; i < n; ++i)
      for (int d = 0; d < 3; ++d) {
        bb_lo[d] = std::min(bb_lo[d], pts[3 * i + d]);
        bb_hi[d] = std::max(bb_hi[d], pts[3 * i + d]);
      }
    if (n > 0) {
      build_rec(0, static_cast<int>(n));
    }
  }

  int build_rec(int lo, int hi) {
    const int id = static_cast<int>(nodes.size());
    nodes.push_back(Node{lo, hi, -1, -1, 0, 0, 0});
    if (hi - lo <= kLeaf) return id;
    // exact bounds of this subset, widest dimension, split at the median
    double mn[3], mx[3];
    for (int d = 0; d < 3; ++d) { mn[d] = std::numeric_limits<double>::infinity(); mx[d] = -mn[d]; }
    for (int i = lo; i < hi; ++i)
      for (int d = 0; d < 3; ++d) {
        const double v = pts[3 * perm[i] + d];
        mn[d] = std::min(mn[d], v); mx[d] = std::max(mx[d], v);
      }
    int dim = 0;
    for (int d = 1; d < 3; ++d) if (mx[d] - mn[d] > mx[dim] - mn[dim]) dim = d;
    if (mx[dim] == mn[dim]) return id;  // all points identical: keep as (big) leaf
    const int mid = lo + (hi - lo) / 2;
    std::nth_element(perm.begin() + lo, perm.begin() + mid, perm.begin() + hi,
                     [&](int a, int b) { return pts[3 * a + dim] < pts[3 * b + dim]; });
    double left_max = -std::numeric_limits<double>::infinity();
    for (int i = lo; i < mid; ++i) left_max = std::max(left_max, pts[3 * perm[i] + dim]);
    const double right_min = pts[3 * perm[mid] + dim];
    nodes[id].dim = dim;
    nodes[id].split_lo = left_max;
    nodes[id].split_hi = right_min;
    const int l = build_rec(lo, mid);
    const int r = build_rec(mid, hi);
    nodes[id].left = l; nodes[id].right = r;
    return id;
  }

  struct Result {
    int k; int count; double worst_d2; int worst_idx;
    int* idx; double* d2;
    inline void insert(double d, int i) {
      int pos = (count < k) ? count : k - 1;
      while (pos > 0 && (d2[pos - 1] > d || (d2[pos - 1] == d && idx[pos - 1] > i))) {
        d2[pos] = d2[pos - 1]; idx[pos] = idx[pos - 1]; --pos;
      }
      d2[pos] = d; idx[pos] = i;
      if (count < k) ++count;
      if (count == k) { worst_d2 = d2[k - 1]; worst_idx = idx[k - 1]; }
    }
  };

  // returns count; idx/d2 have room for k entries.
  int search_hybrid(const double q[3], double radius, int k, int* idx, double* d2) const {
    if (n == 0 || k <= 0) return 0;
    Result res{k, 0, radius * radius, std::numeric_limits<int>::max(), idx, d2};
    // distance from q to the root bounding box
    double off[3], mind2 = 0;
    for (int d = 0; d < 3; ++d) {
      off[d] = 0;
      if (q[d] < bb_lo[d]) off[d] = (q[d] - bb_lo[d]) * (q[d] - bb_lo[d]);
      if (q[d] > bb_hi[d]) off[d] = (q[d] - bb_hi[d]) * (q[d] - bb_hi[d]);
      mind2 += off[d];
    }
    // strict d2 < r^2 (std::lower_bound truncation in KDTreeFlann::SearchHybrid)
    search_rec_strict(0, q, mind2, off, res, radius * radius);
    return res.count;
  }

  void search_rec_strict(int id, const double q[3], double mind2, double off[3], Result& res, double r2) const {
    const Node& nd = nodes[id];
    if (nd.left < 0) {
      for (int i = nd.lo; i < nd.hi; ++i) {
        const int pi = perm[i];
        const double dx = pts[3 * pi] - q[0], dy = pts[3 * pi + 1] - q[1], dz = pts[3 * pi + 2] - q[2];
        const double d = std::fma(dz, dz, std::fma(dy, dy, dx * dx));   // spelled out: the GPU uses the same three operations
        if (d >= r2) continue;
        if (res.count < res.k || d < res.worst_d2 || (d == res.worst_d2 && pi < res.worst_idx)) res.insert(d, pi);
      }
      return;
    }
    const int dim = nd.dim;
    const double v = q[dim];
    const double d_lo = v - nd.split_lo, d_hi = v - nd.split_hi;
    int best, other; double cut;
    if (d_lo + d_hi < 0) { best = nd.left; other = nd.right; cut = (d_hi < 0) ? d_hi * d_hi : 0.0; }
    else { best = nd.right; other = nd.left; cut = (d_lo > 0) ? d_lo * d_lo : 0.0; }
    search_rec_strict(best, q, mind2, off, res, r2);
    const double saved = off[dim];
    const double new_min = mind2 - saved + cut;
    const double bound = (res.count < res.k) ? r2 : res.worst_d2;
    if (new_min <= bound) {
      off[dim] = cut;
      search_rec_strict(other, q, new_min, off, res, r2);
      off[dim] = saved;
    }
  }
};

int brute_hybrid(const double* pts, size_t n, const double q[3], double radius, int k, int* idx, double* d2) {
  KDTree::Result res{k, 0, radius * radius, std::numeric_limits<int>::max(), idx, d2};
  const double r2 = radius * radius;
  for (size_t i = 0; i < n; ++i) {
    const double dx = pts[3 * i] - q[0], dy = pts[3 * i + 1] - q[1], dz = pts[3 * i + 2] - q[2];
    const double d = std::fma(dz, dz, std::fma(dy, dy, dx * dx));   // spelled out: the GPU uses the same three operations
    if (d >= r2) continue;
    const int pi = static_cast<int>(i);
    if (res.count < res.k || d < res.worst_d2 || (d == res.worst_d2 && pi < res.worst_idx)) res.insert(d, pi);
  }
  return res.count;
}

// ---------------------------------------------------------------------------------------------
// Primitive fits.
// ---------------------------------------------------------------------------------------------
// LocalRegistration::fitBestPlane, ref: registration.cpp:303-368.
void fit_best_plane(const V3* pts, int n, double out[4]) {
  if (n <= 0) { out[0] = out[1] = out[2] = out[3] = 0; return; }
  const double total = static_cast<double>(n);
  V3 c{0, 0, 0};
  for (int i = 0; i < n; ++i) c = c + pts[i];
  c = {c.x / total, c.y / total, c.z / total};
  double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
  for (int i = 0; i < n; ++i) {
    const V3 d = pts[i] - c;
    xx += d.x * d.x; xy += d.x * d.y; xz += d.x * d.z;
    yy += d.y * d.y; yz += d.y * d.z; zz += d.z * d.z;
  }
  xx /= total; xy /= total; xz /= total; yy /= total; yz /= total; zz /= total;
  V3 wd{0, 0, 0};
  {
    const double det_x = yy * zz - yz * yz;
    const V3 axis{det_x, xz * yz - xy * zz, xy * yz - xz * yy};
    double w = det_x * det_x;
    if (dot(wd, axis) < 0.0) w = -w;
    wd = wd + w * axis;
  }
  {
    const double det_y = xx * zz - xz * xz;
    const V3 axis{xz * yz - xy * zz, det_y, xy * xz - yz * xx};
    double w = det_y * det_y;
    if (dot(wd, axis) < 0.0) w = -w;
    wd = wd + w * axis;
  }
  {
    const double det_z = xx * yy - xy * xy;
    const V3 axis{xy * yz - xz * yy, xy * xz - yz * xx, det_z};
    double w = det_z * det_z;
    if (dot(wd, axis) < 0.0) w = -w;
    wd = wd + w * axis;
  }
  const double nn = norm(wd);
  if (nn == 0) { out[0] = out[1] = out[2] = out[3] = 0; return; }
  wd = {wd.x / nn, wd.y / nn, wd.z / nn};
  out[0] = wd.x; out[1] = wd.y; out[2] = wd.z; out[3] = -dot(wd, c);
}

// Line fit + test of addEdgeCostFactor, ref: registration.cpp:451-485.
int fit_line(const V3* pts, int n, double dir_thres, V3* a, V3* b, V3* mean_out, V3* dir_out, double eig_out[3]) {
  double cum[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    const V3 p = pts[i];
    cum[0] += p.x; cum[1] += p.y; cum[2] += p.z;
    cum[3] += p.x * p.x; cum[4] += p.x * p.y; cum[5] += p.x * p.z;
    cum[6] += p.y * p.y; cum[7] += p.y * p.z; cum[8] += p.z * p.z;
  }
  for (double& c : cum) c /= static_cast<double>(n);
  double cov[9];
  cov[0] = cum[3] - cum[0] * cum[0];
  cov[4] = cum[6] - cum[1] * cum[1];
  cov[8] = cum[8] - cum[2] * cum[2];
  cov[1] = cov[3] = cum[4] - cum[0] * cum[1];
  cov[2] = cov[6] = cum[5] - cum[0] * cum[2];
  cov[5] = cov[7] = cum[7] - cum[1] * cum[2];
  double eig[3], vec[9];
  sym_eig3(cov, eig, vec);
  const V3 dir{vec[6], vec[7], vec[8]};  // eigenvectors().col(2)
  const V3 mean{cum[0], cum[1], cum[2]};
  *a = mean + 0.1 * dir;
  *b = mean + (-0.1) * dir;
  if (mean_out) *mean_out = mean;
  if (dir_out) *dir_out = dir;
  if (eig_out) { eig_out[0] = eig[0]; eig_out[1] = eig[1]; eig_out[2] = eig[2]; }
  return (eig[2] > 3 * eig[1] && std::fabs(dir.z) > dir_thres) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Cost functors. J is row-major (rows x 6), columns = (translation 3, rotation 3).
// ---------------------------------------------------------------------------------------------
// PointToPointErr::Evaluate, ref: registration.cpp:19-47.
inline void eval_p2p(const SE3& T, V3 p, V3 q, double w, double r[3], double* J, double* slot) {
  const V3 s = se3_act(T, p);
  const V3 res = q - s;
  r[0] = res.x * w; r[1] = res.y * w; r[2] = res.z * w;
  *slot = (r[0] + r[1] + r[2]) * (r[0] + r[1] + r[2]);  // std::pow(sum, 2), Q3
  if (J) {
    // [-I*w | hat(s)*w]
    const double h[9] = {0, -s.z, s.y, s.z, 0, -s.x, -s.y, s.x, 0};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        J[i * 6 + j] = (i == j) ? -w : 0.0 * w;
        J[i * 6 + 3 + j] = h[i * 3 + j] * w;
      }
  }
}

// PointToLineErr::Evaluate, ref: registration.cpp:55-88.
inline void eval_p2l(const SE3& T, V3 p, V3 a, V3 b, double w, double r[3], double* J, double* slot) {
  const V3 c = se3_act(T, p);
  const V3 nu = cross(c - a, c - b);
  const V3 de = a - b;
  const double den = norm(de);
  r[0] = nu.x / den * w; r[1] = nu.y / den * w; r[2] = nu.z / den * w;
  *slot = (r[0] + r[1] + r[2]) * (r[0] + r[1] + r[2]);
  if (J) {
    // skew(b - a) * [I*w | -hat(c)*w] / |a-b|
    double D[18];
    const double h[9] = {0, -c.z, c.y, c.z, 0, -c.x, -c.y, c.x, 0};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        D[i * 6 + j] = (i == j) ? w : 0.0;
        D[i * 6 + 3 + j] = -h[i * 3 + j] * w;
      }
    const V3 re = b - a;
    const double S[9] = {0, -re.z, re.y, re.z, 0, -re.x, -re.y, re.x, 0};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 6; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += S[i * 3 + k] * D[k * 6 + j];
        J[i * 6 + j] = s / den;
      }
  }
}

// PointToPlaneErr::Evaluate, ref: registration.cpp:96-117. NOTE residual is NOT weighted (Q4).
inline void eval_p2pl(const SE3& T, V3 p, V3 n, double d, double w, double r[1], double* J, double* slot) {
  const V3 c = se3_act(T, p);
  r[0] = dot(n, c) + d;
  *slot = r[0] * r[0];
  if (J) {
    // n^T * [I*w | -hat(c)*w]
    const double h[9] = {0, -c.z, c.y, c.z, 0, -c.x, -c.y, c.x, 0};
    const double nv[3] = {n.x, n.y, n.z};
    for (int j = 0; j < 3; ++j) {
      J[j] = nv[j] * w;
      double s = 0;
      for (int k = 0; k < 3; ++k) s += nv[k] * (-h[k * 3 + j] * w);
      J[3 + j] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The registration object.
// ---------------------------------------------------------------------------------------------
enum Cloud { kEdge = 0, kSphere = 1, kPlanar = 2, kGround = 3 };

struct Factor {
  int cloud; int index;  // feature index inside its cloud
  double prim[6];        // plane: n,d ; line: a,b ; point: q
  double w;              // weight captured at build time (functor member)
};

struct Oracle {
  oracle_config cfg;
  std::vector<double> src[4], tgt[4];
  KDTree tree[4];
  std::vector<double> weights[4], slots[4];
  std::vector<Factor> factors;           // residual-block order: planar, ground, edge, sphere
  std::vector<int> row_offset;           // first residual row per factor
  int num_rows = 0;
  double params[6] = {0, 0, 0, 0, 0, 0}; // (translation, rotation), ref: registration.hpp:327-329
  double curr_pose[16], last_pose[16];
  Oracle() { for (int i = 0; i < 16; ++i) curr_pose[i] = last_pose[i] = (i % 5 == 0) ? 1.0 : 0.0; }

  int nthreads() const {
#ifdef _OPENMP
    return cfg.num_threads > 0 ? cfg.num_threads : omp_get_max_threads();
#else
    return 1;
#endif
  }
  double radius_of(int c) const {
    return c == kEdge ? cfg.edge_dist_thres : c == kSphere ? cfg.sphere_dist_thres
         : c == kPlanar ? cfg.planar_dist_thres : cfg.ground_dist_thres;
  }
  int maxnum_of(int c) const {
    return c == kEdge ? cfg.edge_maxnum : c == kSphere ? cfg.sphere_maxnum
         : c == kPlanar ? cfg.planar_maxnum : cfg.ground_maxnum;
  }

  // One feature's correspondence search + primitive fit. Returns:
  //  0 = not a candidate (skipped by a `continue` before the cap test or failed the fit test)
  //  1 = candidate factor (passes all tests; subject to the cap)
  // and `counted` = whether the reference's counter advances / cap test is reached at this feature.
  // ref: registration.cpp:440-493 (edge), 531-551 (sphere), 584-625 (planar), 727-768 (ground).
  int correspond(int c, size_t i, const SE3& T, double prim[6], int* cap_checked, int* counted) const {
    const double* s = src[c].data();
    const V3 p{s[3 * i], s[3 * i + 1], s[3 * i + 2]};
    const V3 pw = se3_act(T, p);
    const double q[3] = {pw.x, pw.y, pw.z};
    int idx[5]; double d2[5];
    const double* tp = tgt[c].data();
    *cap_checked = 0; *counted = 0;
    for (int j = 0; j < 6; ++j) prim[j] = 0;
    if (c == kSphere) {
      const int k = tree[c].search_hybrid(q, radius_of(c), 1, idx, d2);
      if (k > 0) {
        if (d2[0] > 0.2) return 0;       // squared distance vs 0.2 (Q6); `continue` skips the counter
        *cap_checked = 1; *counted = 1;  // cap test at :538 happens here, sphere_sum++ at :551
        prim[0] = tp[3 * idx[0]]; prim[1] = tp[3 * idx[0] + 1]; prim[2] = tp[3 * idx[0] + 2];
        return 1;
      }
      *counted = 1;                      // not found: sphere_sum++ still executes (:551)
      return 0;
    }
    const int k = tree[c].search_hybrid(q, radius_of(c), 5, idx, d2);
    if (k <= 0) return 0;
    V3 nb[5];
    for (int j = 0; j < k; ++j) nb[j] = {tp[3 * idx[j]], tp[3 * idx[j] + 1], tp[3 * idx[j] + 2]};
    if (c == kEdge) {
      if (k <= 3) return 0;              // :445
      *cap_checked = 1;                  // :448
      V3 a, b;
      if (!fit_line(nb, k, cfg.edge_dir_thres, &a, &b, nullptr, nullptr, nullptr)) return 0;
      prim[0] = a.x; prim[1] = a.y; prim[2] = a.z; prim[3] = b.x; prim[4] = b.y; prim[5] = b.z;
      *counted = 1;                      // edge_num++ only when a factor is added (:492)
      return 1;
    }
    // planar / ground
    if (k <= 4) return 0;                // :589 / :732
    *cap_checked = 1;                    // :592 / :735
    double nd[4];
    fit_best_plane(nb, k, nd);
    const V3 nrm{nd[0], nd[1], nd[2]};
    for (int j = 0; j < k; ++j)
      if (dot(nrm, nb[j]) + nd[3] > 0.2) return 0;  // one-sided (Q7), :605-613
    prim[0] = nd[0]; prim[1] = nd[1]; prim[2] = nd[2]; prim[3] = nd[3];
    *counted = 1;                        // surf_num++ / ground_num++ (:623, :766)
    return 1;
  }

  // Build the factor list of one cloud in feature-index order with the `*_maxnum` cap (Q8).
  void build_cloud(int c, const SE3& T, std::vector<Factor>& out, bool parallel) const {
    const size_t n = src[c].size() / 3;
    const int cap = maxnum_of(c);
    out.clear();
    if (!parallel) {
      int counter = 0;  // edge_num / sphere_sum / surf_num / ground_num
      for (size_t i = 0; i < n; ++i) {
        double prim[6]; int cap_checked, counted;
        const int cand = correspond(c, i, T, prim, &cap_checked, &counted);
        if (cap_checked && counter >= cap) return;  // early `return true`
        if (cand) {
          Factor f; f.cloud = c; f.index = static_cast<int>(i); f.w = weights[c][i];
          std::memcpy(f.prim, prim, sizeof(prim));
          out.push_back(f);
        }
        counter += counted;
      }
      return;
    }
    std::vector<unsigned char> cand(n), chk(n), cnt(n);
    std::vector<double> prims(6 * n);
#pragma omp parallel for schedule(dynamic, 256) num_threads(std::min(nthreads(), 32))
    for (long long i = 0; i < static_cast<long long>(n); ++i) {
      int cap_checked, counted;
      cand[i] = static_cast<unsigned char>(correspond(c, static_cast<size_t>(i), T, &prims[6 * i], &cap_checked, &counted));
      chk[i] = static_cast<unsigned char>(cap_checked); cnt[i] = static_cast<unsigned char>(counted);
    }
    int counter = 0;
    for (size_t i = 0; i < n; ++i) {
      if (chk[i] && counter >= cap) return;
      if (cand[i]) {
        Factor f; f.cloud = c; f.index = static_cast<int>(i); f.w = weights[c][i];
        std::memcpy(f.prim, &prims[6 * i], 6 * sizeof(double));
        out.push_back(f);
      }
      counter += cnt[i];
    }
  }

  // ---- Ceres-like evaluation of the whole problem at tangent x ----
  // Returns cost = sum 0.5*rho(|r|^2), rho = CauchyLoss(1.0). If residuals/jac given, they are the
  // CORRECTED (robustified) values: both scaled by sqrt(rho') because rho'' < 0 (corrector.cc).
  // Side effect: writes the per-feature slots (Q3, Q5).
  double evaluate(const double x[6], double* residuals, double* jac, int threads) {
    const SE3 T = se3_exp(x);
    const long long nf = static_cast<long long>(factors.size());
    double cost = 0;
#pragma omp parallel for schedule(static) reduction(+ : cost) num_threads(threads > 0 ? threads : 1)
    for (long long fi = 0; fi < nf; ++fi) {
      const Factor& f = factors[fi];
      const double* s = src[f.cloud].data() + 3 * f.index;
      const V3 p{s[0], s[1], s[2]};
      double r[3], J[18];
      const int nr = (f.cloud == kPlanar || f.cloud == kGround) ? 1 : 3;
      double* slot = &slots[f.cloud][f.index];
      if (f.cloud == kSphere) eval_p2p(T, p, V3{f.prim[0], f.prim[1], f.prim[2]}, f.w, r, jac ? J : nullptr, slot);
      else if (f.cloud == kEdge)
        eval_p2l(T, p, V3{f.prim[0], f.prim[1], f.prim[2]}, V3{f.prim[3], f.prim[4], f.prim[5]}, f.w, r, jac ? J : nullptr, slot);
      else eval_p2pl(T, p, V3{f.prim[0], f.prim[1], f.prim[2]}, f.prim[3], f.w, r, jac ? J : nullptr, slot);
      double sq = 0;
      for (int k = 0; k < nr; ++k) sq += r[k] * r[k];
      // CauchyLoss::Evaluate with a = 1 (b = c = 1)
      const double sum = 1.0 + sq;
      const double inv = 1.0 / sum;
      const double rho0 = std::log(sum);
      const double rho1 = std::max(std::numeric_limits<double>::min(), inv);
      cost += 0.5 * rho0;
      if (residuals || jac) {
        const double sc = std::sqrt(rho1);  // Corrector: rho[2] <= 0 -> residual_scaling = sqrt_rho1, alpha = 0
        const int row = row_offset[fi];
        if (jac) for (int k = 0; k < nr * 6; ++k) jac[row * 6 + k] = J[k] * sc;
        if (residuals) for (int k = 0; k < nr; ++k) residuals[row + k] = r[k] * sc;
      }
    }
    return cost;
  }
};

// Least squares min |A y - b|^2 via Householder QR (DenseQRSolver: lhs.householderQr().solve(rhs)).
// A is m x 6 row-major (destroyed), b length m (destroyed); y length 6.
void householder_ls(std::vector<double>& A, std::vector<double>& b, int m, double y[6]) {
  const int n = 6;
  for (int k = 0; k < n; ++k) {
    double sigma = 0;
    for (int i = k; i < m; ++i) sigma += A[i * n + k] * A[i * n + k];
    const double alpha = std::sqrt(sigma);
    if (alpha == 0) continue;
    const double akk = A[k * n + k];
    const double beta = (akk > 0) ? -alpha : alpha;
    // v = x - beta e1 ; store v in column k (v_k separately)
    const double vk = akk - beta;
    A[k * n + k] = beta;
    const double vnorm2 = sigma - akk * akk + vk * vk;
    if (vnorm2 == 0) continue;
    for (int j = k + 1; j < n; ++j) {
      double s = vk * A[k * n + j];
      for (int i = k + 1; i < m; ++i) s += A[i * n + k] * A[i * n + j];
      const double f = 2.0 * s / vnorm2;
      A[k * n + j] -= f * vk;
      for (int i = k + 1; i < m; ++i) A[i * n + j] -= f * A[i * n + k];
    }
    {
      double s = vk * b[k];
      for (int i = k + 1; i < m; ++i) s += A[i * n + k] * b[i];
      const double f = 2.0 * s / vnorm2;
      b[k] -= f * vk;
      for (int i = k + 1; i < m; ++i) b[i] -= f * A[i * n + k];
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < n; ++j) s -= A[k * n + j] * y[j];
    y[k] = (A[k * n + k] != 0) ? s / A[k * n + k] : 0.0;
  }
}

// 2-D trust-region boundary problem of the subspace dogleg (dogleg_strategy.cc
// FindMinimumOnTrustRegionBoundary): minimise 0.5 y^T B y + g^T y subject to |y| = radius.
// Ceres finds the roots of a quartic; we bracket on a fine angular grid and polish with bisection on
// the derivative -- the same global minimiser to ~1e-15.
void min_on_boundary_2d(const double B[4], const double g[2], double radius, double y[2]) {
  auto f = [&](double t) {
    const double c = radius * std::cos(t), s = radius * std::sin(t);
    return 0.5 * (B[0] * c * c + (B[1] + B[2]) * c * s + B[3] * s * s) + g[0] * c + g[1] * s;
  };
  auto df = [&](double t) {
    const double c = std::cos(t), s = std::sin(t), r = radius;
    return r * r * ((B[3] - B[0]) * c * s + 0.5 * (B[1] + B[2]) * (c * c - s * s)) + r * (-g[0] * s + g[1] * c);
  };
  const int N = 2048;
  double best_t = 0, best_f = std::numeric_limits<double>::infinity();
  for (int i = 0; i < N; ++i) {
    const double t0 = 2 * kPi * i / N, t1 = 2 * kPi * (i + 1) / N;
    double lo = t0, hi = t1;
    const double dlo = df(lo), dhi = df(hi);
    double cand;
    if (dlo < 0 && dhi > 0) {  // a local minimum inside: bisection on f'
      for (int it = 0; it < 200; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (df(mid) < 0) lo = mid; else hi = mid;
      }
      cand = 0.5 * (lo + hi);
    } else {
      cand = (f(t0) < f(t1)) ? t0 : t1;
    }
    const double fc = f(cand);
    if (fc < best_f) { best_f = fc; best_t = cand; }
  }
  y[0] = radius * std::cos(best_t);
  y[1] = radius * std::sin(best_t);
}

// ---------------------------------------------------------------------------------------------
// Ceres 2.0 TrustRegionMinimizer + DoglegStrategy(SUBSPACE_DOGLEG) + DENSE_QR, restated for one
// 6-parameter block with the options of ref: registration.cpp:1036-1047 (everything else default):
//   initial_trust_region_radius 1e4, max 1e16, min 1e-32, min_relative_decrease 1e-3,
//   min_lm_diagonal 1e-6, max_lm_diagonal 1e32, function_tolerance 1e-6, gradient_tolerance 1e-10,
//   parameter_tolerance 1e-8, jacobi_scaling, monotonic steps, max_num_consecutive_invalid_steps 5.
// ---------------------------------------------------------------------------------------------
void ceres_solve(Oracle& O, int eval_threads, oracle_outer_trace* tr) {
  const int m = O.num_rows;
  const int max_it = O.cfg.ceres_max_num_iterations;
  double* params = O.params;
  if (tr) {
    std::memcpy(tr->x_start, params, sizeof(double) * 6);
    tr->n_inner = 0; tr->termination = 0;
  }
  if (O.factors.empty()) {  // no residual blocks: Ceres returns immediately, parameters untouched
    if (tr) { tr->termination = 5; tr->initial_cost = tr->final_cost = 0; std::memcpy(tr->x_end, params, 48); }
    return;
  }
  std::vector<double> r(m), J(static_cast<size_t>(m) * 6), model(m);
  double x[6], cand[6];
  std::memcpy(x, params, sizeof(x));
  auto vnorm6 = [](const double* v) { double s = 0; for (int i = 0; i < 6; ++i) s += v[i] * v[i]; return std::sqrt(s); };

  double scale[6], grad[6];
  double x_norm = vnorm6(x);
  double x_cost = 0, gradient_max_norm = 0;
  // EvaluateGradientAndJacobian
  auto eval_grad_jac = [&](bool first) {
    x_cost = O.evaluate(x, r.data(), J.data(), eval_threads);
    for (int j = 0; j < 6; ++j) grad[j] = 0;
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < 6; ++j) grad[j] += J[static_cast<size_t>(i) * 6 + j] * r[i];
    if (first) {
      double cn[6] = {0, 0, 0, 0, 0, 0};
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < 6; ++j) cn[j] += J[static_cast<size_t>(i) * 6 + j] * J[static_cast<size_t>(i) * 6 + j];
      for (int j = 0; j < 6; ++j) scale[j] = 1.0 / (1.0 + std::sqrt(cn[j]));
      if (tr) {
        for (int a = 0; a < 6; ++a) {
          tr->g0[a] = grad[a];
          for (int b = 0; b < 6; ++b) {
            double s = 0;
            for (int i = 0; i < m; ++i) s += J[static_cast<size_t>(i) * 6 + a] * J[static_cast<size_t>(i) * 6 + b];
            tr->H0[a * 6 + b] = s;
          }
        }
      }
    }
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < 6; ++j) J[static_cast<size_t>(i) * 6 + j] *= scale[j];
    // gradient_max_norm = |x - Plus(x, -g)|_inf
    double ng[6], proj[6];
    for (int j = 0; j < 6; ++j) ng[j] = -grad[j];
    se3_plus(x, ng, proj);
    gradient_max_norm = 0;
    for (int j = 0; j < 6; ++j) gradient_max_norm = std::max(gradient_max_norm, std::fabs(x[j] - proj[j]));
  };

  eval_grad_jac(true);  // IterationZero
  if (tr) tr->initial_cost = x_cost;
  // iteration 0 is "successful": parameters <- x (unchanged)
  double radius = O.cfg.initial_trust_region_radius, mu = 1e-8;
  const double min_mu = 1e-8, max_mu = 1.0;
  bool reuse = false;
  double D[6], sgrad[6], gn[6];  // diagonal_, gradient_ (scaled), gauss_newton_step_ (scaled)
  double sub_basis[12], sub_g[2], sub_B[4];
  bool sub_1d = false;
  int num_invalid = 0;
  int iteration = 0;
  int termination = 0;
  auto finish = [&]() {
    if (tr) { tr->termination = termination; tr->final_cost = x_cost; std::memcpy(tr->x_end, params, 48); }
  };
  if (gradient_max_norm <= 1e-10) { termination = 3; finish(); return; }

  while (true) {
    if (iteration >= max_it) { termination = 0; break; }
    ++iteration;
    oracle_inner_trace* it = (tr && iteration <= ORACLE_MAX_INNER) ? &tr->inner[iteration - 1] : nullptr;
    if (tr) tr->n_inner = iteration;
    if (it) std::memset(it, 0, sizeof(*it));
    // ---- DoglegStrategy::ComputeStep ----
    double step[6];
    bool solver_failed = false;
    bool used_gn = false;
    double step_norm = 0;
    if (!reuse) {
      reuse = true;
      double cn[6] = {0, 0, 0, 0, 0, 0};
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < 6; ++j) cn[j] += J[static_cast<size_t>(i) * 6 + j] * J[static_cast<size_t>(i) * 6 + j];
      for (int j = 0; j < 6; ++j) D[j] = std::sqrt(std::min(std::max(cn[j], 1e-6), 1e32));
      for (int j = 0; j < 6; ++j) sgrad[j] = 0;
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < 6; ++j) sgrad[j] += J[static_cast<size_t>(i) * 6 + j] * r[i];
      for (int j = 0; j < 6; ++j) sgrad[j] /= D[j];
      // ComputeGaussNewtonStep: min |J y - r|^2 + |sqrt(mu) D y|^2 by QR of the stacked system
      solver_failed = true;
      while (mu < max_mu) {
        std::vector<double> A(static_cast<size_t>(m + 6) * 6), b(m + 6);
        std::memcpy(A.data(), J.data(), sizeof(double) * static_cast<size_t>(m) * 6);
        std::memcpy(b.data(), r.data(), sizeof(double) * m);
        for (int j = 0; j < 6; ++j) {
          for (int k = 0; k < 6; ++k) A[static_cast<size_t>(m + j) * 6 + k] = (j == k) ? D[j] * std::sqrt(mu) : 0.0;
          b[m + j] = 0;
        }
        double yv[6];
        householder_ls(A, b, m + 6, yv);
        bool valid = true;
        for (int j = 0; j < 6; ++j) if (!std::isfinite(yv[j])) valid = false;
        if (!valid) { mu *= 10.0; continue; }
        for (int j = 0; j < 6; ++j) gn[j] = -D[j] * yv[j];
        solver_failed = false;
        break;
      }
      if (!solver_failed) {
        // ComputeSubspaceModel: orthonormal basis of span{sgrad, gn} (Gram-Schmidt with pivoting)
        double n0 = vnorm6(sgrad), n1 = vnorm6(gn);
        const double* first = (n0 >= n1) ? sgrad : gn;
        const double* second = (n0 >= n1) ? gn : sgrad;
        const double nf = std::max(n0, n1);
        if (nf == 0) { solver_failed = true; }
        else {
          double u0[6], u1[6];
          for (int j = 0; j < 6; ++j) u0[j] = first[j] / nf;
          double pr = 0;
          for (int j = 0; j < 6; ++j) pr += u0[j] * second[j];
          for (int j = 0; j < 6; ++j) u1[j] = second[j] - pr * u0[j];
          const double n2 = vnorm6(u1);
          // rank test of ColPivHouseholderQR: |R11| <= eps * size * |R00|
          sub_1d = !(n2 > nf * 6 * std::numeric_limits<double>::epsilon());
          if (!sub_1d) {
            for (int j = 0; j < 6; ++j) u1[j] /= n2;
            for (int j = 0; j < 6; ++j) { sub_basis[j] = u0[j]; sub_basis[6 + j] = u1[j]; }
            sub_g[0] = sub_g[1] = 0;
            for (int j = 0; j < 6; ++j) { sub_g[0] += u0[j] * sgrad[j]; sub_g[1] += u1[j] * sgrad[j]; }
            double t0[6], t1[6];
            for (int j = 0; j < 6; ++j) { t0[j] = u0[j] / D[j]; t1[j] = u1[j] / D[j]; }
            double b00 = 0, b01 = 0, b11 = 0;
            for (int i = 0; i < m; ++i) {
              double a0 = 0, a1 = 0;
              for (int j = 0; j < 6; ++j) { a0 += J[static_cast<size_t>(i) * 6 + j] * t0[j]; a1 += J[static_cast<size_t>(i) * 6 + j] * t1[j]; }
              b00 += a0 * a0; b01 += a0 * a1; b11 += a1 * a1;
            }
            sub_B[0] = b00; sub_B[1] = sub_B[2] = b01; sub_B[3] = b11;
          }
        }
      }
    }
    if (!solver_failed) {
      // ComputeSubspaceDoglegStep
      const double gn_norm = vnorm6(gn);
      if (gn_norm <= radius) {
        for (int j = 0; j < 6; ++j) step[j] = gn[j] / D[j];
        step_norm = gn_norm; used_gn = true;
      } else if (sub_1d) {
        const double gnm = vnorm6(sgrad);
        for (int j = 0; j < 6; ++j) step[j] = -(radius / gnm) * sgrad[j] / D[j];
        step_norm = radius;
      } else {
        double y2[2];
        min_on_boundary_2d(sub_B, sub_g, radius, y2);
        for (int j = 0; j < 6; ++j) step[j] = (sub_basis[j] * y2[0] + sub_basis[6 + j] * y2[1]) / D[j];
        step_norm = radius;
      }
    }
    if (it) { it->radius = radius; it->step_norm_scaled = step_norm; it->used_gauss_newton = used_gn ? 1 : 0; }
    // ---- ComputeTrustRegionStep: model cost change and validity ----
    bool valid = false;
    double model_cost_change = 0;
    double delta[6] = {0, 0, 0, 0, 0, 0};
    if (!solver_failed) {
      double acc = 0;
      for (int i = 0; i < m; ++i) {
        double mr = 0;
        for (int j = 0; j < 6; ++j) mr += J[static_cast<size_t>(i) * 6 + j] * step[j];
        acc += mr * (r[i] + mr / 2.0);
      }
      model_cost_change = -acc;
      valid = model_cost_change > 0.0;
      if (valid) { for (int j = 0; j < 6; ++j) delta[j] = step[j] * scale[j]; num_invalid = 0; }
    }
    if (it) it->model_cost_change = model_cost_change;
    if (!valid) {
      // HandleInvalidStep
      ++num_invalid;
      if (it) it->accepted = -1;
      if (num_invalid >= 5) { termination = 6; break; }
      mu *= 10.0; reuse = false;  // StepIsInvalid
      continue;
    }
    // ---- ComputeCandidatePointAndEvaluateCost ----
    se3_plus(x, delta, cand);
    const double cand_cost = O.evaluate(cand, nullptr, nullptr, eval_threads);
    if (it) { std::memcpy(it->x_candidate, cand, 48); it->candidate_cost = cand_cost; }
    // ParameterToleranceReached
    {
      double d[6];
      for (int j = 0; j < 6; ++j) d[j] = x[j] - cand[j];
      const double sn = vnorm6(d);
      if (sn <= 1e-8 * (x_norm + 1e-8)) { termination = 2; if (it) it->accepted = 2; break; }
    }
    // FunctionToleranceReached
    if (std::fabs(x_cost - cand_cost) <= 1e-6 * x_cost) { termination = 1; if (it) it->accepted = 2; break; }
    // IsStepSuccessful (monotonic evaluator: quality = (x_cost - cand_cost) / model_cost_change)
    const double rel = (x_cost - cand_cost) / model_cost_change;
    if (it) it->relative_decrease = rel;
    if (rel > 1e-3) {
      // HandleSuccessfulStep
      std::memcpy(x, cand, sizeof(x));
      x_norm = vnorm6(x);
      eval_grad_jac(false);
      if (rel < 0.25) radius *= 0.5;                       // DoglegStrategy::StepAccepted
      if (rel > 0.75) radius = std::max(radius, 3.0 * step_norm);
      mu = std::max(min_mu, 2.0 * mu / 10.0);
      reuse = false;
      std::memcpy(params, x, sizeof(x));  // x_cost < minimum_cost (monotonic)
      if (it) it->accepted = 1;
      if (gradient_max_norm <= 1e-10) { termination = 3; break; }
    } else {
      radius *= 0.5; reuse = true;  // StepRejected
      if (it) it->accepted = 0;
    }
    if (radius <= 1e-32) { termination = 4; break; }
  }
  finish();
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// LocalRegistration::scanMatching, ref: registration.cpp:879-1133.
int scan_match(Oracle& O, const double predict[16], double result[16], oracle_stats* st) {
  const oracle_config& cfg = O.cfg;
  const double t_begin = now_s();
  if (st) std::memset(st, 0, sizeof(*st));
  for (int c = 0; c < 4; ++c)  // :928-929 (asserts, live in the reference build, Q9)
    if (O.src[c].size() / 3 < 10 || O.tgt[c].size() / 3 < 10) return 1;

  se3_log(se3_from_matrix(predict), O.params);           // :881
  std::memcpy(O.last_pose, O.curr_pose, sizeof(O.last_pose));  // :882
  {
    const double* w = O.params + 3;
    if (std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]) < 1e-2) {  // :884-886 (Q1)
      const double* d = cfg.reinit_dir;
      const double nn = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      for (int j = 0; j < 3; ++j) O.params[3 + j] = d[j] / nn * 1e-4;
    }
  }
  if (st) std::memcpy(st->x_init, O.params, 48);

  // :888-915 KD-trees rebuilt on every call, 4 OpenMP sections.
  double t0 = now_s();
#pragma omp parallel sections num_threads(4)
  {
#pragma omp section
    O.tree[kEdge].build(O.tgt[kEdge].data(), O.tgt[kEdge].size() / 3);
#pragma omp section
    O.tree[kSphere].build(O.tgt[kSphere].data(), O.tgt[kSphere].size() / 3);
#pragma omp section
    O.tree[kPlanar].build(O.tgt[kPlanar].data(), O.tgt[kPlanar].size() / 3);
#pragma omp section
    O.tree[kGround].build(O.tgt[kGround].data(), O.tgt[kGround].size() / 3);
  }
  if (st) st->t_kdtree = now_s() - t0;

  for (int c = 0; c < 4; ++c) {  // :931-949
    O.weights[c].assign(O.src[c].size() / 3, 1.0);
    O.slots[c].assign(O.src[c].size() / 3, 0.0);
  }
  const double inf = std::numeric_limits<double>::infinity();
  double planar_prev_cost = inf;
  double mu = 1.0;                                        // :961
  double noise_bound_sq = cfg.noise_bound * cfg.noise_bound;
  if (noise_bound_sq < 1e-16) noise_bound_sq = 1e-2;      // :963-964

  // small loops (40k features): more than ~32 OpenMP threads only adds fork/join cost on a 128-core host
  const int nth = std::min(O.nthreads(), cfg.threads_mode == 1 ? 32 : O.nthreads());
  const bool par = cfg.threads_mode == 1;
  const int eval_threads = par ? nth : std::max(1, nth / 2);  // options.num_threads = num_threads_/2 (:1044)
  // factor_num -> which builders run (:979-1016): 4: planar,ground,edge,sphere; 3: -sphere; 2: planar,ground
  const int order[4] = {kPlanar, kGround, kEdge, kSphere};
  const int nbuild = (cfg.factor_num >= 2 && cfg.factor_num <= 4) ? cfg.factor_num : 0;

  for (int iter = 0; iter < cfg.max_iterations; ++iter) {  // :966
    oracle_outer_trace* tr = (st && iter < ORACLE_MAX_OUTER) ? &st->outer[iter] : nullptr;
    if (tr) std::memset(tr, 0, sizeof(*tr));
    if (st) st->n_outer = iter + 1;
    t0 = now_s();
    const SE3 T = se3_exp(O.params);
    std::vector<Factor> built[4];
    if (par) {
      for (int b = 0; b < nbuild; ++b) O.build_cloud(order[b], T, built[b], true);
    } else {
      // std::async x factor_num, one serial builder per cloud (:976-1020)
#pragma omp parallel for schedule(static, 1) num_threads(4)
      for (int b = 0; b < nbuild; ++b) O.build_cloud(order[b], T, built[b], false);
    }
    O.factors.clear(); O.row_offset.clear(); O.num_rows = 0;
    for (int b = 0; b < nbuild; ++b) {
      if (tr) tr->n_factors[order[b]] = static_cast<int>(built[b].size());
      for (const Factor& f : built[b]) {
        O.factors.push_back(f);
        O.row_offset.push_back(O.num_rows);
        O.num_rows += (f.cloud == kPlanar || f.cloud == kGround) ? 1 : 3;
      }
    }
    if (st) st->t_factors += now_s() - t0;

    if (iter == 0) {  // :1027-1033 (Q2: slots are still all zero here)
      auto maxc = [](const std::vector<double>& v) { double m = -std::numeric_limits<double>::infinity(); for (double d : v) m = std::max(m, d); return m; };
      const double mp = maxc(O.slots[kPlanar]), me = maxc(O.slots[kEdge]), ms = maxc(O.slots[kSphere]);
      const double max_residual = mp > me ? (mp > ms ? mp : ms) : (me > ms ? me : ms);
      mu = 1 / (2 * max_residual / noise_bound_sq - 1.0);
      if (mu <= 0) mu = 1e-10;
    }

    t0 = now_s();
    ceres_solve(O, eval_threads, tr);                     // :1036-1047
    if (st) st->t_solve += now_s() - t0;

    t0 = now_s();
    const double th1 = (mu + 1) / mu * noise_bound_sq;    // :1049-1050
    const double th2 = mu / (mu + 1) * noise_bound_sq;
    if (tr) { tr->mu = mu; tr->th1 = th1; tr->th2 = th2; }
    for (int b = 0; b < nbuild; ++b) {                    // :1053-1086
      const int c = order[b];
      oracle_update_weight(O.weights[c].data(), O.slots[c].data(), O.weights[c].size(), noise_bound_sq, th1, th2, mu);
    }
    mu = mu * std::exp(double(iter + 1) * cfg.gnc_factor);  // :1089
    double sums[4];
    for (int c = 0; c < 4; ++c) { double s = 0; for (double v : O.slots[c]) s += v; sums[c] = s; if (tr) tr->slot_sum[c] = s; }
    const double planar_cost = sums[kPlanar];
    const double planar_cost_diff = std::fabs(planar_cost - planar_prev_cost);
    if (st) st->t_weights += now_s() - t0;
    if (planar_cost_diff < cfg.cost_threshold) {          // :1108 (Q11)
      if (st) st->converged_early = 1;
      break;
    }
    planar_prev_cost = planar_cost;
    for (int c = 0; c < 4; ++c) std::fill(O.slots[c].begin(), O.slots[c].end(), 0.0);  // :1118-1121
  }

  se3_to_matrix(se3_exp(O.params), result);               // :1124
  std::memcpy(O.curr_pose, result, sizeof(O.curr_pose));
  if (st) { std::memcpy(st->x_final, O.params, 48); st->t_total = now_s() - t_begin; }
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C API
// ---------------------------------------------------------------------------------------------
extern "C" {

void oracle_default_config(oracle_config* c) {  // ref: config/mapping/lidar_odometry.yaml:23-39
  c->k_corr = 10; c->factor_num = 4;
  c->edge_dist_thres = 1.0; c->sphere_dist_thres = 0.5; c->planar_dist_thres = 0.5; c->ground_dist_thres = 0.5;
  c->edge_dir_thres = 0.85;
  c->edge_maxnum = 1200; c->sphere_maxnum = 200; c->planar_maxnum = 2500; c->ground_maxnum = 2000;
  c->max_iterations = 4; c->cost_threshold = 0.000000005; c->gnc_factor = 11.8; c->noise_bound = 0.01;
  c->fitness_thres = 0.02;
  c->ceres_max_num_iterations = 4;
  c->reinit_dir[0] = 1.0; c->reinit_dir[1] = 1.0; c->reinit_dir[2] = 1.0;
  c->initial_trust_region_radius = 1e4;
  c->threads_mode = 0; c->num_threads = 0;
}

void* oracle_create(const oracle_config* cfg) {
  Oracle* o = new Oracle();
  o->cfg = *cfg;
  return o;
}
void oracle_destroy(void* h) { delete static_cast<Oracle*>(h); }

static int set_clouds(std::vector<double>* dst, const double* const xyz[4], const size_t n[4]) {
  for (int c = 0; c < 4; ++c) dst[c].assign(xyz[c], xyz[c] + 3 * n[c]);
  return 0;
}
int oracle_set_source(void* h, const double* const xyz[4], const size_t n[4]) {  // ref: registration.cpp:232-239
  return set_clouds(static_cast<Oracle*>(h)->src, xyz, n);
}
int oracle_set_target(void* h, const double* const xyz[4], const size_t n[4]) {  // ref: registration.cpp:241-248
  return set_clouds(static_cast<Oracle*>(h)->tgt, xyz, n);
}
int oracle_scan_match(void* h, const double predict[16], double result[16], oracle_stats* stats) {
  return scan_match(*static_cast<Oracle*>(h), predict, result, stats);
}

// LocalRegistration::getFitnessScore, ref: registration.cpp:257-296 (untransformed scan points, Q14).
int oracle_fitness(void* h, double* fitness, double* rmse) {
  Oracle& O = *static_cast<Oracle*>(h);
  *fitness = 0; *rmse = 0;
  if (O.cfg.fitness_thres <= 0.0) return 0;
  const int order[4] = {kEdge, kSphere, kPlanar, kGround};
  for (int oi = 0; oi < 4; ++oi) {
    const int c = order[oi];
    if (O.tree[c].n != O.tgt[c].size() / 3 || O.tree[c].pts != O.tgt[c].data()) O.tree[c].build(O.tgt[c].data(), O.tgt[c].size() / 3);
    double err = 0; int corr = 0;
    const size_t n = O.src[c].size() / 3;
    for (size_t i = 0; i < n; ++i) {
      int idx; double d2;
      if (O.tree[c].search_hybrid(&O.src[c][3 * i], O.cfg.fitness_thres, 1, &idx, &d2) > 0) { err += d2; ++corr; }
    }
    if (corr > 0) { *fitness += double(corr) / double(n); *rmse += std::sqrt(err / double(corr)); }
  }
  return 0;
}

// getPoseIncrement, ref: registration.cpp:374-376: last^-1 * curr.
void oracle_get_pose_increment(void* h, double out[16]) {
  Oracle& O = *static_cast<Oracle*>(h);
  const double* L = O.last_pose; const double* C = O.curr_pose;
  // inverse of rigid L: R^T, -R^T t
  double Li[16];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Li[c * 4 + r] = L[r * 4 + c];
  for (int r = 0; r < 3; ++r) Li[12 + r] = -(Li[0 * 4 + r] * L[12] + Li[1 * 4 + r] * L[13] + Li[2 * 4 + r] * L[14]);
  Li[3] = Li[7] = Li[11] = 0; Li[15] = 1;
  for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) {
    double s = 0;
    for (int k = 0; k < 4; ++k) s += Li[k * 4 + r] * C[c * 4 + k];
    out[c * 4 + r] = s;
  }
}

int oracle_get_weights(void* h, int cloud, double* w, size_t n) {
  Oracle& O = *static_cast<Oracle*>(h);
  if (cloud < 0 || cloud > 3 || n != O.weights[cloud].size()) return 1;
  std::memcpy(w, O.weights[cloud].data(), n * sizeof(double));
  return 0;
}

void oracle_se3_exp(const double a[6], double T[16]) { se3_to_matrix(se3_exp(a), T); }
void oracle_se3_log(const double T[16], double a[6]) { se3_log(se3_from_matrix(T), a); }
void oracle_se3_plus(const double x[6], const double d[6], double out[6]) { se3_plus(x, d, out); }
void oracle_se3_exp_quat(const double a[6], double q[7]) {
  const SE3 T = se3_exp(a);
  q[0] = T.w; q[1] = T.x; q[2] = T.y; q[3] = T.z; q[4] = T.t.x; q[5] = T.t.y; q[6] = T.t.z;
}

int oracle_knn(const double* pts, size_t n, const double* queries, size_t nq, double radius, int k, int* idx,
               double* d2, int* count, int brute_force) {
  KDTree tree;
  if (!brute_force) tree.build(pts, n);
  // capped like the all-cores mode of scan_match: 128 spinning threads on the GPU boxes' hosts are 10x slower than 32
#pragma omp parallel for schedule(dynamic, 64) num_threads(std::min(omp_get_max_threads(), 32))
  for (long long i = 0; i < static_cast<long long>(nq); ++i) {
    int* id = idx + i * k; double* dd = d2 + i * k;
    for (int j = 0; j < k; ++j) { id[j] = -1; dd[j] = std::numeric_limits<double>::infinity(); }
    count[i] = brute_force ? brute_hybrid(pts, n, queries + 3 * i, radius, k, id, dd)
                           : tree.search_hybrid(queries + 3 * i, radius, k, id, dd);
    for (int j = count[i]; j < k; ++j) { id[j] = -1; dd[j] = std::numeric_limits<double>::infinity(); }
  }
  return 0;
}

void oracle_fit_plane(const double* pts, int n, double out[4]) {
  std::vector<V3> v(n);
  for (int i = 0; i < n; ++i) v[i] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  fit_best_plane(v.data(), n, out);
}
int oracle_fit_line(const double* pts, int n, double dir_thres, double a[3], double b[3], double mean[3],
                    double dir[3], double eig[3]) {
  std::vector<V3> v(n);
  for (int i = 0; i < n; ++i) v[i] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  V3 A, B, M, Dv;
  const int ok = fit_line(v.data(), n, dir_thres, &A, &B, &M, &Dv, eig);
  a[0] = A.x; a[1] = A.y; a[2] = A.z; b[0] = B.x; b[1] = B.y; b[2] = B.z;
  mean[0] = M.x; mean[1] = M.y; mean[2] = M.z; dir[0] = Dv.x; dir[1] = Dv.y; dir[2] = Dv.z;
  return ok;
}
void oracle_sym_eig3(const double cov[9], double eig[3], double vec[9]) { sym_eig3(cov, eig, vec); }

void oracle_eval_point_to_point(const double x[6], const double p[3], const double q[3], double w, double r[3],
                                double J[18], double* cost) {
  eval_p2p(se3_exp(x), V3{p[0], p[1], p[2]}, V3{q[0], q[1], q[2]}, w, r, J, cost);
}
void oracle_eval_point_to_line(const double x[6], const double p[3], const double a[3], const double b[3], double w,
                               double r[3], double J[18], double* cost) {
  eval_p2l(se3_exp(x), V3{p[0], p[1], p[2]}, V3{a[0], a[1], a[2]}, V3{b[0], b[1], b[2]}, w, r, J, cost);
}
void oracle_eval_point_to_plane(const double x[6], const double p[3], const double n[3], double d, double w,
                                double r[1], double J[6], double* cost) {
  eval_p2pl(se3_exp(x), V3{p[0], p[1], p[2]}, V3{n[0], n[1], n[2]}, d, w, r, J, cost);
}

int oracle_build_factors(void* h, int cloud, const double x[6], int* valid, double* prim, size_t n) {
  Oracle& O = *static_cast<Oracle*>(h);
  if (cloud < 0 || cloud > 3 || n != O.src[cloud].size() / 3) return 1;
  if (O.tree[cloud].pts != O.tgt[cloud].data() || O.tree[cloud].n != O.tgt[cloud].size() / 3)
    O.tree[cloud].build(O.tgt[cloud].data(), O.tgt[cloud].size() / 3);
  O.weights[cloud].assign(n, 1.0);
  std::vector<Factor> out;
  O.build_cloud(cloud, se3_exp(x), out, true);
  for (size_t i = 0; i < n; ++i) { valid[i] = 0; for (int j = 0; j < 6; ++j) prim[6 * i + j] = 0; }
  for (const Factor& f : out) { valid[f.index] = 1; std::memcpy(prim + 6 * f.index, f.prim, 48); }
  return 0;
}

// LocalRegistration::updateWeight, ref: registration.cpp:858-876.
void oracle_update_weight(double* weights, const double* slots, size_t n, double noise_bound_sq, double th1,
                          double th2, double mu) {
  for (size_t i = 0; i < n; ++i) {
    if (slots[i] == 0) continue;  // Q13
    if (slots[i] >= th1) weights[i] = 0.0;
    else if (slots[i] <= th2) weights[i] = 1.0;
    else weights[i] = std::sqrt(noise_bound_sq * mu * (mu + 1) / slots[i]) - mu;
  }
}

// dogleg_strategy.cc FindMinimumOnTrustRegionBoundary (2-D subspace problem), exposed for the quartic-root pin
void oracle_min_on_boundary_2d(const double B[4], const double g[2], double radius, double y[2]) {
  min_on_boundary_2d(B, g, radius, y);
}

// ---------------------------------------------------------------------------------------------
// (f)-1 submap maintenance
// ---------------------------------------------------------------------------------------------
void oracle_submap_default_config(oracle_submap_config* c) {   // ref: config/mapping/lidar_odometry.yaml:6-17
  c->ground_down_sample = 0.3; c->ground_down_sample_submap = 0.45; c->edge_down_sample_submap = 0.3;
  c->planar_frame_size = 3; c->sphere_frame_size = 3;
  c->edge_crop_box_length = 100.0; c->ground_crop_box_length = 100.0;
}

// PointCloud2::VoxelDownSample, ref: src/open3d/PointCloud2.cpp:358-403.
size_t oracle_voxel_down_sample(const double* pts, size_t n, double voxel, double* out) {
  if (n == 0) return 0;
  double mn[3] = {pts[0], pts[1], pts[2]};
  for (size_t i = 1; i < n; ++i) for (int d = 0; d < 3; ++d) mn[d] = std::min(mn[d], pts[3 * i + d]);
  for (int d = 0; d < 3; ++d) mn[d] -= voxel * 0.5;                        // voxel_min_bound (:367)
  struct Acc { long long key[3]; double s[3]; long long cnt; };
  std::vector<std::pair<std::array<long long, 3>, size_t>> keyed(n);
  for (size_t i = 0; i < n; ++i) {
    std::array<long long, 3> k;
    for (int d = 0; d < 3; ++d) k[d] = (long long)std::floor((pts[3 * i + d] - mn[d]) / voxel);   // :381-383
    keyed[i] = {k, i};
  }
  std::stable_sort(keyed.begin(), keyed.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  size_t m = 0;
  for (size_t i = 0; i < n;) {
    size_t j = i;
    double s[3] = {0, 0, 0};
    while (j < n && keyed[j].first == keyed[i].first) {                   // AddPoint in input order (:254)
      for (int d = 0; d < 3; ++d) s[d] += pts[3 * keyed[j].second + d];
      ++j;
    }
    for (int d = 0; d < 3; ++d) out[3 * m + d] = s[d] / double(j - i);     // GetAveragePoint (:271-273)
    ++m;
    i = j;
  }
  return m;
}

// PointCloud2::Crop(AxisAlignedBoundingBox), ref: PointCloud2.cpp:551-559 (bounds inclusive, order kept).
size_t oracle_crop(const double* pts, size_t n, const double lo[3], const double hi[3], double* out) {
  size_t m = 0;
  for (size_t i = 0; i < n; ++i) {
    bool in = true;
    for (int d = 0; d < 3; ++d) in = in && pts[3 * i + d] >= lo[d] && pts[3 * i + d] <= hi[d];
    if (in) { for (int d = 0; d < 3; ++d) out[3 * m + d] = pts[3 * i + d]; ++m; }
  }
  return m;
}

namespace {
struct SubmapOracle {
  oracle_submap_config cfg;
  std::vector<double> cloud[4];                       // edge, sphere, planar, ground (world frame)
  std::vector<std::vector<double>> planar_buf, sphere_buf;   // per-frame clouds, already transformed? no: raw + pose
  std::vector<std::array<double, 16>> planar_pose, sphere_pose;
};
void transform_into(const double T[16], const double* in, size_t n, std::vector<double>& out) {   // Transform + operator+=
  for (size_t i = 0; i < n; ++i) {
    const double x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];
    out.push_back(T[0] * x + T[4] * y + T[8] * z + T[12]);
    out.push_back(T[1] * x + T[5] * y + T[9] * z + T[13]);
    out.push_back(T[2] * x + T[6] * y + T[10] * z + T[14]);
  }
}
}  // namespace

void* oracle_submap_create(const oracle_submap_config* c) { auto* s = new SubmapOracle(); s->cfg = *c; return s; }
void oracle_submap_destroy(void* s) { delete static_cast<SubmapOracle*>(s); }

// first frame, ref: front_end.cpp:285-305 (clouds stay in the sensor frame of frame 0).
int oracle_submap_init(void* sp, const double* edge, size_t ne, const double* ground_raw, size_t ng,
                       const double* planar_sub, size_t np, const double* sphere_sub, size_t ns) {
  SubmapOracle& S = *static_cast<SubmapOracle*>(sp);
  S.cloud[0].assign(edge, edge + 3 * ne);                                              // :286
  std::vector<double> g(3 * ng);
  const size_t m = oracle_voxel_down_sample(ground_raw, ng, S.cfg.ground_down_sample, g.data());   // :287
  S.cloud[3].assign(g.begin(), g.begin() + 3 * m);
  S.cloud[2].assign(planar_sub, planar_sub + 3 * np);                                  // :291
  S.cloud[1].assign(sphere_sub, sphere_sub + 3 * ns);                                  // :292
  // NOTE: the reference does NOT push frame 0 into the sliding-window buffers (:285-305)
  return 0;
}

// ref: front_end.cpp:201-267.
int oracle_submap_update(void* sp, const double pose[16], const double* edge_scan, size_t ne, const double* ground_scan,
                         size_t ng, const double* planar_sub, size_t np, const double* sphere_sub, size_t ns) {
  SubmapOracle& S = *static_cast<SubmapOracle*>(sp);
  std::array<double, 16> P;
  std::copy(pose, pose + 16, P.begin());
  S.sphere_buf.emplace_back(sphere_sub, sphere_sub + 3 * ns); S.sphere_pose.push_back(P);        // :202-205
  S.planar_buf.emplace_back(planar_sub, planar_sub + 3 * np); S.planar_pose.push_back(P);        // :207-210
  while ((int)S.sphere_buf.size() > S.cfg.sphere_frame_size) { S.sphere_buf.erase(S.sphere_buf.begin()); S.sphere_pose.erase(S.sphere_pose.begin()); }
  while ((int)S.planar_buf.size() > S.cfg.planar_frame_size) { S.planar_buf.erase(S.planar_buf.begin()); S.planar_pose.erase(S.planar_pose.begin()); }
  // :220-230 -- the SPHERE submap is rebuilt from the PLANAR buffer (sic)
  S.cloud[1].clear();
  for (size_t f = 0; f < S.planar_buf.size(); ++f) transform_into(S.planar_pose[f].data(), S.planar_buf[f].data(), S.planar_buf[f].size() / 3, S.cloud[1]);
  S.cloud[2].clear();                                                                              // :232-242
  for (size_t f = 0; f < S.planar_buf.size(); ++f) transform_into(S.planar_pose[f].data(), S.planar_buf[f].data(), S.planar_buf[f].size() / 3, S.cloud[2]);
  transform_into(pose, edge_scan, ne, S.cloud[0]);                                                 // :245
  transform_into(pose, ground_scan, ng, S.cloud[3]);                                               // :246
  const double c[3] = {pose[12], pose[13], pose[14]};                                              // :250
  {
    const double L = S.cfg.edge_crop_box_length;
    const double lo[3] = {c[0] - L, c[1] - L, c[2] - L}, hi[3] = {c[0] + L, c[1] + L, c[2] + L};
    std::vector<double> cr(S.cloud[0].size()), ds(S.cloud[0].size());
    const size_t m = oracle_crop(S.cloud[0].data(), S.cloud[0].size() / 3, lo, hi, cr.data());     // :257
    const size_t k = oracle_voxel_down_sample(cr.data(), m, S.cfg.edge_down_sample_submap, ds.data());
    S.cloud[0].assign(ds.begin(), ds.begin() + 3 * k);
  }
  {
    const double L = S.cfg.ground_crop_box_length;
    const double lo[3] = {c[0] - L, c[1] - L, c[2] - L}, hi[3] = {c[0] + L, c[1] + L, c[2] + L};
    std::vector<double> cr(S.cloud[3].size()), ds(S.cloud[3].size());
    const size_t m = oracle_crop(S.cloud[3].data(), S.cloud[3].size() / 3, lo, hi, cr.data());     // :264
    const size_t k = oracle_voxel_down_sample(cr.data(), m, S.cfg.ground_down_sample_submap, ds.data());
    S.cloud[3].assign(ds.begin(), ds.begin() + 3 * k);
  }
  return 0;
}
size_t oracle_submap_size(void* s, int cloud) { return static_cast<SubmapOracle*>(s)->cloud[cloud].size() / 3; }
const double* oracle_submap_data(void* s, int cloud) { return static_cast<SubmapOracle*>(s)->cloud[cloud].data(); }

}  // extern "C"
