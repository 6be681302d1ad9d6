/*
 * feature_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT) for "next" row SURVEY 8(f)-2:
 * featureExtract::calculatePCAInfo / extractPlanarSphere
 * (ref: src/models/feature_extraction/feature_extract.cpp:47-122, 133-197; parameters
 *  config/mapping/feature.yaml; caller src/front_end/front_end.cpp:181-199).
 *
 * PARITY UNPINNED (same reason as tloam_oracle.h): the reference has no tests or fixtures for this function and
 * cannot be built here.  Eigen::SelfAdjointEigenSolver<Matrix3d>::compute is restated as a cyclic Jacobi iteration
 * (eigenvalues agree with any backward-stable solver to ~1e-16 |cov|).
 *
 * This translation unit is compiled with -ffp-contract=off and the CUDA kernel spells out the same operation order
 * with round-to-nearest intrinsics, so that GPU and oracle results are bit-identical and the index lists can be
 * compared exactly.
 */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

#include <omp.h>

#include "tloam_oracle.h"

namespace {

// eigenvalues ascending + eigenvector of the smallest one; cyclic Jacobi, fixed sweep/rotation order
void jacobi3(const double c[6] /* xx xy xz yy yz zz */, double eig[3], double nvec[3]) {
  double a[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 32; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = cs * akp - sn * akq;
          a[k][q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = cs * apk - sn * aqk;
          a[q][k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = cs * vkp - sn * vkq;
          v[k][q] = sn * vkp + cs * vkq;
        }
      }
  }
  // ascending order, ties keep the lower axis first
  int o0 = 0, o1 = 1, o2 = 2;
  if (a[o1][o1] < a[o0][o0]) std::swap(o0, o1);
  if (a[o2][o2] < a[o1][o1]) std::swap(o1, o2);
  if (a[o1][o1] < a[o0][o0]) std::swap(o0, o1);
  eig[0] = a[o0][o0]; eig[1] = a[o1][o1]; eig[2] = a[o2][o2];
  for (int r = 0; r < 3; ++r) nvec[r] = v[r][o0];
}

}  // namespace

extern "C" {

void oracle_feature_default_config(oracle_feature_config* c) {   // ref: config/mapping/feature.yaml
  c->radius = 0.2; c->K = 20; c->min_neigh = 10; c->planar_num = 500; c->sphere_num = 300;
  c->cvr_scan = 0.25; c->cvr_submap = 0.15; c->planar_scan_thres = 0.75; c->planar_submap_thres = 0.65;
  c->planar_vertic_thres = 0.25;
}

// ref: feature_extract.cpp:47-122.  Points with <= min_neigh neighbours keep the value-initialised PCAInfo
// (all zero, no neighbours), like the `continue` at :70-71.
int oracle_pca_info(const double* pts, size_t n, const oracle_feature_config* c, double* cvr, double* flatness,
                    double* sphericity, double* normal, int* num_sum, int* neigh) {
  const int K = c->K;
  if (n == 0 || K < 3 || !(c->radius >= 0.0)) return 1;      // :49-54
  std::vector<int> idx(n * K), cnt(n);
  std::vector<double> d2(n * K);
  oracle_knn(pts, n, pts, n, c->radius, K, idx.data(), d2.data(), cnt.data(), 0);   // SearchHybrid(cur_pt, r, K), :67
#pragma omp parallel for schedule(dynamic, 256) num_threads(std::min(omp_get_max_threads(), 32))
  for (long long i = 0; i < (long long)n; ++i) {
    cvr[i] = flatness[i] = sphericity[i] = 0.0;
    normal[3 * i] = normal[3 * i + 1] = normal[3 * i + 2] = 0.0;
    num_sum[i] = 0;
    for (int j = 0; j < K; ++j) neigh[i * K + j] = -1;
    const int m = cnt[i];
    if (m <= 0 || m <= c->min_neigh) continue;               // :67-71
    double cum[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < m; ++j) {                             // :77-88, neighbours in ascending distance
      const double* p = pts + 3 * (size_t)idx[i * K + j];
      cum[0] += p[0]; cum[1] += p[1]; cum[2] += p[2];
      cum[3] += p[0] * p[0]; cum[4] += p[0] * p[1]; cum[5] += p[0] * p[2];
      cum[6] += p[1] * p[1]; cum[7] += p[1] * p[2]; cum[8] += p[2] * p[2];
    }
    for (int k = 0; k < 9; ++k) cum[k] /= (double)m;          // :89
    const double cov[6] = {cum[3] - cum[0] * cum[0], cum[4] - cum[0] * cum[1], cum[5] - cum[0] * cum[2],
                           cum[6] - cum[1] * cum[1], cum[7] - cum[1] * cum[2], cum[8] - cum[2] * cum[2]};   // :90-98
    double ev[3], nv[3];
    jacobi3(cov, ev, nv);                                     // :100-104
    const double sum = (ev[0] + ev[1]) + ev[2];
    cvr[i] = (sum == 0.0) ? 0.0 : ev[0] / sum;                // :106-111
    flatness[i] = (ev[1] - ev[0]) / ev[2];                    // :113
    sphericity[i] = ev[0] / ev[2];                            // :114
    normal[3 * i] = nv[0]; normal[3 * i + 1] = nv[1]; normal[3 * i + 2] = nv[2];
    num_sum[i] = m;
    for (int j = 0; j < m; ++j) neigh[i * K + j] = idx[i * K + j];   // :115
  }
  return 0;
}

// ref: feature_extract.cpp:133-197.  Outputs hold up to n entries each.
//  * std::sort is not stable in the reference; ties in flatness are ordered by ascending point index here.
//  * QUIRK FE-1 (:181-185): the sphere lists receive the RANK `id`, not `sphere_info[id].second`, so they are
//    0..count-1; QUIRK FE-2 (:156, :182): sphere candidates carry and are thresholded by FLATNESS against cvr_scan.
//    sphere_candidates (optional) returns the point indices the ranks refer to.
int oracle_extract_planar_sphere(const double* pts, size_t n, const oracle_feature_config* c, size_t* planar_scan,
                                 size_t* n_planar_scan, size_t* planar_submap, size_t* n_planar_submap,
                                 size_t* sphere_scan, size_t* n_sphere_scan, size_t* sphere_submap,
                                 size_t* n_sphere_submap, size_t* sphere_candidates) {
  *n_planar_scan = *n_planar_submap = *n_sphere_scan = *n_sphere_submap = 0;
  const int K = c->K;
  std::vector<double> cvr(n), flat(n), sph(n), nrm(3 * n);
  std::vector<int> num(n), neigh(n * (size_t)K);
  if (oracle_pca_info(pts, n, c, cvr.data(), flat.data(), sph.data(), nrm.data(), num.data(), neigh.data()) != 0)
    return 0;                                                 // calculatePCAInfo returned false: nothing selected, :141
  std::vector<std::pair<double, size_t>> planar, sphere;
  for (size_t id = 0; id < n; ++id) {                         // :148-164
    if (flat[id] > c->planar_submap_thres && std::fabs(nrm[3 * id + 2]) < c->planar_vertic_thres) {
      planar.emplace_back(flat[id], id);
    } else if (cvr[id] > c->cvr_submap) {
      bool max_uniform = true;
      for (int j = 0; j < K; ++j) {
        const int item = neigh[id * K + j];
        if (item < 0) break;
        if (cvr[id] < cvr[item]) { max_uniform = false; break; }
      }
      if (max_uniform) sphere.emplace_back(flat[id], id);
    }
  }
  auto desc = [](const std::pair<double, size_t>& a, const std::pair<double, size_t>& b) {
    return a.first > b.first || (a.first == b.first && a.second < b.second);
  };
  std::sort(planar.begin(), planar.end(), desc);              // :167-173
  std::sort(sphere.begin(), sphere.end(), desc);
  for (size_t id = 0; id < planar.size(); ++id) {             // :177-181
    if (id < (size_t)c->planar_num || planar[id].first > c->planar_scan_thres) planar_scan[(*n_planar_scan)++] = planar[id].second;
    planar_submap[(*n_planar_submap)++] = planar[id].second;
  }
  for (size_t id = 0; id < sphere.size(); ++id) {             // :183-188
    if (id < (size_t)c->sphere_num || sphere[id].first > c->cvr_scan) sphere_scan[(*n_sphere_scan)++] = id;
    sphere_submap[(*n_sphere_submap)++] = id;
    if (sphere_candidates) sphere_candidates[id] = sphere[id].second;
  }
  return 1;
}

}  // extern "C"
