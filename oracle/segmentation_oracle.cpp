/*
 * segmentation_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT) for "next" row SURVEY 8(f)-4: the three compute
 * steps of the segmentation nodelet.  Further down in this file: edge extraction (Segmentation::extractEdgePoint, ref:
 * :1144-1304) and object segmentation (DCVC, ref: :772-1112), each with its own header; they are PARITY UNPINNED by the
 * reference too, and pinned by literal Python transcriptions and an independent voxel-level model
 * (tests/test_edge_extract.py, tests/test_object_segmentation.py).  First part:
 * the multi-region ground extraction of the segmentation nodelet,
 *   Segmentation::groundRemove                    ref: src/models/segmentation/segmentation.cpp:738-770
 *   -> initSections / getSection                  ref: :174-238
 *   -> estimateRingsAndTimes2 (HDL_64E branch)    ref: :334-384
 *   -> filterByHeight                             ref: :454-470
 *   -> fillSectionIndex (cv::fastAtan2)           ref: :507-541
 *   -> segmentGroundThread / findBestPlane        ref: :626-730, :551-616
 * with the parameters of config/mapping/segmentation.yaml.
 *
 * PARITY UNPINNED by the reference (no tests, cannot be built here: Eigen / Open3D / OpenCV / ROS).  cv::fastAtan2 is
 * restated from OpenCV's published polynomial and pinned against cv2.fastAtan2 in tests/test_ground_extract.py.
 *
 * What is restated literally, quirks included:
 *   - initSections `continue`s past the angle increment once two successive ring radii differ by >= 5 m, so the table
 *     stalls and only TWO section bounds are ever produced; getSection then reads sectionBounds[2] out of range -- either
 *     outcome of that read returns section 2, which is what is restated;
 *   - the beam estimate counts quadrant 4 -> 1 transitions in point order, saturating at sensorModel - 1;
 *   - seeds are drawn from every 10th point of a region, the refits from every 5th ground point;
 *   - a region with <= 3 seeds is skipped ENTIRELY (`continue`): its points reach neither output;
 *   - theta == 360 (possible only through float underflow) drops the point.
 * Where the reference's result depends on an unspecified order, ONE order is fixed (and mirrored by the GPU):
 *   - output order: regions in (quadrant, section) order -- the reference appends per region under a mutex from four racing
 *     threads, so any interleaving of the quadrants can occur; within a region, point-index order as in the reference;
 *   - mean height: sum of z in chunks of 256 consecutive points, then over the chunk sums (the reference: one running
 *     sum; difference ~1e-13 m);
 *   - plane fit of the SEED set: summed in point-index order (the reference: ascending-z order from an unstable
 *     std::sort; difference ~1e-16 relative); the 20 lowest seeds are summed in ascending (z, index) order like the
 *     reference (ties have equal z, so the sum does not depend on them);
 *   - Vector4d::dot: ((a0 b0 + a1 b1) + a2 b2) + a3 b3.
 * Compiled with -ffp-contract=off; the CUDA kernels spell the same operations with round-to-nearest intrinsics, so
 * index lists are compared exactly.
 */
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "tloam_oracle.h"

namespace {

// cv::fastAtan2(y, x), degrees in [0, 360] (OpenCV modules/core/src/mathfuncs_core.simd.hpp)
float fast_atan2(float y, float x) {
  const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
  const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
  const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
  const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
  const float ax = std::fabs(x), ay = std::fabs(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// initSections, ref: :174-221.  Returns the section bounds actually produced (2 with the shipped configuration).
std::vector<float> init_sections(const oracle_ground_config& c) {
  std::vector<float> bounds;
  const int num_sec = c.num_sec;
  std::vector<int> boundary(num_sec);
  const int section_width = static_cast<int>(std::ceil(1.0 * c.sensor_model) / num_sec);   // :176
  for (int i = 0; i < num_sec; ++i) boundary[i] = section_width * (i + 1) - 1;
  double prev_radius = 0.0, angle = c.init_angle;
  int sec = 0;
  for (int i = 0; i < c.sensor_model; ++i) {
    if (c.sensor_model == 64 && i == 31) angle += 1.7;                                       // :194-196
    double cur = c.sensor_height / std::tan(std::fabs(angle) / 180.0 * M_PI);
    cur = cur < c.sensor_max_range ? cur : c.sensor_max_range;
    if (i >= 1) {
      const double dis = std::fabs(cur - prev_radius);
      if (dis >= 5.0 || dis <= 0.0) continue;                                                // :203-206 (skips everything below)
    }
    if (sec < num_sec && i == boundary[sec] && sec <= 3) {                                   // :209
      const double theta = std::fabs(angle / 180 * M_PI);
      if (theta != 0 && i < c.sensor_model) bounds.push_back(static_cast<float>(c.sensor_height / std::tan(theta)));
      else bounds.push_back(static_cast<float>(c.sensor_max_range));
      ++sec;
    }
    prev_radius = cur;
    angle += c.vertical_res;
  }
  return bounds;
}

int get_section(const std::vector<float>& bounds, int num_sec, double radius) {            // :230-238
  for (int i = 0; i < num_sec; ++i) {
    if (i >= static_cast<int>(bounds.size())) return i < num_sec - 1 ? num_sec - 1 : i;    // out-of-range read: only ever at the
    if (radius < bounds[i]) return i;                                                      // last index, where both outcomes give it
  }
  return num_sec - 1;
}

// findBestPlane, ref: :551-616 (sequential sums in list order)
void find_best_plane(const double* pts, const std::vector<int>& list, double out[4]) {
  const double n = static_cast<double>(list.size());
  double cx = 0, cy = 0, cz = 0;
  for (int i : list) { cx += pts[3 * i]; cy += pts[3 * i + 1]; cz += pts[3 * i + 2]; }
  cx /= n; cy /= n; cz /= n;
  double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
  for (int i : list) {
    const double rx = pts[3 * i] - cx, ry = pts[3 * i + 1] - cy, rz = pts[3 * i + 2] - cz;
    xx += rx * rx; xy += rx * ry; xz += rx * rz; yy += ry * ry; yz += ry * rz; zz += rz * rz;
  }
  xx /= n; xy /= n; xz /= n; yy /= n; yz /= n; zz /= n;
  double wx = 0, wy = 0, wz = 0;
  {
    const double det = yy * zz - yz * yz;
    const double ax = det, ay = xz * yz - xy * zz, az = xy * yz - xz * yy;
    double w = det * det;
    if ((wx * ax + wy * ay) + wz * az < 0.0) w = -w;
    wx += ax * w; wy += ay * w; wz += az * w;
  }
  {
    const double det = xx * zz - xz * xz;
    const double ax = xz * yz - xy * zz, ay = det, az = xy * xz - yz * xx;
    double w = det * det;
    if ((wx * ax + wy * ay) + wz * az < 0.0) w = -w;
    wx += ax * w; wy += ay * w; wz += az * w;
  }
  {
    const double det = xx * yy - xy * xy;
    const double ax = xy * yz - xz * yy, ay = xy * xz - yz * xx, az = det;
    double w = det * det;
    if ((wx * ax + wy * ay) + wz * az < 0.0) w = -w;
    wx += ax * w; wy += ay * w; wz += az * w;
  }
  // :608-613: a zero norm writes the zero plane, then the code falls through to normalize(), which Eigen 3.3 guards
  // (squaredNorm() > 0): the direction stays zero, d = -0, and every point of the region is within planeDis
  const double n2 = (wx * wx + wy * wy) + wz * wz;
  if (n2 > 0.0) { const double norm = std::sqrt(n2); wx /= norm; wy /= norm; wz /= norm; }
  out[0] = wx; out[1] = wy; out[2] = wz; out[3] = -((wx * cx + wy * cy) + wz * cz);
}

}  // namespace

extern "C" {

void oracle_ground_default_config(oracle_ground_config* c) {   // ref: config/mapping/segmentation.yaml
  c->sensor_model = 64; c->sensor_height = 1.73; c->vertical_res = 0.4; c->init_angle = -24.9;
  c->sensor_min_range = 1.0; c->sensor_max_range = 120.0;
  c->quadrant = 4; c->num_sec = 3; c->plane_dis = 0.3; c->max_iter = 3; c->ground_seed_num = 20;
}

float oracle_fast_atan2(float y, float x) { return fast_atan2(y, x); }

int oracle_ground_section_bounds(const oracle_ground_config* c, float* out, int cap) {
  const std::vector<float> b = init_sections(*c);
  for (int i = 0; i < static_cast<int>(b.size()) && i < cap; ++i) out[i] = b[i];
  return static_cast<int>(b.size());
}

// Segmentation::groundRemove on one scan (n points, AoS xyz).  Outputs (each buffer holds n entries):
//   ground_index / object_index: indices into the input cloud, in the order the reference's ground_scan / object_scan
//   receive the points (regions in (quadrant, section) order; object = region leftovers, then the height-filtered points);
//   beam: per-point beam estimate (what the reference stores in the intensity channel); region: q * num_sec + s, 12 = above
//   the height threshold, 13 = dropped; mean_height_out: the height threshold (mean z + 0.5); planes: 12 x 8 x 4 plane
//   models [region][iteration] (NaN where a region / iteration was skipped).  Any of beam / region / planes may be NULL.
int oracle_ground_extract(const double* pts, size_t n_, const oracle_ground_config* cfg, size_t* ground_index,
                          size_t* n_ground, size_t* object_index, size_t* n_object, int* beam, int* region,
                          double* mean_height_out, double* planes) {
  const oracle_ground_config& c = *cfg;
  const int n = static_cast<int>(n_);
  *n_ground = 0; *n_object = 0;
  if (c.quadrant != 4 || c.num_sec < 1 || c.num_sec > 3 || c.max_iter < 1 || c.max_iter > 8) return 1;
  const int nreg = 4 * c.num_sec;
  if (planes) for (int i = 0; i < 12 * 8 * 4; ++i) planes[i] = std::nan("");
  if (n == 0) { if (mean_height_out) *mean_height_out = 1.0; return 0; }                    // :335-338
  const std::vector<float> bounds = init_sections(c);                                        // step 1 (:740)

  // step 2 (:743): beams + mean height, HDL_64E branch (:341-384)
  std::vector<int> beams(n);
  {
    int prev_q = 0, b = 0;
    for (int i = 0; i < n; ++i) {
      const double x = pts[3 * i], y = pts[3 * i + 1];
      int q;
      if (x > 0 && y >= 0) q = 1;
      else if (x <= 0 && y > 0) q = 2;
      else if (x < 0 && y <= 0) q = 3;
      else q = 4;
      if (q == 1 && prev_q == 4 && b < c.sensor_model - 1) ++b;
      beams[i] = b;
      prev_q = q;
    }
  }
  double total = 0.0;
  for (int c0 = 0; c0 < n; c0 += 256) {
    double s = 0.0;
    for (int i = c0; i < std::min(n, c0 + 256); ++i) s += pts[3 * i + 2];
    total += s;
  }
  const double mean_height = total / static_cast<double>(n) + 0.5;                           // :743
  if (mean_height_out) *mean_height_out = mean_height;
  if (beam) std::memcpy(beam, beams.data(), sizeof(int) * n);

  // steps 2b + 3 (:744, :747): height filter, then quadrant x section of what is left
  std::vector<int> reg(n);
  std::vector<std::vector<int>> members(nreg);
  std::vector<int> above;
  for (int i = 0; i < n; ++i) {
    const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    if (z > mean_height) { reg[i] = 12; above.push_back(i); continue; }                      // :460-464
    const double r = std::sqrt(x * x + y * y);
    const float theta = fast_atan2(static_cast<float>(-y), static_cast<float>(x));           // :522
    const int s = get_section(bounds, c.num_sec, r);
    int q = -1;
    if (theta >= 0.0f && theta < 90.0f) q = 0;
    else if (theta >= 90.0f && theta < 180.0f) q = 1;
    else if (theta >= 180.0f && theta < 270.0f) q = 2;
    else if (theta >= 270.0f && theta < 360.0f) q = 3;
    if (q < 0) { reg[i] = 13; continue; }                                                   // :537-539
    reg[i] = q * c.num_sec + s;
    members[q * c.num_sec + s].push_back(i);
  }
  if (region) std::memcpy(region, reg.data(), sizeof(int) * n);

  // step 4 (:749-758): per region, ref: :626-730
  std::vector<size_t> ground, object;
  for (int rg = 0; rg < nreg; ++rg) {
    const std::vector<int>& m = members[rg];
    const int sz = static_cast<int>(m.size());
    struct ZK { double z; int k; };
    std::vector<ZK> info;
    for (int k = 0; k < sz; ++k) {                                                          // :641-647
      const double x = pts[3 * m[k]], y = pts[3 * m[k] + 1], z = pts[3 * m[k] + 2];
      const double r = std::sqrt((x * x + y * y) + z * z);
      if (k % 10 == 0 && z >= -1.5 * c.sensor_height && r >= c.sensor_min_range && r <= c.sensor_max_range) info.push_back({z, k});
    }
    std::vector<ZK> sorted = info;
    std::sort(sorted.begin(), sorted.end(), [](const ZK& a, const ZK& b) { return a.z < b.z || (a.z == b.z && a.k < b.k); });   // :649-651
    double sum_z = 0.0;
    int count = 0;
    for (size_t k = 0; k < sorted.size() && count < c.ground_seed_num; ++k, ++count) sum_z += sorted[k].z;   // :653-657
    const double av_height = count != 0 ? sum_z / count : 0;
    std::vector<int> cur;                                                                    // point ids of the current ground set
    for (const ZK& e : info)                                                                 // seeds, point-index order (see header)
      if (e.z < av_height + c.plane_dis) cur.push_back(m[e.k]);
    if (cur.size() <= 3) continue;                                                           // :665-666: the region is skipped
    std::vector<int> vertical;
    for (int iter = 0; iter < c.max_iter; ++iter) {
      if (cur.size() <= 3) continue;                                                         // :670-672
      double plane[4];
      find_best_plane(pts, cur, plane);
      if (planes) std::memcpy(planes + (rg * 8 + iter) * 4, plane, sizeof(plane));
      cur.clear(); vertical.clear();
      for (int i = 0; i < sz; ++i) {
        const double x = pts[3 * m[i]], y = pts[3 * m[i] + 1], z = pts[3 * m[i] + 2];
        const double dis = std::fabs(((plane[0] * x + plane[1] * y) + plane[2] * z) + plane[3] * 1.0);
        if (dis < c.plane_dis) {
          if (iter < c.max_iter - 1 && i % 5 == 0) cur.push_back(m[i]);
          else if (iter == c.max_iter - 1) cur.push_back(m[i]);
        } else if (iter == c.max_iter - 1) {
          vertical.push_back(m[i]);
        }
      }
    }
    for (int i : cur) ground.push_back(static_cast<size_t>(i));                              // :722-723
    for (int i : vertical) object.push_back(static_cast<size_t>(i));
  }
  for (int i : above) object.push_back(static_cast<size_t>(i));                              // :762
  std::copy(ground.begin(), ground.end(), ground_index);
  std::copy(object.begin(), object.end(), object_index);
  *n_ground = ground.size(); *n_object = object.size();
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// Edge extraction (ref: segmentation.cpp:1144-1304).  Compiled with -ffp-contract=off like the rest of this file: the
// device side uses __dadd_rn / __dmul_rn, so curvatures, the 0.1 threshold and the 0.05 gap test are bit-exact.
// Reference behaviour kept literally:
//   * points are distributed to beams by (int)intensity in input order (:1230-1237);  a beam id equal to sensorModel
//     passes the reference's `<=` test and writes one past the end of ringScans (undefined behaviour): dropped here;
//   * beams with fewer than ringMinNum points are skipped (:1240);
//   * curvature of ring point j (5 <= j < size - 5) = |sum of the 10 neighbours - 10 p_j|^2, summed left to right (:1248-1285);
//   * six sectors per beam; the sub-range handed to extractFromSection ends ONE BEFORE sector_end (iterator end is
//     exclusive, :1292), so the last curvature of every sector is in neither output;
//   * per sector: ascending sort by curvature; from the top, up to 20 picks with curvature > 0.1 (the 21st candidate ends
//     the loop), every pick marks its +-5 ring neighbours as picked until a gap^2 > 0.05 (:1153-1199); everything not
//     picked goes to the non-edge list in ascending-curvature order (:1202-1209).
// Left to the implementation by the reference and fixed here (identically on the device): std::sort is not stable ->
// ties are ordered by ring position; an empty sector (cannot happen with ringMinNum >= 16) is skipped instead of
// running the reference's `i <= size - 1` loop on an unsigned zero.
// ---------------------------------------------------------------------------------------------------------------------
#include <algorithm>

extern "C" int oracle_extract_edge(const double* pts, const double* intensity, size_t n, int sensor_model, int ring_min_num, int max_section,
                                   size_t* edge_index, size_t* n_edge, size_t* non_edge_index, size_t* n_non_edge) {
  *n_edge = 0; *n_non_edge = 0;
  if (sensor_model <= 0) return 0;
  std::vector<std::vector<size_t>> rings((size_t)sensor_model);
  for (size_t i = 0; i < n; ++i) {
    const int beam = (int)intensity[i];                                  // :1232
    if (beam >= 0 && beam < sensor_model) rings[(size_t)beam].push_back(i);
  }
  for (int r = 0; r < sensor_model; ++r) {
    const std::vector<size_t>& ring = rings[(size_t)r];
    const int total = (int)ring.size();
    if (total < ring_min_num) continue;                                  // :1240
    const int total_points = total - 10;                                 // :1246
    if (total_points <= 0) continue;
    auto P = [&](int j, int d) { return pts[3 * ring[(size_t)j] + (size_t)d]; };
    std::vector<double> curv((size_t)total_points);                      // curv[c] belongs to ring position c + 5
    for (int j = 5; j < total - 5; ++j) {
      double diff[3];
      for (int d = 0; d < 3; ++d)
        diff[d] = P(j - 5, d) + P(j - 4, d) + P(j - 3, d) + P(j - 2, d) + P(j - 1, d) - 10 * P(j, d) + P(j + 1, d) + P(j + 2, d) +
                  P(j + 3, d) + P(j + 4, d) + P(j + 5, d);
      curv[(size_t)(j - 5)] = diff[0] * diff[0] + diff[1] * diff[1] + diff[2] * diff[2];
    }
    for (int sct = 0; sct < 6; ++sct) {
      const int sector_length = total_points / 6;                        // :1288
      const int start = sector_length * sct;
      const int end = sct != 5 ? sector_length * (sct + 1) - 1 : total_points - 1;
      const int cnt = end - start;                                       // [start, end): exclusive end, :1292
      if (cnt <= 0) continue;
      if (max_section > 0 && cnt > max_section) return -1;
      std::vector<std::pair<double, int>> sub((size_t)cnt);              // (curvature, ring position)
      for (int c = 0; c < cnt; ++c) sub[(size_t)c] = {curv[(size_t)(start + c)], start + c + 5};
      std::sort(sub.begin(), sub.end());                                 // ascending; ties by ring position
      std::vector<char> picked((size_t)total, 0);
      int largest = 0;
      for (int i = cnt - 1; i >= 0; --i) {
        const int id = sub[(size_t)i].second;
        if (picked[(size_t)id]) continue;
        if (sub[(size_t)i].first <= 0.1) break;                          // :1161
        ++largest;
        picked[(size_t)id] = 1;
        if (largest <= 20) edge_index[(*n_edge)++] = ring[(size_t)id];   // :1169-1173
        else break;
        for (int k = 1; k <= 5; ++k) {                                   // :1177-1187
          double g = 0.0;
          {
            const double dx = P(id + k, 0) - P(id + k - 1, 0), dy = P(id + k, 1) - P(id + k - 1, 1), dz = P(id + k, 2) - P(id + k - 1, 2);
            g = dx * dx + dy * dy + dz * dz;
          }
          if (g > 0.05) break;
          picked[(size_t)(id + k)] = 1;
        }
        for (int k = -1; k >= -5; --k) {                                 // :1189-1199
          const double dx = P(id + k, 0) - P(id + k + 1, 0), dy = P(id + k, 1) - P(id + k + 1, 1), dz = P(id + k, 2) - P(id + k + 1, 2);
          if (dx * dx + dy * dy + dz * dz > 0.05) break;
          picked[(size_t)(id + k)] = 1;
        }
      }
      for (int i = 0; i < cnt; ++i) {                                    // :1202-1209
        const int id = sub[(size_t)i].second;
        if (!picked[(size_t)id]) non_edge_index[(*n_non_edge)++] = ring[(size_t)id];
      }
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Object segmentation: Dynamic Curved-Voxel Clustering (ref: segmentation.cpp:772-1112: getPolarIndex, convertToPolar,
// createHashTable, searchKNN, DCVC, labelAnalysis, colorSegmentation, objectSegmentation).  LITERAL restatement: the
// label vector is relabelled by full sweeps exactly like the reference (`for (auto& seg : label_info)`), the hash table
// is an unordered_map from the reference's integer voxel index to the point list.  Kept literally:
//   * min/max pitch and polar start from the member values left by resetParams (0.0; 5.0 for the polar pair on the very
//     first frame, segmentation.hpp:330-333) -> cfg.{min,max}_{pitch,polar}_init;
//   * points outside (sensorMinRange, sensorMaxRange) keep the value-initialised polar triple (0, 0, 0) and are still
//     hashed and clustered (polarCor.resize leaves Eigen vectors uninitialised in the reference: zero here);
//   * searchKNN: pitch layers > height and polar rings > polarNum are skipped, azimuth -1 wraps to width - 1 but
//     azimuth > 300 is CLAMPED to the literal 300 (so a voxel can be listed twice, and a voxel in pitch layer height + 1
//     is not its own neighbour);
//   * DCVC visits only points that are still unlabelled; neighbours seen before the first labelled one stay unlabelled.
// Fixed where the reference leaves the order to the implementation (std::sort over unordered_map iteration order):
// clusters of equal size are ordered by their smallest point index.  Trigonometry is libm's (std::asin, std::atan2);
// oracle_dcvc_from_polar takes the polar triples as input so that the integer part can be checked bit-exactly whatever
// the last bit of asin / atan2 is.
// ---------------------------------------------------------------------------------------------------------------------
#include <unordered_map>
#include <limits>

extern "C" {

void oracle_dcvc_default_config(oracle_dcvc_config* c) {   // ref: config/mapping/segmentation.yaml
  c->start_r = 0.35; c->delta_r = 0.0004; c->delta_p = 1.2; c->delta_a = 1.2; c->min_seg = 80;
  c->sensor_min_range = 1.0; c->sensor_max_range = 120.0;
  c->min_pitch_init = 0.0; c->max_pitch_init = 0.0; c->min_polar_init = 0.0; c->max_polar_init = 0.0;
}

// convertToPolar, first half (:790-822): polar triples + the four extrema.  polar: [n][3]; ext: minPitch, maxPitch, minPolar, maxPolar
void oracle_dcvc_polar(const double* pts, size_t n, const oracle_dcvc_config* c, double* polar, double* ext) {
  double min_pitch = c->min_pitch_init, max_pitch = c->max_pitch_init, min_polar = c->min_polar_init, max_polar = c->max_polar_init;
  for (size_t i = 0; i < n; ++i) {
    const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    const double r = std::sqrt((x * x + y * y) + z * z);                              // cur.norm()
    const double pitch = std::asin(z / r) * 180.0 / M_PI;
    const double angle = std::atan2(y, x);
    const double az = angle > 0.0 ? angle * 180 / M_PI : (angle + 2 * M_PI) * 180 / M_PI;
    polar[3 * i] = polar[3 * i + 1] = polar[3 * i + 2] = 0.0;
    if (r >= c->sensor_max_range || r <= c->sensor_min_range) continue;                // :808-809
    min_pitch = pitch < min_pitch ? pitch : min_pitch;
    max_pitch = pitch > max_pitch ? pitch : max_pitch;
    min_polar = r < min_polar ? r : min_polar;
    max_polar = r > max_polar ? r : max_polar;
    polar[3 * i] = r; polar[3 * i + 1] = pitch; polar[3 * i + 2] = az;
  }
  ext[0] = min_pitch; ext[1] = max_pitch; ext[2] = min_polar; ext[3] = max_polar;
}

// everything after the trigonometry.  root[i] = smallest point index of i's DCVC class (a canonical form of label_info);
// cluster[i] = 1-based position of i's class in labelRecords, 0 if the class has <= minSeg points; seg_index = the
// segmented scan (indices, cluster by cluster); sizes / boxes: [n_clusters], [n_clusters][6] (centre xyz, dimensions xyz).
// Returns 0; -1 if polarNum exceeds max_bounds (device limit mirrored; <= 0: no limit).
int oracle_dcvc_from_polar(const double* pts, const double* polar, const double* ext, size_t n_, const oracle_dcvc_config* c,
                           int max_bounds, int* voxel, int* root, int* cluster, size_t* seg_index, size_t* n_seg, int* n_clusters,
                           int* sizes, double* boxes) {
  const int n = (int)n_;
  *n_seg = 0; *n_clusters = 0;
  if (n == 0) return 0;
  const double min_pitch = ext[0], max_pitch = ext[1], min_polar = ext[2], max_polar = ext[3];
  // convertToPolar, second half (:825-836)
  int polar_num = 0;
  std::vector<double> bounds;
  const int width = (int)(std::round(360.0 / c->delta_a) + 1);
  const int height = (int)((max_pitch - min_pitch) / c->delta_p);
  {
    double range = min_polar;
    int step = 1;
    while (range <= max_polar) {
      range += (c->start_r - step * c->delta_r);
      bounds.push_back(range);
      polar_num++, step++;
      if (max_bounds > 0 && polar_num > max_bounds) return -1;
      if (polar_num > (1 << 24)) return -1;                                              // non-terminating configuration
    }
  }
  auto polar_index = [&](double radius) {                                                // :777-784
    for (int r = 0; r < polar_num; ++r)
      if (radius < bounds[(size_t)r]) return r;
    return polar_num - 1;
  };
  // createHashTable (:843-874)
  std::unordered_map<int, std::vector<int>> voxel_map;
  std::vector<int> pi((size_t)n), ti((size_t)n), ai((size_t)n), vi((size_t)n);
  for (int i = 0; i < n; ++i) {
    pi[(size_t)i] = polar_index(polar[3 * i]);
    ti[(size_t)i] = (int)std::round((polar[3 * i + 1] - min_pitch) / c->delta_p);
    ai[(size_t)i] = (int)std::round(polar[3 * i + 2] / c->delta_a);
    vi[(size_t)i] = (ai[(size_t)i] * (polar_num + 1) + pi[(size_t)i]) + ti[(size_t)i] * (polar_num + 1) * (width + 1);
    voxel_map[vi[(size_t)i]].push_back(i);
    if (voxel) voxel[i] = vi[(size_t)i];
  }
  // DCVC (:915-990)
  std::vector<int> label((size_t)n, -1);
  int label_count = 0;
  std::vector<int> knn, neighbors;
  for (int i = 0; i < n; ++i) {
    if (label[(size_t)i] != -1) continue;
    knn.clear(); neighbors.clear();
    const int p = pi[(size_t)i], t = ti[(size_t)i], a = ai[(size_t)i];
    for (int z = t - 1; z <= t + 1; ++z) {                                               // searchKNN (:886-908)
      if (z < 0 || z > height) continue;
      for (int y = p - 1; y <= p + 1; ++y) {
        if (y < 0 || y > polar_num) continue;
        for (int x = a - 1; x <= a + 1; ++x) {
          int ax = x;
          if (ax < 0) ax = width - 1;
          if (ax > 300) ax = 300;
          knn.push_back((ax * (polar_num + 1) + y) + z * (polar_num + 1) * (width + 1));
        }
      }
    }
    for (int k : knn) {
      auto it = voxel_map.find(k);
      if (it != voxel_map.end()) neighbors.insert(neighbors.end(), it->second.begin(), it->second.end());
    }
    for (int id : neighbors) {                                                           // :952-968
      const int curr = label[(size_t)i], neigh = label[(size_t)id];
      if (curr != -1 && neigh != -1 && curr != neigh) {
        for (int& seg : label)
          if (seg == curr) seg = neigh;
      } else if (neigh != -1) {
        label[(size_t)i] = neigh;
      } else if (curr != -1) {
        label[(size_t)id] = curr;
      }
    }
    if (label[(size_t)i] == -1) {                                                        // :972-978
      label_count++;
      label[(size_t)i] = label_count;
      for (int id : neighbors) label[(size_t)id] = label_count;
    }
  }
  // labelAnalysis (:998-1025) with the tie rule of the header
  std::unordered_map<int, std::vector<int>> hist;
  for (int i = 0; i < n; ++i) hist[label[(size_t)i]].push_back(i);
  if (root)
    for (auto& kv : hist)
      for (int i : kv.second) root[i] = kv.second.front();
  std::vector<const std::vector<int>*> recs;
  for (auto& kv : hist) recs.push_back(&kv.second);
  std::sort(recs.begin(), recs.end(), [](const std::vector<int>* a, const std::vector<int>* b) {
    return a->size() > b->size() || (a->size() == b->size() && a->front() < b->front());
  });
  if (cluster) std::fill(cluster, cluster + n, 0);
  int count = 0;
  for (const std::vector<int>* rec : recs) {
    if (!((int)rec->size() > c->min_seg)) continue;                                      // :1016
    // colorSegmentation (:1032-1078)
    double mn[3] = {std::numeric_limits<double>::max(), std::numeric_limits<double>::max(), std::numeric_limits<double>::max()};
    double mx[3] = {-std::numeric_limits<double>::max(), -std::numeric_limits<double>::max(), -std::numeric_limits<double>::max()};
    for (int id : *rec) {
      seg_index[(*n_seg)++] = (size_t)id;
      if (cluster) cluster[id] = count + 1;
      for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], pts[3 * id + d]); mx[d] = std::max(mx[d], pts[3 * id + d]); }
    }
    if (sizes) sizes[count] = (int)rec->size();
    if (boxes)
      for (int d = 0; d < 3; ++d) {
        const double len = mx[d] - mn[d];
        boxes[6 * count + d] = mn[d] + len / 2.0;
        boxes[6 * count + 3 + d] = len < 0 ? -1 * len : len;
      }
    ++count;
  }
  *n_clusters = count;
  return 0;
}

int oracle_dcvc(const double* pts, size_t n, const oracle_dcvc_config* c, int max_bounds, double* polar_out, int* voxel, int* root,
                int* cluster, size_t* seg_index, size_t* n_seg, int* n_clusters, int* sizes, double* boxes) {
  std::vector<double> polar(3 * n + 3);
  double ext[4];
  oracle_dcvc_polar(pts, n, c, polar.data(), ext);
  if (polar_out) { std::memcpy(polar_out, polar.data(), sizeof(double) * 3 * n); std::memcpy(polar_out + 3 * n, ext, sizeof(ext)); }
  return oracle_dcvc_from_polar(pts, polar.data(), ext, n, c, max_bounds, voxel, root, cluster, seg_index, n_seg, n_clusters, sizes, boxes);
}

}  // extern "C"
