// Minimal stand-ins for the host-side types the C++ shim touches, so that it can be compiled and exercised in
// an image without Eigen / Open3D / ROS.  Written for this repository's tests; only the members the shim uses
// exist: Vector3d = 3 contiguous doubles with operator[], Isometry3d::matrix().data() = 16 doubles column-major,
// PointCloud2::points_ / intensity_, Frame's five shared_ptrs, CloudData::cloud_ptr, RegistrationInterface's four virtuals.
#pragma once
#include <array>
#include <memory>
#include <utility>
#include <vector>

namespace Eigen {
struct Vector3d {
  double v[3];
  double& operator[](int i) { return v[i]; }
  const double& operator[](int i) const { return v[i]; }
};
static_assert(sizeof(Vector3d) == 24, "Vector3d must be 3 packed doubles");
struct Matrix4dStorage {
  double m[16];
  double* data() { return m; }
  const double* data() const { return m; }
};
struct Isometry3d {
  Matrix4dStorage s;
  Isometry3d() { for (int i = 0; i < 16; ++i) s.m[i] = (i % 5 == 0) ? 1.0 : 0.0; }
  Matrix4dStorage& matrix() { return s; }
  const Matrix4dStorage& matrix() const { return s; }
};
}  // namespace Eigen

namespace open3d { namespace geometry {
struct PointCloud2 { std::vector<Eigen::Vector3d> points_; std::vector<double> intensity_; };   // ref: PointCloud2.hpp:396-408
}}  // namespace open3d::geometry

namespace tloam {
struct Frame {
  Frame() : scan_cloud(new open3d::geometry::PointCloud2), edge_feature(new open3d::geometry::PointCloud2),
            sphere_feature(new open3d::geometry::PointCloud2), planar_feature(new open3d::geometry::PointCloud2),
            ground_feature(new open3d::geometry::PointCloud2) {}
  std::shared_ptr<open3d::geometry::PointCloud2> scan_cloud, edge_feature, sphere_feature, planar_feature, ground_feature;
};
struct CloudData {   // ref: include/tloam/models/utils/sensor_data.hpp:17-43 (the members the shims touch)
  CloudData() : time(0.0), cloud_ptr(new open3d::geometry::PointCloud2) {}
  double time;
  std::shared_ptr<open3d::geometry::PointCloud2> cloud_ptr;
};
class RegistrationInterface {
 public:
  virtual ~RegistrationInterface() = default;
  virtual bool setInputSource(Frame&) = 0;
  virtual bool setInputTarget(Frame&) = 0;
  virtual bool scanMatching(Frame&, Eigen::Isometry3d&, Eigen::Isometry3d&) = 0;
  virtual std::pair<double, double> getFitnessScore() = 0;
};
}  // namespace tloam
