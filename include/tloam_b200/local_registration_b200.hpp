// local_registration_b200.hpp -- C++ shim: tloam::LocalRegistrationB200, a drop-in RegistrationInterface
// (ref: include/tloam/models/registration/registration_interface.hpp:40-48) implemented on the C ABI of
// libtloam_b200.so.  Header-only; it needs only `points_` (contiguous std::vector<Eigen::Vector3d>,
// ref: include/tloam/open3d/PointCloud2.hpp:396) and `Eigen::Isometry3d::matrix().data()` from the host side.
//
// In the reference tree it is used by changing ONE line of FrontEnd::initRegistraton
// (ref: src/front_end/front_end.cpp:160):
//     registration_ptr_ = std::make_shared<LocalRegistrationB200>(config_node["TLS"]);
// (see INTEGRATION.md).  When the reference headers are not available (this repository's tests), define
// TLOAM_B200_MOCK_HOST_TYPES and include a header that provides tloam::Frame / tloam::RegistrationInterface /
// Eigen::Isometry3d with the same members (tests/mock/mock_tloam.hpp).
#ifndef TLOAM_B200_LOCAL_REGISTRATION_B200_HPP
#define TLOAM_B200_LOCAL_REGISTRATION_B200_HPP

#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>

#include "../tloam_b200.h"

#ifndef TLOAM_B200_MOCK_HOST_TYPES
#include <yaml-cpp/yaml.h>
#include "tloam/models/registration/registration_interface.hpp"
#endif

namespace tloam {

class LocalRegistrationB200 : public RegistrationInterface {
 public:
  // From an already-filled POD (same 16 fields as the YAML "TLS:" block).
  explicit LocalRegistrationB200(const tloam_tls_config& cfg, int device = 0, void* stream = nullptr) { create(cfg, device, stream); }

#ifndef TLOAM_B200_MOCK_HOST_TYPES
  // Same constructor signature as tloam::LocalRegistration (ref: registration.cpp:182-230): reads the TLS node.
  explicit LocalRegistrationB200(const YAML::Node& node, int device = 0, void* stream = nullptr) {
    tloam_tls_config c;
    tloam_b200_default_config(&c);
    c.k_corr = node["k_corr"].as<int>();
    c.factor_num = node["factor_num"].as<int>();
    c.edge_dist_thres = node["edge_dist_thres"].as<double>();
    c.sphere_dist_thres = node["sphere_dist_thres"].as<double>();
    c.planar_dist_thres = node["planar_dist_thres"].as<double>();
    c.ground_dist_thres = node["ground_dist_thres"].as<double>();
    c.edge_dir_thres = node["edge_dir_thres"].as<double>();
    c.edge_maxnum = node["edge_maxnum"].as<int>();
    c.sphere_maxnum = node["sphere_maxnum"].as<int>();
    c.planar_maxnum = node["planar_maxnum"].as<int>();
    c.ground_maxnum = node["ground_maxnum"].as<int>();
    c.max_iterations = node["max_iterations"].as<int>();
    c.cost_threshold = node["cost_threshold"].as<double>();
    c.gnc_factor = node["gnc_factor"].as<double>();
    c.noise_bound = node["noise_bound"].as<double>();
    c.fitness_thres = node["fitness_thres"].as<double>();
    // extras of the POD (not in the reference YAML): optional keys, Ceres defaults otherwise
    if (node["initial_trust_region_radius"]) c.initial_trust_region_radius = node["initial_trust_region_radius"].as<double>();
    create(c, device, stream);
  }
#endif

  ~LocalRegistrationB200() override { tloam_b200_destroy(h_); }
  LocalRegistrationB200(const LocalRegistrationB200&) = delete;
  LocalRegistrationB200& operator=(const LocalRegistrationB200&) = delete;

  // ref: registration.cpp:232-239.  The reference aliases the caller's shared_ptrs; this copies to the device.
  bool setInputSource(Frame& cloud_in_) override { return report(set(cloud_in_, true), "setInputSource"); }
  // ref: registration.cpp:241-248 (+ the KD-tree build of :888-915, done once per map here).
  bool setInputTarget(Frame& cloud_in_) override { return report(set(cloud_in_, false), "setInputTarget"); }

  // ref: registration.cpp:879-1133.  predict_pose_ / result_pose_ are Eigen::Isometry3d (4x4 column-major).
  bool scanMatching(Frame& out_result_, Eigen::Isometry3d& predict_pose_, Eigen::Isometry3d& result_pose_) override {
    double result[16];
    const int rc = tloam_b200_scan_match(h_, predict_pose_.matrix().data(), result, nullptr);
    if (rc != TLOAM_B200_OK) return report(rc, "scanMatching");
    for (int i = 0; i < 16; ++i) result_pose_.matrix().data()[i] = result[i];
    // ref: registration.cpp:1126-1128 -- transform the (normally empty) scan cloud of the result frame
    if (out_result_.scan_cloud && !out_result_.scan_cloud->points_.empty()) {
      for (auto& p : out_result_.scan_cloud->points_) {
        const double x = p[0], y = p[1], z = p[2];
        p[0] = result[0] * x + result[4] * y + result[8] * z + result[12];
        p[1] = result[1] * x + result[5] * y + result[9] * z + result[13];
        p[2] = result[2] * x + result[6] * y + result[10] * z + result[14];
      }
    }
    return true;
  }

  // (f)-3: scanMatching with the front end's constant-velocity prediction (ref: front_end.cpp:329-330) computed on
  // the device from the two last results; `FrontEnd::updateLidarOdometry` then needs no predicate_pose at all.
  bool scanMatchingPredicted(Eigen::Isometry3d& result_pose_) {
    double result[16];
    const int rc = tloam_b200_scan_match_predicted(h_, result, nullptr);
    if (rc != TLOAM_B200_OK) return report(rc, "scanMatchingPredicted");
    for (int i = 0; i < 16; ++i) result_pose_.matrix().data()[i] = result[i];
    return true;
  }

  std::pair<double, double> getFitnessScore() override {   // ref: registration.cpp:257-296
    double f = 0.0, e = 0.0;
    report(tloam_b200_fitness(h_, &f, &e), "getFitnessScore");
    return std::make_pair(f, e);
  }

  Eigen::Isometry3d getTransform() {                        // ref: registration.cpp:370-372
    Eigen::Isometry3d T;
    report(tloam_b200_get_transform(h_, T.matrix().data()), "getTransform");
    return T;
  }
  Eigen::Isometry3d getPoseIncrement() {                    // ref: registration.cpp:374-376
    Eigen::Isometry3d T;
    report(tloam_b200_get_pose_increment(h_, T.matrix().data()), "getPoseIncrement");
    return T;
  }
  void resetKDTree() {}                                     // ref: registration.cpp:250-255 -- the map lives on the device

  int lastStatus() const { return last_status_; }
  tloam_b200_handle* handle() { return h_; }

 private:
  void create(const tloam_tls_config& cfg, int device, void* stream) {
    const int rc = tloam_b200_create(&cfg, device, stream, &h_);
    if (rc != TLOAM_B200_OK) throw std::runtime_error(std::string("tloam_b200_create: ") + tloam_b200_status_string(rc));
  }
  // order at the ABI: edge, sphere, planar, ground (ref: registration.cpp:233-236, 242-245)
  int set(Frame& f, bool source) {
    const double* xyz[4] = {data(f.edge_feature), data(f.sphere_feature), data(f.planar_feature), data(f.ground_feature)};
    const size_t n[4] = {size(f.edge_feature), size(f.sphere_feature), size(f.planar_feature), size(f.ground_feature)};
    return source ? tloam_b200_set_source(h_, xyz, n) : tloam_b200_set_target(h_, xyz, n);
  }
  template <class CloudPtr>
  static const double* data(const CloudPtr& c) {
    return (c && !c->points_.empty()) ? reinterpret_cast<const double*>(c->points_.data()) : nullptr;
  }
  template <class CloudPtr>
  static size_t size(const CloudPtr& c) { return c ? c->points_.size() : 0; }
  // the reference returns true unconditionally and logs; keep `bool`, remember the status, log to stderr
  bool report(int rc, const char* where) {
    last_status_ = rc;
    if (rc != TLOAM_B200_OK)
      std::fprintf(stderr, "[tloam_b200] %s: %s %s\n", where, tloam_b200_status_string(rc), tloam_b200_last_error(h_));
    return rc == TLOAM_B200_OK;
  }
  tloam_b200_handle* h_ = nullptr;
  int last_status_ = TLOAM_B200_OK;
};

}  // namespace tloam
#endif
