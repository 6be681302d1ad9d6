// feature_extract_b200.hpp -- C++ shim: tloam::featureExtractB200, same public surface as tloam::featureExtract
// (ref: include/tloam/models/feature_extraction/feature_extract.hpp:29-75) for the one member the front end calls,
// extractPlanarSphere (ref: src/front_end/front_end.cpp:194, :289), implemented on the C ABI of libtloam_b200.so
// ("next" row 8(f)-2).  Header-only; from the host side it needs only CloudData::cloud_ptr->points_
// (contiguous std::vector<Eigen::Vector3d>, ref: include/tloam/models/utils/sensor_data.hpp:17-43).
//
// In the reference tree: replace `feature_extract_ptr_ = std::make_shared<featureExtract>();`
// (ref: src/front_end/front_end.cpp:34) by `std::make_shared<featureExtractB200>(registration handle or nullptr)`.
// Without the reference headers (this repository's tests) define TLOAM_B200_MOCK_HOST_TYPES and provide
// tloam::CloudData with the same member (tests/mock/mock_tloam.hpp).
#ifndef TLOAM_B200_FEATURE_EXTRACT_B200_HPP
#define TLOAM_B200_FEATURE_EXTRACT_B200_HPP

#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../tloam_b200.h"

#ifndef TLOAM_B200_MOCK_HOST_TYPES
#include <yaml-cpp/yaml.h>
#include "tloam/models/utils/sensor_data.hpp"
#include "tloam/models/utils/work_space_path.h"
#endif

namespace tloam {

class featureExtractB200 {
 public:
  // `shared` = a handle that already exists (e.g. LocalRegistrationB200::handle()): the extraction then runs on its
  // device and stream.  nullptr = own handle on `device`.
  explicit featureExtractB200(const tloam_feature_config& cfg, tloam_b200_handle* shared = nullptr, int device = 0)
      : cfg_(cfg) { attach(shared, device); }

#ifndef TLOAM_B200_MOCK_HOST_TYPES
  // Same behaviour as featureExtract::featureExtract() + initConfig (ref: feature_extract.cpp:13-37): reads
  // config/mapping/feature.yaml.
  explicit featureExtractB200(tloam_b200_handle* shared = nullptr, int device = 0) {
    const YAML::Node node = YAML::LoadFile(WORK_SPACE_PATH + "/config/mapping/feature.yaml")["feature"];
    tloam_b200_feature_default_config(&cfg_);
    cfg_.radius = node["radius"].as<double>();
    cfg_.K = node["K"].as<int>();
    cfg_.planar_num = node["planar_num"].as<int>();
    cfg_.sphere_num = node["sphere_num"].as<int>();
    cfg_.min_neigh = node["min_neigh"].as<int>();
    cfg_.cvr_scan = node["cvr_scan"].as<double>();
    cfg_.cvr_submap = node["cvr_submap"].as<double>();
    cfg_.planar_scan_thres = node["planar_scan_thres"].as<double>();
    cfg_.planar_submap_thres = node["planar_submap_thres"].as<double>();
    cfg_.planar_vertic_thres = node["planar_vertic_thres"].as<double>();
    attach(shared, device);
  }
#endif

  ~featureExtractB200() { if (own_) tloam_b200_destroy(h_); }
  featureExtractB200(const featureExtractB200&) = delete;
  featureExtractB200& operator=(const featureExtractB200&) = delete;

  // ref: feature_extract.cpp:133-197.  Appends to the four lists like the reference's emplace_back; returns true
  // like the reference (which returns true unconditionally) unless the device call failed.
  bool extractPlanarSphere(CloudData& cloud_in_, std::vector<size_t>& planar_scan_index,
                           std::vector<size_t>& planar_submap_index, std::vector<size_t>& sphere_scan_index,
                           std::vector<size_t>& sphere_submap_index) {
    const auto& pts = cloud_in_.cloud_ptr->points_;
    const size_t n = pts.size();
    if (n == 0) return true;                      // calculatePCAInfo logs and fails: nothing appended (:49-54, :141)
    buf_[0].resize(n); buf_[1].resize(n); buf_[2].resize(n); buf_[3].resize(n);
    size_t cnt[4] = {0, 0, 0, 0};
    last_status_ = tloam_b200_extract_planar_sphere(h_, &cfg_, reinterpret_cast<const double*>(pts.data()), n, buf_[0].data(),
                                                    &cnt[0], buf_[1].data(), &cnt[1], buf_[2].data(), &cnt[2],
                                                    buf_[3].data(), &cnt[3], nullptr);
    if (last_status_ != TLOAM_B200_OK) {
      std::fprintf(stderr, "[tloam_b200] extractPlanarSphere: %s %s\n", tloam_b200_status_string(last_status_),
                   tloam_b200_last_error(h_));
      return false;
    }
    planar_scan_index.insert(planar_scan_index.end(), buf_[0].begin(), buf_[0].begin() + cnt[0]);
    planar_submap_index.insert(planar_submap_index.end(), buf_[1].begin(), buf_[1].begin() + cnt[1]);
    sphere_scan_index.insert(sphere_scan_index.end(), buf_[2].begin(), buf_[2].begin() + cnt[2]);
    sphere_submap_index.insert(sphere_submap_index.end(), buf_[3].begin(), buf_[3].begin() + cnt[3]);
    return true;
  }

  int lastStatus() const { return last_status_; }

 private:
  void attach(tloam_b200_handle* shared, int device) {
    if (shared) { h_ = shared; own_ = false; return; }
    tloam_tls_config c;
    tloam_b200_default_config(&c);
    const int rc = tloam_b200_create(&c, device, nullptr, &h_);
    if (rc != TLOAM_B200_OK) throw std::runtime_error(std::string("tloam_b200_create: ") + tloam_b200_status_string(rc));
    own_ = true;
  }
  tloam_feature_config cfg_;
  tloam_b200_handle* h_ = nullptr;
  bool own_ = false;
  int last_status_ = TLOAM_B200_OK;
  std::vector<size_t> buf_[4];
};

}  // namespace tloam
#endif
