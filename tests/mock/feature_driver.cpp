// Drives tloam::featureExtractB200 the way FrontEnd::processCloud does (ref: src/front_end/front_end.cpp:187-194):
// reads a cloud written by the Python test (binary: count, points), calls extractPlanarSphere with four vectors that
// already hold one sentinel each (the reference appends) and prints the four lists.
#define TLOAM_B200_MOCK_HOST_TYPES
#include "mock_tloam.hpp"
#include "../../include/tloam_b200/feature_extract_b200.hpp"

#include <cstdint>
#include <cstdio>
#include <memory>

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: feature_driver cloud.bin\n"); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  uint64_t n = 0;
  if (fread(&n, sizeof(n), 1, f) != 1) return 2;
  tloam::CloudData cloud;
  cloud.cloud_ptr->points_.resize(n);
  if (n && fread(cloud.cloud_ptr->points_.data(), sizeof(Eigen::Vector3d), n, f) != n) return 2;
  std::fclose(f);
  tloam_feature_config cfg;
  tloam_b200_feature_default_config(&cfg);
  std::unique_ptr<tloam::featureExtractB200> fe;
  try {
    fe.reset(new tloam::featureExtractB200(cfg));
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 3;
  }
  std::vector<size_t> lists[4];
  for (auto& l : lists) l.push_back(987654321);
  if (!fe->extractPlanarSphere(cloud, lists[0], lists[1], lists[2], lists[3])) return 4;
  for (const auto& l : lists) {
    std::printf("%zu\n", l.size());
    for (size_t v : l) std::printf("%zu\n", v);
  }
  return 0;
}
