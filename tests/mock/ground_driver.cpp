// Drives tloam::GroundExtractB200 the way Segmentation::spinOnce does (ref: src/models/segmentation/segmentation.cpp:47-52):
// reads a scan written by the Python test (binary: count, points), calls groundRemove with ground / object clouds that
// already hold one sentinel point each (the reference appends with +=) and prints sizes, the height threshold and the
// ground / object points' FIRST coordinates bit patterns (enough to identify each point).
#define TLOAM_B200_MOCK_HOST_TYPES
#include "mock_tloam.hpp"
#include "../../include/tloam_b200/ground_extract_b200.hpp"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: ground_driver scan.bin\n"); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  uint64_t n = 0;
  if (fread(&n, sizeof(n), 1, f) != 1) return 2;
  tloam::CloudData scan, ground, object;
  scan.cloud_ptr->points_.resize(n);
  if (n && fread(scan.cloud_ptr->points_.data(), sizeof(Eigen::Vector3d), n, f) != n) return 2;
  std::fclose(f);
  tloam_ground_config cfg;
  tloam_b200_ground_default_config(&cfg);
  std::unique_ptr<tloam::GroundExtractB200> ge;
  try {
    ge.reset(new tloam::GroundExtractB200(cfg));
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 3;
  }
  Eigen::Vector3d s;
  s[0] = s[1] = s[2] = 12345.0;
  ground.cloud_ptr->points_.push_back(s); ground.cloud_ptr->intensity_.push_back(-1.0);
  object.cloud_ptr->points_.push_back(s); object.cloud_ptr->intensity_.push_back(-1.0);
  if (!ge->groundRemove(scan, ground, object)) return 4;
  std::printf("%zu %zu %zu %.17g\n", ground.cloud_ptr->points_.size(), object.cloud_ptr->points_.size(), scan.cloud_ptr->points_.size(),
              ge->heightThreshold());
  auto dump = [](const tloam::CloudData& c) {
    for (size_t i = 0; i < c.cloud_ptr->points_.size(); ++i) {
      uint64_t bits;
      std::memcpy(&bits, &c.cloud_ptr->points_[i].v[0], 8);
      std::printf("%llu %.17g\n", (unsigned long long)bits, c.cloud_ptr->intensity_[i]);
    }
  };
  dump(ground);
  dump(object);
  return 0;
}
