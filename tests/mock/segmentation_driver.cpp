// Drives tloam::SegmentationB200 the way Segmentation::spinOnce does (ref: src/models/segmentation/segmentation.cpp:47-66):
// groundRemove -> objectSegmentation -> extractEdgePoint, twice (two frames: the second one runs with the members
// resetParams() leaves).  Reads a scan written by the Python test (binary: count, points) and prints, per frame, the sizes of
// the clouds and boxes, then for every edge / general point the bit pattern of its first coordinate and its intensity.
#define TLOAM_B200_MOCK_HOST_TYPES
#include "mock_tloam.hpp"
#include "../../include/tloam_b200/segmentation_b200.hpp"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: segmentation_driver scan.bin\n"); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  uint64_t n = 0;
  if (fread(&n, sizeof(n), 1, f) != 1) return 2;
  std::vector<Eigen::Vector3d> pts(n);
  if (n && fread(pts.data(), sizeof(Eigen::Vector3d), n, f) != n) return 2;
  std::fclose(f);
  tloam_ground_config gcfg;
  tloam_b200_ground_default_config(&gcfg);
  tloam_dcvc_config dcfg;
  tloam_b200_dcvc_default_config(&dcfg);
  dcfg.min_polar_init = dcfg.max_polar_init = 5.0;             // first frame (segmentation.hpp:332-333)
  std::unique_ptr<tloam::SegmentationB200> seg;
  try {
    seg.reset(new tloam::SegmentationB200(gcfg, dcfg, 131));
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 3;
  }
  for (int frame = 0; frame < 2; ++frame) {
    tloam::CloudData scan, ground, object, segmented, edge, general;
    scan.cloud_ptr->points_ = pts;
    std::vector<tloam::BoxB200> boxes;
    if (!seg->groundRemove(scan, ground, object)) return 4;
    if (!seg->objectSegmentation(object, segmented, &boxes)) return 5;
    if (!seg->extractEdgePoint(segmented, edge, general)) return 6;
    std::printf("%zu %zu %zu %zu %zu %zu\n", ground.cloud_ptr->points_.size(), object.cloud_ptr->points_.size(),
                segmented.cloud_ptr->points_.size(), boxes.size(), edge.cloud_ptr->points_.size(), general.cloud_ptr->points_.size());
    for (const tloam::BoxB200& b : boxes)
      std::printf("%d %d %.17g %.17g %.17g %.17g %.17g %.17g\n", b.label, b.points, b.position[0], b.position[1], b.position[2],
                  b.dimensions[0], b.dimensions[1], b.dimensions[2]);
    auto dump = [](const tloam::CloudData& c) {
      for (size_t i = 0; i < c.cloud_ptr->points_.size(); ++i) {
        uint64_t bits;
        std::memcpy(&bits, &c.cloud_ptr->points_[i].v[0], 8);
        std::printf("%llu %.17g\n", (unsigned long long)bits, c.cloud_ptr->intensity_[i]);
      }
    };
    dump(edge);
    dump(general);
  }
  // the single-call form must hand out the same clouds as the last frame above (members as resetParams() leaves them)
  {
    tloam::CloudData scan, ground, edge, general;
    scan.cloud_ptr->points_ = pts;
    std::vector<tloam::BoxB200> boxes;
    if (!seg->segmentScan(scan, ground, edge, general, &boxes)) return 7;
    std::printf("%zu %zu %zu %zu\n", ground.cloud_ptr->points_.size(), boxes.size(), edge.cloud_ptr->points_.size(), general.cloud_ptr->points_.size());
    auto dump = [](const tloam::CloudData& c) {
      for (size_t i = 0; i < c.cloud_ptr->points_.size(); ++i) {
        uint64_t bits;
        std::memcpy(&bits, &c.cloud_ptr->points_[i].v[0], 8);
        std::printf("%llu %.17g\n", (unsigned long long)bits, c.cloud_ptr->intensity_[i]);
      }
    };
    dump(edge);
    dump(general);
  }
  return 0;
}
