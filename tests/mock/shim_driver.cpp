// Drives tloam::LocalRegistrationB200 through the RegistrationInterface base pointer exactly the way
// FrontEnd::updateLidarOdometry does (ref: src/front_end/front_end.cpp:314-321): reads a frame written by the
// Python test (binary: 8 counts, 8 clouds, predict), runs setInputTarget / setInputSource / scanMatching and
// prints the resulting pose (16 doubles, column-major).
#define TLOAM_B200_MOCK_HOST_TYPES
#include "mock_tloam.hpp"
#include "../../include/tloam_b200/local_registration_b200.hpp"

#include <cstdint>
#include <cstdio>
#include <cstdlib>

static bool read_cloud(FILE* f, uint64_t n, std::shared_ptr<open3d::geometry::PointCloud2>& c) {
  c->points_.resize(n);
  return n == 0 || fread(c->points_.data(), sizeof(Eigen::Vector3d), n, f) == n;
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: shim_driver frame.bin\n"); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  uint64_t n[8];
  if (fread(n, sizeof(uint64_t), 8, f) != 8) return 2;
  tloam::Frame map, scan, out;
  bool ok = read_cloud(f, n[0], map.edge_feature) && read_cloud(f, n[1], map.sphere_feature) &&
            read_cloud(f, n[2], map.planar_feature) && read_cloud(f, n[3], map.ground_feature) &&
            read_cloud(f, n[4], scan.edge_feature) && read_cloud(f, n[5], scan.sphere_feature) &&
            read_cloud(f, n[6], scan.planar_feature) && read_cloud(f, n[7], scan.ground_feature);
  Eigen::Isometry3d predict, result;
  ok = ok && fread(predict.matrix().data(), sizeof(double), 16, f) == 16;
  std::fclose(f);
  if (!ok) return 2;
  tloam_tls_config cfg;
  tloam_b200_default_config(&cfg);
  std::shared_ptr<tloam::RegistrationInterface> reg;
  try {
    reg = std::make_shared<tloam::LocalRegistrationB200>(cfg);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 3;
  }
  if (!reg->setInputTarget(map) || !reg->setInputSource(scan)) return 4;
  if (!reg->scanMatching(out, predict, result)) return 5;
  for (int i = 0; i < 16; ++i) std::printf("%.17g\n", result.matrix().data()[i]);
  const std::pair<double, double> fit = reg->getFitnessScore();
  std::printf("%.17g\n%.17g\n", fit.first, fit.second);
  return 0;
}
