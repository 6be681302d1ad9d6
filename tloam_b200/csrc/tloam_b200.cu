// tloam_b200.cu -- kernels + C ABI of libtloam_b200.so (sm_100a). See include/tloam_b200.h for the boundary
// and registration.cuh for the execution model.  No CPU fallback exists anywhere in this file.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <new>
#include <utility>
#include <vector>

#include "registration.cuh"
#include "solver.cuh"
#include "map_build.cuh"
#include "frame_kernels.cuh"
#include "dense_search.cuh"
#include "fine_search.cuh"
#include "submap.cuh"
#include "feature_extract.cuh"
#include "ground_extract.cuh"
#include "edge_extract.cuh"
#include "object_segment.cuh"
#include "host_stage.h"



// =================================================================================================
// Host side: handle + C ABI
// =================================================================================================
using namespace tloam;

#define CU_TRY(expr)                                                                                   \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess) {                                                                          \
      snprintf(h->last_error, sizeof(h->last_error), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
               __FILE__, __LINE__);                                                                    \
      return TLOAM_B200_ERR_CUDA;                                                                      \
    }                                                                                                  \
  } while (0)

struct tloam_b200_handle {
  tloam_tls_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t copy_stream = nullptr;              // H2D of the map clouds, overlapped with the build (set_target)
  cudaEvent_t ev_copy[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  int last_uploaded = 0;
  cudaEvent_t ev_src = nullptr;                    // set_source: the H2D copies have landed
  cudaStream_t fit_stream = nullptr;               // per-frame getFitnessScore runs beside the registration (fork / join in the frame graph)
  cudaEvent_t ev_fit[2] = {nullptr, nullptr};
  char last_error[512] = {0};
  HostStage hstage;                                // pageable host inputs: chunked, multi-threaded staging through pinned slots
  // tloam_b200_segment_scan: the three segmentation stages chained on the device.  A stage that finds `active` reads its
  // input from dev_xyz / dev_intensity (no upload) and, instead of copying its index lists home, leaves the device
  // pointers and counts here for the gather kernel that feeds the next stage.
  struct SegChain {
    bool active = false;
    const double* dev_xyz = nullptr; const double* dev_intensity = nullptr;
    const unsigned* ground = nullptr; const unsigned* object = nullptr; const int* beam = nullptr; unsigned n_ground = 0, n_object = 0;
    const unsigned long long* seg = nullptr; unsigned n_seg = 0;
    const unsigned long long* edge = nullptr; const unsigned long long* non_edge = nullptr; unsigned n_edge = 0, n_non = 0;
  } seg;
  unsigned char* d_chain = nullptr; size_t cap_chain = 0;
  long long launches = 0;
  int launches_frame = 0;
  // source
  size_t n_src[4] = {0, 0, 0, 0};
  bool have_src = false, have_tgt = false, frame_pending = false;
  double* d_stage_src = nullptr; size_t cap_stage_src = 0;      // points (the CURRENT staging buffer: one of d_stage_buf[])
  // pipelined handles (set_async_inputs) upload scan k+1 on a stream of its own while frame k is still being registered:
  // two staging buffers alternate; a buffer is free again once its last reader (k_stage_source, or the submap update
  // that appends the staged edge / ground features) has run
  double* d_stage_buf[2] = {nullptr, nullptr}; size_t cap_stage_buf[2] = {0, 0};
  int stage_cur = 0;
  cudaStream_t src_stream = nullptr;
  cudaEvent_t ev_stage_free[2] = {nullptr, nullptr};
  bool stage_free_valid[2] = {false, false};
  double* d_feat = nullptr; size_t cap_pad = 0;                 // 3 + 2 + 6 arrays of cap_pad doubles
  unsigned char* d_flags = nullptr;                             // flags + active
  int* d_blk_count = nullptr; double* d_partial = nullptr; size_t cap_blocks = 0;
  unsigned* d_counter = nullptr;
  FrameState* d_state = nullptr; bool own_state = true;         // a batch owns the states of its handles (one array)
  tloam_b200_stats* d_stats = nullptr;
  // target
  size_t n_tgt[4] = {0, 0, 0, 0};
  double* d_stage_tgt = nullptr; size_t cap_stage_tgt = 0;
  unsigned* d_scratch = nullptr; size_t cap_scratch = 0;        // slot_of + rank_of
  unsigned char* d_blob = nullptr; size_t cap_blob = 0, blob_bytes = 0;
  // second blob: a map being RECEIVED (shared-map broadcast) while frames still register against the active one
  unsigned char* d_blob_in = nullptr; size_t cap_blob_in = 0, blob_in_bytes = 0; MapHeader hdr_in; bool blob_in_ready = false;
  MapHeader hdr;                                                // host copy of the layout (origin filled lazily)
  bool origin_known = false;
  // pinned host staging for results
  double* h_result = nullptr;                                   // 16 result + 2 (status, done)
  tloam_b200_stats* h_stats = nullptr;
  DeviceCtx ctx;
  int total_blocks = 0;
  // whole-frame CUDA graph (re-captured only when the device context changes)
  Predict* h_predict = nullptr; Predict* d_predict = nullptr;
  cudaGraphExec_t gexec = nullptr; DeviceCtx gctx; bool gvalid = false; int glaunches = 0; bool use_graph = true;
  bool use_fused = false;   // k_first (search + fit + first evaluation in one kernel): opt-in, TLOAM_B200_FUSE=1
  bool gfused = false;      // topology of the instantiated graph
  // optional per-kernel-class timing (CUDA events around every launch; off by default)
  bool profiling = false;
  bool traced_last = false;
  bool trace = false;                                           // record tloam_b200_stats traces
  std::vector<cudaEvent_t> ev_pool;
  struct Span { int cls; cudaEvent_t a, b; };
  std::vector<Span> spans;
  size_t ev_next = 0;
  tloam_b200_profile prof;
  unsigned long long* d_dbg = nullptr;
  // ---- device-resident submap ((f)-1) ----
  tloam_submap_config scfg;
  bool submap_ready = false;
  double* d_acc[2] = {nullptr, nullptr};   size_t cap_acc[2] = {0, 0}, n_acc[2] = {0, 0};     // edge, ground accumulators
  double* d_acc_tmp = nullptr;             size_t cap_acc_tmp = 0;
  // n_acc is the host's UPPER BOUND of each accumulator; the exact counts live on the device (no host round trip per
  // frame): d_cnt[0..3] unused, [4 + 2k + acc_cur[k]] = points in accumulator k, [8] scratch
  unsigned* d_cnt = nullptr;               int acc_cur[2] = {0, 0};
  unsigned long long cum_add[2] = {0, 0}, known_cum[2] = {0, 0};  size_t known_cnt[2] = {0, 0};
  struct CntProbe { cudaEvent_t ev = nullptr; unsigned* h_vals = nullptr; unsigned long long cum[2] = {0, 0}; bool pending = false; };
  CntProbe probes[4];                      int probe_next = 0;
  bool src_staged = false;                 // d_stage_src holds the current source (needed by submap_update)
  // pipelined results (async_inputs): two pinned result slots + events, so that the host can stay one frame ahead
  cudaEvent_t ev_res[2] = {nullptr, nullptr};
  long long frames_enqueued = 0, frames_fetched = 0;
  bool frame_fitness = false, gfitness = false;   // k_fitness + reduce appended to every frame (graph topology flag)
  double last_fitness = 0.0, last_rmse = 0.0;
  bool async_inputs = false;               // tloam_b200_set_async_inputs: host buffers stay valid until the next sync
  // per-frame health metric (getFitnessScore) without allocation: block partials + the reduced pair in the state
  double* d_fit = nullptr;                 size_t cap_fit = 0;
  std::vector<double*> ring;               std::vector<size_t> ring_n, ring_cap;               // planar sliding window
  double* d_cat = nullptr;                 size_t cap_cat = 0, n_cat = 0;                       // concatenated planar window
  double* d_sphere0 = nullptr;             size_t n_sphere0 = 0; bool sphere_is_init = false;  // frame-0 sphere submap
  double* d_up = nullptr;                  size_t cap_up = 0;                                   // upload staging
  unsigned char* d_vox = nullptr;          size_t cap_vox = 0;                                  // voxel hash scratch
  // submap update: the ground accumulator is cropped / down-sampled on a stream of its own beside the edge accumulator
  // (own scratch and output buffer); the newest planar frame is uploaded on src_stream while the frame is registered
  unsigned char* d_vox1 = nullptr;         size_t cap_vox1 = 0;
  double* d_acc_tmp1 = nullptr;            size_t cap_acc_tmp1 = 0;
  double* d_up_planar = nullptr;           size_t cap_up_planar = 0;
  cudaStream_t sub_stream = nullptr;
  cudaEvent_t ev_sub[2] = {nullptr, nullptr};
  cudaEvent_t ev_planar_in = nullptr, ev_planar_free = nullptr, ev_planar_done = nullptr;
  bool planar_free_valid = false;
  // ---- PCA feature extraction ((f)-2): one arena, carved up per call ----
  unsigned char* d_fe = nullptr;           size_t cap_fe = 0;  bool fe_attr_set = false;
  double* d_pose = nullptr;
  unsigned char* d_ge = nullptr;           size_t cap_ge = 0;   // ground extraction ((f)-4) arena
  // ---- dense-map correspondence path (dense_search.cuh): query binning scratch + the search-path decision ----
  unsigned char* d_dense = nullptr;        size_t cap_dense = 0, dense_zero_bytes = 0;
  DenseArgs dargs, gdargs;                 // current / captured in the graph
  unsigned* h_mapstats = nullptr;          // pinned: occupied bricks per cloud of the newest map
  cudaEvent_t ev_stats = nullptr;          bool stats_pending = false, stats_known = false;
  unsigned nbricks[4] = {0, 0, 0, 0};
  int dense_mode = 0;                      // TLOAM_B200_DENSE: "auto" -> -1 (points per brick), unset -> 0 never, "1" always
  // two-level grid (fine_search.cuh): TLOAM_B200_FINE unset / "auto" -> -1 (by the density of the previous map),
  // "0" never, "1" always.  map_fine_mask: clouds of the ACTIVE map that were built with the second level
  int dense_kernel = 0, gdense_kernel = 0; // which kernel serves ctx.dense_mask: 1 = TMA-staged (dense_search.cuh), 2 = two-level grid
  int fine_mode = -1;
  int map_fine_mask = 0;
  float4* d_fine_tmp = nullptr;            size_t cap_fine_tmp = 0;
  bool dense_attr_set = false;
  bool fe_sort_attr_set = false, ge_attr_set = false, ee_attr_set = false, os_attr_set = false;   // per handle: function attributes are per device
  bool dense_check = false;                // TLOAM_B200_DENSE_CHECK=1: every dense query is re-searched by the plain path and compared
  int num_sms = 148;
};

// launch bookkeeping: counts the kernel and, in profiling mode, brackets it with events
struct LaunchScope {
  tloam_b200_handle* h; int cls; cudaEvent_t a = nullptr, b = nullptr;
  LaunchScope(tloam_b200_handle* hh, int c) : h(hh), cls(c) {
    h->launches++;
    if (h->profiling) {
      while (h->ev_pool.size() < h->ev_next + 2) { cudaEvent_t e; cudaEventCreate(&e); h->ev_pool.push_back(e); }
      a = h->ev_pool[h->ev_next++]; b = h->ev_pool[h->ev_next++];
      cudaEventRecord(a, h->stream);
    }
  }
  ~LaunchScope() {
    if (h->profiling) { cudaEventRecord(b, h->stream); h->spans.push_back({cls, a, b}); }
  }
};
#define TL_LAUNCH(cls, ...) do { LaunchScope ls__(h, cls); __VA_ARGS__; } while (0)

static size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

// per-sequence grids (the same function of the cloud sizes in single and batched mode => identical reduction trees)
static int eval_grid_of(int nb) {      // k_eval: grid-stride over the feature blocks, a whole number of clusters
  return ((nb < kEvalGridCap ? nb : kEvalGridCap) + kEvalCluster - 1) / kEvalCluster * kEvalCluster;
}
static int first_grid_of(int nb) {     // k_first: one block per 64 features, rounded up to the cluster size
  return (2 * nb + kEvalCluster - 1) / kEvalCluster * kEvalCluster;
}

extern "C" {

void tloam_b200_default_config(tloam_tls_config* c) {   // ref: config/mapping/lidar_odometry.yaml:23-39
  c->k_corr = 10; c->factor_num = 4;
  c->edge_dist_thres = 1.0; c->sphere_dist_thres = 0.5; c->planar_dist_thres = 0.5; c->ground_dist_thres = 0.5;
  c->edge_dir_thres = 0.85;
  c->edge_maxnum = 1200; c->sphere_maxnum = 200; c->planar_maxnum = 2500; c->ground_maxnum = 2000;
  c->max_iterations = 4; c->cost_threshold = 0.000000005; c->gnc_factor = 11.8; c->noise_bound = 0.01;
  c->fitness_thres = 0.02;
  c->ceres_max_num_iterations = 4;
  c->reinit_dir[0] = 1.0; c->reinit_dir[1] = 1.0; c->reinit_dir[2] = 1.0;
  c->initial_trust_region_radius = 1e4;
}

const char* tloam_b200_status_string(int s) {
  switch (s) {
    case TLOAM_B200_OK: return "ok";
    case TLOAM_B200_ERR_INVALID_ARG: return "invalid argument";
    case TLOAM_B200_ERR_TOO_FEW_POINTS: return "a cloud has fewer than 10 points";
    case TLOAM_B200_ERR_BAD_POSE: return "predicted pose is not a rigid transform";
    case TLOAM_B200_ERR_CUDA: return "CUDA error";
    case TLOAM_B200_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    case TLOAM_B200_ERR_NOT_READY: return "source or target not set";
    case TLOAM_B200_ERR_NUMERIC: return "non-finite value in the solve";
    case TLOAM_B200_ERR_MAP_DENSITY: return "a map cell holds more than 65535 points";
    default: return "unknown status";
  }
}

static double radius_of(const tloam_tls_config& c, int cloud) {
  return cloud == 0 ? c.edge_dist_thres : cloud == 1 ? c.sphere_dist_thres : cloud == 2 ? c.planar_dist_thres : c.ground_dist_thres;
}

int tloam_b200_create(const tloam_tls_config* cfg, int device, void* stream, tloam_b200_handle** out) {
  if (!cfg || !out) return TLOAM_B200_ERR_INVALID_ARG;
  *out = nullptr;
  if (cfg->factor_num < 2 || cfg->factor_num > 4 || cfg->max_iterations < 1 ||
      cfg->max_iterations > TLOAM_B200_MAX_OUTER || cfg->ceres_max_num_iterations < 0 ||
      cfg->ceres_max_num_iterations > TLOAM_B200_MAX_INNER || !(cfg->initial_trust_region_radius > 0.0))
    return TLOAM_B200_ERR_INVALID_ARG;
  for (int c = 0; c < 4; ++c) if (!(radius_of(*cfg, c) > 0.0)) return TLOAM_B200_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return TLOAM_B200_ERR_NO_DEVICE;
  }
  tloam_b200_handle* h = new (std::nothrow) tloam_b200_handle();
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  h->cfg = *cfg;
  h->device = device;
  auto fail = [&](int code) { tloam_b200_destroy(h); return code; };
  if (cudaSetDevice(device) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (stream) { h->stream = (cudaStream_t)stream; h->own_stream = false; }
  else {
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
    h->own_stream = true;
  }
  if (cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  for (int i = 0; i < 5; ++i)
    if (cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaEventCreateWithFlags(&h->ev_src, cudaEventDisableTiming) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&h->fit_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&h->src_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&h->sub_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  for (int i = 0; i < 2; ++i)
    if (cudaEventCreateWithFlags(&h->ev_sub[i], cudaEventDisableTiming) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaEventCreateWithFlags(&h->ev_planar_in, cudaEventDisableTiming) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaEventCreateWithFlags(&h->ev_planar_free, cudaEventDisableTiming) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaEventCreateWithFlags(&h->ev_planar_done, cudaEventDisableTiming) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  for (int i = 0; i < 2; ++i)
    if (cudaEventCreateWithFlags(&h->ev_stage_free[i], cudaEventDisableTiming) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  for (int i = 0; i < 2; ++i)
    if (cudaEventCreateWithFlags(&h->ev_fit[i], cudaEventDisableTiming) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  for (int i = 0; i < 2; ++i)
    if (cudaEventCreateWithFlags(&h->ev_res[i], cudaEventDisableTiming) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMalloc(&h->d_cnt, 32 * sizeof(unsigned)) != cudaSuccess || cudaMemset(h->d_cnt, 0, 32 * sizeof(unsigned)) != cudaSuccess)
    return fail(TLOAM_B200_ERR_CUDA);
  for (auto& pr : h->probes) {
    if (cudaEventCreateWithFlags(&pr.ev, cudaEventDisableTiming) != cudaSuccess || cudaMallocHost(&pr.h_vals, 2 * sizeof(unsigned)) != cudaSuccess)
      return fail(TLOAM_B200_ERR_CUDA);
  }
  if (cudaMalloc(&h->d_state, sizeof(FrameState)) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMalloc(&h->d_stats, sizeof(tloam_b200_stats)) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMalloc(&h->d_counter, 256) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMallocHost(&h->h_result, 64 * sizeof(double)) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);   // [0..31] slot 0 + scratch, [32..63] slot 1
  if (cudaMallocHost(&h->h_stats, sizeof(tloam_b200_stats)) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMallocHost(&h->h_predict, sizeof(Predict)) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMalloc(&h->d_predict, sizeof(Predict)) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  { const char* e = getenv("TLOAM_B200_NO_GRAPH"); h->use_graph = !(e && e[0] == '1'); }
  // the fused first evaluation (k_first) is opt-in: measured on B200 it is no faster than k_correspond + k_eval<first>
  // (single stream 0.435 vs 0.419 ms per frame, batch of 8: 1.83 vs 1.79 ms)
  { const char* e = getenv("TLOAM_B200_FUSE"); h->use_fused = (e && e[0] == '1'); }
  { const char* e = getenv("TLOAM_B200_NO_FUSE"); if (e && e[0] == '1') h->use_fused = false; }
  { const char* e = getenv("TLOAM_B200_DENSE_CHECK"); h->dense_check = (e && e[0] == '1'); }
  // dense-map search path (dense_search.cuh): off unless asked for.  "1" = always for the K = 5 clouds, "auto" = by the
  // points-per-brick statistics of the map.  Measured on config 3 it is still SLOWER than the lane-pair search (4.1 vs
  // 3.0 ms per launch): correct and TMA-staged, not yet a win.
  { const char* e = getenv("TLOAM_B200_DENSE"); h->dense_mode = 0; if (e && e[0] == '1') h->dense_mode = 1; else if (e && e[0] == 'a') h->dense_mode = -1; }
  { const char* e = getenv("TLOAM_B200_FINE"); h->fine_mode = -1; if (e && e[0] == '1') h->fine_mode = 1; else if (e && e[0] == '0') h->fine_mode = 0; }
  if (cudaMallocHost(&h->h_mapstats, 4 * sizeof(unsigned)) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaEventCreateWithFlags(&h->ev_stats, cudaEventDisableTiming) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) h->num_sms = v; }
  memset(&h->dargs, 0, sizeof(h->dargs)); memset(&h->gdargs, 0, sizeof(h->gdargs));
  if (kEvalCluster > 8) {                          // cluster sizes above 8 are "non-portable": opt in per kernel
    if (cudaFuncSetAttribute(k_eval<true, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(k_eval<false, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(k_eval<true, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(k_eval<false, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(k_first<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(k_first<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
      return fail(TLOAM_B200_ERR_CUDA);
  }
  // identity curr/last pose (the reference leaves them uninitialised until the first scanMatching)
  FrameState init;
  memset(&init, 0, sizeof(init));
  for (int i = 0; i < 16; ++i) init.curr_pose[i] = init.last_pose[i] = init.result[i] = (i % 5 == 0) ? 1.0 : 0.0;
  init.frame_done = 1;
  if (cudaMemcpy(h->d_state, &init, sizeof(init), cudaMemcpyHostToDevice) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMemset(h->d_counter, 0, 256) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  memset(&h->ctx, 0, sizeof(h->ctx));
  memset(&h->hdr, 0, sizeof(h->hdr));
  *out = h;
  return TLOAM_B200_OK;
}

int tloam_b200_destroy(tloam_b200_handle* h) {
  if (!h) return TLOAM_B200_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  h->hstage.shutdown();
  cudaFree(h->d_stage_buf[0]); cudaFree(h->d_stage_buf[1]); cudaFree(h->d_feat); cudaFree(h->d_flags); cudaFree(h->d_blk_count);
  cudaFree(h->d_partial); cudaFree(h->d_counter); if (h->own_state) cudaFree(h->d_state); cudaFree(h->d_stats);
  cudaFree(h->d_stage_tgt); cudaFree(h->d_scratch); cudaFree(h->d_blob); cudaFree(h->d_blob_in); cudaFree(h->d_dbg);
  cudaFree(h->d_acc[0]); cudaFree(h->d_acc[1]); cudaFree(h->d_acc_tmp); cudaFree(h->d_cat); cudaFree(h->d_sphere0);
  cudaFree(h->d_cnt); cudaFree(h->d_fit); cudaFree(h->d_ge);
  for (auto& pr : h->probes) { if (pr.ev) cudaEventDestroy(pr.ev); if (pr.h_vals) cudaFreeHost(pr.h_vals); }
  cudaFree(h->d_up); cudaFree(h->d_vox); cudaFree(h->d_pose); cudaFree(h->d_fe);
  for (double* p : h->ring) cudaFree(p);
  if (h->h_result) cudaFreeHost(h->h_result);
  if (h->h_stats) cudaFreeHost(h->h_stats);
  if (h->h_predict) cudaFreeHost(h->h_predict);
  if (h->h_mapstats) cudaFreeHost(h->h_mapstats);
  if (h->ev_stats) cudaEventDestroy(h->ev_stats);
  cudaFree(h->d_dense);
  cudaFree(h->d_fine_tmp);
  cudaFree(h->d_predict);
  if (h->gexec) cudaGraphExecDestroy(h->gexec);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (int i = 0; i < 5; ++i) if (h->ev_copy[i]) cudaEventDestroy(h->ev_copy[i]);
  if (h->ev_src) cudaEventDestroy(h->ev_src);
  if (h->fit_stream) { cudaStreamSynchronize(h->fit_stream); cudaStreamDestroy(h->fit_stream); }
  if (h->src_stream) { cudaStreamSynchronize(h->src_stream); cudaStreamDestroy(h->src_stream); }
  if (h->sub_stream) { cudaStreamSynchronize(h->sub_stream); cudaStreamDestroy(h->sub_stream); }
  for (int i = 0; i < 2; ++i) if (h->ev_sub[i]) cudaEventDestroy(h->ev_sub[i]);
  if (h->ev_planar_in) cudaEventDestroy(h->ev_planar_in);
  if (h->ev_planar_free) cudaEventDestroy(h->ev_planar_free);
  if (h->ev_planar_done) cudaEventDestroy(h->ev_planar_done);
  cudaFree(h->d_vox1); cudaFree(h->d_acc_tmp1); cudaFree(h->d_up_planar); cudaFree(h->d_chain);
  for (int i = 0; i < 2; ++i) if (h->ev_stage_free[i]) cudaEventDestroy(h->ev_stage_free[i]);
  for (int i = 0; i < 2; ++i) if (h->ev_fit[i]) cudaEventDestroy(h->ev_fit[i]);
  for (int i = 0; i < 2; ++i) if (h->ev_res[i]) cudaEventDestroy(h->ev_res[i]);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return TLOAM_B200_OK;
}

// fills the configuration part of the device context
static void fill_ctx_config(tloam_b200_handle* h) {
  DeviceCtx& c = h->ctx;
  const tloam_tls_config& f = h->cfg;
  for (int k = 0; k < 4; ++k) { const double r = radius_of(f, k); c.r2[k] = r * r; }
  c.maxnum[0] = f.edge_maxnum; c.maxnum[1] = f.sphere_maxnum; c.maxnum[2] = f.planar_maxnum; c.maxnum[3] = f.ground_maxnum;
  c.factor_num = f.factor_num; c.max_iterations = f.max_iterations; c.ceres_max_it = f.ceres_max_num_iterations;
  c.edge_dir_thres = f.edge_dir_thres; c.cost_threshold = f.cost_threshold; c.gnc_factor = f.gnc_factor;
  c.noise_bound = f.noise_bound; c.fitness_thres = f.fitness_thres;
  for (int k = 0; k < 3; ++k) c.reinit_dir[k] = f.reinit_dir[k];
  c.initial_radius = f.initial_trust_region_radius;
  c.st = h->d_state; c.stats = h->d_stats; c.counter = h->d_counter;
}

static int set_source_impl(tloam_b200_handle* h, const double* const xyz[4], const size_t n[4], bool on_device) {
  if (!h || !xyz || !n) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  size_t total = 0, pad = 0;
  int blocks = 0;
  for (int c = 0; c < 4; ++c) {
    if (n[c] > 0 && !xyz[c]) return TLOAM_B200_ERR_INVALID_ARG;
    if (n[c] > (size_t)1 << 30) return TLOAM_B200_ERR_INVALID_ARG;
    total += n[c];
    pad += round_up(n[c], kBlk);
  }
  if (pad == 0) pad = kBlk;
  blocks = (int)(pad / kBlk);
  // host input is staged; device input is read in place -- unless the device-side submap is in use, whose update
  // appends the staged edge / ground features after the frame (tloam_b200_submap_update)
  const bool stage = !on_device || h->submap_ready;
  // (re)allocations: a pointer is nulled and its capacity zeroed right after the free, and the new capacity is only
  // committed once the allocation succeeded, so a failed cudaMalloc leaves the handle consistent
  h->have_src = false; h->src_staged = false;
  static const bool no_prefetch = getenv("TLOAM_B200_NO_PREFETCH") != nullptr;   // A/B knob
  const bool prefetch = stage && !on_device && h->async_inputs && !no_prefetch;   // upload beside the frame that is still running
  if (prefetch) h->stage_cur ^= 1;
  const int sb = h->stage_cur;
  if (stage && total > h->cap_stage_buf[sb]) {
    cudaFree(h->d_stage_buf[sb]); h->d_stage_buf[sb] = nullptr; h->cap_stage_buf[sb] = 0;   // (cudaFree waits for its readers)
    h->d_stage_src = nullptr; h->cap_stage_src = 0;
    const size_t ncap = total + total / 4 + 1024;
    CU_TRY(cudaMalloc(&h->d_stage_buf[sb], ncap * 3 * sizeof(double)));
    h->cap_stage_buf[sb] = ncap;
  }
  h->d_stage_src = h->d_stage_buf[sb]; h->cap_stage_src = h->cap_stage_buf[sb];
  cudaStream_t up = prefetch ? h->src_stream : h->stream;
  if (prefetch && h->stage_free_valid[sb]) CU_TRY(cudaStreamWaitEvent(up, h->ev_stage_free[sb], 0));
  if (pad > h->cap_pad) {
    cudaFree(h->d_feat); cudaFree(h->d_flags); h->d_feat = nullptr; h->d_flags = nullptr; h->cap_pad = 0;
    const size_t ncap = round_up(pad + pad / 4, kBlk);
    CU_TRY(cudaMalloc(&h->d_feat, ncap * 11 * sizeof(double)));
    CU_TRY(cudaMalloc(&h->d_flags, ncap * 2));
    h->cap_pad = ncap;
  }
  if ((size_t)blocks > h->cap_blocks) {
    cudaFree(h->d_blk_count); cudaFree(h->d_partial); cudaFree(h->d_fit);
    h->d_blk_count = nullptr; h->d_partial = nullptr; h->d_fit = nullptr; h->cap_blocks = 0; h->cap_fit = 0;
    const size_t ncap = h->cap_pad / kBlk + 8;
    CU_TRY(cudaMalloc(&h->d_blk_count, 2 * ncap * sizeof(int)));
    CU_TRY(cudaMemsetAsync(h->d_blk_count, 0, 2 * ncap * sizeof(int), h->stream));
    CU_TRY(cudaMalloc(&h->d_partial, ncap * kNRed * sizeof(double)));
    CU_TRY(cudaMalloc(&h->d_fit, (ncap * 2 + 2) * sizeof(double)));   // + {fitness, rmse} of the side-stream reduction
    h->cap_blocks = ncap; h->cap_fit = ncap;
  }
  DeviceCtx& c = h->ctx;
  fill_ctx_config(h);
  size_t off = 0, poff = 0;
  const double* src[4];
  c.blk_off[0] = 0;
  for (int k = 0; k < 4; ++k) {
    src[k] = stage ? h->d_stage_src + 3 * off : xyz[k];
    if (n[k] > 0 && stage)
    {
      static const bool stage_pageable = getenv("TLOAM_B200_NO_HOST_STAGE") == nullptr;
      if (!on_device && stage_pageable && n[k] * 24 >= (128u << 10) && HostStage::pageable(xyz[k])) {
        CU_TRY(h->hstage.upload(h->d_stage_src + 3 * off, xyz[k], n[k] * 3 * sizeof(double), up));
      } else {
        CU_TRY(cudaMemcpyAsync(h->d_stage_src + 3 * off, xyz[k], n[k] * 3 * sizeof(double),
                               on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, up));
      }
    }
    c.n[k] = (int)n[k];
    c.pad_off[k] = (int)poff;
    off += n[k];
    poff += round_up(n[k], kBlk);
    c.blk_off[k + 1] = (int)(poff / kBlk);
    h->n_src[k] = n[k];
  }
  double* f = h->d_feat;
  const size_t cp = h->cap_pad;
  c.px = f; c.py = f + cp; c.pz = f + 2 * cp; c.w = f + 3 * cp; c.slot = f + 4 * cp;
  for (int j = 0; j < 6; ++j) c.prim[j] = f + (5 + j) * cp;
  c.flags = h->d_flags; c.active = h->d_flags + cp;
  c.blk_count = h->d_blk_count; c.blk_cap = (int)h->cap_blocks; c.partial = h->d_partial;
  h->total_blocks = c.blk_off[4];
  if (!on_device) CU_TRY(cudaEventRecord(h->ev_src, up));             // the uploads are in; the caller's buffers are free
  if (prefetch) CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_src, 0));
  if (h->total_blocks > 0) {
    TL_LAUNCH(TLOAM_B200_K_STAGE_SOURCE, (k_stage_source<<<h->total_blocks, kBlk, 0, h->stream>>>(src[0], src[1], src[2], src[3], c, f, f + cp, f + 2 * cp)));
    CU_TRY(cudaGetLastError());
  }
  if (stage) { CU_TRY(cudaEventRecord(h->ev_stage_free[sb], h->stream)); h->stage_free_valid[sb] = true; }
  h->have_src = true;
  h->src_staged = stage;
  if (!on_device && !h->async_inputs) CU_TRY(cudaEventSynchronize(h->ev_src));    // caller buffers may be freed on return
  return TLOAM_B200_OK;
}

int tloam_b200_set_source(tloam_b200_handle* h, const double* const xyz[4], const size_t n[4]) {
  return set_source_impl(h, xyz, n, false);
}
int tloam_b200_set_source_device(tloam_b200_handle* h, const double* const xyz[4], const size_t n[4]) {
  return set_source_impl(h, xyz, n, true);
}

static unsigned next_pow2(size_t v) { unsigned p = 256; while ((size_t)p < v) p <<= 1; return p; }

// the blob layout is a pure function of the configuration and the four point counts (so every rank of a shared-map
// broadcast can lay the incoming blob out without reading its header back)
static size_t layout_header(const tloam_tls_config& cfg, const size_t n[4], MapHeader& hd) {
  memset(&hd, 0, sizeof(hd));
  hd.magic = kMapMagic;
  size_t off = sizeof(MapHeader);
  for (int c = 0; c < 4; ++c) {
    hd.n[c] = (unsigned)n[c];
    hd.tsize[c] = next_pow2(n[c] + 1);      // >= bricks + 1 even if every point sits in its own brick
    hd.cell[c] = radius_of(cfg, c);
    hd.pts_off[c] = off; off += round_up(n[c] * sizeof(float4), 256);
  }
  // second-level tables: at most n / kFineMin dense cells per cloud (written by the build, never zeroed)
  for (int c = 0; c < 4; ++c) { hd.fine_off[c] = off; off += round_up((n[c] / kFineMin + 1) * (size_t)kFineEntryBytes, 256); }
  for (int c = 0; c < 4; ++c) { hd.table_off[c] = off; off += (size_t)hd.tsize[c] * kBrickBytes; }
  for (int d = 0; d < 3; ++d) { hd.bbox_enc[d] = ~0ull; hd.bbox_enc[3 + d] = 0ull; }
  return off;
}

static int layout_map(tloam_b200_handle* h, const size_t n[4]) {
  const size_t off = layout_header(h->cfg, n, h->hdr);
  h->blob_bytes = off;
  if (off > h->cap_blob) {
    cudaFree(h->d_blob); h->d_blob = nullptr; h->cap_blob = 0;
    CU_TRY(cudaMalloc(&h->d_blob, off + off / 4));
    h->cap_blob = off + off / 4;
  }
  return TLOAM_B200_OK;
}

static void bind_map(tloam_b200_handle* h) {
  DeviceCtx& c = h->ctx;
  for (int k = 0; k < 4; ++k) {
    c.grid[k].pts = reinterpret_cast<const float4*>(h->d_blob + h->hdr.pts_off[k]);
    c.grid[k].table = reinterpret_cast<const uint4*>(h->d_blob + h->hdr.table_off[k]);
    c.grid[k].mask = h->hdr.tsize[k] - 1u;
    c.grid[k].n = h->hdr.n[k];
    c.grid[k].cell = h->hdr.cell[k];
    c.grid[k].inv_cell = 1.0 / h->hdr.cell[k];
    c.grid[k].fine = reinterpret_cast<const unsigned short*>(h->d_blob + h->hdr.fine_off[k]);
  }
  c.origin = reinterpret_cast<const double*>(h->d_blob + offsetof(MapHeader, origin));
  c.map_flags = reinterpret_cast<const unsigned long long*>(h->d_blob + offsetof(MapHeader, build_flags));
  c.map_bricks = reinterpret_cast<const unsigned*>(h->d_blob + offsetof(MapHeader, nbricks));
}

// the origin lives in the device header; fetch it once per map (tiny D2H) so that it can be passed by value
// (and the build flags with it: a map whose cell counters overflowed must not be searched)
static int fetch_origin(tloam_b200_handle* h) {
  if (!h->origin_known) {
    CU_TRY(cudaMemcpyAsync(h->h_result + 24, h->d_blob + offsetof(MapHeader, origin), 3 * sizeof(double),
                           cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaMemcpyAsync(h->h_result + 27, h->d_blob + offsetof(MapHeader, build_flags), sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    for (int d = 0; d < 3; ++d) h->hdr.origin[d] = h->h_result[24 + d];
    memcpy(&h->hdr.build_flags, h->h_result + 27, sizeof(unsigned long long));
    h->origin_known = true;
  }
  return (h->hdr.build_flags & 1ull) ? TLOAM_B200_ERR_MAP_DENSITY : TLOAM_B200_OK;
}

static void harvest_map_stats(tloam_b200_handle* h, bool wait);

// Which clouds of the map about to be built get the second level (map_grid.cuh: kFineMin, fine_search.cuh).  The
// build kernel itself finds the dense cells; this only decides whether it is launched at all, so that sparse maps
// (BASELINE config 2: 8-25 points per occupied brick) do not pay for an empty launch.  Either choice leaves every
// search path exact.
constexpr double kFinePointsPerBrick = 256.0;    // config 2 maps: 8-25; config 3: ~1700
constexpr size_t kFineFirstMapPoints = 65536;     // no statistics yet (first map of a handle): big clouds only
static int fine_build_mask(tloam_b200_handle* h, const size_t n[4]) {
  if (h->fine_mode == 0) return 0;
  harvest_map_stats(h, false);
  int mask = 0;
  for (int c = 0; c < 4; ++c) {
    if (n[c] < kFineMin) continue;
    bool on = h->fine_mode == 1;
    if (!on) {
      if (h->stats_known) on = h->nbricks[c] > 0 && (double)h->n_tgt[c] / (double)h->nbricks[c] >= kFinePointsPerBrick;   // previous map
      else on = n[c] >= kFineFirstMapPoints;
    }
    if (on) mask |= 1 << c;
  }
  return mask;
}

static int set_target_impl(tloam_b200_handle* h, const double* const xyz[4], const size_t n[4], bool on_device,
                           const unsigned* const* n_dev = nullptr) {
  if (!h || !xyz || !n) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  size_t total = 0;
  for (int c = 0; c < 4; ++c) {
    if (n[c] > 0 && !xyz[c]) return TLOAM_B200_ERR_INVALID_ARG;
    if (n[c] > (size_t)1 << 30) return TLOAM_B200_ERR_INVALID_ARG;
    total += n[c];
  }
  h->have_tgt = false;
  bool all_staged = !on_device;                    // every host cloud went through the pageable staging ring
  if (total > h->cap_scratch) {
    cudaFree(h->d_scratch); h->d_scratch = nullptr; h->cap_scratch = 0;
    const size_t ncap = total + total / 4 + 1024;
    CU_TRY(cudaMalloc(&h->d_scratch, ncap * 2 * sizeof(unsigned)));
    h->cap_scratch = ncap;
  }
  if (!on_device && total > h->cap_stage_tgt) {
    cudaFree(h->d_stage_tgt); h->d_stage_tgt = nullptr; h->cap_stage_tgt = 0;
    const size_t ncap = total + total / 4 + 1024;
    CU_TRY(cudaMalloc(&h->d_stage_tgt, ncap * 3 * sizeof(double)));
    h->cap_stage_tgt = ncap;
  }
  const int fine_mask = fine_build_mask(h, n);
  if (fine_mask && total > h->cap_fine_tmp) {
    cudaFree(h->d_fine_tmp); h->d_fine_tmp = nullptr; h->cap_fine_tmp = 0;
    const size_t ncap = total + total / 4 + 1024;
    CU_TRY(cudaMalloc(&h->d_fine_tmp, ncap * sizeof(float4)));
    h->cap_fine_tmp = ncap;
  }
  int rc = layout_map(h, n);
  if (rc != TLOAM_B200_OK) return rc;
  h->hdr.fine_build = (unsigned)fine_mask;
  MapBuildArgs a;
  a.fine_mask = fine_mask; a.fine_tmp = h->d_fine_tmp;
  a.blob = h->d_blob; a.slot_of = h->d_scratch; a.rank_of = h->d_scratch + h->cap_scratch;
  a.only_cloud = -1;
  for (int c = 0; c < 4; ++c) a.n_dev[c] = n_dev ? n_dev[c] : nullptr;
  for (int c = 0; c < 4; ++c) h->ctx.tgt_cnt[c] = a.n_dev[c];
  size_t off = 0;
  int first = -1;                                  // first non-empty cloud: its bounding box defines the origin
  for (int c = 0; c < 4; ++c) {
    a.stage_off[c] = (unsigned)off;
    // host input is staged; device input is read in place by the build kernels (stream-ordered, no extra copy)
    a.src[c] = on_device ? xyz[c] : h->d_stage_tgt + 3 * off;
    if (first < 0 && n[c] > 0) first = c;
    off += n[c];
    h->n_tgt[c] = n[c];
  }
  a.stage_off[4] = (unsigned)off;
  // header + zeroed tables (key 0 == empty)
  CU_TRY(cudaMemcpyAsync(h->d_blob, &h->hdr, sizeof(MapHeader), cudaMemcpyHostToDevice, h->stream));
  const size_t tables_bytes = h->blob_bytes - h->hdr.table_off[0];
  CU_TRY(cudaMemsetAsync(h->d_blob + h->hdr.table_off[0], 0, tables_bytes, h->stream));
  const unsigned tb = 256;
  auto blocks_for = [&](size_t cnt) { return (unsigned)((cnt + tb - 1) / tb); };
  if (total == 0) {
    TL_LAUNCH(TLOAM_B200_K_MAP_ORIGIN, (k_map_origin<<<1, 32, 0, h->stream>>>(a)));
  } else if (on_device) {
    // inputs already in HBM: one pass over all clouds per kernel
    MapBuildArgs ab = a;
    ab.i_beg = a.stage_off[first]; ab.i_end = a.stage_off[first + 1];
    const unsigned gbb = blocks_for(n[first]);
    TL_LAUNCH(TLOAM_B200_K_MAP_BBOX, (k_map_bbox<<<(gbb < 592u ? gbb : 592u), tb, 0, h->stream>>>(ab)));   // + origin
    a.i_beg = 0; a.i_end = (unsigned)total;
    const unsigned gb = blocks_for(total);
    unsigned tslots = 0;
    for (int c = 0; c < 4; ++c) tslots += h->hdr.tsize[c];
    TL_LAUNCH(TLOAM_B200_K_MAP_INSERT, (k_map_insert<<<gb, tb, 0, h->stream>>>(a)));
    TL_LAUNCH(TLOAM_B200_K_MAP_OFFSETS, (k_map_offsets<<<(tslots + tb - 1) / tb, tb, 0, h->stream>>>(a)));
    TL_LAUNCH(TLOAM_B200_K_MAP_SCATTER, (k_map_scatter<<<gb, tb, 0, h->stream>>>(a)));
    if (fine_mask) TL_LAUNCH(TLOAM_B200_K_MAP_FINE, (k_map_fine<<<4 * h->num_sms, 128, 0, h->stream>>>(a)));
    CU_TRY(cudaGetLastError());
  } else {
    // host inputs: the clouds cross PCIe one after the other on a copy stream; cloud c is inserted, offset and
    // scattered on the compute stream while cloud c+1 is still in flight.  The origin cloud goes first, then the
    // others by descending size, so that the build left exposed after the last copy is the smallest one.
    int order[4], m = 0;
    order[m++] = first;
    for (int c = 0; c < 4; ++c) if (c != first && n[c] > 0) order[m++] = c;
    for (int i = 2; i < m; ++i)
      for (int j = i; j > 1 && n[order[j]] > n[order[j - 1]]; --j) { const int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
    CU_TRY(cudaEventRecord(h->ev_copy[0], h->stream));                  // the staging buffer is free once earlier work is done
    CU_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_copy[0], 0));
    static const bool stage_pageable = getenv("TLOAM_B200_NO_HOST_STAGE") == nullptr;   // A/B knob
    for (int k = 0; k < m; ++k) {
      const int c = order[k];
      if (stage_pageable && HostStage::pageable(xyz[c])) {
        CU_TRY(h->hstage.upload(h->d_stage_tgt + 3 * (size_t)a.stage_off[c], xyz[c], n[c] * 3 * sizeof(double), h->copy_stream));
      } else {
        CU_TRY(cudaMemcpyAsync(h->d_stage_tgt + 3 * (size_t)a.stage_off[c], xyz[c], n[c] * 3 * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
        all_staged = false;
      }
      CU_TRY(cudaEventRecord(h->ev_copy[1 + c], h->copy_stream));
      h->last_uploaded = c;
    }
    for (int k = 0; k < m; ++k) {
      const int c = order[k];
      CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_copy[1 + c], 0));
      MapBuildArgs ac = a;
      ac.i_beg = a.stage_off[c]; ac.i_end = a.stage_off[c + 1]; ac.only_cloud = c;
      const unsigned gb = blocks_for(n[c]);
      if (k == 0) TL_LAUNCH(TLOAM_B200_K_MAP_BBOX, (k_map_bbox<<<(gb < 592u ? gb : 592u), tb, 0, h->stream>>>(ac)));   // + origin
      TL_LAUNCH(TLOAM_B200_K_MAP_INSERT, (k_map_insert<<<gb, tb, 0, h->stream>>>(ac)));
      TL_LAUNCH(TLOAM_B200_K_MAP_OFFSETS, (k_map_offsets<<<(h->hdr.tsize[c] + tb - 1) / tb, tb, 0, h->stream>>>(ac)));
      TL_LAUNCH(TLOAM_B200_K_MAP_SCATTER, (k_map_scatter<<<gb, tb, 0, h->stream>>>(ac)));
      if ((fine_mask >> c) & 1) {
        ac.fine_mask = 1 << c;
        TL_LAUNCH(TLOAM_B200_K_MAP_FINE, (k_map_fine<<<4 * h->num_sms, 128, 0, h->stream>>>(ac)));
      }
    }
    CU_TRY(cudaGetLastError());
  }
  bind_map(h);
  h->map_fine_mask = fine_mask;
  h->origin_known = false;
  h->have_tgt = true;
  // occupied bricks per cloud pick the search path (points per brick).  They come home with every frame's result
  // (FrameState::map_bricks); only the FIRST map of a handle is read back on its own, so that the very first frame can
  // already be routed -- both paths return the same exact neighbours, so the choice never changes a pose
  if (!h->stats_known) {
    CU_TRY(cudaMemcpyAsync(h->h_mapstats, h->d_blob + offsetof(MapHeader, nbricks), 4 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaEventRecord(h->ev_stats, h->stream));
    h->stats_pending = true;
  }
  // host path: the caller's buffers are free once the LAST upload has landed; the build of the last cloud may
  // still be running on the compute stream (everything that follows is ordered behind it on that stream)
  // (staged pageable inputs have been read completely already: nothing to wait for)
  if (!on_device && total > 0 && !all_staged) CU_TRY(cudaEventSynchronize(h->ev_copy[1 + h->last_uploaded]));
  return TLOAM_B200_OK;
}

int tloam_b200_set_target(tloam_b200_handle* h, const double* const xyz[4], const size_t n[4]) {
  return set_target_impl(h, xyz, n, false);
}
int tloam_b200_set_target_device(tloam_b200_handle* h, const double* const xyz[4], const size_t n[4]) {
  return set_target_impl(h, xyz, n, true);
}

int tloam_b200_get_map_origin(tloam_b200_handle* h, double origin[3]) {
  if (!h || !origin) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->have_tgt) return TLOAM_B200_ERR_NOT_READY;
  CU_TRY(cudaSetDevice(h->device));
  int rc = fetch_origin(h);
  if (rc != TLOAM_B200_OK) return rc;
  for (int d = 0; d < 3; ++d) origin[d] = h->hdr.origin[d];
  return TLOAM_B200_OK;
}

int tloam_b200_map_blob_size(tloam_b200_handle* h, size_t* bytes) {
  if (!h || !bytes) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->have_tgt) return TLOAM_B200_ERR_NOT_READY;
  *bytes = h->blob_bytes;
  return TLOAM_B200_OK;
}

int tloam_b200_map_export(tloam_b200_handle* h, void* d_dst, size_t bytes) {
  if (!h || !d_dst) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->have_tgt) return TLOAM_B200_ERR_NOT_READY;
  if (bytes < h->blob_bytes) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaMemcpyAsync(d_dst, h->d_blob, h->blob_bytes, cudaMemcpyDeviceToDevice, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return TLOAM_B200_OK;
}

int tloam_b200_map_import(tloam_b200_handle* h, const void* d_src, size_t bytes) {
  if (!h || !d_src || bytes < sizeof(MapHeader)) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  MapHeader hd;
  CU_TRY(cudaMemcpy(&hd, d_src, sizeof(hd), cudaMemcpyDeviceToHost));
  if (hd.magic != kMapMagic) return TLOAM_B200_ERR_INVALID_ARG;
  size_t need = hd.table_off[3] + (size_t)hd.tsize[3] * kBrickBytes;
  if (bytes < need) return TLOAM_B200_ERR_INVALID_ARG;
  for (int c = 0; c < 4; ++c)
    if (hd.cell[c] != radius_of(h->cfg, c)) return TLOAM_B200_ERR_INVALID_ARG;   // grid cell must equal this handle's radius
  if (need > h->cap_blob) {
    cudaFree(h->d_blob);
    h->cap_blob = need + need / 4;
    CU_TRY(cudaMalloc(&h->d_blob, h->cap_blob));
  }
  CU_TRY(cudaMemcpyAsync(h->d_blob, d_src, need, cudaMemcpyDeviceToDevice, h->stream));
  h->hdr = hd;
  h->blob_bytes = need;
  for (int c = 0; c < 4; ++c) h->n_tgt[c] = hd.n[c];
  fill_ctx_config(h);
  bind_map(h);
  h->origin_known = true;
  h->have_tgt = true;
  for (int c = 0; c < 4; ++c) h->nbricks[c] = hd.nbricks[c];
  h->map_fine_mask = (int)hd.fine_build;
  h->stats_known = true; h->stats_pending = false;
  CU_TRY(cudaStreamSynchronize(h->stream));
  return TLOAM_B200_OK;
}


// ---- zero-copy shared-map transport (config 4): ONE collective, no size handshake, no host synchronisation ----
// sender: the built blob itself (no export copy).  The consumer stream must be ordered behind the build:
// tloam_b200_signal_stream(h, consumer_stream).
int tloam_b200_map_send_buffer(tloam_b200_handle* h, void** d_ptr, size_t* bytes) {
  if (!h || !d_ptr || !bytes) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->have_tgt) return TLOAM_B200_ERR_NOT_READY;
  *d_ptr = h->d_blob; *bytes = h->blob_bytes;
  return TLOAM_B200_OK;
}

int tloam_b200_map_layout_bytes(tloam_b200_handle* h, const size_t n[4], size_t* bytes) {
  if (!h || !n || !bytes) return TLOAM_B200_ERR_INVALID_ARG;
  MapHeader hd;
  *bytes = layout_header(h->cfg, n, hd);
  return TLOAM_B200_OK;
}

// receiver: where the broadcast of a map with these point counts must land -- a SECOND blob of the handle, so frames
// keep registering against the active map while the next one is in flight.
int tloam_b200_map_recv_buffer(tloam_b200_handle* h, const size_t n[4], void** d_ptr, size_t* bytes) {
  if (!h || !n || !d_ptr || !bytes) return TLOAM_B200_ERR_INVALID_ARG;
  for (int c = 0; c < 4; ++c) if (n[c] > (size_t)1 << 30) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  const size_t need = layout_header(h->cfg, n, h->hdr_in);
  if (need > h->cap_blob_in) {
    CU_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_blob_in); h->d_blob_in = nullptr; h->cap_blob_in = 0;
    CU_TRY(cudaMalloc(&h->d_blob_in, need + need / 4));
    h->cap_blob_in = need + need / 4;
  }
  h->blob_in_bytes = need; h->blob_in_ready = true;
  *d_ptr = h->d_blob_in; *bytes = need;
  return TLOAM_B200_OK;
}

// receiver: the collective that fills the receive buffer has been ENQUEUED on producer_stream.  Orders the handle's
// stream behind it (device-side wait) and makes the received blob the active map: no copy, no host synchronisation.
int tloam_b200_map_adopt(tloam_b200_handle* h, void* producer_stream) {
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->blob_in_ready) return TLOAM_B200_ERR_NOT_READY;
  CU_TRY(cudaSetDevice(h->device));
  if ((cudaStream_t)producer_stream != h->stream) {
    CU_TRY(cudaEventRecord(h->ev_copy[0], (cudaStream_t)producer_stream));
    CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_copy[0], 0));
  }
  std::swap(h->d_blob, h->d_blob_in);
  std::swap(h->cap_blob, h->cap_blob_in);
  h->blob_bytes = h->blob_in_bytes;
  h->hdr = h->hdr_in;
  h->blob_in_ready = false;
  for (int c = 0; c < 4; ++c) { h->n_tgt[c] = h->hdr.n[c]; h->ctx.tgt_cnt[c] = nullptr; }
  fill_ctx_config(h);
  bind_map(h);
  h->map_fine_mask = 0xF;                  // unknown without reading the header: bricks without a second level say so themselves
  h->origin_known = false;                 // lives in the received header on the device; fetched lazily if ever asked for
  h->stats_pending = false;                // occupied-brick statistics stay those of the previous map
  h->have_tgt = true;
  return TLOAM_B200_OK;
}

// orders `consumer_stream` behind everything enqueued so far on the handle's stream (device-side wait)
int tloam_b200_signal_stream(tloam_b200_handle* h, void* consumer_stream) {
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  if ((cudaStream_t)consumer_stream == h->stream) return TLOAM_B200_OK;
  CU_TRY(cudaEventRecord(h->ev_copy[1], h->stream));
  CU_TRY(cudaStreamWaitEvent((cudaStream_t)consumer_stream, h->ev_copy[1], 0));
  return TLOAM_B200_OK;
}

static int check_ready(tloam_b200_handle* h) {
  if (!h->have_src || !h->have_tgt) return TLOAM_B200_ERR_NOT_READY;
  for (int c = 0; c < 4; ++c)
    if (h->n_src[c] < 10 || h->n_tgt[c] < 10) return TLOAM_B200_ERR_TOO_FEW_POINTS;   // ref: :928-929
  return TLOAM_B200_OK;
}


// ---- dense-map path: decision + scratch ----
static void harvest_map_stats(tloam_b200_handle* h, bool wait) {
  if (!h->stats_pending) return;
  if (wait) cudaEventSynchronize(h->ev_stats);
  else if (cudaEventQuery(h->ev_stats) != cudaSuccess) { cudaGetLastError(); return; }
  for (int c = 0; c < 4; ++c) h->nbricks[c] = h->h_mapstats[c];
  h->stats_pending = false; h->stats_known = true;
}

// exact accumulator counts (device) -> tighter host bounds, whenever an asynchronous read-back has landed
static void harvest_counts(tloam_b200_handle* h) {
  for (int i = 0; i < 4; ++i) {
    tloam_b200_handle::CntProbe& pr = h->probes[i];
    if (!pr.pending) continue;
    if (cudaEventQuery(pr.ev) != cudaSuccess) { cudaGetLastError(); continue; }
    pr.pending = false;
    for (int k = 0; k < 2; ++k)
      if (pr.cum[k] >= h->known_cum[k]) { h->known_cum[k] = pr.cum[k]; h->known_cnt[k] = pr.h_vals[k]; }
  }
  for (int k = 0; k < 2; ++k) {
    const size_t bound = h->known_cnt[k] + (size_t)(h->cum_add[k] - h->known_cum[k]);
    if (bound < h->n_acc[k]) h->n_acc[k] = bound;
  }
}

constexpr double kDensePointsPerBrick = 256.0;   // config 2 maps: 8-25; config 3: ~1700
constexpr size_t kDenseMinQueries = 2048;

// clouds served by the two-level search (k_correspond_fine): built with the second level and dense (or not yet known)
static int fine_mask_of(tloam_b200_handle* h) {
  if (h->fine_mode == 0 || h->map_fine_mask == 0) return 0;
  harvest_map_stats(h, false);
  int mask = 0;
  for (int c = 0; c < 4; ++c) {
    const bool enabled = (c == kPlanar || c == kGround) ? true : (c == kEdge ? h->cfg.factor_num >= 3 : h->cfg.factor_num == 4);
    if (!enabled || h->n_src[c] == 0 || h->n_tgt[c] == 0 || !((h->map_fine_mask >> c) & 1)) continue;
    const bool dense = !h->stats_known || (h->nbricks[c] > 0 && (double)h->n_tgt[c] / (double)h->nbricks[c] >= kFinePointsPerBrick);
    if (h->fine_mode == 1 || dense) mask |= 1 << c;
  }
  return mask;
}

static int dense_mask_of(tloam_b200_handle* h) {
  h->dense_kernel = 0;
  if (h->dense_mode == 0) {
    const int fm = fine_mask_of(h);
    if (fm) h->dense_kernel = 2;
    return fm;
  }
  h->dense_kernel = 1;
  harvest_map_stats(h, !h->stats_known);         // the very first map of a handle: wait once for its statistics
  int mask = 0;
  for (int c = 0; c < 4; ++c) {
    if (c == kSphere) continue;                   // K = 1 search: lane-pair path only
    const bool enabled = (c == kPlanar || c == kGround) ? true : h->cfg.factor_num >= 3;
    if (!enabled || h->n_src[c] == 0 || h->n_tgt[c] == 0) continue;
    const bool dense = h->nbricks[c] > 0 && (double)h->n_tgt[c] / (double)h->nbricks[c] >= kDensePointsPerBrick &&
                       h->n_src[c] >= kDenseMinQueries;
    if (h->dense_mode == 1 || dense) mask |= 1 << c;
  }
  return mask;
}

static int prepare_dense(tloam_b200_handle* h, int mask) {
  DenseArgs a;
  memset(&a, 0, sizeof(a));
  a.mask = mask;
  size_t off = 256;                              // ctl words first
  size_t total_q = 0;
  unsigned toff = 0;
  size_t o_keys[4] = {0, 0, 0, 0}, o_cnt[4] = {0, 0, 0, 0};
  unsigned ts[4] = {0, 0, 0, 0};
  for (int c = 0; c < 4; ++c) {
    if (!((mask >> c) & 1)) continue;
    ts[c] = next_pow2(2 * h->n_src[c] + 1);
    o_keys[c] = off; off += (size_t)ts[c] * 8;
    o_cnt[c] = off; off += (size_t)ts[c] * 4;
    total_q += h->n_src[c];
  }
  const size_t zero_bytes = off;
  size_t o_base[4], o_slot[4], o_rank[4], o_order[4];
  for (int c = 0; c < 4; ++c) {
    if (!((mask >> c) & 1)) continue;
    o_base[c] = off; off += (size_t)ts[c] * 4;
    o_slot[c] = off; off += round_up(h->n_src[c] * 4, 256);
    o_rank[c] = off; off += round_up(h->n_src[c] * 4, 256);
    o_order[c] = off; off += round_up(h->n_src[c] * 4, 256);
  }
  const size_t o_work = off; off += total_q * sizeof(DenseWork);
  if (off > h->cap_dense) {
    CU_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_dense); h->d_dense = nullptr; h->cap_dense = 0;
    CU_TRY(cudaMalloc(&h->d_dense, off + off / 4));
    h->cap_dense = off + off / 4;
  }
  unsigned char* b = h->d_dense;
  for (int c = 0; c < 4; ++c) {
    if (!((mask >> c) & 1)) continue;
    DenseCloud& q = a.cl[c];
    q.keys = (unsigned long long*)(b + o_keys[c]); q.cnt = (unsigned*)(b + o_cnt[c]); q.base = (unsigned*)(b + o_base[c]);
    q.q_slot = (unsigned*)(b + o_slot[c]); q.q_rank = (unsigned*)(b + o_rank[c]); q.q_order = (unsigned*)(b + o_order[c]);
    q.tmask = ts[c] - 1u; q.toff = toff; toff += ts[c];
  }
  a.work = (DenseWork*)(b + o_work);
  a.ctl = (unsigned*)b;
  a.tslots = toff;
  a.dbg = h->dense_check ? h->d_cnt + 16 : nullptr;          // d_cnt[16..23]: self-check counters
  h->dargs = a;
  h->dense_zero_bytes = zero_bytes;
  if (!h->dense_attr_set) {
    CU_TRY(cudaFuncSetAttribute(k_correspond_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDenseSmemBytes));
    h->dense_attr_set = true;
  }
  return TLOAM_B200_OK;
}

// true when no `*_maxnum` cap can bind for any enabled cloud: the fused first evaluation (k_first) applies
static bool caps_cannot_bind(const tloam_b200_handle* h) {
  const DeviceCtx& c = h->ctx;
  for (int k = 0; k < 4; ++k) {
    const bool enabled = (k == kPlanar || k == kGround) ? true : (k == kEdge ? c.factor_num >= 3 : c.factor_num == 4);
    if (enabled && c.maxnum[k] < c.n[k]) return false;
  }
  return true;
}

static constexpr size_t kFrameResultBytes = 16 * sizeof(double) + 2 * sizeof(int) + 2 * sizeof(double) + 4 * sizeof(unsigned);
static BatchTab no_batch() { BatchTab t; memset(&t, 0, sizeof(t)); t.S = 1; return t; }

// correspondence search + fit of one outer iteration as kernels of their own (un-fused sequence, build_factors)
static int enqueue_correspond(tloam_b200_handle* h, const DeviceCtx& c) {
  const int nb = h->total_blocks;
  if (c.dense_mask && h->dense_kernel == 2) {
    // dense map clouds: two-level grid, one thread per feature (fine_search.cuh)
    TL_LAUNCH(TLOAM_B200_K_FINE, (k_correspond_fine<false><<<nb, kBlk, kFineSmemBytes, h->stream>>>(c, no_batch(), h->dense_check ? h->d_cnt + 16 : nullptr)));
  } else if (c.dense_mask) {
    // dense map clouds: bin the queries by map cell, then the TMA-staged block-per-cell search (dense_search.cuh)
    const DenseArgs& da = h->dargs;
    CU_TRY(cudaMemsetAsync(h->d_dense, 0, h->dense_zero_bytes, h->stream));
    TL_LAUNCH(TLOAM_B200_K_DENSE_BIN, (k_qbin_count<<<nb, kBlk, 0, h->stream>>>(c, da)));
    TL_LAUNCH(TLOAM_B200_K_DENSE_BIN, (k_qbin_offsets<<<(da.tslots + 255) / 256, 256, 0, h->stream>>>(c, da)));
    TL_LAUNCH(TLOAM_B200_K_DENSE_BIN, (k_qbin_scatter<<<nb, kBlk, 0, h->stream>>>(c, da)));
    TL_LAUNCH(TLOAM_B200_K_DENSE, (k_correspond_dense<<<h->num_sms, kDenseThreads, kDenseSmemBytes, h->stream>>>(c, da)));
  }
  TL_LAUNCH(TLOAM_B200_K_CORRESPOND, (k_correspond<false><<<nb * 2, kBlk, 0, h->stream>>>(c, no_batch())));
  return TLOAM_B200_OK;
}

// enqueues the frame's fixed launch sequence on h->stream (also used under stream capture)
static int enqueue_frame(tloam_b200_handle* h, const DeviceCtx& c, bool fused) {
  const int nb = h->total_blocks;
  const int ne = eval_grid_of(nb), nf = first_grid_of(nb);
  const BatchTab nt = no_batch();
  // per-frame health metric (ref: registration.cpp:257-296): the reference queries the UNTRANSFORMED scan, so the two
  // kernels depend on the staged scan and the map only -- they run on a side stream beside the registration (a fork /
  // join inside the captured graph; the frame leaves 85 % of the SMs idle) and join before the result is copied.
  // With profiling on (events around every launch on h->stream) they stay in line.
  const bool fit = h->frame_fitness && nb > 0;
  static const bool fit_inline = getenv("TLOAM_B200_FIT_INLINE") != nullptr;      // A/B knob
  cudaStream_t fs = (h->profiling || fit_inline) ? h->stream : h->fit_stream;
  if (fit) {
    if (fs != h->stream) {
      CU_TRY(cudaEventRecord(h->ev_fit[0], h->stream));
      CU_TRY(cudaStreamWaitEvent(fs, h->ev_fit[0], 0));
    }
    TL_LAUNCH(TLOAM_B200_K_FITNESS, (k_fitness<<<nb, kBlk, 0, fs>>>(c, h->cfg.fitness_thres * h->cfg.fitness_thres, h->d_fit)));
    TL_LAUNCH(TLOAM_B200_K_FITNESS, (k_fitness_reduce<<<1, 128, 0, fs>>>(c, h->d_fit, h->d_fit + 2 * h->cap_fit)));
    if (fs != h->stream) CU_TRY(cudaEventRecord(h->ev_fit[1], fs));
  }
  CU_TRY(cudaMemcpyAsync(h->d_predict, h->h_predict, sizeof(Predict), cudaMemcpyHostToDevice, h->stream));
  TL_LAUNCH(TLOAM_B200_K_BEGIN_FRAME, (k_begin_frame<false><<<1, 256, 0, h->stream>>>(c, nt, h->d_predict)));
  for (int outer = 0; outer < h->cfg.max_iterations; ++outer) {
    if (fused) {
      TL_LAUNCH(TLOAM_B200_K_FIRST, (k_first<false><<<nf, kBlk, 0, h->stream>>>(c, nt)));
    } else {
      const int crc = enqueue_correspond(h, c);
      if (crc != TLOAM_B200_OK) return crc;
      TL_LAUNCH(TLOAM_B200_K_EVAL_FIRST, (k_eval<true, false><<<ne, kBlk, 0, h->stream>>>(c, nt)));
    }
    for (int it = 0; it < h->cfg.ceres_max_num_iterations; ++it)
      TL_LAUNCH(TLOAM_B200_K_EVAL, (k_eval<false, false><<<ne, kBlk, 0, h->stream>>>(c, nt)));
  }
  if (fit) {
    if (fs != h->stream) CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_fit[1], 0));        // join
    // into the frame state only now: the solver blocks write the whole state back while the frame runs
    CU_TRY(cudaMemcpyAsync((char*)h->d_state + offsetof(FrameState, fitness), h->d_fit + 2 * h->cap_fit, 2 * sizeof(double),
                           cudaMemcpyDeviceToDevice, h->stream));
  }
  // result[16] + {frame_done, status} + {fitness, rmse}: contiguous in FrameState.  Pipelined handles copy it after the
  // graph launch into alternating slots instead (scan_match_enqueue)
  if (!h->async_inputs)
    CU_TRY(cudaMemcpyAsync(h->h_result, (const char*)h->d_state + offsetof(FrameState, result), kFrameResultBytes,
                           cudaMemcpyDeviceToHost, h->stream));
  return TLOAM_B200_OK;
}

static int scan_match_enqueue(tloam_b200_handle* h, const double* predict);

int tloam_b200_scan_match_async(tloam_b200_handle* h, const double predict[16]) {
  if (!h || !predict) return TLOAM_B200_ERR_INVALID_ARG;
  return scan_match_enqueue(h, predict);
}

// (f)-3: the frame is predicted on the device from the two last results (no host input at all)
int tloam_b200_scan_match_predicted_async(tloam_b200_handle* h) {
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  return scan_match_enqueue(h, nullptr);
}

int tloam_b200_scan_match_predicted(tloam_b200_handle* h, double result[16], tloam_b200_stats* stats) {
  const int rc = tloam_b200_scan_match_predicted_async(h);
  if (rc != TLOAM_B200_OK) return rc;
  return tloam_b200_get_result(h, result, stats);
}

int tloam_b200_set_pose_history(tloam_b200_handle* h, const double last_pose[16], const double curr_pose[16]) {
  if (!h || !last_pose || !curr_pose) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaStreamSynchronize(h->stream));
  CU_TRY(cudaMemcpy((char*)h->d_state + offsetof(FrameState, last_pose), last_pose, 16 * sizeof(double), cudaMemcpyHostToDevice));
  CU_TRY(cudaMemcpy((char*)h->d_state + offsetof(FrameState, curr_pose), curr_pose, 16 * sizeof(double), cudaMemcpyHostToDevice));
  return TLOAM_B200_OK;
}

static int scan_match_enqueue(tloam_b200_handle* h, const double* predict) {
  const int rc = check_ready(h);
  if (rc != TLOAM_B200_OK) return rc;
  if (h->async_inputs && h->frames_enqueued - h->frames_fetched >= 2) return TLOAM_B200_ERR_NOT_READY;   // fetch a result first
  if (h->frame_fitness) {
    if (h->cfg.fitness_thres <= 0.0) return TLOAM_B200_ERR_INVALID_ARG;
    for (int c = 0; c < 4; ++c) if (h->cfg.fitness_thres > radius_of(h->cfg, c)) return TLOAM_B200_ERR_INVALID_ARG;
  }
  CU_TRY(cudaSetDevice(h->device));
  if (predict) { memcpy(h->h_predict->m, predict, 16 * sizeof(double)); h->h_predict->from_state = 0.0; }
  else h->h_predict->from_state = 1.0;
  DeviceCtx c = h->ctx;
  if (!h->trace) c.stats = nullptr;             // skip the per-iteration trace (fewer instructions in the serial solver)
  c.dense_mask = dense_mask_of(h);
  if (c.dense_mask && h->dense_kernel == 1) { const int drc = prepare_dense(h, c.dense_mask); if (drc != TLOAM_B200_OK) return drc; }
  const bool fused = h->use_fused && caps_cannot_bind(h) && c.dense_mask == 0;
  const int per_frame = 1 + h->cfg.max_iterations * ((fused ? 1 : 2) + (c.dense_mask ? (h->dense_kernel == 1 ? 4 : 1) : 0) + h->cfg.ceres_max_num_iterations) +
                        (h->frame_fitness ? 2 : 0);
  CU_TRY(cudaEventRecord(h->ev0, h->stream));
  if (h->use_graph && !h->profiling) {
    // one graph launch per frame; the graph is re-captured only when the device context changed
    if (!h->gvalid || h->gfused != fused || h->gfitness != h->frame_fitness || memcmp(&h->gctx, &c, sizeof(DeviceCtx)) != 0 ||
        h->gdense_kernel != h->dense_kernel ||
        (c.dense_mask && h->dense_kernel == 1 && memcmp(&h->gdargs, &h->dargs, sizeof(DenseArgs)) != 0)) {
      h->gvalid = false;
      cudaGraph_t graph = nullptr;
      CU_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
      const long long l0 = h->launches;
      const int erc = enqueue_frame(h, c, fused);
      h->launches = l0;                           // capture does not execute anything
      cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
      if (erc != TLOAM_B200_OK || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        snprintf(h->last_error, sizeof(h->last_error), "graph capture failed: %s", cudaGetErrorString(ce));
        return TLOAM_B200_ERR_CUDA;
      }
      // same topology, new kernel parameters / grid sizes (cloud sizes change from frame to frame in a real
      // stream): update the instantiated graph in place, which is much cheaper than instantiating a new one
      bool updated = false;
      if (h->gexec && h->gfused == fused && h->gfitness == h->frame_fitness && h->gctx.dense_mask == c.dense_mask &&
          h->gdense_kernel == h->dense_kernel) {
        cudaGraphExecUpdateResultInfo info;
        updated = cudaGraphExecUpdate(h->gexec, graph, &info) == cudaSuccess;
        if (!updated) { cudaGetLastError(); cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
      }
      if (!updated) ce = cudaGraphInstantiate(&h->gexec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) { snprintf(h->last_error, sizeof(h->last_error), "graph instantiate: %s", cudaGetErrorString(ce)); return TLOAM_B200_ERR_CUDA; }
      h->gctx = c;
      h->gdargs = h->dargs;
      h->gdense_kernel = h->dense_kernel;
      h->gfused = fused;
      h->gfitness = h->frame_fitness;
      h->gvalid = true;
    }
    CU_TRY(cudaGraphLaunch(h->gexec, h->stream));
    h->launches += per_frame;
  } else {
    const int erc = enqueue_frame(h, c, fused);
    if (erc != TLOAM_B200_OK) return erc;
    CU_TRY(cudaGetLastError());
  }
  CU_TRY(cudaEventRecord(h->ev1, h->stream));
  if (h->async_inputs) {                         // pipelined: result -> slot (frame & 1), one event per slot
    const int p = (int)(h->frames_enqueued & 1);
    CU_TRY(cudaMemcpyAsync(h->h_result + 32 * p, (const char*)h->d_state + offsetof(FrameState, result), kFrameResultBytes,
                           cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaEventRecord(h->ev_res[p], h->stream));
  }
  h->frames_enqueued++;
  h->launches_frame = per_frame;
  h->traced_last = h->trace;
  h->frame_pending = true;
  return TLOAM_B200_OK;
}

int tloam_b200_get_result(tloam_b200_handle* h, double result[16], tloam_b200_stats* stats) {
  if (!h || !result) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->frame_pending) return TLOAM_B200_ERR_NOT_READY;
  CU_TRY(cudaSetDevice(h->device));
  if (stats && !h->traced_last) memset(h->h_stats, 0, sizeof(tloam_b200_stats));
  const double* slot = h->h_result;
  if (h->async_inputs) {
    // pipelined: the OLDEST un-fetched frame; waits for that frame only (the host may already have enqueued the next)
    if (h->frames_fetched >= h->frames_enqueued) return TLOAM_B200_ERR_NOT_READY;
    const int p = (int)(h->frames_fetched & 1);
    CU_TRY(cudaEventSynchronize(h->ev_res[p]));
    slot = h->h_result + 32 * p;
    h->frames_fetched++;
    h->frame_pending = h->frames_fetched < h->frames_enqueued;
    if (stats) memset(h->h_stats, 0, sizeof(tloam_b200_stats));      // no per-iteration trace in pipelined mode
  } else {
    if (stats && h->traced_last) CU_TRY(cudaMemcpyAsync(h->h_stats, h->d_stats, sizeof(tloam_b200_stats), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    h->frame_pending = false;
    h->frames_fetched = h->frames_enqueued;
  }
  harvest_counts(h);
  memcpy(result, slot, 16 * sizeof(double));
  int flags[2];
  memcpy(flags, slot + 16, sizeof(flags));   // frame_done, status
  h->last_fitness = slot[17]; h->last_rmse = slot[18];
  memcpy(h->nbricks, slot + 19, 4 * sizeof(unsigned));   // statistics of the map this frame used: route the next frames
  h->stats_known = true; h->stats_pending = false;
  if (stats) {
    *stats = *h->h_stats;
    stats->gpu_launches = h->launches_frame;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) stats->gpu_ms = ms;
  }
  if (flags[1] != TLOAM_B200_OK) return flags[1];
  if (!flags[0]) { snprintf(h->last_error, sizeof(h->last_error), "frame did not complete"); return TLOAM_B200_ERR_CUDA; }
  return TLOAM_B200_OK;
}

int tloam_b200_scan_match(tloam_b200_handle* h, const double predict[16], double result[16], tloam_b200_stats* stats) {
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  const bool saved = h->trace;
  if (stats) h->trace = true;                   // the blocking form knows whether the caller wants the trace
  int rc = tloam_b200_scan_match_async(h, predict);
  h->trace = saved;
  if (rc != TLOAM_B200_OK) return rc;
  return tloam_b200_get_result(h, result, stats);
}

int tloam_b200_set_trace(tloam_b200_handle* h, int on) {
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  h->trace = on != 0;
  return TLOAM_B200_OK;
}

// Orders the handle's stream behind everything enqueued so far on `producer_stream` (a cudaStream_t; NULL = the legacy
// default stream): device inputs (set_*_device) are read IN PLACE by kernels on the handle's stream, so whoever
// produced them on another stream must be waited for -- on the device, the host does not block.
int tloam_b200_wait_stream(tloam_b200_handle* h, void* producer_stream) {
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  if ((cudaStream_t)producer_stream == h->stream) return TLOAM_B200_OK;
  CU_TRY(cudaEventRecord(h->ev_src, (cudaStream_t)producer_stream));
  CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_src, 0));
  return TLOAM_B200_OK;
}

// self-check counters of the dense correspondence path (TLOAM_B200_DENSE_CHECK=1), accumulated since creation:
// [0] queries searched, [1] kNN lists that differ from the plain search, [2] work items, [3] staging passes
int tloam_b200_dense_check_counters(tloam_b200_handle* h, unsigned out[16]) {
  if (!h || !out) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaStreamSynchronize(h->stream));
  CU_TRY(cudaMemcpy(out, h->d_cnt + 16, 16 * sizeof(unsigned), cudaMemcpyDeviceToHost));
  return TLOAM_B200_OK;
}

int tloam_b200_synchronize(tloam_b200_handle* h) {
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return TLOAM_B200_OK;
}

long long tloam_b200_launch_count(tloam_b200_handle* h) { return h ? h->launches : 0; }

static int read_state_pose(tloam_b200_handle* h, size_t offset, double out[16]) {
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaMemcpyAsync(h->h_result, (const char*)h->d_state + offset, 16 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  memcpy(out, h->h_result, 16 * sizeof(double));
  return TLOAM_B200_OK;
}

int tloam_b200_get_transform(tloam_b200_handle* h, double pose[16]) {   // ref: registration.cpp:370-372
  if (!h || !pose) return TLOAM_B200_ERR_INVALID_ARG;
  return read_state_pose(h, offsetof(FrameState, curr_pose), pose);
}

int tloam_b200_get_pose_increment(tloam_b200_handle* h, double pose[16]) {   // ref: registration.cpp:374-376
  if (!h || !pose) return TLOAM_B200_ERR_INVALID_ARG;
  double L[16], C[16];
  int rc = read_state_pose(h, offsetof(FrameState, last_pose), L);
  if (rc != TLOAM_B200_OK) return rc;
  rc = read_state_pose(h, offsetof(FrameState, curr_pose), C);
  if (rc != TLOAM_B200_OK) return rc;
  double Li[16];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Li[c * 4 + r] = L[r * 4 + c];
  for (int r = 0; r < 3; ++r) Li[12 + r] = -(Li[0 * 4 + r] * L[12] + Li[1 * 4 + r] * L[13] + Li[2 * 4 + r] * L[14]);
  Li[3] = Li[7] = Li[11] = 0.0; Li[15] = 1.0;
  for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) {
    double s = 0.0;
    for (int k = 0; k < 4; ++k) s += Li[k * 4 + r] * C[c * 4 + k];
    pose[c * 4 + r] = s;
  }
  return TLOAM_B200_OK;
}


// ---------------------------------------------------------------------------------------------
// Batched registration: S independent sequences stepped together, ONE launch sequence per batch frame.
// A single 40k-feature frame is 8.5 warps per SM and every evaluation ends in a ~9 us serial tail on one SM
// (DESIGN.md section 4); batching fills the machine with the sequences' parallel phases and runs their tails
// concurrently.  Each sequence keeps its own handle (map, scan, pose history, submap); only the frame kernels
// are shared.  Poses are bit-identical to registering each sequence alone (same per-sequence reduction tree).
// ---------------------------------------------------------------------------------------------
struct tloam_b200_batch {
  int S = 0, device = 0;
  std::vector<tloam_b200_handle*> hs;
  cudaStream_t stream = nullptr;
  FrameState* d_states = nullptr;
  DeviceCtx* d_ctxs = nullptr;
  std::vector<DeviceCtx> ctx_sent;   bool ctx_valid = false;
  Predict* h_pred = nullptr; Predict* d_pred = nullptr;
  unsigned char* h_results = nullptr;          // pinned [S][kResultBytes]
  cudaEvent_t ev_done = nullptr, ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> ev_ready;
  cudaGraphExec_t gexec = nullptr;
  BatchTab g_first, g_corr, g_eval, g_fine; bool gfused = false, gvalid = false;
  bool any_fine = false, gany_fine = false;          // some sequence has clouds served by the two-level search
  bool use_graph = true, use_fused = false, pending = false;
  long long launches = 0; int launches_frame = 0;
  char last_error[512] = {0};
  // optional per-kernel-class timing (CUDA events around every launch of the batch frame; no graph in this mode)
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool; size_t ev_next = 0;
  struct Span { int cls; cudaEvent_t a, b; };
  std::vector<Span> spans;
  tloam_b200_profile prof;
};
struct BatchLaunchScope {
  tloam_b200_batch* b; int cls; cudaEvent_t a = nullptr, e = nullptr;
  BatchLaunchScope(tloam_b200_batch* bb, int c) : b(bb), cls(c) {
    if (b->profiling) {
      while (b->ev_pool.size() < b->ev_next + 2) { cudaEvent_t ev; if (cudaEventCreate(&ev) != cudaSuccess) { b->profiling = false; return; } b->ev_pool.push_back(ev); }
      a = b->ev_pool[b->ev_next++]; e = b->ev_pool[b->ev_next++];
      cudaEventRecord(a, b->stream);
    }
  }
  ~BatchLaunchScope() { if (a && e) { cudaEventRecord(e, b->stream); b->spans.push_back({cls, a, e}); } }
};
#define TLB_LAUNCH(cls, ...) do { BatchLaunchScope ls__(b, cls); __VA_ARGS__; } while (0)
static constexpr size_t kResultBytes = 16 * sizeof(double) + 2 * sizeof(int);   // result + {frame_done, status}

#define CUB_TRY(expr)                                                                                  \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess) {                                                                          \
      snprintf(b->last_error, sizeof(b->last_error), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
               __FILE__, __LINE__);                                                                    \
      return TLOAM_B200_ERR_CUDA;                                                                      \
    }                                                                                                  \
  } while (0)

int tloam_b200_batch_destroy(tloam_b200_batch* b) {
  if (!b) return TLOAM_B200_OK;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  for (tloam_b200_handle* h : b->hs) tloam_b200_destroy(h);
  if (b->gexec) cudaGraphExecDestroy(b->gexec);
  cudaFree(b->d_states); cudaFree(b->d_ctxs); cudaFree(b->d_pred);
  if (b->h_pred) cudaFreeHost(b->h_pred);
  if (b->h_results) cudaFreeHost(b->h_results);
  for (cudaEvent_t e : b->ev_ready) cudaEventDestroy(e);
  for (cudaEvent_t e : b->ev_pool) cudaEventDestroy(e);
  if (b->ev_done) cudaEventDestroy(b->ev_done);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
  return TLOAM_B200_OK;
}

int tloam_b200_batch_create(const tloam_tls_config* cfg, int device, int S, tloam_b200_batch** out) {
  if (!cfg || !out || S < 1 || S > kMaxBatch) return TLOAM_B200_ERR_INVALID_ARG;
  *out = nullptr;
  tloam_b200_batch* b = new (std::nothrow) tloam_b200_batch();
  if (!b) return TLOAM_B200_ERR_INVALID_ARG;
  b->S = S; b->device = device;
  auto fail = [&](int code) { tloam_b200_batch_destroy(b); return code; };
  for (int s = 0; s < S; ++s) {
    tloam_b200_handle* h = nullptr;
    const int rc = tloam_b200_create(cfg, device, nullptr, &h);     // own non-blocking stream: map builds of the sequences overlap
    if (rc != TLOAM_B200_OK) return fail(rc);
    b->hs.push_back(h);
  }
  if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
  if (cudaMalloc(&b->d_states, S * sizeof(FrameState)) != cudaSuccess || cudaMalloc(&b->d_ctxs, S * sizeof(DeviceCtx)) != cudaSuccess ||
      cudaMalloc(&b->d_pred, S * sizeof(Predict)) != cudaSuccess || cudaMallocHost(&b->h_pred, S * sizeof(Predict)) != cudaSuccess ||
      cudaMallocHost(&b->h_results, S * kResultBytes) != cudaSuccess)
    return fail(TLOAM_B200_ERR_CUDA);
  if (cudaEventCreateWithFlags(&b->ev_done, cudaEventDisableTiming) != cudaSuccess || cudaEventCreate(&b->ev0) != cudaSuccess ||
      cudaEventCreate(&b->ev1) != cudaSuccess)
    return fail(TLOAM_B200_ERR_CUDA);
  for (int s = 0; s < S; ++s) {
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
    b->ev_ready.push_back(e);
    // the sequences' frame states live in ONE array (one strided D2H copy fetches every result)
    tloam_b200_handle* h = b->hs[s];
    if (cudaMemcpy(b->d_states + s, h->d_state, sizeof(FrameState), cudaMemcpyDeviceToDevice) != cudaSuccess) return fail(TLOAM_B200_ERR_CUDA);
    cudaFree(h->d_state);
    h->d_state = b->d_states + s; h->own_state = false;
  }
  { const char* e = getenv("TLOAM_B200_NO_GRAPH"); b->use_graph = !(e && e[0] == '1'); }
  { const char* e = getenv("TLOAM_B200_FUSE"); b->use_fused = (e && e[0] == '1'); }
  { const char* e = getenv("TLOAM_B200_NO_FUSE"); if (e && e[0] == '1') b->use_fused = false; }
  *out = b;
  return TLOAM_B200_OK;
}

int tloam_b200_batch_size(tloam_b200_batch* b) { return b ? b->S : 0; }
tloam_b200_handle* tloam_b200_batch_handle(tloam_b200_batch* b, int i) { return (b && i >= 0 && i < b->S) ? b->hs[i] : nullptr; }
const char* tloam_b200_batch_last_error(tloam_b200_batch* b) { return b ? b->last_error : ""; }
long long tloam_b200_batch_launch_count(tloam_b200_batch* b) {
  if (!b) return 0;
  long long n = b->launches;
  for (tloam_b200_handle* h : b->hs) n += h->launches;
  return n;
}

static int batch_set(tloam_b200_batch* b, const double* const* xyz, const size_t* n, bool target, bool on_device) {
  if (!b || !xyz || !n) return TLOAM_B200_ERR_INVALID_ARG;
  for (int s = 0; s < b->S; ++s) {
    const int rc = target ? set_target_impl(b->hs[s], xyz + 4 * s, n + 4 * s, on_device) : set_source_impl(b->hs[s], xyz + 4 * s, n + 4 * s, on_device);
    if (rc != TLOAM_B200_OK) {
      snprintf(b->last_error, sizeof(b->last_error), "sequence %d: %.400s", s, b->hs[s]->last_error);
      return rc;
    }
  }
  return TLOAM_B200_OK;
}
int tloam_b200_batch_set_target(tloam_b200_batch* b, const double* const* xyz, const size_t* n) { return batch_set(b, xyz, n, true, false); }
int tloam_b200_batch_set_source(tloam_b200_batch* b, const double* const* xyz, const size_t* n) { return batch_set(b, xyz, n, false, false); }
int tloam_b200_batch_set_target_device(tloam_b200_batch* b, const double* const* xyz, const size_t* n) { return batch_set(b, xyz, n, true, true); }
int tloam_b200_batch_set_source_device(tloam_b200_batch* b, const double* const* xyz, const size_t* n) { return batch_set(b, xyz, n, false, true); }

static int batch_enqueue_frame(tloam_b200_batch* b, bool fused, const tloam_tls_config& cfg) {
  static const DeviceCtx zero_ctx = {};
  const int S = b->S;
  CUB_TRY(cudaMemcpyAsync(b->d_pred, b->h_pred, S * sizeof(Predict), cudaMemcpyHostToDevice, b->stream));
  TLB_LAUNCH(TLOAM_B200_K_BEGIN_FRAME, (k_begin_frame<true><<<S, 256, 0, b->stream>>>(zero_ctx, b->g_eval, b->d_pred)));
  for (int outer = 0; outer < cfg.max_iterations; ++outer) {
    if (fused) {
      TLB_LAUNCH(TLOAM_B200_K_FIRST, (k_first<true><<<b->g_first.off[S], kBlk, 0, b->stream>>>(zero_ctx, b->g_first)));
    } else {
      if (b->any_fine) TLB_LAUNCH(TLOAM_B200_K_FINE, (k_correspond_fine<true><<<b->g_fine.off[S], kBlk, kFineSmemBytes, b->stream>>>(zero_ctx, b->g_fine, nullptr)));
      TLB_LAUNCH(TLOAM_B200_K_CORRESPOND, (k_correspond<true><<<b->g_corr.off[S], kBlk, 0, b->stream>>>(zero_ctx, b->g_corr)));
      TLB_LAUNCH(TLOAM_B200_K_EVAL_FIRST, (k_eval<true, true><<<b->g_eval.off[S], kBlk, 0, b->stream>>>(zero_ctx, b->g_eval)));
    }
    for (int it = 0; it < cfg.ceres_max_num_iterations; ++it)
      TLB_LAUNCH(TLOAM_B200_K_EVAL, (k_eval<false, true><<<b->g_eval.off[S], kBlk, 0, b->stream>>>(zero_ctx, b->g_eval)));
  }
  CUB_TRY(cudaGetLastError());
  // every sequence's result[16] + {frame_done, status} with ONE strided copy
  CUB_TRY(cudaMemcpy2DAsync(b->h_results, kResultBytes, (const char*)b->d_states + offsetof(FrameState, result), sizeof(FrameState),
                            kResultBytes, S, cudaMemcpyDeviceToHost, b->stream));
  return TLOAM_B200_OK;
}

// predicts: S x 16 doubles (4x4 column-major each), or NULL = device-side constant-velocity prediction per sequence
int tloam_b200_batch_scan_match_async(tloam_b200_batch* b, const double* predicts) {
  if (!b) return TLOAM_B200_ERR_INVALID_ARG;
  const int S = b->S;
  for (int s = 0; s < S; ++s) {
    const int rc = check_ready(b->hs[s]);
    if (rc != TLOAM_B200_OK) { snprintf(b->last_error, sizeof(b->last_error), "sequence %d is not ready", s); return rc; }
  }
  CUB_TRY(cudaSetDevice(b->device));
  const tloam_tls_config& cfg = b->hs[0]->cfg;
  // the frame kernels run behind everything the sequences have enqueued on their own streams (map build, staging)
  std::vector<DeviceCtx> ctxs(S);
  BatchTab tf, tc, te, tq;
  memset(&tf, 0, sizeof(tf)); memset(&tc, 0, sizeof(tc)); memset(&te, 0, sizeof(te)); memset(&tq, 0, sizeof(tq));
  tf.S = tc.S = te.S = tq.S = S; tf.ctxs = tc.ctxs = te.ctxs = tq.ctxs = b->d_ctxs;
  bool fused = b->use_fused;
  bool any_fine = false;
  for (int s = 0; s < S; ++s) {
    tloam_b200_handle* h = b->hs[s];
    CUB_TRY(cudaEventRecord(b->ev_ready[s], h->stream));
    CUB_TRY(cudaStreamWaitEvent(b->stream, b->ev_ready[s], 0));
    ctxs[s] = h->ctx;
    ctxs[s].stats = nullptr; ctxs[s].dbg = nullptr;
    ctxs[s].dense_mask = fine_mask_of(h);            // dense map clouds: two-level search (the TMA-staged path is single-sequence only)
    any_fine = any_fine || ctxs[s].dense_mask != 0;
    fused = fused && caps_cannot_bind(h);
    const int nb = h->total_blocks;
    tq.off[s + 1] = tq.off[s] + nb;
    tf.off[s + 1] = tf.off[s] + first_grid_of(nb);
    tc.off[s + 1] = tc.off[s] + 2 * nb;
    te.off[s + 1] = te.off[s] + eval_grid_of(nb);
    if (predicts) { memcpy(b->h_pred[s].m, predicts + 16 * s, 16 * sizeof(double)); b->h_pred[s].from_state = 0.0; }
    else b->h_pred[s].from_state = 1.0;
  }
  if (!b->ctx_valid || memcmp(b->ctx_sent.data(), ctxs.data(), S * sizeof(DeviceCtx)) != 0) {
    CUB_TRY(cudaMemcpyAsync(b->d_ctxs, ctxs.data(), S * sizeof(DeviceCtx), cudaMemcpyHostToDevice, b->stream));   // pageable source: staged before return
    b->ctx_sent = ctxs; b->ctx_valid = true;
  }
  if (any_fine) fused = false;
  b->any_fine = any_fine;
  const int per_frame = 1 + cfg.max_iterations * ((fused ? 1 : 2) + (any_fine ? 1 : 0) + cfg.ceres_max_num_iterations);
  CUB_TRY(cudaEventRecord(b->ev0, b->stream));
  if (b->use_graph && !b->profiling) {
    const bool same = b->gvalid && b->gfused == fused && b->gany_fine == any_fine && memcmp(&b->g_first, &tf, sizeof(tf)) == 0 &&
                      memcmp(&b->g_corr, &tc, sizeof(tc)) == 0 && memcmp(&b->g_eval, &te, sizeof(te)) == 0;
    if (!same) {
      b->gvalid = false;
      b->g_first = tf; b->g_corr = tc; b->g_eval = te; b->g_fine = tq;
      cudaGraph_t graph = nullptr;
      CUB_TRY(cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal));
      const int erc = batch_enqueue_frame(b, fused, cfg);
      cudaError_t ce = cudaStreamEndCapture(b->stream, &graph);
      if (erc != TLOAM_B200_OK || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (erc == TLOAM_B200_OK) snprintf(b->last_error, sizeof(b->last_error), "graph capture failed: %s", cudaGetErrorString(ce));
        return TLOAM_B200_ERR_CUDA;
      }
      bool updated = false;
      if (b->gexec && b->gfused == fused && b->gany_fine == any_fine) {
        cudaGraphExecUpdateResultInfo info;
        updated = cudaGraphExecUpdate(b->gexec, graph, &info) == cudaSuccess;
        if (!updated) { cudaGetLastError(); cudaGraphExecDestroy(b->gexec); b->gexec = nullptr; }
      } else if (b->gexec) { cudaGraphExecDestroy(b->gexec); b->gexec = nullptr; }
      if (!updated) ce = cudaGraphInstantiate(&b->gexec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) { snprintf(b->last_error, sizeof(b->last_error), "graph instantiate: %s", cudaGetErrorString(ce)); return TLOAM_B200_ERR_CUDA; }
      b->gfused = fused; b->gany_fine = any_fine; b->gvalid = true;
    }
    CUB_TRY(cudaGraphLaunch(b->gexec, b->stream));
  } else {
    b->g_first = tf; b->g_corr = tc; b->g_eval = te; b->g_fine = tq;
    const int erc = batch_enqueue_frame(b, fused, cfg);
    if (erc != TLOAM_B200_OK) return erc;
  }
  CUB_TRY(cudaEventRecord(b->ev1, b->stream));
  // whatever the sequences enqueue next on their own streams (the next map build overwrites the map in use) waits
  CUB_TRY(cudaEventRecord(b->ev_done, b->stream));
  for (int s = 0; s < S; ++s) CUB_TRY(cudaStreamWaitEvent(b->hs[s]->stream, b->ev_done, 0));
  b->launches += per_frame; b->launches_frame = per_frame;
  b->pending = true;
  return TLOAM_B200_OK;
}

// results: S x 16 doubles; statuses: S ints (tloam_b200_status per sequence).  Returns OK when every sequence is OK,
// else the first non-OK status.
int tloam_b200_batch_get_results(tloam_b200_batch* b, double* results, int* statuses, float* gpu_ms) {
  if (!b || !results) return TLOAM_B200_ERR_INVALID_ARG;
  if (!b->pending) return TLOAM_B200_ERR_NOT_READY;
  CUB_TRY(cudaSetDevice(b->device));
  CUB_TRY(cudaStreamSynchronize(b->stream));
  b->pending = false;
  int worst = TLOAM_B200_OK;
  for (int s = 0; s < b->S; ++s) {
    const unsigned char* r = b->h_results + s * kResultBytes;
    memcpy(results + 16 * s, r, 16 * sizeof(double));
    int flags[2];
    memcpy(flags, r + 16 * sizeof(double), sizeof(flags));
    int st = flags[1];
    if (st == TLOAM_B200_OK && !flags[0]) st = TLOAM_B200_ERR_CUDA;      // the frame did not complete
    if (statuses) statuses[s] = st;
    if (worst == TLOAM_B200_OK && st != TLOAM_B200_OK) worst = st;
  }
  if (gpu_ms) { float ms = 0.f; if (cudaEventElapsedTime(&ms, b->ev0, b->ev1) == cudaSuccess) *gpu_ms = ms; }
  return worst;
}

int tloam_b200_batch_set_profiling(tloam_b200_batch* b, int on) {
  if (!b) return TLOAM_B200_ERR_INVALID_ARG;
  CUB_TRY(cudaSetDevice(b->device));
  CUB_TRY(cudaStreamSynchronize(b->stream));
  b->profiling = on != 0;
  b->spans.clear(); b->ev_next = 0;
  memset(&b->prof, 0, sizeof(b->prof));
  return TLOAM_B200_OK;
}

int tloam_b200_batch_get_profile(tloam_b200_batch* b, tloam_b200_profile* out) {
  if (!b || !out) return TLOAM_B200_ERR_INVALID_ARG;
  CUB_TRY(cudaSetDevice(b->device));
  CUB_TRY(cudaStreamSynchronize(b->stream));
  for (const auto& sp : b->spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess && sp.cls >= 0 && sp.cls < TLOAM_B200_K_COUNT) {
      b->prof.launches[sp.cls] += 1;
      b->prof.total_ms[sp.cls] += ms;
    }
  }
  b->spans.clear(); b->ev_next = 0;
  *out = b->prof;
  return TLOAM_B200_OK;
}

int tloam_b200_batch_scan_match(tloam_b200_batch* b, const double* predicts, double* results, int* statuses) {
  const int rc = tloam_b200_batch_scan_match_async(b, predicts);
  if (rc != TLOAM_B200_OK) return rc;
  return tloam_b200_batch_get_results(b, results, statuses, nullptr);
}

int tloam_b200_fitness(tloam_b200_handle* h, double* fitness, double* rmse) {
  if (!h || !fitness || !rmse) return TLOAM_B200_ERR_INVALID_ARG;
  *fitness = 0.0; *rmse = 0.0;
  if (!h->have_src || !h->have_tgt) return TLOAM_B200_ERR_NOT_READY;
  if (h->cfg.fitness_thres <= 0.0) return TLOAM_B200_OK;                 // ref: :258-261
  for (int c = 0; c < 4; ++c) if (h->cfg.fitness_thres > radius_of(h->cfg, c)) return TLOAM_B200_ERR_INVALID_ARG;
  const int nb = h->total_blocks;
  if (nb == 0) return TLOAM_B200_OK;                                     // every source cloud is empty: nothing matches
  CU_TRY(cudaSetDevice(h->device));
  // block partials live in a buffer sized with the source (no allocation on this per-frame health metric); the
  // per-cloud sums are formed on the device in block order and land in the frame state
  TL_LAUNCH(TLOAM_B200_K_FITNESS, (k_fitness<<<nb, kBlk, 0, h->stream>>>(h->ctx, h->cfg.fitness_thres * h->cfg.fitness_thres, h->d_fit)));
  TL_LAUNCH(TLOAM_B200_K_FITNESS, (k_fitness_reduce<<<1, 128, 0, h->stream>>>(h->ctx, h->d_fit, reinterpret_cast<double*>(reinterpret_cast<char*>(h->d_state) + offsetof(FrameState, fitness)))));
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaMemcpyAsync(h->h_result + 20, (const char*)h->d_state + offsetof(FrameState, fitness), 2 * sizeof(double),
                         cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  *fitness = h->h_result[20]; *rmse = h->h_result[21];
  return TLOAM_B200_OK;
}

// per-frame health metric in the asynchronous flow: every scan_match also evaluates getFitnessScore of its scan
// (two more kernels in the frame graph, no allocation); the pair comes back with the frame's result
int tloam_b200_set_frame_fitness(tloam_b200_handle* h, int on) {
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  h->frame_fitness = on != 0;
  return TLOAM_B200_OK;
}
int tloam_b200_get_frame_fitness(tloam_b200_handle* h, double* fitness, double* rmse) {   // of the last fetched result
  if (!h || !fitness || !rmse) return TLOAM_B200_ERR_INVALID_ARG;
  *fitness = h->last_fitness; *rmse = h->last_rmse;
  return TLOAM_B200_OK;
}
// Pipelined use: set_source / submap_update return without waiting for their uploads (the host buffers must then stay
// valid until the frame's result has been fetched) and get_result waits for the OLDEST un-fetched frame only, so the
// host can enqueue frame k+1 while the GPU still runs frame k (at most 2 frames in flight).
int tloam_b200_set_async_inputs(tloam_b200_handle* h, int on) {
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaStreamSynchronize(h->stream));
  h->async_inputs = on != 0;
  h->frames_fetched = h->frames_enqueued; h->frame_pending = false;
  h->gvalid = false;                             // the frame graph holds (or not) the result copy
  return TLOAM_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// piecewise entry points (tests)
// ---------------------------------------------------------------------------------------------
int tloam_b200_knn(tloam_b200_handle* h, int cloud, const double* queries, size_t nq, double radius, int k, int* idx,
                   double* d2, int* count) {
  if (!h || cloud < 0 || cloud > 3 || !queries || !idx || !d2 || !count) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->have_tgt) return TLOAM_B200_ERR_NOT_READY;
  if (!(radius > 0.0) || radius > h->hdr.cell[cloud] || (k != 1 && k != 3 && k != 5)) return TLOAM_B200_ERR_INVALID_ARG;
  if (nq == 0) return TLOAM_B200_OK;
  CU_TRY(cudaSetDevice(h->device));
  { const int rc = fetch_origin(h); if (rc != TLOAM_B200_OK) return rc; }
  double *dq = nullptr, *dd = nullptr; int *di = nullptr, *dc = nullptr;
  cudaError_t e = cudaMalloc(&dq, nq * 3 * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&dd, nq * k * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&di, nq * k * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&dc, nq * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpyAsync(dq, queries, nq * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    const unsigned tb = 128, gb = (unsigned)((nq + tb - 1) / tb);
    const GridDesc g = h->ctx.grid[cloud];
    const double* o = h->ctx.origin;
    if (k == 1) k_knn<1><<<gb, tb, 0, h->stream>>>(g, o, dq, (unsigned)nq, radius * radius, di, dd, dc);
    else if (k == 3) k_knn<3><<<gb, tb, 0, h->stream>>>(g, o, dq, (unsigned)nq, radius * radius, di, dd, dc);
    else k_knn<5><<<gb, tb, 0, h->stream>>>(g, o, dq, (unsigned)nq, radius * radius, di, dd, dc);
    h->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(idx, di, nq * k * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d2, dd, nq * k * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(count, dc, nq * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(dq); cudaFree(dd); cudaFree(di); cudaFree(dc);
  if (e != cudaSuccess) { snprintf(h->last_error, sizeof(h->last_error), "knn: %s", cudaGetErrorString(e)); return TLOAM_B200_ERR_CUDA; }
  return TLOAM_B200_OK;
}

int tloam_b200_build_factors(tloam_b200_handle* h, int cloud, const double x[6], int* valid, double* prim, size_t n) {
  if (!h || cloud < 0 || cloud > 3 || !x || !valid || !prim) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->have_src || !h->have_tgt) return TLOAM_B200_ERR_NOT_READY;
  if (n != h->n_src[cloud]) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  { const int rc = fetch_origin(h); if (rc != TLOAM_B200_OK) return rc; }
  Predict pr;
  memset(&pr, 0, sizeof(pr));
  memcpy(pr.m, x, 6 * sizeof(double));
  DeviceCtx c = h->ctx;
  c.factor_num = 4;   // build every cloud regardless of the configured subset
  c.dense_mask = dense_mask_of(h);
  if (c.dense_mask && h->dense_kernel == 1) { const int drc = prepare_dense(h, c.dense_mask); if (drc != TLOAM_B200_OK) return drc; }
  k_set_pose<<<1, 256, 0, h->stream>>>(c, pr);
  { const int crc = enqueue_correspond(h, c); if (crc != TLOAM_B200_OK) return crc; }
  k_caps<<<h->total_blocks, kBlk, 0, h->stream>>>(c);
  h->launches += 2;
  CU_TRY(cudaGetLastError());
  std::vector<unsigned char> act(n);
  std::vector<double> col(n);
  const size_t off = (size_t)h->ctx.pad_off[cloud];
  CU_TRY(cudaMemcpyAsync(act.data(), h->ctx.active + off, n, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i < n; ++i) valid[i] = act[i];
  for (int j = 0; j < 6; ++j) {
    CU_TRY(cudaMemcpyAsync(col.data(), h->ctx.prim[j] + off, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < n; ++i) prim[6 * i + j] = act[i] ? col[i] : 0.0;
  }
  // leave the handle idle
  int one = 1;
  CU_TRY(cudaMemcpyAsync((char*)h->d_state + offsetof(FrameState, frame_done), &one, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return TLOAM_B200_OK;
}

static int run_functor(tloam_b200_handle* h, int type, const double x[6], size_t m, const double* p, const double* a,
                       const double* b, size_t b_stride, const double* w, double* r, size_t r_stride, double* J,
                       size_t j_stride, double* cost) {
  if (!h || !x || !p || !a || !b || !w || !r || !J || !cost) return TLOAM_B200_ERR_INVALID_ARG;
  if (m == 0) return TLOAM_B200_OK;
  CU_TRY(cudaSetDevice(h->device));
  Predict pr;
  memset(&pr, 0, sizeof(pr));
  memcpy(pr.m, x, 6 * sizeof(double));
  const size_t in_doubles = m * (3 + 3 + b_stride + 1), out_doubles = m * (r_stride + j_stride + 1);
  double* d = nullptr;
  cudaError_t e = cudaMalloc(&d, (in_doubles + out_doubles) * sizeof(double));
  if (e != cudaSuccess) { snprintf(h->last_error, sizeof(h->last_error), "functor: %s", cudaGetErrorString(e)); return TLOAM_B200_ERR_CUDA; }
  double *dp = d, *da = dp + 3 * m, *db = da + 3 * m, *dw = db + b_stride * m, *dr = dw + m, *dJ = dr + r_stride * m, *dc = dJ + j_stride * m;
  e = cudaMemcpyAsync(dp, p, 3 * m * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(da, a, 3 * m * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(db, b, b_stride * m * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dw, w, m * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    k_functor<<<(unsigned)((m + 127) / 128), 128, 0, h->stream>>>(type, pr, (unsigned)m, dp, da, db, dw, dr, dJ, dc);
    h->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(r, dr, r_stride * m * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(J, dJ, j_stride * m * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(cost, dc, m * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  if (e != cudaSuccess) { snprintf(h->last_error, sizeof(h->last_error), "functor: %s", cudaGetErrorString(e)); return TLOAM_B200_ERR_CUDA; }
  return TLOAM_B200_OK;
}

int tloam_b200_eval_point_to_point(tloam_b200_handle* h, const double x[6], size_t m, const double* p, const double* q,
                                   const double* w, double* r, double* J, double* cost) {
  return run_functor(h, 0, x, m, p, q, q, 3, w, r, 3, J, 18, cost);
}
int tloam_b200_eval_point_to_line(tloam_b200_handle* h, const double x[6], size_t m, const double* p, const double* a,
                                  const double* b, const double* w, double* r, double* J, double* cost) {
  return run_functor(h, 1, x, m, p, a, b, 3, w, r, 3, J, 18, cost);
}
int tloam_b200_eval_point_to_plane(tloam_b200_handle* h, const double x[6], size_t m, const double* p, const double* n,
                                   const double* d, const double* w, double* r, double* J, double* cost) {
  return run_functor(h, 2, x, m, p, n, d, 1, w, r, 1, J, 6, cost);
}

static int run_se3(tloam_b200_handle* h, int op, const double* in, int nin, double* out, int nout, double* extra) {
  if (!h || !in || !out) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  Predict pr;
  memset(&pr, 0, sizeof(pr));
  memcpy(pr.m, in, nin * sizeof(double));
  double* d = nullptr;
  CU_TRY(cudaMalloc(&d, 32 * sizeof(double)));
  k_se3<<<1, 32, 0, h->stream>>>(op, pr, d);
  h->launches++;
  double tmp[32];
  cudaError_t e = cudaMemcpyAsync(tmp, d, 32 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  if (e != cudaSuccess) { snprintf(h->last_error, sizeof(h->last_error), "se3: %s", cudaGetErrorString(e)); return TLOAM_B200_ERR_CUDA; }
  memcpy(out, tmp, nout * sizeof(double));
  if (extra) *extra = tmp[6];
  return TLOAM_B200_OK;
}

int tloam_b200_se3_exp(tloam_b200_handle* h, const double a[6], double T[16]) { return run_se3(h, 0, a, 6, T, 16, nullptr); }
int tloam_b200_se3_log(tloam_b200_handle* h, const double T[16], double a[6]) {
  double ok = 1.0;
  int rc = run_se3(h, 1, T, 16, a, 6, &ok);
  if (rc != TLOAM_B200_OK) return rc;
  return ok != 0.0 ? TLOAM_B200_OK : TLOAM_B200_ERR_BAD_POSE;
}
int tloam_b200_se3_plus(tloam_b200_handle* h, const double x[6], const double delta[6], double out[6]) {
  double in[12];
  memcpy(in, x, 48); memcpy(in + 6, delta, 48);
  return run_se3(h, 2, in, 12, out, 6, nullptr);
}

int tloam_b200_min_on_boundary_2d(tloam_b200_handle* h, const double B[4], const double g[2], double radius, double y[2]) {
  if (!B || !g || !y) return TLOAM_B200_ERR_INVALID_ARG;
  double in[7] = {B[0], B[1], B[2], B[3], g[0], g[1], radius};
  return run_se3(h, 3, in, 7, y, 2, nullptr);
}

// ---------------------------------------------------------------------------------------------
// (f)-1 device-side submap maintenance
// ---------------------------------------------------------------------------------------------
void tloam_b200_submap_default_config(tloam_submap_config* c) {   // ref: config/mapping/lidar_odometry.yaml:6-17
  c->ground_down_sample = 0.3; c->ground_down_sample_submap = 0.45; c->edge_down_sample_submap = 0.3;
  c->planar_frame_size = 3; c->sphere_frame_size = 3;
  c->edge_crop_box_length = 100.0; c->ground_crop_box_length = 100.0;
}

static int ensure_dev(tloam_b200_handle* h, double** p, size_t* cap, size_t need_points, bool keep) {
  if (need_points <= *cap) return TLOAM_B200_OK;
  const size_t ncap = need_points + need_points / 2 + 1024;
  double* q = nullptr;
  CU_TRY(cudaMalloc(&q, ncap * 3 * sizeof(double)));
  if (keep && *p && *cap) CU_TRY(cudaMemcpyAsync(q, *p, *cap * 3 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  if (*p) { CU_TRY(cudaStreamSynchronize(h->stream)); cudaFree(*p); }
  *p = q; *cap = ncap;
  return TLOAM_B200_OK;
}

// crop + VoxelDownSample of d_in into d_out, enqueued without any host round trip: the voxel count lands in
// *out_count (device).  The input holds n_bound points at most; its exact count is n_bound itself (n_dev == nullptr)
// or *n_dev + n_add.  The crop box is lo/hi (host values; nullptr = none) or pose.t +- box_len with the pose in
// device memory.
static int voxel_pipeline(tloam_b200_handle* h, const double* d_in, size_t n_bound, const unsigned* n_dev, unsigned n_add,
                          const double* lo, const double* hi, const double* box_pose, double box_len, double voxel,
                          double* d_out, unsigned* out_count, cudaStream_t stream = nullptr, int scratch = 0) {
  if (!stream) stream = h->stream;
  if (n_bound == 0) { CU_TRY(cudaMemsetAsync(out_count, 0, sizeof(unsigned), stream)); return TLOAM_B200_OK; }
  if (!(voxel > 0.0)) return TLOAM_B200_ERR_INVALID_ARG;
  const unsigned tsize = next_pow2(2 * n_bound + 1);
  const size_t bytes = 256 + (size_t)tsize * (8 + 24 + 4);
  unsigned char*& d_vox = scratch ? h->d_vox1 : h->d_vox;
  size_t& cap_vox = scratch ? h->cap_vox1 : h->cap_vox;
  if (bytes > cap_vox) {
    CU_TRY(cudaStreamSynchronize(stream));
    cudaFree(d_vox); d_vox = nullptr; cap_vox = 0;
    CU_TRY(cudaMalloc(&d_vox, bytes + bytes / 2));
    cap_vox = bytes + bytes / 2;
  }
  VoxArgs a;
  a.in = d_in; a.n = (unsigned)n_bound; a.voxel = voxel;
  a.n_dev = n_dev; a.n_add = n_add; a.box_pose = box_pose; a.box_len = box_len;
  for (int d = 0; d < 3; ++d) { a.lo[d] = lo ? lo[d] : -DBL_MAX; a.hi[d] = hi ? hi[d] : DBL_MAX; }
  a.minenc = reinterpret_cast<unsigned long long*>(d_vox);            // [0..2] min bound
  a.out_count = out_count;
  a.keys = reinterpret_cast<unsigned long long*>(d_vox + 256);
  a.sums = reinterpret_cast<long long*>(d_vox + 256 + (size_t)tsize * 8);
  a.cnt = reinterpret_cast<unsigned*>(d_vox + 256 + (size_t)tsize * 32);
  a.mask = tsize - 1u;
  a.out = d_out;
  CU_TRY(cudaMemsetAsync(d_vox, 0, 256 + (size_t)tsize * 36, stream));   // ONE memset: min bound (complemented), keys, sums, counts
  const unsigned tb = 256, gb = (unsigned)((n_bound + tb - 1) / tb);
  TL_LAUNCH(TLOAM_B200_K_SUBMAP, (k_vox_min<<<(gb < 592u ? gb : 592u), tb, 0, stream>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_SUBMAP, (k_vox_accum<<<gb, tb, 0, stream>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_SUBMAP, (k_vox_emit<<<(tsize + tb - 1) / tb, tb, 0, stream>>>(a)));
  CU_TRY(cudaGetLastError());
  return TLOAM_B200_OK;
}

static int upload_points(tloam_b200_handle* h, const double* host, size_t n) {
  int rc = ensure_dev(h, &h->d_up, &h->cap_up, n, false);
  if (rc != TLOAM_B200_OK) return rc;
  if (n) CU_TRY(cudaMemcpyAsync(h->d_up, host, n * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return TLOAM_B200_OK;
}

static unsigned* acc_count(tloam_b200_handle* h, int k) { return h->d_cnt + 4 + 2 * k + h->acc_cur[k]; }

static int submap_set_target(tloam_b200_handle* h) {
  // order at the ABI: edge, sphere, planar, ground.  After the first update the sphere map IS the planar window
  // (ref: front_end.cpp:220-230 iterates submap_planar_buffer).  Edge / ground: the host knows an upper bound, the
  // exact counts are read on the device.
  const double* xyz[4] = {h->d_acc[0], h->sphere_is_init ? h->d_sphere0 : h->d_cat, h->d_cat, h->d_acc[1]};
  const size_t n[4] = {h->n_acc[0], h->sphere_is_init ? h->n_sphere0 : h->n_cat, h->n_cat, h->n_acc[1]};
  const unsigned* nd[4] = {acc_count(h, 0), nullptr, nullptr, acc_count(h, 1)};
  return set_target_impl(h, xyz, n, true, nd);
}

// ---------------------------------------------------------------------------------------------
// "next" row (f)-2: PCA feature extraction (feature_extract.cuh)
// ---------------------------------------------------------------------------------------------
void tloam_b200_feature_default_config(tloam_feature_config* c) {   // ref: config/mapping/feature.yaml
  c->radius = 0.2; c->K = 20; c->min_neigh = 10; c->planar_num = 500; c->sphere_num = 300;
  c->cvr_scan = 0.25; c->cvr_submap = 0.15; c->planar_scan_thres = 0.75; c->planar_submap_thres = 0.65;
  c->planar_vertic_thres = 0.25;
}

namespace {
struct FeArena {
  double* stage; unsigned* scratch; unsigned char* blob; MapHeader hdr;
  FeOut out;
  unsigned long long *key_p_sorted, *key_s_sorted;
  unsigned *val_p_sorted, *val_s_sorted, *counts;
  unsigned *host_val_p = nullptr, *host_val_s = nullptr;   // optional: the sorted index lists land here (one sync)
};
}  // namespace

// carves the arena for n points and enqueues grid build + k_fe_pca (+ classification and sorts when `select`)
static int fe_run(tloam_b200_handle* h, const tloam_feature_config* cfg, const double* xyz, size_t n, bool select, FeArena& A) {
  if (n == 0 || n > ((size_t)1 << 30)) return TLOAM_B200_ERR_INVALID_ARG;
  if (!(cfg->radius > 0.0) || cfg->K < 3 || cfg->K > kFeK) return TLOAM_B200_ERR_INVALID_ARG;   // :56 asserts K >= 3
  CU_TRY(cudaSetDevice(h->device));
  MapHeader& hd = A.hdr;
  memset(&hd, 0, sizeof(hd));
  hd.magic = kMapMagic;
  size_t boff = sizeof(MapHeader);
  for (int c = 0; c < 4; ++c) {
    hd.n[c] = c == 0 ? (unsigned)n : 0u;
    hd.tsize[c] = next_pow2((c == 0 ? n : 0) + 1);
    hd.cell[c] = cfg->radius;                       // cell edge == search radius: the 27-cell search is exact
    hd.pts_off[c] = boff; boff += round_up(hd.n[c] * sizeof(FePoint), 256);
  }
  for (int c = 0; c < 4; ++c) { hd.table_off[c] = boff; boff += (size_t)hd.tsize[c] * kBrickBytes; }
  for (int d = 0; d < 3; ++d) { hd.bbox_enc[d] = ~0ull; hd.bbox_enc[3 + d] = 0ull; }
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += round_up(bytes, 256); return o; };
  const size_t o_stage = take(n * 3 * sizeof(double)), o_scr = take(n * 2 * sizeof(unsigned)), o_blob = take(boff);
  const size_t o_cvr = take(n * 8), o_flat = take(n * 8), o_sph = take(n * 8), o_nrm = take(n * 24), o_num = take(n * 4),
               o_nei = take(n * kFeK * 4);
  size_t npad = 1;                                     // the bitonic network pads each candidate list to a power of two
  while (npad < n) npad <<= 1;
  const size_t o_kps = take(npad * 8), o_kss = take(npad * 8), o_vps = take(npad * 4), o_vss = take(npad * 4), o_cnt = take(256);
  const size_t o_kpr = take(npad * 8), o_ksr = take(npad * 8), o_vpr = take(npad * 4), o_vsr = take(npad * 4);   // compacted, unsorted
  const size_t o_rk = take(npad * 8);                  // ranks of the two lists
  if (off > h->cap_fe) {
    cudaFree(h->d_fe);
    h->cap_fe = off + off / 4;
    CU_TRY(cudaMalloc(&h->d_fe, h->cap_fe));
  }
  unsigned char* b = h->d_fe;
  A.stage = (double*)(b + o_stage); A.scratch = (unsigned*)(b + o_scr); A.blob = b + o_blob;
  A.out.cvr = (double*)(b + o_cvr); A.out.flatness = (double*)(b + o_flat); A.out.sphericity = (double*)(b + o_sph);
  A.out.normal = (double*)(b + o_nrm); A.out.num_sum = (int*)(b + o_num); A.out.neigh = (int*)(b + o_nei);
  A.key_p_sorted = (unsigned long long*)(b + o_kps); A.key_s_sorted = (unsigned long long*)(b + o_kss);
  A.val_p_sorted = (unsigned*)(b + o_vps); A.val_s_sorted = (unsigned*)(b + o_vss);
  A.counts = (unsigned*)(b + o_cnt);
  unsigned long long* key_p_raw = (unsigned long long*)(b + o_kpr); unsigned long long* key_s_raw = (unsigned long long*)(b + o_ksr);
  unsigned* val_p_raw = (unsigned*)(b + o_vpr); unsigned* val_s_raw = (unsigned*)(b + o_vsr);
  unsigned* rank_p = (unsigned*)(b + o_rk); unsigned* rank_s = rank_p + npad;

  CU_TRY(cudaMemcpyAsync(A.stage, xyz, n * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CU_TRY(cudaMemcpyAsync(A.blob, &hd, sizeof(MapHeader), cudaMemcpyHostToDevice, h->stream));
  CU_TRY(cudaMemsetAsync(A.blob + hd.table_off[0], 0, boff - hd.table_off[0], h->stream));
  CU_TRY(cudaMemsetAsync(A.counts, 0, 256, h->stream));
  MapBuildArgs ma;
  ma.blob = A.blob; ma.slot_of = A.scratch; ma.rank_of = A.scratch + n;
  ma.i_beg = 0; ma.i_end = (unsigned)n; ma.only_cloud = -1;
  ma.fine_mask = 0; ma.fine_tmp = nullptr;
  for (int c = 0; c < 4; ++c) ma.n_dev[c] = nullptr;
  for (int c = 0; c < 4; ++c) ma.src[c] = A.stage;
  ma.stage_off[0] = 0;
  for (int c = 1; c <= 4; ++c) ma.stage_off[c] = (unsigned)n;
  FeBuildArgs fa;
  fa.stage = A.stage; fa.n = (unsigned)n; fa.blob = A.blob; fa.slot_of = A.scratch; fa.rank_of = A.scratch + n;
  const unsigned tb = 256, gb = (unsigned)((n + tb - 1) / tb);
  unsigned tslots = 0;
  for (int c = 0; c < 4; ++c) tslots += hd.tsize[c];
  TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_map_bbox<<<(gb < 592u ? gb : 592u), tb, 0, h->stream>>>(ma)));   // + origin
  TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_fe_insert<<<gb, tb, 0, h->stream>>>(fa)));
  TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_map_offsets<<<(tslots + tb - 1) / tb, tb, 0, h->stream>>>(ma)));
  TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_fe_scatter<<<gb, tb, 0, h->stream>>>(fa)));
  FeGrid g;
  g.pts = reinterpret_cast<const FePoint*>(A.blob + hd.pts_off[0]);
  g.table = reinterpret_cast<const uint4*>(A.blob + hd.table_off[0]);
  g.mask = hd.tsize[0] - 1u; g.n = (unsigned)n; g.cell = cfg->radius; g.inv_cell = 1.0 / cfg->radius;
  g.origin = reinterpret_cast<const double*>(A.blob + offsetof(MapHeader, origin));
  FeParams prm;
  prm.dbg = h->profiling ? h->d_dbg + 8 : nullptr;     // slots 8..11 (k_correspond uses them in frame profiling)
  prm.r2 = cfg->radius * cfg->radius; prm.K = cfg->K; prm.min_neigh = cfg->min_neigh;
  prm.cvr_submap = cfg->cvr_submap; prm.planar_submap_thres = cfg->planar_submap_thres;
  prm.planar_vertic_thres = cfg->planar_vertic_thres;
  if (!h->fe_attr_set) {
    CU_TRY(cudaFuncSetAttribute(k_fe_pca, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFeSmemBytes));
    h->fe_attr_set = true;
  }
  TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_fe_pca<<<(unsigned)((n + kFeBlk - 1) / kFeBlk), kFeBlk, kFeSmemBytes, h->stream>>>(g, prm, A.out)));
  if (select) {
    TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_fe_classify<<<gb, tb, 0, h->stream>>>((unsigned)n, prm, A.out, key_p_raw, key_s_raw, val_p_raw, val_s_raw,
                                                                              A.counts)));
    // candidates first compacted, then ordered by (flatness descending, point index ascending): rank sort over the whole
    // GPU for lists up to 32 768 candidates, the shared-memory bitonic network for longer ones (each is a no-op otherwise)
    CU_TRY(cudaMemsetAsync(rank_p, 0, npad * 8, h->stream));
    TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_fe_rank<<<dim3(gb, 2, kFeRankSplit), 256, 0, h->stream>>>(key_p_raw, val_p_raw, key_s_raw, val_s_raw, rank_p,
                                                                                                  rank_s, A.counts)));
    TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_fe_rank_scatter<<<dim3(gb, 2), 256, 0, h->stream>>>(key_p_raw, val_p_raw, key_s_raw, val_s_raw, rank_p, rank_s,
                                                                                            A.key_p_sorted, A.val_p_sorted, A.key_s_sorted,
                                                                                            A.val_s_sorted, A.counts)));
    if (!h->fe_sort_attr_set) {
      CU_TRY(cudaFuncSetAttribute(k_fe_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFeSortSmemBytes));
      h->fe_sort_attr_set = true;
    }
    TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_fe_sort<<<2, 1024, kFeSortSmemBytes, h->stream>>>(A.key_p_sorted, A.val_p_sorted, A.key_s_sorted, A.val_s_sorted, A.counts)));
    TL_LAUNCH(TLOAM_B200_K_FEATURE, (k_fe_counts<<<1, 32, 0, h->stream>>>(A.key_p_sorted, A.key_s_sorted, A.counts, cfg->planar_num,
                                                                           cfg->sphere_num, cfg->planar_scan_thres, cfg->cvr_scan)));
  }
  CU_TRY(cudaGetLastError());
  // counts + build flags -> pinned scratch (h_result[28..31])
  CU_TRY(cudaMemcpyAsync(h->h_result + 28, A.counts, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaMemcpyAsync(h->h_result + 30, A.blob + offsetof(MapHeader, build_flags), sizeof(unsigned long long),
                         cudaMemcpyDeviceToHost, h->stream));
  if (select && A.host_val_p) CU_TRY(cudaMemcpyAsync(A.host_val_p, A.val_p_sorted, n * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  if (select && A.host_val_s) CU_TRY(cudaMemcpyAsync(A.host_val_s, A.val_s_sorted, n * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  unsigned long long flags;
  memcpy(&flags, h->h_result + 30, sizeof(flags));
  return (flags & 1ull) ? TLOAM_B200_ERR_MAP_DENSITY : TLOAM_B200_OK;
}

int tloam_b200_extract_planar_sphere(tloam_b200_handle* h, const tloam_feature_config* cfg, const double* xyz, size_t n,
                                     size_t* planar_scan_index, size_t* n_planar_scan, size_t* planar_submap_index,
                                     size_t* n_planar_submap, size_t* sphere_scan_index, size_t* n_sphere_scan,
                                     size_t* sphere_submap_index, size_t* n_sphere_submap, size_t* sphere_candidates) {
  if (!h || !cfg || !planar_scan_index || !n_planar_scan || !planar_submap_index || !n_planar_submap || !sphere_scan_index ||
      !n_sphere_scan || !sphere_submap_index || !n_sphere_submap)
    return TLOAM_B200_ERR_INVALID_ARG;
  *n_planar_scan = *n_planar_submap = *n_sphere_scan = *n_sphere_submap = 0;
  if (n == 0) return TLOAM_B200_OK;              // calculatePCAInfo fails on an empty cloud: nothing is selected (:49-54, :141)
  if (!xyz) return TLOAM_B200_ERR_INVALID_ARG;
  FeArena A;
  std::vector<unsigned> vp(n), vs(n);            // sorted point indices (candidates first), fetched with the counts
  A.host_val_p = vp.data(); A.host_val_s = vs.data();
  const int rc = fe_run(h, cfg, xyz, n, true, A);
  if (rc != TLOAM_B200_OK) return rc;
  unsigned counts[4];
  memcpy(counts, h->h_result + 28, sizeof(counts));
  const unsigned np = counts[0], ns = counts[1], nps = counts[2], nss = counts[3];
  for (unsigned i = 0; i < np; ++i) planar_submap_index[i] = vp[i];                    // :180
  for (unsigned i = 0; i < nps; ++i) planar_scan_index[i] = vp[i];                     // :177-178
  *n_planar_submap = np; *n_planar_scan = nps;
  for (unsigned i = 0; i < ns; ++i) sphere_submap_index[i] = i;                          // :187 (rank, not index)
  for (unsigned i = 0; i < nss; ++i) sphere_scan_index[i] = i;                           // :184-185
  *n_sphere_submap = ns; *n_sphere_scan = nss;
  if (sphere_candidates) for (unsigned i = 0; i < ns; ++i) sphere_candidates[i] = vs[i];
  return TLOAM_B200_OK;
}

int tloam_b200_pca_info(tloam_b200_handle* h, const tloam_feature_config* cfg, const double* xyz, size_t n, double* cvr,
                        double* flatness, double* sphericity, double* normal, int* num_sum, int* neigh) {
  if (!h || !cfg || !xyz || n == 0) return TLOAM_B200_ERR_INVALID_ARG;
  FeArena A;
  const int rc = fe_run(h, cfg, xyz, n, false, A);
  if (rc != TLOAM_B200_OK) return rc;
  if (cvr) CU_TRY(cudaMemcpyAsync(cvr, A.out.cvr, n * 8, cudaMemcpyDeviceToHost, h->stream));
  if (flatness) CU_TRY(cudaMemcpyAsync(flatness, A.out.flatness, n * 8, cudaMemcpyDeviceToHost, h->stream));
  if (sphericity) CU_TRY(cudaMemcpyAsync(sphericity, A.out.sphericity, n * 8, cudaMemcpyDeviceToHost, h->stream));
  if (normal) CU_TRY(cudaMemcpyAsync(normal, A.out.normal, n * 24, cudaMemcpyDeviceToHost, h->stream));
  if (num_sum) CU_TRY(cudaMemcpyAsync(num_sum, A.out.num_sum, n * 4, cudaMemcpyDeviceToHost, h->stream));
  std::vector<int> nb;
  if (neigh) {
    nb.resize(n * kFeK);
    CU_TRY(cudaMemcpyAsync(nb.data(), A.out.neigh, n * kFeK * 4, cudaMemcpyDeviceToHost, h->stream));
  }
  CU_TRY(cudaStreamSynchronize(h->stream));
  if (neigh)
    for (size_t i = 0; i < n; ++i)
      for (int j = 0; j < cfg->K; ++j) neigh[i * cfg->K + j] = nb[i * kFeK + j];
  return TLOAM_B200_OK;
}

int tloam_b200_voxel_down_sample(tloam_b200_handle* h, const double* pts, size_t n, double voxel, double* out, size_t* n_out) {
  if (!h || (!pts && n) || !out || !n_out) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  int rc = upload_points(h, pts, n);
  if (rc != TLOAM_B200_OK) return rc;
  rc = ensure_dev(h, &h->d_acc_tmp, &h->cap_acc_tmp, n, false);
  if (rc != TLOAM_B200_OK) return rc;
  rc = voxel_pipeline(h, h->d_up, n, nullptr, 0u, nullptr, nullptr, nullptr, 0.0, voxel, h->d_acc_tmp, h->d_cnt + 8);
  if (rc != TLOAM_B200_OK) return rc;
  CU_TRY(cudaMemcpyAsync(h->h_result + 28, h->d_cnt + 8, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  unsigned cnt;
  memcpy(&cnt, h->h_result + 28, sizeof(cnt));
  *n_out = cnt;
  if (cnt) CU_TRY(cudaMemcpyAsync(out, h->d_acc_tmp, (size_t)cnt * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return TLOAM_B200_OK;
}

int tloam_b200_submap_init(tloam_b200_handle* h, const tloam_submap_config* cfg, const double* edge, size_t ne,
                           const double* ground_raw, size_t ng, const double* planar_sub, size_t np,
                           const double* sphere_sub, size_t ns) {
  if (!h || !cfg || (!edge && ne) || (!ground_raw && ng) || (!planar_sub && np) || (!sphere_sub && ns)) return TLOAM_B200_ERR_INVALID_ARG;
  if (cfg->planar_frame_size < 1 || cfg->planar_frame_size > 64) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  h->scfg = *cfg;
  if (!h->d_pose) CU_TRY(cudaMalloc(&h->d_pose, 16 * sizeof(double)));
  int rc;
  h->acc_cur[0] = h->acc_cur[1] = 0;
  // edge: raw copy (front_end.cpp:286)
  if ((rc = ensure_dev(h, &h->d_acc[0], &h->cap_acc[0], ne, false)) != TLOAM_B200_OK) return rc;
  if (ne) CU_TRY(cudaMemcpyAsync(h->d_acc[0], edge, ne * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->n_acc[0] = ne;
  k_set_counts<<<1, 32, 0, h->stream>>>(h->d_cnt + 4, 0, (unsigned)ne, -1, 0u);
  // ground: VoxelDownSample(ground_down_sample) (:287)
  if ((rc = upload_points(h, ground_raw, ng)) != TLOAM_B200_OK) return rc;
  if ((rc = ensure_dev(h, &h->d_acc[1], &h->cap_acc[1], ng, false)) != TLOAM_B200_OK) return rc;
  if ((rc = voxel_pipeline(h, h->d_up, ng, nullptr, 0u, nullptr, nullptr, nullptr, 0.0, cfg->ground_down_sample, h->d_acc[1],
                           acc_count(h, 1))) != TLOAM_B200_OK) return rc;
  // planar / sphere: the submap-index selections (:291-292); the sliding-window buffers stay empty (:285-305)
  if ((rc = ensure_dev(h, &h->d_cat, &h->cap_cat, np, false)) != TLOAM_B200_OK) return rc;
  if (np) CU_TRY(cudaMemcpyAsync(h->d_cat, planar_sub, np * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->n_cat = np;
  cudaFree(h->d_sphere0); h->d_sphere0 = nullptr;
  if (ns) {
    CU_TRY(cudaMalloc(&h->d_sphere0, ns * 3 * sizeof(double)));
    CU_TRY(cudaMemcpyAsync(h->d_sphere0, sphere_sub, ns * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  h->n_sphere0 = ns; h->sphere_is_init = true;
  for (double* p : h->ring) cudaFree(p);
  h->ring.clear(); h->ring_n.clear(); h->ring_cap.clear();
  // once per sequence: wait for the uploads and read the exact ground count
  CU_TRY(cudaMemcpyAsync(h->h_result + 28, acc_count(h, 1), sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  unsigned gcnt;
  memcpy(&gcnt, h->h_result + 28, sizeof(gcnt));
  h->n_acc[1] = gcnt;
  for (int k = 0; k < 2; ++k) { h->cum_add[k] = h->known_cum[k] = 0; h->known_cnt[k] = h->n_acc[k]; }
  for (auto& pr : h->probes) pr.pending = false;
  h->submap_ready = true;
  return submap_set_target(h);
}

// FrontEnd::updateSubmap (ref: front_end.cpp:201-267) enqueued WITHOUT a host round trip: the voxel counts that size the
// next map stay on the device (the host only tracks upper bounds, tightened by asynchronous read-backs), and in the
// chained form the pose is the device-resident result of the frame that was just enqueued.
static int submap_update_impl(tloam_b200_handle* h, const double* pose_host, const double* planar_sub, size_t np) {
  if (!h || (!planar_sub && np)) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->submap_ready || !h->have_src) return TLOAM_B200_ERR_NOT_READY;
  if (!h->src_staged) return TLOAM_B200_ERR_NOT_READY;           // the staged source is what gets appended
  CU_TRY(cudaSetDevice(h->device));
  harvest_counts(h);
  const tloam_submap_config& cf = h->scfg;
  int rc;
  const double* d_pose = h->d_pose;
  if (pose_host) {
    double tmp[16];
    memcpy(tmp, pose_host, sizeof(tmp));
    CU_TRY(cudaMemcpyAsync(h->d_pose, tmp, sizeof(tmp), cudaMemcpyHostToDevice, h->stream));   // pageable source: staged before return
  } else {
    d_pose = reinterpret_cast<const double*>(reinterpret_cast<const char*>(h->d_state) + offsetof(FrameState, result));
  }
  const unsigned tb = 256;
  // ---- planar sliding window (:207-217, 232-242): newest frame transformed by its pose ----
  // pipelined handles: the upload runs on src_stream (it does not depend on the frame that is still being registered);
  // its buffer is free again once the transform below has read it
  static const bool no_prefetch = getenv("TLOAM_B200_NO_PREFETCH") != nullptr;
  const bool side = h->async_inputs && !h->profiling && !no_prefetch;
  const double* d_planar_in = nullptr;
  if (side) {
    if (np > h->cap_up_planar) {
      CU_TRY(cudaStreamSynchronize(h->stream));
      cudaFree(h->d_up_planar); h->d_up_planar = nullptr; h->cap_up_planar = 0;
      CU_TRY(cudaMalloc(&h->d_up_planar, (np + np / 2 + 1024) * 3 * sizeof(double)));
      h->cap_up_planar = np + np / 2 + 1024;
      h->planar_free_valid = false;
    }
    if (h->planar_free_valid) CU_TRY(cudaStreamWaitEvent(h->src_stream, h->ev_planar_free, 0));
    if (np) CU_TRY(cudaMemcpyAsync(h->d_up_planar, planar_sub, np * 3 * sizeof(double), cudaMemcpyHostToDevice, h->src_stream));
    CU_TRY(cudaEventRecord(h->ev_planar_in, h->src_stream));
    d_planar_in = h->d_up_planar;
  } else {
    if ((rc = upload_points(h, planar_sub, np)) != TLOAM_B200_OK) return rc;
    d_planar_in = h->d_up;
  }
  // ---- fork: the ground accumulator (k = 1 below) is appended, cropped and down-sampled on sub_stream, the planar
  //      window is assembled on fit_stream (idle between frames), the edge accumulator stays on the handle's stream ----
  cudaStream_t gs = side ? h->sub_stream : h->stream;
  cudaStream_t ps = side ? h->fit_stream : h->stream;
  if (side) {
    CU_TRY(cudaEventRecord(h->ev_sub[0], h->stream));
    CU_TRY(cudaStreamWaitEvent(gs, h->ev_sub[0], 0));
    CU_TRY(cudaStreamWaitEvent(ps, h->ev_sub[0], 0));
    CU_TRY(cudaStreamWaitEvent(ps, h->ev_planar_in, 0));
  }
  double* slot = nullptr; size_t slot_cap = 0;
  if ((int)h->ring.size() >= cf.planar_frame_size) {            // recycle the oldest buffer
    slot = h->ring.front(); slot_cap = h->ring_cap.front();
    h->ring.erase(h->ring.begin()); h->ring_n.erase(h->ring_n.begin()); h->ring_cap.erase(h->ring_cap.begin());
  }
  if ((rc = ensure_dev(h, &slot, &slot_cap, np, false)) != TLOAM_B200_OK) return rc;
  if (np) TL_LAUNCH(TLOAM_B200_K_SUBMAP, (k_transform_append<<<(unsigned)((np + tb - 1) / tb), tb, 0, ps>>>(d_planar_in, (unsigned)np, slot, d_pose, nullptr)));
  if (side) { CU_TRY(cudaEventRecord(h->ev_planar_free, ps)); h->planar_free_valid = true; }
  h->ring.push_back(slot); h->ring_n.push_back(np); h->ring_cap.push_back(slot_cap);
  size_t tot = 0;
  for (size_t k : h->ring_n) tot += k;
  if ((rc = ensure_dev(h, &h->d_cat, &h->cap_cat, tot, false)) != TLOAM_B200_OK) return rc;
  size_t off = 0;
  for (size_t f = 0; f < h->ring.size(); ++f) {
    if (h->ring_n[f]) CU_TRY(cudaMemcpyAsync(h->d_cat + 3 * off, h->ring[f], h->ring_n[f] * 3 * sizeof(double), cudaMemcpyDeviceToDevice, ps));
    off += h->ring_n[f];
  }
  if (side) CU_TRY(cudaEventRecord(h->ev_planar_done, ps));     // the planar window is assembled
  h->n_cat = tot;
  h->sphere_is_init = false;
  // ---- edge / ground: append the current source features in the world frame (:245-246), crop (:248-264),
  //      VoxelDownSample ----
  const int src_cloud[2] = {0, 3};
  const double vox[2] = {cf.edge_down_sample_submap, cf.ground_down_sample_submap};
  const double len[2] = {cf.edge_crop_box_length, cf.ground_crop_box_length};
  size_t soff[4], o = 0;
  for (int c = 0; c < 4; ++c) { soff[c] = o; o += h->n_src[c]; }
  for (int k = 0; k < 2; ++k) {
    const size_t nadd = h->n_src[src_cloud[k]];
    const size_t nall = h->n_acc[k] + nadd;                       // bound
    cudaStream_t ks = k == 1 ? gs : h->stream;                    // ground on the side stream, own scratch + output buffer
    double*& tmp = (k == 1 && side) ? h->d_acc_tmp1 : h->d_acc_tmp;
    size_t& tmp_cap = (k == 1 && side) ? h->cap_acc_tmp1 : h->cap_acc_tmp;
    if ((rc = ensure_dev(h, &h->d_acc[k], &h->cap_acc[k], nall, true)) != TLOAM_B200_OK) return rc;
    if ((rc = ensure_dev(h, &tmp, &tmp_cap, nall, false)) != TLOAM_B200_OK) return rc;
    unsigned* cur = acc_count(h, k);
    if (nadd) TL_LAUNCH(TLOAM_B200_K_SUBMAP, (k_transform_append<<<(unsigned)((nadd + tb - 1) / tb), tb, 0, ks>>>(
        h->d_stage_src + 3 * soff[src_cloud[k]], (unsigned)nadd, h->d_acc[k], d_pose, cur)));
    h->acc_cur[k] ^= 1;
    if ((rc = voxel_pipeline(h, h->d_acc[k], nall, cur, (unsigned)nadd, nullptr, nullptr, d_pose, len[k], vox[k], tmp,
                             acc_count(h, k), ks, (k == 1 && side) ? 1 : 0)) != TLOAM_B200_OK) return rc;
    // the down-sampled cloud becomes the accumulator (swap buffers)
    double* t = h->d_acc[k]; h->d_acc[k] = tmp; tmp = t;
    size_t tc = h->cap_acc[k]; h->cap_acc[k] = tmp_cap; tmp_cap = tc;
    h->n_acc[k] = nall;
    h->cum_add[k] += nadd;
  }
  if (side) {                                                     // join
    CU_TRY(cudaEventRecord(h->ev_sub[1], gs));
    CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_sub[1], 0));
    CU_TRY(cudaStreamWaitEvent(h->stream, h->ev_planar_done, 0));
  }
  CU_TRY(cudaEventRecord(h->ev_stage_free[h->stage_cur], h->stream));   // the staged source has been appended: its buffer is free
  // asynchronous read-back of the two exact counts (tightens the bounds of later frames)
  {
    tloam_b200_handle::CntProbe& pr = h->probes[h->probe_next];
    if (!pr.pending || cudaEventQuery(pr.ev) == cudaSuccess) {
      if (pr.pending) { pr.pending = false; for (int k = 0; k < 2; ++k) if (pr.cum[k] >= h->known_cum[k]) { h->known_cum[k] = pr.cum[k]; h->known_cnt[k] = pr.h_vals[k]; } }
      for (int k = 0; k < 2; ++k) {
        CU_TRY(cudaMemcpyAsync(pr.h_vals + k, acc_count(h, k), sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
        pr.cum[k] = h->cum_add[k];
      }
      CU_TRY(cudaEventRecord(pr.ev, h->stream));
      pr.pending = true;
      h->probe_next = (h->probe_next + 1) & 3;
    } else {
      cudaGetLastError();
    }
  }
  return submap_set_target(h);                                   // :267
}

int tloam_b200_submap_update(tloam_b200_handle* h, const double pose[16], const double* planar_sub, size_t np,
                             const double* sphere_sub, size_t ns) {
  (void)sphere_sub; (void)ns;   // stored but never used by the reference (front_end.cpp:202-205, 220-230)
  if (!pose) return TLOAM_B200_ERR_INVALID_ARG;
  return submap_update_impl(h, pose, planar_sub, np);
}

// pose = the result of the frame that was just enqueued on this handle, read on the device: frames chain with no host
// round trip (set_source -> scan_match_predicted_async -> submap_update_chained -> next frame)
int tloam_b200_submap_update_chained(tloam_b200_handle* h, const double* planar_sub, size_t np) {
  return submap_update_impl(h, nullptr, planar_sub, np);
}

// exact sizes (synchronises: inspection / tests)
int tloam_b200_submap_sizes(tloam_b200_handle* h, size_t n[4]) {
  if (!h || !n) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->submap_ready) return TLOAM_B200_ERR_NOT_READY;
  CU_TRY(cudaSetDevice(h->device));
  unsigned cnt[2];
  for (int k = 0; k < 2; ++k) CU_TRY(cudaMemcpyAsync(h->h_result + 28 + k, acc_count(h, k), sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  for (int k = 0; k < 2; ++k) memcpy(&cnt[k], h->h_result + 28 + k, sizeof(unsigned));
  n[0] = cnt[0]; n[1] = h->sphere_is_init ? h->n_sphere0 : h->n_cat; n[2] = h->n_cat; n[3] = cnt[1];
  return TLOAM_B200_OK;
}

int tloam_b200_submap_download(tloam_b200_handle* h, int cloud, double* out, size_t capacity_points) {
  if (!h || !out || cloud < 0 || cloud > 3) return TLOAM_B200_ERR_INVALID_ARG;
  if (!h->submap_ready) return TLOAM_B200_ERR_NOT_READY;
  size_t n[4];
  const int rc = tloam_b200_submap_sizes(h, n);
  if (rc != TLOAM_B200_OK) return rc;
  if (capacity_points < n[cloud]) return TLOAM_B200_ERR_INVALID_ARG;
  const double* src = cloud == 0 ? h->d_acc[0] : cloud == 3 ? h->d_acc[1] : (cloud == 1 && h->sphere_is_init) ? h->d_sphere0 : h->d_cat;
  CU_TRY(cudaSetDevice(h->device));
  if (n[cloud]) CU_TRY(cudaMemcpyAsync(out, src, n[cloud] * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return TLOAM_B200_OK;
}


// ---------------------------------------------------------------------------------------------
// "next" row (f)-4, first part: multi-region ground extraction (ground_extract.cuh)
// ---------------------------------------------------------------------------------------------
void tloam_b200_ground_default_config(tloam_ground_config* c) {   // ref: config/mapping/segmentation.yaml
  c->sensor_model = 64; c->sensor_height = 1.73; c->vertical_res = 0.4; c->init_angle = -24.9;
  c->sensor_min_range = 1.0; c->sensor_max_range = 120.0;
  c->quadrant = 4; c->num_sec = 3; c->plane_dis = 0.3; c->max_iter = 3; c->ground_seed_num = 20;
}

// Segmentation::initSections, ref: segmentation.cpp:174-221 (host side: 64 iterations of scalar arithmetic).  Restated
// literally: the `continue` at :203-206 also skips the angle increment, so the table stalls at the first >= 5 m jump and
// only two section bounds are produced with the shipped configuration.
static int ground_section_bounds(const tloam_ground_config& c, float out[4]) {
  int nb = 0;
  int boundary[4] = {0, 0, 0, 0};
  const int section_width = static_cast<int>(std::ceil(1.0 * c.sensor_model) / c.num_sec);
  for (int i = 0; i < c.num_sec; ++i) boundary[i] = section_width * (i + 1) - 1;
  double prev_radius = 0.0, angle = c.init_angle;
  int sec = 0;
  for (int i = 0; i < c.sensor_model; ++i) {
    if (c.sensor_model == 64 && i == 31) angle += 1.7;
    double cur = c.sensor_height / std::tan(std::fabs(angle) / 180.0 * M_PI);
    cur = cur < c.sensor_max_range ? cur : c.sensor_max_range;
    if (i >= 1) {
      const double dis = std::fabs(cur - prev_radius);
      if (dis >= 5.0 || dis <= 0.0) continue;
    }
    if (sec < c.num_sec && i == boundary[sec] && sec <= 3) {
      const double theta = std::fabs(angle / 180 * M_PI);
      out[nb++] = (theta != 0 && i < c.sensor_model) ? static_cast<float>(c.sensor_height / std::tan(theta))
                                                      : static_cast<float>(c.sensor_max_range);
      ++sec;
    }
    prev_radius = cur;
    angle += c.vertical_res;
  }
  return nb;
}

int tloam_b200_ground_extract(tloam_b200_handle* h, const tloam_ground_config* cfg, const double* xyz, size_t n,
                              size_t* ground_index, size_t* n_ground, size_t* object_index, size_t* n_object, int* beam,
                              int* region, double* height_threshold, double* planes) {
  if (!h || !cfg || !ground_index || !n_ground || !object_index || !n_object) return TLOAM_B200_ERR_INVALID_ARG;
  *n_ground = *n_object = 0;
  if (cfg->sensor_model != 64 || cfg->quadrant != 4 || cfg->num_sec < 1 || cfg->num_sec > 3 || cfg->max_iter < 1 ||
      cfg->max_iter > kGeMaxIter || cfg->ground_seed_num < 1)
    return TLOAM_B200_ERR_INVALID_ARG;                            // only the HDL-64E branch is built (:439-441)
  if (planes) for (int i = 0; i < 12 * kGeMaxIter * 4; ++i) planes[i] = std::nan("");
  if (n == 0) { if (height_threshold) *height_threshold = 1.0; return TLOAM_B200_OK; }   // :335-338
  if (!xyz || n > ((size_t)1 << 30)) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  GeArgs a;
  memset(&a, 0, sizeof(a));
  a.n = (unsigned)n; a.nchunk = (unsigned)((n + kGeChunk - 1) / kGeChunk);
  a.sensor_model = cfg->sensor_model; a.num_sec = cfg->num_sec; a.max_iter = cfg->max_iter; a.seed_num = cfg->ground_seed_num;
  a.sensor_height = cfg->sensor_height; a.min_range = cfg->sensor_min_range; a.max_range = cfg->sensor_max_range;
  a.plane_dis = cfg->plane_dis;
  a.nbounds = ground_section_bounds(*cfg, a.bounds);
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += round_up(bytes, 256); return o; };
  const size_t o_pts = take(n * 24), o_ct = take(a.nchunk * 4), o_cs = take(a.nchunk * 8), o_sc = take(64), o_beam = take(n * 4),
               o_key = take(n), o_cc = take((size_t)a.nchunk * kGeKeys * 4), o_kb = take((kGeKeys + 1) * 4), o_ord = take(n * 4),
               o_flag = take(n), o_lists = take(2 * n * 4), o_rc = take(12 * 2 * 4), o_pl = take(12 * kGeMaxIter * 4 * 8),
               o_og = take(n * 4), o_oo = take(n * 4), o_oc = take(64);
  if (off > h->cap_ge) {
    CU_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_ge); h->d_ge = nullptr; h->cap_ge = 0;
    CU_TRY(cudaMalloc(&h->d_ge, off + off / 4));
    h->cap_ge = off + off / 4;
  }
  unsigned char* b = h->d_ge;
  a.pts = (const double*)(b + o_pts); a.chunk_trans = (unsigned*)(b + o_ct); a.chunk_sum = (double*)(b + o_cs);
  a.scal = (double*)(b + o_sc); a.beam = (int*)(b + o_beam); a.key = b + o_key; a.chunk_cnt = (unsigned*)(b + o_cc);
  a.key_base = (unsigned*)(b + o_kb); a.order = (unsigned*)(b + o_ord); a.flag = b + o_flag; a.lists = (unsigned*)(b + o_lists);
  a.reg_cnt = (unsigned*)(b + o_rc); a.planes = (double*)(b + o_pl); a.out_ground = (unsigned*)(b + o_og);
  a.out_object = (unsigned*)(b + o_oo); a.out_counts = (unsigned*)(b + o_oc);
  if (h->seg.active) a.pts = h->seg.dev_xyz;
  else CU_TRY(cudaMemcpyAsync(b + o_pts, xyz, n * 24, cudaMemcpyHostToDevice, h->stream));
  TL_LAUNCH(TLOAM_B200_K_GROUND, (k_ge_pre<<<a.nchunk, kGeChunk, 0, h->stream>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_GROUND, (k_ge_scan1<<<1, 1024, 0, h->stream>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_GROUND, (k_ge_region<<<a.nchunk, kGeChunk, 0, h->stream>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_GROUND, (k_ge_scan2<<<1, 1024, 0, h->stream>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_GROUND, (k_ge_scatter<<<a.nchunk, kGeChunk, 0, h->stream>>>(a)));
  constexpr size_t kGeFitSmem = (6 * kGeTile + kGeSeedCache) * sizeof(double);
  if (!h->ge_attr_set) {
    CU_TRY(cudaFuncSetAttribute(k_ge_fit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGeFitSmem));
    h->ge_attr_set = true;
  }
  TL_LAUNCH(TLOAM_B200_K_GROUND, (k_ge_fit<<<4 * cfg->num_sec, kGeFitThreads, kGeFitSmem, h->stream>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_GROUND, (k_ge_emit<<<dim3(32, 25), 256, 0, h->stream>>>(a)));
  CU_TRY(cudaGetLastError());
  // counts + threshold first (one small copy each, one synchronisation), then exactly the list prefixes
  CU_TRY(cudaMemcpyAsync(h->h_result + 28, a.out_counts, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaMemcpyAsync(h->h_result + 29, a.scal, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  unsigned counts[2];
  memcpy(counts, h->h_result + 28, sizeof(counts));
  if (height_threshold) *height_threshold = h->h_result[29];
  if (h->seg.active) {                                                     // chained: the lists stay on the device
    h->seg.ground = a.out_ground; h->seg.object = a.out_object; h->seg.beam = a.beam;
    h->seg.n_ground = counts[0]; h->seg.n_object = counts[1];
    *n_ground = counts[0]; *n_object = counts[1];
    return TLOAM_B200_OK;
  }
  std::vector<unsigned> gi(counts[0]), oi(counts[1]);
  if (counts[0]) CU_TRY(cudaMemcpyAsync(gi.data(), a.out_ground, counts[0] * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  if (counts[1]) CU_TRY(cudaMemcpyAsync(oi.data(), a.out_object, counts[1] * sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
  if (beam) CU_TRY(cudaMemcpyAsync(beam, a.beam, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  std::vector<unsigned char> keys;
  if (region) { keys.resize(n); CU_TRY(cudaMemcpyAsync(keys.data(), a.key, n, cudaMemcpyDeviceToHost, h->stream)); }
  if (planes) CU_TRY(cudaMemcpyAsync(planes, a.planes, 12 * kGeMaxIter * 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  for (unsigned i = 0; i < counts[0]; ++i) ground_index[i] = gi[i];
  for (unsigned i = 0; i < counts[1]; ++i) object_index[i] = oi[i];
  if (region) for (size_t i = 0; i < n; ++i) region[i] = keys[i];
  *n_ground = counts[0]; *n_object = counts[1];
  return TLOAM_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// "next" row (f)-4, second part: LOAM-style edge extraction (edge_extract.cuh)
// ---------------------------------------------------------------------------------------------
int tloam_b200_extract_edge(tloam_b200_handle* h, int sensor_model, int ring_min_num, const double* xyz, const double* intensity,
                            size_t n, size_t* edge_index, size_t* n_edge, size_t* non_edge_index, size_t* n_non_edge) {
  if (!h || !edge_index || !n_edge || !non_edge_index || !n_non_edge) return TLOAM_B200_ERR_INVALID_ARG;
  *n_edge = *n_non_edge = 0;
  if (sensor_model < 1 || sensor_model > kEeKeys || ring_min_num < 0) return TLOAM_B200_ERR_INVALID_ARG;
  if (n == 0) return TLOAM_B200_OK;                                        // ref: :1222-1225 (empty input: nothing extracted)
  if (!xyz || !intensity || n > ((size_t)1 << 30)) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  EeArgs a;
  memset(&a, 0, sizeof(a));
  a.n = (unsigned)n; a.nchunk = (unsigned)((n + kEeChunk - 1) / kEeChunk);
  a.sensor_model = sensor_model; a.ring_min = ring_min_num;
  const int nsec = kEeKeys * kEeSectors;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += round_up(bytes, 256); return o; };
  const size_t o_pts = take(n * 24), o_int = take(n * 8), o_key = take(n), o_cc = take((size_t)a.nchunk * kEeKeys * 4),
               o_rb = take((kEeKeys + 1) * 4), o_ord = take(n * 4), o_se = take(n * 4), o_sn = take(n * 4), o_sc = take(nsec * 2 * 4),
               o_so = take((nsec + 1) * 2 * 4), o_oe = take(n * 8), o_on = take(n * 8), o_st = take(64);
  if (off > h->cap_ge) {                                                   // shares the segmentation arena
    CU_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_ge); h->d_ge = nullptr; h->cap_ge = 0;
    CU_TRY(cudaMalloc(&h->d_ge, off + off / 4));
    h->cap_ge = off + off / 4;
  }
  unsigned char* b = h->d_ge;
  a.pts = (const double*)(b + o_pts); a.intensity = (const double*)(b + o_int); a.key = b + o_key;
  a.chunk_cnt = (unsigned*)(b + o_cc); a.ring_base = (unsigned*)(b + o_rb); a.order = (unsigned*)(b + o_ord);
  a.sec_edge = (unsigned*)(b + o_se); a.sec_non = (unsigned*)(b + o_sn); a.sec_cnt = (unsigned*)(b + o_sc);
  a.sec_off = (unsigned*)(b + o_so); a.out_edge = (unsigned long long*)(b + o_oe); a.out_non = (unsigned long long*)(b + o_on);
  a.status = (int*)(b + o_st);
  if (!h->ee_attr_set) {
    CU_TRY(cudaFuncSetAttribute(k_ee_section, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEeSmemBytes));
    h->ee_attr_set = true;
  }
  if (h->seg.active) { a.pts = h->seg.dev_xyz; a.intensity = h->seg.dev_intensity; }
  else {
    CU_TRY(cudaMemcpyAsync(b + o_pts, xyz, n * 24, cudaMemcpyHostToDevice, h->stream));
    CU_TRY(cudaMemcpyAsync(b + o_int, intensity, n * 8, cudaMemcpyHostToDevice, h->stream));
  }
  CU_TRY(cudaMemsetAsync(b + o_sc, 0, nsec * 2 * 4, h->stream));         // beams >= sensor_model have no sections
  TL_LAUNCH(TLOAM_B200_K_EDGE, (k_ee_key<<<a.nchunk, kEeChunk, 0, h->stream>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_EDGE, (k_ee_scan<<<1, kEeKeys, 0, h->stream>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_EDGE, (k_ee_scatter<<<a.nchunk, kEeChunk, 0, h->stream>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_EDGE, (k_ee_section<<<dim3(kEeSectors, sensor_model), kEeThreads, kEeSmemBytes, h->stream>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_EDGE, (k_ee_offsets<<<1, 32, 0, h->stream>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_EDGE, (k_ee_copy<<<dim3(kEeSectors, sensor_model), kEeThreads, 0, h->stream>>>(a)));
  CU_TRY(cudaGetLastError());
  unsigned tot[2];
  int st = 0;
  CU_TRY(cudaMemcpyAsync(tot, a.sec_off + 2 * (sensor_model * kEeSectors), sizeof(tot), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaMemcpyAsync(&st, a.status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  if (st != 0) {
    snprintf(h->last_error, sizeof(h->last_error), "extract_edge: a beam sector holds more than %d curvature values", kEeMaxSection);
    return TLOAM_B200_ERR_INVALID_ARG;
  }
  static_assert(sizeof(size_t) == sizeof(unsigned long long), "index lists are copied straight into size_t arrays");
  if (h->seg.active) {                                                     // chained: the lists stay on the device
    h->seg.edge = a.out_edge; h->seg.non_edge = a.out_non; h->seg.n_edge = tot[0]; h->seg.n_non = tot[1];
    *n_edge = tot[0]; *n_non_edge = tot[1];
    return TLOAM_B200_OK;
  }
  if (tot[0]) CU_TRY(cudaMemcpyAsync(edge_index, a.out_edge, tot[0] * sizeof(size_t), cudaMemcpyDeviceToHost, h->stream));
  if (tot[1]) CU_TRY(cudaMemcpyAsync(non_edge_index, a.out_non, tot[1] * sizeof(size_t), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  *n_edge = tot[0]; *n_non_edge = tot[1];
  return TLOAM_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// "next" row (f)-4, third part: object segmentation = DCVC (object_segment.cuh)
// ---------------------------------------------------------------------------------------------
void tloam_b200_dcvc_default_config(tloam_dcvc_config* c) {                  // ref: config/mapping/segmentation.yaml
  if (!c) return;
  c->start_r = 0.35; c->delta_r = 0.0004; c->delta_p = 1.2; c->delta_a = 1.2; c->min_seg = 80;
  c->sensor_min_range = 1.0; c->sensor_max_range = 120.0;
  c->min_pitch_init = 0.0; c->max_pitch_init = 0.0; c->min_polar_init = 0.0; c->max_polar_init = 0.0;
}

int tloam_b200_object_segmentation(tloam_b200_handle* h, const tloam_dcvc_config* cfg, const double* xyz, size_t n,
                                   size_t* seg_index, size_t* n_seg, int* n_clusters, int* sizes, double* boxes, int* root,
                                   int* cluster, int* voxel, double* polar) {
  if (!h || !cfg || !seg_index || !n_seg || !n_clusters) return TLOAM_B200_ERR_INVALID_ARG;
  *n_seg = 0; *n_clusters = 0;
  if (!(cfg->delta_p > 0.0) || !(cfg->delta_a > 0.0)) return TLOAM_B200_ERR_INVALID_ARG;
  if (n == 0) return TLOAM_B200_OK;                                         // ref: :1088-1093 (nothing to convert)
  if (!xyz || n > ((size_t)1 << 26)) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  OsArgs a;
  memset(&a, 0, sizeof(a));
  a.n = (unsigned)n; a.nchunk = (unsigned)((n + kOsChunk - 1) / kOsChunk);
  a.start_r = cfg->start_r; a.delta_r = cfg->delta_r; a.delta_p = cfg->delta_p; a.delta_a = cfg->delta_a;
  a.min_range = cfg->sensor_min_range; a.max_range = cfg->sensor_max_range; a.min_seg = cfg->min_seg;
  a.init[0] = cfg->min_pitch_init; a.init[1] = cfg->max_pitch_init; a.init[2] = cfg->min_polar_init; a.init[3] = cfg->max_polar_init;
  size_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  a.cap_mask = (unsigned)(cap - 1);
  const size_t n32 = round_up(n, 32);
  const size_t state_words = (n + 15) / 16 + 1;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += round_up(bytes, 256); return o; };
  const size_t o_pts = take(n * 24), o_pol = take(n * 24), o_enc = take(64), o_ext = take(64), o_par = take(64),
               o_bnd = take(kOsMaxBounds * 8), o_key = take(n * 4), o_crd = take(n * 12), o_tab = take(cap * 8), o_sv = take(cap * 4),
               o_ps = take(n * 4), o_vid = take(n * 4), o_f1 = take(n * 4), o_f2 = take(n * 4),
               o_ek = take(n * 4), o_ept = take(n * 4), o_ei = take(n32 * 4), o_rows = take((n32 + 32 * kOsSeqStages) * kOsRowStride * 4), o_ep = take(n32),
               o_stg = take(state_words * 4), o_par2 = take(n * 4), o_root = take(n * 4), o_cnt = take(n * 4), o_clr = take(n * 4),
               o_crr = take(n * 4), o_cl = take(n * 4), o_sz = take(n * 4), o_pk = take(n * 4), o_box = take(n * 48), o_benc = take(n * 48),
               o_cc = take((size_t)a.nchunk * kOsKeys * 4), o_kb = take((kOsKeys + 1) * 4), o_pa = take(n * 4), o_pb = take(n * 4),
               o_seg = take(n * 8);
  if (off > h->cap_ge) {                                                   // shares the segmentation arena
    CU_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_ge); h->d_ge = nullptr; h->cap_ge = 0;
    CU_TRY(cudaMalloc(&h->d_ge, off + off / 4));
    h->cap_ge = off + off / 4;
  }
  unsigned char* b = h->d_ge;
  a.pts = (const double*)(b + o_pts); a.polar = (double*)(b + o_pol); a.ext_enc = (unsigned long long*)(b + o_enc);
  a.ext = (double*)(b + o_ext); a.params = (int*)(b + o_par); a.bounds = (double*)(b + o_bnd); a.key = (int*)(b + o_key);
  a.coord = (int*)(b + o_crd); a.table = (unsigned long long*)(b + o_tab); a.slot_vid = (int*)(b + o_sv); a.pslot = (int*)(b + o_ps);
  a.vid = (int*)(b + o_vid); a.f1 = (int*)(b + o_f1); a.f2 = (int*)(b + o_f2); a.evkey = (int*)(b + o_ek);
  a.ev_pt = (int*)(b + o_ept); a.ev_info = (int*)(b + o_ei); a.rows = (int*)(b + o_rows); a.ev_p = (signed char*)(b + o_ep);
  a.state_g = (unsigned*)(b + o_stg); a.parent = (int*)(b + o_par2); a.root = (int*)(b + o_root); a.cnt = (int*)(b + o_cnt);
  a.cl_root = (int*)(b + o_clr); a.cl_rank_of_root = (int*)(b + o_crr); a.cluster = (int*)(b + o_cl); a.sizes = (int*)(b + o_sz);
  a.pkey = (int*)(b + o_pk); a.boxes = (double*)(b + o_box); a.box_enc = (unsigned long long*)(b + o_benc); a.chunk_cnt = (unsigned*)(b + o_cc); a.key_base = (unsigned*)(b + o_kb);
  a.perm_a = (int*)(b + o_pa); a.perm_b = (int*)(b + o_pb); a.out_seg = (unsigned long long*)(b + o_seg);
  const size_t seq_fixed = (size_t)kOsSeqStages * 32 * kOsRowStride * sizeof(int);
  const size_t state_bytes = round_up(n, 4) + 8;                           // one byte per voxel (at most n voxels)
  const bool seq_bytes = seq_fixed + state_bytes <= (size_t)kOsSeqSmemBytes && !getenv("TLOAM_B200_DCVC_PACKED");
  const int use_smem = seq_fixed + state_words * 4 <= (size_t)kOsSeqSmemBytes && !getenv("TLOAM_B200_DCVC_GLOBAL") ? 1 : 0;
  const size_t seq_smem = seq_fixed + (seq_bytes ? state_bytes : (use_smem ? state_words * 4 : 0));
  if (!h->os_attr_set) {
    CU_TRY(cudaFuncSetAttribute(k_os_seq<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kOsSeqSmemBytes));
    CU_TRY(cudaFuncSetAttribute(k_os_seq<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kOsSeqSmemBytes));
    h->os_attr_set = true;
  }
  cudaStream_t st = h->stream;
  if (h->seg.active) a.pts = h->seg.dev_xyz;
  else CU_TRY(cudaMemcpyAsync(b + o_pts, xyz, n * 24, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemsetAsync(b + o_tab, 0, cap * 8, st));
  if (!seq_bytes && !use_smem) CU_TRY(cudaMemsetAsync(b + o_stg, 0, state_words * 4, st));
  const unsigned ev_blocks = (unsigned)((n32 * 32 + 255) / 256);
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_init<<<1, 32, 0, st>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_polar<<<a.nchunk, kOsChunk, 0, st>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_bounds<<<1, 32, 0, st>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_key<<<a.nchunk, kOsChunk, 0, st>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_vid<<<(unsigned)(cap / 256), 256, 0, st>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_first<<<a.nchunk, kOsChunk, 0, st>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_second<<<a.nchunk, kOsChunk, 0, st>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_evkey<<<a.nchunk, kOsChunk, 0, st>>>(a)));
  auto partition = [&](const int* keys, const int* perm_in, int* perm_out, const int* n_items, int shift, int* total_out) {
    OsPart p;
    p.keys = keys; p.perm_in = perm_in; p.perm_out = perm_out; p.n_items = n_items; p.n_fixed = a.n; p.shift = shift;
    p.chunk_cnt = a.chunk_cnt; p.key_base = a.key_base; p.nchunk = a.nchunk; p.total_out = total_out;
    TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_part_hist<<<a.nchunk, kOsChunk, 0, st>>>(p)));
    TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_part_scan<<<1, kOsKeys, 0, st>>>(p)));
    TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_part_scatter<<<a.nchunk, kOsChunk, 0, st>>>(p)));
    return 0;
  };
  partition(a.evkey, nullptr, a.ev_pt, nullptr, 0, a.params + 5);          // events in point order
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_rows<<<ev_blocks, 256, 0, st>>>(a)));
  if (seq_bytes) TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_seq<true><<<1, 32, seq_smem, st>>>(a, 1)));
  else TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_seq<false><<<1, 32, seq_smem, st>>>(a, use_smem)));
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_union<<<ev_blocks, 256, 0, st>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_label<<<a.nchunk, kOsChunk, 0, st>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_clusters<<<a.nchunk, kOsChunk, 0, st>>>(a)));
  const size_t cl_bound = cfg->min_seg >= 0 ? n / ((size_t)cfg->min_seg + 1) + 1 : n;   // classes with > min_seg points
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_rank<<<(unsigned)((cl_bound + 255) / 256), 256, 0, st>>>(a)));
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_pkey<<<a.nchunk, kOsChunk, 0, st>>>(a)));
  partition(a.pkey, nullptr, a.perm_a, nullptr, 0, a.params + 7);           // stable LSD partition by cluster rank
  const int* seg = a.perm_a;
  if (cl_bound > 256) { partition(a.pkey, a.perm_a, a.perm_b, a.params + 7, 8, nullptr); seg = a.perm_b; }
  if (cl_bound > 65536) { partition(a.pkey, a.perm_b, a.perm_a, a.params + 7, 16, nullptr); seg = a.perm_a; }
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_boxes_acc<<<a.nchunk, 256, 0, st>>>(a, seg)));
  TL_LAUNCH(TLOAM_B200_K_OBJECT, (k_os_boxes_fin<<<(unsigned)((cl_bound + 255) / 256), 256, 0, st>>>(a)));
  CU_TRY(cudaGetLastError());
  int par[8];
  CU_TRY(cudaMemcpyAsync(par, a.params, sizeof(par), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  if (par[3] != 0) {
    snprintf(h->last_error, sizeof(h->last_error), "object_segmentation: more than %d polar rings", kOsMaxBounds);
    return TLOAM_B200_ERR_INVALID_ARG;
  }
  static_assert(sizeof(size_t) == sizeof(unsigned long long), "index lists are copied straight into size_t arrays");
  const size_t nc = (size_t)par[6], ns = (size_t)par[7];
  if (h->seg.active) { h->seg.seg = a.out_seg; h->seg.n_seg = (unsigned)ns; }   // chained: the scan stays on the device
  else if (ns) CU_TRY(cudaMemcpyAsync(seg_index, a.out_seg, ns * sizeof(size_t), cudaMemcpyDeviceToHost, st));
  if (sizes && nc) CU_TRY(cudaMemcpyAsync(sizes, a.sizes, nc * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (boxes && nc) CU_TRY(cudaMemcpyAsync(boxes, a.boxes, nc * 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (root) CU_TRY(cudaMemcpyAsync(root, a.root, n * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (cluster) CU_TRY(cudaMemcpyAsync(cluster, a.cluster, n * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (voxel) CU_TRY(cudaMemcpyAsync(voxel, a.key, n * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (polar) {
    CU_TRY(cudaMemcpyAsync(polar, a.polar, n * 24, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(polar + 3 * n, a.ext, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  CU_TRY(cudaStreamSynchronize(st));
  *n_seg = ns; *n_clusters = (int)nc;
  return TLOAM_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// "next" row (f)-4: the three steps of Segmentation::spinOnce (ref: segmentation.cpp:47-66) as ONE call: the scan is
// uploaded once, every stage is fed by a gather kernel from the previous stage's device-side index list, and only the
// final lists come home -- as indices into the ORIGINAL scan.
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_chain_gather_object(const double* pts, const int* beam, const unsigned* idx, unsigned n, double* out_pts,
                                                             double* out_beam) {
  const unsigned j = blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  const unsigned i = idx[j];
  out_pts[3ull * j] = pts[3ull * i]; out_pts[3ull * j + 1] = pts[3ull * i + 1]; out_pts[3ull * j + 2] = pts[3ull * i + 2];
  out_beam[j] = (double)beam[i];                                           // the intensity channel carries the beam id (:707-709)
}
__global__ void __launch_bounds__(256) k_chain_gather_segmented(const double* obj_pts, const double* obj_beam, const unsigned* obj_idx,
                                                                const unsigned long long* seg, unsigned n, double* out_pts, double* out_beam,
                                                                unsigned* out_orig) {
  const unsigned j = blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  const unsigned i = (unsigned)seg[j];
  out_pts[3ull * j] = obj_pts[3ull * i]; out_pts[3ull * j + 1] = obj_pts[3ull * i + 1]; out_pts[3ull * j + 2] = obj_pts[3ull * i + 2];
  out_beam[j] = obj_beam[i];
  out_orig[j] = obj_idx[i];
}
// blockIdx.y: 0 ground, 1 edge, 2 general -- index lists into the original scan, as 64-bit values
__global__ void __launch_bounds__(256) k_chain_final(const unsigned* ground, unsigned n_ground, const unsigned long long* edge, unsigned n_edge,
                                                     const unsigned long long* non_edge, unsigned n_non, const unsigned* seg_orig,
                                                     unsigned long long* out_ground, unsigned long long* out_edge, unsigned long long* out_general) {
  const unsigned j = blockIdx.x * 256 + threadIdx.x;
  if (blockIdx.y == 0) { if (j < n_ground) out_ground[j] = ground[j]; }
  else if (blockIdx.y == 1) { if (j < n_edge) out_edge[j] = seg_orig[(unsigned)edge[j]]; }
  else { if (j < n_non) out_general[j] = seg_orig[(unsigned)non_edge[j]]; }
}
}  // namespace

int tloam_b200_segment_scan(tloam_b200_handle* h, const tloam_ground_config* gcfg, const tloam_dcvc_config* dcfg, int ring_min_num,
                            const double* xyz, size_t n, size_t* ground_index, size_t* n_ground, size_t* edge_index, size_t* n_edge,
                            size_t* general_index, size_t* n_general, int* n_clusters, int* sizes, double* boxes, int* beam) {
  if (!h || !gcfg || !dcfg || !ground_index || !n_ground || !edge_index || !n_edge || !general_index || !n_general || !n_clusters)
    return TLOAM_B200_ERR_INVALID_ARG;
  *n_ground = *n_edge = *n_general = 0; *n_clusters = 0;
  if (n == 0) return TLOAM_B200_OK;
  if (!xyz || n > ((size_t)1 << 26)) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  // chain buffer: what must outlive a stage's arena
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += round_up(bytes, 256); return o; };
  const size_t o_scan = take(n * 24), o_gnd = take(n * 4), o_oidx = take(n * 4), o_opts = take(n * 24), o_obeam = take(n * 8),
               o_spts = take(n * 24), o_sbeam = take(n * 8), o_sorig = take(n * 4), o_fg = take(n * 8), o_fe = take(n * 8), o_fn = take(n * 8),
               o_beam = take(n * 4);
  if (off > h->cap_chain) {
    CU_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_chain); h->d_chain = nullptr; h->cap_chain = 0;
    CU_TRY(cudaMalloc(&h->d_chain, off + off / 4));
    h->cap_chain = off + off / 4;
  }
  unsigned char* c = h->d_chain;
  double* d_scan = (double*)(c + o_scan);
  unsigned* d_gnd = (unsigned*)(c + o_gnd); unsigned* d_oidx = (unsigned*)(c + o_oidx);
  double* d_opts = (double*)(c + o_opts); double* d_obeam = (double*)(c + o_obeam);
  double* d_spts = (double*)(c + o_spts); double* d_sbeam = (double*)(c + o_sbeam); unsigned* d_sorig = (unsigned*)(c + o_sorig);
  unsigned long long* d_fg = (unsigned long long*)(c + o_fg); unsigned long long* d_fe = (unsigned long long*)(c + o_fe);
  unsigned long long* d_fn = (unsigned long long*)(c + o_fn);
  if (HostStage::pageable(xyz) && getenv("TLOAM_B200_NO_HOST_STAGE") == nullptr) CU_TRY(h->hstage.upload(d_scan, xyz, n * 24, h->stream));
  else CU_TRY(cudaMemcpyAsync(d_scan, xyz, n * 24, cudaMemcpyHostToDevice, h->stream));
  struct Reset { tloam_b200_handle* h; ~Reset() { h->seg = tloam_b200_handle::SegChain(); } } reset{h};
  h->seg.active = true;
  // ---- 1. groundRemove ----
  h->seg.dev_xyz = d_scan;
  size_t ng = 0, no = 0;
  int rc = tloam_b200_ground_extract(h, gcfg, xyz, n, ground_index, &ng, general_index /*scratch: not written when chained*/, &no, nullptr, nullptr,
                                     nullptr, nullptr);
  if (rc != TLOAM_B200_OK) return rc;
  if (ng) CU_TRY(cudaMemcpyAsync(d_gnd, h->seg.ground, ng * 4, cudaMemcpyDeviceToDevice, h->stream));
  if (beam) CU_TRY(cudaMemcpyAsync(c + o_beam, h->seg.beam, n * 4, cudaMemcpyDeviceToDevice, h->stream));
  if (no) {
    CU_TRY(cudaMemcpyAsync(d_oidx, h->seg.object, no * 4, cudaMemcpyDeviceToDevice, h->stream));
    k_chain_gather_object<<<(unsigned)((no + 255) / 256), 256, 0, h->stream>>>(d_scan, h->seg.beam, h->seg.object, (unsigned)no, d_opts, d_obeam);
    CU_TRY(cudaGetLastError());
  }
  size_t ne = 0, nn = 0;
  if (no) {
    // ---- 2. objectSegmentation ----
    h->seg.dev_xyz = d_opts;
    size_t ns = 0;
    rc = tloam_b200_object_segmentation(h, dcfg, d_opts /*placeholder: read on the device*/, no, general_index, &ns, n_clusters, sizes, boxes, nullptr,
                                        nullptr, nullptr, nullptr);
    if (rc != TLOAM_B200_OK) return rc;
    if (ns) {
      k_chain_gather_segmented<<<(unsigned)((ns + 255) / 256), 256, 0, h->stream>>>(d_opts, d_obeam, d_oidx, h->seg.seg, (unsigned)ns, d_spts, d_sbeam,
                                                                                     d_sorig);
      CU_TRY(cudaGetLastError());
      // ---- 3. extractEdgePoint ----
      h->seg.dev_xyz = d_spts; h->seg.dev_intensity = d_sbeam;
      rc = tloam_b200_extract_edge(h, gcfg->sensor_model, ring_min_num, d_spts, d_sbeam, ns, edge_index, &ne, general_index, &nn);
      if (rc != TLOAM_B200_OK) return rc;
    }
  }
  const size_t most = ng > ne ? (ng > nn ? ng : nn) : (ne > nn ? ne : nn);
  if (most) {
    k_chain_final<<<dim3((unsigned)((most + 255) / 256), 3), 256, 0, h->stream>>>(d_gnd, (unsigned)ng, h->seg.edge, (unsigned)ne, h->seg.non_edge,
                                                                                   (unsigned)nn, d_sorig, d_fg, d_fe, d_fn);
    CU_TRY(cudaGetLastError());
    static_assert(sizeof(size_t) == sizeof(unsigned long long), "index lists are copied straight into size_t arrays");
    if (ng) CU_TRY(cudaMemcpyAsync(ground_index, d_fg, ng * 8, cudaMemcpyDeviceToHost, h->stream));
    if (ne) CU_TRY(cudaMemcpyAsync(edge_index, d_fe, ne * 8, cudaMemcpyDeviceToHost, h->stream));
    if (nn) CU_TRY(cudaMemcpyAsync(general_index, d_fn, nn * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  if (beam) CU_TRY(cudaMemcpyAsync(beam, c + o_beam, n * 4, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  *n_ground = ng; *n_edge = ne; *n_general = nn;
  return TLOAM_B200_OK;
}

int tloam_b200_host_alloc(void** p, size_t bytes) {
  if (!p) return TLOAM_B200_ERR_INVALID_ARG;
  return cudaMallocHost(p, bytes) == cudaSuccess ? TLOAM_B200_OK : TLOAM_B200_ERR_CUDA;
}
int tloam_b200_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? TLOAM_B200_OK : TLOAM_B200_ERR_CUDA; }

const char* tloam_b200_last_error(tloam_b200_handle* h) { return h ? h->last_error : ""; }

int tloam_b200_set_profiling(tloam_b200_handle* h, int on) {
  if (!h) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaStreamSynchronize(h->stream));
  h->profiling = on != 0;
  h->spans.clear(); h->ev_next = 0;
  memset(&h->prof, 0, sizeof(h->prof));
  if (on && !h->d_dbg) CU_TRY(cudaMalloc(&h->d_dbg, 16 * sizeof(unsigned long long)));
  if (h->d_dbg) {
    unsigned long long init[16];
    memset(init, 0, sizeof(init));
    init[0] = ~0ull;
    CU_TRY(cudaMemcpy(h->d_dbg, init, sizeof(init), cudaMemcpyHostToDevice));
  }
  h->ctx.dbg = on ? h->d_dbg : nullptr;
  return TLOAM_B200_OK;
}

int tloam_b200_get_profile(tloam_b200_handle* h, tloam_b200_profile* out) {
  if (!h || !out) return TLOAM_B200_ERR_INVALID_ARG;
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaStreamSynchronize(h->stream));
  for (const auto& sp : h->spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess && sp.cls >= 0 && sp.cls < TLOAM_B200_K_COUNT) {
      h->prof.launches[sp.cls] += 1;
      h->prof.total_ms[sp.cls] += ms;
    }
  }
  h->spans.clear(); h->ev_next = 0;
  if (h->d_dbg) CU_TRY(cudaMemcpy(h->prof.dbg, h->d_dbg, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  *out = h->prof;
  return TLOAM_B200_OK;
}

}  // extern "C"
