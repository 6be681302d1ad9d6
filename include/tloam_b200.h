/*
 * tloam_b200.h -- C ABI of the B200-native TLS scan-to-map registration library (libtloam_b200.so).
 *
 * This is the drop-in boundary for T-LOAM's pose-optimisation hot path.  Each entry point names the
 * reference interface it replaces ("ref:" paths are relative to the zhoupengwei/tloam tree):
 *
 *   tloam_b200_create           <- LocalRegistration::LocalRegistration(YAML::Node) + initConfig
 *                                  ref: src/models/registration/registration.cpp:182-230
 *   tloam_b200_set_source       <- RegistrationInterface::setInputSource(Frame&)
 *                                  ref: include/tloam/models/registration/registration_interface.hpp:44,
 *                                       registration.cpp:232-239
 *   tloam_b200_set_target       <- RegistrationInterface::setInputTarget(Frame&)   (+ the per-call KD-tree
 *                                  rebuild of registration.cpp:888-915, done ONCE per map here)
 *                                  ref: registration_interface.hpp:45, registration.cpp:241-248
 *   tloam_b200_scan_match       <- RegistrationInterface::scanMatching(Frame&, Isometry3d&, Isometry3d&)
 *                                  ref: registration_interface.hpp:46, registration.cpp:879-1133
 *   tloam_b200_fitness          <- RegistrationInterface::getFitnessScore()
 *                                  ref: registration_interface.hpp:47, registration.cpp:257-296
 *   tloam_b200_get_transform / _get_pose_increment <- getTransform() / getPoseIncrement()
 *                                  ref: registration.cpp:370-376
 *   tloam_b200_eval_point_to_{point,line,plane}    <- PointTo{Point,Line,Plane}Err::Evaluate
 *                                  ref: registration.cpp:19-47, 55-88, 96-117
 *
 * Conventions
 *   - clouds: contiguous AoS FP64 xyz (the layout of std::vector<Eigen::Vector3d>,
 *     ref: include/tloam/open3d/PointCloud2.hpp:396); array index 0..3 = edge, sphere, planar, ground
 *     (order of registration.cpp:233-236).
 *   - poses: 4x4 FP64 COLUMN-major (Eigen::Isometry3d::matrix().data()).
 *   - HOST inputs are copied at set_*(): caller buffers may be freed on return (unless tloam_b200_set_async_inputs).
 *     DEVICE inputs (set_*_device) are read in place by kernels on the handle's stream: keep them unchanged until
 *     that work has run, and order their producer with tloam_b200_wait_stream.
 *   - one handle = one CUDA device + one stream, single caller, not re-entrant (like the reference,
 *     ref: registration.hpp:327-329).  Distinct handles are independent.
 *   - no function aborts or throws; all return a tloam_b200_status.  There is NO CPU fallback: without a
 *     CUDA device tloam_b200_create returns TLOAM_B200_ERR_NO_DEVICE.
 */
#ifndef TLOAM_B200_H
#define TLOAM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TLOAM_B200_MAX_OUTER 16
#define TLOAM_B200_MAX_INNER 8

typedef struct tloam_b200_handle tloam_b200_handle;

typedef enum tloam_b200_status {
  TLOAM_B200_OK = 0,
  TLOAM_B200_ERR_INVALID_ARG = 1,
  TLOAM_B200_ERR_TOO_FEW_POINTS = 2, /* a cloud has < 10 points: the reference asserts (registration.cpp:928-929) */
  TLOAM_B200_ERR_BAD_POSE = 3,       /* predict is not a rigid transform: Sophus would abort (so3.hpp:469-472) */
  TLOAM_B200_ERR_CUDA = 4,
  TLOAM_B200_ERR_NO_DEVICE = 5,
  TLOAM_B200_ERR_NOT_READY = 6,      /* scan_match before set_source / set_target */
  TLOAM_B200_ERR_NUMERIC = 7,        /* non-finite value met inside the solve */
  TLOAM_B200_ERR_MAP_DENSITY = 8     /* a map cell (edge = search radius) holds more than 65535 points */
} tloam_b200_status;

/* The "TLS:" YAML block (ref: config/mapping/lidar_odometry.yaml:23-39, read at registration.cpp:212-230)
 * as a POD, same names, same defaults, plus explicit switches for reference quirks. */
typedef struct tloam_tls_config {
  int k_corr;            /* on the API surface, unused by the executed path */
  int factor_num;        /* 2 = planar+ground, 3 = +edge, 4 = +sphere (registration.hpp:144-148) */
  double edge_dist_thres, sphere_dist_thres, planar_dist_thres, ground_dist_thres;
  double edge_dir_thres;
  int edge_maxnum, sphere_maxnum, planar_maxnum, ground_maxnum;
  int max_iterations;    /* outer GNC iterations */
  double cost_threshold, gnc_factor, noise_bound, fitness_thres;
  /* extras */
  int ceres_max_num_iterations; /* options.max_num_iterations, registration.cpp:1043 */
  double reinit_dir[3];  /* replaces the unseeded Eigen::Vector3d::Random() of registration.cpp:885 */
  double initial_trust_region_radius; /* Ceres Solver::Options default 1e4 (the reference does not set it,
                          * registration.cpp:1036-1047); smaller values force the dogleg / rejected-step branches (tests) */
} tloam_tls_config;

typedef struct tloam_b200_inner_trace {
  double x_candidate[6];
  double candidate_cost, model_cost_change, relative_decrease, step_norm_scaled, radius;
  int accepted;          /* 1 accepted, 0 rejected, -1 invalid step, 2 terminated by a tolerance */
  int used_gauss_newton;
} tloam_b200_inner_trace;

typedef struct tloam_b200_outer_trace {
  double x_start[6], x_end[6];
  double initial_cost, final_cost;
  double H0[36], g0[6];  /* J^T J / J^T r (robustified) at x_start, row-major */
  double mu, th1, th2;
  double slot_sum[4];
  int n_factors[4];
  int n_inner, termination; /* 0 max-iter, 1 function tol, 2 parameter tol, 3 gradient tol, 4 radius, 5 no residuals, 6 invalid steps */
  tloam_b200_inner_trace inner[TLOAM_B200_MAX_INNER];
} tloam_b200_outer_trace;

typedef struct tloam_b200_stats {
  int n_outer, converged_early;
  double x_init[6], x_final[6];
  int gpu_launches;      /* kernels launched by this scan_match call */
  float gpu_ms;          /* device time of the call (CUDA events on the handle's stream) */
  tloam_b200_outer_trace outer[TLOAM_B200_MAX_OUTER];
} tloam_b200_stats;

void tloam_b200_default_config(tloam_tls_config* cfg);
const char* tloam_b200_status_string(int status);
const char* tloam_b200_last_error(tloam_b200_handle* h); /* text of the last CUDA failure of this handle */

/* device: CUDA ordinal. stream: a cudaStream_t to enqueue on (e.g. torch's current stream), or NULL to let
 * the handle create its own non-blocking stream. */
int tloam_b200_create(const tloam_tls_config* cfg, int device, void* stream, tloam_b200_handle** out);
int tloam_b200_destroy(tloam_b200_handle* h);

/* HOST buffers (pageable or pinned). set_target also builds the voxel-hash grids on the device.  Pinned (or registered)
 * buffers are DMA'd directly; ordinary pageable buffers -- what an unmodified front end holds, std::vector<Eigen::Vector3d>
 * -- are detected and staged by the library itself: 2 MB chunks copied by a small pool of host threads (started on first
 * use) into a ring of pinned slots while the DMA engine drains them (tloam_b200/csrc/host_stage.h; 35 GB/s instead of the
 * ~11 GB/s of cudaMemcpyAsync from pageable memory; TLOAM_B200_NO_HOST_STAGE=1 in the environment restores the latter).
 * Either way the caller's buffers have been read completely when the call returns. */
int tloam_b200_set_source(tloam_b200_handle* h, const double* const xyz[4], const size_t n[4]);
int tloam_b200_set_target(tloam_b200_handle* h, const double* const xyz[4], const size_t n[4]);
/* DEVICE buffers (inputs already resident in HBM), same layout.  Returns without waiting: the buffers are read IN
 * PLACE (no staging copy) by kernels enqueued on the handle's stream, so they must stay unchanged until that work
 * has run (tloam_b200_synchronize, or the get_result of the frame that follows). */
int tloam_b200_set_source_device(tloam_b200_handle* h, const double* const d_xyz[4], const size_t n[4]);
int tloam_b200_set_target_device(tloam_b200_handle* h, const double* const d_xyz[4], const size_t n[4]);
/* Stream ordering of device inputs: makes the handle's stream wait (on the device) for everything enqueued so far on
 * `producer_stream` (cudaStream_t; NULL = the legacy default stream).  Call it before set_*_device when the buffers were
 * written on another stream; the handle's own stream (tloam_b200_create's `stream`) needs no call. */
int tloam_b200_wait_stream(tloam_b200_handle* h, void* producer_stream);

/* Blocking: enqueues the frame, waits, returns the pose (and optionally the trace). */
int tloam_b200_scan_match(tloam_b200_handle* h, const double predict[16], double result[16], tloam_b200_stats* stats);
/* Split form: enqueue only / wait + fetch. Lets one host thread drive several handles (one per GPU). */
int tloam_b200_scan_match_async(tloam_b200_handle* h, const double predict[16]);
int tloam_b200_get_result(tloam_b200_handle* h, double result[16], tloam_b200_stats* stats);
/* "Next" row (f)-3: constant-velocity prediction on the device (ref: src/front_end/front_end.cpp:329-330:
 * step = last^-1 * pose; predict = pose * step).  The frame is enqueued with NO host input: the prediction is computed
 * by the frame's first kernel from the two last results, which live in device memory (the pose returned by frame k
 * and the one returned by frame k-1; identity before the first frame).  set_pose_history seeds / overrides them. */
int tloam_b200_scan_match_predicted_async(tloam_b200_handle* h);
int tloam_b200_scan_match_predicted(tloam_b200_handle* h, double result[16], tloam_b200_stats* stats);
int tloam_b200_set_pose_history(tloam_b200_handle* h, const double last_pose[16], const double curr_pose[16]);
/* The per-iteration trace in tloam_b200_stats costs device time; scan_match records it iff stats != NULL, the
 * async form iff it was switched on here (default off; without it get_result fills only gpu_launches / gpu_ms). */
int tloam_b200_set_trace(tloam_b200_handle* h, int on);

/* ---- batched registration: S independent sequences (one handle each: own map, scan, pose history, submap) whose
 * frames are registered TOGETHER -- one launch sequence per batch frame instead of S.  The reference runs one
 * LocalRegistration per nodelet; this is S of them stepped in lock-step on one GPU (a single
 * 40k-feature frame cannot fill 148 SMs).  Per-sequence poses are bit-identical to the un-batched calls.
 * All sequences share `cfg`.  Feed the sequences through tloam_b200_batch_handle(b, i) (any per-handle set_* /
 * submap_* entry point) or the batch_set_* forms: xyz = S*4 pointers, n = S*4 counts, sequence-major, cloud order
 * edge, sphere, planar, ground. ---- */
typedef struct tloam_b200_batch tloam_b200_batch;
int tloam_b200_batch_create(const tloam_tls_config* cfg, int device, int S, tloam_b200_batch** out);   /* 1 <= S <= 32 */
int tloam_b200_batch_destroy(tloam_b200_batch* b);
int tloam_b200_batch_size(tloam_b200_batch* b);
tloam_b200_handle* tloam_b200_batch_handle(tloam_b200_batch* b, int i);   /* owned by the batch: do not destroy */
int tloam_b200_batch_set_target(tloam_b200_batch* b, const double* const* xyz, const size_t* n);          /* HOST */
int tloam_b200_batch_set_source(tloam_b200_batch* b, const double* const* xyz, const size_t* n);
int tloam_b200_batch_set_target_device(tloam_b200_batch* b, const double* const* d_xyz, const size_t* n); /* DEVICE */
int tloam_b200_batch_set_source_device(tloam_b200_batch* b, const double* const* d_xyz, const size_t* n);
/* predicts: S x 16 doubles (column-major 4x4 each) or NULL = constant-velocity prediction on the device per sequence
 * (front_end.cpp:329-330).  results: S x 16; statuses (optional): S tloam_b200_status values.  Returns OK iff every
 * sequence returned OK, else the first failing status. */
int tloam_b200_batch_scan_match(tloam_b200_batch* b, const double* predicts, double* results, int* statuses);
int tloam_b200_batch_scan_match_async(tloam_b200_batch* b, const double* predicts);
int tloam_b200_batch_get_results(tloam_b200_batch* b, double* results, int* statuses, float* gpu_ms /* optional */);
long long tloam_b200_batch_launch_count(tloam_b200_batch* b);   /* kernels launched by the batch and its handles */
const char* tloam_b200_batch_last_error(tloam_b200_batch* b);
/* per-kernel-class timing of the batch frame kernels (see tloam_b200_set_profiling; declared below) */
struct tloam_b200_profile;
int tloam_b200_batch_set_profiling(tloam_b200_batch* b, int on);
int tloam_b200_batch_get_profile(tloam_b200_batch* b, struct tloam_b200_profile* out);

int tloam_b200_fitness(tloam_b200_handle* h, double* fitness, double* rmse);
/* getFitnessScore as a per-frame health metric of the asynchronous flow (ref: registration.cpp:257-296; the reference
 * declares it on the interface and never calls it): when switched on, every scan_match also scores its scan against
 * the map (two kernels in the frame graph, no allocation, no extra synchronisation) and the pair comes back with the
 * frame's result. */
int tloam_b200_set_frame_fitness(tloam_b200_handle* h, int on);
int tloam_b200_get_frame_fitness(tloam_b200_handle* h, double* fitness, double* rmse);   /* of the last fetched result */
/* Pipelined use (front end one frame ahead of the GPU): set_source / submap_update return without waiting for their
 * uploads -- the HOST buffers must stay valid until the frame's result has been fetched -- and get_result waits for the
 * OLDEST un-fetched frame only (at most 2 frames in flight; no per-iteration trace in this mode). */
int tloam_b200_set_async_inputs(tloam_b200_handle* h, int on);
int tloam_b200_get_transform(tloam_b200_handle* h, double pose[16]);
int tloam_b200_get_pose_increment(tloam_b200_handle* h, double pose[16]);
/* self-check of the dense-map correspondence path (env TLOAM_B200_DENSE_CHECK=1 at create): every query it searches is
 * searched again by the plain path and compared.  out: [0] queries, [1] differing kNN lists, [2] work items, [3] TMA
 * staging passes, [4..7] details of the first mismatch, [8..11] SM cycles / 64 per phase (item total, TMA wait, fine
 * sort, search), [12] work items with <= 8 queries; accumulated since creation. */
int tloam_b200_dense_check_counters(tloam_b200_handle* h, unsigned out[16]);
int tloam_b200_synchronize(tloam_b200_handle* h);
/* total kernels launched by this handle so far */
long long tloam_b200_launch_count(tloam_b200_handle* h);

/* ---- shared-map transport (multi-GPU, config 4): the built map (cell-sorted float4 points + hash
 * tables + header) is one contiguous device blob so that a single ncclBroadcast moves it. ---- */
int tloam_b200_map_blob_size(tloam_b200_handle* h, size_t* bytes);
int tloam_b200_map_export(tloam_b200_handle* h, void* d_dst, size_t bytes);       /* D2D copy out */
int tloam_b200_map_import(tloam_b200_handle* h, const void* d_src, size_t bytes); /* D2D copy in, adopt */
int tloam_b200_get_map_origin(tloam_b200_handle* h, double origin[3]);
/* Zero-copy form (config 4: ONE collective per map epoch, no size handshake, no host synchronisation, no staging
 * copies).  The blob layout is a pure function of the configuration and the four point counts, so every rank can lay
 * the incoming blob out from n[4] alone.
 *   sender:    tloam_b200_map_send_buffer  -> the built blob itself; order the stream of the collective behind the build
 *              with tloam_b200_signal_stream(h, that_stream)
 *   receiver:  tloam_b200_map_recv_buffer  -> a SECOND blob of the handle (frames keep registering against the active map
 *              while the next one is in flight); enqueue the collective into it on any stream; then
 *              tloam_b200_map_adopt(h, that_stream): device-side wait + pointer swap, nothing else */
int tloam_b200_map_layout_bytes(tloam_b200_handle* h, const size_t n[4], size_t* bytes);
int tloam_b200_map_send_buffer(tloam_b200_handle* h, void** d_ptr, size_t* bytes);
int tloam_b200_map_recv_buffer(tloam_b200_handle* h, const size_t n[4], void** d_ptr, size_t* bytes);
int tloam_b200_map_adopt(tloam_b200_handle* h, void* producer_stream);
int tloam_b200_signal_stream(tloam_b200_handle* h, void* consumer_stream);

/* ---- piecewise entry points (parity tests; host arrays in, host arrays out, computed on the GPU) ---- */
/* Exact radius-truncated kNN on the built map of `cloud` (KDTreeFlann::SearchHybrid semantics): idx/d2 are
 * nq*k, ascending (d2, index), padded with -1 / +inf; count[i] = neighbours strictly inside the radius. */
int tloam_b200_knn(tloam_b200_handle* h, int cloud, const double* queries, size_t nq, double radius, int k,
                   int* idx, double* d2, int* count);
/* Correspondence search + primitive fit + caps for one cloud at tangent x (all weights 1):
 * valid[i] in {0,1}; prim = n*6 doubles: plane (n,d,0,0), line (a,b), point (q,0,0,0). */
int tloam_b200_build_factors(tloam_b200_handle* h, int cloud, const double x[6], int* valid, double* prim, size_t n);
/* Batched cost functors: arrays of m factors; r is m*3 (m*1 for plane), J is m*18 (m*6), cost is m. */
int tloam_b200_eval_point_to_point(tloam_b200_handle* h, const double x[6], size_t m, const double* p,
                                   const double* q, const double* w, double* r, double* J, double* cost);
int tloam_b200_eval_point_to_line(tloam_b200_handle* h, const double x[6], size_t m, const double* p,
                                  const double* a, const double* b, const double* w, double* r, double* J,
                                  double* cost);
int tloam_b200_eval_point_to_plane(tloam_b200_handle* h, const double x[6], size_t m, const double* p,
                                   const double* n, const double* d, const double* w, double* r, double* J,
                                   double* cost);
/* SE(3) helpers evaluated on the device (exp: tangent -> 4x4 col-major; log: inverse; plus: left update). */
int tloam_b200_se3_exp(tloam_b200_handle* h, const double a[6], double T[16]);
int tloam_b200_se3_log(tloam_b200_handle* h, const double T[16], double a[6]);
int tloam_b200_se3_plus(tloam_b200_handle* h, const double x[6], const double delta[6], double out[6]);
/* the 2-D trust-region boundary problem of the subspace dogleg as the device solver runs it:
 * minimise 0.5 y^T B y + g^T y on |y| = radius (B row-major 2x2) */
int tloam_b200_min_on_boundary_2d(tloam_b200_handle* h, const double B[4], const double g[2], double radius, double y[2]);

/* ---- per-kernel timing (off by default): CUDA events on the handle's stream around EVERY launch.  The
 * bracketing adds ~1-2 us of event overhead per launch, so profiled durations are upper bounds. ---- */
enum {
  TLOAM_B200_K_MAP_BBOX = 0, TLOAM_B200_K_MAP_ORIGIN, TLOAM_B200_K_MAP_INSERT, TLOAM_B200_K_MAP_OFFSETS,
  TLOAM_B200_K_MAP_SCATTER, TLOAM_B200_K_STAGE_SOURCE, TLOAM_B200_K_BEGIN_FRAME, TLOAM_B200_K_CORRESPOND,
  TLOAM_B200_K_EVAL_FIRST, TLOAM_B200_K_EVAL, TLOAM_B200_K_SUBMAP, TLOAM_B200_K_FEATURE, TLOAM_B200_K_FIRST,
  TLOAM_B200_K_DENSE_BIN, TLOAM_B200_K_DENSE, TLOAM_B200_K_FITNESS, TLOAM_B200_K_GROUND, TLOAM_B200_K_MAP_FINE,
  TLOAM_B200_K_FINE, TLOAM_B200_K_EDGE, TLOAM_B200_K_OBJECT, TLOAM_B200_K_COUNT
};
typedef struct tloam_b200_profile {
  long long launches[TLOAM_B200_K_COUNT];
  double total_ms[TLOAM_B200_K_COUNT];
  /* in-kernel timers (profiling mode): [1] ns k_eval parallel phase, [2] ns final partial sum, [3] ns solver state
   * machine, [4] number of active k_eval launches, [8] SM cycles k_correspond kNN phase summed over thread blocks,
   * [9] number of k_correspond thread blocks */
  unsigned long long dbg[16];
} tloam_b200_profile;
int tloam_b200_set_profiling(tloam_b200_handle* h, int on);   /* also clears the accumulated profile */
int tloam_b200_get_profile(tloam_b200_handle* h, tloam_b200_profile* out);

/* ---- device-side local-map maintenance ("next" row (f)-1): FrontEnd::updateSubmap on the GPU
 * (ref: src/front_end/front_end.cpp:201-267, first frame :285-305; PointCloud2 Transform / += / Crop /
 * VoxelDownSample, ref: src/open3d/PointCloud2.cpp:71-75, 96-132, 358-403, 551-559).  The map stays in HBM between
 * frames; each call ends with the same voxel-hash build as tloam_b200_set_target. ---- */
typedef struct tloam_submap_config {   /* ref: config/mapping/lidar_odometry.yaml:6-17 */
  double ground_down_sample;           /* 0.3  */
  double ground_down_sample_submap;    /* 0.45 */
  double edge_down_sample_submap;      /* 0.3  */
  int planar_frame_size;               /* 3 */
  int sphere_frame_size;               /* 3 (kept for parity; the reference builds the sphere submap from the planar buffer) */
  double edge_crop_box_length, ground_crop_box_length;   /* 100, 100 */
} tloam_submap_config;
void tloam_b200_submap_default_config(tloam_submap_config* c);
/* First frame: edge = raw edge cloud, ground_raw = raw ground cloud (voxel-down-sampled at ground_down_sample
 * inside), planar_sub / sphere_sub = the "submap index" selections of the general cloud.  HOST pointers. */
int tloam_b200_submap_init(tloam_b200_handle* h, const tloam_submap_config* cfg, const double* edge, size_t ne,
                           const double* ground_raw, size_t ng, const double* planar_sub, size_t np,
                           const double* sphere_sub, size_t ns);
/* Later frames, after scan_match: pose = the new lidar_odom_pose (4x4 column-major).  The edge and ground
 * features appended to the map are the ones of the CURRENT SOURCE (tloam_b200_set_source), already on the device;
 * planar_sub is this frame's planar submap selection (HOST pointer, sensor frame); sphere_sub is accepted for
 * interface parity and ignored, as in the reference (front_end.cpp:220-230 iterates the planar buffer). */
int tloam_b200_submap_update(tloam_b200_handle* h, const double pose[16], const double* planar_sub, size_t np,
                             const double* sphere_sub, size_t ns);
/* Chained form: pose = the result of the frame that was just ENQUEUED on this handle (scan_match_async /
 * scan_match_predicted_async), read on the device -- no host round trip between registration and map update. */
int tloam_b200_submap_update_chained(tloam_b200_handle* h, const double* planar_sub, size_t np);
int tloam_b200_submap_sizes(tloam_b200_handle* h, size_t n[4]);   /* exact sizes (synchronises) */
/* copies map cloud `cloud` (0 edge, 1 sphere, 2 planar, 3 ground; world frame, FP64 AoS) to the host */
int tloam_b200_submap_download(tloam_b200_handle* h, int cloud, double* out, size_t capacity_points);
/* PointCloud2::VoxelDownSample on the device (HOST in / out; out must hold n points). */
int tloam_b200_voxel_down_sample(tloam_b200_handle* h, const double* pts, size_t n, double voxel, double* out, size_t* n_out);

/* ------------------------------------------------------------------------------------------------
 * "Next" row (f)-2: PCA feature extraction on the device.  Replaces featureExtract::extractPlanarSphere /
 * calculatePCAInfo (ref: src/models/feature_extraction/feature_extract.cpp:133-197, 47-122; called from
 * FrontEnd::processCloud, src/front_end/front_end.cpp:194 and :289).  Bit-exact against oracle/feature_oracle.cpp.
 * The handle is only used for its device, stream and scratch memory: no map or scan needs to be set. */
typedef struct tloam_feature_config {   /* ref: config/mapping/feature.yaml */
  double radius;                 /* 0.2  neighbour search radius */
  int K;                         /* 20   neighbours per point (3 <= K <= 20 supported) */
  int min_neigh;                 /* 10   points with <= min_neigh neighbours are skipped */
  int planar_num, sphere_num;    /* 500, 300 */
  double cvr_scan, cvr_submap;   /* 0.25, 0.15 */
  double planar_scan_thres, planar_submap_thres, planar_vertic_thres;   /* 0.75, 0.65, 0.25 */
} tloam_feature_config;
void tloam_b200_feature_default_config(tloam_feature_config* c);
/* extractPlanarSphere: xyz = the general cloud (HOST, n x 3 FP64).  Each index buffer must hold n entries; the four
 * lists are what the reference's four std::vector<size_t> receive, INCLUDING its quirk that the two sphere lists
 * hold ranks 0..count-1 instead of point indices (feature_extract.cpp:183-188).  sphere_candidates (optional, n
 * entries) receives the point indices those ranks refer to (sphere candidates by descending flatness). */
int tloam_b200_extract_planar_sphere(tloam_b200_handle* h, const tloam_feature_config* cfg, const double* xyz, size_t n,
                                     size_t* planar_scan_index, size_t* n_planar_scan, size_t* planar_submap_index,
                                     size_t* n_planar_submap, size_t* sphere_scan_index, size_t* n_sphere_scan,
                                     size_t* sphere_submap_index, size_t* n_sphere_submap, size_t* sphere_candidates);
/* calculatePCAInfo (inspection / tests): per point cvr, flatness, sphericity, normal (n x 3), num_sum and the
 * neighbour list (n x K, ascending distance, -1 padded).  Any output pointer may be NULL. */
int tloam_b200_pca_info(tloam_b200_handle* h, const tloam_feature_config* cfg, const double* xyz, size_t n, double* cvr,
                        double* flatness, double* sphericity, double* normal, int* num_sum, int* neigh);

/* ------------------------------------------------------------------------------------------------
 * "Next" row (f)-4, first part: multi-region ground extraction of the segmentation nodelet on the device.  Replaces
 * Segmentation::groundRemove (ref: src/models/segmentation/segmentation.cpp:738-770) with everything it calls:
 * initSections / getSection (:174-238), estimateRingsAndTimes2 HDL_64E (:341-384), filterByHeight (:454-470),
 * fillSectionIndex (:507-541), segmentGroundThread (:626-730), findBestPlane (:551-616).  Bit-exact against
 * oracle/segmentation_oracle.cpp.  The handle is only used for its device, stream and scratch memory. */
typedef struct tloam_ground_config {   /* ref: config/mapping/segmentation.yaml (velodyne: / groundSeg:) */
  int sensor_model;                    /* 64: HDL-64E, the only branch built */
  double sensor_height;                /* 1.73 */
  double vertical_res, init_angle;     /* 0.4, -24.9 */
  double sensor_min_range, sensor_max_range;   /* 1.0, 120.0 */
  int quadrant, num_sec;               /* 4, 3 */
  double plane_dis;                    /* groundSeg.dis 0.3 */
  int max_iter, ground_seed_num;       /* 3, 20 */
} tloam_ground_config;
void tloam_b200_ground_default_config(tloam_ground_config* c);
/* xyz: the scan AFTER RemoveClosedNonFinitePoints, in acquisition order (HOST, n x 3 FP64).  Index buffers hold n entries.
 * ground_index / object_index receive the indices (into xyz) of the points the reference's ground_scan / object_scan
 * receive, in that order with regions taken in (quadrant, section) order (the reference appends regions from four
 * racing threads); points of a region with <= 3 seeds reach neither list, as in the reference.  Optional outputs:
 * beam (n): the beam estimate stored in the intensity channel; region (n): quadrant * num_sec + section, 12 = above the
 * height threshold, 13 = dropped; height_threshold: mean z + 0.5; planes: 12 x 8 x 4 plane models [region][iteration]. */
int tloam_b200_ground_extract(tloam_b200_handle* h, const tloam_ground_config* cfg, const double* xyz, size_t n,
                              size_t* ground_index, size_t* n_ground, size_t* object_index, size_t* n_object, int* beam,
                              int* region, double* height_threshold, double* planes);

/* ---- "next" row (f)-4, second part: LOAM-style edge extraction of the segmentation nodelet,
 * Segmentation::extractEdgePoint + extractFromSection (ref: src/models/segmentation/segmentation.cpp:1144-1304, called at
 * :64 on the clustered object points).  xyz: n AoS points; intensity: the beam id of every point (the reference keeps it in
 * the intensity channel, (int)intensity must lie in [0, sensor_model), sensor_model <= 64); beams with fewer than ring_min_num
 * points are skipped (config/mapping/segmentation.yaml: 131).  Outputs are index lists into the input, in the order in which
 * the reference appends the points: edge_index (capacity n) by (beam, sector, descending curvature), <= 20 per sector;
 * non_edge_index (capacity n) by (beam, sector, ascending curvature).  Bit-exact against oracle/segmentation_oracle.cpp.
 * A sector with more than 4096 curvature values (a beam with more than 24 586 points) is rejected: INVALID_ARG. */
int tloam_b200_extract_edge(tloam_b200_handle* h, int sensor_model, int ring_min_num, const double* xyz, const double* intensity,
                            size_t n, size_t* edge_index, size_t* n_edge, size_t* non_edge_index, size_t* n_non_edge);

/* ---- "next" row (f)-4, third part: object segmentation = Dynamic Curved-Voxel Clustering of the non-ground points,
 * Segmentation::objectSegmentation (ref: src/models/segmentation/segmentation.cpp:1085-1112) and what it calls:
 * convertToPolar (:790-837), getPolarIndex (:777-784), createHashTable (:843-874), searchKNN (:886-908), DCVC (:915-990),
 * labelAnalysis (:998-1025), colorSegmentation (:1032-1078).  The sequential labelling is replayed exactly (see
 * tloam_b200/csrc/object_segment.cuh); integer outputs are bit-exact against oracle/segmentation_oracle.cpp given the
 * same polar triples, the triples themselves agree with libm's to <= 4 ulp. */
typedef struct tloam_dcvc_config {      /* ref: config/mapping/segmentation.yaml (DCVC: / velodyne:) */
  double start_r, delta_r, delta_p, delta_a;   /* 0.35, 0.0004, 1.2, 1.2 */
  int min_seg;                                 /* 80: classes with <= min_seg points are filtered out */
  double sensor_min_range, sensor_max_range;   /* 1.0, 120.0 */
  /* the members minPitch / maxPitch / minPolar / maxPolar before the scan: resetParams() leaves 0.0
   * (segmentation.cpp:1121-1123); on the very first frame the polar pair is 5.0 (segmentation.hpp:332-333) */
  double min_pitch_init, max_pitch_init, min_polar_init, max_polar_init;
} tloam_dcvc_config;
void tloam_b200_dcvc_default_config(tloam_dcvc_config* c);
/* xyz: the object scan (HOST, n x 3 FP64, finite).  seg_index (capacity n): indices of the reference's segmented_scan,
 * cluster after cluster (size descending; equal sizes by smallest point index -- the reference leaves that to
 * unordered_map order), index order inside a cluster; sizes (capacity n) / boxes (capacity n x 6: centre xyz,
 * dimensions xyz) per cluster.  Optional outputs: root (n): smallest point index of the point's DCVC class (a canonical
 * form of label_info); cluster (n): 1-based cluster number, 0 = filtered; voxel (n): the reference's voxelIndex;
 * polar (3 n + 4): range / pitch / azimuth triples followed by minPitch, maxPitch, minPolar, maxPolar.
 * More than 4096 polar rings (the default configuration has ~500 at 120 m): INVALID_ARG. */
int tloam_b200_object_segmentation(tloam_b200_handle* h, const tloam_dcvc_config* cfg, const double* xyz, size_t n,
                                   size_t* seg_index, size_t* n_seg, int* n_clusters, int* sizes, double* boxes, int* root,
                                   int* cluster, int* voxel, double* polar);

/* The three steps of Segmentation::spinOnce (ref: src/models/segmentation/segmentation.cpp:47-66: groundRemove,
 * objectSegmentation, extractEdgePoint(segmented_scan, edge_scan, general_scan)) as ONE call: the scan crosses PCIe once,
 * every stage is fed on the device from the previous stage's index list, only the final lists come home.  All three
 * lists index the ORIGINAL scan: ground_index = the reference's ground_scan, edge_index / general_index = its edge_scan /
 * general_scan (capacity n each), in the reference's order.  sizes / boxes (capacity n / n x 6, optional): per cluster as
 * in tloam_b200_object_segmentation; beam (capacity n, optional): the beam estimate of every point of the scan (what the
 * reference keeps in the intensity channel of the object / segmented / edge / general clouds).  Identical to calling the
 * three functions one after the other with host gathers in between (tests/test_object_segmentation.py). */
int tloam_b200_segment_scan(tloam_b200_handle* h, const tloam_ground_config* gcfg, const tloam_dcvc_config* dcfg, int ring_min_num,
                            const double* xyz, size_t n, size_t* ground_index, size_t* n_ground, size_t* edge_index, size_t* n_edge,
                            size_t* general_index, size_t* n_general, int* n_clusters, int* sizes, double* boxes, int* beam);

/* Pinned host memory helpers (optional; pinned inputs make set_* a direct DMA, no staging threads). */
int tloam_b200_host_alloc(void** p, size_t bytes);
int tloam_b200_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
