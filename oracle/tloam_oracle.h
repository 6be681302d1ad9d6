/*
 * tloam_oracle.h -- C API of the CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * The oracle is a dependency-free C++17 restatement of T-LOAM's TLS scan-to-map
 * registration (reference: src/models/registration/registration.cpp:14-117,
 * 162-179, 232-376, 427-635, 714-778, 858-1133, plus the vendored Sophus
 * se3.hpp/so3.hpp and the documented behaviour of the un-vendored Ceres 2.0 /
 * Open3D 0.12 KDTreeFlann calls made on that path).
 *
 * PARITY UNPINNED: the reference has no tests, golden vectors or fixtures for
 * this path and cannot be built in this container (Eigen, Ceres, Open3D, ROS,
 * yaml-cpp are absent), so this oracle is pinned only by independent
 * cross-checks (scipy cKDTree, scipy expm/logm, numpy eigh, central
 * differences) -- see tests/test_oracle_*.py and DESIGN.md.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may load this library.  The product path (tloam_b200/) never does.
 */
#ifndef TLOAM_ORACLE_H
#define TLOAM_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAX_OUTER 16
#define ORACLE_MAX_INNER 8

/* Cloud order everywhere: 0 = edge, 1 = sphere, 2 = planar, 3 = ground
 * (order of registration.cpp:233-236). */

/* The 16 keys of the YAML "TLS:" block (config/mapping/lidar_odometry.yaml:23-39,
 * read at registration.cpp:212-230) + explicit switches for reference quirks. */
typedef struct oracle_config {
  int k_corr;              /* unused on the executed path (Q17) */
  int factor_num;          /* 2 planar+ground, 3 +edge, 4 +sphere */
  double edge_dist_thres, sphere_dist_thres, planar_dist_thres, ground_dist_thres;
  double edge_dir_thres;
  int edge_maxnum, sphere_maxnum, planar_maxnum, ground_maxnum;
  int max_iterations;      /* outer GNC iterations (4) */
  double cost_threshold, gnc_factor, noise_bound, fitness_thres;
  /* --- extras (not in the YAML) --- */
  int ceres_max_num_iterations; /* registration.cpp:1043 (4) */
  double reinit_dir[3];    /* Q1: replaces Eigen::Vector3d::Random(); normalised inside */
  double initial_trust_region_radius; /* Ceres default 1e4 (not set by the reference); tests shrink it */
  int threads_mode;        /* 0 = reference-faithful thread structure, 1 = all cores */
  int num_threads;         /* 0 = omp_get_max_threads() */
} oracle_config;

typedef struct oracle_inner_trace {
  double x_candidate[6];
  double candidate_cost;
  double model_cost_change;
  double relative_decrease;
  double step_norm_scaled;   /* |D*step| of the dogleg step */
  double radius;             /* trust-region radius when the step was computed */
  int accepted;              /* 1 accepted, 0 rejected, -1 invalid, 2 terminated by tolerance */
  int used_gauss_newton;     /* 1 if |GN| <= radius */
} oracle_inner_trace;

typedef struct oracle_outer_trace {
  double x_start[6];
  double x_end[6];
  double initial_cost, final_cost;
  double H0[36];             /* J^T J (corrected, unscaled) at x_start, row-major 6x6 */
  double g0[6];              /* J^T r at x_start */
  double mu;                 /* mu used by this iteration's weight update */
  double th1, th2;
  double slot_sum[4];        /* sum of residual slots per cloud after the solve */
  int n_factors[4];          /* residual blocks added per cloud */
  int n_inner;               /* trust-region iterations executed (>=0) */
  int termination;           /* 0 max-iter, 1 function tol, 2 parameter tol, 3 gradient tol, 4 radius, 5 no residuals */
  oracle_inner_trace inner[ORACLE_MAX_INNER];
} oracle_outer_trace;

typedef struct oracle_stats {
  int n_outer;               /* outer iterations executed */
  int converged_early;       /* 1 if the planar-cost test broke the loop */
  double x_init[6];          /* tangent after log() and the Q1 re-init */
  double x_final[6];
  /* wall-clock stage breakdown (seconds) */
  double t_kdtree, t_factors, t_solve, t_weights, t_total;
  oracle_outer_trace outer[ORACLE_MAX_OUTER];
} oracle_stats;

void oracle_default_config(oracle_config* cfg);

void* oracle_create(const oracle_config* cfg);
void oracle_destroy(void* h);
/* xyz[c] = contiguous AoS doubles (x,y,z), n[c] points; copied. */
int oracle_set_source(void* h, const double* const xyz[4], const size_t n[4]);
int oracle_set_target(void* h, const double* const xyz[4], const size_t n[4]);
/* predict/result: 4x4 column-major (Eigen::Isometry3d layout). Returns 0 on success,
 * 1 if a cloud has < 10 points (the reference asserts, registration.cpp:928-929). */
int oracle_scan_match(void* h, const double predict[16], double result[16], oracle_stats* stats);
int oracle_fitness(void* h, double* fitness, double* rmse);
void oracle_get_pose_increment(void* h, double out[16]);
/* per-feature state after the last scan_match (weights as used by the last outer iteration's build,
 * i.e. before the final updateWeight; weights_after = after it). */
int oracle_get_weights(void* h, int cloud, double* weights_after, size_t n);

/* ---- piecewise entry points for unit / parity tests ---- */
void oracle_se3_exp(const double a[6], double T[16]);
void oracle_se3_log(const double T[16], double a[6]);
void oracle_se3_plus(const double x[6], const double delta[6], double out[6]); /* log(exp(delta)*exp(x)) */
void oracle_se3_exp_quat(const double a[6], double q_wxyz_t[7]);

/* Exact radius-truncated kNN (KDTreeFlann::SearchHybrid semantics): for each query writes up to k
 * indices / squared distances in ascending (d2, index) order; count[i] = number found (< k if fewer
 * points lie strictly within radius). idx/d2 are nq*k, padded with -1 / inf. */
int oracle_knn(const double* pts, size_t n, const double* queries, size_t nq, double radius, int k,
               int* idx, double* d2, int* count, int brute_force);

void oracle_fit_plane(const double* pts, int n, double out_nd[4]);             /* registration.cpp:303-368 */
/* registration.cpp:451-485: returns 1 if the line test passes; always fills mean/dir/eig. */
int oracle_fit_line(const double* pts, int n, double dir_thres, double a[3], double b[3], double mean[3],
                    double dir[3], double eig[3]);
void oracle_sym_eig3(const double cov[9], double eig[3], double vec_colmajor[9]);

void oracle_eval_point_to_point(const double x[6], const double p[3], const double q[3], double w,
                                double r[3], double J[18], double* cost);
void oracle_eval_point_to_line(const double x[6], const double p[3], const double a[3], const double b[3],
                               double w, double r[3], double J[18], double* cost);
void oracle_eval_point_to_plane(const double x[6], const double p[3], const double n[3], double d, double w,
                                double r[1], double J[6], double* cost);

/* Factor build for one cloud at tangent x with the handle's current weights = 1 (fresh):
 * valid[i] in {0,1}; prim is n*6 doubles: plane (n,d,0,0), line (a,b), point (q,0,0,0). */
int oracle_build_factors(void* h, int cloud, const double x[6], int* valid, double* prim, size_t n);

/* minimise 0.5 y^T B y + g^T y on |y| = radius (B row-major 2x2): the boundary problem of the subspace dogleg */
void oracle_min_on_boundary_2d(const double B[4], const double g[2], double radius, double y[2]);

/* GNC weight update, registration.cpp:858-876. */
void oracle_update_weight(double* weights, const double* slots, size_t n, double noise_bound_sq, double th1,
                          double th2, double mu);

/* ------------------------------------------------------------------------------------------------
 * "Next" row (f)-1: submap maintenance, FrontEnd::updateSubmap (ref: src/front_end/front_end.cpp:201-267) and the
 * first-frame seeding (ref: front_end.cpp:285-305), with PointCloud2::Transform / operator+= / Crop /
 * VoxelDownSample (ref: src/open3d/PointCloud2.cpp:71-75, 96-132, 358-403, 551-559).
 * VoxelDownSample's output order in the reference is std::unordered_map iteration order (implementation
 * defined); the oracle emits voxels in ascending (ix,iy,iz) order and tests compare as sets. */
typedef struct oracle_submap_config {
  double ground_down_sample;          /* 0.3  (scan ground, processCloud, front_end.cpp:183) */
  double ground_down_sample_submap;   /* 0.45 */
  double edge_down_sample_submap;     /* 0.3  */
  int planar_frame_size;              /* 3 */
  int sphere_frame_size;              /* 3 (maintained but unused by the reference) */
  double edge_crop_box_length, ground_crop_box_length;   /* 100, 100 */
} oracle_submap_config;
void oracle_submap_default_config(oracle_submap_config* c);
/* out must hold n points; returns the number of voxels. */
size_t oracle_voxel_down_sample(const double* pts, size_t n, double voxel, double* out);
size_t oracle_crop(const double* pts, size_t n, const double lo[3], const double hi[3], double* out);
void* oracle_submap_create(const oracle_submap_config* c);
void oracle_submap_destroy(void* s);
/* first frame: edge = raw edge cloud, ground = raw ground cloud (down-sampled inside), planar / sphere = the
 * "submap index" selections of the general cloud. */
int oracle_submap_init(void* s, const double* edge, size_t ne, const double* ground_raw, size_t ng,
                       const double* planar_sub, size_t np, const double* sphere_sub, size_t ns);
/* later frames: pose = lidar_odom_pose (4x4 column-major); edge_scan / ground_scan = current_scan features. */
int oracle_submap_update(void* s, const double pose[16], const double* edge_scan, size_t ne, const double* ground_scan,
                         size_t ng, const double* planar_sub, size_t np, const double* sphere_sub, size_t ns);
size_t oracle_submap_size(void* s, int cloud);
const double* oracle_submap_data(void* s, int cloud);

/* ------------------------------------------------------------------------------------------------
 * "Next" row (f)-2: PCA feature extraction, featureExtract::calculatePCAInfo / extractPlanarSphere
 * (ref: src/models/feature_extraction/feature_extract.cpp:47-122, 133-197; config/mapping/feature.yaml).
 * Implemented in feature_oracle.cpp (compiled with -ffp-contract=off, see there). */
typedef struct oracle_feature_config {
  double radius;                 /* 0.2 */
  int K;                         /* 20 */
  int min_neigh;                 /* 10 */
  int planar_num, sphere_num;    /* 500, 300 */
  double cvr_scan, cvr_submap;   /* 0.25, 0.15 */
  double planar_scan_thres, planar_submap_thres, planar_vertic_thres;   /* 0.75, 0.65, 0.25 */
} oracle_feature_config;
void oracle_feature_default_config(oracle_feature_config* c);
/* per-point PCAInfo: normal is n*3, neigh is n*K (ascending distance, -1 padded). Returns 1 for an empty cloud. */
int oracle_pca_info(const double* pts, size_t n, const oracle_feature_config* c, double* cvr, double* flatness,
                    double* sphericity, double* normal, int* num_sum, int* neigh);
/* the four index lists of extractPlanarSphere (each buffer holds n entries); sphere_candidates may be NULL. */
int oracle_extract_planar_sphere(const double* pts, size_t n, const oracle_feature_config* c, size_t* planar_scan,
                                 size_t* n_planar_scan, size_t* planar_submap, size_t* n_planar_submap,
                                 size_t* sphere_scan, size_t* n_sphere_scan, size_t* sphere_submap,
                                 size_t* n_sphere_submap, size_t* sphere_candidates);

/* ------------------------------------------------------------------------------------------------
 * "Next" row (f)-4, first part: multi-region ground extraction, Segmentation::groundRemove
 * (ref: src/models/segmentation/segmentation.cpp:738-770 and :174-238, 334-384, 454-470, 507-541, 551-730;
 * config/mapping/segmentation.yaml).  Implemented in segmentation_oracle.cpp (-ffp-contract=off, see there). */
typedef struct oracle_ground_config {
  int sensor_model;                 /* 64 (HDL-64E; the only branch restated) */
  double sensor_height;             /* 1.73 */
  double vertical_res, init_angle;  /* 0.4, -24.9 */
  double sensor_min_range, sensor_max_range;   /* 1.0, 120.0 */
  int quadrant, num_sec;            /* 4, 3 */
  double plane_dis;                 /* groundSeg.dis 0.3 */
  int max_iter, ground_seed_num;    /* 3, 20 */
} oracle_ground_config;
void oracle_ground_default_config(oracle_ground_config* c);
float oracle_fast_atan2(float y, float x);                                   /* cv::fastAtan2 */
int oracle_ground_section_bounds(const oracle_ground_config* c, float* out, int cap);
int oracle_ground_extract(const double* pts, size_t n, const oracle_ground_config* cfg, size_t* ground_index,
                          size_t* n_ground, size_t* object_index, size_t* n_object, int* beam, int* region,
                          double* mean_height_out, double* planes);

/* Segmentation::extractEdgePoint + extractFromSection (ref: src/models/segmentation/segmentation.cpp:1144-1304): LOAM-style
 * edge features by the smoothness of the local surface, per beam and per sixth of a beam.  pts: AoS xyz; intensity: the
 * beam id of every point (the reference keeps it in the intensity channel).  Outputs are INDEX lists into the input
 * (the reference emits the points themselves): edge_index in (beam, sector, descending curvature) order, non_edge_index in
 * (beam, sector, ascending curvature) order.  Returns 0, or -1 if a sector holds more than max_section curvature values
 * (device limit mirrored so that both sides reject the same inputs; max_section <= 0: no limit). */
int oracle_extract_edge(const double* pts, const double* intensity, size_t n, int sensor_model, int ring_min_num, int max_section,
                        size_t* edge_index, size_t* n_edge, size_t* non_edge_index, size_t* n_non_edge);

/* Segmentation::objectSegmentation = Dynamic Curved-Voxel Clustering (ref: src/models/segmentation/segmentation.cpp:772-1112;
 * config/mapping/segmentation.yaml DCVC + velodyne ranges).  Literal restatement, see segmentation_oracle.cpp. */
typedef struct oracle_dcvc_config {
  double start_r, delta_r, delta_p, delta_a;   /* 0.35, 0.0004, 1.2, 1.2 */
  int min_seg;                                 /* 80 */
  double sensor_min_range, sensor_max_range;   /* 1.0, 120.0 */
  double min_pitch_init, max_pitch_init, min_polar_init, max_polar_init;   /* member values left by resetParams: 0 (polar: 5.0 on the first frame) */
} oracle_dcvc_config;
void oracle_dcvc_default_config(oracle_dcvc_config* c);
void oracle_dcvc_polar(const double* pts, size_t n, const oracle_dcvc_config* c, double* polar, double* ext);
int oracle_dcvc_from_polar(const double* pts, const double* polar, const double* ext, size_t n, const oracle_dcvc_config* c,
                           int max_bounds, int* voxel, int* root, int* cluster, size_t* seg_index, size_t* n_seg, int* n_clusters,
                           int* sizes, double* boxes);
/* polar_out (optional): [n][3] triples followed by the 4 extrema */
int oracle_dcvc(const double* pts, size_t n, const oracle_dcvc_config* c, int max_bounds, double* polar_out, int* voxel, int* root,
                int* cluster, size_t* seg_index, size_t* n_seg, int* n_clusters, int* sizes, double* boxes);

#ifdef __cplusplus
}
#endif
#endif
