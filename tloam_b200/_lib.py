"""ctypes loader for libtloam_b200.so (the C-ABI CUDA library). Fails loudly: there is no CPU fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# TLOAM_B200_LIB: alternative build of the same library (A/B experiments with -D variants); never a different backend
LIB_PATH = os.environ.get("TLOAM_B200_LIB") or os.path.join(HERE, "libtloam_b200.so")

MAX_OUTER = 16
MAX_INNER = 8

OK, ERR_INVALID_ARG, ERR_TOO_FEW_POINTS, ERR_BAD_POSE, ERR_CUDA, ERR_NO_DEVICE, ERR_NOT_READY, ERR_NUMERIC, ERR_MAP_DENSITY = range(9)


class TlsConfig(C.Structure):
    """tloam_tls_config (include/tloam_b200.h) = the YAML "TLS:" block of the reference."""
    _fields_ = [
        ("k_corr", C.c_int), ("factor_num", C.c_int),
        ("edge_dist_thres", C.c_double), ("sphere_dist_thres", C.c_double),
        ("planar_dist_thres", C.c_double), ("ground_dist_thres", C.c_double),
        ("edge_dir_thres", C.c_double),
        ("edge_maxnum", C.c_int), ("sphere_maxnum", C.c_int), ("planar_maxnum", C.c_int), ("ground_maxnum", C.c_int),
        ("max_iterations", C.c_int),
        ("cost_threshold", C.c_double), ("gnc_factor", C.c_double), ("noise_bound", C.c_double),
        ("fitness_thres", C.c_double),
        ("ceres_max_num_iterations", C.c_int),
        ("reinit_dir", C.c_double * 3),
        ("initial_trust_region_radius", C.c_double),
    ]


class SubmapConfig(C.Structure):
    """tloam_submap_config (ref: config/mapping/lidar_odometry.yaml:6-17)."""
    _fields_ = [("ground_down_sample", C.c_double), ("ground_down_sample_submap", C.c_double),
                ("edge_down_sample_submap", C.c_double), ("planar_frame_size", C.c_int), ("sphere_frame_size", C.c_int),
                ("edge_crop_box_length", C.c_double), ("ground_crop_box_length", C.c_double)]


class InnerTrace(C.Structure):
    _fields_ = [
        ("x_candidate", C.c_double * 6), ("candidate_cost", C.c_double), ("model_cost_change", C.c_double),
        ("relative_decrease", C.c_double), ("step_norm_scaled", C.c_double), ("radius", C.c_double),
        ("accepted", C.c_int), ("used_gauss_newton", C.c_int),
    ]


class OuterTrace(C.Structure):
    _fields_ = [
        ("x_start", C.c_double * 6), ("x_end", C.c_double * 6),
        ("initial_cost", C.c_double), ("final_cost", C.c_double),
        ("H0", C.c_double * 36), ("g0", C.c_double * 6),
        ("mu", C.c_double), ("th1", C.c_double), ("th2", C.c_double),
        ("slot_sum", C.c_double * 4), ("n_factors", C.c_int * 4),
        ("n_inner", C.c_int), ("termination", C.c_int),
        ("inner", InnerTrace * MAX_INNER),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_outer", C.c_int), ("converged_early", C.c_int),
        ("x_init", C.c_double * 6), ("x_final", C.c_double * 6),
        ("gpu_launches", C.c_int), ("gpu_ms", C.c_float),
        ("outer", OuterTrace * MAX_OUTER),
    ]


KERNEL_CLASSES = ("map_bbox", "map_origin", "map_insert", "map_offsets", "map_scatter", "stage_source",
                  "begin_frame", "correspond", "eval_first", "eval", "submap", "feature", "first", "dense_bin", "dense", "fitness", "ground", "map_fine", "fine", "edge", "object")


class FeatureConfig(C.Structure):
    """tloam_feature_config (include/tloam_b200.h)."""
    _fields_ = [("radius", C.c_double), ("K", C.c_int), ("min_neigh", C.c_int), ("planar_num", C.c_int),
                ("sphere_num", C.c_int), ("cvr_scan", C.c_double), ("cvr_submap", C.c_double),
                ("planar_scan_thres", C.c_double), ("planar_submap_thres", C.c_double),
                ("planar_vertic_thres", C.c_double)]


class DcvcConfig(C.Structure):
    """tloam_dcvc_config (ref: config/mapping/segmentation.yaml DCVC + velodyne ranges)."""
    _fields_ = [("start_r", C.c_double), ("delta_r", C.c_double), ("delta_p", C.c_double), ("delta_a", C.c_double),
                ("min_seg", C.c_int), ("sensor_min_range", C.c_double), ("sensor_max_range", C.c_double),
                ("min_pitch_init", C.c_double), ("max_pitch_init", C.c_double), ("min_polar_init", C.c_double),
                ("max_polar_init", C.c_double)]


class GroundConfig(C.Structure):
    """tloam_ground_config (ref: config/mapping/segmentation.yaml)."""
    _fields_ = [("sensor_model", C.c_int), ("sensor_height", C.c_double), ("vertical_res", C.c_double), ("init_angle", C.c_double),
                ("sensor_min_range", C.c_double), ("sensor_max_range", C.c_double), ("quadrant", C.c_int), ("num_sec", C.c_int),
                ("plane_dis", C.c_double), ("max_iter", C.c_int), ("ground_seed_num", C.c_int)]


class Profile(C.Structure):
    _fields_ = [("launches", C.c_longlong * len(KERNEL_CLASSES)), ("total_ms", C.c_double * len(KERNEL_CLASSES)),
                ("dbg", C.c_ulonglong * 16)]


EXPORTS = [
    "tloam_b200_default_config", "tloam_b200_status_string", "tloam_b200_last_error", "tloam_b200_create",
    "tloam_b200_destroy", "tloam_b200_set_source", "tloam_b200_set_target", "tloam_b200_set_source_device",
    "tloam_b200_set_target_device", "tloam_b200_scan_match", "tloam_b200_scan_match_async", "tloam_b200_get_result",
    "tloam_b200_fitness", "tloam_b200_get_transform", "tloam_b200_get_pose_increment", "tloam_b200_synchronize",
    "tloam_b200_launch_count", "tloam_b200_map_blob_size", "tloam_b200_map_export", "tloam_b200_map_import",
    "tloam_b200_get_map_origin", "tloam_b200_knn", "tloam_b200_build_factors", "tloam_b200_eval_point_to_point",
    "tloam_b200_eval_point_to_line", "tloam_b200_eval_point_to_plane", "tloam_b200_se3_exp", "tloam_b200_se3_log",
    "tloam_b200_se3_plus", "tloam_b200_min_on_boundary_2d", "tloam_b200_host_alloc", "tloam_b200_host_free", "tloam_b200_set_profiling",
    "tloam_b200_get_profile", "tloam_b200_set_trace", "tloam_b200_submap_default_config", "tloam_b200_submap_init",
    "tloam_b200_submap_update", "tloam_b200_submap_sizes", "tloam_b200_submap_download", "tloam_b200_voxel_down_sample",
    "tloam_b200_scan_match_predicted_async", "tloam_b200_scan_match_predicted", "tloam_b200_set_pose_history",
    "tloam_b200_feature_default_config", "tloam_b200_extract_planar_sphere", "tloam_b200_pca_info",
    "tloam_b200_batch_create", "tloam_b200_batch_destroy", "tloam_b200_batch_size", "tloam_b200_batch_handle",
    "tloam_b200_batch_set_target", "tloam_b200_batch_set_source", "tloam_b200_batch_set_target_device",
    "tloam_b200_batch_set_source_device", "tloam_b200_batch_scan_match", "tloam_b200_batch_scan_match_async",
    "tloam_b200_batch_get_results", "tloam_b200_batch_launch_count", "tloam_b200_batch_last_error",
    "tloam_b200_batch_set_profiling", "tloam_b200_batch_get_profile",
    "tloam_b200_submap_update_chained", "tloam_b200_set_frame_fitness", "tloam_b200_get_frame_fitness",
    "tloam_b200_set_async_inputs", "tloam_b200_wait_stream", "tloam_b200_dense_check_counters",
    "tloam_b200_ground_default_config", "tloam_b200_ground_extract", "tloam_b200_extract_edge", "tloam_b200_dcvc_default_config", "tloam_b200_object_segmentation", "tloam_b200_segment_scan", "tloam_b200_map_layout_bytes",
    "tloam_b200_map_send_buffer", "tloam_b200_map_recv_buffer", "tloam_b200_map_adopt", "tloam_b200_signal_stream",
]

_lib = None


def load():
    """Load the CUDA library. Raises if it has not been built (python -m tloam_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m tloam_b200.build` "
                           "(there is no CPU fallback for the registration path)")
    L = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    vp = C.c_void_p
    L.tloam_b200_default_config.argtypes = [C.POINTER(TlsConfig)]
    L.tloam_b200_default_config.restype = None
    L.tloam_b200_status_string.argtypes = [C.c_int]
    L.tloam_b200_status_string.restype = C.c_char_p
    L.tloam_b200_last_error.argtypes = [vp]
    L.tloam_b200_last_error.restype = C.c_char_p
    L.tloam_b200_create.argtypes = [C.POINTER(TlsConfig), C.c_int, vp, C.POINTER(vp)]
    L.tloam_b200_destroy.argtypes = [vp]
    for name in ("set_source", "set_target"):
        f = getattr(L, "tloam_b200_" + name)
        f.argtypes = [vp, C.POINTER(dp), C.POINTER(C.c_size_t)]
    for name in ("set_source_device", "set_target_device"):
        f = getattr(L, "tloam_b200_" + name)
        f.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.tloam_b200_scan_match.argtypes = [vp, dp, dp, C.POINTER(Stats)]
    L.tloam_b200_scan_match_async.argtypes = [vp, dp]
    L.tloam_b200_get_result.argtypes = [vp, dp, C.POINTER(Stats)]
    L.tloam_b200_scan_match_predicted_async.argtypes = [vp]
    L.tloam_b200_scan_match_predicted.argtypes = [vp, dp, C.POINTER(Stats)]
    L.tloam_b200_set_pose_history.argtypes = [vp, dp, dp]
    L.tloam_b200_fitness.argtypes = [vp, dp, dp]
    L.tloam_b200_get_transform.argtypes = [vp, dp]
    L.tloam_b200_get_pose_increment.argtypes = [vp, dp]
    L.tloam_b200_synchronize.argtypes = [vp]
    L.tloam_b200_launch_count.argtypes = [vp]
    L.tloam_b200_launch_count.restype = C.c_longlong
    L.tloam_b200_map_blob_size.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.tloam_b200_map_export.argtypes = [vp, vp, C.c_size_t]
    L.tloam_b200_map_import.argtypes = [vp, vp, C.c_size_t]
    L.tloam_b200_get_map_origin.argtypes = [vp, dp]
    L.tloam_b200_knn.argtypes = [vp, C.c_int, dp, C.c_size_t, C.c_double, C.c_int, ip, dp, ip]
    L.tloam_b200_build_factors.argtypes = [vp, C.c_int, dp, ip, dp, C.c_size_t]
    L.tloam_b200_eval_point_to_point.argtypes = [vp, dp, C.c_size_t, dp, dp, dp, dp, dp, dp]
    L.tloam_b200_eval_point_to_line.argtypes = [vp, dp, C.c_size_t, dp, dp, dp, dp, dp, dp, dp]
    L.tloam_b200_eval_point_to_plane.argtypes = [vp, dp, C.c_size_t, dp, dp, dp, dp, dp, dp, dp]
    L.tloam_b200_se3_exp.argtypes = [vp, dp, dp]
    L.tloam_b200_se3_log.argtypes = [vp, dp, dp]
    L.tloam_b200_se3_plus.argtypes = [vp, dp, dp, dp]
    L.tloam_b200_min_on_boundary_2d.argtypes = [vp, dp, dp, C.c_double, dp]
    L.tloam_b200_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.tloam_b200_host_free.argtypes = [vp]
    L.tloam_b200_set_profiling.argtypes = [vp, C.c_int]
    L.tloam_b200_get_profile.argtypes = [vp, C.POINTER(Profile)]
    L.tloam_b200_set_trace.argtypes = [vp, C.c_int]
    L.tloam_b200_submap_default_config.argtypes = [C.POINTER(SubmapConfig)]
    L.tloam_b200_submap_default_config.restype = None
    L.tloam_b200_submap_init.argtypes = [vp, C.POINTER(SubmapConfig), dp, C.c_size_t, dp, C.c_size_t, dp, C.c_size_t, dp, C.c_size_t]
    L.tloam_b200_submap_update.argtypes = [vp, dp, dp, C.c_size_t, dp, C.c_size_t]
    L.tloam_b200_submap_sizes.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.tloam_b200_submap_download.argtypes = [vp, C.c_int, dp, C.c_size_t]
    L.tloam_b200_voxel_down_sample.argtypes = [vp, dp, C.c_size_t, C.c_double, dp, C.POINTER(C.c_size_t)]
    szp = C.POINTER(C.c_size_t)
    L.tloam_b200_feature_default_config.argtypes = [C.POINTER(FeatureConfig)]
    L.tloam_b200_feature_default_config.restype = None
    L.tloam_b200_extract_planar_sphere.argtypes = [vp, C.POINTER(FeatureConfig), dp, C.c_size_t, szp, szp, szp, szp, szp,
                                                   szp, szp, szp, szp]
    L.tloam_b200_pca_info.argtypes = [vp, C.POINTER(FeatureConfig), dp, C.c_size_t, dp, dp, dp, dp,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.tloam_b200_batch_create.argtypes = [C.POINTER(TlsConfig), C.c_int, C.c_int, C.POINTER(vp)]
    L.tloam_b200_batch_destroy.argtypes = [vp]
    L.tloam_b200_batch_size.argtypes = [vp]
    L.tloam_b200_batch_handle.argtypes = [vp, C.c_int]
    L.tloam_b200_batch_handle.restype = vp
    for name in ("set_target", "set_source"):
        getattr(L, "tloam_b200_batch_" + name).argtypes = [vp, C.POINTER(dp), szp]
        getattr(L, "tloam_b200_batch_" + name + "_device").argtypes = [vp, C.POINTER(vp), szp]
    L.tloam_b200_batch_scan_match.argtypes = [vp, dp, dp, ip]
    L.tloam_b200_batch_scan_match_async.argtypes = [vp, dp]
    L.tloam_b200_batch_get_results.argtypes = [vp, dp, ip, C.POINTER(C.c_float)]
    L.tloam_b200_batch_launch_count.argtypes = [vp]
    L.tloam_b200_batch_launch_count.restype = C.c_longlong
    L.tloam_b200_batch_last_error.argtypes = [vp]
    L.tloam_b200_batch_last_error.restype = C.c_char_p
    L.tloam_b200_batch_set_profiling.argtypes = [vp, C.c_int]
    L.tloam_b200_submap_update_chained.argtypes = [vp, dp, C.c_size_t]
    L.tloam_b200_set_frame_fitness.argtypes = [vp, C.c_int]
    L.tloam_b200_get_frame_fitness.argtypes = [vp, dp, dp]
    L.tloam_b200_set_async_inputs.argtypes = [vp, C.c_int]
    L.tloam_b200_wait_stream.argtypes = [vp, vp]
    L.tloam_b200_dense_check_counters.argtypes = [vp, C.POINTER(C.c_uint)]
    L.tloam_b200_map_layout_bytes.argtypes = [vp, szp, szp]
    L.tloam_b200_map_send_buffer.argtypes = [vp, C.POINTER(vp), szp]
    L.tloam_b200_map_recv_buffer.argtypes = [vp, szp, C.POINTER(vp), szp]
    L.tloam_b200_map_adopt.argtypes = [vp, vp]
    L.tloam_b200_signal_stream.argtypes = [vp, vp]
    L.tloam_b200_ground_default_config.argtypes = [C.POINTER(GroundConfig)]
    L.tloam_b200_ground_default_config.restype = None
    L.tloam_b200_ground_extract.argtypes = [vp, C.POINTER(GroundConfig), dp, C.c_size_t, szp, szp, szp, szp, ip, ip, dp, dp]
    L.tloam_b200_extract_edge.argtypes = [vp, C.c_int, C.c_int, dp, dp, C.c_size_t, szp, szp, szp, szp]
    L.tloam_b200_dcvc_default_config.argtypes = [C.POINTER(DcvcConfig)]
    L.tloam_b200_dcvc_default_config.restype = None
    L.tloam_b200_object_segmentation.argtypes = [vp, C.POINTER(DcvcConfig), dp, C.c_size_t, szp, szp, ip, ip, dp, ip, ip, ip, dp]
    L.tloam_b200_segment_scan.argtypes = [vp, C.POINTER(GroundConfig), C.POINTER(DcvcConfig), C.c_int, dp, C.c_size_t, szp, szp, szp, szp, szp, szp,
                                          ip, ip, dp, ip]
    L.tloam_b200_batch_get_profile.argtypes = [vp, C.POINTER(Profile)]
    _lib = L
    return L
