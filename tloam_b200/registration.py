"""Host-side mirror of the reference's registration interface, over the C-ABI CUDA library.

`LocalRegistration` mirrors tloam::LocalRegistration / RegistrationInterface
(ref: include/tloam/models/registration/registration_interface.hpp:40-48,
      include/tloam/models/registration/registration.hpp:142-165): same method names
(snake_case), same argument meaning, status codes instead of aborts.  It is a thin ctypes shell:
all arithmetic happens in libtloam_b200.so on the GPU.  There is no CPU path.
"""
import ctypes as C

import numpy as np

from . import _lib

CLOUDS = ("edge", "sphere", "planar", "ground")  # ABI order (ref: registration.cpp:233-236)


class RegistrationError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        msg = _lib.load().tloam_b200_status_string(status).decode()
        super().__init__(f"{where}: {msg} (status {status}) {detail}")


def default_config(**overrides):
    """The "TLS:" block defaults (ref: config/mapping/lidar_odometry.yaml:23-39)."""
    cfg = _lib.TlsConfig()
    _lib.load().tloam_b200_default_config(C.byref(cfg))
    for k, v in overrides.items():
        if k == "reinit_dir":
            for i in range(3):
                cfg.reinit_dir[i] = float(v[i])
        else:
            if not hasattr(cfg, k):
                raise KeyError(k)
            setattr(cfg, k, v)
    return cfg


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Frame:
    """Mirror of tloam::Frame (ref: registration_interface.hpp:19-38): the four feature clouds, (n,3) float64.
    scan_cloud is accepted and ignored, as in the reference (registration.cpp:232-239)."""

    def __init__(self, edge_feature, sphere_feature, planar_feature, ground_feature, scan_cloud=None):
        self.edge_feature = edge_feature
        self.sphere_feature = sphere_feature
        self.planar_feature = planar_feature
        self.ground_feature = ground_feature
        self.scan_cloud = scan_cloud

    def clouds(self):
        return [self.edge_feature, self.sphere_feature, self.planar_feature, self.ground_feature]


class LocalRegistration:
    def __init__(self, config=None, device=0, stream=None, _borrowed=None, **overrides):
        self._L = _lib.load()
        self.cfg = config if config is not None else default_config(**overrides)
        self._owned = _borrowed is None
        if _borrowed is None:
            h = C.c_void_p()
            rc = self._L.tloam_b200_create(C.byref(self.cfg), int(device), C.c_void_p(stream or 0), C.byref(h))
            if rc != _lib.OK:
                raise RegistrationError(rc, "tloam_b200_create")
        else:
            h = C.c_void_p(_borrowed)     # a sequence of a BatchRegistration: the batch owns the handle
        self._h = h
        self._keep = []
        self.n_source = [0, 0, 0, 0]

    def close(self):
        if getattr(self, "_h", None):
            if self._owned:
                self._L.tloam_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, where):
        if rc != _lib.OK:
            detail = self._L.tloam_b200_last_error(self._h).decode() if rc == _lib.ERR_CUDA else ""
            raise RegistrationError(rc, where, detail)

    # ---- RegistrationInterface ----
    @staticmethod
    def _host_args(frame):
        clouds = frame.clouds() if isinstance(frame, Frame) else list(frame)
        arrs = [_f64(c).reshape(-1, 3) for c in clouds]
        ptrs = (C.POINTER(C.c_double) * 4)(*[_dp(a) for a in arrs])
        ns = (C.c_size_t * 4)(*[a.shape[0] for a in arrs])
        return arrs, ptrs, ns

    def set_input_source(self, frame):
        arrs, ptrs, ns = self._host_args(frame)
        self._check(self._L.tloam_b200_set_source(self._h, ptrs, ns), "set_input_source")
        self.n_source = [a.shape[0] for a in arrs]
        self._keep_source_host = arrs         # pipelined mode (set_async_inputs): the upload may still be reading them
        return True

    def set_input_target(self, frame):
        arrs, ptrs, ns = self._host_args(frame)
        self._check(self._L.tloam_b200_set_target(self._h, ptrs, ns), "set_input_target")
        return True

    @staticmethod
    def _device_args(tensors):
        ptrs = (C.c_void_p * 4)(*[int(t.data_ptr()) for t in tensors])
        ns = (C.c_size_t * 4)(*[int(t.shape[0]) for t in tensors])
        return ptrs, ns

    def _order_after_producer(self, tensors, producer_stream):
        """The tensors are read IN PLACE by kernels on the handle's stream: order that stream behind the stream that
        produced them (default: torch's current stream on the tensors' device) -- a device-side wait, the host does not
        block -- and tell torch's caching allocator that another stream uses the memory."""
        if producer_stream is None:
            import torch
            producer_stream = torch.cuda.current_stream(tensors[0].device).cuda_stream
        self._check(self._L.tloam_b200_wait_stream(self._h, C.c_void_p(int(producer_stream))), "wait_stream")

    def set_input_source_device(self, tensors, producer_stream=None):
        """tensors: 4 CUDA float64 (n,3) contiguous torch tensors on this handle's device."""
        ptrs, ns = self._device_args(tensors)
        self._order_after_producer(tensors, producer_stream)
        self._check(self._L.tloam_b200_set_source_device(self._h, ptrs, ns), "set_input_source_device")
        self.n_source = [int(t.shape[0]) for t in tensors]
        self._keep_source = list(tensors)     # read in place by kernels still in flight: keep them alive until replaced
        return True

    def set_input_target_device(self, tensors, producer_stream=None):
        ptrs, ns = self._device_args(tensors)
        self._order_after_producer(tensors, producer_stream)
        self._check(self._L.tloam_b200_set_target_device(self._h, ptrs, ns), "set_input_target_device")
        self._keep_target = list(tensors)     # idem
        return True

    def scan_matching(self, predict_pose, want_stats=False):
        """predict_pose: 4x4 numpy (row-major view of the Isometry3d). Returns result 4x4 [, Stats]."""
        p = _f64(np.asarray(predict_pose).T).reshape(16)
        out = np.zeros(16)
        st = _lib.Stats() if want_stats else None
        rc = self._L.tloam_b200_scan_match(self._h, _dp(p), _dp(out), C.byref(st) if st is not None else None)
        self._check(rc, "scan_matching")
        T = out.reshape(4, 4).T.copy()
        return (T, st) if want_stats else T

    def scan_matching_predicted(self, want_stats=False):
        """(f)-3: scanMatching with the constant-velocity prediction of FrontEnd::updateLidarOdometry
        (ref: front_end.cpp:329-330) computed on the device from the two last results."""
        out = np.zeros(16)
        st = _lib.Stats() if want_stats else None
        rc = self._L.tloam_b200_scan_match_predicted(self._h, _dp(out), C.byref(st) if st is not None else None)
        self._check(rc, "scan_matching_predicted")
        T = out.reshape(4, 4).T.copy()
        return (T, st) if want_stats else T

    def set_pose_history(self, last_pose, curr_pose):
        a = _f64(np.asarray(last_pose).T).reshape(16)
        b = _f64(np.asarray(curr_pose).T).reshape(16)
        self._check(self._L.tloam_b200_set_pose_history(self._h, _dp(a), _dp(b)), "set_pose_history")

    def scan_matching_async(self, predict_pose):
        p = _f64(np.asarray(predict_pose).T).reshape(16)
        self._check(self._L.tloam_b200_scan_match_async(self._h, _dp(p)), "scan_matching_async")

    def get_result(self, want_stats=False):
        out = np.zeros(16)
        st = _lib.Stats() if want_stats else None
        rc = self._L.tloam_b200_get_result(self._h, _dp(out), C.byref(st) if st is not None else None)
        self._check(rc, "get_result")
        T = out.reshape(4, 4).T.copy()
        return (T, st) if want_stats else T

    def get_fitness_score(self):
        a, b = C.c_double(0), C.c_double(0)
        self._check(self._L.tloam_b200_fitness(self._h, C.byref(a), C.byref(b)), "get_fitness_score")
        return a.value, b.value

    def get_transform(self):
        out = np.zeros(16)
        self._check(self._L.tloam_b200_get_transform(self._h, _dp(out)), "get_transform")
        return out.reshape(4, 4).T.copy()

    def get_pose_increment(self):
        out = np.zeros(16)
        self._check(self._L.tloam_b200_get_pose_increment(self._h, _dp(out)), "get_pose_increment")
        return out.reshape(4, 4).T.copy()

    def dense_check_counters(self):
        out = (C.c_uint * 16)()
        self._check(self._L.tloam_b200_dense_check_counters(self._h, out), "dense_check_counters")
        return [int(v) for v in out]

    def synchronize(self):
        self._check(self._L.tloam_b200_synchronize(self._h), "synchronize")

    def launch_count(self):
        return int(self._L.tloam_b200_launch_count(self._h))

    def set_profiling(self, on):
        self._check(self._L.tloam_b200_set_profiling(self._h, 1 if on else 0), "set_profiling")

    def get_profile(self):
        """dict kernel-class -> (launches, total_ms), from CUDA events around every launch."""
        p = _lib.Profile()
        self._check(self._L.tloam_b200_get_profile(self._h, C.byref(p)), "get_profile")
        out = {k: (int(p.launches[i]), float(p.total_ms[i])) for i, k in enumerate(_lib.KERNEL_CLASSES)}
        self.last_dbg = [int(v) for v in p.dbg]
        return out

    # ---- device-side submap maintenance ((f)-1, mirrors FrontEnd::updateSubmap) ----
    def submap_init(self, edge, ground_raw, planar_sub, sphere_sub, **cfg_overrides):
        cfg = _lib.SubmapConfig()
        self._L.tloam_b200_submap_default_config(C.byref(cfg))
        for k, v in cfg_overrides.items():
            setattr(cfg, k, v)
        a = [_f64(x).reshape(-1, 3) for x in (edge, ground_raw, planar_sub, sphere_sub)]
        self._check(self._L.tloam_b200_submap_init(self._h, C.byref(cfg), _dp(a[0]), a[0].shape[0], _dp(a[1]), a[1].shape[0],
                                                   _dp(a[2]), a[2].shape[0], _dp(a[3]), a[3].shape[0]), "submap_init")

    def submap_update(self, pose, planar_sub, sphere_sub=None):
        p = _f64(np.asarray(pose).T).reshape(16)
        a = _f64(planar_sub).reshape(-1, 3)
        b = _f64(sphere_sub if sphere_sub is not None else np.zeros((0, 3))).reshape(-1, 3)
        self._check(self._L.tloam_b200_submap_update(self._h, _dp(p), _dp(a), a.shape[0], _dp(b), b.shape[0]), "submap_update")

    def submap_update_chained(self, planar_sub):
        """FrontEnd::updateSubmap with the pose of the frame that was just enqueued, read on the device."""
        a = _f64(planar_sub).reshape(-1, 3)
        self._keep_planar = a
        self._check(self._L.tloam_b200_submap_update_chained(self._h, _dp(a), a.shape[0]), "submap_update_chained")

    def set_async_inputs(self, on):
        self._check(self._L.tloam_b200_set_async_inputs(self._h, 1 if on else 0), "set_async_inputs")

    def set_frame_fitness(self, on):
        self._check(self._L.tloam_b200_set_frame_fitness(self._h, 1 if on else 0), "set_frame_fitness")

    def get_frame_fitness(self):
        a, b = C.c_double(0), C.c_double(0)
        self._check(self._L.tloam_b200_get_frame_fitness(self._h, C.byref(a), C.byref(b)), "get_frame_fitness")
        return a.value, b.value

    def scan_matching_predicted_async(self):
        self._check(self._L.tloam_b200_scan_match_predicted_async(self._h), "scan_matching_predicted_async")

    def submap_cloud(self, cloud):
        n = (C.c_size_t * 4)()
        self._check(self._L.tloam_b200_submap_sizes(self._h, n), "submap_sizes")
        out = np.zeros((n[cloud], 3))
        self._check(self._L.tloam_b200_submap_download(self._h, cloud, _dp(out), n[cloud]), "submap_download")
        return out

    def voxel_down_sample(self, pts, voxel):
        a = _f64(pts).reshape(-1, 3)
        out = np.zeros_like(a)
        n = C.c_size_t(0)
        self._check(self._L.tloam_b200_voxel_down_sample(self._h, _dp(a), a.shape[0], float(voxel), _dp(out), C.byref(n)), "voxel_down_sample")
        return out[:n.value].copy()

    # ---- "next" row (f)-2: PCA feature extraction (ref: feature_extract.cpp:47-122, 133-197) ----
    def _feature_config(self, overrides):
        c = _lib.FeatureConfig()
        self._L.tloam_b200_feature_default_config(C.byref(c))
        for k, v in overrides.items():
            setattr(c, k, v)
        return c

    def extract_planar_sphere(self, general_cloud, **overrides):
        """featureExtract::extractPlanarSphere on the device.  Returns (planar_scan_index, planar_submap_index,
        sphere_scan_index, sphere_submap_index, sphere_candidates); the two sphere lists hold ranks (reference quirk),
        sphere_candidates the point indices those ranks refer to."""
        a = _f64(general_cloud).reshape(-1, 3)
        c = self._feature_config(overrides)
        n = a.shape[0]
        bufs = [np.zeros(max(n, 1), dtype=np.uintp) for _ in range(5)]
        cnt = [C.c_size_t(0) for _ in range(4)]
        szp = C.POINTER(C.c_size_t)
        self._check(self._L.tloam_b200_extract_planar_sphere(
            self._h, C.byref(c), _dp(a), n, bufs[0].ctypes.data_as(szp), C.byref(cnt[0]), bufs[1].ctypes.data_as(szp),
            C.byref(cnt[1]), bufs[2].ctypes.data_as(szp), C.byref(cnt[2]), bufs[3].ctypes.data_as(szp), C.byref(cnt[3]),
            bufs[4].ctypes.data_as(szp)), "extract_planar_sphere")
        return (bufs[0][:cnt[0].value].copy(), bufs[1][:cnt[1].value].copy(), bufs[2][:cnt[2].value].copy(),
                bufs[3][:cnt[3].value].copy(), bufs[4][:cnt[3].value].copy())

    def pca_info(self, general_cloud, **overrides):
        """featureExtract::calculatePCAInfo on the device: dict of cvr, flatness, sphericity, normal (n,3), num_sum, neigh (n,K)."""
        a = _f64(general_cloud).reshape(-1, 3)
        c = self._feature_config(overrides)
        n = a.shape[0]
        out = {"cvr": np.zeros(n), "flatness": np.zeros(n), "sphericity": np.zeros(n), "normal": np.zeros((n, 3)),
               "num_sum": np.zeros(n, dtype=np.int32), "neigh": np.full((n, c.K), -1, dtype=np.int32)}
        ip = C.POINTER(C.c_int)
        self._check(self._L.tloam_b200_pca_info(self._h, C.byref(c), _dp(a), n, _dp(out["cvr"]), _dp(out["flatness"]),
                                                _dp(out["sphericity"]), _dp(out["normal"]),
                                                out["num_sum"].ctypes.data_as(ip), out["neigh"].ctypes.data_as(ip)),
                    "pca_info")
        return out

    # ---- "next" row (f)-4, first part: ground extraction (ref: segmentation.cpp:738-770) ----
    def ground_extract(self, scan, **overrides):
        """Segmentation::groundRemove on the device.  Returns dict(ground, object, beam, region, height_threshold, planes)."""
        a = _f64(scan).reshape(-1, 3)
        n = a.shape[0]
        c = _lib.GroundConfig()
        self._L.tloam_b200_ground_default_config(C.byref(c))
        for k, v in overrides.items():
            setattr(c, k, v)
        g = np.zeros(max(n, 1), dtype=np.uintp)
        o = np.zeros(max(n, 1), dtype=np.uintp)
        ng, no = C.c_size_t(0), C.c_size_t(0)
        beam = np.zeros(max(n, 1), dtype=np.int32)
        region = np.zeros(max(n, 1), dtype=np.int32)
        thr = C.c_double(0)
        planes = np.zeros((12, 8, 4))
        szp, ip = C.POINTER(C.c_size_t), C.POINTER(C.c_int)
        self._check(self._L.tloam_b200_ground_extract(self._h, C.byref(c), _dp(a), n, g.ctypes.data_as(szp), C.byref(ng),
                                                      o.ctypes.data_as(szp), C.byref(no), beam.ctypes.data_as(ip),
                                                      region.ctypes.data_as(ip), C.byref(thr), _dp(planes)), "ground_extract")
        return dict(ground=g[:ng.value].copy(), object=o[:no.value].copy(), beam=beam[:n].copy(), region=region[:n].copy(),
                    height_threshold=thr.value, planes=planes)

    # ---- "next" row (f)-4, second part: edge extraction (ref: segmentation.cpp:1144-1304) ----
    def extract_edge(self, points, intensity, sensor_model=64, ring_min_num=16):
        """Segmentation::extractEdgePoint on the device.  intensity holds the beam id of every point (as groundRemove leaves
        it).  Returns dict(edge, non_edge): index lists into the input in the reference's append order."""
        a = _f64(points).reshape(-1, 3)
        it = _f64(intensity).reshape(-1)
        n = a.shape[0]
        if it.shape[0] != n:
            raise ValueError("one intensity (beam id) per point")
        e = np.zeros(max(n, 1), dtype=np.uintp)
        o = np.zeros(max(n, 1), dtype=np.uintp)
        ne, no = C.c_size_t(0), C.c_size_t(0)
        szp = C.POINTER(C.c_size_t)
        self._check(self._L.tloam_b200_extract_edge(self._h, sensor_model, ring_min_num, _dp(a), _dp(it), n, e.ctypes.data_as(szp),
                                                    C.byref(ne), o.ctypes.data_as(szp), C.byref(no)), "extract_edge")
        return dict(edge=e[:ne.value].copy(), non_edge=o[:no.value].copy())

    # ---- "next" row (f)-4, third part: object segmentation = DCVC (ref: segmentation.cpp:772-1112) ----
    def object_segmentation(self, points, details=True, **overrides):
        """Segmentation::objectSegmentation on the device.  Returns dict(segmented, sizes, boxes[, root, cluster, voxel,
        polar, ext]): the segmented scan as indices (cluster after cluster), cluster sizes / boxes, and with `details` per
        point the smallest index of its DCVC class, its 1-based cluster number (0 = filtered), the reference's voxel index
        and the polar triple (four more downloads; the C++ shim does not ask for them)."""
        a = _f64(points).reshape(-1, 3)
        n = a.shape[0]
        c = _lib.DcvcConfig()
        self._L.tloam_b200_dcvc_default_config(C.byref(c))
        for k, v in overrides.items():
            setattr(c, k, v)
        m = max(n, 1)
        seg = np.zeros(m, dtype=np.uintp)
        sizes = np.zeros(m, dtype=np.int32)
        boxes = np.zeros((m, 6))
        nseg, ncl = C.c_size_t(0), C.c_int(0)
        szp, ip = C.POINTER(C.c_size_t), C.POINTER(C.c_int)
        if details:
            root, cluster, voxel = (np.zeros(m, dtype=np.int32) for _ in range(3))
            polar = np.zeros(3 * m + 4)
            extra = (root.ctypes.data_as(ip), cluster.ctypes.data_as(ip), voxel.ctypes.data_as(ip), _dp(polar))
        else:
            extra = (None, None, None, None)
        self._check(self._L.tloam_b200_object_segmentation(self._h, C.byref(c), _dp(a), n, seg.ctypes.data_as(szp), C.byref(nseg),
                                                           C.byref(ncl), sizes.ctypes.data_as(ip), _dp(boxes), *extra),
                    "object_segmentation")
        k = ncl.value
        out = dict(segmented=seg[:nseg.value].copy(), sizes=sizes[:k].copy(), boxes=boxes[:k].copy())
        if details:
            out.update(root=root[:n].copy(), cluster=cluster[:n].copy(), voxel=voxel[:n].copy(),
                       polar=polar[:3 * n].reshape(-1, 3).copy(), ext=polar[3 * n:3 * n + 4].copy())
        return out

    # ---- "next" row (f)-4: the three segmentation steps as one call (ref: segmentation.cpp:47-66) ----
    def segment_scan(self, scan, ring_min_num=131, ground=None, dcvc=None):
        """groundRemove -> objectSegmentation -> extractEdgePoint chained on the device (one upload, one download).  Returns
        dict(ground, edge, general, sizes, boxes, beam): index lists into `scan`, the cluster table, the beam estimate per point.  ground / dcvc: dicts of
        configuration overrides."""
        a = _f64(scan).reshape(-1, 3)
        n = a.shape[0]
        gc, dc = _lib.GroundConfig(), _lib.DcvcConfig()
        self._L.tloam_b200_ground_default_config(C.byref(gc))
        self._L.tloam_b200_dcvc_default_config(C.byref(dc))
        for k, v in (ground or {}).items():
            setattr(gc, k, v)
        for k, v in (dcvc or {}).items():
            setattr(dc, k, v)
        m = max(n, 1)
        g, e, o = (np.zeros(m, dtype=np.uintp) for _ in range(3))
        ng, ne, no = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        ncl = C.c_int(0)
        sizes = np.zeros(m, dtype=np.int32)
        beam = np.zeros(m, dtype=np.int32)
        boxes = np.zeros((m, 6))
        szp, ip = C.POINTER(C.c_size_t), C.POINTER(C.c_int)
        self._check(self._L.tloam_b200_segment_scan(self._h, C.byref(gc), C.byref(dc), ring_min_num, _dp(a), n, g.ctypes.data_as(szp), C.byref(ng),
                                                    e.ctypes.data_as(szp), C.byref(ne), o.ctypes.data_as(szp), C.byref(no), C.byref(ncl),
                                                    sizes.ctypes.data_as(ip), _dp(boxes), beam.ctypes.data_as(ip)), "segment_scan")
        k = ncl.value
        return dict(ground=g[:ng.value].copy(), edge=e[:ne.value].copy(), general=o[:no.value].copy(), sizes=sizes[:k].copy(),
                    boxes=boxes[:k].copy(), beam=beam[:n].copy())

    # ---- shared map (multi-GPU) ----
    def map_blob_size(self):
        n = C.c_size_t(0)
        self._check(self._L.tloam_b200_map_blob_size(self._h, C.byref(n)), "map_blob_size")
        return n.value

    def map_export(self, dev_ptr, nbytes):
        self._check(self._L.tloam_b200_map_export(self._h, C.c_void_p(dev_ptr), nbytes), "map_export")

    def map_import(self, dev_ptr, nbytes):
        self._check(self._L.tloam_b200_map_import(self._h, C.c_void_p(dev_ptr), nbytes), "map_import")

    def map_send_buffer(self):
        p, n = C.c_void_p(), C.c_size_t(0)
        self._check(self._L.tloam_b200_map_send_buffer(self._h, C.byref(p), C.byref(n)), "map_send_buffer")
        return p.value, n.value

    def map_recv_buffer(self, n_map):
        ns = (C.c_size_t * 4)(*[int(v) for v in n_map])
        p, n = C.c_void_p(), C.c_size_t(0)
        self._check(self._L.tloam_b200_map_recv_buffer(self._h, ns, C.byref(p), C.byref(n)), "map_recv_buffer")
        return p.value, n.value

    def map_adopt(self, producer_stream):
        self._check(self._L.tloam_b200_map_adopt(self._h, C.c_void_p(int(producer_stream))), "map_adopt")

    def signal_stream(self, consumer_stream):
        self._check(self._L.tloam_b200_signal_stream(self._h, C.c_void_p(int(consumer_stream))), "signal_stream")

    def map_origin(self):
        o = np.zeros(3)
        self._check(self._L.tloam_b200_get_map_origin(self._h, _dp(o)), "map_origin")
        return o

    # ---- piecewise (parity tests) ----
    def knn(self, cloud, queries, radius, k):
        q = _f64(queries).reshape(-1, 3)
        nq = q.shape[0]
        idx = np.full((nq, k), -1, dtype=np.int32)
        d2 = np.full((nq, k), np.inf)
        cnt = np.zeros(nq, dtype=np.int32)
        rc = self._L.tloam_b200_knn(self._h, cloud, _dp(q), nq, float(radius), int(k),
                                    idx.ctypes.data_as(C.POINTER(C.c_int)), _dp(d2),
                                    cnt.ctypes.data_as(C.POINTER(C.c_int)))
        self._check(rc, "knn")
        return idx, d2, cnt

    def build_factors(self, cloud, x):
        n = self.n_source[cloud]
        valid = np.zeros(n, dtype=np.int32)
        prim = np.zeros((n, 6))
        x = _f64(x)
        rc = self._L.tloam_b200_build_factors(self._h, cloud, _dp(x), valid.ctypes.data_as(C.POINTER(C.c_int)),
                                              _dp(prim), n)
        self._check(rc, "build_factors")
        return valid, prim

    def eval_point_to_point(self, x, p, q, w):
        x, p, q, w = _f64(x), _f64(p).reshape(-1, 3), _f64(q).reshape(-1, 3), _f64(w).reshape(-1)
        m = p.shape[0]
        r, J, c = np.zeros((m, 3)), np.zeros((m, 3, 6)), np.zeros(m)
        self._check(self._L.tloam_b200_eval_point_to_point(self._h, _dp(x), m, _dp(p), _dp(q), _dp(w), _dp(r), _dp(J), _dp(c)),
                    "eval_point_to_point")
        return r, J, c

    def eval_point_to_line(self, x, p, a, b, w):
        x, p, a, b, w = _f64(x), _f64(p).reshape(-1, 3), _f64(a).reshape(-1, 3), _f64(b).reshape(-1, 3), _f64(w).reshape(-1)
        m = p.shape[0]
        r, J, c = np.zeros((m, 3)), np.zeros((m, 3, 6)), np.zeros(m)
        self._check(self._L.tloam_b200_eval_point_to_line(self._h, _dp(x), m, _dp(p), _dp(a), _dp(b), _dp(w), _dp(r), _dp(J), _dp(c)),
                    "eval_point_to_line")
        return r, J, c

    def eval_point_to_plane(self, x, p, n, d, w):
        x, p, n, d, w = _f64(x), _f64(p).reshape(-1, 3), _f64(n).reshape(-1, 3), _f64(d).reshape(-1), _f64(w).reshape(-1)
        m = p.shape[0]
        r, J, c = np.zeros((m, 1)), np.zeros((m, 1, 6)), np.zeros(m)
        self._check(self._L.tloam_b200_eval_point_to_plane(self._h, _dp(x), m, _dp(p), _dp(n), _dp(d), _dp(w), _dp(r), _dp(J), _dp(c)),
                    "eval_point_to_plane")
        return r, J, c

    def se3_exp(self, a):
        a = _f64(a)
        T = np.zeros(16)
        self._check(self._L.tloam_b200_se3_exp(self._h, _dp(a), _dp(T)), "se3_exp")
        return T.reshape(4, 4).T.copy()

    def se3_log(self, T):
        t = _f64(np.asarray(T).T).reshape(16)
        a = np.zeros(6)
        self._check(self._L.tloam_b200_se3_log(self._h, _dp(t), _dp(a)), "se3_log")
        return a

    def min_on_boundary_2d(self, B, g, radius):
        B, g = _f64(B).reshape(4), _f64(g).reshape(2)
        y = np.zeros(2)
        self._check(self._L.tloam_b200_min_on_boundary_2d(self._h, _dp(B), _dp(g), float(radius), _dp(y)), "min_on_boundary_2d")
        return y

    def se3_plus(self, x, d):
        x, d = _f64(x), _f64(d)
        out = np.zeros(6)
        self._check(self._L.tloam_b200_se3_plus(self._h, _dp(x), _dp(d), _dp(out)), "se3_plus")
        return out


class BatchRegistration:
    """S independent sequences registered together (tloam_b200_batch_*): one launch sequence per batch frame.
    `self.seq[i]` is a LocalRegistration view of sequence i (set_input_*, submap_*, get_transform ... per sequence);
    scan_matching() steps all of them.  Per-sequence poses are bit-identical to the un-batched path."""

    def __init__(self, S, config=None, device=0, **overrides):
        self._L = _lib.load()
        self.cfg = config if config is not None else default_config(**overrides)
        self.S = int(S)
        b = C.c_void_p()
        rc = self._L.tloam_b200_batch_create(C.byref(self.cfg), int(device), self.S, C.byref(b))
        if rc != _lib.OK:
            raise RegistrationError(rc, "tloam_b200_batch_create")
        self._b = b
        self.seq = [LocalRegistration(config=self.cfg, _borrowed=self._L.tloam_b200_batch_handle(b, i)) for i in range(self.S)]
        self._keep = {}

    def close(self):
        if getattr(self, "_b", None):
            for r in self.seq:
                r.close()
            self._L.tloam_b200_batch_destroy(self._b)
            self._b = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, where):
        if rc != _lib.OK:
            raise RegistrationError(rc, where, self._L.tloam_b200_batch_last_error(self._b).decode())

    def pack_device(self, per_seq_tensors):
        """Marshal S x 4 CUDA tensors once (outside a timed loop); pass the result to set_input_*_device."""
        flat = [t for seq in per_seq_tensors for t in seq]
        assert len(flat) == 4 * self.S
        ptrs = (C.c_void_p * (4 * self.S))(*[int(t.data_ptr()) for t in flat])
        ns = (C.c_size_t * (4 * self.S))(*[int(t.shape[0]) for t in flat])
        return ptrs, ns, flat

    def pack_host(self, per_seq_clouds):
        flat = [_f64(c).reshape(-1, 3) for seq in per_seq_clouds for c in seq]
        assert len(flat) == 4 * self.S
        ptrs = (C.POINTER(C.c_double) * (4 * self.S))(*[_dp(a) for a in flat])
        ns = (C.c_size_t * (4 * self.S))(*[a.shape[0] for a in flat])
        return ptrs, ns, flat

    def _set(self, fn, packed, what):
        ptrs, ns, flat = packed
        self._check(fn(self._b, ptrs, ns), what)
        self._keep[what] = flat            # device inputs are read in place: keep them alive until replaced
        if what.startswith("source"):
            for i, r in enumerate(self.seq):
                r.n_source = [int(ns[4 * i + c]) for c in range(4)]

    def set_input_target_device(self, packed):
        self._set(self._L.tloam_b200_batch_set_target_device, packed, "target_device")

    def set_input_source_device(self, packed):
        self._set(self._L.tloam_b200_batch_set_source_device, packed, "source_device")

    def set_input_target(self, packed):
        self._set(self._L.tloam_b200_batch_set_target, packed, "target")

    def set_input_source(self, packed):
        self._set(self._L.tloam_b200_batch_set_source, packed, "source")

    def scan_matching(self, predicts=None):
        """predicts: (S,4,4) or None (device-side constant-velocity prediction).  Returns (S,4,4) poses, statuses."""
        out = np.zeros((self.S, 16))
        st = np.zeros(self.S, dtype=np.int32)
        p = None
        if predicts is not None:
            p = _f64(np.transpose(np.asarray(predicts).reshape(self.S, 4, 4), (0, 2, 1))).reshape(-1)
        rc = self._L.tloam_b200_batch_scan_match(self._b, _dp(p) if p is not None else None, _dp(out),
                                                 st.ctypes.data_as(C.POINTER(C.c_int)))
        self._check(rc, "batch_scan_matching")
        return np.transpose(out.reshape(self.S, 4, 4), (0, 2, 1)).copy(), st

    def scan_matching_async(self, predicts=None):
        """Enqueue only (several BatchRegistration objects can be in flight on one GPU: their serial solver tails then
        overlap with the others' parallel phases); fetch with get_results()."""
        p = None
        if predicts is not None:
            p = _f64(np.transpose(np.asarray(predicts).reshape(self.S, 4, 4), (0, 2, 1))).reshape(-1)
            self._keep["predicts"] = p
        self._check(self._L.tloam_b200_batch_scan_match_async(self._b, _dp(p) if p is not None else None), "batch_scan_matching_async")

    def get_results(self):
        out = np.zeros((self.S, 16))
        st = np.zeros(self.S, dtype=np.int32)
        rc = self._L.tloam_b200_batch_get_results(self._b, _dp(out), st.ctypes.data_as(C.POINTER(C.c_int)), None)
        self._check(rc, "batch_get_results")
        return np.transpose(out.reshape(self.S, 4, 4), (0, 2, 1)).copy(), st

    def launch_count(self):
        return int(self._L.tloam_b200_batch_launch_count(self._b))

    def set_profiling(self, on):
        self._check(self._L.tloam_b200_batch_set_profiling(self._b, 1 if on else 0), "batch_set_profiling")

    def get_profile(self):
        p = _lib.Profile()
        self._check(self._L.tloam_b200_batch_get_profile(self._b, C.byref(p)), "batch_get_profile")
        return {k: (int(p.launches[i]), float(p.total_ms[i])) for i, k in enumerate(_lib.KERNEL_CLASSES)}
