"""ctypes access to the C++ host adapters through tests/cpp/host_shim.cpp (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "tests", "cpp", "_build", "libhost_shim.so")
ITER_LOG_DT = np.dtype([("level", "<i4"), ("iter", "<i4"), ("n_pts", "<i4"), ("flags", "<i4"), ("chi2", "<f8"), ("x", "<f8", 6)])
_lib = None


def build():
    from dsdtm_b200 import build as gb
    from dsdtm_b200.host import build as hb
    gb.build()
    host = hb.build()
    src = os.path.join(ROOT, "tests", "cpp", "host_shim.cpp")
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    deps = [src, host, os.path.join(ROOT, "dsdtm_b200", "host", "dsdtm_host.h")]
    if not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
        libdir = os.path.join(ROOT, "dsdtm_b200", "lib")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                               "-o", _SO, src, host, os.path.join(libdir, "libdsdtm_gpu.so"), "-Wl,-rpath,$ORIGIN/../../../dsdtm_b200/lib"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.hs_last_error.restype = C.c_char_p
        L.hs_config_get.restype = C.c_double
        L.hs_config_get.argtypes = [C.c_char_p]
        L.hs_config_get_int.argtypes = [C.c_char_p]
        for f in ("hs_camera_new", "hs_frame_new", "hs_keyframe_new"):
            getattr(L, f).restype = C.c_void_p
        L.hs_frame_new.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.hs_frame_free.argtypes = [C.c_void_p]
        L.hs_frame_set_pose.argtypes = [C.c_void_p, C.c_void_p]
        L.hs_frame_get_pose.argtypes = [C.c_void_p, C.c_void_p]
        L.hs_frame_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hs_frame_n_features.argtypes = [C.c_void_p]
        L.hs_frame_get_features.argtypes = [C.c_void_p] * 4
        L.hs_frame_mask.argtypes = [C.c_void_p, C.c_void_p]
        L.hs_detect.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.hs_frame_attach_points.argtypes = [C.c_void_p] * 3
        L.hs_sparse_align_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.hs_keyframe_new.argtypes = [C.c_void_p]
        L.hs_search_local_points.argtypes = [C.c_void_p] * 5
        L.hs_mappoint_found.argtypes = [C.c_int]
        L.hs_mappoint_pose.argtypes = [C.c_int, C.c_void_p]
        L.hs_mappoint_increase_found.argtypes = [C.c_int, C.c_int]
        L.hs_set_use_store.argtypes = [C.c_int]
        L.hs_set_speculate.argtypes = [C.c_int]
        L.hs_mappoint_set_bad.argtypes = [C.c_int, C.c_int]
        L.hs_mappoint_set_pose.argtypes = [C.c_int, C.c_void_p]
        L.hs_search_local_points_multi.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.hs_frame_attach_points_from.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hs_frame_feature_mp_ids.argtypes = [C.c_void_p, C.c_void_p]
        L.hs_align2d_single.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.hs_align2d_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.hs_frame_keyframe_lift.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        L.hs_warp_affine_single.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hs_circle.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int]
        L.hs_map_add_keyframe.argtypes = [C.c_void_p]
        L.hs_map_mark_moved.argtypes = [C.c_void_p]
        L.hs_keyframe_set_pose.argtypes = [C.c_void_p, C.c_void_p]
        L.hs_track_local_map.argtypes = [C.c_void_p] * 5
        L.hs_pose_optimization.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.hs_frame_set_feature.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def configure(cam, levels=5, cell=15, max_fts=300, min_fts=15, min_dist=15, max_frames=16, dist=None):
    L = lib()
    L.hs_reset()
    if dist is not None:
        for k, v in zip(("Camera.k1", "Camera.k2", "Camera.p1", "Camera.p2", "Camera.k3"), dist):
            L.hs_config_set(k.encode(), repr(float(v)).encode())
    for k, v in (("Camera.width", cam["width"]), ("Camera.height", cam["height"]), ("Camera.fx", cam["fx"]), ("Camera.fy", cam["fy"]),
                 ("Camera.cx", cam["cx"]), ("Camera.cy", cam["cy"]), ("Camera.f", cam["f"]), ("Camera.MaxPyraLevels", levels),
                 ("Camera.MinPyraLevels", 0), ("Camera.CellSize", cell), ("Camera.Max_fts", max_fts), ("Camera.Max_tkfts", 200),
                 ("Camera.Min_fts", min_fts), ("Camera.Min_dist", min_dist), ("Gpu.MaxFrames", max_frames)):
        L.hs_config_set(k.encode(), repr(v).encode())
    return L.hs_camera_new()


class HFrame:
    def __init__(self, cam_h, img, pose):
        L = lib()
        img = np.ascontiguousarray(img, np.uint8)
        self.h = L.hs_frame_new(cam_h, _p(img), img.shape[1], img.shape[0], _p(np.ascontiguousarray(pose, np.float64)))
        if not self.h:
            raise RuntimeError(L.hs_last_error().decode())
        self.shape = img.shape

    def level(self, l, shape):
        out = np.empty(shape, np.uint8)
        lib().hs_frame_level(self.h, l, _p(out))
        return out

    def features(self):
        n = lib().hs_frame_n_features(self.h)
        px = np.empty((n, 2), np.float32); lv = np.empty(n, np.int32); ini = np.empty(n, np.int32)
        lib().hs_frame_get_features(self.h, _p(px), _p(lv), _p(ini))
        return px, lv, ini

    def detect(self, thr=5.0, use_existing=False):
        n = lib().hs_detect(self.h, float(thr), int(use_existing))
        if n < 0:
            raise RuntimeError(lib().hs_last_error().decode())
        return n

    def attach_points(self, pts, has):
        lib().hs_frame_attach_points(self.h, _p(np.ascontiguousarray(pts, np.float64)), _p(np.ascontiguousarray(has, np.uint8)))

    def attach_points_from(self, start, pts, has):
        lib().hs_frame_attach_points_from(self.h, int(start), _p(np.ascontiguousarray(pts, np.float64)), _p(np.ascontiguousarray(has, np.uint8)))

    def keyframe_lift(self, depth_u16, depth_scale):
        """SetDepth + UndistortFeatures + Get_FeatureDetph + UnProject -> (n x 9 table, n x 3 normals), see host_shim.cpp"""
        n = lib().hs_frame_n_features(self.h)
        out = np.zeros((n, 9)); nrm = np.zeros((n, 3))
        d = np.ascontiguousarray(depth_u16, np.uint16)
        if lib().hs_frame_keyframe_lift(self.h, _p(d), C.c_float(depth_scale), _p(out), _p(nrm)) < 0:
            raise RuntimeError(lib().hs_last_error().decode())
        return out, nrm

    def mp_ids(self):
        n = lib().hs_frame_n_features(self.h)
        ids = np.empty(n, np.int32)
        lib().hs_frame_feature_mp_ids(self.h, _p(ids))
        return ids

    def pose(self):
        p = np.empty(7); lib().hs_frame_get_pose(self.h, _p(p)); return p

    def pose_optimization(self):
        """Optimizer::PoseOptimization(frame, 10) as Tracking calls it -> (pose, summary record, residual norms)"""
        n = lib().hs_frame_n_features(self.h)
        pose = np.empty(7); res = np.zeros(max(n, 1))
        summ = np.zeros(1, np.dtype([("iterations", "<i4"), ("termination", "<i4"), ("n_successful", "<i4"), ("n_obs", "<i4"),
                                     ("initial_cost", "<f8"), ("final_cost", "<f8")]))
        nb = lib().hs_pose_optimization(self.h, _p(pose), _p(summ), _p(res), n)
        if nb < 0:
            raise RuntimeError(lib().hs_last_error().decode())
        return pose, summ[0], res[:nb]

    def set_pose(self, p):
        lib().hs_frame_set_pose(self.h, _p(np.ascontiguousarray(p, np.float64)))

    def mask(self):
        m = np.zeros(self.shape, np.uint8); lib().hs_frame_mask(self.h, _p(m)); return m

    def free(self):
        if self.h:
            lib().hs_frame_free(self.h); self.h = None


def search_local_points_multi(cam_h, cur, kf_handles):
    arr = (C.c_void_p * len(kf_handles))(*kf_handles)
    nrep = C.c_int(0)
    m = lib().hs_search_local_points_multi(cam_h, cur.h, arr, len(kf_handles), C.byref(nrep))
    if m < 0:
        raise RuntimeError(lib().hs_last_error().decode())
    return m, nrep.value


def track_local_map(cam_h, cur):
    """Tracking::UpdateLocalMap (device-side close-key-frame selection over the whole Map) + SearchLocalPoints
    -> (matches, rows of mvpLocalKeyFrames in rank order, reprojected points)"""
    local = np.full(16, -1, np.int32); n_local = C.c_int(0); nrep = C.c_int(0)
    m = lib().hs_track_local_map(cam_h, cur.h, _p(local), C.byref(n_local), C.byref(nrep))
    if m < 0:
        raise RuntimeError(lib().hs_last_error().decode())
    return m, local[:n_local.value].copy(), nrep.value


def sparse_align_run(maxl, minl, iters, cur, ref, want_log=True):
    pose = np.empty(7); log = np.zeros(256 if want_log else 0, ITER_LOG_DT); n_log = C.c_int(0)
    n = lib().hs_sparse_align_run(maxl, minl, iters, cur.h, ref.h, _p(pose), _p(log) if want_log else None, 256 if want_log else 0, C.byref(n_log))
    if n < 0:
        raise RuntimeError(lib().hs_last_error().decode())
    return n, pose, log[:n_log.value].copy()
