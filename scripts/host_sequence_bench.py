"""Per-frame wall-clock of the tracking front end through the C++ host adapters (the reference's class interfaces), as
Tracking::Track_RGBDCam drives them (ref: src/Tracking.cpp:57,199-224): new Frame (upload + pyramid), Sprase_ImgAlign::Run
against the previous frame, UpdateLocalMap (ReprojectPoint per map point) + Feature_Alignment::SearchLocalPoints against the key
frame, Optimizer::PoseOptimization over the matches (ref: src/Tracking.cpp:236). Synthetic relief scene, smooth trajectory, one key frame with 300 map points. Rendering is outside the timed calls."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hostlib as HL
from dsdtm_b200 import synth as S


def trajectory(n, seed=3):
    rng = np.random.default_rng(seed)
    poses = [S.IDENTITY.copy()]
    v = np.array([0.004, -0.002, 0.003, 0.0008, -0.0012, 0.0005])
    for _ in range(n - 1):
        v = v + rng.normal(0, 1, 6) * np.array([3e-4] * 3 + [1e-4] * 3)
        poses.append(S.pose_mul(S.pose_from_xi(v), poses[-1]))
    return poses


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    cam = dict(S.KINECT)
    scene = S.Scene(77)
    poses = trajectory(n_frames)
    cam_h = HL.configure(cam, max_fts=300, max_frames=16)
    L = HL.lib()
    L.hs_config_set(b"Optimization.LocalBAthreshhold", b"2.0")
    img0, _, pts0 = S.render(scene, cam, poses[0], want_points=True)
    g0 = HL.HFrame(cam_h, img0, poses[0])
    n0 = g0.detect(5.0)
    px0, lv0, _ = g0.features()
    g0.attach_points(pts0[px0[:, 1].astype(int), px0[:, 0].astype(int)], np.ones(n0, np.uint8))
    kf_h = L.hs_keyframe_new(g0.h)
    imgs = [S.render(scene, cam, poses[k])[0] for k in range(n_frames)]      # rendering (CPU ray casting) is not part of the loop
    for pace_ms in (0.0, 33.0):
        g_last = g0
        t_frame, t_run, t_search, t_opt, matches, tracked = [], [], [], [], [], []
        for k in range(1, n_frames):
            img = imgs[k]
            pose_last = g_last.pose()
            if pace_ms:
                time.sleep(pace_ms * 1e-3)                                   # a 30 Hz camera: the GPU idles between frames
            t0 = time.perf_counter()
            g_cur = HL.HFrame(cam_h, img, pose_last)                     # Frame ctor: ComputeImagePyramid (H2D + pyramid kernels)
            t1 = time.perf_counter()
            n, pose, _ = HL.sparse_align_run(5, 0, 8, g_cur, g_last)     # production ctor (ref: src/Tracking.cpp:37)
            t2 = time.perf_counter()
            nrep = C.c_int(0)
            m = L.hs_search_local_points(cam_h, g_cur.h, kf_h, None, C.byref(nrep))
            t3 = time.perf_counter()
            if m < 0:
                raise RuntimeError(L.hs_last_error().decode())
            pose, _, _ = g_cur.pose_optimization()                       # ref: src/Tracking.cpp:236
            t4 = time.perf_counter()
            if k > 3:                                                    # skip warm-up frames
                t_frame.append(t1 - t0); t_run.append(t2 - t1); t_search.append(t3 - t2); t_opt.append(t4 - t3); matches.append(m); tracked.append(n)
            if g_last is not g0:
                g_last.free()
            g_last = g_cur
            err = S.pose_dist(pose, poses[k])
            assert err[0] < 2e-3 and err[1] < 5e-3, (k, err)
        if g_last is not g0:
            g_last.free()
        report(pace_ms, t_frame, t_run, t_search, t_opt, matches, tracked)


def report(pace_ms, t_frame, t_run, t_search, t_opt, matches, tracked):
    print("--- %s" % ("back to back" if not pace_ms else "paced: %.0f ms idle before every frame" % pace_ms))
    us = lambda a: (np.median(a) * 1e6, np.percentile(a, 95) * 1e6)
    print("frames timed: %d, tracked features (median) %d, matches (median) %d" % (len(t_run), np.median(tracked), np.median(matches)))
    print("Frame ctor (upload + pyramid)      : median %.1f us, p95 %.1f us" % us(t_frame))
    print("Sprase_ImgAlign::Run               : median %.1f us, p95 %.1f us" % us(t_run))
    print("UpdateLocalMap + SearchLocalPoints : median %.1f us, p95 %.1f us" % us(t_search))
    print("Optimizer::PoseOptimization        : median %.1f us, p95 %.1f us" % us(t_opt))
    tot = np.array(t_frame) + np.array(t_run) + np.array(t_search)
    print("front end per frame (first three)  : median %.1f us, p95 %.1f us" % us(tot))
    print("... including PoseOptimization     : median %.1f us, p95 %.1f us" % us(tot + np.array(t_opt)))


if __name__ == "__main__":
    main()
