"""BASELINE configs[3]: the tracking front end on a synthetic 1000-frame RGB-D sequence at TUM shape (640x480, 16-bit depth,
scale 5000), driven through the C++ adapter classes in Tracking's order (ref: src/Tracking.cpp:57,199-236,412-464):
  new Frame (upload + pyramid) -> Sprase_ImgAlign(5,0,8)::Run(cur, last) -> Tracking::UpdateLocalMap (GetCloseKeyFrames over the
  WHOLE map on the device-resident table, the 10 nearest close key frames) + SearchLocalPoints -> Optimizer::PoseOptimization -> every KF_EVERY frames CraeteKeyframe (detect on the free cells,
  UndistortFeatures / depth lookup / UnProject on the device, new map points).
Frames are ray-cast beforehand by a process pool (CPU work, not part of the loop). Reports the per-call wall clock, the matches,
and the trajectory error against the ground truth. Every CHECK_EVERY-th frame the pose Run returns is compared with the oracle's
Sprase_ImgAlign::Run on the same two images, features, map points and start pose (1e-5 rad / 1e-5 m, tracked counts equal)."""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from dsdtm_b200 import synth as S

SCALE = 5000.0
KF_EVERY = 20
SCENE_SEED = 1000
CHECK_EVERY = 50


def ground_truth(n):
    """Smooth closed-form trajectory: a slow Lissajous sweep over the relief with a few degrees of rotation."""
    poses = []
    for k in range(n):
        t = float(k)
        xi = np.array([0.55 * np.sin(2 * np.pi * t / 500.0), 0.35 * np.sin(2 * np.pi * t / 333.0), 0.12 * np.sin(2 * np.pi * t / 250.0),
                       np.deg2rad(2.5) * np.sin(2 * np.pi * t / 400.0), np.deg2rad(3.0) * np.sin(2 * np.pi * t / 287.0),
                       np.deg2rad(2.0) * np.sin(2 * np.pi * t / 611.0)])
        poses.append(S.pose_from_xi(xi))
    return poses


def _render(args):
    k, pose, want_depth = args
    scene = S.Scene(SCENE_SEED)
    img, z = S.render(scene, S.KINECT, pose)
    d16 = np.clip(np.rint(z * SCALE), 0, 65535).astype(np.uint16) if want_depth else None
    return k, img, d16


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 2)
    cam = dict(S.KINECT)
    poses = ground_truth(n)
    t0 = time.time()
    with mp.Pool(procs) as pool:                       # before any CUDA call: the workers are forked from a CUDA-free process
        out = pool.map(_render, [(k, poses[k], k % KF_EVERY == 0) for k in range(n)], chunksize=4)
    imgs = [o[1] for o in out]; depth = {o[0]: o[2] for o in out if o[2] is not None}
    print("rendered %d frames with %d processes in %.1f s" % (n, procs, time.time() - t0))

    import hostlib as HL
    cam_h = HL.configure(cam, max_fts=300, max_frames=128, dist=(0.0, 0.0, 0.0, 0.0, 0.0))   # every key frame that observes a local map point must stay resident during a call
    L = HL.lib()
    L.hs_config_set(b"Optimization.LocalBAthreshhold", b"2.0")

    def lift(frame, d16, start):
        tab, _ = frame.keyframe_lift(d16, SCALE)
        new = np.arange(start, len(tab))
        frame.attach_points_from(int(start), tab[new, 3:6], (tab[new, 2] > 0).astype(np.uint8))
        return len(new)

    g_last = HL.HFrame(cam_h, imgs[0], poses[0])
    assert g_last.detect(5.0) == 300
    lift(g_last, depth[0], 0)
    kfs = [L.hs_keyframe_new(g_last.h)]
    L.hs_map_add_keyframe(kfs[-1])
    kf_frames = {0: g_last}
    T = {k: [] for k in ("frame", "run", "search", "opt", "keyframe", "kf_detect", "kf_lift", "kf_map")}
    tracked, matches, iters, err_sa, err_po, n_local, n_reproj = [], [], [], [], [], [], []
    lost = 0
    checked = []
    for k in range(1, n):
        t0 = time.perf_counter()
        g_cur = HL.HFrame(cam_h, imgs[k], g_last.pose())
        t1 = time.perf_counter()
        if k % CHECK_EVERY == 0:                                # snapshot Run's inputs for the oracle (outside the timed calls)
            import oracle as O
            pxl, lvl, inil = g_last.features()
            idl = g_last.mp_ids()
            Fo = np.zeros(len(pxl), O.REF_FEAT_DT)
            oc = O.make_cam(cam["width"], cam["height"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["f"])
            for j in range(len(pxl)):
                Fo[j]["px"] = pxl[j]; Fo[j]["level"] = lvl[j]; Fo[j]["initial"] = inil[j]
                Fo[j]["normal"] = O.feature_normal(oc, pxl[j])
                if idl[j] >= 0:
                    P = np.zeros(3); L.hs_mappoint_pose(int(idl[j]), HL._p(P)); Fo[j]["point_w"] = P
            last_pose = g_last.pose()
            t1 = time.perf_counter()
        nt, pose_sa, _ = HL.sparse_align_run(5, 0, 8, g_cur, g_last, want_log=False)
        t2 = time.perf_counter()
        if k % CHECK_EVERY == 0:
            rp, offs, ws, hs = O.pyramid(imgs[k - 1], 5)
            cp = O.pyramid(imgs[k], 5)[0]
            T0 = O.se3_mul(last_pose, O.se3_inv(last_pose))         # cur starts at the last pose (ref: src/Tracking.cpp:201)
            po, no, _ = O.sparse_align(oc, rp, cp, offs, ws, hs, Fo, O.se3_inv(last_pose)[4:], T0, 5, 0, 8)
            dd = S.pose_dist(O.se3_mul(po, last_pose), pose_sa)
            assert no == nt and dd[0] < 1e-5 and dd[1] < 1e-5, ("oracle check failed at frame %d" % k, dd, no, nt)
            checked.append(max(dd))
        m, local, nrep = HL.track_local_map(cam_h, g_cur)
        n_local.append(len(local)); n_reproj.append(nrep)
        t3 = time.perf_counter()
        pose_po, summ, _ = g_cur.pose_optimization()
        t4 = time.perf_counter()
        T["frame"].append(t1 - t0); T["run"].append(t2 - t1); T["search"].append(t3 - t2); T["opt"].append(t4 - t3)
        tracked.append(nt); matches.append(m); iters.append(int(summ["iterations"]))
        err_sa.append(S.pose_dist(pose_sa, poses[k])); err_po.append(S.pose_dist(pose_po, poses[k]))
        if nt < 30 or m < 30:                                   # Tracking's "too few features" thresholds (ref: src/Tracking.cpp:244-256)
            lost += 1
        if k % KF_EVERY == 0:
            t5 = time.perf_counter()
            n_old = len(g_cur.features()[0])
            g_cur.detect(5.0, use_existing=True)
            t6 = time.perf_counter()
            lift(g_cur, depth[k], n_old)
            t7 = time.perf_counter()
            kfs.append(L.hs_keyframe_new(g_cur.h))
            L.hs_map_add_keyframe(kfs[-1])
            kf_frames[k] = g_cur
            t8 = time.perf_counter()
            T["keyframe"].append(t8 - t5); T["kf_detect"].append(t6 - t5); T["kf_lift"].append(t7 - t6); T["kf_map"].append(t8 - t7)
        if g_last is not None and (k - 1) not in kf_frames:
            g_last.free()
        g_last = g_cur
    err_sa, err_po = np.array(err_sa), np.array(err_po)
    us = lambda a: (np.median(a) * 1e6, np.percentile(a, 95) * 1e6)
    print("frames %d, key frames %d, lost %d" % (n, len(kfs), lost))
    print("tracked features (Run): median %d min %d; matches (SearchLocalPoints): median %d min %d; LM iterations: mean %.2f max %d"
          % (np.median(tracked), np.min(tracked), np.median(matches), np.min(matches), np.mean(iters), np.max(iters)))
    print("local key frames per frame: median %d max %d; reprojected local map points: median %d max %d" % (np.median(n_local), np.max(n_local), np.median(n_reproj), np.max(n_reproj)))
    for key, name in (("frame", "Frame ctor (upload + pyramid)"), ("run", "Sprase_ImgAlign::Run"), ("search", "UpdateLocalMap + SearchLocalPoints"),
                      ("opt", "Optimizer::PoseOptimization"), ("keyframe", "CraeteKeyframe (detect + lift), per key frame"),
                      ("kf_detect", "  of which Feature_detector::detect"), ("kf_lift", "  of which depth upload + UndistortFeatures + map points"),
                      ("kf_map", "  of which KeyFrame copy + Map::AddKeyFrame")):
        print("%-46s: median %.1f us, p95 %.1f us" % ((name,) + us(T[key])))
    tot = np.array(T["frame"]) + np.array(T["run"]) + np.array(T["search"]) + np.array(T["opt"])
    print("%-46s: median %.1f us, p95 %.1f us  (%.0f frames/s)" % (("front end per frame",) + us(tot) + (1.0 / np.mean(tot),)))
    print("pose error vs ground truth after Run              : mean %.2e rad %.2e m, max %.2e rad %.2e m" % (*err_sa.mean(0), *err_sa.max(0)))
    print("pose error vs ground truth after PoseOptimization : mean %.2e rad %.2e m, max %.2e rad %.2e m" % (*err_po.mean(0), *err_po.max(0)))
    print("oracle checks of Run (every %d frames): %d frames, max pose difference %.1e (tolerance 1e-5)" % (CHECK_EVERY, len(checked), max(checked) if checked else 0.0))
    q = [n // 4, n // 2, 3 * n // 4, n - 2]
    print("error along the sequence (rad, m) at frames %s: %s" % (q, ["%.1e/%.1e" % tuple(err_po[i]) for i in q]))


if __name__ == "__main__":
    main()
