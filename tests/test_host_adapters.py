"""The drop-in boundary: the reference's class interfaces (DSDTM::Frame::ComputeImagePyramid, Feature_detector::detect,
Sprase_ImgAlign::Run, Feature_Alignment::SearchLocalPoints) driven as Tracking drives them, checked against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
import hostlib as HL
import oracle as O
from dsdtm_b200 import synth as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------ CPU-side host logic
def test_config_parser_reads_reference_style_yaml_with_conflict_markers(built, tmp_path):
    y = tmp_path / "kinect.yaml"
    y.write_text("%YAML:1.0\n# Camera Parameters\n<<<<<<< HEAD\nCamera.fx: 520.9\n=======\nCamera.fx: 517.306408\n>>>>>>> dev\n"
                 "Camera.fy: 516.469215   # fr3\nCamera.width: 640\nCamera.MaxPyraLevels: 5\nCamera.f: 525\nname: \"kinect\"\n")
    L = HL.lib()
    L.hs_reset()
    assert L.hs_config_load(str(y).encode()) == 0
    assert L.hs_config_get(b"Camera.fx") == pytest.approx(517.306408)       # the later block wins, like a resolved merge
    assert L.hs_config_get(b"Camera.fy") == pytest.approx(516.469215)
    assert L.hs_config_get_int(b"Camera.width") == 640 and L.hs_config_get_int(b"Camera.MaxPyraLevels") == 5
    assert L.hs_config_get_int(b"Camera.Missing") == 0                     # cv::FileStorage yields 0 for a missing node
    assert L.hs_config_load(b"/nonexistent.yaml") == -1 and b"does not exist" in L.hs_last_error()


def test_host_circle_matches_cv2_goldens(built, golden):
    g = golden["circle_cv2"]
    Hh, W = g["shape"]
    for (cx, cy, r), bits in zip(g["cases"], g["masks"]):
        m = np.full((Hh, W), 255, np.uint8)
        HL.lib().hs_circle(HL._p(m), int(W), int(Hh), float(cx), float(cy), int(r), 0)
        assert ((m == 0) == np.unpackbits(bits)[:Hh * W].reshape(Hh, W).astype(bool)).all(), (cx, cy, r)


def test_host_adapters_fail_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cam_h = HL.configure(S.KINECT)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        HL.HFrame(cam_h, np.zeros((480, 640), np.uint8), S.IDENTITY)


# ------------------------------------------------------------------------------------------------ GPU: the classes as Tracking calls them
@pytest.fixture(scope="module")
def rig(built, scenario):
    cam_h = HL.configure(scenario["cam"], max_fts=300)
    ref = HL.HFrame(cam_h, scenario["ref_img"], scenario["T_ref"])
    cur = HL.HFrame(cam_h, scenario["cur_img"], scenario["T_ref"])       # cur.Set_Pose(last.Get_Pose()), ref: src/Tracking.cpp:201
    return dict(cam=cam_h, ref=ref, cur=cur)


@pytest.mark.gpu
def test_frame_compute_image_pyramid(rig, scenario):
    packed, offs, ws, hs = scenario["ref_pyr"]
    for l in range(5):
        want = O.pyr_level(packed, offs, ws, hs, l)
        assert (rig["ref"].level(l, want.shape) == want).all(), l


@pytest.mark.gpu
def test_feature_detector_detect_equals_reference_selection(rig, scenario):
    """Initializer path: detect(frame, 5.0) on a fresh frame (ref: src/Initializer.cpp:44)."""
    n = rig["ref"].detect(5.0)
    px, lv, ini = rig["ref"].features()
    want = scenario["corners"]
    assert n == len(want) == 300
    assert (px[:, 0] == want["x"]).all() and (px[:, 1] == want["y"]).all() and (lv == want["level"]).all()
    assert not ini.any()
    assert rig["ref"].mask().max() == 0            # mImgMask.release() at the end of detect (ref: :153)


@pytest.mark.gpu
def test_sprase_imgalign_run(rig, scenario):
    """TrackWithLastFrame: Run(cur, last) (ref: src/Tracking.cpp:199-217) with map points attached to the ref features."""
    ref, cur = rig["ref"], rig["cur"]
    if HL.lib().hs_frame_n_features(ref.h) == 0:
        ref.detect(5.0)
    c = scenario["corners"]
    pts = scenario["ref_points"][c["y"], c["x"]]
    ref.attach_points(pts, np.ones(len(c), np.uint8))
    packed, offs, ws, hs = scenario["ref_pyr"]
    for cfg in ((4, 0, 30), (5, 0, 8)):
        cur.set_pose(scenario["T_ref"])
        n, pose, log = HL.sparse_align_run(*cfg, cur, ref)
        po, no, lo = O.sparse_align(H.ocam(scenario["cam"]), packed, scenario["cur_pyr"][0], offs, ws, hs, scenario["feats"],
                                    scenario["ref_center"], S.IDENTITY, *cfg)
        want = O.se3_mul(po, scenario["T_ref"])                      # ref: src/Sprase_ImageAlign.cpp:57
        d = S.pose_dist(want, pose)
        assert d[0] < 1e-5 and d[1] < 1e-5 and n == no
        assert len(log) == len(lo) and all(abs(a["chi2"] - b["chi2"]) <= 1e-4 * abs(a["chi2"]) for a, b in zip(lo, log))
        e = S.pose_dist(pose, scenario["T_cur"])
        assert e[0] < 2e-4 and e[1] < 5e-4


@pytest.mark.gpu
def test_sprase_imgalign_too_few_features_returns_zero(built, scenario):
    cam_h = HL.configure(scenario["cam"], min_fts=15)
    ref = HL.HFrame(cam_h, scenario["ref_img"], S.IDENTITY)
    cur = HL.HFrame(cam_h, scenario["cur_img"], S.IDENTITY)
    n, pose, log = HL.sparse_align_run(4, 0, 30, cur, ref)              # ref frame has no features (< Camera.Min_fts)
    assert n == 0 and len(log) == 0 and np.allclose(pose, S.IDENTITY)   # ref: src/Sprase_ImageAlign.cpp:34-38


def _emulate_search_local_points(sc, kf_feats, cur_pose, found, cell=15, levels=5):
    """SearchLocalPoints restated with oracle primitives (ref: src/Feature_alignment.cpp:54-158): reproject, bin, sort by
    found count (stable list::sort), first match per cell, mask painting, stop at 200."""
    cam = sc["cam"]; oc = H.ocam(cam)
    w, h = cam["width"], cam["height"]
    gcols = -(-w // cell)
    cells = {}
    fx, fy, cx, cy = (float(np.float32(cam[k])) for k in ("fx", "fy", "cx", "cy"))
    for i, f in enumerate(kf_feats):
        q = O.se3_act(cur_pose, f["point_w"])
        px = np.array([fx * q[0] / q[2] + cx, fy * q[1] / q[2] + cy])
        rx, ry = O.cvround(np.float32(px[0])), O.cvround(np.float32(px[1]))
        if rx >= 8 and rx < w - 8 and ry >= 8 and ry < h - 8:
            cells.setdefault(int(px[1] / cell) * gcols + int(px[0] / cell), []).append((i, px))
    mask = np.full((h, w), 255, np.uint8)
    packed, offs, ws, hs = sc["ref_pyr"]
    cpacked = sc["cur_pyr"][0]
    kf_center = sc["ref_center"]
    T_c2r = O.se3_mul(cur_pose, O.se3_inv(sc["T_ref"]))
    out = []
    matches = 0
    for k in sorted(cells):
        cands = sorted(cells[k], key=lambda c: -found[c[0]])            # stable, like std::list::sort
        for i, px in cands:
            if mask[O.cvround(np.float32(px[1])), O.cvround(np.float32(px[0]))] != 255:
                continue
            f = kf_feats[i]
            # Get_ClosetObs: one observation, cos angle between (kf centre - P) and (cur centre - P)
            cur_center = O.se3_inv(cur_pose)[4:]
            a = kf_center - f["point_w"]; b = cur_center - f["point_w"]
            if np.dot(a / np.linalg.norm(a), b / np.linalg.norm(b)) < 0.5:
                continue
            L0 = int(f["level"])
            rpx = f["px"] / np.float32(1 << L0)
            if not (O.cvround(rpx[0]) >= 5 and O.cvround(rpx[0]) < w // (1 << L0) - 5 and O.cvround(rpx[1]) >= 5 and O.cvround(rpx[1]) < h // (1 << L0) - 5):
                continue
            A = O.solve_affine(oc, kf_center, f["point_w"], f["normal"], f["px"], L0, T_c2r)
            SL = O.best_search_level(A, levels - 3)
            patch = O.warp_affine(A, O.pyr_level(packed, offs, ws, hs, L0), f["px"], L0, SL)
            p, conv, _ = O.align2d(O.pyr_level(cpacked, offs, ws, hs, SL), patch, 10, px / (1 << SL))
            p = p * (1 << SL)
            if not conv:
                continue
            out.append((i, p, SL))
            cx_, cy_ = O.cvround(np.float32(p[0])), O.cvround(np.float32(p[1]))
            O.circle_fill(mask, cx_, cy_, cell, 0)
            matches += 1
            break
        if matches >= 200:
            break
    return out, mask


@pytest.mark.gpu
def test_feature_alignment_search_local_points(built, scenario):
    """TrackWithLocalMap: UpdateLocalMap's ReprojectPoint loop + SearchLocalPoints (ref: src/Tracking.cpp:219-313) against a
    one-keyframe map; the speculative GPU batch + host replay must reproduce the reference's sequential greedy matching."""
    sc = scenario
    cam_h = HL.configure(sc["cam"], max_fts=300)
    ref = HL.HFrame(cam_h, sc["ref_img"], sc["T_ref"])
    assert ref.detect(5.0) == 300
    c = sc["corners"]
    ref.attach_points(sc["ref_points"][c["y"], c["x"]], np.ones(len(c), np.uint8))
    kf = HL.lib().hs_keyframe_new(ref.h)
    cur = HL.HFrame(cam_h, sc["cur_img"], sc["T_cur"])                   # pose after sparse alignment ~ ground truth
    rng = np.random.default_rng(5)
    found = rng.integers(1, 6, len(c)).astype(np.int32)
    nrep = C.c_int(0)
    m = HL.lib().hs_search_local_points(cam_h, cur.h, kf, HL._p(found), C.byref(nrep))
    assert m >= 0, HL.lib().hs_last_error()
    px, lv, ini = cur.features()
    want, want_mask = _emulate_search_local_points(sc, sc["feats"], sc["T_cur"], found)
    assert m == len(want) == len(px) and m > 100 and m <= 200
    for (i, p, SL), gp, gl in zip(want, px, lv):
        assert gl == SL and np.abs(np.float32(p) - gp).max() <= 1e-3      # refined feature position tolerance (north_star)
    assert ini.all()
    assert (cur.mask() == want_mask).all()
    # refined positions land on the true reprojections
    truth = np.array([want_i[1] for want_i in want])
    assert np.median(np.linalg.norm(px - truth, axis=1)) < 1e-3


@pytest.mark.gpu
def test_frame_rgbd_keyframe_path(built, scenario):
    """Tracking::CraeteKeyframe's per-feature arithmetic through the Frame interface (ref: src/Tracking.cpp:412-464):
    detect -> UndistortFeatures (ref: src/Frame.cpp:94-150) -> Get_FeatureDetph (:200-224) -> UnProject (:152-157), with the
    EuRoC distortion set on the kinect geometry, against the oracle (cv2-pinned undistortPoints)."""
    sc = scenario
    dist = (-0.28368365, 0.07451284, -0.00010473, -3.55590700e-05, 0.0)
    cam_h = HL.configure(sc["cam"], max_fts=300, dist=dist)
    oc = H.ocam(sc["cam"])
    pose = S.pose_from_xi([0.1, -0.05, 0.02, 0.02, -0.03, 0.01])
    fr = HL.HFrame(cam_h, sc["ref_img"], pose)
    assert fr.detect(5.0) == 300
    px0, lv, _ = fr.features()
    rng = np.random.default_rng(12)
    d16 = rng.integers(500, 40000, sc["ref_img"].shape).astype(np.uint16)
    d16[rng.uniform(size=d16.shape) < 0.4] = 0
    df = O.depth_convert(d16, 5000.0)
    tab, nrm = fr.keyframe_lift(d16, 5000.0)
    px1, _, _ = fr.features()
    und = O.undistort_points(oc, dist, px0)
    assert (px1.view(np.uint32) == und.view(np.uint32)).all()                  # mpx rewritten in place, bit-equal to cv2's result
    n_ok = 0
    for i in range(len(und)):
        assert np.allclose(nrm[i], O.feature_normal(oc, und[i]), rtol=0, atol=1e-15)
        z = O.feature_depth(df, und[i])
        assert tab[i, 2] == np.float32(z)
        if z >= 0:
            want = O.unproject(oc, pose, und[i], z)
            assert np.allclose(tab[i, 3:6], want, rtol=0, atol=1e-12) and np.allclose(tab[i, 6:9], want, rtol=0, atol=1e-12)
            n_ok += 1
    assert 100 < n_ok < 300
    fr.free()


@pytest.mark.gpu
def test_single_candidate_warp_affine_and_patch_forms(built, scenario):
    """Feature_Alignment::WarpAffine + GetPatchNoBoarder as public single-candidate calls (ref: src/Feature_alignment.cpp:206-275)."""
    sc = scenario
    cam_h = HL.configure(sc["cam"], max_fts=300)
    ref = HL.HFrame(cam_h, sc["ref_img"], sc["T_ref"])
    assert ref.detect(5.0) == 300
    corners = sc["corners"]
    ref.attach_points(sc["ref_points"][corners["y"], corners["x"]], np.ones(len(corners), np.uint8))
    kf = HL.lib().hs_keyframe_new(ref.h)
    packed, offs, ws, hs = sc["ref_pyr"]
    rng = np.random.default_rng(4)
    for idx in (0, 7, 123, 299):
        A = np.eye(2) + rng.uniform(-0.2, 0.2, (2, 2))
        L0 = int(corners[idx]["level"])
        p10 = np.zeros(100, np.uint8); p8 = np.zeros(64, np.uint8)
        assert HL.lib().hs_warp_affine_single(cam_h, kf, idx, HL._p(np.ascontiguousarray(A)), 0, HL._p(p10), HL._p(p8)) == 0, HL.lib().hs_last_error()
        want = O.warp_affine(A, O.pyr_level(packed, offs, ws, hs, L0), sc["feats"][idx]["px"], L0, 0)
        assert (p10 == want).all() and (p8 == O.patch_no_border(want)).all(), idx
    ref.free()


@pytest.mark.gpu
def test_align2d_static_signature_on_an_arbitrary_image(built, scenario):
    """Feature_Alignment::Align2DGaussNewton(const Mat&, uchar*, uchar*, int, Vector2d&) -- the reference's own static signature
    (ref: include/Feature_alignment.h:85) -- with the recipe of ref Test/test_Feature_alignment.cpp:56-81: reference patch interpolated
    at px_true = (130.2, 120.3), start offset (+1.1, +0.8), 3 iterations; the image is NOT a Frame (here: level 0 and level 1 of the
    scenario pyramid handed in as plain images). Against the oracle and the pinned reference value."""
    sc = scenario
    HL.configure(sc["cam"], max_fts=300)
    packed, offs, ws, hs = sc["cur_pyr"]
    px_true = np.array([130.2, 120.3])
    for L in (0, 1):
        img = np.ascontiguousarray(O.pyr_level(packed, offs, ws, hs, L))
        # generateRefPatchNoWarpInterpolate (ref: Test/test_Feature_alignment.cpp:22-45): bilinear 10x10 around px_true, truncated to u8
        u_r, v_r = int(np.floor(px_true[0])), int(np.floor(px_true[1]))
        sx, sy = px_true[0] - u_r, px_true[1] - v_r
        wTL, wTR, wBL, wBR = (1 - sx) * (1 - sy), sx * (1 - sy), (1 - sx) * sy, sx * sy
        I = img.astype(np.float64)
        p10 = np.zeros((10, 10), np.uint8)
        for y in range(10):
            for x in range(10):
                yy, xx = v_r + y - 5, u_r + x - 5
                p10[y, x] = np.uint8(wTL * I[yy, xx] + wTR * I[yy, xx + 1] + wBL * I[yy + 1, xx] + wBR * I[yy + 1, xx + 1])
        p8 = O.patch_no_border(p10)
        px = px_true + (1.1, 0.8)
        got = px.copy()
        rc = HL.lib().hs_align2d_image(HL._p(img), img.shape[1], img.shape[0], HL._p(np.ascontiguousarray(p10.reshape(-1))), HL._p(p8), 3, HL._p(got))
        assert rc >= 0, HL.lib().hs_last_error()
        want, conv, _ = O.align2d(img, p10, 3, px)
        assert bool(rc) == conv and np.abs(got - want).max() <= 1e-3
        assert np.linalg.norm(got - px_true) < 0.5                   # three iterations from 1.36 px away (the reference prints 0.015 px on its own, not shipped, image)
    bad = np.zeros((100, 123), np.uint8)
    assert HL.lib().hs_align2d_image(HL._p(bad), 123, 100, HL._p(np.zeros(100, np.uint8)), HL._p(np.zeros(64, np.uint8)), 3, HL._p(np.zeros(2))) == -1
    assert b"pyramid level" in HL.lib().hs_last_error()


@pytest.mark.gpu
def test_optimizer_pose_optimization_after_search_local_points(built, scenario):
    """Tracking's next call after SearchLocalPoints (ref: src/Tracking.cpp:236): Optimizer::PoseOptimization through the adapter
    class against the oracle's restatement of ceres::Solve -- pose, stopping rule, residuals, and the reference's literal
    EraseFound bookkeeping (residual BLOCK i is charged to the map point of FEATURE i, ref: src/Optimizer.cpp:81-94)."""
    sc = scenario
    cam_h = HL.configure(sc["cam"], max_fts=300)
    L = HL.lib()
    ref = HL.HFrame(cam_h, sc["ref_img"], sc["T_ref"])
    assert ref.detect(5.0) == 300
    c = sc["corners"]
    ref.attach_points(sc["ref_points"][c["y"], c["x"]], np.ones(len(c), np.uint8))
    kf = L.hs_keyframe_new(ref.h)
    cur = HL.HFrame(cam_h, sc["cur_img"], sc["T_cur"])
    found0 = np.full(len(c), 3, np.int32)
    m = L.hs_search_local_points(cam_h, cur.h, kf, HL._p(found0), None)
    assert m > 100, L.hs_last_error()
    px, lv, ini = cur.features()
    ids = cur.mp_ids()
    assert (ids >= 0).all() and ini.all()
    # make the numbering of features and residual blocks differ: one matched map point goes bad before the solve
    bad_feature = 7
    L.hs_mappoint_set_bad(int(ids[bad_feature]), 1)
    # start away from the optimum
    start = O.se3_mul(O.se3_exp(np.array([0.01, -0.008, 0.012, 0.004, -0.003, 0.005])), sc["T_cur"])
    cur.set_pose(start)
    found_before = np.array([L.hs_mappoint_found(int(i)) for i in ids])
    # the oracle on the same snapshot
    oc = H.ocam(sc["cam"])
    keep = np.arange(len(px)) != bad_feature
    normals = np.array([O.feature_normal(oc, p) for p in px])
    pts = sc["ref_points"][c["y"], c["x"]][ids]
    a, ra, sa = O.pose_optimization(normals[keep], lv[keep], pts[keep], start)
    # Optimization.LocalBAthreshhold (pixels; 2.0 in the reference's configs) chosen between two residuals of this frame so that
    # both branches of the outlier loop run
    srt = np.sort(ra)
    thr_px = float(np.float32(0.5 * (srt[len(srt) // 2] + srt[len(srt) // 2 + 1]) * float(np.float32(sc["cam"]["f"]))))
    L.hs_config_set(b"Optimization.LocalBAthreshhold", repr(thr_px).encode())
    pose = np.empty(7); summ = np.zeros(1, O.BA_SUMMARY_DT); res = np.zeros(len(px))
    nb = L.hs_pose_optimization(cur.h, HL._p(pose), HL._p(summ), HL._p(res), len(px))
    assert nb == len(px) - 1, L.hs_last_error()
    s = summ[0]
    assert (s["iterations"], s["termination"], s["n_successful"]) == (sa["iterations"], sa["termination"], sa["n_successful"])
    assert np.abs(a - pose).max() < 1e-9 and np.abs(ra - res[:nb]).max() < 1e-9
    assert np.abs(cur.pose() - pose).max() == 0
    # back near the ground truth (sub-pixel matches): well inside the north_star pose budget of the front end
    assert np.abs(pose - sc["T_cur"]).max() < 2e-4
    # EraseFound, literally: block i -> feature i; feature `bad_feature` has no entry (skipped), later blocks are charged to the
    # feature one position BEFORE their own
    thr = float(np.float32(thr_px)) / float(np.float32(sc["cam"]["f"]))
    assert np.abs(ra - thr).min() > 1e-8                      # no residual sits on the threshold
    expect = found_before.copy()
    for i in range(nb):
        if ra[i] > thr and i != bad_feature:
            expect[i] -= 1
    found_after = np.array([L.hs_mappoint_found(int(i)) for i in ids])
    assert (found_after == expect).all()
    assert 0 < (expect != found_before).sum() < nb            # the threshold splits the matches: both branches ran


@pytest.mark.gpu
def test_pyramid_levels_are_fetched_on_first_read_even_after_the_slot_was_recycled(built, scenario):   # last: it re-configures the runtime
    """mvImg_Pyr[l] (l >= 1) is a lazy host copy: nothing is downloaded in the constructor; a read after the frame's slot went
    to another frame re-uploads level 0, rebuilds the pyramid and returns the same bytes (ref: src/Frame.cpp:74-81)."""
    cam_h = HL.configure(scenario["cam"], max_fts=300, max_frames=4)
    first = HL.HFrame(cam_h, scenario["ref_img"], scenario["T_ref"])
    others = [HL.HFrame(cam_h, scenario["cur_img"], scenario["T_ref"]) for _ in range(6)]     # 6 > 4 slots: `first` is evicted
    packed, offs, ws, hs = scenario["ref_pyr"]
    for l in (4, 1, 3, 2, 0):
        want = O.pyr_level(packed, offs, ws, hs, l)
        assert (first.level(l, want.shape) == want).all(), l
    packed, offs, ws, hs = scenario["cur_pyr"]
    for l in range(5):
        want = O.pyr_level(packed, offs, ws, hs, l)
        assert (others[-1].level(l, want.shape) == want).all(), l
