"""configs[3] / configs[4] of BASELINE.json at test scale.

* test_tracking_front_end_sequence: the front-end calls of Tracking (ref: src/Tracking.cpp:45-145,199-313,412-464) over a
  synthetic RGB-D sequence at TUM shape through the C++ adapters -- pyramid, FAST + grid selection on keyframes, sparse
  alignment against the last frame, reprojection + feature alignment against the keyframe map -- checked frame by frame
  against the same loop restated with the oracle. (The 1000-frame length of configs[3] is a bench-scale number; parity is
  per frame, so a short sequence exercises every transition: init keyframe, tracking, new keyframe.)
* test_batched_sweep_752x480: configs[4] at reduced count -- independent EuRoC-geometry pairs through the batched entry
  point, each pair against the oracle, plus the size-independent properties used at full size (batch == singles,
  permutation invariance, determinism)."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
import hostlib as HL
import oracle as O
from dsdtm_b200 import synth as S

from oracle_seq import CELL, LEVELS, OMap, OracleFrame, _oracle_search, _oracle_search_multi, _oracle_sparse_align, _trajectory

pytestmark = pytest.mark.gpu


def test_tracking_front_end_sequence(built):
    cam = dict(S.KINECT)
    oc = H.ocam(cam)
    scene = S.Scene(77)
    n_frames = 7
    poses = _trajectory(n_frames)
    cam_h = HL.configure(cam, max_fts=300, max_frames=16)
    cfg = (5, 0, 8)                                       # production ctor (ref: src/Tracking.cpp:37)

    # ---- frame 0: Initializer::Init_RGBDCam = detect + map points from depth (ref: src/Initializer.cpp:40-57,147-196)
    img0, _, pts0 = S.render(scene, cam, poses[0], want_points=True)
    g0 = HL.HFrame(cam_h, img0, poses[0])
    assert g0.detect(5.0) == 300
    px0, lv0, _ = g0.features()
    o0 = OracleFrame(img0, poses[0])
    corners, _ = H.detect_oracle(img0, LEVELS, CELL, 300)
    assert (px0[:, 0] == corners["x"]).all() and (px0[:, 1] == corners["y"]).all() and (lv0 == corners["level"]).all()
    world = pts0[corners["y"], corners["x"]]
    g0.attach_points(world, np.ones(len(corners), np.uint8))
    o0.feats = H.ref_feats_from_corners(cam, corners, pts0)
    kf_h = HL.lib().hs_keyframe_new(g0.h)
    o_kf = o0
    found = np.ones(len(corners), np.int32)

    g_last, o_last = g0, o0
    for k in range(1, n_frames):
        img, _ = S.render(scene, cam, poses[k])
        # Track_RGBDCam: new Frame (pyramid), cur.Set_Pose(last.Get_Pose()), Run(cur, last)  (ref: src/Tracking.cpp:57,199-205)
        g_cur = HL.HFrame(cam_h, img, g_last.pose())
        o_cur = OracleFrame(img, o_last.pose)
        for l in range(LEVELS):
            packed, offs, ws, hs = o_cur.pyr
            assert (g_cur.level(l, (hs[l], ws[l])) == O.pyr_level(packed, offs, ws, hs, l)).all()
        n_g, pose_g, _ = HL.sparse_align_run(*cfg, g_cur, g_last)
        pose_o, n_o = _oracle_sparse_align(oc, o_cur, o_last, cfg)
        o_cur.pose = pose_o
        d = S.pose_dist(pose_o, pose_g)
        assert d[0] < 1e-5 and d[1] < 1e-5 and n_g == n_o and n_g >= 20, (k, d, n_g, n_o)      # < 20 would mean Lost (ref: :208)
        e = S.pose_dist(pose_g, poses[k])
        assert e[0] < 1e-3 and e[1] < 3e-3, (k, e)
        # TrackWithLocalMap: reproject the keyframe's map points, SearchLocalPoints (ref: :219-313)
        nrep = C.c_int(0)
        m = HL.lib().hs_search_local_points(cam_h, g_cur.h, kf_h, HL._p(found), C.byref(nrep))
        want = _oracle_search(oc, cam, o_cur, o_kf, found)
        px, lv, ini = g_cur.features()
        assert m == len(want) == len(px) and m > 100, (k, m, len(want))
        for (i, p, SL), gp, gl in zip(want, px, lv):
            assert gl == SL and np.abs(p - gp).max() <= 1e-3
        # MapPoint::IncreaseFound on every match (ref: src/Feature_alignment.cpp:106); the matched features become the next
        # frame's reference features (px refined, level, bearing from px, same map point)
        F = np.zeros(len(want), O.REF_FEAT_DT)
        for j, (i, p, SL) in enumerate(want):
            found[i] += 1
            F[j]["px"] = px[j]; F[j]["level"] = SL; F[j]["initial"] = 1
            F[j]["normal"] = O.feature_normal(oc, px[j]); F[j]["point_w"] = o_kf.feats[i]["point_w"]
        o_cur.feats = F
        g_last, o_last = g_cur, o_cur


def test_batched_sweep_752x480(built):
    from dsdtm_b200 import capi
    cam = dict(S.EUROC)
    n, stride = 6, 320
    ctx = capi.Context(cam, levels=5, cell_size=15, max_feats=stride, max_patches=8, max_frames=2 * n, max_batch=n)
    assert ctx.ws == [752, 376, 188, 94, 47] and ctx.hs == [480, 240, 120, 60, 30]        # SURVEY 3.5 (tile kernel for 94, 47)
    scs = [H.make_scenario(100 + i, cam) for i in range(3)]
    feats = np.zeros((n, stride), O.REF_FEAT_DT); nf = np.zeros(n, np.int32)
    centers = np.zeros((n, 3)); poses = np.tile(S.IDENTITY, (n, 1))
    rng = np.random.default_rng(1)
    for i in range(n):
        sc = scs[i % 3]
        ctx.upload(2 * i, sc["ref_img"]); ctx.upload(2 * i + 1, sc["cur_img"])
        nf[i] = len(sc["feats"]); feats[i, :nf[i]] = sc["feats"]; centers[i] = sc["ref_center"]
        if i >= 3:
            poses[i] = S.pose_from_xi(rng.uniform(-0.003, 0.003, 6))
    for l in range(5):
        packed, offs, ws, hs = scs[0]["cur_pyr"]
        assert (ctx.download_level(1, l) == O.pyr_level(packed, offs, ws, hs, l)).all()
    ref_slots = 2 * np.arange(n); cur_slots = ref_slots + 1
    pb, tb, log, nlog = ctx.sparse_align_batch(ref_slots, cur_slots, feats, nf, centers, poses, 5, 0, 8, log_cap=64)
    for i in range(n):
        sc = scs[i % 3]
        packed, offs, ws, hs = sc["ref_pyr"]
        po, no, lo = O.sparse_align(H.ocam(cam), packed, sc["cur_pyr"][0], offs, ws, hs, sc["feats"], sc["ref_center"], poses[i], 5, 0, 8)
        d = S.pose_dist(po, pb[i])
        assert d[0] < 1e-5 and d[1] < 1e-5 and no == tb[i]
        assert len(lo) == nlog[i] and all(abs(a["chi2"] - b["chi2"]) <= 1e-4 * abs(a["chi2"]) for a, b in zip(lo, log[i]))
        e = S.pose_dist(pb[i], sc["T_c2r"])
        assert e[0] < 3e-4 and e[1] < 1e-3
    # size-independent properties (what the full 4096-pair sweep is checked with): permutation invariance and determinism
    perm = rng.permutation(n)
    pp, tp, _, _ = ctx.sparse_align_batch(ref_slots[perm], cur_slots[perm], feats[perm], nf[perm], centers[perm], poses[perm], 5, 0, 8)
    assert (pp == pb[perm]).all() and (tp == tb[perm]).all()
    p2, t2, _, _ = ctx.sparse_align_batch(ref_slots, cur_slots, feats, nf, centers, poses, 5, 0, 8)
    assert (p2 == pb).all()
    ctx.close()


def test_front_end_sequence_with_second_keyframe(built):
    """Same loop as above, plus Tracking::CraeteKeyframe in the middle (ref: src/Tracking.cpp:412-464): Set_ExistingFeatures +
    detect on a frame that already carries matched features (occupancy grid, Frame::Set_Mask with Min_dist circles), new map
    points, a second keyframe, and afterwards a local map of two keyframes where MapPoint::Get_ClosetObs picks the
    observation by viewing angle."""
    _run_keyframe_sequence(8, (4,), seed=9)


@pytest.mark.slow
def test_front_end_sequence_100_frames_five_keyframes(built):
    """BASELINE configs[3] at a length where the map matters: 100 frames, a new key frame every 20 frames (five
    besides the initial one), every frame compared with the oracle-side model: pose after Run, match list of SearchLocalPoints over
    the growing local map (ids, levels, refined pixels), new corners at every key frame."""
    _run_keyframe_sequence(100, (20, 40, 60, 80, 95), seed=9, step_scale=0.25)


def _run_keyframe_sequence(n_frames, kf_frames, seed, step_scale=1.0):
    cam = dict(S.KINECT)
    oc = H.ocam(cam)
    scene = S.Scene(91)
    kf_at = kf_frames[0]
    poses = _trajectory(n_frames, seed=seed, scale=step_scale)
    cam_h = HL.configure(cam, max_fts=300, max_frames=max(24, 2 * len(kf_frames) + 8))
    cfg = (5, 0, 8)
    omap = OMap()

    img0, _, pts0 = S.render(scene, cam, poses[0], want_points=True)
    g0 = HL.HFrame(cam_h, img0, poses[0]); o0 = OracleFrame(img0, poses[0])
    assert g0.detect(5.0) == 300
    corners, _ = H.detect_oracle(img0, LEVELS, CELL, 300)
    g0.attach_points(pts0[corners["y"], corners["x"]], np.ones(len(corners), np.uint8))
    o0.feats = H.ref_feats_from_corners(cam, corners, pts0)
    o0.feat_mp = [omap.new_mp(o0.feats[i]["point_w"]) for i in range(len(corners))]
    for i, mp in enumerate(o0.feat_mp):
        omap.mp_obs[mp].append((0, i))
    omap.kfs.append(o0)
    kf_handles = [HL.lib().hs_keyframe_new(g0.h)]

    g_last, o_last = g0, o0
    for k in range(1, n_frames):
        img, _, pts = S.render(scene, cam, poses[k], want_points=True)
        g_cur = HL.HFrame(cam_h, img, g_last.pose()); o_cur = OracleFrame(img, o_last.pose)
        n_g, pose_g, _ = HL.sparse_align_run(*cfg, g_cur, g_last)
        pose_o, n_o = _oracle_sparse_align(oc, o_cur, o_last, cfg)
        o_cur.pose = pose_o
        d = S.pose_dist(pose_o, pose_g)
        assert d[0] < 1e-5 and d[1] < 1e-5 and n_g == n_o and n_g >= 20, (k, d, n_g, n_o)
        m, nrep = HL.search_local_points_multi(cam_h, g_cur, kf_handles)
        want = _oracle_search_multi(oc, cam, o_cur, omap, list(range(len(omap.kfs))))
        px, lv, ini = g_cur.features()
        ids = g_cur.mp_ids()
        assert m == len(want) == len(px) and m > 100, (k, m, len(want))
        for (mp, p, SL, _), gp, gl, gid in zip(want, px, lv, ids):
            assert gid == mp and gl == SL and np.abs(p - gp).max() <= 1e-3
        F = np.zeros(len(want), O.REF_FEAT_DT)
        for j, (mp, p, SL, _) in enumerate(want):
            omap.mp_found[mp] += 1
            F[j]["px"] = px[j]; F[j]["level"] = SL; F[j]["initial"] = 1
            F[j]["normal"] = O.feature_normal(oc, px[j]); F[j]["point_w"] = omap.mp_point[mp]
        o_cur.feats = F
        o_cur.feat_mp = [w_[0] for w_ in want]
        if k >= kf_at + 1 and len(kf_frames) == 1:
            assert len({w_[3] for w_ in want}) == 2, "both keyframes should serve as closest observation somewhere"

        if k in kf_frames:
            # ---- CraeteKeyframe: Set_ExistingFeatures(cur features) + detect(cur, 5.0), then map points for the new features
            n_old = len(px)
            n_all = g_cur.detect(5.0, use_existing=True)
            px2, lv2, ini2 = g_cur.features()
            # oracle: occupancy from the existing features (ref: src/Feature_detection.cpp:43-50), Set_Mask circles of Min_dist at
            # features WITH map points (ref: src/Frame.cpp:286-298), then the usual sort + selection
            packed, offs, ws, hs = o_cur.pyr
            rows_, cols_ = O.grid_dims(cam["width"], cam["height"], CELL)
            occ = np.zeros(rows_ * cols_, np.uint8)
            for p in px:
                occ[int(p[1] / np.float32(CELL)) * cols_ + int(p[0] / np.float32(CELL))] = 1
            cells = O.detect_cells(packed, offs, ws, hs, CELL, occ, 5.0)
            mask = np.full((cam["height"], cam["width"]), 255, np.uint8)
            for p in px:
                O.circle_fill(mask, O.cvround(p[0]), O.cvround(p[1]), 15, 0)          # Camera.Min_dist = 15
            new_c, _ = O.detect_select(cells, mask, CELL, 300, n_existing=n_old)
            assert n_all == n_old + len(new_c) and n_all > n_old
            assert (px2[n_old:, 0] == new_c["x"]).all() and (px2[n_old:, 1] == new_c["y"]).all() and (lv2[n_old:] == new_c["level"]).all()
            newpts = pts[new_c["y"], new_c["x"]]
            g_cur.attach_points_from(n_old, newpts, np.ones(len(new_c), np.uint8))
            Fn = H.ref_feats_from_corners(cam, new_c, pts)
            o_cur.feats = np.concatenate([o_cur.feats, Fn])
            new_ids = [omap.new_mp(Fn[i]["point_w"]) for i in range(len(new_c))]
            o_cur.feat_mp = o_cur.feat_mp + new_ids
            kfi = len(omap.kfs)
            for i, mp in enumerate(o_cur.feat_mp):
                omap.mp_obs[mp].append((kfi, i))
            omap.kfs.append(o_cur)
            kf_handles.append(HL.lib().hs_keyframe_new(g_cur.h))
            assert (g_cur.mp_ids() == np.array(o_cur.feat_mp)).all()
        g_last, o_last = g_cur, o_cur


def test_frame_pool_too_small_fails_loudly(built):
    """A pool smaller than (local keyframes + current frame) would recycle a slot that an earlier candidate of the same batch
    points to. The adapter must refuse (never sample the wrong image silently)."""
    cam = dict(S.KINECT)
    scene = S.Scene(5)
    poses = _trajectory(3, seed=2)
    cam_h = HL.configure(cam, max_fts=120, max_frames=2)
    kfs = []
    frames = []
    for k in range(2):
        img, _, pts = S.render(scene, cam, poses[k], want_points=True)
        g = HL.HFrame(cam_h, img, poses[k])
        n = g.detect(5.0)
        px, lv, _ = g.features()
        g.attach_points(pts[px[:, 1].astype(int), px[:, 0].astype(int)], np.ones(n, np.uint8))
        kfs.append(HL.lib().hs_keyframe_new(g.h)); frames.append(g)
    img, _ = S.render(scene, cam, poses[2])
    cur = HL.HFrame(cam_h, img, poses[2])
    with pytest.raises(RuntimeError, match="frame pool too small"):
        HL.search_local_points_multi(cam_h, cur, kfs)


# ------------------------------------------------------------------------------------------------ RGB-D sequence with the key-frame lift
@pytest.mark.gpu
def test_rgbd_sequence_with_device_keyframe_lift(built):
    """A 25-frame RGB-D run through the adapter classes in which every map point comes from the device key-frame path
    (ref: src/Tracking.cpp:412-464: detect -> UndistortFeatures -> Get_FeatureDetph -> UnProject): TUM-style 16-bit depth
    (scale 5000), key frames every 8 frames, tracked pose checked against the ground-truth trajectory and the lifted points
    against the analytic scene. No oracle here: this is the functional end-to-end check of SURVEY 8f-3 inside the loop; the
    per-call parity tests are in test_gpu_ingest.py / test_host_adapters.py."""
    cam = dict(S.KINECT)
    scene = S.Scene(123)
    n_frames, kf_every, scale = 25, 8, 5000.0
    poses = _trajectory(n_frames, seed=21)
    cam_h = HL.configure(cam, max_fts=300, max_frames=24, dist=(0.0, 0.0, 0.0, 0.0, 0.0))
    cfg = (5, 0, 8)

    def depth16(z):
        return np.clip(np.rint(z * scale), 0, 65535).astype(np.uint16)

    def lift(frame, z, pts, start):
        """CraeteKeyframe on `frame`: device lift of all features, map points for the new ones (index >= start) with a depth."""
        tab, _ = frame.keyframe_lift(depth16(z), scale)
        px, lv, ini = frame.features()
        new = np.arange(start, len(px))
        has = (tab[new, 2] > 0).astype(np.uint8)
        assert has.sum() == len(new)                                     # the synthetic depth has no holes
        want = pts[np.rint(px[new, 1]).astype(int), np.rint(px[new, 0]).astype(int)]
        got = tab[new, 3:6]
        assert np.abs(got - want).max() < 2e-3, np.abs(got - want).max()  # 16-bit depth quantisation (0.2 mm) + pixel-centre sampling
        frame.attach_points_from(int(start), got, has)
        return len(new)

    img0, z0, pts0 = S.render(scene, cam, poses[0], want_points=True)
    g0 = HL.HFrame(cam_h, img0, poses[0])
    assert g0.detect(5.0) == 300
    assert lift(g0, z0, pts0, 0) == 300
    kf_handles = [HL.lib().hs_keyframe_new(g0.h)]
    g_last = g0
    n_kf = 1
    for k in range(1, n_frames):
        img, z, pts = S.render(scene, cam, poses[k], want_points=True)
        g_cur = HL.HFrame(cam_h, img, g_last.pose())
        n, pose, _ = HL.sparse_align_run(*cfg, g_cur, g_last)
        e = S.pose_dist(pose, poses[k])
        assert n >= 100 and e[0] < 1.5e-3 and e[1] < 4e-3, (k, n, e)      # tracks the ground truth (no drift beyond the per-frame noise)
        m, nrep = HL.search_local_points_multi(cam_h, g_cur, kf_handles)
        assert m >= 150, (k, m)
        if k % kf_every == 0:
            n_old = len(g_cur.features()[0])
            n_all = g_cur.detect(5.0, use_existing=True)
            assert n_all > n_old
            assert lift(g_cur, z, pts, n_old) == n_all - n_old
            kf_handles.append(HL.lib().hs_keyframe_new(g_cur.h))
            n_kf += 1
        g_last = g_cur
    assert n_kf == 4


# ------------------------------------------------------------------------------------------------ TrackWithLocalMap including the pose refinement
@pytest.mark.gpu
def test_sequence_with_pose_optimization_after_matching(built):
    """Tracking::TrackWithLocalMap's order on a 25-frame RGB-D run (ref: src/Tracking.cpp:199-236): Sprase_ImgAlign::Run ->
    SearchLocalPoints -> Optimizer::PoseOptimization, the refined pose handed to the next frame as in the reference.
    Functional end-to-end check against the ground-truth trajectory: the geometric refinement over the sub-pixel matches must not
    lose what the photometric alignment found (and on average tightens it); with LocalBAthreshhold = 2 px (the value of every
    reference config) no accurate match is charged. The per-call parity of the solve is in test_gpu_pose_opt.py and
    test_host_adapters.py."""
    cam = dict(S.KINECT)
    scene = S.Scene(321)
    n_frames, kf_every, scale = 25, 8, 5000.0
    poses = _trajectory(n_frames, seed=33)
    cam_h = HL.configure(cam, max_fts=300, max_frames=24, dist=(0.0, 0.0, 0.0, 0.0, 0.0))
    L = HL.lib()
    L.hs_config_set(b"Optimization.LocalBAthreshhold", b"2.0")
    cfg = (5, 0, 8)

    def lift(frame, z, start):
        tab, _ = frame.keyframe_lift(np.clip(np.rint(z * scale), 0, 65535).astype(np.uint16), scale)
        n = len(tab)
        new = np.arange(start, n)
        frame.attach_points_from(int(start), tab[new, 3:6], (tab[new, 2] > 0).astype(np.uint8))

    img0, z0, _ = S.render(scene, cam, poses[0], want_points=True)
    g_last = HL.HFrame(cam_h, img0, poses[0])
    assert g_last.detect(5.0) == 300
    lift(g_last, z0, 0)
    kf_handles = [L.hs_keyframe_new(g_last.h)]
    err_sa, err_po, its = [], [], []
    for k in range(1, n_frames):
        img, z, _ = S.render(scene, cam, poses[k], want_points=True)
        g_cur = HL.HFrame(cam_h, img, g_last.pose())
        n, pose_sa, _ = HL.sparse_align_run(*cfg, g_cur, g_last)
        assert n >= 100, (k, n)
        m, _ = HL.search_local_points_multi(cam_h, g_cur, kf_handles)
        assert m >= 150, (k, m)
        ids = g_cur.mp_ids()
        found_before = np.array([L.hs_mappoint_found(int(i)) for i in ids])
        pose_po, summ, res = g_cur.pose_optimization()
        assert len(res) == m and summ["n_obs"] == m
        assert summ["termination"] in (0, 1, 2) and summ["iterations"] <= 30, (k, summ)     # converged, not the iteration cap
        assert summ["final_cost"] <= summ["initial_cost"]
        assert res.max() < 2.0 / cam["f"]                                                     # every match within 2 px ...
        assert (np.array([L.hs_mappoint_found(int(i)) for i in ids]) == found_before).all()    # ... so EraseFound never ran
        assert np.abs(g_cur.pose() - pose_po).max() == 0                                      # Set_Pose happened
        e0, e1 = S.pose_dist(pose_sa, poses[k]), S.pose_dist(pose_po, poses[k])
        err_sa.append(e0); err_po.append(e1); its.append(int(summ["iterations"]))
        assert e1[0] < 1.5e-3 and e1[1] < 4e-3, (k, e0, e1)
        if k % kf_every == 0:
            n_old = len(g_cur.features()[0])
            assert g_cur.detect(5.0, use_existing=True) > n_old
            lift(g_cur, z, n_old)
            kf_handles.append(L.hs_keyframe_new(g_cur.h))
        g_last = g_cur
    err_sa, err_po = np.array(err_sa), np.array(err_po)
    # the refinement over ~190 sub-pixel matches keeps (on average tightens) the photometric estimate
    print("mean pose error (rad, m): sparse alignment", err_sa.mean(0), "after PoseOptimization", err_po.mean(0), "LM iterations", np.mean(its))
    # measured: 1.6e-5 rad / 3.4e-5 m after the refinement, 3.1e-5 / 7.1e-5 before
    assert err_po[:, 0].mean() <= err_sa[:, 0].mean() and err_po[:, 1].mean() <= err_sa[:, 1].mean(), (err_sa.mean(0), err_po.mean(0))


# ------------------------------------------------------------------------------------------------ UpdateLocalMap with the device-side key-frame selection
@pytest.mark.gpu
def test_update_local_map_selects_the_oracles_keyframes(built):
    """Tracking::UpdateLocalMap through the adapter (ref: src/Tracking.cpp:257-345): eight key frames strung along 4.2 m, the current
    frame near the third. GetCloseKeyFrames runs on the device-resident map table; the close set, the distance ranking and the 10-nearest
    cut must equal the oracle's, the map points of the chosen key frames are reprojected once each, and SearchLocalPoints matches
    against them. Then a key frame is moved (bundle adjustment) and the table row is rewritten."""
    cam = dict(S.KINECT)
    scene = S.Scene(77)
    cam_h = HL.configure(cam, max_fts=300, max_frames=24, dist=(0.0, 0.0, 0.0, 0.0, 0.0))
    L = HL.lib()
    oc = H.ocam(cam)
    kf_poses = [S.pose_from_xi(np.array([0.6 * k, 0.05 * (k % 3), 0.02 * k, 0.0, 0.01 * k, 0.0])) for k in range(8)]
    kfs, frames, tables = [], [], []
    for k, pose in enumerate(kf_poses):
        img, _, pts = S.render(scene, cam, pose, want_points=True)
        g = HL.HFrame(cam_h, img, pose)
        n = g.detect(5.0)
        px, _, _ = g.features()
        P = pts[px[:, 1].astype(int), px[:, 0].astype(int)]
        has = np.ones(n, np.uint8); has[::9] = 0                      # some features without a map point: zero rows of the table
        g.attach_points(P, has)
        kf = L.hs_keyframe_new(g.h)
        assert L.hs_map_add_keyframe(kf) == k + 1
        kfs.append(kf); frames.append(g)
        tables.append(np.where(has[:, None] > 0, P, 0.0))
    pt_count = np.array([len(t) for t in tables], np.int32)
    pt_begin = np.concatenate([[0], np.cumsum(pt_count)[:-1]]).astype(np.int32)
    kf_t = np.array([p[4:] for p in kf_poses])
    points = np.concatenate(tables)
    pose_cur = S.pose_mul(S.pose_from_xi(np.array([0.012, -0.008, 0.006, 0.002, -0.003, 0.001])), kf_poses[2])
    img, _ = S.render(scene, cam, pose_cur)
    cur = HL.HFrame(cam_h, img, pose_cur)
    m, local, nrep = HL.track_local_map(cam_h, cur)
    vo, do, lo = O.close_keyframes(oc, pose_cur, pt_begin, pt_count, kf_t, points)
    assert (local == lo).all() and 2 <= len(lo) < 8 and lo[0] == 2               # the far key frames see none of the current view
    assert m >= 150 and nrep > 300
    # every reprojected point belongs to a chosen key frame (the others were never walked)
    ids = cur.mp_ids()
    first_id = np.concatenate([[0], np.cumsum([int((np.abs(t).sum(1) > 0).sum()) for t in tables])])
    owner = np.searchsorted(first_id, ids, side="right") - 1
    assert set(owner.tolist()) <= set(lo.tolist())
    # the device path (dsdtm_store_track, default) against the literal host loop of the reference (one ReprojectPoint per map point,
    # snapshot of the observations per candidate): same local key frames, same matches in the same order, same pixels, same levels
    px_a, lv_a, _ = cur.features()
    for i in ids:
        L.hs_mappoint_increase_found(int(i), -1)                      # undo the first run's IncreaseFound: same sort keys for the second
    L.hs_set_use_store(0)
    try:
        cur_b = HL.HFrame(cam_h, img, pose_cur)
        m_b, local_b, nrep_b = HL.track_local_map(cam_h, cur_b)
    finally:
        L.hs_set_use_store(1)
    px_b, lv_b, _ = cur_b.features()
    assert m_b == m and nrep_b == nrep and (local_b == local).all()
    assert (cur_b.mp_ids() == ids).all() and (lv_b == lv_a).all() and (px_b == px_a).all()
    for i in ids:
        L.hs_mappoint_increase_found(int(i), -1)
    # LocalBundleAdjustment moves the nearest key frame far away: its row is rewritten, it drops out of the ranking
    moved = kf_poses[2].copy(); moved[4:] += np.array([50.0, 0.0, 0.0])
    L.hs_keyframe_set_pose(kfs[2], HL._p(np.ascontiguousarray(moved)))
    L.hs_map_mark_moved(kfs[2])
    cur2 = HL.HFrame(cam_h, img, pose_cur)
    m2, local2, _ = HL.track_local_map(cam_h, cur2)
    kf_t2 = kf_t.copy(); kf_t2[2] = moved[4:]
    _, _, lo2 = O.close_keyframes(oc, pose_cur, pt_begin, pt_count, kf_t2, points)
    assert (local2 == lo2).all() and local2[-1] == 2 and m2 >= 100


@pytest.mark.gpu
def test_map_store_follows_recycled_slots_bad_flags_and_moved_points(built):
    """The device-resident map tables are maintained incrementally (Map::SyncStore): key frames whose pyramids were evicted from the
    frame pool come back in OTHER slots (dsdtm_store_set_keyframe), map points flagged bad or moved by a bundle adjustment are
    rewritten in place (dsdtm_store_update_points). After each kind of change the device path must still equal the reference's literal
    per-object loop (ref: src/Tracking.cpp:283-297, src/MapPoint.cpp:133-174) run on the same host objects."""
    cam = dict(S.KINECT)
    scene = S.Scene(77)
    cam_h = HL.configure(cam, max_fts=300, max_frames=12, dist=(0.0, 0.0, 0.0, 0.0, 0.0))
    L = HL.lib()
    kf_poses = [S.pose_from_xi(np.array([0.3 * k, 0.04 * (k % 3), 0.02 * k, 0.0, 0.01 * k, 0.0])) for k in range(8)]
    frames = []
    for k, pose in enumerate(kf_poses):
        img, _, pts = S.render(scene, cam, pose, want_points=True)
        g = HL.HFrame(cam_h, img, pose)
        n = g.detect(5.0)
        px, _, _ = g.features()
        g.attach_points(pts[px[:, 1].astype(int), px[:, 0].astype(int)], np.ones(n, np.uint8))
        L.hs_map_add_keyframe(L.hs_keyframe_new(g.h))
        frames.append(g)
    pose_cur = S.pose_mul(S.pose_from_xi(np.array([0.012, -0.008, 0.006, 0.002, -0.003, 0.001])), kf_poses[3])
    img, _ = S.render(scene, cam, pose_cur)

    def track(use_store):
        L.hs_set_use_store(1 if use_store else 0)
        try:
            cur = HL.HFrame(cam_h, img, pose_cur)
            m, local, nrep = HL.track_local_map(cam_h, cur)
        finally:
            L.hs_set_use_store(1)
        px, lv, _ = cur.features()
        ids = cur.mp_ids()
        for i in ids:
            L.hs_mappoint_increase_found(int(i), -1)                  # every run starts from the same found counters
        cur.free()
        return m, local, nrep, px, lv, ids

    def same(a, b):
        return a[0] == b[0] and (a[1] == b[1]).all() and a[2] == b[2] and (a[3] == b[3]).all() and (a[4] == b[4]).all() and (a[5] == b[5]).all()

    base = track(True)
    assert base[0] >= 150 and len(base[1]) >= 3 and same(base, track(False))
    # 1. ten frames that stay alive push the oldest key-frame pyramids out of the 12-slot pool; once they are gone the key frames are
    #    re-uploaded into whatever slots are free: the rows of the table must follow
    junk = [HL.HFrame(cam_h, img, pose_cur) for _ in range(10)]
    for j in junk:
        j.free()
    again = track(True)
    assert same(base, again)
    # 2. a fifth of the matched points is flagged bad, another fifth moves by 1-2 cm (LocalBundleAdjustment): the bad ones disappear
    #    from the candidates, the moved ones are projected from their new positions -- in both paths alike
    ids = base[5]
    bad_ids, moved_ids = ids[::5], ids[2::5]
    for i in bad_ids:
        L.hs_mappoint_set_bad(int(i), 1)
    rng = np.random.default_rng(3)
    for i in moved_ids:
        p = np.zeros(3)
        L.hs_mappoint_pose(int(i), HL._p(p))
        p += rng.uniform(-0.02, 0.02, 3)
        L.hs_mappoint_set_pose(int(i), HL._p(np.ascontiguousarray(p)))
    dev, host = track(True), track(False)
    assert same(dev, host)
    assert dev[2] < base[2] and not (set(dev[5].tolist()) & set(int(i) for i in bad_ids))
    assert dev[0] >= 100
    # 3. the flags are cleared again: the first result comes back except for the moved points
    for i in bad_ids:
        L.hs_mappoint_set_bad(int(i), 0)
    dev2, host2 = track(True), track(False)
    assert same(dev2, host2) and dev2[2] > dev[2] and abs(dev2[2] - base[2]) <= len(moved_ids)


def test_run_and_update_local_map_as_one_submission_equal_the_separate_calls(built):
    """Sprase_ImgAlign::Run also runs Tracking::UpdateLocalMap's device stage in the same submission (dsdtm_track_frame_store) and
    Tracking::UpdateLocalMap takes the parked records: poses, tracked counts, local key frames, matches, pixels, levels and map-point
    ids must be bit-equal to the two separate calls, frame after frame (the found counters evolve identically)."""
    cam = dict(S.KINECT)
    scene = S.Scene(77)
    L = HL.lib()
    kf_poses = [S.pose_from_xi(np.array([0.25 * k, 0.03 * (k % 3), 0.01 * k, 0.0, 0.01 * k, 0.0])) for k in range(4)]
    cur_poses = [S.pose_mul(S.pose_from_xi(np.array([0.01 * (j + 1), -0.006, 0.004, 0.002, -0.002 * j, 0.001])), kf_poses[1]) for j in range(4)]
    runs = []
    for spec in (1, 0):
        cam_h = HL.configure(cam, max_fts=300, max_frames=24, dist=(0.0, 0.0, 0.0, 0.0, 0.0))       # hs_reset inside: fresh map, fresh points
        L.hs_set_speculate(spec)
        try:
            frames = []
            for k, pose in enumerate(kf_poses):
                img, _, pts = S.render(scene, cam, pose, want_points=True)
                g = HL.HFrame(cam_h, img, pose)
                n = g.detect(5.0)
                px, _, _ = g.features()
                g.attach_points(pts[px[:, 1].astype(int), px[:, 0].astype(int)], np.ones(n, np.uint8))
                L.hs_map_add_keyframe(L.hs_keyframe_new(g.h))
                frames.append(g)
            g_last = frames[1]
            rec = []
            img0, _ = S.render(scene, cam, cur_poses[0])
            warm = HL.HFrame(cam_h, img0, g_last.pose())
            HL.track_local_map(cam_h, warm)                                # creates the Tracking object (registers the map) in both runs
            for i in warm.mp_ids():
                L.hs_mappoint_increase_found(int(i), -1)
            for j, pose in enumerate(cur_poses):
                img, _ = S.render(scene, cam, pose)
                g_cur = HL.HFrame(cam_h, img, g_last.pose())
                nt, pose_sa, _ = HL.sparse_align_run(5, 0, 8, g_cur, g_last, want_log=False)
                m, local, nrep = HL.track_local_map(cam_h, g_cur)
                px, lv, _ = g_cur.features()
                rec.append((nt, pose_sa.copy(), m, local.copy(), nrep, px.copy(), lv.copy(), g_cur.mp_ids().copy()))
                g_last = g_cur
            runs.append(rec)
        finally:
            L.hs_set_speculate(1)
    for a, b in zip(*runs):
        assert a[0] == b[0] and (a[1] == b[1]).all() and a[2] == b[2] and (a[3] == b[3]).all() and a[4] == b[4]
        assert (a[5] == b[5]).all() and (a[6] == b[6]).all() and (a[7] == b[7]).all()
        assert a[2] >= 100 and a[0] >= 100
