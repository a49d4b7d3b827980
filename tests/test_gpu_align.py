"""GPU parity (tolerances of BASELINE.json north_star): kernel (c) sparse alignment, kernel (d) Align2D, WarpAffine."""
import numpy as np
import pytest

import helpers as H
import oracle as O
from dsdtm_b200 import synth as S

pytestmark = pytest.mark.gpu

POSE_TOL = 1e-5        # rad and m (north_star)
CHI2_RTOL = 1e-4       # per-iteration chi2, relative (north_star)
PX_TOL = 1e-3          # refined feature position, px (north_star)


def _upload(ctx, sc, ref_slot=0, cur_slot=1):
    ctx.upload(ref_slot, sc["ref_img"])
    ctx.upload(cur_slot, sc["cur_img"])


def _compare_traces(lo, lg):
    """Iteration traces agree within tolerance; a +-1 iteration difference is tolerated only at stagnation (chi2 equal to
    ~1e-13), which did not occur on any seed so far -- so we assert equality of the structure."""
    assert len(lo) == len(lg)
    for a, b in zip(lo, lg):
        assert (a["level"], a["iter"], a["n_pts"], a["flags"]) == (b["level"], b["iter"], b["n_pts"], b["flags"])
        assert abs(a["chi2"] - b["chi2"]) <= CHI2_RTOL * abs(a["chi2"])
        assert np.allclose(a["x"], b["x"], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("cfg", [(4, 0, 30), (5, 0, 8), (5, 2, 8), (3, 1, 4)])
def test_sparse_align_matches_oracle(ctx, scenario, cfg):
    """configs[0]: one synthetic 640x480 pair, kinect intrinsics; ctor (4,0,30) of Test/test_SpraseImg_alignment.cpp:110 and
    the production (5,0,8) of src/Tracking.cpp:37."""
    _upload(ctx, scenario)
    packed, offs, ws, hs = scenario["ref_pyr"]
    ml, mn, it = cfg
    po, no, lo = O.sparse_align(H.ocam(scenario["cam"]), packed, scenario["cur_pyr"][0], offs, ws, hs, scenario["feats"],
                                scenario["ref_center"], S.IDENTITY, ml, mn, it)
    pg, ng, lg = ctx.sparse_align(0, 1, scenario["feats"], scenario["ref_center"], S.IDENTITY, ml, mn, it)
    d = S.pose_dist(po, pg)
    assert d[0] < POSE_TOL and d[1] < POSE_TOL and no == ng
    _compare_traces(lo, lg)
    if cfg[:2] == (4, 0) or cfg[:2] == (5, 0):
        e = S.pose_dist(pg, scenario["T_c2r"])
        assert e[0] < 2e-4 and e[1] < 5e-4          # converged to the ground-truth motion


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_sparse_align_other_seeds_and_nonidentity_ref(ctx, seed):
    sc = H.make_scenario(seed, trans=0.03, rot_deg=0.8)
    _upload(ctx, sc, 2, 3)
    packed, offs, ws, hs = sc["ref_pyr"]
    start = S.pose_from_xi(np.random.default_rng(seed).uniform(-0.002, 0.002, 6))     # not exactly identity
    po, no, lo = O.sparse_align(H.ocam(sc["cam"]), packed, sc["cur_pyr"][0], offs, ws, hs, sc["feats"], sc["ref_center"], start, 5, 0, 8)
    pg, ng, lg = ctx.sparse_align(2, 3, sc["feats"], sc["ref_center"], start, 5, 0, 8)
    d = S.pose_dist(po, pg)
    assert d[0] < POSE_TOL and d[1] < POSE_TOL and no == ng
    _compare_traces(lo, lg)


@pytest.mark.parametrize("opt", [("sa_variant", 1), ("sa_warps_per_pair", 1), ("sa_warps_per_pair", 2), ("sa_warps_per_pair", 3), ("sa_warps_per_pair", 4), ("sa_warps_per_pair", 5)])
def test_sparse_align_kernel_variants_agree_with_oracle(ctx, scenario, opt):
    """Tuning knobs never change results beyond the reduction-order tolerance: the L2-workspace variant and every
    warps-per-pair instantiation against the oracle."""
    _upload(ctx, scenario)
    packed, offs, ws, hs = scenario["ref_pyr"]
    po, no, lo = O.sparse_align(H.ocam(scenario["cam"]), packed, scenario["cur_pyr"][0], offs, ws, hs, scenario["feats"],
                                scenario["ref_center"], S.IDENTITY, 5, 0, 8)
    ctx.set_option(*opt)
    try:
        pg, ng, lg = ctx.sparse_align(0, 1, scenario["feats"], scenario["ref_center"], S.IDENTITY, 5, 0, 8)
    finally:
        ctx.set_option("sa_variant", 0)
        ctx.set_option("sa_warps_per_pair", 0)
    d = S.pose_dist(po, pg)
    assert d[0] < POSE_TOL and d[1] < POSE_TOL and no == ng
    _compare_traces(lo, lg)


def test_sparse_align_skips_uninitialised_zero_and_border_features(ctx, scenario):
    """ref: src/Sprase_ImageAlign.cpp:86,95-100 -- mbInitial false, P_w == 0 and features within 3 px of the level border."""
    _upload(ctx, scenario)
    F = scenario["feats"].copy()
    F["initial"][::7] = 0
    F["point_w"][3::11] = 0.0
    F["px"][5] = (2.0, 100.0); F["px"][6] = (638.0, 100.0); F["px"][8] = (100.0, 477.5)
    packed, offs, ws, hs = scenario["ref_pyr"]
    po, no, lo = O.sparse_align(H.ocam(scenario["cam"]), packed, scenario["cur_pyr"][0], offs, ws, hs, F, scenario["ref_center"], S.IDENTITY, 4, 0, 30)
    pg, ng, lg = ctx.sparse_align(0, 1, F, scenario["ref_center"], S.IDENTITY, 4, 0, 30)
    d = S.pose_dist(po, pg)
    assert d[0] < POSE_TOL and d[1] < POSE_TOL and no == ng and no < 270
    _compare_traces(lo, lg)


def test_sparse_align_many_features_and_other_geometry(built):
    """Camera.Max_fts = 500 (Config/kinect.yaml:68) on EuRoC geometry: several features per lane in every kernel variant,
    capacity 512, 4-level pyramid, cell size 30."""
    from dsdtm_b200 import capi
    cam = dict(S.EUROC)
    sc = H.make_scenario(41, cam, levels=4, max_fts=500)
    assert len(sc["feats"]) > 320
    c = capi.Context(cam, levels=4, cell_size=30, max_feats=512, max_patches=8, max_frames=2, max_batch=1)
    c.upload(0, sc["ref_img"]); c.upload(1, sc["cur_img"])
    packed, offs, ws, hs = sc["ref_pyr"]
    po, no, lo = O.sparse_align(H.ocam(cam), packed, sc["cur_pyr"][0], offs, ws, hs, sc["feats"], sc["ref_center"], S.IDENTITY, 4, 0, 30)
    for wpp in (0, 1, 4):
        c.set_option("sa_warps_per_pair", wpp)
        pg, ng, lg = c.sparse_align(0, 1, sc["feats"], sc["ref_center"], S.IDENTITY, 4, 0, 30)
        d = S.pose_dist(po, pg)
        assert d[0] < POSE_TOL and d[1] < POSE_TOL and no == ng, (wpp, d)
        _compare_traces(lo, lg)
    with pytest.raises(capi.DsdtmError):                      # more features than the capacity is an error, never a silent truncation
        c.sparse_align(0, 1, np.zeros(600, O.REF_FEAT_DT), sc["ref_center"], S.IDENTITY, 4, 0, 30)
    # a SECOND context with a smaller feature table must not shrink what this one may launch (the shared-memory limit is a property
    # of the kernel, shared by every context of the process)
    small = capi.Context(cam, levels=4, cell_size=30, max_feats=64, max_patches=8, max_frames=2, max_batch=1)
    c.set_option("sa_warps_per_pair", 0)
    pg, ng, _ = c.sparse_align(0, 1, sc["feats"], sc["ref_center"], S.IDENTITY, 4, 0, 30)
    d = S.pose_dist(po, pg)
    assert d[0] < POSE_TOL and d[1] < POSE_TOL and no == ng
    small.close()
    c.close()


def test_sparse_align_levels_too_small_for_any_feature(built):
    """A 160x96 image with five levels ends in a 10x6 level: no feature passes the 3-pixel border test there (ref:
    src/Sprase_ImageAlign.cpp:95-96), so the level staging stages nothing -- its unconditional window loads must stay inside the level
    (the reference frame sits in the LAST slot of the pool here) -- chi2 is NaN (Q2) and the pose moves on to the next level unchanged."""
    from dsdtm_b200 import capi
    cam = dict(width=160, height=96, fx=130.0, fy=130.0, cx=79.5, cy=47.5, f=130.0)
    sc = H.make_scenario(5, cam, levels=5, max_fts=150, trans=0.004, rot_deg=0.1)
    c = capi.Context(cam, levels=5, cell_size=15, max_feats=160, max_patches=8, max_frames=2, max_batch=1)
    c.upload(1, sc["ref_img"]); c.upload(0, sc["cur_img"])
    packed, offs, ws, hs = sc["ref_pyr"]
    assert hs[4] < 7
    po, no, lo = O.sparse_align(H.ocam(cam), packed, sc["cur_pyr"][0], offs, ws, hs, sc["feats"], sc["ref_center"], S.IDENTITY, 5, 0, 8)
    for wpp in (0, 1, 3):
        c.set_option("sa_warps_per_pair", wpp)
        pg, ng, lg = c.sparse_align(1, 0, sc["feats"], sc["ref_center"], S.IDENTITY, 5, 0, 8)
        pg2, ng2, lg2 = c.sparse_align(1, 0, sc["feats"], sc["ref_center"], S.IDENTITY, 5, 0, 8)
        assert (pg == pg2).all() and ng == ng2 and np.isfinite(pg).all()
        assert lg[0]["level"] == 4 and lg[0]["n_pts"] == 0 == lo[0]["n_pts"] and np.isnan(lg[0]["chi2"]) and lg[0]["flags"] == lo[0]["flags"]
        first3 = [e for e in lg if e["level"] == 3][0]; ofirst3 = [e for e in lo if e["level"] == 3][0]
        assert first3["n_pts"] == ofirst3["n_pts"] and abs(first3["chi2"] - ofirst3["chi2"]) <= 1e-9 * abs(ofirst3["chi2"])
    c.close()


def test_sparse_align_subnormal_byte_encoding_keeps_full_precision(ctx):
    """The kernel feeds image bytes to the fp64 arithmetic as subnormals b * 2^-1034 with weights scaled by 2^1010 (csrc/sparse_align.cu,
    DSDTM_SA_CVT 3); the claim is that every sample keeps the reference's rounding. Stress it where a loss would show: images of
    extreme bytes (0 / 255 blocks, so that products with tiny weights sit next to full-scale ones), refined (non-dyadic float) feature
    positions, a start pose with arbitrary fractions. Agreement with the oracle must be at summation-order level (1e-12), six orders
    tighter than the north_star tolerances, at every logged iteration."""
    sc = H.make_scenario(11, trans=0.015, rot_deg=0.4)
    rng = np.random.default_rng(5)
    hard = {}
    for k in ("ref_img", "cur_img"):
        img = sc[k].copy()
        img[(img > 150)] = 255
        img[(img < 90)] = 0
        hard[k] = img
    ctx.upload(4, hard["ref_img"])
    ctx.upload(5, hard["cur_img"])
    ref_pyr = O.pyramid(hard["ref_img"], 5)
    cur_pyr = O.pyramid(hard["cur_img"], 5)
    feats = sc["feats"].copy()
    feats["px"] = (feats["px"] + rng.uniform(0.0, 1.0, feats["px"].shape)).astype(np.float32)      # refined positions: arbitrary float fractions
    start = S.pose_from_xi(rng.uniform(-0.003, 0.003, 6))
    packed, offs, ws, hs = ref_pyr
    po, no, lo = O.sparse_align(H.ocam(sc["cam"]), packed, cur_pyr[0], offs, ws, hs, feats, sc["ref_center"], start, 4, 0, 30)
    pg, ng, lg = ctx.sparse_align(4, 5, feats, sc["ref_center"], start, 4, 0, 30)
    assert no == ng and len(lo) == len(lg)
    assert np.abs(po - pg).max() < 1e-12
    for a, b in zip(lo, lg):
        assert (a["level"], a["iter"], a["n_pts"], a["flags"]) == (b["level"], b["iter"], b["n_pts"], b["flags"])
        assert abs(a["chi2"] - b["chi2"]) <= 1e-12 * abs(a["chi2"])


def test_sparse_align_rank_deficient_system_is_handled(ctx, scenario):
    """Two visible features give a rank-4 H. In floating point the two 'zero' pivots come out as rounding noise (1e-17
    relative), Eigen's ldlt().solve divides by them (its zero rule only triggers below 5e-309), and the step is dominated
    by that noise IN THE REFERENCE ITSELF: it changes with the summation order, so no implementation can match it bit for
    bit. What must hold: the SPD fast path rejects the system and hands it to the pivoted LDL^T (no NaN, no crash), the
    pose-independent quantities agree (visible count, first chi2), and the result is deterministic."""
    _upload(ctx, scenario)
    F = scenario["feats"][:2].copy()
    packed, offs, ws, hs = scenario["ref_pyr"]
    po, no, lo = O.sparse_align(H.ocam(scenario["cam"]), packed, scenario["cur_pyr"][0], offs, ws, hs, F, scenario["ref_center"], S.IDENTITY, 2, 0, 4)
    pg, ng, lg = ctx.sparse_align(0, 1, F, scenario["ref_center"], S.IDENTITY, 2, 0, 4)
    pg2, ng2, lg2 = ctx.sparse_align(0, 1, F, scenario["ref_center"], S.IDENTITY, 2, 0, 4)
    assert lo[0]["n_pts"] == lg[0]["n_pts"] == 2
    assert abs(lo[0]["chi2"] - lg[0]["chi2"]) <= 1e-9 * lo[0]["chi2"]
    assert np.isfinite(pg).all() and np.isfinite(lg["x"]).all()
    assert (pg == pg2).all() and ng == ng2 and (lg["x"] == lg2["x"]).all()


def test_sparse_align_nothing_visible_nan_chi2(ctx, scenario):
    """Q2: chi2/tResNum is NaN when no feature is visible; the pose is left unchanged."""
    _upload(ctx, scenario)
    far = S.pose_from_xi([50.0, 0, 0, 0, 0, 0])
    pg, ng, lg = ctx.sparse_align(0, 1, scenario["feats"], scenario["ref_center"], far, 2, 0, 5)
    packed, offs, ws, hs = scenario["ref_pyr"]
    po, no, lo = O.sparse_align(H.ocam(scenario["cam"]), packed, scenario["cur_pyr"][0], offs, ws, hs, scenario["feats"], scenario["ref_center"], far, 2, 0, 5)
    assert ng == 0 == no and np.isnan(lg[0]["chi2"]) and np.allclose(pg, far) and len(lg) == len(lo)
    assert [int(e["flags"]) for e in lg] == [int(e["flags"]) for e in lo]


def test_sparse_align_batch_equals_single_and_is_deterministic(ctx):
    scs = [H.make_scenario(s) for s in (11, 12, 13, 14)]
    n = len(scs)
    for i, sc in enumerate(scs):
        ctx.upload(2 * i, sc["ref_img"]); ctx.upload(2 * i + 1, sc["cur_img"])
    stride = 320
    feats = np.zeros((n, stride), O.REF_FEAT_DT)
    nf = np.zeros(n, np.int32)
    for i, sc in enumerate(scs):
        nf[i] = len(sc["feats"]); feats[i, :nf[i]] = sc["feats"]
    centers = np.stack([sc["ref_center"] for sc in scs]); poses = np.tile(S.IDENTITY, (n, 1))
    ref_slots = np.arange(0, 2 * n, 2); cur_slots = ref_slots + 1
    p1, t1, log, nlog = ctx.sparse_align_batch(ref_slots, cur_slots, feats, nf, centers, poses, 5, 0, 8, log_cap=64)
    p2, t2, _, _ = ctx.sparse_align_batch(ref_slots, cur_slots, feats, nf, centers, poses, 5, 0, 8)
    assert (p1 == p2).all() and (t1 == t2).all()                     # bitwise run-to-run determinism
    for i, sc in enumerate(scs):
        packed, offs, ws, hs = sc["ref_pyr"]
        po, no, lo = O.sparse_align(H.ocam(sc["cam"]), packed, sc["cur_pyr"][0], offs, ws, hs, sc["feats"], sc["ref_center"], S.IDENTITY, 5, 0, 8)
        d = S.pose_dist(po, p1[i])
        assert d[0] < POSE_TOL and d[1] < POSE_TOL and no == t1[i]
        _compare_traces(lo, log[i][:nlog[i]])
        ps, ns, _ = ctx.sparse_align(2 * i, 2 * i + 1, sc["feats"], sc["ref_center"], S.IDENTITY, 5, 0, 8)
        assert (ps == p1[i]).all() and ns == t1[i]


@pytest.mark.parametrize("iters", [3, 10])
def test_align2d_300_patches(ctx, scenario, iters):
    """configs[1]: 300 synthetic 8x8 patches, 2-D LK refinement vs the reference; MaxIters 3 (test) and 10 (production)."""
    _upload(ctx, scenario)
    packed, offs, ws, hs = scenario["cur_pyr"]
    levels, patches, truth, start = H.make_patches(scenario["cur_pyr"], 300, 21, max_level=2)
    px, conv = ctx.align2d(1, levels, patches, start, iters)
    n_flag_diff = 0
    for i in range(300):
        p, c, _ = O.align2d(O.pyr_level(packed, offs, ws, hs, int(levels[i])), patches[i], iters, start[i])
        assert np.abs(p - px[i]).max() <= PX_TOL
        n_flag_diff += int(c != conv[i])
    assert n_flag_diff == 0          # may differ only when |d|^2 straddles 9e-4 within fp32 rounding; not on this data
    if iters == 10:
        err = np.linalg.norm(px - truth, axis=1)[conv]
        assert conv.sum() >= 290 and np.median(err) < 0.05


def test_align2d_border_nan_and_q4_semantics(ctx, scenario):
    """u_r == cols-4 / v_r == rows-4 iterate (Q4) with linear addressing + zero past the end; NaN and out-of-range break."""
    _upload(ctx, scenario)
    packed, offs, ws, hs = scenario["cur_pyr"]
    rng = np.random.default_rng(4)
    starts = np.array([[636.3, 200.2], [300.5, 476.4], [636.0, 476.0], [3.9, 100.0], [100.0, 3.2], [637.2, 50.0], [50.0, 477.1],
                       [np.nan, 100.0], [1e9, 5.0], [4.0, 4.0], [-5.0, 30.0]])
    n = len(starts)
    patches = rng.integers(0, 256, (n, 100), dtype=np.uint8)
    levels = np.zeros(n, np.int32)
    px, conv = ctx.align2d(1, levels, patches, starts, 10)
    img = O.pyr_level(packed, offs, ws, hs, 0)
    for i in range(n):
        p, c, _ = O.align2d(img, patches[i], 10, starts[i])
        assert c == conv[i], i
        assert np.allclose(p, px[i], atol=PX_TOL, equal_nan=True), (i, p, px[i])


def test_align2d_unused_entries_and_level_checks(ctx, scenario):
    from dsdtm_b200 import capi
    _upload(ctx, scenario)
    levels, patches, truth, start = H.make_patches(scenario["cur_pyr"], 8, 2)
    levels[3] = -1
    px, conv = ctx.align2d(1, levels, patches, start, 10)
    assert not conv[3] and (px[3] == start[3]).all()
    levels[3] = 9
    with pytest.raises(capi.DsdtmError):
        ctx.align2d(1, levels, patches, start, 10)


def test_warp_affine_bit_exact(ctx, scenario):
    ctx.upload(0, scenario["ref_img"])
    packed, offs, ws, hs = scenario["ref_pyr"]
    rng = np.random.default_rng(3)
    n = 300
    A = np.tile(np.eye(2), (n, 1, 1)) + rng.uniform(-0.3, 0.3, (n, 2, 2))
    A[:20] *= rng.uniform(1.5, 3.0, (20, 1, 1))
    rl = rng.integers(0, 4, n).astype(np.int32)
    sl = rng.integers(0, 3, n).astype(np.int32); sl[: n // 2] = 0
    rpx = np.stack([rng.uniform(1, 639, n), rng.uniform(1, 479, n)], 1).astype(np.float32)
    rpx[:6] = [[0.2, 0.3], [639, 479], [638.9, 100], [320, 478.99], [3, 3], [639.0, 0.0]]
    got = ctx.warp_affine(np.zeros(n, np.int32), A, rpx, rl, sl)
    for i in range(n):
        want = O.warp_affine(A[i], O.pyr_level(packed, offs, ws, hs, int(rl[i])), rpx[i], int(rl[i]), int(sl[i]))
        assert (want == got[i]).all(), i
    # Q3: search level >= 1 -> constant patches
    assert all(len(set(got[i])) == 1 for i in range(n) if sl[i] >= 1)


def test_fused_candidate_pipeline_matches_oracle_stage_by_stage(ctx, scenario):
    """SolveAffineMatrix + GetBestSearchLevel + WarpAffine + Align2D as one device pipeline
    (dsdtm_feature_align_batch) against the oracle's per-candidate chain (ref: src/Feature_alignment.cpp:142-156)."""
    from dsdtm_b200 import capi
    sc = scenario
    _upload(ctx, sc)                     # slot 0 = reference keyframe, slot 1 = current frame
    packed, offs, ws, hs = sc["ref_pyr"]
    oc = H.ocam(sc["cam"])
    F = sc["feats"]
    n = len(F)
    rng = np.random.default_rng(8)
    # current pose = ground truth, plus a few exaggerated motions so that det(A) > 3 exercises search levels >= 1 (Q3 path)
    poses = np.tile(sc["T_c2r"], (n, 1))
    for j in range(0, n, 9):
        poses[j] = S.pose_from_xi([0.0, 0.0, -1.05 - 0.2 * rng.uniform(), 0, 0, 0])       # move towards the scene: magnification > sqrt(3)
    kf_center = sc["ref_center"]
    cands = np.zeros(n, capi.CANDIDATE_DT)
    fx, fy, cx, cy = (float(np.float32(sc["cam"][k])) for k in ("fx", "fy", "cx", "cy"))
    for i in range(n):
        q = O.se3_act(poses[i], F[i]["point_w"])
        cands[i]["ref_slot"] = 0; cands[i]["ref_level"] = F[i]["level"]; cands[i]["ref_px"] = F[i]["px"]
        cands[i]["ref_normal"] = F[i]["normal"]; cands[i]["ref_point_w"] = F[i]["point_w"]; cands[i]["kf_center"] = kf_center
        cands[i]["pose_c2r"] = poses[i]
        cands[i]["px"] = (fx * q[0] / q[2] + cx + rng.uniform(-1, 1), fy * q[1] / q[2] + cy + rng.uniform(-1, 1))
    px, lv, conv, A = ctx.feature_align_batch(1, cands, 2, 10, want_A=True)
    levels_seen = set()
    for i in range(n):
        Ao = O.solve_affine(oc, kf_center, F[i]["point_w"], F[i]["normal"], F[i]["px"], int(F[i]["level"]), poses[i])
        assert np.allclose(A[i], Ao, rtol=1e-13, atol=1e-15), i
        Lo = O.best_search_level(Ao, 2)
        assert lv[i] == Lo
        levels_seen.add(Lo)
        patch = O.warp_affine(Ao, O.pyr_level(packed, offs, ws, hs, int(F[i]["level"])), F[i]["px"], int(F[i]["level"]), Lo)
        p, c, _ = O.align2d(O.pyr_level(sc["cur_pyr"][0], offs, ws, hs, Lo), patch, 10, cands[i]["px"] / (1 << Lo))
        assert c == conv[i], i
        assert np.allclose(p * (1 << Lo), px[i], atol=PX_TOL * (1 << Lo), equal_nan=True), i
    assert len(levels_seen) >= 2 and conv.sum() > 200
    # Q3: no candidate with search level >= 1 can converge (constant patch -> singular H)
    assert not conv[lv >= 1].any()


def test_local_map_stage_matches_oracle(ctx, scenario):
    """ReprojectPoint + Get_ClosetObs + the IsInImage gate + FindMatchDirect's arithmetic as one device call
    (dsdtm_local_map_align_batch) against the oracle's per-point chain (ref: src/Feature_alignment.cpp:54-69,128-158,
    src/MapPoint.cpp:133-174). Integer outputs (flags, cell, chosen observation, level) exact; px within the north_star tolerance."""
    from dsdtm_b200 import capi
    sc = scenario
    _upload(ctx, sc)                     # slot 0 = keyframe image, slot 1 = current frame
    packed, offs, ws, hs = sc["ref_pyr"]
    cam = sc["cam"]; oc = H.ocam(cam)
    rng = np.random.default_rng(17)
    T_cur = sc["T_cur"]
    cur_center = O.se3_inv(T_cur)[4:]
    # keyframe table: the real reference frame, the same image seen from a shifted pose, and a keyframe far to the side
    kf_pose = [sc["T_ref"], S.pose_mul(S.pose_from_xi([0.25, -0.1, 0.05, 0.0, 0.03, 0.0]), sc["T_ref"]),
               S.pose_mul(S.pose_from_xi([3.0, 0.5, 1.5, 0.0, -0.9, 0.0]), sc["T_ref"]), S.pose_mul(S.pose_from_xi([-0.1, 0.05, 0.0, 0.01, 0.0, 0.02]), sc["T_ref"])]
    kfs = np.zeros(len(kf_pose), capi.KF_VIEW_DT)
    for k, T in enumerate(kf_pose):
        kfs[k]["slot"] = 0 if k != 3 else 1
        kfs[k]["pose_c2w"] = T; kfs[k]["center"] = O.se3_inv(T)[4:]
    F = sc["feats"]
    pts_w = [F[i]["point_w"] for i in range(len(F))]
    pts_w += [np.array([rng.uniform(-4, 4), rng.uniform(-3, 3), rng.uniform(-1.0, 6.0)]) for _ in range(200)]   # some behind / outside
    pts_w += [np.array([0.3, 0.2, 0.0])]                                                                       # z = 0 in the ref camera
    obs, pts = [], np.zeros(len(pts_w), capi.MAP_POINT_DT)
    for i, P in enumerate(pts_w):
        pts[i]["point_w"] = P; pts[i]["obs_begin"] = len(obs)
        cnt = 0 if i % 37 == 36 else int(rng.integers(1, 5))
        for k in rng.permutation(len(kf_pose))[:cnt]:
            ob = np.zeros((), capi.OBS_DT)
            f = F[i % len(F)]
            ob["kf"] = k; ob["level"] = f["level"] if k == 0 else rng.integers(0, 4)
            ob["px"] = f["px"] if rng.uniform() < 0.8 else (rng.uniform(0, 640), rng.uniform(0, 480))       # some fail the :138 gate
            ob["normal"] = O.feature_normal(oc, ob["px"]); ob["point_w"] = P
            obs.append(ob)
        pts[i]["obs_count"] = cnt
    obs = np.array(obs, capi.OBS_DT)
    out = ctx.local_map_align_batch(1, T_cur, cur_center, kfs, obs, pts, 2, 10)
    rows, cols = O.grid_dims(cam["width"], cam["height"], 15)
    n_aligned = n_conv = 0
    seen = set()
    mism = []
    for i, P in enumerate(pts_w):
        r = out[i]
        inimg, px, cell = O.reproject_point(oc, T_cur, P, 15, cols)
        assert np.array_equal(px, r["px_proj"], equal_nan=True), (i, px, r["px_proj"])
        assert cell == r["cell"] and bool(r["flags"] & capi.LM_IN_IMAGE) == inimg, i
        b, c = int(pts[i]["obs_begin"]), int(pts[i]["obs_count"])
        ok, best = O.closest_obs(cur_center, P, [kfs[int(obs[b + j]["kf"])]["center"] for j in range(c)])
        assert r["obs"] == (b + best if c else -1) and bool(r["flags"] & capi.LM_OBS_OK) == ok, i
        ref_ok = False
        if c:
            ob = obs[b + best]
            L0 = int(ob["level"])
            rp = ob["px"] / np.float32(1 << L0)
            ref_ok = O.is_in_image(oc, rp[0], rp[1], 5, L0)
        assert bool(r["flags"] & capi.LM_REF_OK) == ref_ok, i
        seen.add(int(r["flags"]) & 7)
        if not (inimg and ok and ref_ok):
            assert r["level"] == -1 and not (r["flags"] & capi.LM_CONVERGED) and np.array_equal(r["px"], r["px_proj"], equal_nan=True), i
            continue
        n_aligned += 1
        kf = kfs[int(ob["kf"])]
        T_c2r = O.se3_mul(T_cur, O.se3_inv(kf["pose_c2w"]))
        A = O.solve_affine(oc, kf["center"], ob["point_w"], ob["normal"], ob["px"], L0, T_c2r)
        SL = O.best_search_level(A, 2)
        assert r["level"] == SL, i
        img_pyr = sc["ref_pyr"][0] if kf["slot"] == 0 else sc["cur_pyr"][0]
        patch = O.warp_affine(A, O.pyr_level(img_pyr, offs, ws, hs, L0), ob["px"], L0, SL)
        p, conv, _ = O.align2d(O.pyr_level(sc["cur_pyr"][0], offs, ws, hs, SL), patch, 10, px / (1 << SL))
        assert conv == bool(r["flags"] & capi.LM_CONVERGED), i
        if not np.allclose(p * (1 << SL), r["px"], atol=PX_TOL * (1 << SL), equal_nan=True):
            # a non-converged Gauss-Newton run on a deliberately wrong patch wanders for 10 iterations and amplifies the
            # reduction-order difference of the fp32 sums; only those may differ
            wandered = (not conv) and np.abs(p * (1 << SL) - px).max() > 2.0
            mism.append((i, conv, wandered, SL, int(ob["kf"]), p * (1 << SL), r["px"].copy(), px))
        n_conv += conv
    assert all(m[2] for m in mism) and len(mism) <= n_aligned // 20, (len(mism), n_aligned, mism[:12])
    assert n_aligned > 200 and n_conv > 100 and len(seen) >= 5, (n_aligned, n_conv, seen)
    # bad tables are refused
    bad = obs.copy(); bad[0]["kf"] = len(kfs)
    with pytest.raises(capi.DsdtmError):
        ctx.local_map_align_batch(1, T_cur, cur_center, kfs, bad, pts, 2, 10)
    badp = pts.copy(); badp[-1]["obs_begin"] = len(obs); badp[-1]["obs_count"] = 1
    with pytest.raises(capi.DsdtmError):
        ctx.local_map_align_batch(1, T_cur, cur_center, kfs, obs, badp, 2, 10)


def test_staged_batch_run_graph_replay_and_e2e(ctx):
    """dsdtm_batch_stage/run/fetch (CUDA-graph replay on HBM-resident inputs) and the host-buffer e2e call give the same
    results as the single-pair entry points."""
    from dsdtm_b200 import capi
    scs = [H.make_scenario(s) for s in (31, 32)]
    n, stride, ppp = 2, 320, 40
    for i, sc in enumerate(scs):
        ctx.upload(2 * i, sc["ref_img"]); ctx.upload(2 * i + 1, sc["cur_img"])
    feats = np.zeros((n, stride), O.REF_FEAT_DT); nf = np.zeros(n, np.int32)
    lv = np.zeros((n, ppp), np.int32); pt = np.zeros((n, ppp, 100), np.uint8); st = np.zeros((n, ppp, 2))
    for i, sc in enumerate(scs):
        nf[i] = len(sc["feats"]); feats[i, :nf[i]] = sc["feats"]
        lv[i], pt[i], _, st[i] = H.make_patches(sc["cur_pyr"], ppp, 50 + i, max_level=1)
    lv[1, 5] = -1
    centers = np.stack([sc["ref_center"] for sc in scs]); poses = np.tile(S.IDENTITY, (n, 1))
    ref_slots = np.array([0, 2], np.int32); cur_slots = np.array([1, 3], np.int32)
    ctx.batch_stage(ref_slots, cur_slots, feats, nf, centers, poses, 5, 0, 8, pt, st, lv, 10)
    for rep in range(3):                              # graph replay: same answer every time
        ctx.batch_run(1)
        poses_b, nt_b, px_b, conv_b = ctx.batch_fetch()
        if rep:
            assert (poses_b == last[0]).all() and (px_b == last[2]).all()
        last = (poses_b, nt_b, px_b, conv_b)
    assert ctx.last_run_ms() > 0
    for i, sc in enumerate(scs):
        ps, ns, _ = ctx.sparse_align(2 * i, 2 * i + 1, sc["feats"], sc["ref_center"], S.IDENTITY, 5, 0, 8)
        assert (ps == poses_b[i]).all() and ns == nt_b[i]
        pxs, cs = ctx.align2d(2 * i + 1, lv[i], pt[i], st[i], 10)
        assert (pxs == px_b[i]).all() and (cs == conv_b[i].astype(bool)).all()
    # e2e with host buffers: cur images uploaded inside the call
    cur_imgs = np.stack([sc["cur_img"] for sc in scs])
    out = dict(poses=np.empty((n, 7)), n_tracked=np.empty(n, np.int32), px=np.empty((n, ppp, 2)), conv=np.empty((n, ppp), np.uint8))
    ctx.pair_batch_e2e(cur_imgs, ref_slots, cur_slots, feats, stride, nf, centers, poses, 5, 0, 8, pt, st, lv, ppp, 10, out)
    assert (out["poses"] == poses_b).all() and (out["n_tracked"] == nt_b).all()
    assert (out["px"] == px_b).all() and (out["conv"] == conv_b).all()
    with pytest.raises(capi.DsdtmError):
        ctx.sparse_align_batch(np.arange(20), np.arange(20), np.zeros((20, 4), O.REF_FEAT_DT), np.zeros(20, np.int32),
                               np.zeros((20, 3)), np.tile(S.IDENTITY, (20, 1)), 5, 0, 8)     # > max_batch


def test_large_calls_take_the_direct_copy_path_and_agree_with_the_packed_arena_path(built, scenario):
    """The per-frame entry points pack small calls into a 1 MB pinned arena (one copy each way); calls that do not fit use the
    caller's buffers directly. Both paths must give identical results: 9000 patches / 8000 map points against the same data in
    small calls."""
    from dsdtm_b200 import capi
    sc = scenario
    c = capi.Context(sc["cam"], levels=5, cell_size=15, max_feats=320, max_patches=320, max_frames=4, max_batch=32)
    try:
        c.upload(0, sc["ref_img"]); c.upload(1, sc["cur_img"])
        levels, patches, truth, start = H.make_patches(sc["cur_pyr"], 300, 21, max_level=2)
        px_s, conv_s = c.align2d(1, levels, patches, start, 10)
        reps = 30
        px_l, conv_l = c.align2d(1, np.tile(levels, reps), np.tile(patches, (reps, 1)), np.tile(start, (reps, 1)), 10)
        assert (px_l.reshape(reps, 300, 2) == px_s).all() and (conv_l.reshape(reps, 300) == conv_s).all()

        F = sc["feats"]
        T_cur = sc["T_cur"]; cen = O.se3_inv(T_cur)[4:]
        kfs = np.zeros(1, capi.KF_VIEW_DT); kfs[0]["slot"] = 0; kfs[0]["pose_c2w"] = sc["T_ref"]; kfs[0]["center"] = O.se3_inv(sc["T_ref"])[4:]
        n = len(F)
        obs = np.zeros(n, capi.OBS_DT); pts = np.zeros(n, capi.MAP_POINT_DT)
        for i in range(n):
            obs[i]["kf"] = 0; obs[i]["level"] = F[i]["level"]; obs[i]["px"] = F[i]["px"]; obs[i]["normal"] = F[i]["normal"]; obs[i]["point_w"] = F[i]["point_w"]
            pts[i]["point_w"] = F[i]["point_w"]; pts[i]["obs_begin"] = i; pts[i]["obs_count"] = 1
        small = c.local_map_align_batch(1, T_cur, cen, kfs, obs, pts, 2, 10)
        reps = 8000 // n + 1
        obs_l = np.tile(obs, reps); pts_l = np.tile(pts, reps)
        pts_l["obs_begin"] = np.arange(len(pts_l))
        large = c.local_map_align_batch(1, T_cur, cen, kfs, obs_l, pts_l, 2, 10)
        for k in ("px_proj", "px", "cell", "flags", "level"):
            L = large[k].reshape((reps, n) + large[k].shape[1:])
            for r in range(reps):
                assert np.array_equal(L[r], small[k], equal_nan=(small[k].dtype.kind == "f")), (k, r)
        assert (large["obs"] == np.arange(len(pts_l))).all()
        assert (small["flags"] & capi.LM_CONVERGED).astype(bool).sum() > 200
    finally:
        c.close()


def test_track_frame_single_call_equals_the_three_separate_calls(ctx, scenario):
    """dsdtm_track_frame = new Frame + Sprase_ImgAlign::Run + UpdateLocalMap/SearchLocalPoints' device part in one call with one
    synchronisation (ref: src/Tracking.cpp:57,199-224); the pose found by the sparse alignment is composed with the reference pose
    on the device and feeds the local-map kernels without visiting the host. Must equal the separate calls bit for bit."""
    from dsdtm_b200 import capi
    sc = scenario
    ctx.upload(0, sc["ref_img"])
    F = sc["feats"]; n = len(F)
    T_ref = sc["T_ref"]
    pose_in = O.se3_mul(T_ref, O.se3_inv(T_ref))                       # cur.Set_Pose(last.Get_Pose()) -> T_c2r = identity
    kfs = np.zeros(1, capi.KF_VIEW_DT); kfs[0]["slot"] = 0; kfs[0]["pose_c2w"] = T_ref; kfs[0]["center"] = sc["ref_center"]
    obs = np.zeros(n, capi.OBS_DT); pts = np.zeros(n, capi.MAP_POINT_DT)
    for i in range(n):
        obs[i]["kf"] = 0; obs[i]["level"] = F[i]["level"]; obs[i]["px"] = F[i]["px"]; obs[i]["normal"] = F[i]["normal"]; obs[i]["point_w"] = F[i]["point_w"]
        pts[i]["point_w"] = F[i]["point_w"]; pts[i]["obs_begin"] = i; pts[i]["obs_count"] = 1
    # --- separate calls (slot 1), pose composition on the host with the oracle's Sophus arithmetic
    ctx.upload(1, sc["cur_img"])
    p_out, ntr, _ = ctx.sparse_align(0, 1, F, sc["ref_center"], pose_in, 4, 0, 30)
    T_cur = O.se3_mul(p_out, T_ref); cen = O.se3_inv(T_cur)[4:]
    rep_sep = ctx.local_map_align_batch(1, T_cur, cen, kfs, obs, pts, 2, 10)
    # --- one call (slot 2 gets the frame)
    out, rep = ctx.track_frame(0, 2, sc["cur_img"], F, sc["ref_center"], T_ref, pose_in, (4, 0, 30), kfs, obs, pts, 2, 10)
    assert (out["pose_c2r"] == p_out).all() and out["n_tracked"] == ntr
    assert (out["pose_cur_c2w"] == T_cur).all() and (out["cur_center"] == cen).all()      # device composition == Sophus restatement
    for k in rep.dtype.names:
        assert np.array_equal(rep[k], rep_sep[k], equal_nan=(rep[k].dtype.kind == "f")), k
    packed, offs, ws, hs = sc["cur_pyr"]
    for l in range(5):
        assert (ctx.download_level(2, l) == O.pyr_level(packed, offs, ws, hs, l)).all()
    assert (rep["flags"] & capi.LM_CONVERGED).astype(bool).sum() > 200
    d = S.pose_dist(out["pose_c2r"], sc["T_c2r"])
    assert d[0] < 1e-3 and d[1] < 3e-3
    # no local map: alignment only
    out2, rep2 = ctx.track_frame(0, 2, sc["cur_img"], F, sc["ref_center"], T_ref, pose_in, (4, 0, 30), kfs[:0], obs[:0], pts[:0], 2, 10)
    assert (out2["pose_c2r"] == p_out).all() and len(rep2) == 0


def test_track_frame_refuses_bad_arguments(ctx, scenario):
    from dsdtm_b200 import capi
    sc = scenario
    ctx.upload(0, sc["ref_img"])
    F = sc["feats"]
    kfs = np.zeros(1, capi.KF_VIEW_DT); kfs[0]["pose_c2w"] = sc["T_ref"]; kfs[0]["center"] = sc["ref_center"]
    obs = np.zeros(2, capi.OBS_DT); obs["point_w"] = F[:2]["point_w"]; obs["px"] = F[:2]["px"]; obs["normal"] = F[:2]["normal"]
    pts = np.zeros(2, capi.MAP_POINT_DT); pts["point_w"] = F[:2]["point_w"]; pts["obs_begin"] = [0, 1]; pts["obs_count"] = 1
    ok = lambda **kw: ctx.track_frame(kw.get("ref", 0), kw.get("cur", 2), sc["cur_img"], kw.get("feats", F), sc["ref_center"], sc["T_ref"], S.IDENTITY,
                                      kw.get("cfg", (4, 0, 30)), kw.get("kfs", kfs), kw.get("obs", obs), kw.get("pts", pts), kw.get("msl", 2), 10)
    ok()                                                                  # the well-formed call works
    bad_obs = obs.copy(); bad_obs[1]["kf"] = 3
    bad_pts = pts.copy(); bad_pts[1]["obs_begin"] = 2
    bad_kfs = kfs.copy(); bad_kfs[0]["slot"] = 99
    for kw in (dict(cur=99), dict(ref=-1), dict(cfg=(9, 0, 30)), dict(cfg=(4, 4, 30)), dict(msl=7), dict(obs=bad_obs), dict(pts=bad_pts), dict(kfs=bad_kfs)):
        with pytest.raises(capi.DsdtmError):
            ok(**kw)
    ok()                                                                  # and the context is still usable afterwards


def test_large_batch_uses_the_throughput_kernel_and_matches_the_oracle(built):
    """A batch of >= 4 pairs per SM switches sparse_align to its 3-warps-per-pair instantiation (four CTAs per SM). 640 pairs =
    replicas of two scenes with per-replica start poses: every pair against the oracle (on the distinct inputs), replicas with equal
    inputs bit-equal, staged graph replay equal to the direct batch call."""
    from dsdtm_b200 import capi
    cam = dict(S.KINECT)
    n, stride = 640, 320
    ctx = capi.Context(cam, levels=5, cell_size=15, max_feats=stride, max_patches=8, max_frames=2 * n, max_batch=n)
    try:
        scs = [H.make_scenario(200 + i, cam) for i in range(2)]
        rng = np.random.default_rng(5)
        starts = [S.IDENTITY] + [S.pose_from_xi(rng.uniform(-0.003, 0.003, 6)) for _ in range(3)]
        imgs = np.stack([scs[(i // 2) % 2]["ref_img" if i % 2 == 0 else "cur_img"] for i in range(2 * n)])
        for c0 in range(0, 2 * n, 256):
            ctx.upload_batch(c0, imgs[c0:c0 + 256])
        feats = np.zeros((n, stride), O.REF_FEAT_DT); nf = np.zeros(n, np.int32); centers = np.zeros((n, 3)); poses = np.zeros((n, 7))
        for i in range(n):
            sc = scs[i % 2]
            nf[i] = len(sc["feats"]); feats[i, :nf[i]] = sc["feats"]; centers[i] = sc["ref_center"]; poses[i] = starts[(i // 2) % 4]
        ref_slots = 2 * np.arange(n); cur_slots = ref_slots + 1
        pb, tb, _, _ = ctx.sparse_align_batch(ref_slots, cur_slots, feats, nf, centers, poses, 4, 0, 30)
        for k in range(8):                                        # the 8 distinct (scene, start pose) combinations
            sc = scs[k % 2]
            packed, offs, ws, hs = sc["ref_pyr"]
            po, no, _ = O.sparse_align(H.ocam(cam), packed, sc["cur_pyr"][0], offs, ws, hs, sc["feats"], sc["ref_center"], starts[(k // 2) % 4], 4, 0, 30)
            same = np.arange(k, n, 8)
            d = S.pose_dist(po, pb[k])
            assert d[0] < POSE_TOL and d[1] < POSE_TOL and no == tb[k], (k, d)
            assert (pb[same] == pb[k]).all() and (tb[same] == tb[k]).all(), k
        ctx.batch_stage(ref_slots, cur_slots, feats, nf, centers, poses, 4, 0, 30)
        ctx.batch_run(0); ctx.batch_run(0)
        pg, tg, _, _ = ctx.batch_fetch()
        assert (pg == pb).all() and (tg == tb).all()
    finally:
        ctx.close()


def test_chained_batch_is_safe_under_chunking_and_in_the_e2e_pipeline(built):
    """in a chained batch (ref[i] == cur[i-1]) a pair's reference slot is another pair's current slot. With
    step_chunks > 1 all pyramids are built before the fork, and dsdtm_pair_batch_e2e runs such a batch as one chunk: results must
    equal the unchunked step bit for bit (before the fix another chunk's stream could still be writing levels >= 1 / uploading
    level 0 of a slot while this chunk's alignment read it)."""
    from dsdtm_b200 import capi
    cam = dict(S.KINECT)
    scene = S.Scene(5)
    poses3 = [S.IDENTITY, S.pose_from_xi([0.01, 0.004, -0.003, 0.002, -0.003, 0.001]), S.pose_from_xi([0.018, -0.006, 0.004, -0.003, 0.004, 0.002])]
    imgs, feats3, centers3 = [], [], []
    for p in poses3:
        img, _, pts = S.render(scene, cam, p, want_points=True)
        corners, _ = H.detect_oracle(img, 5, 15, 60)
        imgs.append(img); feats3.append(H.ref_feats_from_corners(cam, corners, pts)); centers3.append(O.se3_inv(p)[4:])
    n, stride = 640, 64                                              # >= 4 pairs per SM: the chunked path is taken
    c = capi.Context(cam, levels=5, cell_size=15, max_feats=stride, max_patches=1, max_frames=n + 1, max_batch=n)
    try:
        c.upload_batch(0, np.stack([imgs[i % 3] for i in range(n + 1)]))
        ref_slots = np.arange(n, dtype=np.int32); cur_slots = ref_slots + 1
        feats = np.zeros((n, stride), O.REF_FEAT_DT); nf = np.zeros(n, np.int32); centers = np.zeros((n, 3)); poses = np.zeros((n, 7))
        for i in range(n):
            f = feats3[i % 3]
            nf[i] = len(f); feats[i, :nf[i]] = f; centers[i] = centers3[i % 3]
            poses[i] = O.se3_mul(poses3[i % 3], O.se3_inv(poses3[i % 3]))          # start at the reference pose: T_c2r = identity
        pt = np.zeros((n, 0, 100), np.uint8); st = np.zeros((n, 0, 2)); lv = np.zeros((n, 0), np.int32)
        c.set_option("sa_warps_per_pair", 4)                          # singles and batches reduce in the same order
        singles = [c.sparse_align(int(ref_slots[i]), int(cur_slots[i]), feats[i, :nf[i]], centers[i], poses[i], 5, 0, 8) for i in (0, 1, 2, 317, 639)]
        c.batch_stage(ref_slots, cur_slots, feats, nf, centers, poses, 5, 0, 8, None, None, None, 10)
        c.set_option("step_chunks", 1)
        c.batch_run(1); base = c.batch_fetch()
        for (p, k, _), i in zip(singles, (0, 1, 2, 317, 639)):
            assert (p == base[0][i]).all() and k == base[1][i]
        c.set_option("step_chunks", 4)
        for rep in range(3):
            c.upload_batch(0, np.stack([imgs[i % 3] for i in range(n + 1)]))          # level 0 again; run() rebuilds levels >= 1
            c.batch_run(1); got = c.batch_fetch()
            assert (got[0] == base[0]).all() and (got[1] == base[1]).all()
        c.set_option("step_chunks", 1)
        cur_imgs = np.ascontiguousarray(np.stack([imgs[(i + 1) % 3] for i in range(n)]))
        out = dict(poses=np.empty((n, 7)), n_tracked=np.empty(n, np.int32))
        for rep in range(2):
            c.pair_batch_e2e(cur_imgs, ref_slots, cur_slots, feats, stride, nf, centers, poses, 5, 0, 8, None, None, None, 0, 10, out)
            assert (out["poses"] == base[0]).all() and (out["n_tracked"] == base[1]).all()
    finally:
        c.close()


def test_batched_refinement_chain_matches_oracle(built):
    """dsdtm_batch_run(flags | 2) / dsdtm_track_batch_e2e: pyramid -> Run -> pose composition -> ReprojectPoint / Get_ClosetObs / gates ->
    SolveAffineMatrix -> WarpAffine -> Align2D for every reference feature of every pair, against the same chain restated with the
    oracle (orc_pair_batch_map): flags, cells, search levels exact, refined pixels within 1e-3 px, poses within 1e-5."""
    from dsdtm_b200 import capi
    cam = dict(S.KINECT)
    oc = H.ocam(cam)
    T_ref = S.pose_from_xi(np.r_[0.2, -0.1, 0.05, 0.03, -0.02, 0.04])
    prs = [S.make_pair(61 + k, cam, trans=0.03, rot_deg=0.8, ref_pose=(T_ref if k % 2 else None)) for k in range(3)]
    n, stride, ppp = 6, 320, 300
    c = capi.Context(cam, levels=5, cell_size=15, max_feats=stride, max_patches=ppp, max_frames=2 * n, max_batch=n)
    try:
        feats = np.zeros((n, stride), O.REF_FEAT_DT); nf = np.zeros(n, np.int32); centers = np.zeros((n, 3))
        poses_ref = np.zeros((n, 7)); poses_in = np.tile(S.IDENTITY, (n, 1))
        ref_pyrs = []; cur_imgs = []
        for i in range(n):
            pr = prs[i % 3]
            c.upload(2 * i, pr["ref_img"]); c.upload(2 * i + 1, pr["cur_img"])
            corners, pyr = H.detect_oracle(pr["ref_img"], 5, 15, 300)
            F = H.ref_feats_from_corners(cam, corners, pr["ref_points"])
            if i >= 3:
                F["initial"][::11] = 0                                   # features without a map point are no candidates
                F["px"][3] = (2.0, 2.0)                                  # a reference feature too close to the border: REF_OK fails
            nf[i] = len(F); feats[i, :nf[i]] = F; centers[i] = O.se3_inv(pr["T_ref"])[4:]; poses_ref[i] = pr["T_ref"]
            ref_pyrs.append(pyr[0]); cur_imgs.append(pr["cur_img"])
        ref_slots = 2 * np.arange(n, dtype=np.int32); cur_slots = ref_slots + 1
        c.batch_stage(ref_slots, cur_slots, feats, nf, centers, poses_in, 5, 0, 8, None, None, None, 10)
        c.batch_stage_map(poses_ref, ppp, 2, 10)
        c.batch_run(3)
        poses_g, nt_g, _, _ = c.batch_fetch()
        rep_g = c.batch_fetch_map()
        poses_o, nt_o, rep_o = O.pair_batch_map(oc, 5, 15, np.stack(ref_pyrs), np.stack(cur_imgs), feats, nf, centers, poses_ref, poses_in,
                                                5, 0, 8, ppp, 2, 10, 4)
        for i in range(n):
            d = S.pose_dist(poses_o[i], poses_g[i])
            assert d[0] < 1e-5 and d[1] < 1e-5 and nt_g[i] == nt_o[i]
        assert (rep_g["flags"] == rep_o["flags"]).all() and (rep_g["cell"] == rep_o["cell"]).all()
        assert (rep_g["level"] == rep_o["level"]).all() and (rep_g["obs"] == rep_o["obs"]).all()
        assert np.abs(rep_g["px_proj"] - rep_o["px_proj"]).max() < 1e-6        # the aligned poses differ by ~1e-15; projections follow
        ok = (rep_g["flags"] & capi.LM_CONVERGED) != 0
        assert ok.sum() > 0.8 * nf.sum() * 0.9 and np.abs(rep_g["px"][ok] - rep_o["px"][ok]).max() <= 1e-3
        assert (rep_g["level"][:, :][rep_g["flags"] & 7 != 7] == -1).all()
        # replay + the host-buffer entry give the same bits
        c.batch_run(3)
        assert c.batch_fetch_map().tobytes() == rep_g.tobytes()           # byte-wise: Q3 candidates carry NaN positions
        out = dict(poses=np.empty((n, 7)), n_tracked=np.empty(n, np.int32), reproj=np.zeros((n, ppp), capi.REPROJ_DT))
        c.track_batch_e2e(np.ascontiguousarray(np.stack(cur_imgs)), ref_slots, cur_slots, feats, stride, nf, centers, poses_ref, poses_in, 5, 0, 8, ppp, 2, 10, out)
        assert (out["poses"] == poses_g).all() and (out["n_tracked"] == nt_g).all() and out["reproj"].tobytes() == rep_g.tobytes()
        with pytest.raises(capi.DsdtmError):
            c.batch_stage(ref_slots, cur_slots, feats, nf, centers, poses_in, 5, 0, 8, None, None, None, 10)
            c.batch_run(2)                                             # chain requested without dsdtm_batch_stage_map
    finally:
        c.close()
