"""GPU parity for the keyframe-ingest stage: depth convertTo, UndistortFeatures, Get_FeatureDetph,
UnProject (ref: src/Tracking.cpp:56,412-464; src/Frame.cpp:94-157,200-224) against the oracle and the cv2 golden."""
import numpy as np
import pytest

import helpers as H
import oracle as O
from dsdtm_b200 import synth as S

pytestmark = pytest.mark.gpu

EUROC_DIST = (-0.28368365, 0.07451284, -0.00010473, -3.55590700e-05, 0.0)      # ref: Config/EuRoc.yaml:16-20
DEFAULT_DIST = (0.231222, -0.784899, -0.003257, -0.000105, 0.917205)            # ref: Config/default.yaml:39-43


def _depth_image(rng, h, w):
    d = rng.integers(400, 60000, (h, w)).astype(np.uint16)
    d[rng.uniform(size=(h, w)) < 0.35] = 0                    # holes: exercises the 4-neighbourhood fallback and the -1 case
    d[100:140, 200:260] = 0
    return d


def test_depth_convert_bit_exact(ctx):
    rng = np.random.default_rng(1)
    d = [_depth_image(rng, ctx.height, ctx.width) for _ in range(3)]
    for k in range(3):
        ctx.depth_upload(k, d[k])
    for scale in (5000.0, 1000.0):
        got = ctx.depth_convert_f32(0, 3, scale)
        for k in range(3):
            assert (got[k] == O.depth_convert(d[k], scale)).all()
    one = ctx.depth_convert_f32(2, 1, 5000.0)
    assert (one[0] == O.depth_convert(d[2], 5000.0)).all()
    from dsdtm_b200 import capi
    with pytest.raises(capi.DsdtmError):
        ctx.depth_convert_f32(3, 2, 5000.0)                  # past the depth pool (default 4 slots)
    with pytest.raises(capi.DsdtmError):
        ctx.depth_upload(4, d[0])


@pytest.mark.parametrize("dist", [EUROC_DIST, DEFAULT_DIST, (0.0, 0.0, 0.0, 0.0, 0.0)])
def test_keyframe_lift_matches_oracle(ctx, dist):
    """Undistorted pixels bit-equal (they feed cvRound), normals / world points to 1e-12, depth and status exact."""
    from dsdtm_b200 import capi
    cam = S.KINECT; oc = H.ocam(cam)
    rng = np.random.default_rng(7)
    h, w = ctx.height, ctx.width
    d16 = _depth_image(rng, h, w)
    ctx.depth_upload(1, d16)
    scale = 5000.0
    df = O.depth_convert(d16, scale)
    n = 700
    px = np.stack([rng.integers(3, w - 3, n), rng.integers(3, h - 3, n)], 1).astype(np.float32)   # detector output: integers
    px[-100:] += rng.uniform(-0.5, 0.5, (100, 2)).astype(np.float32)                             # refined (sub-pixel) features
    px[:4] = [[0, 0], [w - 1, h - 1], [3, h - 4], [w - 4, 3]]                                       # corners: lookups leave the image
    initial = (rng.uniform(size=n) < 0.3).astype(np.uint8)
    pose = S.pose_from_xi([0.3, -0.2, 0.1, 0.05, -0.1, 0.2])
    out = ctx.keyframe_lift(1, pose, dist, scale, px, initial)
    und = O.undistort_points(oc, dist, px)
    seen = set()
    for i in range(n):
        o = out[i]
        seen.add(int(o["status"]))
        if initial[i]:
            assert o["status"] == capi.LIFT_SKIPPED and (o["px"] == px[i]).all()
            continue
        assert (o["px"].view(np.uint32) == und[i].view(np.uint32)).all(), (i, o["px"], und[i])
        assert np.allclose(o["normal"], O.feature_normal(oc, und[i]), rtol=0, atol=1e-15)
        z = O.feature_depth(df, und[i])
        assert o["depth"] == np.float32(z), (i, o["depth"], z)
        if z < 0:
            assert o["status"] == capi.LIFT_NO_DEPTH
        else:
            assert o["status"] == capi.LIFT_OK
            assert np.allclose(o["point_w"], O.unproject(oc, pose, und[i], z), rtol=0, atol=1e-12), i
    assert seen == {capi.LIFT_SKIPPED, capi.LIFT_OK, capi.LIFT_NO_DEPTH}
    # no depth slot: undistort + normal only
    out2 = ctx.keyframe_lift(-1, pose, dist, scale, px[:50])
    assert (out2["status"] == capi.LIFT_NO_DEPTH).all() and (out2["px"].view(np.uint32) == und[:50].view(np.uint32)).all()


def test_keyframe_lift_reproduces_cv2_golden(golden, built):
    """The device undistortion against cv2.undistortPoints itself (752x480 EuRoC geometry from the golden file)."""
    from dsdtm_b200 import capi
    g = golden["undistort_cv2"]
    for name in ("euroc", "strong"):
        K, D, src, dst = g[name + "_K"], g[name + "_D"], g[name + "_src"], g[name + "_dst"]
        w, hgt = (int(v) for v in g[name + "_wh"])
        cam = dict(width=w, height=hgt, fx=float(K[0, 0]), fy=float(K[1, 1]), cx=float(K[0, 2]), cy=float(K[1, 2]), f=float(K[0, 0]))
        c = capi.Context(cam, levels=5, cell_size=15, max_feats=320, max_patches=8, max_frames=2, max_batch=1)
        out = c.keyframe_lift(-1, S.IDENTITY, D, 1.0, src)
        assert (out["px"].view(np.uint32) == dst.view(np.uint32)).all(), name
        c.close()


def test_clahe_upload_is_bit_exact_and_feeds_the_pyramid(golden, built):
    """cv::createCLAHE(clip, tiles)->apply + ComputeImagePyramid as one device call, against the cv2 4.13 golden and the oracle,
    single and batched (ref: Test/test_Feature_detection.cpp:85-89)."""
    from dsdtm_b200 import capi
    g = golden["clahe_cv2"]
    for name, h, w, clip, tiles in H.CLAHE_CASES:
        img = H.clahe_input(name, h, w)
        cam = dict(width=w, height=h, fx=400.0, fy=400.0, cx=w / 2.0, cy=h / 2.0, f=400.0)
        levels = 3 if min(h, w) >= 96 else 2
        c = capi.Context(cam, levels=levels, cell_size=15, max_feats=64, max_patches=8, max_frames=4, max_batch=1)
        try:
            imgs = np.stack([img, img[::-1].copy(), np.roll(img, 5, axis=1)])
            out = c.upload_clahe(1, imgs, clip, tiles)
            assert (out[0] == g[name]).all(), name                                  # cv2 itself
            for k in range(3):
                want = O.clahe(imgs[k], clip, tiles)
                assert (out[k] == want).all(), (name, k)
                packed, offs, ws, hs = O.pyramid(want, levels)
                for l in range(levels):
                    assert (c.download_level(1 + k, l) == O.pyr_level(packed, offs, ws, hs, l)).all(), (name, k, l)
            with pytest.raises(capi.DsdtmError):
                c.upload_clahe(0, img, clip, (7, 8) if w % 7 else (9, 8))         # not divisible: refused, not padded
        finally:
            c.close()
