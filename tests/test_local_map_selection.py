"""SURVEY 8f-1, the caller's half: Tracking::GetCloseKeyFrames + the ranking at the top of Tracking::UpdateLocalMap
(ref: src/Tracking.cpp:261-277,315-345; Frame::isVisible ref: src/Frame.cpp:300-311).
CPU: the oracle against a plain numpy restatement (PARITY UNPINNED: the reference holds no golden for it).
GPU: dsdtm_map_table_upload + dsdtm_close_keyframes against the oracle -- flags and ranking exact, distances bit-equal
(non-contracted fp64 in the reference's operation order), incremental table updates, full-size map."""
import numpy as np
import pytest

import helpers as H
import oracle as O
from dsdtm_b200 import synth as S


def _numpy_close(cam, mt, max_local=10):
    R = S.quat_to_R(mt["pose_cur"][:4]); t = mt["pose_cur"][4:]
    fx, fy, cx, cy = (float(np.float32(cam[k])) for k in ("fx", "fy", "cx", "cy"))
    vis = np.zeros(len(mt["pt_begin"]), np.uint8); dist = np.zeros(len(vis))
    for k, (b, n) in enumerate(zip(mt["pt_begin"], mt["pt_count"])):
        P = mt["points"][b:b + n]
        P = P[np.abs(P).sum(1) > 0]
        c = P @ R.T + t
        c = c[c[:, 2] >= 0]
        u = np.float32(fx * c[:, 0] / c[:, 2] + cx); v = np.float32(fy * c[:, 1] / c[:, 2] + cy)
        ok = (np.rint(u) >= 0) & (np.rint(u) < cam["width"]) & (np.rint(v) >= 0) & (np.rint(v) < cam["height"])
        if ok.any():
            vis[k] = 1; dist[k] = np.linalg.norm(t - mt["kf_t"][k])
    order = [k for k in np.argsort(dist, kind="stable") if vis[k]][:max_local]
    return vis, dist, np.array(order, np.int32)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_close_keyframes_against_numpy(seed):
    cam = S.KINECT
    mt = H.make_map_table(seed)
    vis, dist, local = O.close_keyframes(H.ocam(cam), mt["pose_cur"], mt["pt_begin"], mt["pt_count"], mt["kf_t"], mt["points"])
    v2, d2, l2 = _numpy_close(cam, mt)
    assert (vis == v2).all() and np.abs(dist - d2).max() < 1e-12 and (local == l2).all()
    assert 0 < vis.sum() < len(vis) and len(local) == min(10, vis.sum())        # both outcomes occur
    assert (vis[mt["pt_count"] == 0] == 0).all()
    assert (np.diff(dist[local]) >= 0).all()


def _kf_rows(mt):
    from dsdtm_b200 import capi
    rows = np.zeros(len(mt["pt_begin"]), capi.MAP_KF_DT)
    rows["pt_begin"] = mt["pt_begin"]; rows["pt_count"] = mt["pt_count"]; rows["t"] = mt["kf_t"]
    return rows


@pytest.mark.gpu
def test_close_keyframes_matches_oracle_and_updates_incrementally(built):
    from dsdtm_b200 import capi
    cam = dict(S.KINECT)
    ctx = capi.Context(cam, max_frames=2, max_batch=1)
    try:
        mt = H.make_map_table(5, n_kfs=90)
        rows = _kf_rows(mt)
        oc = H.ocam(cam)
        # the map grows key frame by key frame (appends), queried along the way
        first_half = 45
        npts_half = int(mt["pt_begin"][first_half])
        ctx.map_table_upload(0, rows[:first_half], 0, mt["points"][:npts_half])
        v, d, l = ctx.close_keyframes(mt["pose_cur"], first_half)
        vo, do, lo = O.close_keyframes(oc, mt["pose_cur"], mt["pt_begin"][:first_half], mt["pt_count"][:first_half], mt["kf_t"][:first_half], mt["points"])
        assert (v == vo).all() and (d == do).all() and (l == lo).all()
        ctx.map_table_upload(first_half, rows[first_half:], npts_half, mt["points"][npts_half:])
        for s in range(4):                                      # several current poses against the whole table
            pose = H.make_map_table(100 + s, n_kfs=1)["pose_cur"]
            v, d, l = ctx.close_keyframes(pose, len(rows))
            vo, do, lo = O.close_keyframes(oc, pose, mt["pt_begin"], mt["pt_count"], mt["kf_t"], mt["points"])
            assert (v == vo).all() and (d == do).all() and (l == lo).all(), s
            assert 0 < v.sum() < len(v)
        # a bundle adjustment moves points and a key-frame pose: rewrite those rows only
        k = int(lo[0]); b, n = int(mt["pt_begin"][k]), int(mt["pt_count"][k])
        mt["points"][b:b + n] += 100.0                          # out of sight
        mt["kf_t"][k] += 0.25
        rows = _kf_rows(mt)
        ctx.map_table_upload(k, rows[k:k + 1], b, mt["points"][b:b + n])
        v, d, l = ctx.close_keyframes(pose, len(rows), max_local=3)
        vo, do, lo = O.close_keyframes(oc, pose, mt["pt_begin"], mt["pt_count"], mt["kf_t"], mt["points"], 3)
        assert (v == vo).all() and (d == do).all() and (l == lo).all() and v[k] == 0 and len(l) == 3
        # bad arguments
        with pytest.raises(capi.DsdtmError):
            ctx.close_keyframes(pose, len(rows) + 1)
        with pytest.raises(capi.DsdtmError):
            ctx.map_table_upload(len(rows) + 5, rows[:1], 0, mt["points"][:1])
        bad = rows[:1].copy(); bad["pt_count"] = 10 ** 8
        with pytest.raises(capi.DsdtmError):
            ctx.map_table_upload(0, bad, 0, mt["points"][:1])
        v, d, l = ctx.close_keyframes(pose, 0)
        assert len(v) == 0 and len(l) == 0
    finally:
        ctx.close()


@pytest.mark.gpu
def test_close_keyframes_full_size_map(built):
    """A 4096-key-frame map (1 M point rows): device flags and ranking equal the oracle's; the call is one launch."""
    from dsdtm_b200 import capi
    cam = dict(S.KINECT)
    ctx = capi.Context(cam, max_frames=2, max_batch=1)
    try:
        mt = H.make_map_table(9, n_kfs=4096, pts_per_kf=(200, 300), spread=40.0)
        ctx.map_table_upload(0, _kf_rows(mt), 0, mt["points"])
        n0 = ctx.launch_count()
        v, d, l = ctx.close_keyframes(mt["pose_cur"], 4096)
        assert ctx.launch_count() == n0 + 1
        vo, do, lo = O.close_keyframes(H.ocam(cam), mt["pose_cur"], mt["pt_begin"], mt["pt_count"], mt["kf_t"], mt["points"])
        assert (v == vo).all() and (d == do).all() and (l == lo).all() and len(l) == 10
    finally:
        ctx.close()
