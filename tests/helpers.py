"""Shared test helpers: synthetic scenarios checked by the CPU oracle (test infrastructure)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import oracle as O  # noqa: E402
from dsdtm_b200 import synth as S  # noqa: E402
from dsdtm_b200 import workload as W  # noqa: E402


def ocam(cam):
    return O.make_cam(cam["width"], cam["height"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["f"])


def detect_oracle(img, levels=5, cell=15, max_fts=300, thr=5.0):
    """Feature_detector::detect on the oracle: returns (features CORNER_DT, packed pyramid, offs, ws, hs)."""
    packed, offs, ws, hs = O.pyramid(img, levels)
    cells = O.detect_cells(packed, offs, ws, hs, cell, None, thr)
    mask = np.full(img.shape, 255, np.uint8)
    feats, _ = O.detect_select(cells, mask, cell, max_fts)
    return feats, (packed, offs, ws, hs)


def ref_feats_from_corners(cam, corners, ref_points, n_pad=None):
    """dsdtm_ref_feat records for detected corners with analytic depth (Feature + MapPoint snapshot)."""
    oc = ocam(cam)
    n = len(corners)
    F = np.zeros(n if n_pad is None else n_pad, O.REF_FEAT_DT)
    for i, c in enumerate(corners):
        F[i]["px"] = (c["x"], c["y"])
        F[i]["level"] = c["level"]
        F[i]["initial"] = 1
        F[i]["normal"] = O.feature_normal(oc, F[i]["px"])
        F[i]["point_w"] = ref_points[c["y"], c["x"]]
    return F


def make_scenario(seed, cam=S.KINECT, levels=5, max_fts=300, trans=0.02, rot_deg=0.5):
    pr = S.make_pair(seed, cam, trans, rot_deg)
    corners, pyr = detect_oracle(pr["ref_img"], levels, 15, max_fts)
    pr["corners"] = corners
    pr["feats"] = ref_feats_from_corners(cam, corners, pr["ref_points"])
    pr["ref_pyr"] = pyr
    pr["cur_pyr"] = O.pyramid(pr["cur_img"], levels)
    pr["ref_center"] = -(S.quat_to_R(pr["T_ref"][:4]).T @ pr["T_ref"][4:])
    return pr


def make_patches(cur_pyr, n, seed, max_level=0, pert=1.5, margin=8):
    """test_Feature_alignment recipe (ref: Test/test_Feature_alignment.cpp:47-86) on a synthetic image: 10x10 reference
    patches interpolated at sub-pixel truths, start positions perturbed by U(+-pert)."""
    packed, offs, ws, hs = cur_pyr
    rng = np.random.default_rng(seed)
    levels = rng.integers(0, max_level + 1, n).astype(np.int32)
    patches = np.zeros((n, 100), np.uint8)
    truth = np.zeros((n, 2))
    start = np.zeros((n, 2))
    for i in range(n):
        L = int(levels[i])
        img = O.pyr_level(packed, offs, ws, hs, L).astype(np.float64)
        h, w = img.shape
        tx = rng.uniform(margin + 6, w - margin - 7)
        ty = rng.uniform(margin + 6, h - margin - 7)
        truth[i] = (tx, ty)
        # generateRefPatchNoWarpInterpolate: bilinear sample of a 10x10 window centred on the truth
        xs = tx - 5 + np.arange(10)
        ys = ty - 5 + np.arange(10)
        x0 = np.floor(xs).astype(int); y0 = np.floor(ys).astype(int)
        fx = xs - x0; fy = ys - y0
        I = img
        p = ((1 - fy)[:, None] * ((1 - fx)[None, :] * I[np.ix_(y0, x0)] + fx[None, :] * I[np.ix_(y0, x0 + 1)]) +
             fy[:, None] * ((1 - fx)[None, :] * I[np.ix_(y0 + 1, x0)] + fx[None, :] * I[np.ix_(y0 + 1, x0 + 1)]))
        patches[i] = np.clip(np.rint(p), 0, 255).astype(np.uint8).reshape(-1)
        start[i] = truth[i] + rng.uniform(-pert, pert, 2)
    return levels, patches, truth, start


CLAHE_CASES = [  # (name, h, w, clip limit, tiles)
    ("tex320", 240, 320, 3.0, (8, 8)), ("rand160", 120, 160, 3.0, (8, 8)), ("blocks96", 64, 96, 2.0, (4, 4)),
    ("ramp752", 480, 752, 3.0, (8, 8)), ("flat128", 96, 128, 40.0, (8, 8)), ("tiny", 64, 96, 0.5, (8, 8)),
]


def clahe_input(name, h, w):
    """Deterministic pure-numpy test images for the CLAHE goldens (no cv2 at test time)."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    if name.startswith("tex"):
        v = 120 + 60 * np.sin(x * 0.11) * np.cos(y * 0.07) + 30 * np.sin((x + 2 * y) * 0.31) + 15 * np.cos(x * y * 0.001)
    elif name.startswith("rand"):
        v = np.random.default_rng(160).integers(0, 256, (h, w)).astype(np.float64)
    elif name.startswith("blocks"):
        v = np.zeros((h, w)); v[:, : w // 2] = 17; v[h // 3:, w // 3:] = 200; v[::7, ::5] = 255
    elif name.startswith("ramp"):
        v = 90 + 40 * x / w + 10 * np.sin(y * 0.2) + ((x.astype(int) * 7 + y.astype(int) * 13) % 5)
    elif name.startswith("flat"):
        v = np.full((h, w), 128.0); v[10:20, 10:40] = 131
    else:
        v = (x * 3 + y * 5) % 256
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


make_ba_problem = W.make_ba_problem   # the generator lives with the other synthetic workloads (bench.py uses it too)


def ba_obs_records(pr, dtype, n_pad=None):
    """dsdtm_ba_obs records of a make_ba_problem() frame."""
    n = len(pr["levels"])
    o = np.zeros(n if n_pad is None else n_pad, dtype)
    o["normal"][:n] = pr["normals"]; o["point_w"][:n] = pr["points_w"]; o["level"][:n] = pr["levels"]
    if n_pad is not None and n_pad > n:
        o["normal"][n:] = (0, 0, 1); o["point_w"][n:] = (0, 0, 1)
    return o


BA_CASES = [  # (seed, n, noise, outliers, max_level, start_rot, start_trans)
    (0, 200, 1e-3, 0.1, 0, 0.01, 0.02), (1, 300, 1e-3, 0.1, 3, 0.01, 0.02), (2, 37, 1e-2, 0.4, 0, 0.01, 0.02),
    (3, 5, 1e-3, 0.0, 2, 0.01, 0.02), (4, 1, 1e-3, 0.0, 0, 0.01, 0.02), (5, 1000, 1e-2, 0.4, 3, 0.01, 0.02),
    (6, 200, 1e-3, 0.1, 0, 0.05, 0.5), (7, 300, 1e-3, 0.3, 4, 0.1, 0.3), (8, 64, 0.0, 0.0, 0, 0.02, 0.05),
    (2, 33, 1e-3, 0.1, 1, 0.3, 1.0),     # far start: 20 iterations, 11 accepted
    (0, 33, 1e-3, 0.4, 0, 0.5, 2.0),     # runs into the 100-iteration cap
    (19, 33, 1e-3, 0.4, 3, 0.8, 0.5),    # ends on the parameter tolerance
    (3, 33, 1e-3, 0.4, 3, 0.8, 0.5),     # 65 iterations, 58 accepted
]


def make_map_table(seed, n_kfs=60, pts_per_kf=(40, 300), cam=S.KINECT, spread=6.0):
    """A map for Tracking::GetCloseKeyFrames (caller side): key frames scattered over `spread` metres around the
    origin, each with its own cloud of map points ~2 m in front of it (some null / zero rows, some behind the camera), and a current
    pose near the origin. Returns dict(kfs MAP_KF-like arrays, points, pose_cur)."""
    r = np.random.default_rng(seed)
    pt_begin, pt_count, kf_t, pts = [], [], [], []
    total = 0
    for k in range(n_kfs):
        n = int(r.integers(pts_per_kf[0], pts_per_kf[1] + 1)) if k % 11 else 0          # a few key frames without any point
        q = _rotvec_quat(r.normal(0, 0.15, 3))
        centre = np.r_[r.uniform(-spread, spread, 2), r.uniform(-0.5, 0.5)]
        R = S.quat_to_R(q)
        t = -R @ centre                                                                  # pose c2w: p_cam = R p_w + t
        local = np.c_[r.uniform(-1.5, 1.5, n), r.uniform(-1.1, 1.1, n), r.uniform(1.0, 3.5, n)]
        world = (local - t) @ R
        if n:
            world[r.random(n) < 0.08] = 0.0                                              # null map points / isZero(0)
        pt_begin.append(total); pt_count.append(n); kf_t.append(t); pts.append(world); total += n
    pose_cur = np.r_[_rotvec_quat(r.normal(0, 0.1, 3)), r.normal(0, 0.3, 3)]
    return dict(pt_begin=np.array(pt_begin, np.int32), pt_count=np.array(pt_count, np.int32), kf_t=np.array(kf_t),
                points=np.concatenate(pts) if total else np.zeros((0, 3)), pose_cur=pose_cur)


def _rotvec_quat(w):
    th = float(np.linalg.norm(w))
    return np.array([1.0, 0, 0, 0]) if th < 1e-12 else np.r_[np.cos(th / 2), np.sin(th / 2) * np.asarray(w) / th]
