"""Synthetic bench / sweep workload: batches of independent frame pairs, built without the oracle.

K distinct scenes are ray-cast on the host (dsdtm_b200.synth, a multiprocessing pool), reference features come from the GPU
FAST stage + the host-side grid selection, and the K scenes are tiled to B pairs with per-pair start poses so that no two
pairs run the same Gauss-Newton trace. Every pair owns its own frame slots, features and patches in HBM.
"""
import multiprocessing as mp
import os

import numpy as np

from . import synth as S
from .capi import CORNER_DT, REF_FEAT_DT

BASE_SEED = 20260101


def _render(args):
    seed, cam, trans, rot = args
    pr = S.make_pair(seed, cam, trans, rot)
    # world points at integer pixels are only needed at feature locations; keep the full map (float64 HxWx3) out of the pipe
    return dict(seed=seed, T_c2r=pr["T_c2r"], T_ref=pr["T_ref"], ref_img=pr["ref_img"], cur_img=pr["cur_img"],
                ref_points=pr["ref_points"].astype(np.float64))


def render_scenes(k, cam, trans=0.02, rot_deg=0.5, seed0=BASE_SEED, procs=None):
    procs = procs or min(k, max(1, (os.cpu_count() or 2) // 2), 16)
    args = [(seed0 + i, cam, trans, rot_deg) for i in range(k)]
    if procs <= 1 or k == 1:
        return [_render(a) for a in args]
    with mp.get_context("spawn").Pool(procs) as pool:
        return pool.map(_render, args)


def circle_mask_offsets(r):
    """Row spans of OpenCV's filled midpoint circle of radius r (cv::circle(..., -1)), as {dy: half_width}."""
    spans = {}
    err, dx, dy, plus, minus = 0, r, 0, 1, (r << 1) - 1
    while dx >= dy:
        for yy, hw in ((dy, dx), (-dy, dx), (dx, dy), (-dx, dy)):
            spans[yy] = max(spans.get(yy, -1), hw)
        dy += 1
        err += plus
        plus += 2
        mask = -1 if err > 0 else 0
        err -= minus & mask
        dx += mask
        minus -= mask & 2
    return spans


def select_features(cells, width, height, cell_size, max_fts):
    """Host-side tail of Feature_detector::detect (ref: src/Feature_detection.cpp:111-150): score-descending order, literal
    `> 20` filter, mask test and filled-circle painting. (The C++ adapter in dsdtm_b200/host is the drop-in version and keeps
    std::sort's unstable tie order; numpy's stable sort here only matters for exactly equal float scores.)"""
    order = np.argsort(-cells["score"], kind="stable")
    mask = np.full((height, width), 255, np.uint8)
    spans = circle_mask_offsets(cell_size)
    out = []
    for i in order:
        c = cells[i]
        if c["score"] > 20 and mask[c["y"], c["x"]] == 255:
            out.append(i)
            for dy, hw in spans.items():
                y = c["y"] + dy
                if 0 <= y < height:
                    mask[y, max(c["x"] - hw, 0):min(c["x"] + hw, width - 1) + 1] = 0
        if len(out) >= max_fts:
            break
    return cells[np.array(out, np.int64)] if out else cells[:0]


def feature_normals(cam, px):
    """Frame::Add_Feature's bearing (ref: src/Frame.cpp:83-92; float Pixel2Camera then normalize, src/Camera.cpp:173-178)."""
    fx, fy, cx, cy = (np.float32(cam[k]) for k in ("fx", "fy", "cx", "cy"))
    px = px.astype(np.float32)
    n = np.stack([((px[:, 0] - cx) / fx).astype(np.float64), ((px[:, 1] - cy) / fy).astype(np.float64), np.ones(len(px))], 1)
    return n / np.linalg.norm(n, axis=1, keepdims=True)


def scene_inputs(sc, cam, cells, cell_size, n_feats, feat_stride, patches_per_pair):
    """Per-scene inputs of a pair from the per-cell corner records of its reference frame (`cells`: the FAST / Shi-Tomasi stage's
    output, from the GPU -- dsdtm_fast_cells -- or from any implementation with the same contract): selected features with their
    bearings and map points, and the host-patch variant of the refinement inputs."""
    w, h = cam["width"], cam["height"]
    feats_c = select_features(cells, w, h, cell_size, n_feats)
    nf = len(feats_c)
    F = np.zeros(feat_stride, REF_FEAT_DT)
    px = np.stack([feats_c["x"], feats_c["y"]], 1).astype(np.float32)
    F["px"][:nf] = px
    F["level"][:nf] = feats_c["level"]
    F["initial"][:nf] = 1
    F["normal"][:nf] = feature_normals(cam, px)
    F["point_w"][:nf] = sc["ref_points"][feats_c["y"], feats_c["x"]]
    # feature-alignment inputs: 10x10 reference patches around the ref features (level 0, identity warp) and start
    # positions = true reprojection into cur + U(+-1) px
    P = S.pose_act(S.pose_mul(sc["T_c2r"], sc["T_ref"]), F["point_w"][:nf])
    fxd, fyd, cxd, cyd = (float(np.float32(cam[q])) for q in ("fx", "fy", "cx", "cy"))
    proj = np.stack([fxd * P[:, 0] / P[:, 2] + cxd, fyd * P[:, 1] / P[:, 2] + cyd], 1)
    rng = np.random.default_rng(sc["seed"] + 777)
    npatch = min(patches_per_pair, nf)
    patches = np.zeros((patches_per_pair, 100), np.uint8)
    ppx = np.zeros((patches_per_pair, 2))
    plv = np.full(patches_per_pair, -1, np.int32)
    for j in range(npatch):
        x, y = int(feats_c["x"][j]), int(feats_c["y"][j])
        if 6 <= x < w - 6 and 6 <= y < h - 6 and 8 <= proj[j, 0] < w - 8 and 8 <= proj[j, 1] < h - 8:
            patches[j] = sc["ref_img"][y - 5:y + 5, x - 5:x + 5].reshape(-1)
            ppx[j] = proj[j] + rng.uniform(-1, 1, 2)
            plv[j] = 0
    return dict(F=F, nf=nf, patches=patches, ppx=ppx, plv=plv, truth_px=proj)


def assemble_batch(scenes, per_scene, n_pairs, feat_stride, patches_per_pair, seed0=BASE_SEED, first_slot=0, lo=0, hi=None):
    """Host arrays of pairs [lo, hi) of a GLOBAL batch of n_pairs pairs (pair g uses scene g % k; start poses come from one seeded
    stream over the global index, so a shard of the batch is the same data whichever rank builds it). Slots are local to the shard."""
    k = len(scenes)
    hi = n_pairs if hi is None else hi
    B = hi - lo
    feats = np.zeros((B, feat_stride), REF_FEAT_DT)
    nfe = np.zeros(B, np.int32)
    centers = np.zeros((B, 3))
    poses = np.zeros((B, 7))
    poses_ref = np.zeros((B, 7))
    patches = np.zeros((B, patches_per_pair, 100), np.uint8)
    ppx = np.zeros((B, patches_per_pair, 2))
    plv = np.zeros((B, patches_per_pair), np.int32)
    rng = np.random.default_rng(seed0 + 4242)
    start = np.concatenate([rng.uniform(-0.004, 0.004, (n_pairs, 3)), rng.uniform(-0.002, 0.002, (n_pairs, 3))], 1)
    for i in range(B):
        g = lo + i
        s = g % k
        ps = per_scene[s]
        feats[i] = ps["F"]; nfe[i] = ps["nf"]
        T_ref = scenes[s]["T_ref"]
        centers[i] = -(S.quat_to_R(T_ref[:4]).T @ T_ref[4:])
        poses_ref[i] = T_ref
        # start pose: identity (cur.Set_Pose(last.Get_Pose()), ref: src/Tracking.cpp:201) perturbed per replica
        poses[i] = S.IDENTITY if g < k else S.pose_from_xi(start[g])
        patches[i] = ps["patches"]; ppx[i] = ps["ppx"]; plv[i] = ps["plv"]
    ref_slots = first_slot + 2 * np.arange(B, dtype=np.int32)
    cur_slots = ref_slots + 1
    truth = np.stack([scenes[(lo + i) % k]["T_c2r"] for i in range(B)]) if B else np.zeros((0, 7))
    return dict(n_pairs=B, n_scenes=k, ref_slots=ref_slots, cur_slots=cur_slots, feats=feats, n_feats=nfe, centers=centers,
                poses_in=poses, poses_ref=poses_ref, patches=patches, patch_px=ppx, patch_level=plv, truth=truth, scenes=scenes,
                per_scene=per_scene, feat_stride=feat_stride, patches_per_pair=patches_per_pair, lo=lo)


def upload_frames(ctx, batch, first_slot=0):
    """ref / cur frames of the shard into slots first_slot + 2i / + 2i + 1 (level 0 + pyramid), in chunks of 32 pairs"""
    scenes = batch["scenes"]; k = len(scenes); B = batch["n_pairs"]; lo = batch["lo"]
    h, w = scenes[0]["ref_img"].shape
    chunk = 32
    for i0 in range(0, B, chunk):
        n = min(chunk, B - i0)
        buf = np.empty((2 * n, h, w), np.uint8)
        for j in range(n):
            s = (lo + i0 + j) % k
            buf[2 * j] = scenes[s]["ref_img"]; buf[2 * j + 1] = scenes[s]["cur_img"]
        ctx.upload_batch(first_slot + 2 * i0, buf)


def build_batch(ctx, cam, n_pairs, n_scenes=8, n_feats=300, feat_stride=320, patches_per_pair=300, seed0=BASE_SEED, scenes=None,
                first_slot=0, lo=0, hi=None):
    """Uploads the frames of pairs [lo, hi) of a global batch of n_pairs pairs (default: all of it; ref slots first_slot + 2i, cur
    slots first_slot + 2i + 1) and returns the host-side batch dict. Reference features come from the GPU FAST stage."""
    scenes = scenes or render_scenes(n_scenes, cam, seed0=seed0)
    per_scene = []
    for sc in scenes:
        ctx.upload(first_slot, sc["ref_img"])
        cells = ctx.fast_cells(first_slot, 20, 5.0)
        per_scene.append(scene_inputs(sc, cam, cells, ctx.prm.cell_size, n_feats, feat_stride, patches_per_pair))
    batch = assemble_batch(scenes, per_scene, n_pairs, feat_stride, patches_per_pair, seed0, first_slot, lo, hi)
    upload_frames(ctx, batch, first_slot)
    return batch


# ---------------------------------------------------------------- pose refinement after matching
def _quat_from_rotvec(w):
    th = float(np.linalg.norm(w))
    if th < 1e-12:
        return np.array([1.0, 0, 0, 0])
    return np.r_[np.cos(th / 2), np.sin(th / 2) * np.asarray(w) / th]


def _qmul(a, b):
    w1, x1, y1, z1 = a; w2, x2, y2, z2 = b
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2, w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2])


def make_ba_problem(seed, n=200, noise=1e-3, outliers=0.1, max_level=0, start_rot=0.01, start_trans=0.02):
    """A frame after SearchLocalPoints for Optimizer::PoseOptimization: n matched map points seen by a camera at a
    seeded pose, observations = unit bearing vectors (Feature::mNormal) of the true projections + pixel noise / focal, a share of
    gross outliers, pyramid levels U{0..max_level}, start pose = truth perturbed. Returns dict(normals, levels, points_w, pose_in, truth)."""
    r = np.random.default_rng(seed)
    q = _quat_from_rotvec(r.uniform(-0.2, 0.2, 3)); t = r.uniform(-0.3, 0.3, 3)
    Pc = np.c_[r.uniform(-1.5, 1.5, n), r.uniform(-1, 1, n), r.uniform(1.5, 4, n)]
    R = S.quat_to_R(q)
    Pw = (Pc - t) @ R                                   # cam = R Pw + t
    obs = np.c_[Pc[:, 0] / Pc[:, 2], Pc[:, 1] / Pc[:, 2]] + r.normal(0, noise, (n, 2))
    k = r.random(n) < outliers
    obs[k] += r.normal(0, 0.05, (int(k.sum()), 2))
    nm = np.c_[obs, np.ones(n)]
    nm /= np.linalg.norm(nm, axis=1)[:, None]
    lv = r.integers(0, max_level + 1, n).astype(np.int32)
    q0 = _qmul(_quat_from_rotvec(r.normal(0, start_rot, 3)), q)
    pose0 = np.r_[q0, t + r.normal(0, start_trans, 3)]
    return dict(normals=np.ascontiguousarray(nm), levels=lv, points_w=np.ascontiguousarray(Pw), pose_in=pose0, truth=np.r_[q, t])
