"""Oracle-side model of the tracking loop (test infrastructure): the pieces of Frame / KeyFrame / MapPoint / Map the hot path reads,
and `Tracking`'s front-end calls restated with oracle primitives. Shared by the GPU sequence tests (CUDA path vs this model) and by
tests/test_ref_pin.py (this model vs the reference's own classes compiled into oracle/_ref/libdsdtm_ref.so)."""
import numpy as np

import oracle as O
from dsdtm_b200 import synth as S

LEVELS, CELL = 5, 15


def _trajectory(n, seed=3, scale=1.0):
    rng = np.random.default_rng(seed)
    v = np.concatenate([rng.uniform(-0.012, 0.012, 3), np.deg2rad(rng.uniform(-0.3, 0.3, 3))])
    poses = [S.IDENTITY.copy()]
    for k in range(1, n):
        step = scale * v * (1.0 + 0.2 * np.sin(0.7 * k))
        poses.append(S.pose_mul(S.pose_from_xi(step), poses[-1]))
    return poses


class OracleFrame:
    def __init__(self, img, pose):
        self.img = img
        self.pyr = O.pyramid(img, LEVELS)
        self.pose = np.array(pose)
        self.feats = np.zeros(0, O.REF_FEAT_DT)       # features with map points (px, level, normal, point_w)


def _oracle_sparse_align(oc, cur, ref, cfg):
    packed, offs, ws, hs = ref.pyr
    center = O.se3_inv(ref.pose)[4:]
    T0 = O.se3_mul(cur.pose, O.se3_inv(ref.pose))
    po, n, log = O.sparse_align(oc, packed, cur.pyr[0], offs, ws, hs, ref.feats, center, T0, *cfg)
    return O.se3_mul(po, ref.pose), n


def _oracle_search(oc, cam, cur, kf, found):
    """SearchLocalPoints against a one-keyframe map, restated with oracle primitives (ref: src/Feature_alignment.cpp:54-158)."""
    w, h = cam["width"], cam["height"]
    gcols = -(-w // CELL)
    fx, fy, cx, cy = (float(np.float32(cam[k])) for k in ("fx", "fy", "cx", "cy"))
    cells = {}
    for i, f in enumerate(kf.feats):
        q = O.se3_act(cur.pose, f["point_w"])
        px = np.array([fx * q[0] / q[2] + cx, fy * q[1] / q[2] + cy])
        rx, ry = O.cvround(np.float32(px[0])), O.cvround(np.float32(px[1]))
        if 8 <= rx < w - 8 and 8 <= ry < h - 8:
            cells.setdefault(int(px[1] / CELL) * gcols + int(px[0] / CELL), []).append((i, px))
    mask = np.full((h, w), 255, np.uint8)
    packed, offs, ws, hs = kf.pyr
    kf_center = O.se3_inv(kf.pose)[4:]
    cur_center = O.se3_inv(cur.pose)[4:]
    T_c2r = O.se3_mul(cur.pose, O.se3_inv(kf.pose))
    out = []
    for k in sorted(cells):
        for i, px in sorted(cells[k], key=lambda c: -found[c[0]]):
            if mask[O.cvround(np.float32(px[1])), O.cvround(np.float32(px[0]))] != 255:
                continue
            f = kf.feats[i]
            a = kf_center - f["point_w"]; b = cur_center - f["point_w"]
            if np.dot(a / np.linalg.norm(a), b / np.linalg.norm(b)) < 0.5:
                continue
            L0 = int(f["level"])
            rpx = f["px"] / np.float32(1 << L0)
            if not (5 <= O.cvround(rpx[0]) < w // (1 << L0) - 5 and 5 <= O.cvround(rpx[1]) < h // (1 << L0) - 5):
                continue
            A = O.solve_affine(oc, kf_center, f["point_w"], f["normal"], f["px"], L0, T_c2r)
            SL = O.best_search_level(A, LEVELS - 3)
            patch = O.warp_affine(A, O.pyr_level(packed, offs, ws, hs, L0), f["px"], L0, SL)
            p, conv, _ = O.align2d(O.pyr_level(cur.pyr[0], offs, ws, hs, SL), patch, 10, px / (1 << SL))
            p = p * (1 << SL)
            if not conv:
                continue
            out.append((i, np.float32(p), SL))
            O.circle_fill(mask, O.cvround(np.float32(p[0])), O.cvround(np.float32(p[1])), CELL, 0)
            break
        if len(out) >= 200:
            break
    return out


class OMap:
    """Oracle-side model of the pieces of Map / MapPoint / KeyFrame the hot path reads."""
    def __init__(self):
        self.mp_point = []      # id -> world point
        self.mp_found = []      # id -> found count (MapPoint::mnFound starts at 1)
        self.mp_obs = []        # id -> list of (kf index, feature index)
        self.kfs = []           # OracleFrame + .feat_mp (map point id per feature)

    def new_mp(self, p):
        self.mp_point.append(np.array(p)); self.mp_found.append(1); self.mp_obs.append([])
        return len(self.mp_point) - 1


def _oracle_search_multi(oc, cam, cur, omap, kf_order):
    """UpdateLocalMap + SearchLocalPoints over several keyframes (ref: src/Tracking.cpp:276-305, src/Feature_alignment.cpp:54-158,
    src/MapPoint.cpp:133-174)."""
    w, h = cam["width"], cam["height"]
    gcols = -(-w // CELL)
    fx, fy, cx, cy = (float(np.float32(cam[k])) for k in ("fx", "fy", "cx", "cy"))
    cells, seen = {}, set()
    for q in kf_order:
        for mp in omap.kfs[q].feat_mp:
            if mp < 0 or mp in seen:
                continue
            seen.add(mp)
            qq = O.se3_act(cur.pose, omap.mp_point[mp])
            px = np.array([fx * qq[0] / qq[2] + cx, fy * qq[1] / qq[2] + cy])
            rx, ry = O.cvround(np.float32(px[0])), O.cvround(np.float32(px[1]))
            if 8 <= rx < w - 8 and 8 <= ry < h - 8:
                cells.setdefault(int(px[1] / CELL) * gcols + int(px[0] / CELL), []).append((mp, px))
    mask = np.full((h, w), 255, np.uint8)
    cur_center = O.se3_inv(cur.pose)[4:]
    out = []
    for k in sorted(cells):
        for mp, px in sorted(cells[k], key=lambda c: -omap.mp_found[c[0]]):
            if mask[O.cvround(np.float32(px[1])), O.cvround(np.float32(px[0]))] != 255:
                continue
            P = omap.mp_point[mp]
            b = cur_center - P; b = b / np.linalg.norm(b)
            best, best_obs = 0.0, omap.mp_obs[mp][0]
            for (qk, fi) in omap.mp_obs[mp]:
                a = O.se3_inv(omap.kfs[qk].pose)[4:] - P
                c = float(np.dot(a / np.linalg.norm(a), b))
                if c > best:
                    best, best_obs = c, (qk, fi)
            if best < 0.5:
                continue
            kf = omap.kfs[best_obs[0]]
            f = kf.feats[best_obs[1]]
            L0 = int(f["level"])
            rpx = f["px"] / np.float32(1 << L0)
            if not (5 <= O.cvround(rpx[0]) < w // (1 << L0) - 5 and 5 <= O.cvround(rpx[1]) < h // (1 << L0) - 5):
                continue
            packed, offs, ws, hs = kf.pyr
            kf_center = O.se3_inv(kf.pose)[4:]
            T_c2r = O.se3_mul(cur.pose, O.se3_inv(kf.pose))
            A = O.solve_affine(oc, kf_center, f["point_w"], f["normal"], f["px"], L0, T_c2r)
            SL = O.best_search_level(A, LEVELS - 3)
            patch = O.warp_affine(A, O.pyr_level(packed, offs, ws, hs, L0), f["px"], L0, SL)
            p, conv, _ = O.align2d(O.pyr_level(cur.pyr[0], offs, ws, hs, SL), patch, 10, px / (1 << SL))
            p = p * (1 << SL)
            if not conv:
                continue
            out.append((mp, np.float32(p), SL, best_obs[0]))
            O.circle_fill(mask, O.cvround(np.float32(p[0])), O.cvround(np.float32(p[1])), CELL, 0)
            break
        if len(out) >= 200:
            break
    return out


