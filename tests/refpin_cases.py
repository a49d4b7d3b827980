"""The seeded inputs behind tests/golden/refpin.npz, shared by the generator (tests/golden/make_golden_refpin.py, which runs the
reference's own compiled code on them) and the consumers (tests/test_ref_golden.py: oracle bit-equal on CPU, CUDA path within the
stated tolerances on the GPU box)."""
import hashlib

import numpy as np

import helpers as H
import oracle as O
from dsdtm_b200 import synth as S

SA_SEEDS = [20260101, 4, 57]
SA_CFGS = [(4, 0, 30), (5, 0, 8), (5, 2, 8), (3, 1, 4)]
SA_EUROC_SEED = 100
A2D_ITERS = (3, 10)
DET_CASES = [("kinect", S.KINECT, 20260101), ("kinect8", S.KINECT, 8), ("euroc", S.EUROC, 100)]


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def a2d_case():
    sc = H.make_scenario(6)
    levels, patches, truth, start = H.make_patches(sc["cur_pyr"], 300, 3, max_level=2)
    return sc, levels, patches, truth, start


def warp_cases(n=200):
    rng = np.random.default_rng(8)
    sc = H.make_scenario(6)
    packed, offs, ws, hs = sc["ref_pyr"]
    out = []
    for t in range(n):
        L0 = int(rng.integers(0, 3))
        img = O.pyr_level(packed, offs, ws, hs, L0)
        h, w = img.shape
        A = np.eye(2) * rng.uniform(0.6, 3.2) + rng.normal(0, 0.2, (2, 2))
        px = np.array([rng.uniform(0, w - 1), rng.uniform(0, h - 1)], np.float32) * np.float32(1 << L0)
        if t % 9 == 0:
            px = np.array([rng.uniform(0, 3) * (1 << L0), rng.uniform(0, h - 1) * (1 << L0)], np.float32)
        out.append((img, A, px, L0, int(t % 3 == 2)))      # every third case at search level 1 (Q3: constant patch)
    return out


def shi_case():
    rng = np.random.default_rng(5)
    img = S.make_pair(11)["ref_img"]
    pts = np.c_[rng.integers(0, 640, 400), rng.integers(0, 480, 400)]
    pts = np.r_[pts, [[4, 4], [5, 5], [635, 475], [634, 474], [0, 0], [639, 479], [5, 100], [100, 5]]]
    return img, pts
