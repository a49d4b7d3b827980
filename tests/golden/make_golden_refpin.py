"""Generates tests/golden/refpin.npz: outputs of the REFERENCE'S OWN hot-path code (oracle/_ref/libdsdtm_ref.so = the unmodified
/root/reference translation units compiled against tests/ref_shim by `make -C oracle ref_dsdtm`) on seeded synthetic inputs, so
that the pin established in this container (tests/test_ref_pin.py) travels to machines without /root/reference:

    PYTHONPATH=.:tests python tests/golden/make_golden_refpin.py

The inputs are NOT stored: they are regenerated from the seeds by dsdtm_b200/synth.py + tests/helpers.py (pure numpy); a digest of
every regenerated input is stored and checked by the consumers (tests/test_ref_golden.py), so a drift of the generators is reported
as such and not as a parity failure.
  sa_*      : Sprase_ImgAlign::Run final poses / tracked counts for seeds x constructor configs (kinect geometry), and one EuRoC pair
  a2d_*     : Feature_Alignment::Align2DGaussNewton positions / flags for 300 patches x {3, 10} iterations
  warp_*    : Feature_Alignment::WarpAffine 10x10 patches + GetBestSearchLevel for 200 random affine maps
  shi_*     : Feature_detector::shiTomasiScore at 400 positions
  det_*     : Feature_detector::detect corner lists
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import helpers as H  # noqa: E402
import oracle as O  # noqa: E402
from dsdtm_b200 import synth as S  # noqa: E402
from oracle import refpin as RP  # noqa: E402
import refpin_cases as K  # noqa: E402


def main():
    RP.build()
    out = {}
    R = RP.Ref(dict(S.KINECT), levels=5)
    # ---- sparse alignment
    poses, counts, digests = [], [], []
    for seed in K.SA_SEEDS:
        sc = H.make_scenario(seed)
        digests.append(K.digest(sc["ref_img"], sc["cur_img"], sc["feats"]))
        for cfg in K.SA_CFGS:
            R.reset()
            ref = R.frame(sc["ref_img"], sc["T_ref"]); cur = R.frame(sc["cur_img"], sc["T_ref"])
            kf = R.keyframe(ref)
            for f in sc["feats"]:
                k = R.add_feature(ref, f["px"], f["level"], True)
                R.set_mappoint(ref, k, R.mappoint(f["point_w"], kf))
            p, n = R.sparse_align_run(cur, ref, *cfg)
            poses.append(p); counts.append(n)
    out["sa_pose"] = np.array(poses).reshape(len(K.SA_SEEDS), len(K.SA_CFGS), 7)
    out["sa_n"] = np.array(counts, np.int32).reshape(len(K.SA_SEEDS), len(K.SA_CFGS))
    out["sa_digest"] = np.array(digests)
    # EuRoC geometry (BASELINE configs[4])
    RE = RP.Ref(dict(S.EUROC), levels=5)
    sc = H.make_scenario(K.SA_EUROC_SEED, dict(S.EUROC))
    ref = RE.frame(sc["ref_img"], sc["T_ref"]); cur = RE.frame(sc["cur_img"], sc["T_ref"])
    kf = RE.keyframe(ref)
    for f in sc["feats"]:
        k = RE.add_feature(ref, f["px"], f["level"], True)
        RE.set_mappoint(ref, k, RE.mappoint(f["point_w"], kf))
    p, n = RE.sparse_align_run(cur, ref, 5, 0, 8)
    out["sa_euroc_pose"] = p; out["sa_euroc_n"] = np.int32(n); out["sa_euroc_digest"] = np.array(K.digest(sc["ref_img"], sc["cur_img"], sc["feats"]))
    # ---- Align2D
    sc, levels, patches, truth, start = K.a2d_case()
    packed, offs, ws, hs = sc["cur_pyr"]
    px = np.zeros((2, len(levels), 2)); conv = np.zeros((2, len(levels)), np.uint8)
    for j, iters in enumerate(K.A2D_ITERS):
        for i in range(len(levels)):
            img = O.pyr_level(packed, offs, ws, hs, int(levels[i]))
            p, c = R.align2d(img, patches[i], O.patch_no_border(patches[i]), iters, start[i])
            px[j, i] = p; conv[j, i] = c
    out["a2d_px"] = px; out["a2d_conv"] = conv; out["a2d_digest"] = np.array(K.digest(sc["cur_img"], patches, start))
    # ---- WarpAffine / GetBestSearchLevel
    cases = K.warp_cases()
    wp = np.zeros((len(cases), 100), np.uint8); wl = np.zeros(len(cases), np.int32)
    for i, (img, A, p, L0, sl) in enumerate(cases):
        wl[i] = R.best_search_level(A, 2)
        wp[i] = R.warp_affine(A, img, p, L0, sl)
    out["warp_patch"] = wp; out["warp_level"] = wl
    out["warp_digest"] = np.array(K.digest(*[np.r_[A.reshape(-1), p, L0, sl] for (_, A, p, L0, sl) in cases]))
    # ---- Shi-Tomasi
    img, pts = K.shi_case()
    out["shi"] = np.array([R.shitomasi(img, u, v) for u, v in pts], np.float32)
    # ---- detect
    for name, cam, seed in K.DET_CASES:
        RR = RP.Ref(dict(cam), levels=5)
        pr = S.make_pair(seed, dict(cam))
        fr = RR.frame(pr["ref_img"], S.IDENTITY)
        RR.detect(fr, 5.0, True)
        f = RR.features(fr)
        out["det_%s_px" % name] = f["px"].astype(np.int16); out["det_%s_level" % name] = f["level"].astype(np.int8)
    np.savez_compressed(os.path.join(HERE, "refpin.npz"), **out)
    print("refpin.npz written:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
