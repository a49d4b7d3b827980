"""tests/golden/refpin.npz = outputs of the reference's OWN compiled hot-path code (generator: tests/golden/make_golden_refpin.py;
how that library is built and what it pins: tests/test_ref_pin.py, oracle/ref_dsdtm_wrap.cpp). This file is how the pin reaches
machines without /root/reference:

* CPU (`-m "not gpu"`): the oracle reproduces every golden value BIT FOR BIT (integer decisions, float bits, pose doubles).
* GPU (`-m gpu`): the CUDA path through the C-ABI against the same golden values, within the tolerances BASELINE.json's north_star
  states (final pose 1e-5 rad / 1e-5 m, refined positions 1e-3 px), bit-exact for the integer / byte outputs.
"""
import numpy as np
import pytest

import helpers as H
import oracle as O
import refpin_cases as K
from dsdtm_b200 import synth as S


@pytest.fixture(scope="module")
def G(golden):
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refpin.npz"))


# ------------------------------------------------------------------------------------------------ CPU: oracle == reference
def test_oracle_sparse_align_equals_reference_golden(G):
    for si, seed in enumerate(K.SA_SEEDS):
        sc = H.make_scenario(seed)
        assert K.digest(sc["ref_img"], sc["cur_img"], sc["feats"]) == str(G["sa_digest"][si]), "the synthetic generators drifted: regenerate refpin.npz"
        packed, offs, ws, hs = sc["ref_pyr"]
        for ci, cfg in enumerate(K.SA_CFGS):
            po, no, _ = O.sparse_align(H.ocam(sc["cam"]), packed, sc["cur_pyr"][0], offs, ws, hs, sc["feats"], sc["ref_center"], S.IDENTITY, *cfg)
            assert no == G["sa_n"][si, ci] and (po == G["sa_pose"][si, ci]).all(), (seed, cfg)
    cam = dict(S.EUROC)
    sc = H.make_scenario(K.SA_EUROC_SEED, cam)
    assert K.digest(sc["ref_img"], sc["cur_img"], sc["feats"]) == str(G["sa_euroc_digest"])
    packed, offs, ws, hs = sc["ref_pyr"]
    po, no, _ = O.sparse_align(H.ocam(cam), packed, sc["cur_pyr"][0], offs, ws, hs, sc["feats"], sc["ref_center"], S.IDENTITY, 5, 0, 8)
    assert no == int(G["sa_euroc_n"]) and (po == G["sa_euroc_pose"]).all()


def test_oracle_align2d_warp_shitomasi_detect_equal_reference_golden(G):
    sc, levels, patches, truth, start = K.a2d_case()
    assert K.digest(sc["cur_img"], patches, start) == str(G["a2d_digest"])
    packed, offs, ws, hs = sc["cur_pyr"]
    for j, iters in enumerate(K.A2D_ITERS):
        for i in range(len(levels)):
            p, c, _ = O.align2d(O.pyr_level(packed, offs, ws, hs, int(levels[i])), patches[i], iters, start[i])
            assert c == bool(G["a2d_conv"][j, i]) and (p == G["a2d_px"][j, i]).all()
    assert G["a2d_conv"][1].sum() > 200
    for i, (img, A, p, L0, sl) in enumerate(K.warp_cases()):
        assert O.best_search_level(A, 2) == G["warp_level"][i]
        assert (O.warp_affine(A, img, p, L0, sl) == G["warp_patch"][i]).all()
    img, pts = K.shi_case()
    got = np.array([O.shitomasi(img, u, v) for u, v in pts], np.float32)
    assert (got.view(np.uint32) == G["shi"].view(np.uint32)).all()
    for name, cam, seed in K.DET_CASES:
        pr = S.make_pair(seed, dict(cam))
        oc, _ = H.detect_oracle(pr["ref_img"], 5, 15, 300)
        assert (oc["x"] == G["det_%s_px" % name][:, 0]).all() and (oc["y"] == G["det_%s_px" % name][:, 1]).all()
        assert (oc["level"] == G["det_%s_level" % name]).all()


# ------------------------------------------------------------------------------------------------ GPU: CUDA path vs reference
@pytest.mark.gpu
def test_cuda_sparse_align_against_reference_golden(ctx, G):
    for si, seed in enumerate(K.SA_SEEDS):
        sc = H.make_scenario(seed)
        ctx.upload(0, sc["ref_img"]); ctx.upload(1, sc["cur_img"])
        for ci, cfg in enumerate(K.SA_CFGS):
            pg, ng, _ = ctx.sparse_align(0, 1, sc["feats"], sc["ref_center"], S.IDENTITY, *cfg)
            d = S.pose_dist(pg, G["sa_pose"][si, ci])
            assert ng == G["sa_n"][si, ci] and d[0] < 1e-5 and d[1] < 1e-5, (seed, cfg, d)     # north_star tolerance; measured ~1e-15


@pytest.mark.gpu
def test_cuda_sparse_align_euroc_against_reference_golden(built, G):
    from dsdtm_b200 import capi
    cam = dict(S.EUROC)
    sc = H.make_scenario(K.SA_EUROC_SEED, cam)
    c = capi.Context(cam, levels=5, cell_size=15, max_feats=320, max_patches=8, max_frames=2, max_batch=1)
    c.upload(0, sc["ref_img"]); c.upload(1, sc["cur_img"])
    pg, ng, _ = c.sparse_align(0, 1, sc["feats"], sc["ref_center"], S.IDENTITY, 5, 0, 8)
    d = S.pose_dist(pg, G["sa_euroc_pose"])
    assert ng == int(G["sa_euroc_n"]) and d[0] < 1e-5 and d[1] < 1e-5, d
    cells = c.fast_cells(0, 20, 5.0)
    c.close()


@pytest.mark.gpu
def test_cuda_align2d_warp_detect_against_reference_golden(ctx, G):
    sc, levels, patches, truth, start = K.a2d_case()
    ctx.upload(1, sc["cur_img"])
    for j, iters in enumerate(K.A2D_ITERS):
        px, conv = ctx.align2d(1, levels, patches, start, iters)
        assert (conv == G["a2d_conv"][j].astype(bool)).all()
        assert np.abs(px - G["a2d_px"][j]).max() <= 1e-3                                         # north_star tolerance
    cases = K.warp_cases()
    ctx.upload(0, sc["ref_img"])
    A = np.array([c[1] for c in cases]); px = np.array([c[2] for c in cases]); L0 = np.array([c[3] for c in cases]); sl = np.array([c[4] for c in cases])
    got = ctx.warp_affine(np.zeros(len(cases), np.int32), A, px, L0, sl)
    assert (got == G["warp_patch"]).all()                                                        # bytes: bit-exact incl. the Q3 constant patches
    for name, cam, seed in K.DET_CASES:
        if cam is not S.KINECT:
            continue
        pr = S.make_pair(seed, dict(cam))
        ctx.upload(2, pr["ref_img"])
        cells = ctx.fast_cells(2, 20, 5.0)
        mask = np.full(pr["ref_img"].shape, 255, np.uint8)
        oc = np.zeros(len(cells), O.CORNER_DT)
        for k in ("x", "y", "level", "score"):
            oc[k] = cells[k]
        feats, _ = O.detect_select(oc, mask, 15, 300)          # sort + mask-circle selection (host side of a6) on the DEVICE cells
        assert (feats["x"] == G["det_%s_px" % name][:, 0]).all() and (feats["y"] == G["det_%s_px" % name][:, 1]).all()
        assert (feats["level"] == G["det_%s_level" % name]).all()
