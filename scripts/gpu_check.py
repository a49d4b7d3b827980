"""Scratch GPU parity check (first bring-up). Run under gpurun."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle as O
from dsdtm_b200 import synth as S, capi
import helpers as H

cam = S.KINECT
sc = H.make_scenario(20260101)
ctx = capi.Context(cam, levels=5, max_feats=320, max_patches=320, max_frames=4, max_batch=4)
ctx.upload(0, sc["ref_img"]); ctx.upload(1, sc["cur_img"])
packed, offs, ws, hs = sc["ref_pyr"]
ok = True
for l in range(5):
    g = ctx.download_level(0, l); o = O.pyr_level(packed, offs, ws, hs, l)
    eq = (g == o).all(); ok &= eq
    print("pyramid level", l, g.shape, "bit-exact" if eq else "MISMATCH %d" % (g != o).sum())
for l in range(5):
    img = O.pyr_level(packed, offs, ws, hs, l)
    s, m = ctx.fast_score_map(0, l, 20)
    xy = O.fast10_detect(img, 20); sco = O.fast10_score(img, xy); keep = O.fast_nonmax(xy, sco)
    so = np.zeros_like(s); so[xy[:, 1], xy[:, 0]] = sco
    mo = np.zeros_like(m); mo[xy[keep, 1], xy[keep, 0]] = 1
    e1 = (s == so).all(); e2 = (m == mo).all(); ok &= e1 and e2
    print("fast level", l, "corners", len(xy), "kept", len(keep), "score", e1, "nonmax", e2)
cells = ctx.fast_cells(0, 20, 5.0)
ocells = O.detect_cells(packed, offs, ws, hs, 15, None, 5.0)
eq = all((cells[k] == ocells[k]).all() for k in ("x", "y", "level")) and (cells["score"].view(np.uint32) == ocells["score"].view(np.uint32)).all()
ok &= eq
print("fast cells bit-exact", eq, "n>20:", (cells["score"] > 20).sum())
if not eq:
    bad = np.nonzero((cells["x"] != ocells["x"]) | (cells["score"] != ocells["score"]))[0][:10]
    for b in bad: print("  ", b, cells[b], ocells[b])
# sparse align
oc = H.ocam(cam)
for (ml, it) in ((4, 30), (5, 8)):
    po, no, lo = O.sparse_align(oc, sc["ref_pyr"][0], sc["cur_pyr"][0], offs, ws, hs, sc["feats"], sc["ref_center"], S.IDENTITY, ml, 0, it)
    t = time.time(); pg, ng, lg = ctx.sparse_align(0, 1, sc["feats"], sc["ref_center"], S.IDENTITY, ml, 0, it); dt = time.time() - t
    d = S.pose_dist(po, pg)
    print("sparse align", (ml, it), "oracle n", no, "gpu n", ng, "iters", len(lo), len(lg), "pose diff rad/m", d, "truth err", S.pose_dist(pg, sc["T_c2r"]), "%.2f ms" % (dt * 1e3))
    for a, b in zip(lo, lg):
        rel = abs(a["chi2"] - b["chi2"]) / abs(a["chi2"])
        if rel > 1e-9 or a["flags"] != b["flags"] or a["n_pts"] != b["n_pts"]:
            print("   L%d it%d chi2 %.12g vs %.12g rel %.2e flags %d/%d n %d/%d" % (a["level"], a["iter"], a["chi2"], b["chi2"], rel, a["flags"], b["flags"], a["n_pts"], b["n_pts"]))
    ok &= d[0] < 1e-5 and d[1] < 1e-5 and no == ng and len(lo) == len(lg)
# align2d
levels, patches, truth, start = H.make_patches(sc["cur_pyr"], 300, 7, max_level=2)
px_g, conv_g = ctx.align2d(1, levels, patches, start, 10)
nbad = 0; maxd = 0
for i in range(300):
    L = int(levels[i])
    p, c, nit = O.align2d(O.pyr_level(sc["cur_pyr"][0], offs, ws, hs, L), patches[i], 10, start[i])
    d = np.abs(p - px_g[i]).max(); maxd = max(maxd, d)
    if c != conv_g[i] or d > 1e-3: nbad += 1
print("align2d: max |gpu-oracle| px", maxd, "bad", nbad, "converged", conv_g.sum(), "err vs truth (median)", np.median(np.linalg.norm(px_g - truth, axis=1)[conv_g]))
ok &= nbad == 0
# warp affine
rng = np.random.default_rng(3)
n = 200
A = np.tile(np.eye(2), (n, 1, 1)) + rng.uniform(-0.2, 0.2, (n, 2, 2))
rl = rng.integers(0, 3, n).astype(np.int32); sl = rng.integers(0, 3, n).astype(np.int32)
sl[: n // 2] = 0
rpx = np.stack([rng.uniform(1, 639, n), rng.uniform(1, 479, n)], 1).astype(np.float32)
rpx[:5] = [[0.2, 0.3], [639, 479], [638.9, 100], [320, 478.99], [3, 3]]
wg = ctx.warp_affine(np.zeros(n, np.int32), A, rpx, rl, sl)
nb = 0
for i in range(n):
    wo = O.warp_affine(A[i], O.pyr_level(packed, offs, ws, hs, int(rl[i])), rpx[i], int(rl[i]), int(sl[i]))
    if not (wo == wg[i]).all(): nb += 1
print("warp affine mismatching patches:", nb, "/", n)
ok &= nb == 0
print("ALL OK" if ok else "FAILURES")
