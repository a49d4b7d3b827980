"""Generates the golden fixtures in this directory. Run ONCE in the build container (needs cv2, /root/reference and
oracle/_ref/libfast_ref.so = the reference's own FAST sources compiled by oracle/Makefile):

    PYTHONPATH=. python tests/golden/make_golden.py

Outputs (committed):
  test1_fast.npz   : gray test1.png (ref: Thirdparty/fast/test/data/test1.png, the image of the reference's only KAT,
                     Thirdparty/fast/test/test.cpp:10-13,52), the REFERENCE library's corner lists at barriers 75 (167 corners)
                     and 20, scores and non-max survivors; cv2.pyrDown chain of the image (levels 1..4).
  pyrdown_cv2.npz  : cv2.pyrDown outputs for small seeded images incl. odd sizes (inputs stored too).
  circle_cv2.npz   : cv2.circle(img, c, r, 0, -1) masks (bit-packed) for a list of centres/radii incl. clipped ones.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402

img = cv2.imread("/root/reference/Thirdparty/fast/test/data/test1.png", 0)
assert img.shape == (480, 752)
out = {"img": img}
for b in (75, 20):
    xy = O.ref_fast10_detect(img, b, True)
    xy_plain = O.ref_fast10_detect(img, b, False)
    assert (xy == xy_plain).all()
    sc = O.ref_fast10_score(img, xy, b)
    keep = O.ref_fast_nonmax(xy, sc)
    out["xy%d" % b] = xy; out["score%d" % b] = sc.astype(np.int16); out["keep%d" % b] = keep.astype(np.int32)
assert len(out["xy75"]) == 167
lv = img
for l in range(1, 5):
    lv = cv2.pyrDown(lv)
    out["pyr%d" % l] = lv
np.savez_compressed(os.path.join(HERE, "test1_fast.npz"), **out)

rng = np.random.default_rng(20260101)
pd = {}
for i, (h, w) in enumerate([(60, 94), (30, 47), (61, 81), (7, 5), (3, 3), (16, 128), (33, 140)]):
    a = rng.integers(0, 256, (h, w), dtype=np.uint8)
    if i % 2:
        a = cv2.GaussianBlur(a, (5, 5), 1.2)
    pd["in%d" % i] = a
    pd["out%d" % i] = cv2.pyrDown(a)
np.savez_compressed(os.path.join(HERE, "pyrdown_cv2.npz"), **pd)

cases = []
masks = []
H, W = 60, 80
for r in (1, 2, 3, 7, 15, 20, 30):
    for (cx, cy) in [(40, 30), (0, 0), (79, 59), (5, 55), (-3, 20), (85, 30), (40, -10), (77, 3), (12, 70)]:
        m = np.full((H, W), 255, np.uint8)
        cv2.circle(m, (cx, cy), r, 0, -1)
        cases.append((cx, cy, r)); masks.append(np.packbits(m == 0))
np.savez_compressed(os.path.join(HERE, "circle_cv2.npz"), cases=np.array(cases, np.int32), masks=np.array(masks), shape=np.array([H, W]))
print("golden fixtures written")
