"""Golden fixtures for the keyframe-ingest restatements. Run ONCE in the build container (needs cv2):

    PYTHONPATH=. python tests/golden/make_golden_ingest.py

Output (committed):
  undistort_cv2.npz : cv2.undistortPoints(src, K, dist, None, None, K) called the way ref: src/Frame.cpp:121-122 calls it --
                      K (3x3) and dist (5x1: k1 k2 p1 p2 k3) as CV_32F (ref: src/Camera.cpp:53-68), src CV_32FC2 -- for the
                      distortion sets of the reference's own configs (Config/EuRoc.yaml:16-20, Config/default.yaml:39-43),
                      an all-zero set (Config/Rpg_uzh.yaml) and an exaggerated one that drives the iteration hard.
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CAMS = {
    "euroc":   dict(w=752, h=480, fx=458.654, fy=457.296, cx=367.215, cy=248.375, dist=(-0.28368365, 0.07451284, -0.00010473, -3.55590700e-05, 0.0)),
    "default": dict(w=640, h=480, fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, dist=(0.231222, -0.784899, -0.003257, -0.000105, 0.917205)),
    "zero":    dict(w=752, h=480, fx=329.115520046, fy=329.115520046, cx=376.0, cy=240.0, dist=(0.0, 0.0, 0.0, 0.0, 0.0)),
    "strong":  dict(w=640, h=480, fx=300.0, fy=310.0, cx=320.5, cy=239.5, dist=(-0.45, 0.25, 0.004, -0.003, -0.07)),
}
rng = np.random.default_rng(20260101)
out = {}
for name, c in CAMS.items():
    K = np.array([[c["fx"], 0, c["cx"]], [0, c["fy"], c["cy"]], [0, 0, 1]], np.float32)
    D = np.array(c["dist"], np.float32).reshape(5, 1)
    n = 600
    src = np.stack([rng.uniform(0, c["w"] - 1, n), rng.uniform(0, c["h"] - 1, n)], 1).astype(np.float32)
    src[:4] = [[0, 0], [c["w"] - 1, c["h"] - 1], [c["cx"], c["cy"]], [3.0, c["h"] - 4.0]]     # corners, principal point, FAST border
    src[4:300] = np.round(src[4:300])                                                          # detector output is integer-valued
    dst = cv2.undistortPoints(src.reshape(n, 1, 2), K, D, None, None, K).reshape(n, 2)
    out[name + "_K"] = K; out[name + "_D"] = D.ravel(); out[name + "_src"] = src; out[name + "_dst"] = dst
    out[name + "_wh"] = np.array([c["w"], c["h"]], np.int32)
np.savez_compressed(os.path.join(HERE, "undistort_cv2.npz"), **out)
print("undistort_cv2.npz written, cv2", cv2.__version__)

# ---- CLAHE (ref: Test/test_Feature_detection.cpp:85-86, Test/test_Euroc.cpp:64, Test/test_Optimizer.cpp:75: createCLAHE(3.0, 8x8))
#   clahe_cv2.npz : cv2.createCLAHE(clip, tiles).apply(img) for the deterministic numpy inputs of tests/helpers.py (clahe_input)
import sys
sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
import helpers as Hh  # noqa: E402

cl = {}
for name, h, w, clip, tiles in Hh.CLAHE_CASES:
    img = Hh.clahe_input(name, h, w)
    cl[name] = cv2.createCLAHE(clip, tiles).apply(img)
np.savez_compressed(os.path.join(HERE, "clahe_cv2.npz"), **cl)
print("clahe_cv2.npz written")
