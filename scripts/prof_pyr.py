"""Small driver for ncu: the batched pyramid launches only (n frames of random content, levels 1..4 rebuilt `reps` times)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsdtm_b200 import capi, synth as S


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2072)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--pyr", type=int, default=0)
    ap.add_argument("--cam", default="kinect", choices=["kinect", "euroc"])
    a = ap.parse_args()
    cam = dict(S.EUROC if a.cam == "euroc" else S.KINECT)
    ctx = capi.Context(cam, levels=5, cell_size=15, max_feats=64, max_patches=8, max_frames=a.frames, max_batch=1)
    ctx.set_option("pyramid_kernel", a.pyr)
    rng = np.random.default_rng(0)
    base = rng.integers(0, 256, (8, cam["height"], cam["width"]), dtype=np.uint8)
    imgs = np.ascontiguousarray(np.tile(base, (a.frames // 8 + 1, 1, 1))[:a.frames])
    ctx.upload_batch(0, imgs)                      # 4 launches (levels 1..4 of all frames)
    ctx.profile(True); ctx.profile_get()
    for _ in range(a.reps):
        ctx.build_pyramid(0, a.frames)             # 4 launches each
    ctx.sync()
    print(ctx.profile_get())
    ctx.close()


if __name__ == "__main__":
    main()
