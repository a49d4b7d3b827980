"""Small driver: CLAHE + pyramid over n frames (profile events around the two CLAHE kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsdtm_b200 import capi, synth as S, workload as W


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    cam = dict(S.KINECT)
    ctx = capi.Context(cam, levels=5, cell_size=15, max_feats=64, max_patches=8, max_frames=n, max_batch=1)
    sc = W.render_scenes(1, cam, procs=1)[0]
    imgs = np.ascontiguousarray(np.stack([np.roll(sc["cur_img"], k, axis=1) for k in range(8)] * (n // 8)))
    ctx.upload_clahe(0, imgs, 3.0, (8, 8), fetch=False)
    ctx.profile(True); ctx.profile_get()
    for _ in range(5):
        ctx.upload_clahe(0, imgs, 3.0, (8, 8), fetch=False)
    st = ctx.profile_get()
    ms = st["ingest"][0] / 5
    print("clahe %d frames: %.3f ms per call (lut + apply kernels) = %.1f GB/s algorithmic (3 B/px)" % (n, ms, n * 640 * 480 * 3 / ms / 1e6))
    ctx.close()


if __name__ == "__main__":
    main()
