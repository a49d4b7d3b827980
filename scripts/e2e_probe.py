"""Where does the e2e step spend its time? 2-D strided H2D of the cur images vs 1-D copies vs the whole call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsdtm_b200 import capi, synth as S
import ctypes as C

def main():
    cam = dict(S.KINECT)
    B = 2048
    ctx = capi.Context(cam, levels=5, max_feats=320, max_patches=300, max_frames=B + 2, max_batch=B)
    imgs = capi.pinned_empty((B, 480, 640), np.uint8)
    imgs[:] = 7
    for rep in range(2):
        t = time.perf_counter(); ctx.upload_batch(0, imgs); dt = time.perf_counter() - t
    print("upload_batch (2-D strided H2D + pyramid): %.2f ms  %.1f GB/s" % (dt * 1e3, imgs.nbytes / dt / 1e9))
    pageable = np.full((B, 480, 640), 7, np.uint8)
    t = time.perf_counter(); ctx.upload_batch(0, pageable); dt = time.perf_counter() - t
    print("same from pageable memory: %.2f ms  %.1f GB/s" % (dt * 1e3, imgs.nbytes / dt / 1e9))
    import torch
    d = torch.empty(imgs.nbytes, dtype=torch.uint8, device="cuda")
    h = torch.from_numpy(imgs.reshape(-1))
    for rep in range(2):
        torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("torch 1-D copy of the same pinned buffer: %.2f ms  %.1f GB/s (is_pinned=%s)" % (dt * 1e3, imgs.nbytes / dt / 1e9, h.is_pinned()))
    ctx.close()

if __name__ == "__main__":
    main()
