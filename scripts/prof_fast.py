"""Driver for profiling the FAST stage alone: 256 frames (128 x ref, 128 x cur of one scene), fast_cells_batch x N."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsdtm_b200 import capi, synth as S, workload as W

def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    cam = dict(S.KINECT); ctx = capi.Context(cam, max_frames=260, max_batch=256)
    sc = W.render_scenes(1, cam, procs=1)[0]
    ctx.upload_batch(0, np.stack([sc["ref_img"]] * 128)); ctx.upload_batch(128, np.stack([sc["cur_img"]] * 128))
    cells = np.zeros(256 * ctx.n_cells, capi.CORNER_DT)
    ctx.profile(True)
    for _ in range(reps):
        ctx._ck(ctx.L.dsdtm_fast_cells_batch(ctx.hp, 0, 256, 20, capi.C.c_float(5.0), None, capi._p(cells)))
    ms, n = ctx.profile_get()["fast"]
    print("fast 256 frames: %.3f ms per launch, cells>20: %d" % (ms / n, int((cells["score"] > 20).sum())))

if __name__ == "__main__":
    main()
