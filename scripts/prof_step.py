"""Small driver for ncu: stages a batch of pairs and replays the step a few times (no oracle, no e2e)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsdtm_b200 import capi, synth as S, workload as W
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=148)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--scenes", type=int, default=2)
    ap.add_argument("--fast", type=int, default=0, help="also run fast_cells_batch over this many frames")
    ap.add_argument("--wpp", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--chunks", type=int, default=1)
    ap.add_argument("--pyr", type=int, default=0, help="pyramid_kernel option (0 auto, 1 tile)")
    ap.add_argument("--direct", action="store_true", help="launch kernels directly instead of graph replay")
    a = ap.parse_args()
    cam = dict(S.KINECT)
    B = a.pairs
    ctx = capi.Context(cam, levels=bench.LEVELS, cell_size=15, max_feats=bench.FEAT_STRIDE, max_patches=bench.N_FEATS, max_frames=2 * B + 2, max_batch=B)
    batch = W.build_batch(ctx, cam, B, scenes=W.render_scenes(a.scenes, cam, procs=1), n_feats=bench.N_FEATS, feat_stride=bench.FEAT_STRIDE, patches_per_pair=bench.N_FEATS)
    ctx.batch_stage(batch["ref_slots"], batch["cur_slots"], batch["feats"], batch["n_feats"], batch["centers"], batch["poses_in"],
                    bench.ALIGN_CFG["max_level"], bench.ALIGN_CFG["min_level"], bench.ALIGN_CFG["max_iters"], batch["patches"], batch["patch_px"],
                    batch["patch_level"], bench.ALIGN2D_ITERS)
    ctx.set_option("sa_warps_per_pair", a.wpp)
    ctx.set_option("sa_variant", a.variant)
    ctx.set_option("step_chunks", a.chunks)
    ctx.set_option("pyramid_kernel", a.pyr)
    if a.direct:
        ctx.profile(True)
    if not a.direct:
        for _ in range(3):
            ctx.batch_run(1)
        ctx.sync()
        ctx.timer_start()
    for _ in range(a.steps):
        ctx.batch_run(1)
    if not a.direct:
        print("graph replay: %.4f ms per step (chunks=%d)" % (ctx.timer_stop() / a.steps, a.chunks))
    ctx.sync()
    if a.direct:
        print(ctx.profile_get())
    poses, nt, px, conv = ctx.batch_fetch()
    print("pose err", np.median([S.pose_dist(poses[i], batch["truth"][i]) for i in range(min(B, 32))], 0), "tracked", nt[:4], "conv", conv.mean())
    if a.fast:
        cells = np.zeros(a.fast * ctx.n_cells, capi.CORNER_DT)
        ctx._ck(ctx.L.dsdtm_fast_cells_batch(ctx.hp, 0, a.fast, 20, capi.C.c_float(5.0), None, capi._p(cells)))
        print("fast cells > 20:", int((cells["score"] > 20).sum()))
    ctx.close()


if __name__ == "__main__":
    main()
