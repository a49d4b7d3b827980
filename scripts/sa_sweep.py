"""Sparse-alignment kernel sweep on one staged batch: (variant, warps per pair) -> ms per launch (CUDA events through the stage
profiler) + bitwise comparison of the poses against variant 0 at the same warps-per-pair."""
import argparse, hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsdtm_b200 import capi, synth as S, workload as W
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--combos", default="0:3,0:4,0:5,0:10")
    a = ap.parse_args()
    cam = dict(S.KINECT)
    B = a.pairs
    ctx = capi.Context(cam, levels=bench.LEVELS, cell_size=15, max_feats=bench.FEAT_STRIDE, max_patches=bench.N_FEATS, max_frames=2 * B + 2, max_batch=B)
    batch = W.build_batch(ctx, cam, B, scenes=W.render_scenes(8, cam), n_feats=bench.N_FEATS, feat_stride=bench.FEAT_STRIDE, patches_per_pair=bench.N_FEATS)
    ctx.batch_stage(batch["ref_slots"], batch["cur_slots"], batch["feats"], batch["n_feats"], batch["centers"], batch["poses_in"],
                    bench.ALIGN_CFG["max_level"], bench.ALIGN_CFG["min_level"], bench.ALIGN_CFG["max_iters"], batch["patches"], batch["patch_px"],
                    batch["patch_level"], bench.ALIGN2D_ITERS)
    base = {}
    for combo in a.combos.split(","):
        v, w = (int(t) for t in combo.split(":"))
        ctx.set_option("sa_variant", v); ctx.set_option("sa_warps_per_pair", w)
        ctx.profile(False)
        for _ in range(2):
            ctx.batch_run(1)
        ctx.sync(); ctx.profile(True); ctx.profile_get(reset=True)
        for _ in range(a.steps):
            ctx.batch_run(1)
        ctx.sync()
        st = ctx.profile_get(reset=True)
        poses, nt, px, conv = ctx.batch_fetch()
        if v == 0:
            base[w] = (poses.copy(), nt.copy())
        same = "-" if w not in base else ("bit-equal to variant 0" if (poses == base[w][0]).all() and (nt == base[w][1]).all() else "DIFFERS max %.3g" % np.abs(poses - base[w][0]).max())
        err = np.median([S.pose_dist(poses[i], batch["truth"][i]) for i in range(min(B, 64))], 0)
        dig = hashlib.sha1(poses.tobytes() + nt.tobytes()).hexdigest()[:12]     # equal digests across libraries = bit-equal poses and counts
        print("variant %d wpp %2d: sparse_align %.4f ms per launch (%d pairs) | pyramid %.4f align2d %.4f | %s | median pose err %.2e rad %.2e m | sha1 %s" % (
            v, w, st["sparse_align"][0] / a.steps, B, st["pyramid"][0] / a.steps, st["align2d"][0] / a.steps, same, err[0], err[1], dig), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
