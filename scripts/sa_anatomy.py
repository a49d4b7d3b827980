"""Per-iteration anatomy of sparse_align_kernel (needs a -DDSDTM_SA_TIMING build: the log's x[3..5] carry clock64 deltas of lane 0 of
warp 0): cycles of the feature pass, of the reductions + barrier up to the tail, and of the one-lane tail (solve + SE3 update + log).
Run on the GPU box:  DSDTM_NVCC_FLAGS=-DDSDTM_SA_TIMING python dsdtm_b200/build.py --force && python scripts/sa_anatomy.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsdtm_b200 import capi, synth as S, workload as W


def main():
    cam = dict(S.KINECT)
    for B, wpps in ((1, (10,)), (4096, (3,))):
        ctx = capi.Context(cam, levels=5, max_feats=320, max_patches=300, max_frames=2 * B + 2, max_batch=B)
        batch = W.build_batch(ctx, cam, B, scenes=W.render_scenes(2, cam, procs=1), n_feats=300, feat_stride=320, patches_per_pair=300)
        for wpp in wpps:
            ctx.set_option("sa_warps_per_pair", wpp)
            for rep in range(2):
                _, _, log, nlog = ctx.sparse_align_batch(batch["ref_slots"], batch["cur_slots"], batch["feats"], batch["n_feats"], batch["centers"], batch["poses_in"], 4, 0, 30, log_cap=64)
            rows = np.concatenate([log[i, :nlog[i]] for i in range(0, B, max(1, B // 64))])
            first = rows[rows["iter"] == 0]; later = rows[rows["iter"] > 0]
            for name, r in (("first iteration of a level", first), ("later iterations", later)):
                c = r["x"][:, 3:6]
                print("pairs %4d wpp %2d %-27s n=%4d  pass %7.0f  reduce+barrier %6.0f  tail %6.0f cycles (median) | tail share %.0f %%" % (
                    B, wpp, name, len(r), np.median(c[:, 0]), np.median(c[:, 1]), np.median(c[:, 2]), 100 * np.median(c[:, 2]) / np.median(c.sum(1))))
            print("pairs %4d wpp %2d level staging %7.0f cycles (median), prologue %7.0f" % (B, wpp, np.median(first["x"][:, 2]),
                  np.median(rows[(rows["iter"] == 0) & (rows["level"] == 3)]["x"][:, 1])))
        ctx.close()


if __name__ == "__main__":
    main()
