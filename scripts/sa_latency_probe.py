"""Single-pair sparse-alignment latency vs iteration cap: separates the fixed cost (launch, level staging) from the per-iteration cost."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsdtm_b200 import capi, synth as S, workload as W

def main():
    cam = dict(S.KINECT)
    ctx = capi.Context(cam, levels=5, max_feats=320, max_patches=300, max_frames=4, max_batch=1)
    batch = W.build_batch(ctx, cam, 1, scenes=W.render_scenes(1, cam, procs=1), n_feats=300, feat_stride=320, patches_per_pair=300)
    for iters in (0, 1, 2, 4, 8, 30):
        ctx.batch_stage(batch["ref_slots"], batch["cur_slots"], batch["feats"], batch["n_feats"], batch["centers"], batch["poses_in"], 4, 0, iters,
                        batch["patches"], batch["patch_px"], batch["patch_level"], 10)
        for _ in range(5): ctx.batch_run(0)
        ctx.sync(); ctx.profile(True); ctx.profile_get(reset=True)
        for _ in range(50): ctx.batch_run(0)
        st = ctx.profile_get(reset=True); ctx.profile(False)
        _, _, log, nlog = ctx.sparse_align_batch(batch["ref_slots"], batch["cur_slots"], batch["feats"], batch["n_feats"], batch["centers"], batch["poses_in"], 4, 0, iters, log_cap=64)
        print("max_iters %2d: sparse_align %.1f us, GN iterations executed %d" % (iters, st["sparse_align"][0] / 50 * 1e3, int(nlog[0])))

if __name__ == "__main__":
    main()
