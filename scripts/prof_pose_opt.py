"""Driver for the pose-refinement stage (dsdtm_pose_optimize[_batch]): device time of one frame (300 matches) and
of a 4096-frame sweep from the context's stage timers (CUDA events on the launching stream), wall clock of the per-frame call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from dsdtm_b200 import capi, synth as S
import helpers as H


def main():
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    ctx = capi.Context(dict(S.KINECT), max_frames=2, max_batch=1)
    base = [H.make_ba_problem(300 + i, n=n, max_level=i % 4) for i in range(16)]
    obs = np.stack([H.ba_obs_records(base[i % 16], capi.BA_OBS_DT) for i in range(nf)])
    n_obs = np.full(nf, n, np.int32)
    poses = np.stack([base[i % 16]["pose_in"] for i in range(nf)])
    one = obs[1].copy()
    ctx.pose_optimize_batch(obs, n_obs, poses); ctx.pose_optimize(one, poses[1])       # warm-up
    ctx.profile(True)
    for _ in range(reps):
        out, res, sm = ctx.pose_optimize_batch(obs, n_obs, poses)
    ms, k = ctx.profile_get()["pose_opt"]
    it = float(sm["iterations"].mean())
    print("pose_opt batch: %d frames x %d obs: %.3f ms per launch = %.2f M frames/s, %.2f LM iterations per frame" % (nf, n, ms / k, nf / (ms / k) / 1e3, it))
    for _ in range(50):
        ctx.pose_optimize(one, poses[1])
    ms1, k1 = ctx.profile_get()["pose_opt"]
    ctx.set_option("pose_opt_solo_max", 0)
    for _ in range(50):
        ctx.pose_optimize(one, poses[1])
    msw, kw = ctx.profile_get()["pose_opt"]
    ctx.set_option("pose_opt_solo_max", -1)
    print("one frame, one warp (sweep kernel): %.1f us; CTA of eight warps (solo kernel): %.1f us" % (1e3 * msw / kw, 1e3 * ms1 / k1))
    for nb in (64, 296, 297, 1024):
        for _ in range(5):
            ctx.pose_optimize_batch(obs[:nb], n_obs[:nb], poses[:nb])
        msb, kb = ctx.profile_get()["pose_opt"]
        print("  %d frames: %.1f us per launch" % (nb, 1e3 * msb / kb))
    ctx.profile(False)
    t0 = time.perf_counter()
    for _ in range(200):
        ctx.pose_optimize(one, poses[1])
    wall = (time.perf_counter() - t0) / 200
    _, _, s1 = ctx.pose_optimize(one, poses[1])
    print("pose_opt one frame (%d obs, %d iterations): kernel %.1f us, call through the python binding %.1f us" % (n, s1["iterations"], 1e3 * ms1 / k1, 1e6 * wall))


if __name__ == "__main__":
    main()
