"""Where one frame's pose refinement spends its time: kernel time of the solo kernel against the iteration cap and the number
of observations (stage timers = CUDA events on the launching stream)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dsdtm_b200 import capi, synth as S, workload as W


def main():
    ctx = capi.Context(dict(S.KINECT), max_frames=2, max_batch=1)
    for n in (300, 32, 2048):
        pr = W.make_ba_problem(301, n=n, max_level=1)
        o = np.zeros(n, capi.BA_OBS_DT); o["normal"] = pr["normals"]; o["point_w"] = pr["points_w"]; o["level"] = pr["levels"]
        for cap in (0, 1, 2, 4, 8, 100):
            for _ in range(3):
                ctx.pose_optimize(o, pr["pose_in"], cap)
            ctx.profile(True); ctx.profile_get(reset=True)
            for _ in range(20):
                _, _, s = ctx.pose_optimize(o, pr["pose_in"], cap)
            ms, k = ctx.profile_get()["pose_opt"]
            ctx.profile(False)
            print("n=%4d cap=%3d iterations=%2d accepted=%2d kernel %.1f us" % (n, cap, s["iterations"], s["n_successful"], 1e3 * ms / k))


if __name__ == "__main__":
    main()
