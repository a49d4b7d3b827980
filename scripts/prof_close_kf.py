"""Tracking::GetCloseKeyFrames on the device-resident map table (dsdtm_close_keyframes): kernel and call time against the map
size, with the CPU port (oracle, one thread) beside it."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from dsdtm_b200 import capi, synth as S
import helpers as H
import oracle as O


def main():
    cam = dict(S.KINECT)
    ctx = capi.Context(cam, max_frames=2, max_batch=1)
    for n_kfs in (50, 1000, 4096, 16384):
        mt = H.make_map_table(9, n_kfs=n_kfs, pts_per_kf=(200, 300), spread=max(6.0, 0.6 * n_kfs ** 0.5))
        rows = np.zeros(n_kfs, capi.MAP_KF_DT); rows["pt_begin"] = mt["pt_begin"]; rows["pt_count"] = mt["pt_count"]; rows["t"] = mt["kf_t"]
        t0 = time.perf_counter(); ctx.map_table_upload(0, rows, 0, mt["points"]); up = time.perf_counter() - t0
        for _ in range(3):
            v, d, l = ctx.close_keyframes(mt["pose_cur"], n_kfs)
        ctx.profile(True); ctx.profile_get(reset=True)
        t0 = time.perf_counter()
        for _ in range(20):
            ctx.close_keyframes(mt["pose_cur"], n_kfs)
        wall = (time.perf_counter() - t0) / 20
        ms, k = ctx.profile_get()["local_map"]
        ctx.profile(False)
        t0 = time.perf_counter()
        vo, do, lo = O.close_keyframes(H.ocam(cam), mt["pose_cur"], mt["pt_begin"], mt["pt_count"], mt["kf_t"], mt["points"])
        cpu = time.perf_counter() - t0
        assert (v == vo).all() and (l == lo).all()
        rows_read = int(mt["pt_count"][v == 0].sum()) + 32 * int(v.sum())
        print("%6d key frames, %8d point rows (%5.1f MB), %4d close: kernel %7.1f us (%.0f GB/s of rows actually read), call %7.1f us, one-time upload %.1f ms, CPU port %8.1f us"
              % (n_kfs, len(mt["points"]), len(mt["points"]) * 24 / 1e6, int(v.sum()), 1e3 * ms / k, rows_read * 24 / (ms / k * 1e-3) / 1e9, 1e6 * wall, 1e3 * up, 1e6 * cpu))


if __name__ == "__main__":
    main()
