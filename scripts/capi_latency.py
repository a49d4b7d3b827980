"""Wall-clock cost of the three single-frame C-ABI calls an adapter makes per tracked frame (host buffers, synchronous)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dsdtm_b200 import capi, synth as S, workload as W
import bench


def main():
    cam = dict(S.KINECT)
    ctx = capi.Context(cam, levels=bench.LEVELS, cell_size=15, max_feats=bench.FEAT_STRIDE, max_patches=bench.N_FEATS, max_frames=8, max_batch=2)
    batch = W.build_batch(ctx, cam, 2, scenes=W.render_scenes(1, cam, procs=1), n_feats=bench.N_FEATS, feat_stride=bench.FEAT_STRIDE, patches_per_pair=bench.N_FEATS)
    nf = int(batch["n_feats"][0]); img = np.ascontiguousarray(batch["scenes"][0]["cur_img"])
    rs, cs = int(batch["ref_slots"][0]), int(batch["cur_slots"][0])
    feats = np.ascontiguousarray(batch["feats"][0][:nf]); cen = batch["centers"][0].copy(); pose = batch["poses_in"][0].copy()
    lv = batch["patch_level"][0].copy(); pt = batch["patches"][0].copy(); px = batch["patch_px"][0].copy()
    calls = {"upload+pyramid": lambda: ctx.upload(cs, img),
             "sparse_align": lambda: ctx.sparse_align(rs, cs, feats, cen, pose, 4, 0, 30, log_cap=1),
             "align2d": lambda: ctx.align2d(cs, lv, pt, px, 10)}
    for name, f in calls.items():
        for _ in range(10):
            f()
        t0 = time.perf_counter()
        for _ in range(200):
            f()
        print("%-16s %.1f us per call" % (name, (time.perf_counter() - t0) / 200 * 1e6))
    # the same calls with the ctypes argument marshalling hoisted out of the loop: what a C++ caller pays
    C = capi.C
    L = ctx.L
    pose_out = np.empty(7); ntr = C.c_int(0); nlog = C.c_int(0); log = np.zeros(1, capi.ITER_LOG_DT); conv = np.zeros(len(lv), np.uint8)
    px_io = px.copy()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    raw = {"upload+pyramid": (L.dsdtm_frame_upload_pyramid, (ctx.hp, cs, P(img), img.shape[1])),
           "sparse_align": (L.dsdtm_sparse_align, (ctx.hp, rs, cs, P(feats), nf, P(cen), P(pose), 4, 0, 30, P(pose_out), C.byref(ntr), P(log), 1, C.byref(nlog))),
           "align2d": (L.dsdtm_align2d_batch, (ctx.hp, cs, P(lv), P(pt), P(px_io), len(lv), 10, P(conv)))}
    for name, (fn, args) in raw.items():
        for _ in range(10):
            px_io[...] = px
            assert fn(*args) == 0
        t = 0.0
        for _ in range(200):
            px_io[...] = px
            t0 = time.perf_counter()
            fn(*args)
            t += time.perf_counter() - t0
        print("%-16s %.1f us per call (pre-marshalled arguments)" % (name, t / 200 * 1e6))
    # the production configuration of Tracking (ref: src/Tracking.cpp:37: Sprase_ImgAlign(MaxPyraLevels, MinPyraLevels, 8)) with the log
    log2 = np.zeros(256, capi.ITER_LOG_DT)
    a2 = (ctx.hp, rs, cs, P(feats), nf, P(cen), P(pose), 5, 0, 8, P(pose_out), C.byref(ntr), P(log2), 256, C.byref(nlog))
    for _ in range(10):
        assert L.dsdtm_sparse_align(*a2) == 0
    ctx.profile(True); ctx.profile_get()
    t0 = time.perf_counter()
    for _ in range(100):
        L.dsdtm_sparse_align(*a2)
    dt = (time.perf_counter() - t0) / 100 * 1e6
    st = ctx.profile_get(); ctx.profile(False)
    print("sparse_align(5,0,8) + log: %.1f us per call, kernel %.1f us, %d GN iterations" % (dt, st["sparse_align"][0] / st["sparse_align"][1] * 1e3, nlog.value))
    ctx.profile(True); ctx.profile_get()
    for _ in range(50):
        for f in calls.values():
            f()
    st = ctx.profile_get()
    print({k: round(v[0] / max(v[1], 1) * 1e3, 1) for k, v in st.items() if v[1]})
    ctx.close()


if __name__ == "__main__":
    main()
