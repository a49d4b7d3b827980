"""GPU parity (bit-exact): kernel (a) pyramid and kernel (b) FAST + grid cells against the oracle and the goldens."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def _ctx_for(shape, levels=5, **kw):
    from dsdtm_b200 import capi
    h, w = shape
    cam = dict(width=w, height=h, fx=458.654, fy=457.296, cx=w / 2.0, cy=h / 2.0, f=458.654)
    return capi.Context(cam, levels=levels, cell_size=kw.pop("cell", 15), max_feats=64, max_patches=8, max_frames=kw.pop("frames", 4),
                        max_batch=kw.pop("batch", 4), **kw)


def test_pyramid_bit_exact_kinect(ctx, scenario):
    ctx.upload(0, scenario["ref_img"])
    packed, offs, ws, hs = scenario["ref_pyr"]
    for l in range(5):
        assert (ctx.download_level(0, l) == O.pyr_level(packed, offs, ws, hs, l)).all(), l


def test_pyramid_bit_exact_752x480_against_cv2_golden(golden):
    g = golden["test1_fast"]
    c = _ctx_for((480, 752))
    c.upload(1, g["img"])
    assert (c.download_level(1, 0) == g["img"]).all()
    for l in range(1, 5):
        assert (c.download_level(1, l) == g["pyr%d" % l]).all(), l     # 376, 188, 94, 47 wide: unaligned / odd widths
    c.close()


@pytest.mark.parametrize("shape", [(481, 641), (33, 140), (61, 81), (16, 128), (9, 9), (67, 256), (480, 1920), (37, 96), (12, 32), (200, 4096)])
def test_pyramid_odd_and_tiny_shapes(shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    levels = 3 if min(shape) >= 16 else 2
    c = _ctx_for(shape, levels=levels)
    c.upload(0, img)
    packed, offs, ws, hs = O.pyramid(img, levels)
    for l in range(levels):
        assert (c.download_level(0, l) == O.pyr_level(packed, offs, ws, hs, l)).all(), (shape, l)
    c.close()


def test_pyramid_tile_kernel_option_is_bit_exact_too(ctx, scenario):
    """"pyramid_kernel" = 1 forces the shared-memory tile kernel on every level, 2 the register strip kernel (the fallbacks of
    the default bulk-staged kernel)."""
    from dsdtm_b200 import capi
    packed, offs, ws, hs = scenario["cur_pyr"]
    for opt in (1, 2):
        ctx.set_option("pyramid_kernel", opt)
        try:
            ctx.upload(5, scenario["ref_img"])       # different content first, so a kernel that writes nothing is caught
            ctx.upload(5, scenario["cur_img"])
        finally:
            ctx.set_option("pyramid_kernel", 0)
        for l in range(5):
            assert (ctx.download_level(5, l) == O.pyr_level(packed, offs, ws, hs, l)).all(), (opt, l)
    with pytest.raises(capi.DsdtmError):
        ctx.set_option("pyramid_kernel", 7)
    with pytest.raises(capi.DsdtmError):
        ctx.set_option("no_such_option", 1)


def test_pyramid_async_upload_is_ordered_before_later_calls(ctx, scenario):
    """dsdtm_frame_upload_pyramid_async returns without a synchronisation; the caller's (pageable) buffer may be reused right away and
    every later call on the context sees the finished pyramid (ref: src/Frame.cpp:74-81 through the Frame adapter's constructor)."""
    packed, offs, ws, hs = scenario["ref_pyr"]
    cpacked = scenario["cur_pyr"][0]
    buf = scenario["ref_img"].copy()
    ctx.upload_async(6, buf)
    buf[:] = scenario["cur_img"]                     # overwrite the source immediately: the first image has already left the buffer
    ctx.upload_async(7, buf)
    buf[:] = 0
    for l in range(5):
        assert (ctx.download_level(6, l) == O.pyr_level(packed, offs, ws, hs, l)).all(), l
        assert (ctx.download_level(7, l) == O.pyr_level(cpacked, offs, ws, hs, l)).all(), l


def test_pyramid_batch_and_rebuild(ctx, scenario):
    imgs = np.stack([scenario["ref_img"], scenario["cur_img"], scenario["ref_img"][::-1].copy()])
    ctx.upload_batch(2, imgs)
    for i in range(3):
        packed, offs, ws, hs = O.pyramid(imgs[i], 5)
        for l in range(5):
            assert (ctx.download_level(2 + i, l) == O.pyr_level(packed, offs, ws, hs, l)).all()
    ctx.build_pyramid(2, 3)      # idempotent: rebuilding from level 0 gives the same levels
    ctx.sync()
    packed, offs, ws, hs = O.pyramid(imgs[2], 5)
    assert (ctx.download_level(4, 4) == O.pyr_level(packed, offs, ws, hs, 4)).all()


def _dense_from_lists(shape, xy, scores, keep):
    s = np.zeros(shape, np.uint8); m = np.zeros(shape, np.uint8)
    s[xy[:, 1], xy[:, 0]] = scores
    m[xy[keep, 1], xy[keep, 0]] = 1
    return s, m


def test_fast_kat_167_and_reference_lists_on_test1(golden):
    """FAST-10 @75 on test1.png = 167 corners (ref: Thirdparty/fast/test/test.cpp:52); @20 = 3787 -> 843 after non-max."""
    g = golden["test1_fast"]
    c = _ctx_for((480, 752))
    c.upload(0, g["img"])
    for b in (75, 20):
        s, m = c.fast_score_map(0, 0, b)
        ys, xs = np.nonzero(s)                       # raster order == the reference's output order
        xy = np.stack([xs, ys], 1).astype(np.int16)
        assert xy.shape == g["xy%d" % b].shape and (xy == g["xy%d" % b]).all()
        assert (s[ys, xs] == g["score%d" % b]).all()
        ky, kx = np.nonzero(m)
        kept = g["xy%d" % b][g["keep%d" % b]]
        assert len(ky) == len(kept) and (np.stack([kx, ky], 1) == kept).all()
    assert len(g["xy75"]) == 167
    c.close()


def test_fast_score_maps_all_levels(ctx, scenario):
    ctx.upload(0, scenario["ref_img"])
    packed, offs, ws, hs = scenario["ref_pyr"]
    for l in range(5):
        img = O.pyr_level(packed, offs, ws, hs, l)
        xy = O.fast10_detect(img, 20); sc = O.fast10_score(img, xy); keep = O.fast_nonmax(xy, sc)
        so, mo = _dense_from_lists(img.shape, xy, sc, keep)
        s, m = ctx.fast_score_map(0, l, 20)
        assert (s == so).all() and (m == mo).all(), l


@pytest.mark.parametrize("kind", ["noise", "binary", "blocks", "const"])
def test_fast_adversarial_images(kind):
    """saturation at 0/255, plateaus (mutual >= suppression), corners at x=3 / x=w-4."""
    rng = np.random.default_rng({"noise": 1, "binary": 2, "blocks": 3, "const": 4}[kind])
    shape = (70, 101)
    if kind == "noise":
        img = rng.integers(0, 256, shape, dtype=np.uint8)
    elif kind == "binary":
        img = (rng.integers(0, 2, shape) * 255).astype(np.uint8)
    elif kind == "blocks":
        img = np.repeat(np.repeat(rng.integers(0, 256, (24, 34), dtype=np.uint8), 3, 0), 3, 1)[:70, :101].copy()
    else:
        img = np.full(shape, 200, np.uint8)
    c = _ctx_for(shape, levels=2)
    c.upload(0, img)
    for b in (20, 1, 100):
        xy = O.fast10_detect(img, b)
        sc = O.fast10_score(img, xy) if len(xy) else np.zeros(0, np.int32)
        keep = O.fast_nonmax(xy, sc) if len(xy) else np.zeros(0, np.int32)
        so, mo = _dense_from_lists(shape, xy, sc, keep)
        s, m = c.fast_score_map(0, 0, b)
        assert (s == so).all() and (m == mo).all(), (kind, b)
    c.close()


def _cells_equal(a, b):
    return all((a[k] == b[k]).all() for k in ("x", "y", "level")) and (a["score"].view(np.uint32) == b["score"].view(np.uint32)).all()


def test_fast_cells_bit_exact_and_selection(ctx, scenario):
    """per-cell winners incl. Shi-Tomasi bits, then the host-side sort + mask selection gives the same feature list."""
    ctx.upload(0, scenario["ref_img"])
    packed, offs, ws, hs = scenario["ref_pyr"]
    for thr in (5.0, 20.0):
        want = O.detect_cells(packed, offs, ws, hs, 15, None, thr)
        got = ctx.fast_cells(0, 20, thr)
        assert _cells_equal(got, want), thr
    m1 = np.full((480, 640), 255, np.uint8); m2 = m1.copy()
    f1, _ = O.detect_select(O.detect_cells(packed, offs, ws, hs, 15, None, 5.0), m1, 15, 300)
    f2, _ = O.detect_select(ctx.fast_cells(0, 20, 5.0).view(O.CORNER_DT), m2, 15, 300)
    assert (f1 == f2).all() and (m1 == m2).all()


def test_fast_cells_with_occupancy_and_batch(ctx, scenario):
    imgs = np.stack([scenario["ref_img"], scenario["cur_img"]])
    ctx.upload_batch(0, imgs)
    rng = np.random.default_rng(9)
    occ = (rng.uniform(size=(2, ctx.n_cells)) < 0.4).astype(np.uint8)
    got = ctx.fast_cells(0, 20, 5.0, occupied=occ, n=2)
    for i in range(2):
        packed, offs, ws, hs = O.pyramid(imgs[i], 5)
        want = O.detect_cells(packed, offs, ws, hs, 15, occ[i], 5.0)
        assert _cells_equal(got[i], want), i
        assert (got[i]["score"][occ[i] == 1] == np.float32(5.0)).all()


def test_fast_cells_euroc_geometry(golden):
    g = golden["test1_fast"]
    c = _ctx_for((480, 752), cell=30)
    c.upload(0, g["img"])
    packed, offs, ws, hs = O.pyramid(g["img"], 5)
    assert _cells_equal(c.fast_cells(0, 20, 5.0), O.detect_cells(packed, offs, ws, hs, 30, None, 5.0))
    c.close()


def test_bad_arguments_return_errors(ctx):
    from dsdtm_b200 import capi
    with pytest.raises(capi.DsdtmError):
        ctx.download_level(99, 0)
    import ctypes
    buf = np.zeros(16, np.uint8)
    assert ctx.L.dsdtm_fast_score_map(ctx.hp, 0, 7, 20, capi._p(buf), capi._p(buf)) == -1        # DSDTM_E_ARG
    assert ctx.L.dsdtm_fast_cells(ctx.hp, 0, 0, ctypes.c_float(5.0), None, capi._p(buf)) == -1   # barrier out of range
    assert ctx.L.dsdtm_batch_run(ctx.hp, 0) in (-4, 0)
