"""CPU check of the INDEXING of pyrdown_small_kernel (dsdtm_b200/csrc/pyramid.cu): planes staged at an 8-byte aligned pitch with the
reflect-101 border materialised (columns -2, -1 and w .. w + 10), groups of four outputs read columns c0 - 2 .. c0 + 10 with the
interior formula only, the produced level becomes the next source in shared memory. A numpy emulation of exactly that plan against
the oracle's cv::pyrDown restatement -- the bit-exact GPU run is tests/test_gpu_pyramid_fast.py; this is the host-side proof that
no group ever reads a column the border fill did not write, for ragged and tiny shapes."""
import ctypes as C

import numpy as np
import pytest

import oracle as O

FP = 8


def _reflect101(i, n):
    if n == 1:
        return 0
    while i < 0 or i >= n:
        i = -i if i < 0 else 2 * n - 2 - i
    return i


def _pitch(w):
    return (FP + w + 11 + 7) & ~7


def _fill_border(plane, w, h, p):
    for r in range(h):
        base = r * p + FP
        for lane in range(13):
            col = lane - 2 if lane < 2 else w + (lane - 2)
            plane[base + col] = plane[base + _reflect101(col, w)]


def _chain(img, nlev):
    h, w = img.shape
    POISON = 1 << 20                                         # any read of a byte nobody wrote shows up in the result
    A = np.full(1 << 18, POISON, np.int64); B = np.full(1 << 18, POISON, np.int64)
    p = _pitch(w)
    for r in range(h):
        A[r * p + FP:r * p + FP + w] = img[r]
    _fill_border(A, w, h, p)
    cur, nxt = A, B
    out_levels = []
    for l in range(nlev):
        dw, dh = (w + 1) // 2, (h + 1) // 2
        p = _pitch(w); G = (dw + 3) >> 2
        assert 8 * (G - 1) + 11 <= w + 10 and FP + w + 11 <= p           # the reach of the last group stays inside the filled border
        sh = np.zeros((h, G, 4), np.int64)
        for r in range(h):
            for g in range(G):
                v = cur[r * p + FP + 8 * g - 2:r * p + FP + 8 * g + 11]
                for j in range(4):
                    c = 2 * j + 2
                    sh[r, g, j] = v[c - 2] + 4 * v[c - 1] + 6 * v[c] + 4 * v[c + 1] + v[c + 2]
        out = np.zeros((dh, dw), np.int64); npitch = _pitch(dw)
        keep = l + 1 < nlev
        for y in range(dh):
            r = 2 * y
            rows = [_reflect101(r - 2, h), _reflect101(r - 1, h), r, _reflect101(r + 1, h), _reflect101(r + 2, h)]
            s = sh[rows[0]] + sh[rows[4]] + 4 * (sh[rows[1]] + sh[rows[3]]) + 6 * sh[rows[2]]
            o = ((s + 128) >> 8).reshape(-1)
            out[y] = o[:dw]
            if keep:
                nxt[y * npitch + FP:y * npitch + FP + 4 * G] = o          # columns >= dw: rewritten by the border fill
        out_levels.append(out)
        if keep:
            _fill_border(nxt, dw, dh, npitch)
        cur, nxt = nxt, cur
        w, h = dw, dh
    return out_levels


@pytest.mark.parametrize("shape,levels", [((120, 188), 2), ((60, 94), 1), ((30, 47), 2), ((13, 13), 2), ((5, 7), 1), ((3, 3), 1),
                                          ((32, 12), 2), ((9, 100), 3), ((120, 94), 3)])
def test_small_chain_plan_matches_pyrdown(shape, levels):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    got = _chain(img, levels)
    src = img
    for l in range(levels):
        hh, ww = src.shape
        want = np.zeros(((hh + 1) // 2, (ww + 1) // 2), np.uint8)
        O.lib().orc_pyrdown_u8(src.ctypes.data_as(C.c_void_p), ww, hh, ww, want.ctypes.data_as(C.c_void_p))
        assert (got[l] == want).all(), (shape, l)
        src = want
