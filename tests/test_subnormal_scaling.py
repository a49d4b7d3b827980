"""The arithmetic fact behind DSDTM_SA_CVT 3 of csrc/sparse_align.cu, checked in IEEE double on the CPU (numpy): a byte encoded as the
subnormal b * 2^-1034 times a bilinear weight scaled by 2^1010 gives the reference's product times 2^-24 with the SAME rounding, sums of
such products are the scaled sums, and squares / products of two scaled values are restored exactly by 2^48."""
import numpy as np


def _subnormal(b):
    # the kernel's PRMT: byte into bits 8..15 of the high word, exponent field 0
    return (np.asarray(b, np.uint64) << np.uint64(40)).view(np.float64)


def test_byte_as_subnormal_is_proportional_to_the_byte():
    b = np.arange(256, dtype=np.uint64)
    d = _subnormal(b)
    assert d[0] == 0.0 and not np.signbit(d[0])
    assert (d == b.astype(np.float64) * 2.0 ** -1034).all()          # exact: a subnormal with 8 significant bits


def test_scaled_bilinear_sample_rounds_like_the_reference():
    rng = np.random.default_rng(3)
    n = 200000
    sx = rng.random(n); sy = rng.random(n)
    sx[:1000] = 0.0; sy[500:1500] = 0.0                               # integer-pixel features: weights exactly 0 and 1
    sx[2000:2100] = 2.0 ** -46                                        # the smallest fraction a projected pixel can carry at |u| < 128
    b = rng.integers(0, 256, (4, n)).astype(np.uint64)
    b[:, :5000] = rng.choice([0, 255], (4, 5000))
    w = [(1 - sx) * (1 - sy), sx * (1 - sy), (1 - sx) * sy, sx * sy]                     # ref: src/Sprase_ImageAlign.cpp:129-132
    ref = ((w[0] * b[0] + w[1] * b[1]) + w[2] * b[2]) + w[3] * b[3]                      # ref: :147-148, left to right
    WS = 2.0 ** 1010
    osy, sys_ = (1 - sy) * WS, sy * WS                                                  # the kernel scales ONE factor of each weight
    ws = [(1 - sx) * osy, sx * osy, (1 - sx) * sys_, sx * sys_]
    d = _subnormal(b)
    got = ((ws[0] * d[0] + ws[1] * d[1]) + ws[2] * d[2]) + ws[3] * d[3]
    assert np.isfinite(got).all()
    assert (got * 2.0 ** 24 == ref).all()                             # bit-equal after the exact restore
    # residuals, central differences and their products: scaled by 2^-24 resp. 2^-48, restored exactly
    ref2 = np.roll(ref, 1); got2 = np.roll(got, 1)
    assert ((got - got2) * 2.0 ** 24 == (ref - ref2)).all()
    assert (((got - got2) * (got - got2)) * 2.0 ** 48 == (ref - ref2) * (ref - ref2)).all()
    assert (np.cumsum((got - got2) * got)[-1] * 2.0 ** 48 == np.cumsum((ref - ref2) * ref)[-1])
