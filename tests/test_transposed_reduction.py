"""Host emulation of the warp reductions of csrc/sparse_align.cu (DSDTM_SA_TREDUCE): the transposed reduction (lane l ends with the total of
value l; partner columns that are zero columns take the plain butterfly) pairs partial sums exactly like the xor butterfly it replaced, so
every total carries the SAME bits -- the claim behind "bit-equal" in profiles/r2_sparse_align.md. 32 lanes are emulated with numpy doubles."""
import numpy as np
import pytest


def butterfly(col):
    """the xor-shuffle butterfly over the 32 lanes of one value: every lane ends with the total"""
    v = col.copy()
    for o in (16, 8, 4, 2, 1):
        v = v + v[np.arange(32) ^ o]
    return v


def transposed(vals):
    """vals[lane, i], i < N. Mirrors warp_sum_transposed<N>: returns w[:, 0] (lane l < N holds the total of value l)."""
    n = vals.shape[1]
    w = np.zeros((32, 32))
    w[:, :n] = vals
    lane = np.arange(32)
    for o in (16, 8, 4, 2, 1):
        up = (lane & o) != 0
        for i in range(o):
            if i >= n:
                continue
            if i + o >= n:                                   # partner column is a zero column: plain butterfly
                w[:, i] = w[:, i] + w[lane ^ o, i]
                continue
            lo, hi = w[:, i].copy(), w[:, i + o].copy()
            keep = np.where(up, hi, lo)
            send = np.where(up, lo, hi)
            w[:, i] = keep + send[lane ^ o]
    return w[:, 0]


@pytest.mark.parametrize("n", [7, 21, 28, 32, 1, 5])
def test_transposed_reduction_has_the_butterflys_bits(n):
    rng = np.random.default_rng(n)
    for scale in (1.0, 1e-9, 1e12):
        vals = rng.standard_normal((32, n)) * scale * np.exp(rng.uniform(-20, 20, (32, n)))     # wide dynamic range: rounding matters
        tot = transposed(vals)
        for i in range(n):
            assert tot[i] == butterfly(vals[:, i])[0]        # same pairing tree => bit-equal (== on doubles), not merely close
