"""CPU suite: the product's register-only 6x6 pivoted LDL^T and SE3 update (dsdtm_b200/csrc/se3_ldlt.cuh, compiled for the
host) against the oracle's Eigen/Sophus restatement and numpy."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def prod():
    d = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, "libse3_ldlt_host.so")
    src = os.path.join(ROOT, "tests", "cpp", "se3_ldlt_host.cpp")
    hdr = os.path.join(ROOT, "dsdtm_b200", "csrc", "se3_ldlt.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", src, "-o", so])
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _solve(prod, H, b):
    x = np.empty(6)
    prod.prod_ldlt6_solve(_p(np.ascontiguousarray(H, np.float64)), _p(np.ascontiguousarray(b, np.float64)), _p(x))
    return x


def test_ldlt_matches_oracle_bitwise_and_numpy(prod):
    rng = np.random.default_rng(0)
    for t in range(200):
        J = rng.normal(size=(40, 6)) * rng.uniform(0.01, 100, 6)      # badly scaled columns force pivoting
        H = J.T @ J
        b = rng.normal(size=6)
        x = _solve(prod, H, b)
        xo = O.ldlt6_solve(H, b)
        assert (x == xo).all(), t                                       # same algorithm, same operation order
        assert np.allclose(x, np.linalg.solve(H, b), rtol=1e-7, atol=0)


def test_ldlt_singular_and_zero_cases(prod):
    Z = np.zeros((6, 6))
    assert (_solve(prod, Z, np.ones(6)) == 0).all() and (O.ldlt6_solve(Z, np.ones(6)) == 0).all()   # Q2: nothing visible
    rng = np.random.default_rng(1)
    J = rng.normal(size=(3, 6))                                         # rank 3
    H = J.T @ J
    b = J.T @ rng.normal(size=3)
    x, xo = _solve(prod, H, b), O.ldlt6_solve(H, b)
    assert np.array_equal(x, xo, equal_nan=True)
    Hn = H.copy(); Hn[0, 0] = np.nan
    assert np.array_equal(_solve(prod, Hn, b), O.ldlt6_solve(Hn, b), equal_nan=True)      # NaN handling identical
    Hn = H.copy(); Hn[3, 1] = Hn[1, 3] = np.nan
    xn = _solve(prod, Hn, b)
    assert np.isnan(xn).any() and np.array_equal(xn, O.ldlt6_solve(Hn, b), equal_nan=True)


def test_se3_update_matches_oracle(prod):
    rng = np.random.default_rng(2)
    for t in range(50):
        T = O.se3_exp(rng.uniform(-0.5, 0.5, 6))
        x = rng.uniform(-0.05, 0.05, 6) if t else np.zeros(6)
        if t == 1:
            x[3:] = 1e-12                                                # small-angle branch
        out = np.empty(7)
        prod.prod_se3_mul_exp(_p(T), _p(np.ascontiguousarray(x)), _p(out))
        want = O.se3_mul(T, O.se3_exp(x))
        assert np.allclose(out, want, rtol=0, atol=1e-15)


def test_se3_update_with_series_from_the_coefficient_table_is_bit_equal(prod):
    """The sparse-alignment kernel's tail (csrc/sparse_align.cu, DSDTM_SA_TAIL_CONST) evaluates the four power series of SE3::exp from a
    coefficient table and hands them to se3_mul_exp; the Horner steps and the doubles are those of se3_mul_exp's own chains, so the
    pose must come out with the same bits -- for tracker-sized steps, tiny ones, and large ones that take the generic branch."""
    rng = np.random.default_rng(9)
    for t in range(400):
        T = O.se3_exp(rng.uniform(-0.5, 0.5, 6))
        scale = (0.05, 1e-6, 1e-12, 0.45, 1.5)[t % 5]                   # 1.5: |omega|^2 >= 0.25, no series
        x = rng.uniform(-scale, scale, 6)
        a = np.empty(7); b = np.empty(7)
        prod.prod_se3_mul_exp(_p(T), _p(np.ascontiguousarray(x)), _p(a))
        prod.prod_se3_mul_exp_table(_p(T), _p(np.ascontiguousarray(x)), _p(b))
        assert (a == b).all(), (t, np.abs(a - b).max())
