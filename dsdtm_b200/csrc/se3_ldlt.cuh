// se3_ldlt.cuh -- the serial tail of one Gauss-Newton iteration: 6x6 pivoted LDL^T solve (Eigen's ldlt() semantics)
// and T <- T * SE3::exp(x) (Sophus non-templated semantics). Written so that EVERY array index is a compile-time
// constant after unrolling: the whole state lives in registers (round-1 profile: the first version kept A[36] in local
// memory and this section was 70 % of the kernel, profiles/r1_sparse_align_v1.md).
// __host__ __device__ so that the CPU suite can check it against the oracle without a GPU.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define DSDTM_HD __host__ __device__ __forceinline__
#else
#define DSDTM_HD inline
#endif

namespace dsdtm {

DSDTM_HD double dsdtm_rsqrt(double v)
{
#if defined(__CUDA_ARCH__)
    return rsqrt(v);
#else
    return 1.0 / sqrt(v);
#endif
}

template <class T>
DSDTM_HD void cswap(bool p, T& a, T& b)
{
    const T ta = p ? b : a;
    const T tb = p ? a : b;
    a = ta; b = tb;
}

// H: full symmetric 6x6 (row-major, both triangles filled). Pivoted LDL^T (largest remaining |diagonal|, first wins),
// then x = P^T L^-T D^+ L^-1 P b with Eigen's rule "pivot <= 1/highest -> component 0".
DSDTM_HD void ldlt6_solve_reg(const double (&Hin)[6][6], const double (&bin)[6], double (&x)[6])
{
    double A[6][6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) A[i][j] = Hin[i][j];
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = bin[i];
    int tr[6] = { 0, 1, 2, 3, 4, 5 };
    bool zero_matrix = false;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        if (zero_matrix) continue;      // Eigen: an all-zero matrix stops the factorisation at k == 0 with identity transpositions
        // pivot search over the remaining diagonal (first maximum wins)
        int big = k;
        double bigv = fabs(A[k][k]);
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            const double v = fabs(A[i][i]);
            const bool g = v > bigv;
            bigv = g ? v : bigv;
            big = g ? i : big;
        }
        tr[k] = big;
        // symmetric swap of rows/cols k and big in the lower triangle, written as predicated swaps over every candidate
        // so that all register-array indices stay compile-time constants
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            const bool p = (big == i);
#pragma unroll
            for (int j = 0; j < k; ++j) cswap(p, A[k][j], A[i][j]);        // rows k / big, columns left of k
#pragma unroll
            for (int r = i + 1; r < 6; ++r) cswap(p, A[r][k], A[r][i]);    // columns k / big, rows below big
            cswap(p, A[k][k], A[i][i]);
#pragma unroll
            for (int r = k + 1; r < i; ++r) cswap(p, A[r][k], A[i][r]);    // the "elbow" between k and big
        }
        if (k > 0) {
            double temp[6];
#pragma unroll
            for (int j = 0; j < k; ++j) temp[j] = A[j][j] * A[k][j];
            double s = 0;
#pragma unroll
            for (int j = 0; j < k; ++j) s += A[k][j] * temp[j];
            A[k][k] -= s;
#pragma unroll
            for (int i = k + 1; i < 6; ++i) {
                double s2 = 0;
#pragma unroll
                for (int j = 0; j < k; ++j) s2 += A[i][j] * temp[j];
                A[i][k] -= s2;
            }
        }
        const double akk = A[k][k];
        const bool valid = fabs(akk) > 0.0;
        if (k == 0 && !valid) {
            zero_matrix = true;
            tr[0] = 0;
        } else if (k < 5 && valid) {
#pragma unroll
            for (int i = k + 1; i < 6; ++i) A[i][k] = A[i][k] / akk;
        }
    }
    // P b
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int i = k + 1; i < 6; ++i) cswap(tr[k] == i, y[k], y[i]);
    // L^-1
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) y[i] -= A[i][j] * y[j];
    const double tol = 1.0 / 1.7976931348623157e308;
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = (fabs(A[i][i]) > tol) ? y[i] / A[i][i] : 0.0;
    // L^-T
#pragma unroll
    for (int i = 5; i >= 0; --i)
#pragma unroll
        for (int j = i + 1; j < 6; ++j) y[i] -= A[j][i] * y[j];
    // P^T
#pragma unroll
    for (int k = 5; k >= 0; --k)
#pragma unroll
        for (int i = k + 1; i < 6; ++i) cswap(tr[k] == i, y[k], y[i]);
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = y[i];
}

// Fast path: unpivoted LDL^T in registers (~150 instructions instead of ~2500 for the predicated-swap pivoted version:
// the pivoted code alone is larger than the 32 KB instruction cache, profiles/r1_sparse_align_v2.md). For a symmetric
// positive definite H both factorisations solve the same system and differ only in rounding (backward stable either way).
// Returns false -- caller falls back to ldlt6_solve_reg, i.e. Eigen's exact pivoted semantics -- when a pivot is not
// safely positive (rank-deficient / indefinite / NaN input).
DSDTM_HD bool ldlt6_solve_spd(const double (&A)[6][6], const double (&b)[6], double (&x)[6])
{
    double L[6][6], d[6];
    double maxdiag = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) maxdiag = fmax(maxdiag, fabs(A[i][i]));
    const double thresh = 1e-11 * maxdiag;
    bool ok = maxdiag > 0.0 && maxdiag < 1.7976931348623157e308;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double dk = A[k][k];
#pragma unroll
        for (int j = 0; j < k; ++j) dk -= L[k][j] * L[k][j] * d[j];
        d[k] = dk;
        ok = ok && (dk > thresh);
        const double inv = 1.0 / dk;
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            double v = A[i][k];
#pragma unroll
            for (int j = 0; j < k; ++j) v -= L[i][j] * L[k][j] * d[j];
            L[i][k] = v * inv;
        }
    }
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double v = b[i];
#pragma unroll
        for (int j = 0; j < i; ++j) v -= L[i][j] * y[j];
        y[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = y[i] / d[i];
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double v = y[i];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) v -= L[j][i] * x[j];
        x[i] = v;
    }
    return ok;
}

// Split form of ldlt6_solve_spd for callers that reuse one factorisation for several right-hand sides (the sparse-alignment
// H only changes when the visibility set changes): factor once into 15 strictly-lower entries of L (row-major packed) and the
// 6 pivots d, then substitute per right-hand side. Same arithmetic and operation order as ldlt6_solve_spd.
DSDTM_HD bool ldlt6_factor_spd(const double (&A)[6][6], double (&Lp)[15], double (&dinv)[6])
{
    double d[6];
    double L[6][6];
    double maxdiag = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) maxdiag = fmax(maxdiag, fabs(A[i][i]));
    const double thresh = 1e-11 * maxdiag;
    bool ok = maxdiag > 0.0 && maxdiag < 1.7976931348623157e308;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double dk = A[k][k];
#pragma unroll
        for (int j = 0; j < k; ++j) dk -= L[k][j] * L[k][j] * d[j];
        d[k] = dk;
        ok = ok && (dk > thresh);
        const double inv = 1.0 / dk;
        dinv[k] = inv;                       // the substitution multiplies by the reciprocal pivot (one division per pivot, none per solve)
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            double v = A[i][k];
#pragma unroll
            for (int j = 0; j < k; ++j) v -= L[i][j] * L[k][j] * d[j];
            L[i][k] = v * inv;
        }
    }
#pragma unroll
    for (int i = 1; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) Lp[i * (i - 1) / 2 + j] = L[i][j];
    return ok;
}

DSDTM_HD void ldlt6_subst_spd(const double (&Lp)[15], const double (&dinv)[6], const double (&b)[6], double (&x)[6])
{
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double v = b[i];
#pragma unroll
        for (int j = 0; j < i; ++j) v -= Lp[i * (i - 1) / 2 + j] * y[j];
        y[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = y[i] * dinv[i];
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double v = y[i];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) v -= Lp[j * (j - 1) / 2 + i] * x[j];
        x[i] = v;
    }
}

struct Quat { double w, x, y, z; };

// out = T * exp(x); pose7 = {qw,qx,qy,qz,tx,ty,tz}. (ref: src/Sprase_ImageAlign.cpp:335; Sophus SE3::exp, SE3::operator*=)
// Coefficient k (highest power first) of series j of se3_mul_exp below, as the flat list {k = 0: j = 0..3, k = 1: j = 0..3, ...}: j = 0
// cos(theta/2) and 1 2 sin(theta/2)/theta in (theta/2)^2, 2 (1 - cos theta)/theta^2 and 3 (theta - sin theta)/theta^3 in theta^2. The same
// constant expressions as in the Horner chains of se3_mul_exp: the same doubles (tests/test_host_math.py checks the two forms bit for bit).
#define DSDTM_SE3_SERIES_COEF_LIST \
    -1.0 / 87178291200.0, -1.0 / 1307674368000.0, -1.0 / 20922789888000.0, -1.0 / 355687428096000.0, \
    1.0 / 479001600.0, 1.0 / 6227020800.0, 1.0 / 87178291200.0, 1.0 / 1307674368000.0, \
    -1.0 / 3628800.0, -1.0 / 39916800.0, -1.0 / 479001600.0, -1.0 / 6227020800.0, \
    1.0 / 40320.0, 1.0 / 362880.0, 1.0 / 3628800.0, 1.0 / 39916800.0, \
    -1.0 / 720.0, -1.0 / 5040.0, -1.0 / 40320.0, -1.0 / 362880.0, \
    1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, \
    -0.5, -1.0 / 6.0, -1.0 / 24.0, -1.0 / 120.0, \
    1.0, 1.0, 0.5, 1.0 / 6.0

DSDTM_HD double se3_theta2(const double (&x)[6])
{
    const double o0 = x[3], o1 = x[4], o2 = x[5];
    return o0 * o0 + o1 * o1 + o2 * o2;
}
DSDTM_HD bool se3_exp_uses_series(double t2)
{
#ifndef DSDTM_SE3_SERIES
#define DSDTM_SE3_SERIES 1
#endif
    return DSDTM_SE3_SERIES && t2 < 0.25;
}

// `series4` (optional): the values of the four power series below for se3_theta2(x), evaluated elsewhere with the same Horner steps (the
// sparse-alignment kernel spreads them over four lanes of the warp that runs the solve, csrc/sparse_align.cu); only read when
// se3_exp_uses_series(se3_theta2(x)).
DSDTM_HD void se3_mul_exp(const double (&T)[7], const double (&x)[6], double (&out)[7], const double* series4 = nullptr)
{
    // This runs on ONE lane while the rest of the CTA waits, so the dependent chain is kept short: one sincos (the full-angle
    // values come from the half-angle identities), reciprocals instead of repeated divisions. Against the literal Sophus
    // sequence (the oracle) the result differs by a few ulp (tests/test_host_math.py: <= 1e-15 absolute).
    const double SMALL_EPS = 1e-10;
    const double u0 = x[0], u1 = x[1], u2 = x[2], o0 = x[3], o1 = x[4], o2 = x[5];
    const double t2 = se3_theta2(x);
    double theta, ch, imag, a, b;
    double ew, ex, ey, ez;
    const bool series = se3_exp_uses_series(t2);
    if (series) {
        // |theta| < 0.5 rad (every Gauss-Newton step of a tracker; larger updates take the generic branch below): the four scalar
        // functions of Sophus' exp are even power series in theta,
        //   cos(theta/2), sin(theta/2)/theta, (1 - cos theta)/theta^2, (theta - sin theta)/theta^3,
        // so neither the square root, nor sincos, nor the division by theta is needed: four independent Horner chains of eight terms
        // (truncation < 1e-18 relative at theta = 0.5) instead of a dependent sqrt -> sincos -> 1/theta sequence of several hundred
        // cycles on the one lane the whole CTA waits for. (Sophus' own Taylor terms for theta < 1e-10 vanish in fp64: same imag.)
        const double h2 = 0.25 * t2;                                     // (theta/2)^2
        if (series4) { ch = series4[0]; imag = 0.5 * series4[1]; a = series4[2]; b = series4[3]; }
        else {
        ch = fma(h2, fma(h2, fma(h2, fma(h2, fma(h2, fma(h2, fma(h2, -1.0 / 87178291200.0, 1.0 / 479001600.0), -1.0 / 3628800.0), 1.0 / 40320.0),
                                             -1.0 / 720.0), 1.0 / 24.0), -0.5), 1.0);
        imag = 0.5 * fma(h2, fma(h2, fma(h2, fma(h2, fma(h2, fma(h2, fma(h2, -1.0 / 1307674368000.0, 1.0 / 6227020800.0), -1.0 / 39916800.0), 1.0 / 362880.0),
                                                 -1.0 / 5040.0), 1.0 / 120.0), -1.0 / 6.0), 1.0);
        a = fma(t2, fma(t2, fma(t2, fma(t2, fma(t2, fma(t2, fma(t2, -1.0 / 20922789888000.0, 1.0 / 87178291200.0), -1.0 / 479001600.0), 1.0 / 3628800.0),
                                        -1.0 / 40320.0), 1.0 / 720.0), -1.0 / 24.0), 0.5);
        b = fma(t2, fma(t2, fma(t2, fma(t2, fma(t2, fma(t2, fma(t2, -1.0 / 355687428096000.0, 1.0 / 1307674368000.0), -1.0 / 6227020800.0), 1.0 / 39916800.0),
                                        -1.0 / 362880.0), 1.0 / 5040.0), -1.0 / 120.0), 1.0 / 6.0);
        }
        // only tested against SMALL_EPS below. Sophus switches to V = R(q) for theta < 1e-10 (first-order different from the series:
        // R u = u + w x u, V u = u + w x u / 2); reproduced, the reference's step is what counts
        theta = (t2 < SMALL_EPS * SMALL_EPS) ? 0.0 : 1.0;
        ew = ch; ex = imag * o0; ey = imag * o1; ez = imag * o2;
        // Sophus' SO3(Quaternion) constructor normalises; |q|^2 is within a few ulp of 1 here, where 1/sqrt(n) = 1.5 - 0.5 n to O((n-1)^2)
        const double rn = fma(-0.5, ex * ex + ey * ey + ez * ez + ew * ew, 1.5);
        ex *= rn; ey *= rn; ez *= rn; ew *= rn;
    } else {
        theta = sqrt(t2);
        const double half = 0.5 * theta;
        double sh;
#if defined(__CUDA_ARCH__)
        sincos(half, &sh, &ch);
#else
        sh = sin(half); ch = cos(half);
#endif
        if (theta < SMALL_EPS) {
            const double t4 = t2 * t2;
            imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
            a = 0.0; b = 0.0;
        } else {
            const double it = 1.0 / theta;
            imag = sh * it;
            const double ct = 2.0 * ch * ch - 1.0, st = 2.0 * sh * ch;     // cos(theta), sin(theta)
            const double it2 = it * it;
            a = (1 - ct) * it2;                                            // (1 - cos theta) / theta^2
            b = (theta - st) * (it2 * it);                                 // (theta - sin theta) / theta^3
        }
        ew = ch; ex = imag * o0; ey = imag * o1; ez = imag * o2;
        const double rn = dsdtm_rsqrt(ex * ex + ey * ey + ez * ez + ew * ew);
        ex *= rn; ey *= rn; ez *= rn; ew *= rn;
    }
    // V = I + a*Omega + b*Omega^2 (or R(e) for tiny theta);  et = V * upsilon
    double et0, et1, et2;
    if (theta < SMALL_EPS) {
        const double tx = 2 * ex, ty = 2 * ey, tz = 2 * ez;
        const double twx = tx * ew, twy = ty * ew, twz = tz * ew;
        const double txx = tx * ex, txy = ty * ex, txz = tz * ex;
        const double tyy = ty * ey, tyz = tz * ey, tzz = tz * ez;
        et0 = (1 - (tyy + tzz)) * u0 + (txy - twz) * u1 + (txz + twy) * u2;
        et1 = (txy + twz) * u0 + (1 - (txx + tzz)) * u1 + (tyz - twx) * u2;
        et2 = (txz - twy) * u0 + (tyz + twx) * u1 + (1 - (txx + tyy)) * u2;
    } else {
        // Omega = [0 -o2 o1; o2 0 -o0; -o1 o0 0]; Omega^2 entries
        const double O2_00 = -o2 * o2 - o1 * o1, O2_01 = o1 * o0, O2_02 = o2 * o0;
        const double O2_10 = o0 * o1, O2_11 = -o2 * o2 - o0 * o0, O2_12 = o2 * o1;
        const double O2_20 = o0 * o2, O2_21 = o1 * o2, O2_22 = -o1 * o1 - o0 * o0;
        const double V00 = 1.0 + b * O2_00, V01 = a * -o2 + b * O2_01, V02 = a * o1 + b * O2_02;
        const double V10 = a * o2 + b * O2_10, V11 = 1.0 + b * O2_11, V12 = a * -o0 + b * O2_12;
        const double V20 = a * -o1 + b * O2_20, V21 = a * o0 + b * O2_21, V22 = 1.0 + b * O2_22;
        et0 = V00 * u0 + V01 * u1 + V02 * u2;
        et1 = V10 * u0 + V11 * u1 + V12 * u2;
        et2 = V20 * u0 + V21 * u1 + V22 * u2;
    }
    // T * E : t = t_T + R(q_T) et ; q = normalize(q_T * q_E)
    const double aw = T[0], ax = T[1], ay = T[2], az = T[3];
    double uv0 = ay * et2 - az * et1, uv1 = az * et0 - ax * et2, uv2 = ax * et1 - ay * et0;
    uv0 += uv0; uv1 += uv1; uv2 += uv2;
    const double c0 = ay * uv2 - az * uv1, c1 = az * uv0 - ax * uv2, c2 = ax * uv1 - ay * uv0;
    const double rw = aw * ew - ax * ex - ay * ey - az * ez;
    const double rx = aw * ex + ax * ew + ay * ez - az * ey;
    const double ry = aw * ey + ay * ew + az * ex - ax * ez;
    const double rz = aw * ez + az * ew + ax * ey - ay * ex;
    // |q_T q_E|^2 = |q_T|^2 |q_E|^2: within rounding of 1 whenever the pose handed in is a unit quaternion (every pose this library
    // produces); then 1/sqrt(n) = 1 - d/2 + 3 d^2/8 with d = n - 1 is exact to O(d^3) < 1e-24. A pose that is not normalised takes rsqrt.
    const double n2 = rx * rx + ry * ry + rz * rz + rw * rw;
    const double dn = n2 - 1.0;
    const double rn = (fabs(dn) < 1e-8) ? fma(dn, fma(dn, 0.375, -0.5), 1.0) : dsdtm_rsqrt(n2);
    out[0] = rw * rn; out[1] = rx * rn; out[2] = ry * rn; out[3] = rz * rn;
    out[4] = T[4] + (et0 + aw * uv0 + c0);
    out[5] = T[5] + (et1 + aw * uv1 + c1);
    out[6] = T[6] + (et2 + aw * uv2 + c2);
}

}  // namespace dsdtm
