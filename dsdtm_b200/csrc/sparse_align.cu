// sparse_align.cu -- kernel (c): coarse-to-fine inverse-compositional Gauss-Newton on SE3, one CTA per frame pair.
//
// Replaces Sprase_ImgAlign::GetJocabianMat / ComputeResiduals / GaussNewtonSolver and the level loop of Run
// (ref: src/Sprase_ImageAlign.cpp:43-55, 62-193, 240-344). fp64 throughout, like the reference.
//
// Design (DESIGN.md "Sparse alignment"):
//   * one CTA per frame pair, one THREAD per reference feature; the whole level x iteration loop runs on the device
//     (the chain of <= levels*max_iters dependent GN steps never returns to the host);
//   * inverse-compositional structure is exploited: the Jacobian row of pixel p of feature j is
//         J_jp = (dx_jp * a_j + dy_jp * b_j) * (f * scale)           (ref: :160; a_j, b_j = rows of GetJocabianBA(P_j))
//     so  sum_p J_jp J_jp^T = (f*scale)^2 (Sxx a a^T + Sxy (a b^T + b a^T) + Syy b b^T) does not depend on the pose:
//     it is reduced over the VISIBLE features only when the visibility set changes, and the per-iteration work is
//         b += (f*scale) (a_j sum_p dx_jp r_jp + b_j sum_p dy_jp r_jp),   chi2 += sum_p r_jp^2
//     i.e. 3 FMAs per residual instead of 27. Same mathematics, different summation order than the reference's
//     sequential feature-major/pixel-minor accumulation (documented tolerance: chi2 1e-4 rel, pose 1e-5);
//   * the 16 reference samples and their (dx, dy) per feature live in shared memory as [pixel][feature] columns
//     (conflict-free: thread j only touches column j), everything else in registers;
//   * block reduction = warp xor-shuffle butterfly (bitwise identical in all lanes) + one shared-memory row per warp,
//     summed in warp order by warp 0 => run-to-run deterministic;
//   * lane 0 of warp 0 solves the 6x6 system with the same pivoted LDL^T as Eigen's ldlt(), applies SE3::exp and the
//     reference's accept / revert / converge rules, and publishes the new pose through shared memory.
#include "ctx.cuh"

namespace dsdtm {

namespace {

struct SaArgs {
    const uint8_t* frames; unsigned frame_stride; LevelGeom geo;
    const int* ref_slots; const int* cur_slots;
    const dsdtm_ref_feat* feats; int feat_stride; const int* n_feats;
    const double* centers; const double* poses_in; double* poses_out; int* n_tracked;
    dsdtm_iter_log* log; int* n_log; int log_cap;
    float fx, fy, cx, cy, f;
    int max_level, min_level, max_iters;
    int nf;      // shared-memory column count == blockDim.x
    int pair0;
};

// ---------------------------------------------------------------- SE3 (Sophus non-templated semantics) ----------
struct Quat { double w, x, y, z; };

__device__ __forceinline__ void qrot(const Quat& q, const double v[3], double out[3])  // Eigen _transformVector
{
    double uv0 = __dsub_rn(__dmul_rn(q.y, v[2]), __dmul_rn(q.z, v[1]));
    double uv1 = __dsub_rn(__dmul_rn(q.z, v[0]), __dmul_rn(q.x, v[2]));
    double uv2 = __dsub_rn(__dmul_rn(q.x, v[1]), __dmul_rn(q.y, v[0]));
    uv0 = __dadd_rn(uv0, uv0); uv1 = __dadd_rn(uv1, uv1); uv2 = __dadd_rn(uv2, uv2);
    const double c0 = __dsub_rn(__dmul_rn(q.y, uv2), __dmul_rn(q.z, uv1));
    const double c1 = __dsub_rn(__dmul_rn(q.z, uv0), __dmul_rn(q.x, uv2));
    const double c2 = __dsub_rn(__dmul_rn(q.x, uv1), __dmul_rn(q.y, uv0));
    out[0] = __dadd_rn(__dadd_rn(v[0], __dmul_rn(q.w, uv0)), c0);
    out[1] = __dadd_rn(__dadd_rn(v[1], __dmul_rn(q.w, uv1)), c1);
    out[2] = __dadd_rn(__dadd_rn(v[2], __dmul_rn(q.w, uv2)), c2);
}

__device__ void quat_to_R(const Quat& q, double R[9])
{
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// pose7 = {qw,qx,qy,qz,tx,ty,tz};  out = T * exp(x)   (ref: :335; Sophus SE3::exp, SE3::operator*)
__device__ void se3_mul_exp(const double* T, const double x[6], double* out)
{
    const double SMALL_EPS = 1e-10;
    const double* ups = x;
    const double* om = x + 3;
    const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    const double half = 0.5 * theta;
    double imag;
    const double real = cos(half);
    if (theta < SMALL_EPS) {
        const double t2 = theta * theta, t4 = t2 * t2;
        imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
    } else {
        imag = sin(half) / theta;
    }
    Quat e = { real, imag * om[0], imag * om[1], imag * om[2] };
    {
        const double n = sqrt(e.x * e.x + e.y * e.y + e.z * e.z + e.w * e.w);
        e.x /= n; e.y /= n; e.z /= n; e.w /= n;
    }
    const double O[9] = { 0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0 };
    double V[9];
    if (theta < SMALL_EPS) {
        quat_to_R(e, V);
    } else {
        double O2[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) O2[3 * i + j] = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
        const double t2 = theta * theta;
        const double a = (1 - cos(theta)) / t2;
        const double b = (theta - sin(theta)) / (t2 * theta);
#pragma unroll
        for (int i = 0; i < 9; ++i) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + a * O[i] + b * O2[i];
    }
    double et[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) et[i] = V[3 * i] * ups[0] + V[3 * i + 1] * ups[1] + V[3 * i + 2] * ups[2];
    // T * E : t = t_T + R_T * t_E ; q = normalize(q_T * q_E)
    const Quat a = { T[0], T[1], T[2], T[3] };
    double rt[3];
    qrot(a, et, rt);
    Quat r;
    r.w = a.w * e.w - a.x * e.x - a.y * e.y - a.z * e.z;
    r.x = a.w * e.x + a.x * e.w + a.y * e.z - a.z * e.y;
    r.y = a.w * e.y + a.y * e.w + a.z * e.x - a.x * e.z;
    r.z = a.w * e.z + a.z * e.w + a.x * e.y - a.y * e.x;
    const double n = sqrt(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w);
    out[0] = r.w / n; out[1] = r.x / n; out[2] = r.y / n; out[3] = r.z / n;
    out[4] = T[4] + rt[0]; out[5] = T[5] + rt[1]; out[6] = T[6] + rt[2];
}

// Eigen LDLT<Matrix6d> (pivoted, lower) compute + solve. Hs = 21 packed lower-triangular entries (row-major: (i,j), j<=i).
__device__ void ldlt6_solve(const double* Hs, const double* bin, double x[6])
{
    double A[36];
    int tr[6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) { A[i * 6 + j] = Hs[i * (i + 1) / 2 + j]; A[j * 6 + i] = A[i * 6 + j]; }
#define LL(i, j) A[(i) * 6 + (j)]
    for (int k = 0; k < 6; ++k) {
        int big = k; double bigv = fabs(LL(k, k));
        for (int i = k + 1; i < 6; ++i) { const double v = fabs(LL(i, i)); if (v > bigv) { bigv = v; big = i; } }
        tr[k] = big;
        if (k != big) {
            for (int j = 0; j < k; ++j) { const double t = LL(k, j); LL(k, j) = LL(big, j); LL(big, j) = t; }
            for (int i = big + 1; i < 6; ++i) { const double t = LL(i, k); LL(i, k) = LL(i, big); LL(i, big) = t; }
            { const double t = LL(k, k); LL(k, k) = LL(big, big); LL(big, big) = t; }
            for (int i = k + 1; i < big; ++i) { const double t = LL(i, k); LL(i, k) = LL(big, i); LL(big, i) = t; }
        }
        if (k > 0) {
            double temp[6];
            for (int j = 0; j < k; ++j) temp[j] = LL(j, j) * LL(k, j);
            double s = 0; for (int j = 0; j < k; ++j) s += LL(k, j) * temp[j];
            LL(k, k) -= s;
            for (int i = k + 1; i < 6; ++i) {
                double s2 = 0; for (int j = 0; j < k; ++j) s2 += LL(i, j) * temp[j];
                LL(i, k) -= s2;
            }
        }
        const double akk = LL(k, k);
        const bool valid = fabs(akk) > 0.0;
        if (k == 0 && !valid) { for (int j = 0; j < 6; ++j) tr[j] = j; break; }
        if (k < 5 && valid) { const double inv = akk; for (int i = k + 1; i < 6; ++i) LL(i, k) /= inv; }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) y[i] = bin[i];
    for (int k = 0; k < 6; ++k) if (tr[k] != k) { const double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < i; ++j) y[i] -= LL(i, j) * y[j];
    const double tol = 1.0 / 1.7976931348623157e308;
    for (int i = 0; i < 6; ++i) { if (fabs(LL(i, i)) > tol) y[i] /= LL(i, i); else y[i] = 0; }
    for (int i = 5; i >= 0; --i) for (int j = i + 1; j < 6; ++j) y[i] -= LL(j, i) * y[j];
    for (int k = 5; k >= 0; --k) if (tr[k] != k) { const double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
    for (int i = 0; i < 6; ++i) x[i] = y[i];
#undef LL
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// non-contracted bilinear sample: ((w0*i0 + w1*i1) + w2*i2) + w3*i3   (ref: :147-148, :281)
__device__ __forceinline__ double bil(double w0, double w1, double w2, double w3, int i0, int i1, int i2, int i3)
{
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(w0, (double)i0), __dmul_rn(w1, (double)i1)), __dmul_rn(w2, (double)i2)),
                     __dmul_rn(w3, (double)i3));
}

constexpr int kMaxWarps = 16;
constexpr int kRedCols = 8;    // b[6], chi2, (pad)
constexpr int kHCols = 21;

__global__ void __launch_bounds__(512, 1) sparse_align_kernel(const SaArgs a)
{
    extern __shared__ __align__(16) double s_dyn[];
    const int NF = a.nf;
    double* s_ref = s_dyn;                 // [16][NF]
    double* s_dx = s_dyn + 16 * NF;        // [16][NF]
    double* s_dy = s_dyn + 32 * NF;        // [16][NF]
    __shared__ double s_red[kMaxWarps][kRedCols];
    __shared__ int s_cnt[kMaxWarps];
    __shared__ double s_redH[kMaxWarps][kHCols + 1];
    __shared__ double s_H[kHCols];
    __shared__ double s_T[7], s_Told[7];
    __shared__ double s_chi2prev;
    __shared__ int s_stop, s_npts, s_nlog;

    const int pair = blockIdx.x + a.pair0;
    const int j = threadIdx.x, lane = j & 31, warp = j >> 5, nwarps = blockDim.x >> 5;
    const int nfeat = min(a.n_feats[pair], NF);
    const uint8_t* __restrict__ ref_frame = a.frames + (size_t)a.ref_slots[pair] * a.frame_stride;
    const uint8_t* __restrict__ cur_frame = a.frames + (size_t)a.cur_slots[pair] * a.frame_stride;
    const double cen0 = a.centers[3 * pair], cen1 = a.centers[3 * pair + 1], cen2 = a.centers[3 * pair + 2];
    const double fx = (double)a.fx, fy = (double)a.fy, cx = (double)a.cx, cy = (double)a.cy;

    if (j < 7) { s_T[j] = a.poses_in[7 * pair + j]; s_Told[j] = s_T[j]; }
    if (j == 0) { s_npts = 0; s_nlog = 0; s_stop = 0; s_chi2prev = 0.0; }
    __syncthreads();

    for (int level = a.max_level - 1; level >= a.min_level; --level) {
        const int cols = a.geo.w[level], rows = a.geo.h[level];
        const float tScale = 1.0f / (float)(1 << level);
        const double scale = (double)tScale;
        const double fs = (double)a.f * scale;     // == (v * f) * scale bit-exactly, scale being a power of two

        // ------------------------------------------------ GetJocabianMat for feature j (ref: :62-166)
        bool valid = false;
        double P0 = 0, P1 = 0, P2 = 1;
        double Sxx = 0, Sxy = 0, Syy = 0;
        if (j < nfeat) {
            const dsdtm_ref_feat ft = a.feats[(size_t)pair * a.feat_stride + j];
            if (ft.initial) {
                const double px = (double)ft.px[0] * scale, py = (double)ft.px[1] * scale;          // ref: :89-91
                const bool zero = (ft.point_w[0] == 0.0 && ft.point_w[1] == 0.0 && ft.point_w[2] == 0.0);
                const int boarder = 3;                                                               // ref: :67
                if (!(zero || px - boarder < 0 || py - boarder < 0 || px + boarder >= cols || py + boarder >= rows)) {   // ref: :95-96
                    valid = true;
                    const double d0 = __dsub_rn(ft.point_w[0], cen0), d1 = __dsub_rn(ft.point_w[1], cen1), d2 = __dsub_rn(ft.point_w[2], cen2);
                    const double depth = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2)));   // ref: :117-118
                    P0 = __dmul_rn(ft.normal[0], depth); P1 = __dmul_rn(ft.normal[1], depth); P2 = __dmul_rn(ft.normal[2], depth);   // ref: :119
                    const int fxi = __double2int_rd(px), fyi = __double2int_rd(py);
                    const double sx = px - fxi, sy = py - fyi;
                    const double w00 = __dmul_rn(1.0 - sx, 1.0 - sy), w01 = __dmul_rn(sx, 1.0 - sy);
                    const double w10 = __dmul_rn(1.0 - sx, sy), w11 = __dmul_rn(sx, sy);                // ref: :129-132
                    // 7x7 neighbourhood (rows fyi-3..fyi+3, cols fxi-3..fxi+3) covers every tap of ref/dx/dy
                    const uint8_t* __restrict__ img = ref_frame + a.geo.off[level];
                    int I[7][7];
#pragma unroll
                    for (int r = 0; r < 7; ++r)
#pragma unroll
                        for (int c = 0; c < 7; ++c) I[r][c] = __ldg(img + (size_t)(fyi - 3 + r) * cols + (fxi - 3 + c));
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            // "it" of the reference = I[r+1][c+1]
                            const int R = r + 1, C = c + 1;
                            const double v = bil(w00, w01, w10, w11, I[R][C], I[R][C + 1], I[R + 1][C], I[R + 1][C + 1]);
                            const double dxp = bil(w00, w01, w10, w11, I[R][C + 1], I[R][C + 2], I[R + 1][C + 1], I[R + 1][C + 2]);
                            const double dxm = bil(w00, w01, w10, w11, I[R][C - 1], I[R][C], I[R + 1][C - 1], I[R + 1][C]);
                            const double dyp = bil(w00, w01, w10, w11, I[R + 1][C], I[R + 1][C + 1], I[R + 2][C], I[R + 2][C + 1]);
                            const double dym = bil(w00, w01, w10, w11, I[R - 1][C], I[R - 1][C + 1], I[R][C], I[R][C + 1]);
                            const double dx = __dmul_rn(0.5, __dsub_rn(dxp, dxm));     // ref: :150-153
                            const double dy = __dmul_rn(0.5, __dsub_rn(dyp, dym));     // ref: :155-158
                            const int p = 4 * r + c;
                            s_ref[p * NF + j] = v; s_dx[p * NF + j] = dx; s_dy[p * NF + j] = dy;
                            Sxx = fma(dx, dx, Sxx); Sxy = fma(dx, dy, Sxy); Syy = fma(dy, dy, Syy);
                        }
                }
            }
        }
        // GetJocabianBA(P) rows (ref: :169-193); a1 = b0 = 0
        const double zi = 1.0 / P2, zi2 = zi * zi;
        const double a0 = -zi, a2 = P0 * zi2, a3 = P1 * a2, a4 = -(1.0 + P0 * a2), a5 = P1 * zi;
        const double b1 = -zi, b2 = P1 * zi2, b3 = 1.0 + P1 * b2, b4 = -P0 * b2, b5 = -P0 * zi;
        const double fs2 = fs * fs;

        bool prev_vis = false;
        const uint8_t* __restrict__ cimg = cur_frame + a.geo.off[level];

        // ------------------------------------------------ GaussNewtonSolver (ref: :301-344)
        for (int it = 0; it < a.max_iters; ++it) {
            const Quat q = { s_T[0], s_T[1], s_T[2], s_T[3] };
            const double t0 = s_T[4], t1 = s_T[5], t2 = s_T[6];
            bool vis = false;
            double Sx = 0, Sy = 0, c2 = 0;
            if (valid) {
                const double P[3] = { P0, P1, P2 };
                double Q[3];
                qrot(q, P, Q);                                                             // ref: :254
                Q[0] = __dadd_rn(Q[0], t0); Q[1] = __dadd_rn(Q[1], t1); Q[2] = __dadd_rn(Q[2], t2);
                // Camera2Pixel * tScale (ref: src/Camera.cpp:167-171, :255): (fx*X)/Z + cx
                const double u = __dmul_rn(__dadd_rn(__ddiv_rn(__dmul_rn(fx, Q[0]), Q[2]), cx), scale);
                const double v = __dmul_rn(__dadd_rn(__ddiv_rn(__dmul_rn(fy, Q[1]), Q[2]), cy), scale);
                const double uf = floor(u), vf = floor(v);
                // ref: :262 with border 3; evaluated in double so that NaN / huge values are rejected like the reference's INT_MIN
                if (uf >= 3.0 && vf >= 3.0 && uf < (double)(cols - 3) && vf < (double)(rows - 3)) {
                    vis = true;
                    const int ui = (int)uf, vi = (int)vf;
                    const double su = u - uf, sv = v - vf;
                    const double tl = __dmul_rn(1.0 - su, 1.0 - sv), tr = __dmul_rn(su, 1.0 - sv);
                    const double bl = __dmul_rn(1.0 - su, sv), br = __dmul_rn(su, sv);        // ref: :267-270
                    // 5x5 window rows vi-2..vi+2, cols ui-2..ui+2 : two aligned 32-bit loads per row
                    int W[5][5];
                    const unsigned a0w = (unsigned)(vi - 2) * (unsigned)cols + (unsigned)(ui - 2);
#pragma unroll
                    for (int r = 0; r < 5; ++r) {
                        const unsigned ad = a0w + (unsigned)r * (unsigned)cols;
                        const uint32_t* wp = reinterpret_cast<const uint32_t*>(cimg + (ad & ~3u));
                        const uint32_t lo = __ldg(wp), hi = __ldg(wp + 1);
                        const unsigned long long bits = (((unsigned long long)hi << 32) | lo) >> (8 * (ad & 3u));
#pragma unroll
                        for (int c = 0; c < 5; ++c) W[r][c] = (int)((bits >> (8 * c)) & 0xFFull);
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int p = 4 * r + c;
                            const double cur = bil(tl, tr, bl, br, W[r][c], W[r][c + 1], W[r + 1][c], W[r + 1][c + 1]);   // ref: :281
                            const double res = __dsub_rn(cur, s_ref[p * NF + j]);                                    // ref: :282
                            c2 = __dadd_rn(c2, __dmul_rn(res, res));                                                 // ref: :284
                            Sx = fma(s_dx[p * NF + j], res, Sx);
                            Sy = fma(s_dy[p * NF + j], res, Sy);
                        }
                }
            }
            // b_j = fs * (a * Sx + b * Sy)
            double v0 = fs * (a0 * Sx), v1 = fs * (b1 * Sy), v2 = fs * (a2 * Sx + b2 * Sy), v3 = fs * (a3 * Sx + b3 * Sy);
            double v4 = fs * (a4 * Sx + b4 * Sy), v5 = fs * (a5 * Sx + b5 * Sy);
            v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3); v4 = warp_sum(v4); v5 = warp_sum(v5);
            c2 = warp_sum(c2);
            const int cnt = __reduce_add_sync(0xffffffffu, vis ? 1 : 0);
            if (lane == 0) {
                s_red[warp][0] = v0; s_red[warp][1] = v1; s_red[warp][2] = v2; s_red[warp][3] = v3;
                s_red[warp][4] = v4; s_red[warp][5] = v5; s_red[warp][6] = c2;
                s_cnt[warp] = cnt;
            }
            const int need_H = __syncthreads_or((it == 0) || (vis != prev_vis));
            prev_vis = vis;
            if (need_H) {
                // H_j = fs^2 (Sxx a a^T + Sxy (a b^T + b a^T) + Syy b b^T), lower triangle packed row-major
                const double av[6] = { a0, 0.0, a2, a3, a4, a5 };
                const double bv[6] = { 0.0, b1, b2, b3, b4, b5 };
                const double m = vis ? fs2 : 0.0;
                const double cxx = m * Sxx, cxy = m * Sxy, cyy = m * Syy;
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int c = 0; c <= r; ++c) {
                        double h = cxx * (av[r] * av[c]) + cxy * (av[r] * bv[c] + bv[r] * av[c]) + cyy * (bv[r] * bv[c]);
                        h = warp_sum(h);
                        if (lane == 0) s_redH[warp][r * (r + 1) / 2 + c] = h;
                    }
                __syncthreads();
            }
            if (warp == 0) {
                if (need_H && lane < kHCols) {
                    double h = 0;
                    for (int w = 0; w < nwarps; ++w) h += s_redH[w][lane];
                    s_H[lane] = h;
                }
                double red = 0;
                if (lane < 7) for (int w = 0; w < nwarps; ++w) red += s_red[w][lane];
                int npts = 0;
                for (int w = 0; w < nwarps; ++w) npts += s_cnt[w];
                double bvec[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) bvec[k] = __shfl_sync(0xffffffffu, red, k);
                const double chi2sum = __shfl_sync(0xffffffffu, red, 6);
                __syncwarp();
                if (lane == 0) {
                    const double chi2New = chi2sum / (double)(16 * npts);                  // ref: :298 (NaN if nothing visible)
                    double x[6];
                    ldlt6_solve(s_H, bvec, x);                                             // ref: :318
                    int flags = 0;
                    bool stop = false;
                    if (isnan(x[0])) { stop = true; flags |= 4; }                          // ref: :321-326
                    if ((it > 0 && chi2New > s_chi2prev) || stop) {                        // ref: :328-332
                        for (int k = 0; k < 7; ++k) s_T[k] = s_Told[k];
                        flags |= 2;
                        stop = true;
                    } else {
                        double Tn[7], Tc[7];
                        for (int k = 0; k < 7; ++k) Tc[k] = s_T[k];
                        se3_mul_exp(Tc, x, Tn);                                            // ref: :335
                        for (int k = 0; k < 7; ++k) { s_Told[k] = Tc[k]; s_T[k] = Tn[k]; } // ref: :336-337
                        s_chi2prev = chi2New;                                              // ref: :339
                        flags |= 1;
                        double mx = 0;
                        for (int k = 0; k < 6; ++k) mx = fmax(mx, fabs(x[k]));
                        if (mx <= 1e-8) { stop = true; flags |= 8; }                       // ref: :341
                    }
                    s_npts = npts;
                    s_stop = stop ? 1 : 0;
                    if (a.log) {
                        const int n = s_nlog;
                        if (n < a.log_cap) {
                            dsdtm_iter_log* e = a.log + (size_t)pair * a.log_cap + n;
                            e->level = level; e->iter = it; e->n_pts = npts; e->flags = flags; e->chi2 = chi2New;
                            for (int k = 0; k < 6; ++k) e->x[k] = x[k];
                        }
                        s_nlog = n + 1;
                    }
                }
            }
            __syncthreads();
            if (s_stop) break;
        }
        // ref: :308 tT_c2rOld(tT_c2r) and chi2 = 0 at the start of every level
        __syncthreads();
        if (j < 7) s_Told[j] = s_T[j];
        if (j == 0) { s_stop = 0; s_chi2prev = 0.0; }
        __syncthreads();
    }
    if (j < 7) a.poses_out[7 * pair + j] = s_T[j];
    if (j == 0) {
        a.n_tracked[pair] = s_npts;
        if (a.n_log) a.n_log[pair] = s_nlog;
    }
}

}  // namespace

int sparse_align_smem_bytes(int nf) { return 48 * nf * (int)sizeof(double); }

cudaError_t sparse_align_init(dsdtm_ctx* c)
{
    const int nf = (c->prm.max_feats + 31) / 32 * 32;
    return cudaFuncSetAttribute(sparse_align_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sparse_align_smem_bytes(nf));
}

cudaError_t launch_sparse_align(dsdtm_ctx* c, int n_pairs, int feat_stride, int max_level, int min_level, int max_iters,
                                bool want_log, cudaStream_t s, int pair0)
{
    SaArgs a;
    a.frames = c->frames_d; a.frame_stride = c->geo.frame_stride; a.geo = c->geo;
    a.ref_slots = c->ref_slots_d; a.cur_slots = c->cur_slots_d;
    a.feats = c->feats_d; a.feat_stride = feat_stride; a.n_feats = c->n_feats_d;
    a.centers = c->centers_d; a.poses_in = c->poses_in_d; a.poses_out = c->poses_out_d; a.n_tracked = c->n_tracked_d;
    a.log = want_log ? c->log_d : nullptr; a.n_log = want_log ? c->n_log_d : nullptr; a.log_cap = kLogCap;
    a.fx = c->cam.fx; a.fy = c->cam.fy; a.cx = c->cam.cx; a.cy = c->cam.cy; a.f = c->cam.f;
    a.max_level = max_level; a.min_level = min_level; a.max_iters = max_iters;
    a.nf = (c->prm.max_feats + 31) / 32 * 32;
    a.pair0 = pair0;
    sparse_align_kernel<<<n_pairs, a.nf, sparse_align_smem_bytes(a.nf), s>>>(a);
    c->launches++;
    return cudaGetLastError();
}

}  // namespace dsdtm
