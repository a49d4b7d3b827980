// sparse_align.cu -- kernel (c): coarse-to-fine inverse-compositional Gauss-Newton on SE3, one CTA per frame pair.
//
// Replaces Sprase_ImgAlign::GetJocabianMat / ComputeResiduals / GaussNewtonSolver and the level loop of Run
// (ref: src/Sprase_ImageAlign.cpp:43-55, 62-193, 240-344). fp64 throughout, like the reference.
//
// Design (DESIGN.md "Sparse alignment"; profiles/r1_sparse_align_v1.md explains why v1 was replaced):
//   * one CTA of WPP warps per frame pair, each lane owns ceil(N / (32*WPP)) reference features; the whole
//     level x iteration loop runs on the device (the chain of dependent GN steps never returns to the host).
//     WPP = 1 packs 6 independent pairs on one SM (throughput: other pairs fill the serial solve of this one),
//     WPP = 10 gives one feature per lane (single-pair latency);
//   * per level the 7x7 u8 neighbourhood of every reference feature (49 B) is staged ONCE in shared memory next to its 3-D point, 1 / z,
//     level-0 pixel and parked second moments (121 B / feature, instead of the 384 B of precomputed fp64 patches which limited v1 to one
//     pair per SM); the level-independent part runs once per pair in a prologue, the staging keeps two features' 42 window loads in flight
//     per thread; the bilinear reference samples and their central differences are re-derived from the bytes with the reference's
//     expressions each iteration (2x the flops, 1/3 the shared memory, 6x the residency);
//   * bytes enter the fp64 arithmetic WITHOUT a conversion instruction: as subnormals b * 2^-1034 built by one PRMT, with bilinear weights
//     that carry 2^1010 and sums restored by exact powers of two (DSDTM_SA_CVT 3 below): the reference's roundings, bit for bit;
//   * inverse-compositional structure: the Jacobian row of pixel p of feature j is
//         J_jp = (dx_jp * a_j + dy_jp * b_j) * (f * scale)          (ref: :160; a_j, b_j = rows of GetJocabianBA(P_j))
//     so  sum_p J_jp J_jp^T = (f*scale)^2 (Sxx a a^T + Sxy (a b^T + b a^T) + Syy b b^T) is pose-independent: H is
//     re-reduced over the VISIBLE features only when the visibility set changes (always at iteration 0 of a level),
//     and the per-iteration work is  b += (f*scale)(a_j sum_p dx r + b_j sum_p dy r),  chi2 += sum_p r^2.
//     Same mathematics, different summation order than the reference's sequential accumulation
//     (documented tolerance: chi2 1e-4 relative, pose 1e-5 rad / m);
//   * reductions: per-lane accumulation over its features, then a transposed warp reduction (lane l ends with total l; the pairing
//     tree of the xor butterfly, so the same bits), WPP > 1 adds one shared-memory row per warp summed in warp order => deterministic;
//   * lane 0 solves the 6x6 system with Eigen's pivoted LDL^T entirely in registers (se3_ldlt.cuh), applies SE3::exp and
//     the reference's accept / revert / converge rules, and publishes the pose through shared memory.
#include <algorithm>
#include <mutex>
#include "ctx.cuh"
#include <type_traits>

#include "se3_ldlt.cuh"

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
    double* ws;  // variant 1: per-pair workspace [48][nf] doubles (ref, 2dx, 2dy per patch pixel), L2-resident
    int nf;      // shared-memory column count (multiple of 16, >= every n_feats)
    int pair0;
    int variant;   // 0 recompute, 1 L2 workspace
    double* mom;   // DSDTM_SA_SGLOBAL: per-pair [3][nf] second moments (the head of the pair's workspace block)
};

constexpr int NB_WORDS = 14;   // 7 rows x 2 words (7 bytes) of the reference neighbourhood

__device__ __forceinline__ double u8_to_f64(uint32_t word, int byte)
{
#ifndef DSDTM_SA_CVT
#define DSDTM_SA_CVT 3      // measured (ms per 2072 pairs, round 1): 0 magic-number add 0.847, 1 PRMT + I2F.F64 0.816, 2 shift + I2F.F64 0.828; all exact.
                            // 3 (round 2, main kernel): subnormal encoding, no conversion instruction -- 1.295 -> 1.267 ms per 4096 pairs, bit-equal
#endif
#if DSDTM_SA_CVT == 0
    // exact int -> double without the slow I2F.F64 path: 2^52 + b has b in its low mantissa bits
    return __hiloint2double(0x43300000, (int)__byte_perm(word, 0, 0x4440 | byte)) - 4503599627370496.0;
#elif DSDTM_SA_CVT == 1 || DSDTM_SA_CVT == 3
    return (double)__byte_perm(word, 0, 0x4440 | byte);                 // PRMT + I2F.F64.U32
#else
    return (double)(unsigned char)(word >> (8 * byte));                 // lets ptxas pick I2F.F64.U8 with a byte selector
#endif
}

// DSDTM_SA_CVT == 3 (main kernel only): NO conversion instruction at all. One PRMT puts the byte into bits 8..15 of the HIGH word of
// a double whose exponent field is 0: the subnormal b * 2^-1034, exactly proportional to the byte (0 -> +0.0). The bilinear weights
// carry 2^SA_WEXP, so every sample is the reference's value times 2^(SA_WEXP - 1034) = 2^-24 with the SAME rounding (scaling by a power
// of two commutes with IEEE rounding while nothing under- or overflows: products are >= 2^-46 * 2^-24, far from 2^-1022), and so are
// residuals and central differences; sums of products carry 2^-48 and are restored by one exact multiplication each. Bit-identical
// results, 74 I2F.F64 (XU pipe, quarter rate) less per feature and iteration.
#if DSDTM_SA_CVT == 3
#define SA_WSCALE 0x1p1010
#define SA_RESTORE2 0x1p48          // 1 / (2^-24)^2
__device__ __forceinline__ double u8_to_f64s(uint32_t word, int byte)
{
    return __hiloint2double((int)__byte_perm(word, 0, 0x4404 | (byte << 4)), 0);
}
#else
#define SA_WSCALE 1.0
#define SA_RESTORE2 1.0
__device__ __forceinline__ double u8_to_f64s(uint32_t word, int byte) { return u8_to_f64(word, byte); }
#endif

__device__ __forceinline__ void qrot(double qw, double qx, double qy, double qz, double v0, double v1, double v2,
                                     double& o0, double& o1, double& o2)   // Eigen _transformVector, non-contracted
{
    double uv0 = __dsub_rn(__dmul_rn(qy, v2), __dmul_rn(qz, v1));
    double uv1 = __dsub_rn(__dmul_rn(qz, v0), __dmul_rn(qx, v2));
    double uv2 = __dsub_rn(__dmul_rn(qx, v1), __dmul_rn(qy, v0));
    uv0 = __dadd_rn(uv0, uv0); uv1 = __dadd_rn(uv1, uv1); uv2 = __dadd_rn(uv2, uv2);
    const double c0 = __dsub_rn(__dmul_rn(qy, uv2), __dmul_rn(qz, uv1));
    const double c1 = __dsub_rn(__dmul_rn(qz, uv0), __dmul_rn(qx, uv2));
    const double c2 = __dsub_rn(__dmul_rn(qx, uv1), __dmul_rn(qy, uv0));
    o0 = __dadd_rn(__dadd_rn(v0, __dmul_rn(qw, uv0)), c0);
    o1 = __dadd_rn(__dadd_rn(v1, __dmul_rn(qw, uv1)), c1);
    o2 = __dadd_rn(__dadd_rn(v2, __dmul_rn(qw, uv2)), c2);
}

// bilinear sample ((w0*i0 + w1*i1) + w2*i2) + w3*i3 in the reference's left-to-right order (ref: :147-148, :281).
// Default: the three additions are fused (1 DMUL + 3 DFMA instead of 4 DMUL + 3 DADD: the fp64 pipe is the binding unit,
// profiles/r1_sparse_align_v2.md); each sample then differs from the reference's FMA-free x86 value by <= 1 ulp (1e-16
// relative), far inside the 1e-4 chi2 / 1e-5 pose tolerances. -DDSDTM_SA_STRICT=1 restores the non-contracted form.
#ifndef DSDTM_SA_STRICT
#define DSDTM_SA_STRICT 0
#endif
#ifndef DSDTM_SA_MOM
#define DSDTM_SA_MOM 1      // second moments (needed by the first iteration of a level only): 0 `if (first)` per pixel = 6 FSEL + 3 DFMA per
                            // pixel on EVERY iteration (8.5 % of all executed instructions were FSEL), 1 always accumulate, 2 uniform
                            // branch per patch row. ms per 4096 pairs: 1.346 / 1.295 / 1.305 (profiles/r2_sparse_align.md)
#endif
__device__ __forceinline__ double bil(double w0, double w1, double w2, double w3, double i0, double i1, double i2, double i3)
{
#if DSDTM_SA_STRICT
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(w0, i0), __dmul_rn(w1, i1)), __dmul_rn(w2, i2)), __dmul_rn(w3, i3));
#else
    return fma(w3, i3, fma(w2, i2, fma(w1, i1, __dmul_rn(w0, i0))));
#endif
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#ifndef DSDTM_SA_TREDUCE
#define DSDTM_SA_TREDUCE 1       // 1 = the 7 per-iteration sums and the 21 entries of H meet in a transposed warp reduction (lane l ends with total l:
                                 // 21 resp. 31 exchanges instead of 35 resp. 105) and lanes write their own entry of the shared-memory row
#endif
// Sum N <= 32 per-lane values over the warp so that lane l ends up with the total of value l (the reduction of pose_opt.cuh): at the stage
// with offset o a lane keeps the half of its values whose index has bit o equal to its own lane bit and hands the other half to its partner.
// Deterministic; only lanes < N hold a defined result.
template <int N>
__device__ __forceinline__ double warp_sum_transposed(const double (&v)[N])
{
    static_assert(N <= 32, "one value per lane");
    const int lane = threadIdx.x & 31;
    double w[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) w[i] = i < N ? v[i] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            if (i >= N) continue;
            if (i + o >= N) {
                // the partner column is a zero column: the plain butterfly gives the lanes that keep column i the same sum without the two
                // selects (the other lanes then carry a value nobody reads: later stages only pair lanes with equal upper bits)
                w[i] += __shfl_xor_sync(0xffffffffu, w[i], o);
                continue;
            }
            const double lo = w[i], hi = w[i + o];
            const double keep = up ? hi : lo;
            const double send = up ? lo : hi;
            w[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return w[0];
}

// Reference-side samples of one feature, re-derived from its staged 7x7 neighbourhood. Row r of the 4x4 patch:
//   ref[c] = G[r+1][c+1], dx[c] = 0.5 (G[r+1][c+2] - G[r+1][c]), dy[c] = 0.5 (G[r+2][c+1] - G[r][c+1])
// with G[y][x] = bil(w; N[y][x], N[y][x+1], N[y+1][x], N[y+1][x+1]) -- the reference's expressions (ref: :147-158).
template <bool SCALED>
struct RefRowsT {
    double w00, w01, w10, w11;
    double Nd[2][7];      // two converted neighbourhood rows (y, y+1)
    double G[3][6];       // G rows y-1, y, y+1
    const uint32_t* nb;   // &s_nb[0 * NF + f]
    int NF;

    __device__ __forceinline__ void load_row(int slot, int y)
    {
        const uint32_t lo = nb[(2 * y) * NF], hi = nb[(2 * y + 1) * NF];
#pragma unroll
        for (int c = 0; c < 4; ++c) Nd[slot][c] = SCALED ? u8_to_f64s(lo, c) : u8_to_f64(lo, c);
#pragma unroll
        for (int c = 0; c < 3; ++c) Nd[slot][4 + c] = SCALED ? u8_to_f64s(hi, c) : u8_to_f64(hi, c);
    }
    __device__ __forceinline__ void load_row_words(int slot, uint32_t lo, uint32_t hi)
    {
#pragma unroll
        for (int c = 0; c < 4; ++c) Nd[slot][c] = SCALED ? u8_to_f64s(lo, c) : u8_to_f64(lo, c);
#pragma unroll
        for (int c = 0; c < 3; ++c) Nd[slot][4 + c] = SCALED ? u8_to_f64s(hi, c) : u8_to_f64(hi, c);
    }
    __device__ __forceinline__ void grid_row(int g, int a, int b)   // G[g][x] from Nd[a] (row y) and Nd[b] (row y+1)
    {
#pragma unroll
        for (int x = 0; x < 6; ++x) G[g][x] = bil(w00, w01, w10, w11, Nd[a][x], Nd[a][x + 1], Nd[b][x], Nd[b][x + 1]);
    }
};
using RefRows = RefRowsT<false>;

#ifndef DSDTM_SA_TAIL_LANES
#define DSDTM_SA_TAIL_LANES 0    // 1 = the solve + pose update run redundantly on ALL lanes of warp 0 so that the four power series of SE3::exp are one
                                 // Horner chain on four lanes (7 instead of 28 DFMA, no 64-bit immediates) and the fresh factor is used from registers.
                                 // Bit-equal, but SLOWER (1.205 vs 1.194 ms): an FP64 instruction of a warp with one active lane occupies the pipe for
                                 // one 16-lane pass, with 32 active lanes for two -- the one-lane tail is the cheaper form. 0 = everything on lane 0
#endif
// the series coefficients of se3_mul_exp (se3_ldlt.cuh), [k][j] = coefficient k (highest power first) of series j
__constant__ double kSe3SeriesCoef[32] = { DSDTM_SE3_SERIES_COEF_LIST };

#if DSDTM_SA_TAIL_LANES
// all 32 lanes of the calling warp, uniform x: T * exp(x) with the four series evaluated by lanes 0..3 (every lane runs the chain of its
// lane & 3) -- the same Horner steps as se3_mul_exp's own, so the same bits
__device__ __forceinline__ void pose_update_lanes(const double* T, const double (&x)[6], double* Tn, const double* s_coef, int lane)
{
    double Tc[7], To[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) Tc[q] = T[q];
    const double t2 = se3_theta2(x);
    double s4[4] = { 0.0, 0.0, 0.0, 0.0 };
    if (se3_exp_uses_series(t2)) {
        const int j = lane & 3;
        const double arg = (j < 2) ? 0.25 * t2 : t2;
        double r = s_coef[j];
#pragma unroll
        for (int k = 1; k < 8; ++k) r = fma(arg, r, s_coef[4 * k + j]);
#pragma unroll
        for (int q = 0; q < 4; ++q) s4[q] = __shfl_sync(0xffffffffu, r, q);
    }
    se3_mul_exp(Tc, x, To, s4);
#pragma unroll
    for (int q = 0; q < 7; ++q) Tn[q] = To[q];
}

#endif
#ifndef DSDTM_SA_TAIL_CONST
#define DSDTM_SA_TAIL_CONST 1    // (1.193 -> 1.187 ms, bit-equal) 1 = lane 0 evaluates the four exp series with coefficients read from the constant bank (operands of the DFMA) instead of
                                 // 64-bit immediates moved through uniform registers (two UMOV per coefficient), and uses a fresh factor from registers
#endif
#if DSDTM_SA_TAIL_CONST
// lane 0 (or any single lane): T * exp(x), the four series from the constant table -- the same Horner steps and doubles as se3_mul_exp's own
__device__ __forceinline__ void pose_update_const(const double* T, const double (&x)[6], double* Tn)
{
    double Tc[7], To[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) Tc[q] = T[q];
    const double t2 = se3_theta2(x);
    double s4[4] = { 0.0, 0.0, 0.0, 0.0 };
    if (se3_exp_uses_series(t2)) {
        const double h2 = 0.25 * t2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double arg = (j < 2) ? h2 : t2;
            double r = kSe3SeriesCoef[j];
#pragma unroll
            for (int k = 1; k < 8; ++k) r = fma(arg, r, kSe3SeriesCoef[4 * k + j]);
            s4[j] = r;
        }
    }
    se3_mul_exp(Tc, x, To, s4);
#pragma unroll
    for (int q = 0; q < 7; ++q) Tn[q] = To[q];
}

#endif
// the solve on every lane of the warp (uniform inputs): a fresh factorisation is used from registers and parked by lane 0
__device__ __forceinline__ void solve_all_lanes(const double* __restrict__ sH /*21 packed*/, double* __restrict__ sF /*22*/, bool refactor,
                                                const double (&bvec)[6], double (&x)[6], int lane)
{
    double Lp[15], d[6];
    bool ok;
    if (refactor) {
        double Hm[6][6];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) { Hm[r][c] = sH[r * (r + 1) / 2 + c]; Hm[c][r] = Hm[r][c]; }
        ok = ldlt6_factor_spd(Hm, Lp, d);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 15; ++i) sF[i] = Lp[i];
#pragma unroll
            for (int i = 0; i < 6; ++i) sF[15 + i] = d[i];
            sF[21] = ok ? 1.0 : 0.0;
        }
    } else {
        ok = sF[21] != 0.0;
#pragma unroll
        for (int i = 0; i < 15; ++i) Lp[i] = sF[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) d[i] = sF[15 + i];
    }
    if (ok) {
        ldlt6_subst_spd(Lp, d, bvec, x);                                   // ref: :318 (Eigen ldlt().solve), SPD fast path
    } else {
        double Hm[6][6];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) { Hm[r][c] = sH[r * (r + 1) / 2 + c]; Hm[c][r] = Hm[r][c]; }
        ldlt6_solve_reg(Hm, bvec, x);                                      // Eigen's pivoted LDL^T incl. its zero-pivot rule
    }
}

// the serial tail of one GN iteration, kept out of line: it runs on one lane once per iteration and must not bloat
// (or evict from the instruction cache) the per-feature loop. H only changes when the visibility set changes, so the
// factorisation is cached in shared memory (sF: 15 entries of L, 6 pivots, 1 flag) and most iterations only substitute.
#ifndef DSDTM_SA_TAIL_INLINE
#define DSDTM_SA_TAIL_INLINE 1
#endif
#if DSDTM_SA_TAIL_INLINE
#define DSDTM_TAIL_ATTR __forceinline__
#else
#define DSDTM_TAIL_ATTR __noinline__
#endif
__device__ DSDTM_TAIL_ATTR void solve_and_update(const double* __restrict__ sH /*21 packed*/, double* __restrict__ sF /*22*/, bool refactor,
                                              const double (&bvec)[6], double (&x)[6])
{
    if (refactor) {
        double Hm[6][6], Lp[15], d[6];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) { Hm[r][c] = sH[r * (r + 1) / 2 + c]; Hm[c][r] = Hm[r][c]; }
        const bool ok = ldlt6_factor_spd(Hm, Lp, d);
#pragma unroll
        for (int i = 0; i < 15; ++i) sF[i] = Lp[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) sF[15 + i] = d[i];
        sF[21] = ok ? 1.0 : 0.0;
    }
    if (sF[21] != 0.0) {
        double Lp[15], d[6];
#pragma unroll
        for (int i = 0; i < 15; ++i) Lp[i] = sF[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) d[i] = sF[15 + i];
        ldlt6_subst_spd(Lp, d, bvec, x);                                   // ref: :318 (Eigen ldlt().solve), SPD fast path
    } else {
        double Hm[6][6];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) { Hm[r][c] = sH[r * (r + 1) / 2 + c]; Hm[c][r] = Hm[r][c]; }
        ldlt6_solve_reg(Hm, bvec, x);                                      // Eigen's pivoted LDL^T incl. its zero-pivot rule
    }
}

__device__ DSDTM_TAIL_ATTR void pose_update(const double* T, const double (&x)[6], double* Tn)
{
    double Tc[7], To[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) Tc[q] = T[q];
    se3_mul_exp(Tc, x, To);                                                // ref: :335
#pragma unroll
    for (int q = 0; q < 7; ++q) Tn[q] = To[q];
}

#ifndef DSDTM_SA_PIPELINE
#define DSDTM_SA_PIPELINE 0      // 1 = issue feature k+1's projection + gather before feature k's arithmetic. Round 1: slower (registers). Round 2 again, now
                                 // without spills (158 registers): 1.274 vs 1.220 ms -- the pass does not wait for memory, hiding the gather buys nothing
#endif
#if DSDTM_SA_PIPELINE
struct Pre { bool valid, vis; double su, sv; uint32_t cw0[5], cw1[5]; };     // carried across a feature's arithmetic: fractions instead of the four weights
#else
struct Pre { bool valid, vis; double tl, tr, bl, br; uint32_t cw0[5], cw1[5]; };
#endif

// (Tried and removed: staging an 8 x 7 window of the CURRENT image per feature in shared memory at level start, so that the
// iterations of a level stop re-gathering it from global memory. 173 instead of 113 B/feature of shared memory took more L1
// away than the re-gathers cost: 1.68 vs 1.55 ms per 4096 pairs, 86 vs 84 us for a single pair. profiles/r1_sparse_align_v3.md)
#ifndef DSDTM_SA_STAGE
#define DSDTM_SA_STAGE 1         // 1 = level staging with two features per thread in flight + level-independent prologue (1.295 -> 1.242 ms per 4096 pairs; with CVT 3: 1.216); 0 = one feature at a time
#endif
#ifndef DSDTM_SA_STAGE_G
#define DSDTM_SA_STAGE_G 2       // features per thread whose window loads are in flight together in the level staging (4: 1.227 vs 1.216 ms)
#endif
#ifndef DSDTM_SA_SGLOBAL
#define DSDTM_SA_SGLOBAL 0       // 1 = the parked second moments live in a per-pair global workspace (L2) instead of shared memory: 97 instead of 121 B / feature,
                                 // 96 instead of 64 KB of L1 -- measured slower (1.247 vs 1.216 ms)
#endif
#ifndef DSDTM_SA_COMPACT
#define DSDTM_SA_COMPACT 0       // 1 = the features staged at a level are compacted into an index list: the pass runs ceil(valid / lanes) rounds instead of
                                 // ceil(n / lanes). Measured: no gain on the sweep batch (1.215 vs 1.215 ms: 295-297 of its 300 features are staged at
                                 // every level, the fourth round stays) and a different (equally valid) summation order; off
#endif
#ifndef DSDTM_SA_RMAT
#define DSDTM_SA_RMAT 0          // 1 = the point is rotated with the 3x3 matrix of the pose's quaternion (9 FMA) instead of Eigen's quaternion formula (30 operations):
                                 // 1.210 vs 1.215 ms, results move in the last bits; not worth leaving the reference's expression
#endif
#ifndef DSDTM_SA_LD64
#define DSDTM_SA_LD64 0          // bit 0: level staging, bit 1: feature pass gather with 8-byte aligned 64-bit loads (the second load predicated): fewer L1 requests,
                                 // measured 1.224 (staging) / 1.285 (pass) vs 1.219 ms -- L1 request slots are not what the gathers wait for; off
#endif
#ifndef DSDTM_SA_PRO_G
#define DSDTM_SA_PRO_G 2         // feature records per thread in flight in the prologue (4: 1.185 vs 1.187 ms, within noise)
#endif
#ifndef DSDTM_SA_PREF_NEXT
#define DSDTM_SA_PREF_NEXT 0     // 1 = the staging of a level prefetches the next level's reference rows into L2: 1.260 vs 1.205 ms (a scattered prefetch costs a
                                 // full L1 request per lane like a load); off
#endif
#ifndef DSDTM_SA_PREF
#define DSDTM_SA_PREF 0          // 1 = the level staging prefetches the current-image rows of the first iteration (needs DSDTM_SA_STAGE 1): measured SLOWER, 1.292 vs 1.216 ms
#endif
#ifndef DSDTM_SA_MINB3
#define DSDTM_SA_MINB3 4         // resident CTAs per SM the 3-warp variant is compiled for (5 -> 128 regs, 60 B spills: 1.79 vs 1.52 ms; 6 -> 96 regs: 2.89 ms)
#endif
#ifndef DSDTM_SA_MINB5
#define DSDTM_SA_MINB5 3         // resident CTAs per SM the 5-warp variant is compiled for (3 -> 128 registers, 56 B of spills; 2 -> 168, none)
#endif
#ifndef DSDTM_SA_MINB4
#define DSDTM_SA_MINB4 3         // resident CTAs per SM the 4-warp variant is compiled for (register cap 65536 / (128 * MINB4))
#endif

template <int WPP>
__global__ void __launch_bounds__(32 * WPP, (WPP <= 2) ? 6 : (WPP == 3 ? DSDTM_SA_MINB3 : (WPP == 4 ? DSDTM_SA_MINB4 : (WPP == 5 ? DSDTM_SA_MINB5 : (WPP == 6 ? 2 : 1))))) sparse_align_kernel(const SaArgs a)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int NF = a.nf;
    double* s_P = reinterpret_cast<double*>(s_raw);                                        // [3][NF] point in the ref camera
#if DSDTM_SA_SGLOBAL
    constexpr int SB = 32;                                                                 // bytes per feature in front of the neighbourhood
    double* s_S = a.mom + (size_t)(blockIdx.x + a.pair0) * 48 * NF;                        // [3][NF] Sxx, Sxy, Syy: written and read by the same thread, L2-resident
    double* s_zi = reinterpret_cast<double*>(s_raw + (size_t)24 * NF);                     // [NF] 1 / P.z
#else
    constexpr int SB = 56;
    double* s_S = reinterpret_cast<double*>(s_raw + (size_t)24 * NF);                      // [3][NF] Sxx, Sxy, Syy of the ref patch
    double* s_zi = reinterpret_cast<double*>(s_raw + (size_t)48 * NF);                     // [NF] 1 / P.z (pose-independent: one division per level instead of one per iteration)
#endif
    uint32_t* s_nb = reinterpret_cast<uint32_t*>(s_raw + (size_t)SB * NF);                 // [14][NF] 7x7 u8 neighbourhood
    float2* s_sub = reinterpret_cast<float2*>(s_raw + (size_t)(SB + 4 * NB_WORDS) * NF);   // [NF] sub-pixel offsets (DSDTM_SA_STAGE 0) ...
    float2* s_px = s_sub;                                                                  // ... or the feature's level-0 pixel (DSDTM_SA_STAGE 1)
    uint8_t* s_valid = s_raw + (size_t)(SB + 4 * NB_WORDS + 8) * NF;                       // [NF] bit 0: staged at this level, bit 1: has a map point (all levels)
#if DSDTM_SA_COMPACT
    uint16_t* s_idx = reinterpret_cast<uint16_t*>(s_raw + (size_t)(SB + 4 * NB_WORDS + 8 + 1) * NF);   // [NF] the features staged at this level, ascending
    __shared__ int s_nvalid;
#endif
    __shared__ double s_red[WPP][8];
    __shared__ int s_cnt[WPP];
    __shared__ double s_redH[WPP][22];
    __shared__ double s_H[21];
    __shared__ double s_F[22];
    __shared__ double s_T[7], s_Told[7];
    __shared__ double s_chi2prev;
    __shared__ int s_stop, s_npts, s_nlog;
#if DSDTM_SA_TAIL_LANES
    __shared__ double s_coef[32];       // kSe3SeriesCoef: read with a lane-dependent index (a constant-bank read would serialise)
#endif

    constexpr int NT = 32 * WPP;
    const int pair = blockIdx.x + a.pair0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#if DSDTM_SA_TAIL_LANES
    if (tid < 32) s_coef[tid] = kSe3SeriesCoef[tid];
#endif
    const int nfeat = min(a.n_feats[pair], NF);
    const uint8_t* __restrict__ ref_frame = a.frames + (size_t)a.ref_slots[pair] * a.frame_stride;
    const uint8_t* __restrict__ cur_frame = a.frames + (size_t)a.cur_slots[pair] * a.frame_stride;
    const double cen0 = a.centers[3 * pair], cen1 = a.centers[3 * pair + 1], cen2 = a.centers[3 * pair + 2];
    const double fx = (double)a.fx, fy = (double)a.fy, cx = (double)a.cx, cy = (double)a.cy;

#ifdef DSDTM_SA_TIMING
    const long long tkp0 = clock64();
#endif
    if (tid < 7) { s_T[tid] = a.poses_in[7 * pair + tid]; s_Told[tid] = s_T[tid]; }
    if (tid == 0) { s_npts = 0; s_nlog = 0; s_stop = 0; s_chi2prev = 0.0; }
#if DSDTM_SA_STAGE == 1
    // prologue: what GetJocabianMat derives per feature that does not depend on the level (ref: :86, :95 zero test, :117-119), once;
    // DSDTM_SA_PRO_G 64-byte records per thread in flight
    for (int f0 = tid; f0 < nfeat; f0 += DSDTM_SA_PRO_G * NT) {
        dsdtm_ref_feat ft[DSDTM_SA_PRO_G];
#pragma unroll
        for (int g = 0; g < DSDTM_SA_PRO_G; ++g) ft[g] = a.feats[(size_t)pair * a.feat_stride + min(f0 + g * NT, nfeat - 1)];
#pragma unroll
        for (int g = 0; g < DSDTM_SA_PRO_G; ++g) {
            const int f = f0 + g * NT;
            if (f < nfeat) {
                const bool zero = (ft[g].point_w[0] == 0.0 && ft[g].point_w[1] == 0.0 && ft[g].point_w[2] == 0.0);
                const bool eligible = ft[g].initial && !zero;
                if (eligible) {
                    const double d0 = __dsub_rn(ft[g].point_w[0], cen0), d1 = __dsub_rn(ft[g].point_w[1], cen1), d2 = __dsub_rn(ft[g].point_w[2], cen2);
                    const double depth = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2)));   // ref: :117-118
                    s_P[f] = __dmul_rn(ft[g].normal[0], depth);                                       // ref: :119
                    s_P[NF + f] = __dmul_rn(ft[g].normal[1], depth);
                    s_P[2 * NF + f] = __dmul_rn(ft[g].normal[2], depth);
                    s_zi[f] = 1.0 / __dmul_rn(ft[g].normal[2], depth);
                }
                s_px[f] = make_float2(ft[g].px[0], ft[g].px[1]);
                s_valid[f] = eligible ? 2 : 0;
            }
        }
    }
#endif
    __syncthreads();

    for (int level = a.max_level - 1; level >= a.min_level; --level) {
        const int cols = a.geo.w[level], rows = a.geo.h[level];
        const float tScale = 1.0f / (float)(1 << level);
        const double scale = (double)tScale;
        const double fs = (double)a.f * scale;     // == (v * f) * scale bit-exactly, scale being a power of two
        const double fs2 = fs * fs;

#ifdef DSDTM_SA_TIMING
        const long long tks0 = clock64();     // x[2] of a level's first log entry: cycles of the level staging; x[1] of the very first entry: the prologue
#endif
        // ------------------------------------------------ level staging = the pose-independent part of GetJocabianMat (ref: :62-132)
#if DSDTM_SA_STAGE == 1
        // Two features per thread and step, straight-line: both features' 21 aligned window loads are in flight together (the loop
        // used to pay two dependent DRAM round trips per feature -- record, then window -- four times per level: ~15 k cycles of
        // a 360 k-cycle pair). The level-independent part (point in the reference camera, 1 / z) is done once in the prologue.
        {
            constexpr int SG = DSDTM_SA_STAGE_G;
            const uint8_t* __restrict__ img = ref_frame + a.geo.off[level];
            for (int f0 = tid; f0 < nfeat; f0 += SG * NT) {
                uint32_t w[SG][7][3];
                unsigned a0[SG];
                bool val[SG];
#pragma unroll
                for (int g = 0; g < SG; ++g) {
                    const int f = f0 + g * NT;
                    val[g] = false; a0[g] = 0u;
                    if (f < nfeat && (s_valid[f] & 2)) {                                                  // ref: :86 and the zero-point test of :95
                        const float2 p = s_px[f];
                        const double px = (double)p.x * scale, py = (double)p.y * scale;                  // ref: :89-91
                        const int boarder = 3;                                                             // ref: :67
                        if (!(px - boarder < 0 || py - boarder < 0 || px + boarder >= cols || py + boarder >= rows)) {   // ref: :95-96
                            val[g] = true;
                            const int fxi = __double2int_rd(px), fyi = __double2int_rd(py);
                            // rows fyi-3 .. fyi+3, cols fxi-3 .. fxi+3 : three aligned 32-bit loads + funnel shifts per row
                            a0[g] = (unsigned)(fyi - 3) * (unsigned)cols + (unsigned)(fxi - 3);
                        }
                    }
#if DSDTM_SA_LD64 & 1
                    // 8-byte aligned 64-bit loads: the 7 bytes of a row start at byte s = ad & 7 of the first pair of words and need the second
                    // pair only for s >= 2 -- 1.75 load requests per row instead of three 32-bit ones (the staging is bound by L1 request slots)
#pragma unroll
                    for (int r = 0; r < 7; ++r) {
                        const unsigned ad = a0[g] + (unsigned)r * (val[g] ? (unsigned)cols : 0u);
                        const uint2* wp = reinterpret_cast<const uint2*>(img + (ad & ~7u));
                        const uint2 lo = __ldg(wp);
                        uint2 hi = make_uint2(0u, 0u);
                        if ((ad & 7u) >= 2u) hi = __ldg(wp + 1);
                        const bool up = (ad & 4u) != 0u;
                        w[g][r][0] = up ? lo.y : lo.x; w[g][r][1] = up ? hi.x : lo.y; w[g][r][2] = up ? hi.y : hi.x;
                    }
#else
                    // unconditional: no branch between the loads. A feature that is not staged reads the first 12 bytes of the level seven
                    // times (row stride 0), which exist for every level geometry (a slot ends with 64 zero bytes)
                    const unsigned rstride = val[g] ? (unsigned)cols : 0u;
#pragma unroll
                    for (int r = 0; r < 7; ++r) {
                        const unsigned ad = a0[g] + (unsigned)r * rstride;
                        const uint32_t* wp = reinterpret_cast<const uint32_t*>(img + (ad & ~3u));
                        w[g][r][0] = __ldg(wp); w[g][r][1] = __ldg(wp + 1); w[g][r][2] = __ldg(wp + 2);
                    }
#endif
                }
#if DSDTM_SA_PREF_NEXT
                // while this level's windows are in flight: pull the NEXT (finer) level's reference rows of the same features into L2, so that
                // the next staging finds them there instead of in DRAM. Fire and forget; addresses only.
                if (level > a.min_level) {
                    const int ncols = a.geo.w[level - 1], nrows = a.geo.h[level - 1];
                    const uint8_t* __restrict__ nimg = ref_frame + a.geo.off[level - 1];
                    const float nscale = 2.0f * tScale;
#pragma unroll
                    for (int g = 0; g < SG; ++g) {
                        const int f = f0 + g * NT;
                        if (f < nfeat && (s_valid[f] & 2)) {
                            const float2 p = s_px[f];
                            const int nx = (int)(p.x * nscale), ny = (int)(p.y * nscale);
                            if (nx >= 3 && ny >= 3 && nx + 3 < ncols && ny + 3 < nrows) {
                                const uint8_t* q = nimg + (unsigned)(ny - 3) * (unsigned)ncols + (unsigned)(nx - 3);
#pragma unroll
                                for (int r = 0; r < 7; ++r) {
                                    asm volatile("prefetch.global.L2 [%0];" :: "l"(q + (size_t)r * ncols));
                                    if (((size_t)(q + (size_t)r * ncols) & 31) > 25) asm volatile("prefetch.global.L2 [%0];" :: "l"(q + (size_t)r * ncols + 6));
                                }
                            }
                        }
                    }
                }
#endif
#if DSDTM_SA_PREF
                // while the reference windows are in flight: ask for the rows of the CURRENT image the first iteration of this level
                // will gather (pose at level start). Addresses only -- fp32 projection, nothing here reaches a result.
#pragma unroll
                for (int g = 0; g < SG; ++g) {
                    const int f = f0 + g * NT;
                    if (val[g]) {
                        const float q0 = (float)s_T[0], q1 = (float)s_T[1], q2 = (float)s_T[2], q3 = (float)s_T[3];
                        const float v0 = (float)s_P[f], v1 = (float)s_P[NF + f], v2 = (float)s_P[2 * NF + f];
                        const float uv0 = 2.f * (q2 * v2 - q3 * v1), uv1 = 2.f * (q3 * v0 - q1 * v2), uv2 = 2.f * (q1 * v1 - q2 * v0);
                        const float X = v0 + q0 * uv0 + (q2 * uv2 - q3 * uv1) + (float)s_T[4];
                        const float Y = v1 + q0 * uv1 + (q3 * uv0 - q1 * uv2) + (float)s_T[5];
                        const float Z = v2 + q0 * uv2 + (q1 * uv1 - q2 * uv0) + (float)s_T[6];
                        const float iz = __frcp_rn(Z);
                        const float uu = (a.fx * X * iz + a.cx) * tScale, vv = (a.fy * Y * iz + a.cy) * tScale;
                        if (uu >= 3.f && vv >= 3.f && uu < (float)(cols - 3) && vv < (float)(rows - 3)) {
                            const uint8_t* q = cur_frame + a.geo.off[level] + ((unsigned)((int)vv - 2) * (unsigned)cols + (unsigned)((int)uu - 2));
#pragma unroll
                            for (int r = 0; r < 5; ++r) {
                                asm volatile("prefetch.global.L1 [%0];" :: "l"(q + (size_t)r * cols));
                                asm volatile("prefetch.global.L1 [%0];" :: "l"(q + (size_t)r * cols + 7));
                            }
                        }
                    }
                }
#endif
#pragma unroll
                for (int g = 0; g < SG; ++g) {
                    const int f = f0 + g * NT;
                    if (f < nfeat) {
#pragma unroll
                        for (int r = 0; r < 7; ++r) {
                            const int sh = 8 * ((a0[g] + (unsigned)r * (val[g] ? (unsigned)cols : 0u)) & 3u);
                            s_nb[(2 * r) * NF + f] = __funnelshift_r(w[g][r][0], w[g][r][1], sh);
                            s_nb[(2 * r + 1) * NF + f] = __funnelshift_r(w[g][r][1], w[g][r][2], sh);
                        }
                        s_valid[f] = (uint8_t)((s_valid[f] & 2) | (val[g] ? 1 : 0));
                    }
                }
            }
        }
#else
        {
            const uint8_t* __restrict__ img = ref_frame + a.geo.off[level];
            for (int f = tid; f < nfeat; f += NT) {
                const dsdtm_ref_feat ft = a.feats[(size_t)pair * a.feat_stride + f];
                bool valid = false;
                if (ft.initial) {                                                                      // ref: :86
                    const double px = (double)ft.px[0] * scale, py = (double)ft.px[1] * scale;       // ref: :89-91
                    const bool zero = (ft.point_w[0] == 0.0 && ft.point_w[1] == 0.0 && ft.point_w[2] == 0.0);
                    const int boarder = 3;                                                             // ref: :67
                    if (!(zero || px - boarder < 0 || py - boarder < 0 || px + boarder >= cols || py + boarder >= rows)) {   // ref: :95-96
                        valid = true;
                        const double d0 = __dsub_rn(ft.point_w[0], cen0), d1 = __dsub_rn(ft.point_w[1], cen1), d2 = __dsub_rn(ft.point_w[2], cen2);
                        const double depth = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2)));   // ref: :117-118
                        s_P[f] = __dmul_rn(ft.normal[0], depth);                                      // ref: :119
                        s_P[NF + f] = __dmul_rn(ft.normal[1], depth);
                        s_P[2 * NF + f] = __dmul_rn(ft.normal[2], depth);
                        s_zi[f] = 1.0 / __dmul_rn(ft.normal[2], depth);
                        const int fxi = __double2int_rd(px), fyi = __double2int_rd(py);
                        s_sub[f] = make_float2((float)(px - fxi), (float)(py - fyi));                 // exact: px is a float scaled by 2^-level
                        // rows fyi-3 .. fyi+3, cols fxi-3 .. fxi+3 : three aligned 32-bit loads + funnel shifts per row
                        const unsigned a0 = (unsigned)(fyi - 3) * (unsigned)cols + (unsigned)(fxi - 3);
#pragma unroll
                        for (int r = 0; r < 7; ++r) {
                            const unsigned ad = a0 + (unsigned)r * (unsigned)cols;
                            const uint32_t* wp = reinterpret_cast<const uint32_t*>(img + (ad & ~3u));
                            const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
                            const int sh = 8 * (ad & 3u);
                            s_nb[(2 * r) * NF + f] = __funnelshift_r(w0, w1, sh);
                            s_nb[(2 * r + 1) * NF + f] = __funnelshift_r(w1, w2, sh);
                        }
                    }
                }
                s_valid[f] = valid ? 1 : 0;
            }
        }
#endif
        __syncthreads();
#if DSDTM_SA_COMPACT
        // stable compaction of the staged features (warp 0, ballot + popc per chunk of 32): deterministic, ascending feature order
        if (warp == 0) {
            int base = 0;
            for (int c = 0; c < nfeat; c += 32) {
                const int f = c + lane;
                const bool v = f < nfeat && (s_valid[f] & 1);
                const unsigned m = __ballot_sync(0xffffffffu, v);
                if (v) s_idx[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)f;
                base += __popc(m);
            }
            if (lane == 0) s_nvalid = base;
        }
        __syncthreads();
        const int nvalid = s_nvalid;
#endif
#ifdef DSDTM_SA_TIMING
        const long long tks1 = clock64();
#endif

        unsigned prev_vis = 0;
        const uint8_t* __restrict__ cimg = cur_frame + a.geo.off[level];

        // ------------------------------------------------ GaussNewtonSolver (ref: :301-344)
        for (int it = 0; it < a.max_iters; ++it) {
#ifdef DSDTM_SA_TIMING
            const long long tk0 = clock64();     // anatomy build: cycles of the pass / the reductions / the one-lane tail go into the log's x[3..5]
#endif
            const double qw = s_T[0], qx = s_T[1], qy = s_T[2], qz = s_T[3];
            const double t0 = s_T[4], t1 = s_T[5], t2 = s_T[6];
#if DSDTM_SA_RMAT && !DSDTM_SA_STRICT
            const double r00 = 1.0 - 2.0 * (qy * qy + qz * qz), r01 = 2.0 * (qx * qy - qw * qz), r02 = 2.0 * (qx * qz + qw * qy);
            const double r10 = 2.0 * (qx * qy + qw * qz), r11 = 1.0 - 2.0 * (qx * qx + qz * qz), r12 = 2.0 * (qy * qz - qw * qx);
            const double r20 = 2.0 * (qx * qz - qw * qy), r21 = 2.0 * (qy * qz + qw * qx), r22 = 1.0 - 2.0 * (qx * qx + qy * qy);
#endif
            double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, acc4 = 0, acc5 = 0, accc = 0;
            int cnt = 0;
            static_assert(DSDTM_MAX_FEATS_LIMIT <= 32 * 32, "vis_mask holds one bit per feature of a lane: at most 32 features per lane at WPP = 1");
            unsigned vis_mask = 0;

            // stage 1 of a feature: projection, bounds test and the ten aligned 32-bit loads of its 5x5 current-image window
            auto stage1 = [&](int f, Pre& p) {
                p.valid = false; p.vis = false;
                if (f >= nfeat) return;
                if (!(s_valid[f] & 1)) return;
                p.valid = true;
                const double P0 = s_P[f], P1 = s_P[NF + f], P2 = s_P[2 * NF + f];
                double Q0, Q1, Q2;
#if DSDTM_SA_RMAT && !DSDTM_SA_STRICT
                Q0 = fma(r02, P2, fma(r01, P1, fma(r00, P0, t0)));
                Q1 = fma(r12, P2, fma(r11, P1, fma(r10, P0, t1)));
                Q2 = fma(r22, P2, fma(r21, P1, fma(r20, P0, t2)));
#else
                qrot(qw, qx, qy, qz, P0, P1, P2, Q0, Q1, Q2);                                  // ref: :254
                Q0 = __dadd_rn(Q0, t0); Q1 = __dadd_rn(Q1, t1); Q2 = __dadd_rn(Q2, t2);
#endif
                // Camera2Pixel * tScale (ref: src/Camera.cpp:167-171, :255): (fx*X)/Z + cx
#if DSDTM_SA_STRICT
                const double u = __dmul_rn(__dadd_rn(__ddiv_rn(__dmul_rn(fx, Q0), Q2), cx), scale);
                const double v = __dmul_rn(__dadd_rn(__ddiv_rn(__dmul_rn(fy, Q1), Q2), cy), scale);
#else
                // one reciprocal instead of two divisions: u, v within 1 ulp (1e-13 px) of the reference's (fx*X)/Z
                const double iz = 1.0 / Q2;
                const double u = __dmul_rn(__dadd_rn(__dmul_rn(__dmul_rn(fx, Q0), iz), cx), scale);
                const double v = __dmul_rn(__dadd_rn(__dmul_rn(__dmul_rn(fy, Q1), iz), cy), scale);
#endif
                const double uf = floor(u), vf = floor(v);
                // ref: :262 with border 3; evaluated in double so that NaN / huge values are rejected like the reference's INT_MIN
                if (!(uf >= 3.0 && vf >= 3.0 && uf < (double)(cols - 3) && vf < (double)(rows - 3))) return;
                p.vis = true;
                const int ui = (int)uf, vi = (int)vf;
                const double su = u - uf, sv = v - vf;
#if DSDTM_SA_PIPELINE
                p.su = su; p.sv = sv;
#else
                {   // ref: :267-270; the second factor carries SA_WSCALE (a power of two: the products round exactly as unscaled)
                    const double osv = (1.0 - sv) * SA_WSCALE, svs = sv * SA_WSCALE;
                    p.tl = __dmul_rn(1.0 - su, osv); p.tr = __dmul_rn(su, osv);
                    p.bl = __dmul_rn(1.0 - su, svs); p.br = __dmul_rn(su, svs);
                }
#endif
                const unsigned c0w = (unsigned)(vi - 2) * (unsigned)cols + (unsigned)(ui - 2);
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    const unsigned ad = c0w + (unsigned)r * (unsigned)cols;
#if DSDTM_SA_LD64 & 2
                    // bytes s .. s+4 of the row, s = ad & 7: one 8-byte aligned 64-bit load, plus the next word only for s >= 4
                    const uint2* wp = reinterpret_cast<const uint2*>(cimg + (ad & ~7u));
                    const uint2 w01 = __ldg(wp);
                    uint32_t w2 = 0u;
                    if ((ad & 7u) >= 4u) w2 = __ldg(reinterpret_cast<const uint32_t*>(wp + 1));
                    const bool up = (ad & 4u) != 0u;
                    const uint32_t lo = up ? w01.y : w01.x, hi = up ? w2 : w01.y;
#else
                    const uint32_t* wp = reinterpret_cast<const uint32_t*>(cimg + (ad & ~3u));
                    const uint32_t lo = __ldg(wp), hi = __ldg(wp + 1);
#endif
                    const int sh = 8 * (ad & 3u);
                    p.cw0[r] = __funnelshift_r(lo, hi, sh);
                    p.cw1[r] = hi >> sh;
                }
            };

            // stage 2: the fp64 arithmetic of one feature. `first` (iteration 0 of a level, uniform over the CTA) also derives the
            // pose-independent second moments Sxx, Sxy, Syy of every VALID feature (visible or not) and parks them in shared memory for
            // H. One instantiation serves both cases (round 1 compiled the pass twice: half of the 120 KB of SASS, and the first
            // iteration of every level ran cold code -- profiles/r2_sparse_align.md).
#if DSDTM_SA_MOM == 3
            auto stage2 = [&](auto first_tag, int f, int k, const Pre& me) {
                constexpr bool first = decltype(first_tag)::value;
#else
            auto stage2 = [&](bool first, int f, int k, const Pre& me) {
#endif
                if (!(first ? me.valid : me.vis)) return;
                const bool vis = me.vis;
#if DSDTM_SA_PIPELINE
                const double osv_ = (1.0 - me.sv) * SA_WSCALE, svs_ = me.sv * SA_WSCALE;
                const double me_tl = __dmul_rn(1.0 - me.su, osv_), me_tr = __dmul_rn(me.su, osv_);
                const double me_bl = __dmul_rn(1.0 - me.su, svs_), me_br = __dmul_rn(me.su, svs_);
#else
                const double me_tl = me.tl, me_tr = me.tr, me_bl = me.bl, me_br = me.br;
#endif
                if (vis) { vis_mask |= 1u << k; ++cnt; }
                RefRowsT<true> R;
                {
#if DSDTM_SA_STAGE == 1
                    // the level pixel and its fraction in fp32: exact (a float scaled by 2^-level, minus its floor), the same values
                    // as the reference's double expressions (ref: :89-91, :123-127)
                    const float2 p0 = s_px[f];
                    const float plx = p0.x * tScale, ply = p0.y * tScale;
                    const double sx = (double)(plx - floorf(plx)), sy = (double)(ply - floorf(ply));
#else
                    const float2 sub = s_sub[f];
                    const double sx = (double)sub.x, sy = (double)sub.y;
#endif
                    const double osy = (1.0 - sy) * SA_WSCALE, sys = sy * SA_WSCALE;
                    R.w00 = __dmul_rn(1.0 - sx, osy); R.w01 = __dmul_rn(sx, osy);
                    R.w10 = __dmul_rn(1.0 - sx, sys); R.w11 = __dmul_rn(sx, sys);             // ref: :129-132
                }
                R.nb = s_nb + f; R.NF = NF;
                R.load_row(0, 0); R.load_row(1, 1);
                R.grid_row(0, 0, 1);
                R.load_row(0, 2);
                R.grid_row(1, 1, 0);
                double Cw[2][5];
                auto cvt_cur = [&](int slot, int r) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) Cw[slot][c] = u8_to_f64s(me.cw0[r], c);
                    Cw[slot][4] = u8_to_f64s(me.cw1[r], 0);
                };
                cvt_cur(0, 0);
                double Sx = 0, Sy = 0, c2 = 0, Sxx = 0, Sxy = 0, Syy = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    // after the prologue Nd[0] = row 2, Nd[1] = row 1: row r+2 sits in slot (r & 1), row r+3 goes to the other
                    const int sa = (r & 1), sb = sa ^ 1;
                    R.load_row(sb, r + 3);
                    R.grid_row((r + 2) % 3, sa, sb);
                    cvt_cur((r + 1) & 1, r + 1);
                    const int g0 = r % 3, g1 = (r + 1) % 3, g2 = (r + 2) % 3;   // G rows r, r+1, r+2
                    const int ca = r & 1, cb = ca ^ 1;                          // Cw rows r, r+1
#if DSDTM_SA_MOM == 2
                    double dxa[4], dya[4];
#endif
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double refv = R.G[g1][c + 1];
                        // 2*dx, 2*dy: the reference's 0.5 factor (ref: :150-158) is applied once to the sums below (exact: power of two)
                        const double dx2 = __dsub_rn(R.G[g1][c + 2], R.G[g1][c]);
                        const double dy2 = __dsub_rn(R.G[g2][c + 1], R.G[g0][c + 1]);
#if DSDTM_SA_MOM == 0 || DSDTM_SA_MOM == 3
                        if (first) { Sxx = fma(dx2, dx2, Sxx); Sxy = fma(dx2, dy2, Sxy); Syy = fma(dy2, dy2, Syy); }
#elif DSDTM_SA_MOM == 1
                        Sxx = fma(dx2, dx2, Sxx); Sxy = fma(dx2, dy2, Sxy); Syy = fma(dy2, dy2, Syy);
#else
                        dxa[c] = dx2; dya[c] = dy2;
#endif
                        const double cur = bil(me_tl, me_tr, me_bl, me_br, Cw[ca][c], Cw[ca][c + 1], Cw[cb][c], Cw[cb][c + 1]);   // ref: :281
                        const double res = __dsub_rn(cur, refv);                                          // ref: :282
#if DSDTM_SA_STRICT
                        c2 = __dadd_rn(c2, __dmul_rn(res, res));                                          // ref: :284
#else
                        c2 = fma(res, res, c2);
#endif
                        Sx = fma(dx2, res, Sx);
                        Sy = fma(dy2, res, Sy);
                    }
#if DSDTM_SA_MOM == 2
                    if (first) {                              // uniform over the CTA: a real branch (the asm keeps it from becoming 6 selects per pixel)
                        asm volatile("" ::: "memory");
#pragma unroll
                        for (int c = 0; c < 4; ++c) { Sxx = fma(dxa[c], dxa[c], Sxx); Sxy = fma(dxa[c], dya[c], Sxy); Syy = fma(dya[c], dya[c], Syy); }
                    }
#endif
                }
                if (first) { s_S[f] = (0.25 * SA_RESTORE2) * Sxx; s_S[NF + f] = (0.25 * SA_RESTORE2) * Sxy; s_S[2 * NF + f] = (0.25 * SA_RESTORE2) * Syy; }   // (0.5 d2)^2, exact scaling
                if (first && !vis) return;
                Sx *= 0.5 * SA_RESTORE2; Sy *= 0.5 * SA_RESTORE2;
                // GetJocabianBA(P) rows (ref: :169-193); a1 = b0 = 0.   b_j = fs * (a * Sx + b * Sy)
                const double P0 = s_P[f], P1 = s_P[NF + f];
                const double zi = s_zi[f], zi2 = zi * zi;
                const double a0 = -zi, a2 = P0 * zi2, a3 = P1 * a2, a4 = -(1.0 + P0 * a2), a5 = P1 * zi;
                const double b1 = -zi, b2 = P1 * zi2, b3 = 1.0 + P1 * b2, b4 = -P0 * b2, b5 = -P0 * zi;
                acc0 += fs * (a0 * Sx); acc1 += fs * (b1 * Sy); acc2 += fs * (a2 * Sx + b2 * Sy);
                acc3 += fs * (a3 * Sx + b3 * Sy); acc4 += fs * (a4 * Sx + b4 * Sy); acc5 += fs * (a5 * Sx + b5 * Sy);
                accc += c2;
            };

            {
                const bool first = (it == 0);
                int kq = 0;
#if DSDTM_SA_MOM == 3
                if (first) {
                    for (int f = tid; f < nfeat; f += NT, ++kq) { Pre me; stage1(f, me); stage2(std::true_type{}, f, kq, me); }
                } else {
                    for (int f = tid; f < nfeat; f += NT, ++kq) { Pre me; stage1(f, me); stage2(std::false_type{}, f, kq, me); }
                }
#else
#if DSDTM_SA_PIPELINE
                {
                    Pre me, nx;
                    int f = tid;
                    stage1(f, me);
                    for (; f < nfeat; f += NT, ++kq) {
                        stage1(f + NT, nx);            // the next feature's projection and window loads are in flight during this one's arithmetic
                        stage2(first, f, kq, me);
                        me = nx;
                    }
                }
#elif DSDTM_SA_COMPACT
                for (int k = tid; k < nvalid; k += NT, ++kq) {
                    const int f = s_idx[k];
                    Pre me;
                    stage1(f, me);
                    stage2(first, f, kq, me);
                }
#else
                for (int f = tid; f < nfeat; f += NT, ++kq) {
                    Pre me;
                    stage1(f, me);
                    stage2(first, f, kq, me);
                }
#endif
#endif
            }
#ifdef DSDTM_SA_TIMING
            const long long tk1 = clock64();
#endif
            accc *= SA_RESTORE2;      // the lane's sum of squared residuals carried 2^-48 (exact: a sum of scaled terms is the scaled sum)
            cnt = __reduce_add_sync(0xffffffffu, cnt);
#if DSDTM_SA_TREDUCE
            {
                const double av7[7] = { acc0, acc1, acc2, acc3, acc4, acc5, accc };
                const double tot = warp_sum_transposed(av7);          // lane l < 7: total of value l
                if (WPP > 1) {
                    if (lane < 7) s_red[warp][lane] = tot;
                    if (lane == 0) s_cnt[warp] = cnt;
                } else {
                    acc0 = __shfl_sync(0xffffffffu, tot, 0); acc1 = __shfl_sync(0xffffffffu, tot, 1); acc2 = __shfl_sync(0xffffffffu, tot, 2);
                    acc3 = __shfl_sync(0xffffffffu, tot, 3); acc4 = __shfl_sync(0xffffffffu, tot, 4); acc5 = __shfl_sync(0xffffffffu, tot, 5);
                    accc = __shfl_sync(0xffffffffu, tot, 6);
                }
            }
#else
            acc0 = warp_sum(acc0); acc1 = warp_sum(acc1); acc2 = warp_sum(acc2); acc3 = warp_sum(acc3);
            acc4 = warp_sum(acc4); acc5 = warp_sum(acc5); accc = warp_sum(accc);
            if (WPP > 1 && lane == 0) {
                s_red[warp][0] = acc0; s_red[warp][1] = acc1; s_red[warp][2] = acc2; s_red[warp][3] = acc3;
                s_red[warp][4] = acc4; s_red[warp][5] = acc5; s_red[warp][6] = accc;
                s_cnt[warp] = cnt;
            }
#endif
            const int need_H = __syncthreads_or((it == 0) || (vis_mask != prev_vis));
            prev_vis = vis_mask;
            if (need_H) {
                // H = fs^2 sum_visible (Sxx a a^T + Sxy (a b^T + b a^T) + Syy b b^T), lower triangle packed row-major, assembled from
                // the parked second moments (no image arithmetic here):  H += a u^T + b w^T  with  u = cxx a + cxy b,  w = cxy a + cyy b
                double hacc[21];
#pragma unroll
                for (int i = 0; i < 21; ++i) hacc[i] = 0.0;
                auto add_H = [&](int f) {
                    const double P0 = s_P[f], P1 = s_P[NF + f];
                    const double zi = s_zi[f], zi2 = zi * zi;
                    const double av[6] = { -zi, 0.0, P0 * zi2, P1 * (P0 * zi2), -(1.0 + P0 * (P0 * zi2)), P1 * zi };
                    const double bv[6] = { 0.0, -zi, P1 * zi2, 1.0 + P1 * (P1 * zi2), -P0 * (P1 * zi2), -P0 * zi };
                    const double cxx = fs2 * s_S[f], cxy = fs2 * s_S[NF + f], cyy = fs2 * s_S[2 * NF + f];
                    double u[6], w[6];
                    u[0] = cxx * av[0]; w[0] = cxy * av[0];             // bv[0] == 0
                    u[1] = cxy * bv[1]; w[1] = cyy * bv[1];             // av[1] == 0
#pragma unroll
                    for (int c = 2; c < 6; ++c) { u[c] = fma(cxy, bv[c], cxx * av[c]); w[c] = fma(cyy, bv[c], cxy * av[c]); }
#pragma unroll
                    for (int r = 0; r < 6; ++r)
#pragma unroll
                        for (int c = 0; c <= r; ++c) {
                            if (r == 0) hacc[0] = fma(av[0], u[0], hacc[0]);
                            else if (r == 1) hacc[r * (r + 1) / 2 + c] = fma(bv[1], w[c], hacc[r * (r + 1) / 2 + c]);
                            else hacc[r * (r + 1) / 2 + c] = fma(bv[r], w[c], fma(av[r], u[c], hacc[r * (r + 1) / 2 + c]));
                        }
                };
                int kk = 0;
#if DSDTM_SA_COMPACT
                for (int k = tid; k < nvalid; k += NT, ++kk)
                    if ((vis_mask >> kk) & 1u) add_H(s_idx[k]);
#else
                for (int f = tid; f < nfeat; f += NT, ++kk)
                    if ((vis_mask >> kk) & 1u) add_H(f);
#endif
#if DSDTM_SA_TREDUCE
                {
                    const double h = warp_sum_transposed(hacc);       // lane l < 21: total of entry l
                    if (lane < 21) { if (WPP > 1) s_redH[warp][lane] = h; else s_H[lane] = h; }
                }
#else
#pragma unroll
                for (int i = 0; i < 21; ++i) {
                    const double h = warp_sum(hacc[i]);
                    if (WPP > 1) { if (lane == 0) s_redH[warp][i] = h; }
                    else if (lane == 0) s_H[i] = h;
                }
#endif
                if (WPP > 1) __syncthreads();
            }
            if (warp == 0) {
                if (WPP > 1) {
                    if (need_H && lane < 21) {
                        double h = 0;
                        for (int w = 0; w < WPP; ++w) h += s_redH[w][lane];
                        s_H[lane] = h;
                    }
                    double red = 0;
                    if (lane < 7) for (int w = 0; w < WPP; ++w) red += s_red[w][lane];
                    int npts = 0;
                    for (int w = 0; w < WPP; ++w) npts += s_cnt[w];
                    acc0 = __shfl_sync(0xffffffffu, red, 0); acc1 = __shfl_sync(0xffffffffu, red, 1);
                    acc2 = __shfl_sync(0xffffffffu, red, 2); acc3 = __shfl_sync(0xffffffffu, red, 3);
                    acc4 = __shfl_sync(0xffffffffu, red, 4); acc5 = __shfl_sync(0xffffffffu, red, 5);
                    accc = __shfl_sync(0xffffffffu, red, 6);
                    cnt = npts;
                }
                __syncwarp();
#if DSDTM_SA_TAIL_LANES
                {   // every lane of warp 0 holds the same sums: the tail runs on all of them (no extra cost in SIMT), lane 0 publishes
                    const bool l0 = (lane == 0);
#else
                if (lane == 0) {
                    const bool l0 = true;
#endif
#ifdef DSDTM_SA_TIMING
                    const long long tk2 = clock64();
#endif
                    const double chi2New = accc / (double)(16 * cnt);                      // ref: :298 (NaN if nothing visible)
                    const double bvec[6] = { acc0, acc1, acc2, acc3, acc4, acc5 };
                    double x[6];
#if DSDTM_SA_TAIL_LANES || DSDTM_SA_TAIL_CONST
                    solve_all_lanes(s_H, s_F, need_H != 0, bvec, x, lane);                 // ref: :318
#else
                    solve_and_update(s_H, s_F, need_H != 0, bvec, x);                      // ref: :318
#endif
                    int flags = 0;
                    bool stop = false;
                    if (isnan(x[0])) { stop = true; flags |= 4; }                          // ref: :321-326
                    if ((it > 0 && chi2New > s_chi2prev) || stop) {                        // ref: :328-332
                        if (l0) {
#pragma unroll
                            for (int q = 0; q < 7; ++q) s_T[q] = s_Told[q];
                        }
                        flags |= 2;
                        stop = true;
                    } else {
                        double Tn[7], Tcur[7];
#pragma unroll
                        for (int q = 0; q < 7; ++q) Tcur[q] = s_T[q];
#if DSDTM_SA_TAIL_LANES
                        pose_update_lanes(Tcur, x, Tn, s_coef, lane);                      // ref: :335
                        __syncwarp();                                                      // every lane has read s_T and s_chi2prev
#elif DSDTM_SA_TAIL_CONST
                        pose_update_const(Tcur, x, Tn);                                    // ref: :335
#else
                        pose_update(Tcur, x, Tn);                                          // ref: :335
#endif
                        if (l0) {
#pragma unroll
                            for (int q = 0; q < 7; ++q) { s_Told[q] = Tcur[q]; s_T[q] = Tn[q]; } // ref: :336-337
                            s_chi2prev = chi2New;                                          // ref: :339
                        }
                        flags |= 1;
#if DSDTM_SA_TAIL_CONST
                        // max_q |x_q| <= 1e-8 (ref: :341; Eigen's maxCoeff and fmax both pass over a NaN) as six compares into one predicate
                        bool small = true;
#pragma unroll
                        for (int q = 0; q < 6; ++q) small = small && !(fabs(x[q]) > 1e-8);
                        if (small) { stop = true; flags |= 8; }
#else
                        double mx = 0;
#pragma unroll
                        for (int q = 0; q < 6; ++q) mx = fmax(mx, fabs(x[q]));
                        if (mx <= 1e-8) { stop = true; flags |= 8; }                       // ref: :341
#endif
                    }
                    if (l0) {
                    s_npts = cnt;
                    s_stop = stop ? 1 : 0;
                    }
                    if (l0 && a.log) {
                        const int n = s_nlog;
                        if (n < a.log_cap) {
                            dsdtm_iter_log* e = a.log + (size_t)pair * a.log_cap + n;
                            e->level = level; e->iter = it; e->n_pts = cnt; e->flags = flags; e->chi2 = chi2New;
#pragma unroll
                            for (int q = 0; q < 6; ++q) e->x[q] = x[q];
#ifdef DSDTM_SA_TIMING
                            e->x[3] = (double)(tk1 - tk0); e->x[4] = (double)(tk2 - tk1); e->x[5] = (double)(clock64() - tk2);
                            if (it == 0) { e->x[2] = (double)(tks1 - tks0); if (level == a.max_level - 1) e->x[1] = (double)(tks0 - tkp0); }
#endif
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
        if (tid < 7) s_Told[tid] = s_T[tid];
        if (tid == 0) { s_stop = 0; s_chi2prev = 0.0; }
        __syncthreads();
    }
    if (tid < 7) a.poses_out[7 * pair + tid] = s_T[tid];
    if (tid == 0) {
        a.n_tracked[pair] = s_npts;
        if (a.n_log) a.n_log[pair] = s_nlog;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Variant 1 ("workspace"): the pose-independent reference samples (ref, 2*dx, 2*dy for the 16 patch pixels) are computed
// ONCE per level by the owning lane -- with exactly the same expressions as variant 0 -- and parked in a per-pair global
// workspace laid out [48][nf] so that consecutive lanes read consecutive doubles (256-byte coalesced, ld.global.cg: the
// data lives in L2, 115 KB per resident pair). An iteration then only does the current-image side: 12 loads + 16 bilinear
// samples per feature instead of re-deriving the 6x6 grid from bytes (~465 instead of ~700 instructions per feature and
// iteration), and shared memory drops to 49 B / feature. A lane only ever reads what it wrote itself.
template <int WPP>
__global__ void __launch_bounds__(32 * WPP, (WPP <= 2) ? 8 : (WPP == 3 ? 4 : (WPP == 4 ? 4 : (WPP == 5 ? 3 : 1)))) sparse_align_ws_kernel(const SaArgs a)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int NF = a.nf;
    double* s_P = reinterpret_cast<double*>(s_raw);                                        // [3][NF] point in the ref camera
    double* s_S = reinterpret_cast<double*>(s_raw + (size_t)24 * NF);                      // [3][NF] Sxx, Sxy, Syy of the ref patch
    uint8_t* s_valid = s_raw + (size_t)48 * NF;                                            // [NF]
    __shared__ double s_red[WPP][8];
    __shared__ int s_cnt[WPP];
    __shared__ double s_redH[WPP][22];
    __shared__ double s_H[21];
    __shared__ double s_F[22];
    __shared__ double s_T[7], s_Told[7];
    __shared__ double s_chi2prev;
    __shared__ int s_stop, s_npts, s_nlog;

    constexpr int NT = 32 * WPP;
    const int pair = blockIdx.x + a.pair0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nfeat = min(a.n_feats[pair], NF);
    const uint8_t* __restrict__ ref_frame = a.frames + (size_t)a.ref_slots[pair] * a.frame_stride;
    const uint8_t* __restrict__ cur_frame = a.frames + (size_t)a.cur_slots[pair] * a.frame_stride;
    double* __restrict__ ws = a.ws + (size_t)pair * 48 * NF;
    const double cen0 = a.centers[3 * pair], cen1 = a.centers[3 * pair + 1], cen2 = a.centers[3 * pair + 2];
    const double fx = (double)a.fx, fy = (double)a.fy, cx = (double)a.cx, cy = (double)a.cy;

    if (tid < 7) { s_T[tid] = a.poses_in[7 * pair + tid]; s_Told[tid] = s_T[tid]; }
    if (tid == 0) { s_npts = 0; s_nlog = 0; s_stop = 0; s_chi2prev = 0.0; }
    __syncthreads();

    for (int level = a.max_level - 1; level >= a.min_level; --level) {
        const int cols = a.geo.w[level], rows = a.geo.h[level];
        const float tScale = 1.0f / (float)(1 << level);
        const double scale = (double)tScale;
        const double fs = (double)a.f * scale;     // == (v * f) * scale bit-exactly, scale being a power of two
        const double fs2 = fs * fs;

        // ------------------------------------------------ GetJocabianMat (ref: :62-166), pose-independent: once per level
        {
            const uint8_t* __restrict__ img = ref_frame + a.geo.off[level];
            for (int f = tid; f < nfeat; f += NT) {
                const dsdtm_ref_feat ft = a.feats[(size_t)pair * a.feat_stride + f];
                bool valid = false;
                if (ft.initial) {                                                                      // ref: :86
                    const double px = (double)ft.px[0] * scale, py = (double)ft.px[1] * scale;       // ref: :89-91
                    const bool zero = (ft.point_w[0] == 0.0 && ft.point_w[1] == 0.0 && ft.point_w[2] == 0.0);
                    const int boarder = 3;                                                             // ref: :67
                    if (!(zero || px - boarder < 0 || py - boarder < 0 || px + boarder >= cols || py + boarder >= rows)) {   // ref: :95-96
                        valid = true;
                        const double d0 = __dsub_rn(ft.point_w[0], cen0), d1 = __dsub_rn(ft.point_w[1], cen1), d2 = __dsub_rn(ft.point_w[2], cen2);
                        const double depth = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2)));   // ref: :117-118
                        s_P[f] = __dmul_rn(ft.normal[0], depth);                                      // ref: :119
                        s_P[NF + f] = __dmul_rn(ft.normal[1], depth);
                        s_P[2 * NF + f] = __dmul_rn(ft.normal[2], depth);
                        const int fxi = __double2int_rd(px), fyi = __double2int_rd(py);
                        const double sx = px - fxi, sy = py - fyi;
                        RefRows R;
                        R.w00 = __dmul_rn(1.0 - sx, 1.0 - sy); R.w01 = __dmul_rn(sx, 1.0 - sy);
                        R.w10 = __dmul_rn(1.0 - sx, sy); R.w11 = __dmul_rn(sx, sy);                   // ref: :129-132
                        // rows fyi-3 .. fyi+3, cols fxi-3 .. fxi+3 : three aligned 32-bit loads + funnel shifts per row
                        uint32_t lo[7], hi[7];
                        const unsigned a0 = (unsigned)(fyi - 3) * (unsigned)cols + (unsigned)(fxi - 3);
#pragma unroll
                        for (int r = 0; r < 7; ++r) {
                            const unsigned ad = a0 + (unsigned)r * (unsigned)cols;
                            const uint32_t* wp = reinterpret_cast<const uint32_t*>(img + (ad & ~3u));
                            const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
                            const int sh = 8 * (ad & 3u);
                            lo[r] = __funnelshift_r(w0, w1, sh);
                            hi[r] = __funnelshift_r(w1, w2, sh);
                        }
                        R.load_row_words(0, lo[0], hi[0]); R.load_row_words(1, lo[1], hi[1]);
                        R.grid_row(0, 0, 1);
                        R.load_row_words(0, lo[2], hi[2]);
                        R.grid_row(1, 1, 0);
                        double Sxx = 0, Sxy = 0, Syy = 0;
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const int sa = (r & 1), sb = sa ^ 1;
                            R.load_row_words(sb, lo[r + 3], hi[r + 3]);
                            R.grid_row((r + 2) % 3, sa, sb);
                            const int g0 = r % 3, g1 = (r + 1) % 3, g2 = (r + 2) % 3;
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const double refv = R.G[g1][c + 1];
                                const double dx2 = __dsub_rn(R.G[g1][c + 2], R.G[g1][c]);                // 2*dx (ref: :150-153)
                                const double dy2 = __dsub_rn(R.G[g2][c + 1], R.G[g0][c + 1]);            // 2*dy (ref: :155-158)
                                const int p = 4 * r + c;
                                __stcg(ws + (size_t)(3 * p) * NF + f, refv);
                                __stcg(ws + (size_t)(3 * p + 1) * NF + f, dx2);
                                __stcg(ws + (size_t)(3 * p + 2) * NF + f, dy2);
                                Sxx = fma(dx2, dx2, Sxx); Sxy = fma(dx2, dy2, Sxy); Syy = fma(dy2, dy2, Syy);
                            }
                        }
                        s_S[f] = 0.25 * Sxx; s_S[NF + f] = 0.25 * Sxy; s_S[2 * NF + f] = 0.25 * Syy;   // (0.5 d2)^2, exact scaling
                    }
                }
                s_valid[f] = valid ? 1 : 0;
            }
        }
        __syncthreads();

        unsigned prev_vis = 0;
        const uint8_t* __restrict__ cimg = cur_frame + a.geo.off[level];

        // ------------------------------------------------ GaussNewtonSolver (ref: :301-344)
        for (int it = 0; it < a.max_iters; ++it) {
            const double qw = s_T[0], qx = s_T[1], qy = s_T[2], qz = s_T[3];
            const double t0 = s_T[4], t1 = s_T[5], t2 = s_T[6];
            double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, acc4 = 0, acc5 = 0, accc = 0;
            int cnt = 0;
            unsigned vis_mask = 0;
            int k = 0;
            for (int f = tid; f < nfeat; f += NT, ++k) {
                if (!s_valid[f]) continue;
                const double P0 = s_P[f], P1 = s_P[NF + f], P2 = s_P[2 * NF + f];
                double Q0, Q1, Q2;
                qrot(qw, qx, qy, qz, P0, P1, P2, Q0, Q1, Q2);                                  // ref: :254
                Q0 = __dadd_rn(Q0, t0); Q1 = __dadd_rn(Q1, t1); Q2 = __dadd_rn(Q2, t2);
                // Camera2Pixel * tScale (ref: src/Camera.cpp:167-171, :255): (fx*X)/Z + cx
#if DSDTM_SA_STRICT
                const double u = __dmul_rn(__dadd_rn(__ddiv_rn(__dmul_rn(fx, Q0), Q2), cx), scale);
                const double v = __dmul_rn(__dadd_rn(__ddiv_rn(__dmul_rn(fy, Q1), Q2), cy), scale);
#else
                // one reciprocal instead of two divisions: u, v within 1 ulp (1e-13 px) of the reference's (fx*X)/Z
                const double iz = 1.0 / Q2;
                const double u = __dmul_rn(__dadd_rn(__dmul_rn(__dmul_rn(fx, Q0), iz), cx), scale);
                const double v = __dmul_rn(__dadd_rn(__dmul_rn(__dmul_rn(fy, Q1), iz), cy), scale);
#endif
                const double uf = floor(u), vf = floor(v);
                // ref: :262 with border 3; evaluated in double so that NaN / huge values are rejected like the reference's INT_MIN
                if (!(uf >= 3.0 && vf >= 3.0 && uf < (double)(cols - 3) && vf < (double)(rows - 3))) continue;
                vis_mask |= 1u << k;
                ++cnt;
                const int ui = (int)uf, vi = (int)vf;
                const double su = u - uf, sv = v - vf;
                const double tl = __dmul_rn(1.0 - su, 1.0 - sv), tr = __dmul_rn(su, 1.0 - sv);
                const double bl = __dmul_rn(1.0 - su, sv), br = __dmul_rn(su, sv);            // ref: :267-270
                // current-image 5x5 window rows vi-2 .. vi+2, cols ui-2 .. ui+2
                uint32_t cw0[5], cw1[5];
                {
                    const unsigned c0w = (unsigned)(vi - 2) * (unsigned)cols + (unsigned)(ui - 2);
#pragma unroll
                    for (int r = 0; r < 5; ++r) {
                        const unsigned ad = c0w + (unsigned)r * (unsigned)cols;
                        const uint32_t* wp = reinterpret_cast<const uint32_t*>(cimg + (ad & ~3u));
                        const uint32_t lo = __ldg(wp), hi = __ldg(wp + 1);
                        const int sh = 8 * (ad & 3u);
                        cw0[r] = __funnelshift_r(lo, hi, sh);
                        cw1[r] = hi >> sh;
                    }
                }
                const double* __restrict__ wf = ws + f;
                double Cw[2][5];
                auto cvt_cur = [&](int slot, int r) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) Cw[slot][c] = u8_to_f64(cw0[r], c);
                    Cw[slot][4] = u8_to_f64(cw1[r], 0);
                };
                cvt_cur(0, 0);
                double Sx = 0, Sy = 0, c2 = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    double rv[4], gx[4], gy[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int p = 4 * r + c;
                        rv[c] = __ldcg(wf + (size_t)(3 * p) * NF);
                        gx[c] = __ldcg(wf + (size_t)(3 * p + 1) * NF);
                        gy[c] = __ldcg(wf + (size_t)(3 * p + 2) * NF);
                    }
                    cvt_cur((r + 1) & 1, r + 1);
                    const int ca = r & 1, cb = ca ^ 1;                          // Cw rows r, r+1
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double cur = bil(tl, tr, bl, br, Cw[ca][c], Cw[ca][c + 1], Cw[cb][c], Cw[cb][c + 1]);   // ref: :281
                        const double res = __dsub_rn(cur, rv[c]);                                         // ref: :282
#if DSDTM_SA_STRICT
                        c2 = __dadd_rn(c2, __dmul_rn(res, res));                                          // ref: :284
#else
                        c2 = fma(res, res, c2);
#endif
                        Sx = fma(gx[c], res, Sx);
                        Sy = fma(gy[c], res, Sy);
                    }
                }
                Sx *= 0.5; Sy *= 0.5;
                // GetJocabianBA(P) rows (ref: :169-193); a1 = b0 = 0.   b_j = fs * (a * Sx + b * Sy)
                const double zi = 1.0 / P2, zi2 = zi * zi;
                const double a0 = -zi, a2 = P0 * zi2, a3 = P1 * a2, a4 = -(1.0 + P0 * a2), a5 = P1 * zi;
                const double b1 = -zi, b2 = P1 * zi2, b3 = 1.0 + P1 * b2, b4 = -P0 * b2, b5 = -P0 * zi;
                acc0 += fs * (a0 * Sx); acc1 += fs * (b1 * Sy); acc2 += fs * (a2 * Sx + b2 * Sy);
                acc3 += fs * (a3 * Sx + b3 * Sy); acc4 += fs * (a4 * Sx + b4 * Sy); acc5 += fs * (a5 * Sx + b5 * Sy);
                accc += c2;
            }
            acc0 = warp_sum(acc0); acc1 = warp_sum(acc1); acc2 = warp_sum(acc2); acc3 = warp_sum(acc3);
            acc4 = warp_sum(acc4); acc5 = warp_sum(acc5); accc = warp_sum(accc);
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (WPP > 1 && lane == 0) {
                s_red[warp][0] = acc0; s_red[warp][1] = acc1; s_red[warp][2] = acc2; s_red[warp][3] = acc3;
                s_red[warp][4] = acc4; s_red[warp][5] = acc5; s_red[warp][6] = accc;
                s_cnt[warp] = cnt;
            }
            const int need_H = __syncthreads_or((it == 0) || (vis_mask != prev_vis));
            prev_vis = vis_mask;
            if (need_H) {
                // H = fs^2 sum_visible (Sxx a a^T + Sxy (a b^T + b a^T) + Syy b b^T), lower triangle packed row-major,
                // assembled from the parked second moments (no image arithmetic here)
                double hacc[21];
#pragma unroll
                for (int i = 0; i < 21; ++i) hacc[i] = 0.0;
                int kk = 0;
                for (int f = tid; f < nfeat; f += NT, ++kk) {
                    if (!((vis_mask >> kk) & 1u)) continue;
                    const double P0 = s_P[f], P1 = s_P[NF + f], P2 = s_P[2 * NF + f];
                    const double zi = 1.0 / P2, zi2 = zi * zi;
                    const double av[6] = { -zi, 0.0, P0 * zi2, P1 * (P0 * zi2), -(1.0 + P0 * (P0 * zi2)), P1 * zi };
                    const double bv[6] = { 0.0, -zi, P1 * zi2, 1.0 + P1 * (P1 * zi2), -P0 * (P1 * zi2), -P0 * zi };
                    const double cxx = fs2 * s_S[f], cxy = fs2 * s_S[NF + f], cyy = fs2 * s_S[2 * NF + f];
#pragma unroll
                    for (int r = 0; r < 6; ++r)
#pragma unroll
                        for (int c = 0; c <= r; ++c)
                            hacc[r * (r + 1) / 2 + c] += cxx * (av[r] * av[c]) + cxy * (av[r] * bv[c] + bv[r] * av[c]) + cyy * (bv[r] * bv[c]);
                }
#pragma unroll
                for (int i = 0; i < 21; ++i) {
                    const double h = warp_sum(hacc[i]);
                    if (WPP > 1) { if (lane == 0) s_redH[warp][i] = h; }
                    else if (lane == 0) s_H[i] = h;
                }
                if (WPP > 1) __syncthreads();
            }
            if (warp == 0) {
                if (WPP > 1) {
                    if (need_H && lane < 21) {
                        double h = 0;
                        for (int w = 0; w < WPP; ++w) h += s_redH[w][lane];
                        s_H[lane] = h;
                    }
                    double red = 0;
                    if (lane < 7) for (int w = 0; w < WPP; ++w) red += s_red[w][lane];
                    int npts = 0;
                    for (int w = 0; w < WPP; ++w) npts += s_cnt[w];
                    acc0 = __shfl_sync(0xffffffffu, red, 0); acc1 = __shfl_sync(0xffffffffu, red, 1);
                    acc2 = __shfl_sync(0xffffffffu, red, 2); acc3 = __shfl_sync(0xffffffffu, red, 3);
                    acc4 = __shfl_sync(0xffffffffu, red, 4); acc5 = __shfl_sync(0xffffffffu, red, 5);
                    accc = __shfl_sync(0xffffffffu, red, 6);
                    cnt = npts;
                }
                __syncwarp();
                if (lane == 0) {
                    const double chi2New = accc / (double)(16 * cnt);                      // ref: :298 (NaN if nothing visible)
                    const double bvec[6] = { acc0, acc1, acc2, acc3, acc4, acc5 };
                    double x[6];
                    solve_and_update(s_H, s_F, need_H != 0, bvec, x);                      // ref: :318
                    int flags = 0;
                    bool stop = false;
                    if (isnan(x[0])) { stop = true; flags |= 4; }                          // ref: :321-326
                    if ((it > 0 && chi2New > s_chi2prev) || stop) {                        // ref: :328-332
#pragma unroll
                        for (int q = 0; q < 7; ++q) s_T[q] = s_Told[q];
                        flags |= 2;
                        stop = true;
                    } else {
                        double Tn[7];
                        pose_update(s_T, x, Tn);                                           // ref: :335
#pragma unroll
                        for (int q = 0; q < 7; ++q) { s_Told[q] = s_T[q]; s_T[q] = Tn[q]; } // ref: :336-337
                        s_chi2prev = chi2New;                                              // ref: :339
                        flags |= 1;
                        double mx = 0;
#pragma unroll
                        for (int q = 0; q < 6; ++q) mx = fmax(mx, fabs(x[q]));
                        if (mx <= 1e-8) { stop = true; flags |= 8; }                       // ref: :341
                    }
                    s_npts = cnt;
                    s_stop = stop ? 1 : 0;
                    if (a.log) {
                        const int n = s_nlog;
                        if (n < a.log_cap) {
                            dsdtm_iter_log* e = a.log + (size_t)pair * a.log_cap + n;
                            e->level = level; e->iter = it; e->n_pts = cnt; e->flags = flags; e->chi2 = chi2New;
#pragma unroll
                            for (int q = 0; q < 6; ++q) e->x[q] = x[q];
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
        if (tid < 7) s_Told[tid] = s_T[tid];
        if (tid == 0) { s_stop = 0; s_chi2prev = 0.0; }
        __syncthreads();
    }
    if (tid < 7) a.poses_out[7 * pair + tid] = s_T[tid];
    if (tid == 0) {
        a.n_tracked[pair] = s_npts;
        if (a.n_log) a.n_log[pair] = s_nlog;
    }
}

#ifndef DSDTM_SA_PAD
#define DSDTM_SA_PAD 0           // experiment: bytes of unused dynamic shared memory per CTA (with DSDTM_SA_CARVEOUT3 it caps the resident CTAs per SM)
#endif
int smem_bytes(int nf) { return ((DSDTM_SA_SGLOBAL ? 32 : 56) + 4 * NB_WORDS + 8 + 1 + (DSDTM_SA_COMPACT ? 2 : 0)) * nf + DSDTM_SA_PAD; }
int round_nf(int max_feats) { return (max_feats + 15) / 16 * 16; }

int smem_bytes_ws(int nf) { return (48 + 1) * nf; }

template <int WPP>
cudaError_t launch(const SaArgs& a, int n_pairs, cudaStream_t s)
{
    if (a.ws) sparse_align_ws_kernel<WPP><<<n_pairs, 32 * WPP, smem_bytes_ws(a.nf), s>>>(a);
    else sparse_align_kernel<WPP><<<n_pairs, 32 * WPP, smem_bytes(a.nf), s>>>(a);
    return cudaGetLastError();
}

}  // namespace

int sparse_align_smem_bytes(int nf) { return smem_bytes(nf); }
size_t sparse_align_ws_doubles(int max_feats) { return (size_t)48 * round_nf(max_feats); }

cudaError_t sparse_align_init(dsdtm_ctx* c)
{
    const int nf = round_nf(c->prm.max_feats);
    const int bytes = smem_bytes(nf);
    // the attribute belongs to the kernel, not to the context: several contexts in one process (different max_feats) must only ever
    // RAISE it, or the context with the larger feature table fails its next launch with "invalid argument"
    static std::mutex attr_mutex;
    static int attr_bytes = 0, attr_bytes_ws = 0;
    std::lock_guard<std::mutex> attr_lock(attr_mutex);
    attr_bytes = std::max(attr_bytes, bytes);
    cudaError_t e = cudaFuncSetAttribute(sparse_align_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, attr_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sparse_align_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, attr_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sparse_align_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, attr_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sparse_align_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, attr_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sparse_align_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, attr_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sparse_align_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, attr_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sparse_align_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, attr_bytes);
#ifdef DSDTM_SA_CARVEOUT
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sparse_align_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, DSDTM_SA_CARVEOUT);
#endif
#ifdef DSDTM_SA_CARVEOUT3
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sparse_align_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, DSDTM_SA_CARVEOUT3);
#endif
    if (e == cudaSuccess) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sparse_align_kernel<3>, 96, bytes) == cudaSuccess) c->sa_ctas_per_sm[3] = n;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sparse_align_kernel<4>, 128, bytes) == cudaSuccess) c->sa_ctas_per_sm[4] = n;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sparse_align_kernel<5>, 160, bytes) == cudaSuccess) c->sa_ctas_per_sm[5] = n;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sparse_align_kernel<10>, 320, bytes) == cudaSuccess) c->sa_ctas_per_sm[10] = n;
        (void)cudaGetLastError();
    }
    const int bw = attr_bytes_ws = std::max(attr_bytes_ws, smem_bytes_ws(nf));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sparse_align_ws_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bw);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sparse_align_ws_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bw);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sparse_align_ws_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bw);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sparse_align_ws_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bw);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sparse_align_ws_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, bw);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sparse_align_ws_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, bw);
    return e;
}

// Warps per pair. A CTA owns a pair, so the choice trades the latency of one pair (more warps: fewer rounds of the feature pass per
// iteration) against the pairs resident per SM (registers: 168 per thread). Measured per-pair latencies at full residency on B200
// (profiles/r2_sparse_align.md, 4096 pairs: time / waves): 3 warps (4 CTAs per SM) 195 us, 4 warps (3 per SM) 156 us, 5 warps (3 per
// SM) 152 us, 10 warps (1 per SM) 88 us. The batch runs in ceil(n / resident CTAs) waves, so the pick is the width with the smallest
// waves x latency: a lone pair gets ten warps, batches that fill the chip many times get three (highest throughput), and a batch of
// 512 pairs on 148 SMs -- one eighth of BASELINE configs[4] -- gets three as well (ONE wave of 592 slots; the round-1 rule picked four
// warps = 444 slots = two waves, 0.29 instead of 0.20 ms).
int sparse_align_pick_wpp(const dsdtm_ctx* c, int n_pairs)
{
    if (c->sa_wpp_override > 0) return c->sa_wpp_override;
    static const int widths[4] = { 3, 4, 5, 10 };
    static const double latency_us[4] = { 195.0, 156.0, 152.0, 88.0 };
    int best = 3;
    double best_cost = 1e300;
    for (int k = 0; k < 4; ++k) {
        const int per_sm = c->sa_ctas_per_sm[widths[k]] > 0 ? c->sa_ctas_per_sm[widths[k]] : 1;
        const long long slots = (long long)per_sm * c->sm_count;
        const double cost = (double)((n_pairs + slots - 1) / slots) * latency_us[k];
        if (cost < best_cost) { best_cost = cost; best = widths[k]; }
    }
    return best;
}

cudaError_t launch_sparse_align(dsdtm_ctx* c, int n_pairs, int feat_stride, int max_level, int min_level, int max_iters,
                                bool want_log, cudaStream_t s, int pair0, int n_pairs_total)
{
    SaArgs a;
    a.frames = c->frames_d; a.frame_stride = c->geo.frame_stride; a.geo = c->geo;
    a.ref_slots = c->ref_slots_d; a.cur_slots = c->cur_slots_d;
    a.feats = c->feats_d; a.feat_stride = feat_stride; a.n_feats = c->n_feats_d;
    a.centers = c->centers_d; a.poses_in = c->poses_in_d; a.poses_out = c->poses_out_d; a.n_tracked = c->n_tracked_d;
    a.log = want_log ? c->log_d : nullptr; a.n_log = want_log ? c->n_log_d : nullptr; a.log_cap = kLogCap;
    a.fx = c->cam.fx; a.fy = c->cam.fy; a.cx = c->cam.cx; a.cy = c->cam.cy; a.f = c->cam.f;
    a.max_level = max_level; a.min_level = min_level; a.max_iters = max_iters;
    a.nf = round_nf(c->prm.max_feats);
    a.ws = (c->sa_variant == 1) ? c->sa_ws_d : nullptr;
    a.variant = c->sa_variant;
    a.mom = c->sa_ws_d;
    a.pair0 = pair0;
    c->launches++;
    // warps-per-pair is chosen from the size of the WHOLE batch so that chunked (e2e) and single-launch runs reduce in the same order
    const int wpp = sparse_align_pick_wpp(c, n_pairs_total > 0 ? n_pairs_total : n_pairs);
    switch (wpp) {
    case 1: return launch<1>(a, n_pairs, s);
    case 2: return launch<2>(a, n_pairs, s);
    case 3: return launch<3>(a, n_pairs, s);
    case 4: return launch<4>(a, n_pairs, s);
    case 5: return launch<5>(a, n_pairs, s);
    case 6: if (!a.ws) { sparse_align_kernel<6><<<n_pairs, 192, smem_bytes(a.nf), s>>>(a); return cudaGetLastError(); } return launch<5>(a, n_pairs, s);
    default: return launch<10>(a, n_pairs, s);
    }
}

}  // namespace dsdtm
