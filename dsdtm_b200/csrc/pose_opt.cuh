// pose_opt.cuh -- motion-only bundle adjustment of one frame: what Optimizer::PoseOptimization
// (ref: src/Optimizer.cpp:20-101, include/Optimizer.h:129-258) asks ceres::Solve to do, as one self-contained routine.
//
// The reference's configuration: one pose block x = [t, log(R)] with PoseLocalParameterization (Plus = left multiplication by
// SE3(SO3::exp(d[3..5]), d[0..2]), identity Jacobian), constant map points, FullBA_Problem residuals
// r = (normal.xy / normal.z - proj(R P + t)) / 2^level with its analytic 2 x 6 Jacobian (which omits the 1 / 2^level factor:
// reproduced), CauchyLoss(1.0), DENSE_SCHUR, <= 100 iterations, Ceres defaults otherwise. For that configuration Ceres'
// trust-region minimizer is Levenberg-Marquardt on a 6 x 6 system:
//   * loss correction: rho''(s) <= 0 for Cauchy, so residual and Jacobian of a block are both scaled by sqrt(rho'(s)), cost = rho(s) / 2;
//   * Jacobi scaling s_c = 1 / (1 + ||J_c||), fixed at iteration 0;
//   * D^2 = clamp(diag(Js'Js), 1e-6, 1e32) / radius (recomputed after accepted steps only);
//   * step = -(Js'Js + D^2)^-1 Js'f ; model change = -(Js step)'(f + Js step / 2) ; candidate = Plus(x, step * s);
//   * stop on |x - cand| <= 1e-8 (|x| + 1e-8) or |cost change| <= 1e-6 cost (the candidate is NOT taken), on max |x - Plus(x, -g)| <= 1e-10
//     after an accepted step, on the iteration cap, on radius < 1e-32, after 5 consecutive invalid steps;
//   * accept when cost change / model change > 1e-3: radius /= max(1/3, 1 - (2q - 1)^3); otherwise radius /= 2, 4, 8, ...
//
// B200 mapping: the work per iteration is one pass over <= a few hundred observations producing 28 sums (cost, J'f, the 21
// unique entries of J'J) followed by a 6 x 6 factorisation: a dependent chain, latency-bound. Observations are strided over the
// threads that own the frame (staged once as structure-of-arrays in shared memory: observation, 2^-level, point), the 28 sums
// are reduced with a transposed warp reduction + broadcast (+ one shared-memory exchange when several warps own the frame) that
// leave identical bits in every thread, so every thread runs the scalar tail redundantly and no broadcast is needed, and the candidate pass
// accumulates J'J and J'f speculatively so that an accepted step costs no second pass. A sweep gives each frame ONE warp (the
// other resident warps hide its chain); a lone frame gets a CTA of eight warps (the pass is eight times shorter).
// Differences from Ceres' arithmetic are rounding only (documented in DESIGN.md): sums are per lane then tree instead of
// sequential, J'J is accumulated unscaled and scaled afterwards, the system is solved by Cholesky substitution instead of
// forming the inverse. The routine is __host__ __device__ (lane policy SerialLanes) so the CPU suite checks the same source
// against the oracle without a GPU.
#pragma once

#include <float.h>
#include <math.h>

#include "../../include/dsdtm_gpu.h"

#if defined(__CUDACC__)
#define DSDTM_PO_HD __host__ __device__ __forceinline__
#else
#define DSDTM_PO_HD inline
#endif

namespace dsdtm {

// Lane policies: how the observations of one frame are spread over threads and how the per-thread partial sums meet.
// sum_n leaves the SAME bits in every participating thread, so the scalar tail runs redundantly and control flow stays uniform.
struct SerialLanes {
    DSDTM_PO_HD int lane() const { return 0; }
    DSDTM_PO_HD int count() const { return 1; }
    template <int N> DSDTM_PO_HD void sum_n(double (&)[N]) const {}
    DSDTM_PO_HD void sync() const {}
};

#if defined(__CUDACC__)
// Sum N <= 32 per-lane values over the warp so that lane l ends up with the total of value l ("transposed" reduction): at the
// stage with offset o a lane keeps the half of its values whose index has bit o equal to its own lane bit and hands the other
// half to its partner, so 16 + 8 + 4 + 2 + 1 = 31 exchanges move everything -- instead of 5 exchanges for each of the N values
// in a plain butterfly (140 for N = 28). Lanes >= N end with the total of a zero column.
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
            if (i >= N) continue;                         // both halves are zero columns
            const double lo = w[i], hi = w[i + o];
            const double keep = up ? hi : lo;
            const double send = up ? lo : hi;
            w[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return w[0];
}

// one warp per frame (sweeps: the other resident warps hide this warp's dependent chain)
struct WarpLanes {
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int count() const { return 32; }
    template <int N> __device__ __forceinline__ void sum_n(double (&v)[N]) const
    {
        const double tot = warp_sum_transposed(v);
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = __shfl_sync(0xffffffffu, tot, i);
    }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};
// one CTA of kWarps warps per frame (a lone frame: the pass over the observations is kWarps times shorter). Per-warp transposed
// reductions, then lane l of every warp adds the kWarps warp sums of value l from shared memory in warp order and hands the
// total to its warp; two buffers alternate so that one __syncthreads per reduction suffices.
template <int kWarps, int kMaxN>
struct CtaLanes {
    double* buf;            // 2 * kWarps * kMaxN doubles of shared memory
    mutable int phase = 0;
    __device__ __forceinline__ explicit CtaLanes(double* b) : buf(b) {}
    __device__ __forceinline__ int lane() const { return threadIdx.x; }
    __device__ __forceinline__ int count() const { return 32 * kWarps; }
    template <int N> __device__ __forceinline__ void sum_n(double (&v)[N]) const
    {
        static_assert(N <= kMaxN && N <= 32, "reduction buffer too small");
        double tot = warp_sum_transposed(v);
        double* b = buf + phase * (kWarps * kMaxN);
        phase ^= 1;
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        if (l < N) b[w * kMaxN + l] = tot;
        __syncthreads();
        if (l < N) {
            tot = b[l];
#pragma unroll
            for (int k = 1; k < kWarps; ++k) tot += b[k * kMaxN + l];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = __shfl_sync(0xffffffffu, tot, i);
    }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
#endif

namespace po {

static constexpr double kSmallEps = 1e-10;   // Sophus SMALL_EPS

// The scalar tail is one dependent fp64 chain per iteration; a division costs ~10 dependent instructions, so denominators
// shared by several quotients are inverted once (1 ulp away from the reference's quotients; covered by the stated tolerance).
DSDTM_PO_HD double po_rsqrt(double v)
{
#if defined(__CUDA_ARCH__)
    return rsqrt(v);
#else
    return 1.0 / sqrt(v);
#endif
}
DSDTM_PO_HD void po_sincos(double a, double* s, double* c)
{
#if defined(__CUDA_ARCH__)
    sincos(a, s, c);
#else
    *s = sin(a); *c = cos(a);
#endif
}

// Sophus SO3::expAndTheta + the normalising SO3(Quaternion) constructor; q = {w, x, y, z}
DSDTM_PO_HD void so3_exp(const double* om, double* q)
{
    const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    const double half = 0.5 * theta;
    double imag, real, sn;
    po_sincos(half, &sn, &real);
    if (theta < kSmallEps) {
        const double t2 = theta * theta, t4 = t2 * t2;
        imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
    } else {
        imag = sn / theta;
    }
    double w = real, x = imag * om[0], y = imag * om[1], z = imag * om[2];
    const double rn = po_rsqrt(x * x + y * y + z * z + w * w);
    q[0] = w * rn; q[1] = x * rn; q[2] = y * rn; q[3] = z * rn;
}

// Sophus SO3::logAndTheta (atan form)
DSDTM_PO_HD void so3_log(const double* q, double* out)
{
    const double n = sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double w = q[0];
    double f;
    if (n < kSmallEps) f = 2. / w - 2. * (n * n) / (w * (w * w));
    else f = 2 * atan(n / w) / n;
    out[0] = f * q[1]; out[1] = f * q[2]; out[2] = f * q[3];
}

// Eigen QuaternionBase::_transformVector
DSDTM_PO_HD void qrot(const double* q, double v0, double v1, double v2, double& o0, double& o1, double& o2)
{
    double uv0 = q[2] * v2 - q[3] * v1, uv1 = q[3] * v0 - q[1] * v2, uv2 = q[1] * v1 - q[2] * v0;
    uv0 += uv0; uv1 += uv1; uv2 += uv2;
    const double c0 = q[2] * uv2 - q[3] * uv1, c1 = q[3] * uv0 - q[1] * uv2, c2 = q[1] * uv1 - q[2] * uv0;
    o0 = v0 + q[0] * uv0 + c0;
    o1 = v1 + q[0] * uv1 + c1;
    o2 = v2 + q[0] * uv2 + c2;
}

// PoseLocalParameterization::Plus: [t, log R] of SE3(exp(d.w), d.t) * SE3(exp(x.w), x.t)   (ref: include/Optimizer.h:220-236)
// qo = so3_exp(x + 3), computed once per pose by the caller; q_out = the normalised product quaternion, which the evaluation of
// the new pose uses directly: exp(log(q)) == +-q up to rounding, so re-deriving it from out[3..5] would only add a sincos.
DSDTM_PO_HD void pose_plus(const double* x, const double* qo, const double* d, double* out, double* q_out)
{
    double qd[4];
    so3_exp(d + 3, qd);
    double r0, r1, r2;
    qrot(qd, x[0], x[1], x[2], r0, r1, r2);
    out[0] = d[0] + r0; out[1] = d[1] + r1; out[2] = d[2] + r2;
    double q[4];
    q[0] = qd[0] * qo[0] - qd[1] * qo[1] - qd[2] * qo[2] - qd[3] * qo[3];
    q[1] = qd[0] * qo[1] + qd[1] * qo[0] + qd[2] * qo[3] - qd[3] * qo[2];
    q[2] = qd[0] * qo[2] + qd[2] * qo[0] + qd[3] * qo[1] - qd[1] * qo[3];
    q[3] = qd[0] * qo[3] + qd[3] * qo[0] + qd[1] * qo[2] - qd[2] * qo[1];
    const double rn = po_rsqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + q[0] * q[0]);
    q[0] *= rn; q[1] *= rn; q[2] *= rn; q[3] *= rn;
    so3_log(q, out + 3);
    q_out[0] = q[0]; q_out[1] = q[1]; q_out[2] = q[2]; q_out[3] = q[3];
}

struct Normal {     // cost, J'f and the upper triangle of J'J (loss-corrected, unscaled), row-major packed
    double cost;
    double g[6];
    double H[21];
};

DSDTM_PO_HD int hidx(int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); }   // i <= j

// One pass of ProgramEvaluator::Evaluate over the staged observations (soa = ox | oy | inv | px | py | pz, each n doubles).
// q = so3_exp(x + 3)
template <class Lanes>
DSDTM_PO_HD void evaluate(const Lanes& ln, int n, const double* soa, const double* x, const double* q, Normal& out)
{
    double acc[28];   // cost | J'f | upper triangle of J'J
#pragma unroll
    for (int i = 0; i < 28; ++i) acc[i] = 0.0;
    const double *ox = soa, *oy = soa + n, *inv = soa + 2 * n, *px = soa + 3 * n, *py = soa + 4 * n, *pz = soa + 5 * n;
    for (int k = ln.lane(); k < n; k += ln.count()) {
        double cx, cy, cz;
        qrot(q, px[k], py[k], pz[k], cx, cy, cz);
        cx += x[0]; cy += x[1]; cz += x[2];
        // FullBA_Problem::Evaluate (ref: include/Optimizer.h:141-205)
        const double z_inv = 1.0 / cz;
        const double r0 = (ox[k] - cx * z_inv) * inv[k];
        const double r1 = (oy[k] - cy * z_inv) * inv[k];
        const double z_inv2 = z_inv * z_inv;
        double j0[6], j1[6];
        j0[0] = -z_inv; j0[1] = 0.0; j0[2] = cx * z_inv2; j0[3] = cy * j0[2]; j0[4] = -(1.0 + cx * j0[2]); j0[5] = cy * z_inv;
        j1[0] = 0.0; j1[1] = -z_inv; j1[2] = cy * z_inv2; j1[3] = 1.0 + cy * j1[2]; j1[4] = -cx * j1[2]; j1[5] = -cx * z_inv;
        // CauchyLoss(1.0) + Corrector
        const double s = r0 * r0 + r1 * r1;
        const double sum = 1.0 + s;
        const double rho1 = fmax(DBL_MIN, 1.0 / sum);
        acc[0] += 0.5 * log(sum);
        const double w = sqrt(rho1);
        const double f0 = r0 * w, f1 = r1 * w;
#pragma unroll
        for (int c = 0; c < 6; ++c) { j0[c] *= w; j1[c] *= w; }
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[1 + c] += j0[c] * f0 + j1[c] * f1;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int b = a; b < 6; ++b) acc[7 + hidx(a, b)] += j0[a] * j0[b] + j1[a] * j1[b];
    }
    ln.sum_n(acc);
    out.cost = acc[0];
#pragma unroll
    for (int c = 0; c < 6; ++c) out.g[c] = acc[1 + c];
#pragma unroll
    for (int i = 0; i < 21; ++i) out.H[i] = acc[7 + i];
}

// Cholesky (lower) solve of the symmetric positive definite M (upper triangle packed) ; false when a pivot is not positive
DSDTM_PO_HD bool chol6_solve(const double* Mu, const double* b, double* y)
{
    double L[6][6], rd[6];   // rd[k] = 1 / L[k][k]
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double d = Mu[hidx(k, k)];
#pragma unroll
        for (int j = 0; j < k; ++j) d -= L[k][j] * L[k][j];
        ok = ok && (d > 0.0);          // no early exit: the chain stays one basic block, NaNs of a failed pivot are discarded by the caller
        rd[k] = po_rsqrt(d);
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            double s = Mu[hidx(k, i)];
#pragma unroll
            for (int j = 0; j < k; ++j) s -= L[i][j] * L[k][j];
            L[i][k] = s * rd[k];
        }
    }
    double v[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = b[i];
#pragma unroll
        for (int j = 0; j < i; ++j) s -= L[i][j] * v[j];
        v[i] = s * rd[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double s = v[i];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) s -= L[j][i] * v[j];
        v[i] = s * rd[i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = v[i];
    return ok;
}

DSDTM_PO_HD bool finite_d(double v) { return fabs(v) <= DBL_MAX; }   // false for NaN and infinities

}  // namespace po

// Whole solve of one frame. obs[n] (host or device global memory), soa = 6 * n doubles of scratch (shared memory on the device),
// res_norm may be null. Every lane of `ln` calls this with identical arguments; lane 0's summary / pose writes are the result.
template <class Lanes>
DSDTM_PO_HD void pose_optimize(const Lanes& ln, int n, const dsdtm_ba_obs* obs, double* soa, const double* pose_in, int max_iters,
                               double* pose_out, double* res_norm, dsdtm_ba_summary* summary)
{
    using namespace po;
    const double kFunctionTol = 1e-6, kGradientTol = 1e-10, kParameterTol = 1e-8, kMinRelDecrease = 1e-3;
    const double kMinDiag = 1e-6, kMaxDiag = 1e32, kMaxRadius = 1e16, kMinRadius = 1e-32;
    const int kMaxInvalid = 5;

    // stage: observation = mNormal.xy / mNormal.z, 1 / (1 << level), the map point
    for (int k = ln.lane(); k < n; k += ln.count()) {
        const dsdtm_ba_obs o = obs[k];
        soa[k] = o.normal[0] / o.normal[2];
        soa[n + k] = o.normal[1] / o.normal[2];
        soa[2 * n + k] = 1.0 / (double)(1 << o.level);
        soa[3 * n + k] = o.point_w[0];
        soa[4 * n + k] = o.point_w[1];
        soa[5 * n + k] = o.point_w[2];
    }
    ln.sync();

    // ref: src/Optimizer.cpp:34-36
    double x[6] = { pose_in[4], pose_in[5], pose_in[6], 0, 0, 0 };
    so3_log(pose_in, x + 3);

    double qx[4] = { 1, 0, 0, 0 };   // so3_exp(x + 3) of the current pose
    int iterations = 0, n_successful = 0, term = DSDTM_BA_NO_RESIDUALS;
    double initial_cost = 0.0, final_cost = 0.0;

    if (n > 0) {
        // One loop, rotated so that the pass over the observations appears ONCE in the code (iteration zero evaluates the start
        // pose, every later trip the candidate): the pass is the bulk of the instructions and a second inlined copy costs
        // instruction-cache misses on a lone CTA.
        Normal cur;
        double scale[6] = { 1, 1, 1, 1, 1, 1 };
        double radius = 1e4, decrease_factor = 2.0;
        bool reuse_diagonal = false, last_successful = true, first = true;
        double diagonal[6] = { 0, 0, 0, 0, 0, 0 };
        int n_invalid = 0;
        double x_norm = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3] + x[4] * x[4] + x[5] * x[5]);
        double model_cost_change = 0.0;
        double cand[6], qc[4];
        so3_exp(x + 3, qx);
#pragma unroll
        for (int c = 0; c < 6; ++c) cand[c] = x[c];
#pragma unroll
        for (int c = 0; c < 4; ++c) qc[c] = qx[c];
        for (;;) {
            Normal nxt;
            evaluate(ln, n, soa, cand, qc, nxt);
            if (first) {
                // IterationZero
                first = false;
                cur = nxt;
                initial_cost = final_cost = cur.cost;
                if (!finite_d(cur.cost)) { term = DSDTM_BA_FAILURE; break; }
#pragma unroll
                for (int c = 0; c < 6; ++c) scale[c] = 1.0 / (1.0 + sqrt(cur.H[hidx(c, c)]));
            } else {
                const bool cand_ok = finite_d(nxt.cost);
                const double cand_cost = cand_ok ? nxt.cost : DBL_MAX;
                // ParameterToleranceReached / FunctionToleranceReached: the candidate is NOT taken when they fire
                double step_norm = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) step_norm += (x[c] - cand[c]) * (x[c] - cand[c]);
                step_norm = sqrt(step_norm);
                if (step_norm <= kParameterTol * (x_norm + kParameterTol)) { term = DSDTM_BA_PARAMETER_TOL; break; }
                const double cost_change = cur.cost - cand_cost;
                if (fabs(cost_change) <= kFunctionTol * cur.cost) { term = DSDTM_BA_FUNCTION_TOL; break; }

                const double relative_decrease = cost_change / model_cost_change;
                if (relative_decrease > kMinRelDecrease) {   // HandleSuccessfulStep
#pragma unroll
                    for (int c = 0; c < 6; ++c) x[c] = cand[c];
#pragma unroll
                    for (int c = 0; c < 4; ++c) qx[c] = qc[c];
                    x_norm = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3] + x[4] * x[4] + x[5] * x[5]);
                    cur = nxt;
                    final_cost = cur.cost;
                    ++n_successful;
                    last_successful = true;
                    const double q = 2.0 * relative_decrease - 1.0;
                    radius = radius / fmax(1.0 / 3.0, 1.0 - q * q * q);
                    radius = fmin(kMaxRadius, radius);
                    decrease_factor = 2.0;
                    reuse_diagonal = false;
                } else {
                    radius = radius / decrease_factor; decrease_factor *= 2.0;
                }
            }

            // the next trust-region step; invalid steps shrink the radius and try again without a pass over the observations
            bool stop = false;
            double step[6];
            for (;;) {
                // FinalizeIterationAndCheckIfMinimizerCanContinue
                if (iterations >= max_iters) { term = DSDTM_BA_NO_CONVERGENCE; stop = true; break; }
                // GradientToleranceReached: max |x - Plus(x, -g)| <= 1e-10. The exact test costs a whole Plus (sincos, atan, rsqrt chain),
                // so it is skipped when it cannot fire: with d = -g, the rotation part of x - Plus(x, d) is J_l^-1(x_w) d_w + O(d^2) and
                // every singular value of J_l^-1 is >= 1, the translation part is -d_t + (I - R(d_w)) x_t with |(I - R) x_t| <= |d_w| |x_t|;
                // for 1e-5 (1 + |x_t|) < max |g| < 1 at least one component therefore exceeds 1e-7.
                double gabs = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) gabs = fmax(gabs, fabs(cur.g[c]));
                const bool g_clear = gabs < 1.0 && gabs > 1e-5 * (1.0 + sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]));
                if (last_successful && !g_clear) {
                    double ng[6], proj[6], qp[4];
#pragma unroll
                    for (int c = 0; c < 6; ++c) ng[c] = -cur.g[c];
                    pose_plus(x, qx, ng, proj, qp);
                    double gmax = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) gmax = fmax(gmax, fabs(x[c] - proj[c]));
                    if (gmax <= kGradientTol) { term = DSDTM_BA_GRADIENT_TOL; stop = true; break; }
                }
                if (radius < kMinRadius) { term = DSDTM_BA_MIN_RADIUS; stop = true; break; }
                ++iterations;
                last_successful = false;

                // LevenbergMarquardtStrategy::ComputeStep on the column-scaled system
                double M[21], gs[6];
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    gs[a] = cur.g[a] * scale[a];
#pragma unroll
                    for (int b = a; b < 6; ++b) M[hidx(a, b)] = cur.H[hidx(a, b)] * scale[a] * scale[b];
                }
                if (!reuse_diagonal) {
#pragma unroll
                    for (int c = 0; c < 6; ++c) diagonal[c] = fmin(fmax(M[hidx(c, c)], kMinDiag), kMaxDiag);
                }
                reuse_diagonal = true;
                double A[21];
#pragma unroll
                for (int i = 0; i < 21; ++i) A[i] = M[i];
#pragma unroll
                for (int c = 0; c < 6; ++c) A[hidx(c, c)] += diagonal[c] / radius;
                bool valid = chol6_solve(A, gs, step);
#pragma unroll
                for (int c = 0; c < 6; ++c) { valid = valid && finite_d(step[c]); step[c] = -step[c]; }
                model_cost_change = 0.0;
                if (valid) {
                    // -(Js step)'(f + Js step / 2) = -(step'gs + step'(Js'Js)step / 2)
                    double sg = 0.0, Hs_step_dot = 0.0;
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        sg += step[a] * gs[a];
                        double row = 0.0;
#pragma unroll
                        for (int b = 0; b < 6; ++b) row += M[a <= b ? hidx(a, b) : hidx(b, a)] * step[b];
                        Hs_step_dot += step[a] * row;
                    }
                    model_cost_change = -(sg + 0.5 * Hs_step_dot);
                    valid = model_cost_change > 0.0;
                }
                if (valid) { n_invalid = 0; break; }
                // HandleInvalidStep
                if (++n_invalid >= kMaxInvalid) { term = DSDTM_BA_FAILURE; stop = true; break; }
                radius = radius / decrease_factor; decrease_factor *= 2.0;
            }
            if (stop) break;
            double delta[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) delta[c] = step[c] * scale[c];
            pose_plus(x, qx, delta, cand, qc);
        }
    }

    // ref: src/Optimizer.cpp:79 and :298-318 (raw residual norm of every block at the final parameters)
    double q[4];
    so3_exp(x + 3, q);   // ref: src/Optimizer.cpp:79 (also when n == 0: the pose handed back is exp(log(R)) as in the reference)
    if (res_norm) {
        const double *ox = soa, *oy = soa + n, *inv = soa + 2 * n, *px = soa + 3 * n, *py = soa + 4 * n, *pz = soa + 5 * n;
        for (int k = ln.lane(); k < n; k += ln.count()) {
            double cx, cy, cz;
            qrot(q, px[k], py[k], pz[k], cx, cy, cz);
            cx += x[0]; cy += x[1]; cz += x[2];
            const double r0 = (ox[k] - cx / cz) * inv[k];
            const double r1 = (oy[k] - cy / cz) * inv[k];
            res_norm[k] = sqrt(r0 * r0 + r1 * r1);
        }
    }
    if (ln.lane() == 0) {
        pose_out[0] = q[0]; pose_out[1] = q[1]; pose_out[2] = q[2]; pose_out[3] = q[3];
        pose_out[4] = x[0]; pose_out[5] = x[1]; pose_out[6] = x[2];
        if (summary) {
            summary->iterations = iterations;
            summary->termination = term;
            summary->n_successful = n_successful;
            summary->n_obs = n;
            summary->initial_cost = initial_cost;
            summary->final_cost = final_cost;
        }
    }
}

}  // namespace dsdtm
