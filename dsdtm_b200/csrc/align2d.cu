// align2d.cu -- kernel (d): batched 2-D inverse-compositional patch alignment, one WARP per 8x8 feature patch,
// and the affine patch warp that feeds it.
//
// align2d_kernel replaces Feature_Alignment::Align2DGaussNewton (ref: src/Feature_alignment.cpp:318-417), fp32 like the
// reference: H = sum J J^T over the 8x8 interior of the 10x10 bordered patch (exact in fp32: quarter-integers < 2^22),
// Hinv by the 3x3 cofactor formula (Eigen Matrix3f::inverse), then <= max_iters iterations of bilinear sampling,
// Jres accumulation and the (u, v, mean) update; converged iff du^2 + dv^2 < 0.03^2.
// Lane l owns pixels l and l+32 of the patch; the three Jres sums are reduced with a warp xor-shuffle butterfly (bitwise
// identical in every lane => uniform control flow). The reference sums the 64 terms sequentially in fp32, the butterfly
// is a tree: documented tolerance 1e-3 px on the refined position.
// Q4 (ref: :367-368 uses '>' where SVO uses '>='): with u_r == cols-4 or v_r == rows-4 the reference reads one byte past
// the row / image. We reproduce the linear addressing (next row's first pixel) and read bytes past the end of the level
// image as 0, identically in the oracle.
//
// warp_affine_kernel replaces Feature_Alignment::WarpAffine + GetPatchNoBoarder (ref: :206-275), one warp per candidate,
// fp32 non-contracted with the reference's operation order => bit-exact 10x10 uchar patches (incl. the integer-division
// quirk Q3: the sampling grid collapses to the reference pixel for search levels >= 1).
#include "ctx.cuh"

namespace dsdtm {

namespace {

struct A2dArgs {
    const uint8_t* frames; unsigned frame_stride; LevelGeom geo;
    const int* patch_slot;      // frame slot per patch
    const int* patch_level;     // < 0 = unused entry
    const uint8_t* patch10;     // n x 100
    const double* px_in;        // n x 2 start positions (level coordinates)
    double* px;                 // n x 2 refined positions
    uint8_t* conv;              // n
    int n, max_iters, patch0;
};

__device__ __forceinline__ float wsum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int A2D_WARPS = 8;

__global__ void __launch_bounds__(A2D_WARPS * 32) align2d_kernel(const A2dArgs a)
{
    __shared__ uint8_t s_patch[A2D_WARPS][104];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = a.patch0 + blockIdx.x * A2D_WARPS + warp;
    if (i >= a.patch0 + a.n) return;
    const int L = a.patch_level[i];
    if (L < 0) { if (lane == 0) { a.conv[i] = 0; a.px[2 * i] = a.px_in[2 * i]; a.px[2 * i + 1] = a.px_in[2 * i + 1]; } return; }

    // stage the 10x10 bordered patch (100 bytes = 25 words; patch10 rows are 4-byte aligned because 100 % 4 == 0)
    if (lane < 25) reinterpret_cast<uint32_t*>(s_patch[warp])[lane] = __ldg(reinterpret_cast<const uint32_t*>(a.patch10 + (size_t)i * 100) + lane);
    __syncwarp();

    // lane owns pixels e = lane and lane + 32 : row = e / 8, col = e % 8 of the 8x8 interior
    float rdx[2], rdy[2], rref[2];
    float h00 = 0, h01 = 0, h02 = 0, h11 = 0, h12 = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int e = lane + 32 * k, r = e >> 3, c = e & 7;
        const uint8_t* it = s_patch[warp] + (r + 1) * 10 + 1 + c;
        rdx[k] = 0.5f * (float)((int)it[1] - (int)it[-1]);          // ref: :336 (exact half-integers)
        rdy[k] = 0.5f * (float)((int)it[10] - (int)it[-10]);        // ref: :337
        rref[k] = (float)it[0];
        h00 += rdx[k] * rdx[k]; h01 += rdx[k] * rdy[k]; h02 += rdx[k];
        h11 += rdy[k] * rdy[k]; h12 += rdy[k];
    }
    // all sums are exact in fp32 in any order (multiples of 1/4 below 2^22)
    h00 = wsum(h00); h01 = wsum(h01); h02 = wsum(h02); h11 = wsum(h11); h12 = wsum(h12);
    const float h22 = 64.0f;
    // Eigen 3x3 inverse: cofactor(i,j) = m(i1,j1) m(i2,j2) - m(i1,j2) m(i2,j1); det = sum_i cof(i,0) m(i,0)
    float Hinv[9];
    {
        const float m[9] = { h00, h01, h02, h01, h11, h12, h02, h12, h22 };
#define M(r, c) m[(r) * 3 + (c)]
#define COF(i, j) __fsub_rn(__fmul_rn(M(((i) + 1) % 3, ((j) + 1) % 3), M(((i) + 2) % 3, ((j) + 2) % 3)), \
                            __fmul_rn(M(((i) + 1) % 3, ((j) + 2) % 3), M(((i) + 2) % 3, ((j) + 1) % 3)))
        const float c00 = COF(0, 0), c10 = COF(1, 0), c20 = COF(2, 0);
        const float det = __fadd_rn(__fadd_rn(__fmul_rn(c00, M(0, 0)), __fmul_rn(c10, M(1, 0))), __fmul_rn(c20, M(2, 0)));
        const float invdet = __fdiv_rn(1.0f, det);
        Hinv[0] = __fmul_rn(c00, invdet); Hinv[1] = __fmul_rn(c10, invdet); Hinv[2] = __fmul_rn(c20, invdet);
        Hinv[3] = __fmul_rn(COF(0, 1), invdet); Hinv[4] = __fmul_rn(COF(1, 1), invdet); Hinv[5] = __fmul_rn(COF(2, 1), invdet);
        Hinv[6] = __fmul_rn(COF(0, 2), invdet); Hinv[7] = __fmul_rn(COF(1, 2), invdet); Hinv[8] = __fmul_rn(COF(2, 2), invdet);
#undef COF
#undef M
    }

    const int cols = a.geo.w[L], rows = a.geo.h[L];
    const uint8_t* __restrict__ img = a.frames + (size_t)a.patch_slot[i] * a.frame_stride + a.geo.off[L];
    const unsigned img_bytes = (unsigned)cols * (unsigned)rows;

    float u = (float)a.px_in[2 * i], v = (float)a.px_in[2 * i + 1];      // ref: :349-350
    float mean_diff = 0.f;
    const float min_update_squared = (float)(0.03 * 0.03);          // ref: :352
    bool converged = false;
    for (int it = 0; it < a.max_iters; ++it) {
        const float uf = floorf(u), vf = floorf(v);
        // ref: :367-369 (Q4 '>'); written on floats so that NaN / huge values break like the reference
        if (!(uf >= 4.f && vf >= 4.f && uf <= (float)(cols - 4) && vf <= (float)(rows - 4))) break;
        const int u_r = (int)uf, v_r = (int)vf;
        const float sx = u - uf, sy = v - vf;
        const float wTL = (float)((1.0 - (double)sx) * (1.0 - (double)sy));   // ref: :373 (double arithmetic, narrowed)
        const float wTR = __fmul_rn(sx, 1.0f - sy);                          // ref: :374 (float arithmetic)
        const float wBL = (float)((1.0 - (double)sx) * (double)sy);          // ref: :375
        const float wBR = __fmul_rn(sx, sy);                                 // ref: :376
        float j0 = 0.f, j1 = 0.f, j2 = 0.f;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int e = lane + 32 * k, r = e >> 3, c = e & 7;
            const unsigned o = (unsigned)(v_r + r - 4) * (unsigned)cols + (unsigned)(u_r - 4 + c);
            const unsigned o2 = o + (unsigned)cols;
            const float i00 = (float)(o < img_bytes ? __ldg(img + o) : 0);
            const float i01 = (float)(o + 1 < img_bytes ? __ldg(img + o + 1) : 0);
            const float i10 = (float)(o2 < img_bytes ? __ldg(img + o2) : 0);
            const float i11 = (float)(o2 + 1 < img_bytes ? __ldg(img + o2 + 1) : 0);
            const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(wTL, i00), __fmul_rn(wTR, i01)), __fmul_rn(wBL, i10)), __fmul_rn(wBR, i11));   // ref: :386
            const float res = __fadd_rn(__fsub_rn(s, rref[k]), mean_diff);                                                                  // ref: :387
            j0 = __fsub_rn(j0, __fmul_rn(res, rdx[k]));
            j1 = __fsub_rn(j1, __fmul_rn(res, rdy[k]));
            j2 = __fsub_rn(j2, res);
        }
        j0 = wsum(j0); j1 = wsum(j1); j2 = wsum(j2);
        // ref: :395 tUpdate = Hinv * Jres (row-wise, left to right)
        const float d0 = __fadd_rn(__fadd_rn(__fmul_rn(Hinv[0], j0), __fmul_rn(Hinv[1], j1)), __fmul_rn(Hinv[2], j2));
        const float d1 = __fadd_rn(__fadd_rn(__fmul_rn(Hinv[3], j0), __fmul_rn(Hinv[4], j1)), __fmul_rn(Hinv[5], j2));
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(Hinv[6], j0), __fmul_rn(Hinv[7], j1)), __fmul_rn(Hinv[8], j2));
        u = __fadd_rn(u, d0); v = __fadd_rn(v, d1); mean_diff = __fadd_rn(mean_diff, d2);
        if (__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)) < min_update_squared) { converged = true; break; }   // ref: :400
    }
    if (lane == 0) {
        a.px[2 * i] = (double)u; a.px[2 * i + 1] = (double)v;     // ref: :414
        a.conv[i] = converged ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct WaArgs {
    const uint8_t* frames; unsigned frame_stride; LevelGeom geo;
    const double* A;       // n x 4 row-major
    const float* ref_px;   // n x 2
    const int* meta;       // n x 3 : slot, ref_level, search_level
    uint8_t* out;          // n x 100
    int n;
};

__global__ void __launch_bounds__(256) warp_affine_kernel(const WaArgs a)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (i >= a.n) return;
    const int slot = a.meta[3 * i], rl = a.meta[3 * i + 1], sl = a.meta[3 * i + 2];
    const double A00 = a.A[4 * i], A01 = a.A[4 * i + 1], A10 = a.A[4 * i + 2], A11 = a.A[4 * i + 3];
    // Eigen 2x2 inverse in double, then cast<float>() (ref: :211)
    const double det = __dsub_rn(__dmul_rn(A00, A11), __dmul_rn(A10, A01));
    const double invdet = __ddiv_rn(1.0, det);
    const float a00 = (float)__dmul_rn(A11, invdet), a01 = (float)__dmul_rn(-A01, invdet);
    const float a10 = (float)__dmul_rn(-A10, invdet), a11 = (float)__dmul_rn(A00, invdet);
    const float rx = __fdiv_rn(a.ref_px[2 * i], (float)(1 << rl)), ry = __fdiv_rn(a.ref_px[2 * i + 1], (float)(1 << rl));   // ref: :215-216
    const float kf = (float)(1 / (1 << sl));                                                                                 // ref: :231 (Q3)
    const int w = a.geo.w[rl], h = a.geo.h[rl];
    const uint8_t* __restrict__ img = a.frames + (size_t)slot * a.frame_stride + a.geo.off[rl];
    for (int j = lane; j < 100; j += 32) {
        const float gx = (float)(j % 10 - 5), gy = (float)(j / 10 - 5);
        float wx = __fmul_rn(__fadd_rn(__fmul_rn(a00, gx), __fmul_rn(a01, gy)), kf);
        float wy = __fmul_rn(__fadd_rn(__fmul_rn(a10, gx), __fmul_rn(a11, gy)), kf);
        wx = __fadd_rn(wx, rx); wy = __fadd_rn(wy, ry);
        uint8_t o = 0;
        if (!(wx < 0 || wy < 0 || wx > (float)(w - 1) || wy > (float)(h - 1))) {       // ref: :249
            const int fx = (int)floor((double)wx), fy = (int)floor((double)wy);
            const float sx = __fsub_rn(wx, (float)fx), sy = __fsub_rn(wy, (float)fy);
            const float ox = __fsub_rn(1.0f, sx), oy = __fsub_rn(1.0f, sy);
            const float W00 = __fmul_rn(ox, oy), W01 = __fmul_rn(ox, sy), W10 = __fmul_rn(sx, oy);
            const float W11 = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, W00), W01), W10);      // ref: :244
            const bool xin = fx + 1 < w, yin = fy + 1 < h;
            const int p00 = __ldg(img + (size_t)fy * w + fx);
            const int p01 = yin ? __ldg(img + (size_t)(fy + 1) * w + fx) : 0;
            const int p10 = xin ? __ldg(img + (size_t)fy * w + fx + 1) : 0;
            const int p11 = (xin && yin) ? __ldg(img + (size_t)(fy + 1) * w + fx + 1) : 0;
            const float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(W00, (float)p00), __fmul_rn(W01, (float)p01)), __fmul_rn(W10, (float)p10)),
                                      __fmul_rn(W11, (float)p11));                                                               // ref: :254-255
            o = (uint8_t)(int)v;   // truncating store
        }
        a.out[(size_t)i * 100 + j] = o;
    }
}

}  // namespace

cudaError_t launch_align2d(dsdtm_ctx* c, int n_patches, int max_iters, cudaStream_t s, int patch0)
{
    A2dArgs a;
    a.frames = c->frames_d; a.frame_stride = c->geo.frame_stride; a.geo = c->geo;
    a.patch_slot = c->patch_slot_d; a.patch_level = c->patch_level_d; a.patch10 = c->patches_d;
    a.px_in = c->patch_px_in_d; a.px = c->patch_px_d; a.conv = c->patch_conv_d; a.n = n_patches; a.max_iters = max_iters; a.patch0 = patch0;
    align2d_kernel<<<(n_patches + A2D_WARPS - 1) / A2D_WARPS, A2D_WARPS * 32, 0, s>>>(a);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_warp_affine(dsdtm_ctx* c, int n, uint8_t* out_d, cudaStream_t s)
{
    WaArgs a;
    a.frames = c->frames_d; a.frame_stride = c->geo.frame_stride; a.geo = c->geo;
    a.A = c->wa_A_d; a.ref_px = c->wa_px_d; a.meta = c->wa_meta_d; a.out = out_d; a.n = n;
    warp_affine_kernel<<<(n + 7) / 8, 256, 0, s>>>(a);
    c->launches++;
    return cudaGetLastError();
}

}  // namespace dsdtm
