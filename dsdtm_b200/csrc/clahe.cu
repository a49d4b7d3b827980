// clahe.cu -- SURVEY 8f-4 ingest: cv::createCLAHE(clip, tiles)->apply(img) on CV_8UC1, the preprocessing the reference's
// drivers run in front of the Frame constructor (ref: Test/test_Feature_detection.cpp:85-86, Test/test_Euroc.cpp:64,
// Test/test_Optimizer.cpp:75; "the single image cost almost 10ms on reading and clahe"), fused with the pyramid build.
// OpenCV's algorithm (imgproc/src/clahe.cpp) for sizes divisible by the tile grid:
//   clahe_lut_kernel   : one CTA per tile: 256-bin histogram in shared memory, clip + redistribution (integer, exact), inclusive
//                        scan, lut = saturate(cvRound(sum * 255/tile_area)) (one fp32 multiply, round-half-even).
//   clahe_apply_kernel : per pixel the four surrounding tile LUTs blended bilinearly in non-contracted fp32 in OpenCV's operation
//                        order ((l11*xa1 + l12*xa)*ya1 + (l21*xa1 + l22*xa)*ya), cvRound, 4 pixels per thread. The <= 3 LUT rows
//                        a band of 8 image rows touches are staged in shared memory (6 KB for an 8-wide grid).
// Bit-exact against cv2 4.13 (tests/golden/clahe_cv2.npz). HBM-bound: two passes over the image + one write = 3 B / pixel.
#include "ctx.cuh"

namespace dsdtm {

namespace {

struct ClaheArgs {
    const uint8_t* src;         // n dense w*h images
    uint8_t* frames; unsigned frame_stride; int first_slot; unsigned dst_off;
    int w, h, tiles_x, tiles_y, tw, th;
    int clip; float lut_scale, inv_tw, inv_th;
    uint8_t* lut;               // n * tiles_x*tiles_y * 256
};

__global__ void __launch_bounds__(256) clahe_lut_kernel(const ClaheArgs a)
{
    __shared__ int s_hw[8][256];        // one private histogram per warp: no contention between warps
    __shared__ int s_warp[8];
    const int t = threadIdx.x, tile = blockIdx.x, frame = blockIdx.y;
    const int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
#pragma unroll
    for (int k = 0; k < 8; ++k) s_hw[k][t] = 0;
    int* s_hist = s_hw[t >> 5];
    __syncthreads();
    const uint8_t* __restrict__ base = a.src + (size_t)frame * a.w * a.h + (size_t)(ty * a.th) * a.w + tx * a.tw;
    if (((a.tw | a.w) & 3) == 0) {                                  // 4 pixels per load
        // thread -> (row r0 = t / wq, word xq = t % wq), then down the tile in steps of 256 / wq rows: no division in the loop
        const int wq = a.tw >> 2;
        if (wq <= 256) {
            const int rstep = 256 / wq, r0 = t / wq, xq = t - r0 * wq;
            if (r0 < rstep)
                for (int y = r0; y < a.th; y += rstep) {
                    const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)y * a.w) + xq);
                    atomicAdd(&s_hist[v & 0xFF], 1); atomicAdd(&s_hist[(v >> 8) & 0xFF], 1);
                    atomicAdd(&s_hist[(v >> 16) & 0xFF], 1); atomicAdd(&s_hist[v >> 24], 1);
                }
        } else {
            for (int y = 0; y < a.th; ++y)
                for (int xq = t; xq < wq; xq += 256) {
                    const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)y * a.w) + xq);
                    atomicAdd(&s_hist[v & 0xFF], 1); atomicAdd(&s_hist[(v >> 8) & 0xFF], 1);
                    atomicAdd(&s_hist[(v >> 16) & 0xFF], 1); atomicAdd(&s_hist[v >> 24], 1);
                }
        }
    } else {
        const int np = a.tw * a.th;
        for (int i = t; i < np; i += 256) {
            const int y = i / a.tw, x = i - y * a.tw;
            atomicAdd(&s_hist[__ldg(base + (size_t)y * a.w + x)], 1);
        }
    }
    __syncthreads();
    int hv = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) hv += s_hw[k][t];
    if (a.clip > 0) {                                               // clip and redistribute (clahe.cpp CLAHE_CalcLut_Body)
        int over = max(hv - a.clip, 0);
        hv = min(hv, a.clip);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) over += __shfl_xor_sync(0xffffffffu, over, o);
        if ((t & 31) == 0) s_warp[t >> 5] = over;
        __syncthreads();
        int clipped = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) clipped += s_warp[k];
        __syncthreads();
        const int batch = clipped >> 8;
        const int residual = clipped - (batch << 8);
        hv += batch;
        if (residual != 0) {
            const int step = max(256 / residual, 1);                // for (i = 0; i < 256 && residual > 0; i += step, residual--) hist[i]++
            if (t % step == 0 && t / step < residual) hv++;
        }
    }
    // inclusive scan over the 256 bins
    int sum = hv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, sum, o); if ((t & 31) >= o) sum += v; }
    if ((t & 31) == 31) s_warp[t >> 5] = sum;
    __syncthreads();
    for (int k = 0; k < (t >> 5); ++k) sum += s_warp[k];
    const int r = __float2int_rn(__fmul_rn((float)sum, a.lut_scale));   // saturate_cast<uchar>(float) = cvRound + clamp
    a.lut[((size_t)frame * a.tiles_x * a.tiles_y + tile) * 256 + t] = (uint8_t)min(max(r, 0), 255);
}

constexpr int CL_ROWS = 8;          // image rows per CTA of clahe_apply_kernel
constexpr int CL_MAX_TX = 16;       // widest supported tile grid

// ind1_p / ind2_p / xa_p / xa1_p of CLAHE_Interpolation_Body for one column
struct ClaheCol { int i1, i2; float xa, xa1; };
__device__ __forceinline__ ClaheCol clahe_col(int x, float inv_tw, int tiles_x)
{
    const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f);
    const float fl = floorf(txf);
    ClaheCol c;
    c.xa = __fsub_rn(txf, fl); c.xa1 = __fsub_rn(1.0f, c.xa);
    c.i1 = max((int)fl, 0) * 256; c.i2 = min((int)fl + 1, tiles_x - 1) * 256;
    return c;
}
__device__ __forceinline__ uint32_t clahe_px(const uint8_t* l1, const uint8_t* l2, const ClaheCol& c, int v, float ya, float ya1)
{
    const float top = __fadd_rn(__fmul_rn((float)l1[c.i1 + v], c.xa1), __fmul_rn((float)l1[c.i2 + v], c.xa));
    const float bot = __fadd_rn(__fmul_rn((float)l2[c.i1 + v], c.xa1), __fmul_rn((float)l2[c.i2 + v], c.xa));
    const int r = __float2int_rn(__fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya)));
    return (uint32_t)min(max(r, 0), 255);
}

__global__ void __launch_bounds__(256) clahe_apply_kernel(const ClaheArgs a)
{
    __shared__ __align__(16) uint8_t s_lut[3][CL_MAX_TX * 256];
    const int t = threadIdx.x, frame = blockIdx.y;
    const int y0 = blockIdx.x * CL_ROWS, y1 = min(y0 + CL_ROWS, a.h);
    // LUT rows touched by rows y0 .. y1-1: ty1(y0) .. ty2(y1-1) (at most 3 when CL_ROWS <= tile height)
    const int r_lo = max((int)floorf(__fsub_rn(__fmul_rn((float)y0, a.inv_th), 0.5f)), 0);
    const int r_hi = min((int)floorf(__fsub_rn(__fmul_rn((float)(y1 - 1), a.inv_th), 0.5f)) + 1, a.tiles_y - 1);
    const int row_bytes = a.tiles_x * 256;
    const uint8_t* __restrict__ lut = a.lut + (size_t)frame * a.tiles_x * a.tiles_y * 256;
    for (int r = r_lo; r <= r_hi; ++r)
        for (int i = t; i < row_bytes / 4; i += blockDim.x)
            reinterpret_cast<uint32_t*>(s_lut[r - r_lo])[i] = __ldg(reinterpret_cast<const uint32_t*>(lut + (size_t)r * row_bytes) + i);
    __syncthreads();
    const uint8_t* __restrict__ src = a.src + (size_t)frame * a.w * a.h;
    uint8_t* __restrict__ dst = a.frames + (size_t)(a.first_slot + frame) * a.frame_stride + a.dst_off;
    if ((a.w & 3) == 0) {
        // a thread owns column quads xq, xq + blockDim, ...: the four column terms are computed once and reused down the band
        for (int xq = t; xq < (a.w >> 2); xq += blockDim.x) {
            const int x = 4 * xq;
            const ClaheCol c0 = clahe_col(x, a.inv_tw, a.tiles_x), c1 = clahe_col(x + 1, a.inv_tw, a.tiles_x);
            const ClaheCol c2 = clahe_col(x + 2, a.inv_tw, a.tiles_x), c3 = clahe_col(x + 3, a.inv_tw, a.tiles_x);
#pragma unroll 2
            for (int y = y0; y < y1; ++y) {
                const float tyf = __fsub_rn(__fmul_rn((float)y, a.inv_th), 0.5f);
                const float fl = floorf(tyf);
                const float ya = __fsub_rn(tyf, fl), ya1 = __fsub_rn(1.0f, ya);
                const uint8_t* l1 = s_lut[max((int)fl, 0) - r_lo];
                const uint8_t* l2 = s_lut[min((int)fl + 1, a.tiles_y - 1) - r_lo];
                const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)y * a.w) + xq);
                const uint32_t o = clahe_px(l1, l2, c0, v & 0xFF, ya, ya1) | (clahe_px(l1, l2, c1, (v >> 8) & 0xFF, ya, ya1) << 8) |
                                   (clahe_px(l1, l2, c2, (v >> 16) & 0xFF, ya, ya1) << 16) | (clahe_px(l1, l2, c3, v >> 24, ya, ya1) << 24);
                *reinterpret_cast<uint32_t*>(dst + (size_t)y * a.w + x) = o;
            }
        }
    } else {
        for (int x = t; x < a.w; x += blockDim.x) {
            const ClaheCol c0 = clahe_col(x, a.inv_tw, a.tiles_x);
            for (int y = y0; y < y1; ++y) {
                const float tyf = __fsub_rn(__fmul_rn((float)y, a.inv_th), 0.5f);
                const float fl = floorf(tyf);
                const float ya = __fsub_rn(tyf, fl), ya1 = __fsub_rn(1.0f, ya);
                dst[(size_t)y * a.w + x] = (uint8_t)clahe_px(s_lut[max((int)fl, 0) - r_lo], s_lut[min((int)fl + 1, a.tiles_y - 1) - r_lo], c0,
                                                             __ldg(src + (size_t)y * a.w + x), ya, ya1);
            }
        }
    }
}

}  // namespace

cudaError_t launch_clahe(dsdtm_ctx* c, int first_slot, int n, double clip_limit, int tiles_x, int tiles_y, cudaStream_t s)
{
    ClaheArgs a;
    a.src = c->clahe_src_d; a.frames = c->frames_d; a.frame_stride = c->geo.frame_stride; a.first_slot = first_slot; a.dst_off = c->geo.off[0];
    a.w = c->geo.w[0]; a.h = c->geo.h[0]; a.tiles_x = tiles_x; a.tiles_y = tiles_y; a.tw = a.w / tiles_x; a.th = a.h / tiles_y;
    const int total = a.tw * a.th;
    a.lut_scale = static_cast<float>(255) / total;                  // clahe.cpp: static_cast<float>(histSize - 1) / tileSizeTotal
    a.clip = 0;
    if (clip_limit > 0.0) a.clip = std::max(static_cast<int>(clip_limit * total / 256), 1);
    a.inv_tw = 1.0f / a.tw; a.inv_th = 1.0f / a.th;
    a.lut = c->clahe_lut_d;
    clahe_lut_kernel<<<dim3(tiles_x * tiles_y, n), 256, 0, s>>>(a);
    const int units = (a.w & 3) == 0 ? a.w >> 2 : a.w;              // column quads (or columns) per row
    const int threads = std::min(256, ((units + 31) / 32) * 32);
    clahe_apply_kernel<<<dim3((a.h + CL_ROWS - 1) / CL_ROWS, n), threads, 0, s>>>(a);
    c->launches += 2;
    return cudaGetLastError();
}

int clahe_max_tiles_x() { return CL_MAX_TX; }
int clahe_rows_per_cta() { return CL_ROWS; }

}  // namespace dsdtm
