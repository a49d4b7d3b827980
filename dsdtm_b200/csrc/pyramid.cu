// pyramid.cu -- kernel (a): 8-bit Gaussian pyramid, cv::pyrDown semantics.
//
// Replaces Frame::ComputeImagePyramid (ref: src/Frame.cpp:74-81), i.e. cv::pyrDown on CV_8UC1 with default
// arguments: dst(y,x) = (sum_{i,j in [-2,2]} k_i k_j src(R(2y+i), R(2x+j)) + 128) >> 8, k = [1 4 6 4 1],
// R = BORDER_REFLECT_101, dst size ((w+1)/2, (h+1)/2). Integer arithmetic => bit-exact in any evaluation order.
//
// One launch per level over a batch of frames (grid.z = frame). A 256-thread CTA produces a 64x16 output tile:
//   1. the (2*64+8) x (2*16+3) source window is staged in shared memory with coalesced 32-bit row loads
//      (byte loads with reflect-101 on border tiles / widths that are not a multiple of 4),
//   2. horizontal 5-tap pass into a u16 buffer (35 x 64),
//   3. vertical 5-tap pass, 4 outputs per thread packed into one 32-bit store.
// HBM-bound stage: algorithmic bytes per frame = sum_l (w_{l-1} h_{l-1} + w_l h_l) (510 000 B @640x480x5).
#include "ctx.cuh"

namespace dsdtm {

namespace {

constexpr int TW = 64;               // output tile width
constexpr int TH = 16;               // output tile height
constexpr int SW = 2 * TW + 8;       // staged source columns: [2*x0-4, 2*x0+132)
constexpr int SH = 2 * TH + 3;       // staged source rows:    [2*y0-2, 2*y0+33)
constexpr int SWP = SW + 4;          // padded row pitch in bytes (multiple of 4)

__device__ __forceinline__ int reflect101(int i, int n)
{
    // valid for -n < i < 2n-1 (enough for a 5-tap kernel on n >= 3; n < 3 handled by the loop below)
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
    return i;
}

__global__ void __launch_bounds__(256) pyrdown_kernel(uint8_t* __restrict__ frames, unsigned frame_stride, int first_slot,
                                                      const int* __restrict__ slots, unsigned src_off, unsigned dst_off,
                                                      int w, int h, int dw, int dh)
{
    __shared__ __align__(16) uint8_t s_src[SH][SWP];
    __shared__ uint16_t s_h[SH][TW];

    const int slot = slots ? slots[blockIdx.z] : first_slot + blockIdx.z;
    const uint8_t* __restrict__ src = frames + (size_t)slot * frame_stride + src_off;
    uint8_t* __restrict__ dst = frames + (size_t)slot * frame_stride + dst_off;

    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int sx0 = 2 * x0 - 4, sy0 = 2 * y0 - 2;
    const int tid = threadIdx.x;

    const bool fast = ((w & 3) == 0) && sx0 >= 0 && sx0 + SW <= w && sy0 >= 0 && sy0 + SH <= h;
    if (fast) {
        // coalesced 32-bit loads; (sy*w + sx0) is a multiple of 4 because w % 4 == 0 and sx0 % 4 == 0
        constexpr int WPR = SW / 4;
        for (int i = tid; i < SH * WPR; i += 256) {
            const int r = i / WPR, c = i - r * WPR;
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)(sy0 + r) * w + sx0) + c);
            *reinterpret_cast<uint32_t*>(&s_src[r][4 * c]) = v;
        }
    } else {
        for (int i = tid; i < SH * SW; i += 256) {
            const int r = i / SW, c = i - r * SW;
            const int sy = reflect101(sy0 + r, h), sx = reflect101(sx0 + c, w);
            s_src[r][c] = __ldg(src + (size_t)sy * w + sx);
        }
    }
    __syncthreads();

    // horizontal pass: s_h[r][x] = sum_j k_j * s_src[r][2x + j + 4 - 2 .. ]   (window column of source col 2x0+2x-2 is 2x+2)
    for (int i = tid; i < SH * TW; i += 256) {
        const int r = i / TW, x = i - r * TW;
        const uint8_t* p = &s_src[r][2 * x + 2];
        s_h[r][x] = (uint16_t)(p[0] + 4 * p[1] + 6 * p[2] + 4 * p[3] + p[4]);
    }
    __syncthreads();

    // vertical pass: thread -> row ty, 4 consecutive outputs
    const int ty = tid >> 4, tx = (tid & 15) * 4;
    const int oy = y0 + ty, ox = x0 + tx;
    if (oy < dh && ox < dw) {
        uint32_t packed = 0;
        uint8_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = tx + k;
            const int s = s_h[2 * ty][x] + 4 * s_h[2 * ty + 1][x] + 6 * s_h[2 * ty + 2][x] + 4 * s_h[2 * ty + 3][x] + s_h[2 * ty + 4][x];
            o[k] = (uint8_t)((s + 128) >> 8);
            packed |= (uint32_t)o[k] << (8 * k);
        }
        uint8_t* q = dst + (size_t)oy * dw + ox;
        if (((dw & 3) == 0) && ox + 3 < dw) {
            *reinterpret_cast<uint32_t*>(q) = packed;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (ox + k < dw) q[k] = o[k];
        }
    }
}

cudaError_t launch_levels(dsdtm_ctx* c, int first_slot, const int* slots_d, int n, cudaStream_t s)
{
    const LevelGeom& g = c->geo;
    for (int l = 1; l < g.levels; ++l) {
        dim3 grid((g.w[l] + TW - 1) / TW, (g.h[l] + TH - 1) / TH, n);
        pyrdown_kernel<<<grid, 256, 0, s>>>(c->frames_d, g.frame_stride, first_slot, slots_d, g.off[l - 1], g.off[l],
                                            g.w[l - 1], g.h[l - 1], g.w[l], g.h[l]);
        c->launches++;
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_pyramid(dsdtm_ctx* c, int first_slot, int n, cudaStream_t s) { return launch_levels(c, first_slot, nullptr, n, s); }
cudaError_t launch_pyramid_slots(dsdtm_ctx* c, const int* slots_d, int n, cudaStream_t s) { return launch_levels(c, 0, slots_d, n, s); }

}  // namespace dsdtm
