// fast.cu -- kernel (b): FAST-10 detect + score + 3x3 non-max + Shi-Tomasi + per-cell argmax, all levels, one launch.
//
// Replaces the per-level body of Feature_detector::detect (ref: src/Feature_detection.cpp:76-109) and the FAST library
// it calls (ref: Thirdparty/fast/src/faster_corner_10_sse.cpp:15-202, fast_10_score.cpp:21-3147, nonmax_3x3.cpp:17-112):
//   detect : pixel is a corner at barrier b iff 10 contiguous ring pixels are all > p+b or all < p-b (strict)
//   score  : largest barrier at which it is still a corner = max_arc min_k |d_k| - 1   (closed form of the decision tree)
//   nonmax : keep iff every 8-neighbour corner has a strictly smaller score
//   cells  : Shi-Tomasi min-eigenvalue (fp32, non-contracted, same op order as the reference) and a packed 64-bit
//            atomicMax key per grid cell = (score bits << 32) | ~(level, y, x): highest score, earliest (level, raster) on ties,
//            which is exactly "replace iff strictly greater, visiting levels ascending in raster order" (ref: :104-107).
//
// Work decomposition: one 512-thread CTA per 32x16 tile of one level of one frame (tile table covers all levels, grid.y =
// frame). The (32+10)x(16+10) halo tile is staged once in shared memory; everything else runs out of shared memory.
// Phase A rejects most pixels with the two opposite-pair tests (any 10-arc contains >= 1 of each opposite ring pair),
// survivors are compacted into a shared-memory queue so phase B (ring masks, arc test, score) runs on dense warps.
// Shi-Tomasi runs one WARP per surviving corner (64 pixels = 2 per lane; integer sums, exact in fp32 in any order).
#include "ctx.cuh"

namespace dsdtm {

namespace {

constexpr int FT_W = 32, FT_H = 16;        // interior tile
constexpr int HALO = 5;                    // shi-tomasi needs +-5, fast score of the 1-px nonmax ring needs +-4
constexpr int SMW = FT_W + 2 * HALO;       // 42
constexpr int SMH = FT_H + 2 * HALO;       // 26
constexpr int SMP = 44;                    // pitch
constexpr int SC_W = FT_W + 2, SC_H = FT_H + 2;   // score tile incl. 1-px ring
constexpr int NTHREADS = 512;

__constant__ int c_ring_dx[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
__constant__ int c_ring_dy[16] = { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 };

__device__ __forceinline__ bool has_arc10(unsigned m)
{
    const unsigned d = m | (m << 16);
    unsigned r = d & (d >> 1);
    r &= r >> 2;
    r &= r >> 4;
    r &= d >> 8;
    r &= d >> 9;
    return (r & 0xFFFFu) != 0;
}

// max over the 16 cyclic 10-arcs of the minimum of v[] over the arc
__device__ __forceinline__ int max_arc_min10(const int (&v)[16])
{
    int m2[16], m4[16], best = -100000;
#pragma unroll
    for (int i = 0; i < 16; ++i) m2[i] = min(v[i], v[(i + 1) & 15]);
#pragma unroll
    for (int i = 0; i < 16; ++i) m4[i] = min(m2[i], m2[(i + 2) & 15]);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int m8 = min(m4[i], m4[(i + 4) & 15]);
        best = max(best, min(m8, m2[(i + 8) & 15]));
    }
    return best;
}

struct FastArgs {
    uint8_t* frames; unsigned frame_stride; int first_slot;
    const int* tiles; LevelGeom geo;
    int barrier; float seed_score;
    const uint8_t* occupied;             // n * n_cells or null
    unsigned long long* cells;           // n * n_cells keys
    int n_cells, grid_cols, cell_size;
    // score-map mode (parity helper): single level, dense outputs
    uint8_t* score_out; uint8_t* nonmax_out;
};

__global__ void __launch_bounds__(NTHREADS) fast_kernel(const FastArgs a)
{
    __shared__ uint8_t s_img[SMH][SMP];
    __shared__ uint8_t s_score[SC_H][SC_W + 2];
    __shared__ unsigned short s_queue[SC_H * SC_W];
    __shared__ unsigned short s_kept[FT_H * FT_W];
    __shared__ int s_nq, s_nk;

    const int tid = threadIdx.x;
    const int t = a.tiles[blockIdx.x];
    const int L = t >> 24, tyi = (t >> 12) & 0xFFF, txi = t & 0xFFF;
    const int w = a.geo.w[L], h = a.geo.h[L];
    const int frame = blockIdx.y;
    const uint8_t* __restrict__ img = a.frames + (size_t)(a.first_slot + frame) * a.frame_stride + a.geo.off[L];
    const int x0 = txi * FT_W, y0 = tyi * FT_H;

    if (tid == 0) { s_nq = 0; s_nk = 0; }
    // stage halo tile (zero outside the image; such pixels are never used by a valid computation)
    for (int i = tid; i < SMH * SMW; i += NTHREADS) {
        const int r = i / SMW, c = i - r * SMW;
        const int y = y0 - HALO + r, x = x0 - HALO + c;
        s_img[r][c] = (x >= 0 && x < w && y >= 0 && y < h) ? __ldg(img + (size_t)y * w + x) : (uint8_t)0;
    }
    for (int i = tid; i < SC_H * (SC_W + 2); i += NTHREADS) (&s_score[0][0])[i] = 0;
    __syncthreads();

    // ---- phase A: quick reject over the (32+2)x(16+2) score region, compaction of candidates
    const int b = a.barrier;
    for (int i = tid; i < SC_H * SC_W; i += NTHREADS) {
        const int r = i / SC_W, c = i - r * SC_W;
        const int y = y0 - 1 + r, x = x0 - 1 + c;
        bool cand = false;
        if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) {     // ref: fast_10.cpp:36,41 scan bounds
            const int sr = r + HALO - 1, sc = c + HALO - 1;
            const int p = s_img[sr][sc];
            const int d0 = s_img[sr + 3][sc] - p, d8 = s_img[sr - 3][sc] - p;
            const int d4 = s_img[sr][sc + 3] - p, d12 = s_img[sr][sc - 3] - p;
            const bool br = (d0 > b || d8 > b) && (d4 > b || d12 > b);
            const bool dk = (d0 < -b || d8 < -b) && (d4 < -b || d12 < -b);
            cand = br || dk;
        }
        if (cand) s_queue[atomicAdd(&s_nq, 1)] = (unsigned short)i;
    }
    __syncthreads();

    // ---- phase B: exact arc test + score for candidates
    const int nq = s_nq;
    for (int q = tid; q < nq; q += NTHREADS) {
        const int i = s_queue[q];
        const int r = i / SC_W, c = i - r * SC_W;
        const int sr = r + HALO - 1, sc = c + HALO - 1;
        const int p = s_img[sr][sc];
        int d[16];
        unsigned bright = 0, dark = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            d[k] = (int)s_img[sr + c_ring_dy[k]][sc + c_ring_dx[k]] - p;
            bright |= (d[k] > b) ? (1u << k) : 0u;
            dark |= (d[k] < -b) ? (1u << k) : 0u;
        }
        if (has_arc10(bright) || has_arc10(dark)) {
            const int sb = max_arc_min10(d);
#pragma unroll
            for (int k = 0; k < 16; ++k) d[k] = -d[k];
            const int sd = max_arc_min10(d);
            s_score[r][c] = (uint8_t)(max(sb, sd) - 1);     // in [barrier, 254]
        }
    }
    __syncthreads();

    // ---- non-max over the interior (one thread per pixel; 512 == 32*16)
    {
        const int r = tid >> 5, c = tid & 31;
        const int y = y0 + r, x = x0 + c;
        const int s = s_score[r + 1][c + 1];
        bool keep = false;
        if (s > 0 && x < w && y < h) {
            keep = s_score[r][c] < s && s_score[r][c + 1] < s && s_score[r][c + 2] < s && s_score[r + 1][c] < s &&
                   s_score[r + 1][c + 2] < s && s_score[r + 2][c] < s && s_score[r + 2][c + 1] < s && s_score[r + 2][c + 2] < s;
        }
        if (a.score_out) {
            if (x < w && y < h) {
                a.score_out[(size_t)y * w + x] = (uint8_t)s;
                a.nonmax_out[(size_t)y * w + x] = keep ? 1 : 0;
            }
        } else if (keep) {
            s_kept[atomicAdd(&s_nk, 1)] = (unsigned short)tid;
        }
    }
    if (a.score_out) return;
    __syncthreads();

    // ---- Shi-Tomasi + per-cell argmax: one warp per surviving corner
    const int nk = s_nk;
    const int warp = tid >> 5, lane = tid & 31;
    for (int q = warp; q < nk; q += NTHREADS / 32) {
        const int id = s_kept[q];
        const int r = id >> 5, c = id & 31;
        const int y = y0 + r, x = x0 + c;
        const int k = ((y << L) / a.cell_size) * a.grid_cols + (x << L) / a.cell_size;      // ref: Feature_detection.cpp:97-98
        if (a.occupied && a.occupied[(size_t)frame * a.n_cells + k]) continue;              // ref: :100
        float score = 0.f;
        // ref: :172 "patch too close to the boundary" -> 0
        if (!(x - 4 < 1 || x + 4 >= w - 1 || y - 4 < 1 || y + 4 >= h - 1)) {
            int sxx = 0, syy = 0, sxy = 0;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int pix = lane + 32 * e;                 // 8x8 box: rows y-4..y+3, cols x-4..x+3
                const int sr = r + HALO - 4 + (pix >> 3), sc = c + HALO - 4 + (pix & 7);
                const int dx = (int)s_img[sr][sc + 1] - (int)s_img[sr][sc - 1];
                const int dy = (int)s_img[sr + 1][sc] - (int)s_img[sr - 1][sc];
                sxx += dx * dx; syy += dy * dy; sxy += dx * dy;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sxx += __shfl_xor_sync(0xffffffffu, sxx, o);
                syy += __shfl_xor_sync(0xffffffffu, syy, o);
                sxy += __shfl_xor_sync(0xffffffffu, sxy, o);
            }
            // |sums| <= 64*255^2 < 2^24: the reference's float accumulation is exact, so int -> float is the same value.
            // ref: :194-197, float arithmetic except the two double-typed constants (2.0*box_area = 128.0 and 0.5)
            const float dXX = (float)((double)(float)sxx / 128.0);
            const float dYY = (float)((double)(float)syy / 128.0);
            const float dXY = (float)((double)(float)sxy / 128.0);
            const float tr = __fadd_rn(dXX, dYY);
            const float disc = __fsub_rn(__fmul_rn(tr, tr), __fmul_rn(4.0f, __fsub_rn(__fmul_rn(dXX, dYY), __fmul_rn(dXY, dXY))));
            score = (float)(0.5 * (double)__fsub_rn(tr, __fsqrt_rn(disc)));
        }
        if (lane == 0 && score > a.seed_score) {
            const unsigned order = ((unsigned)L << 28) | ((unsigned)y << 14) | (unsigned)x;
            const unsigned long long key = ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(~order);
            atomicMax(&a.cells[(size_t)frame * a.n_cells + k], key);
        }
    }
}

FastArgs make_args(dsdtm_ctx* c, int first_slot)
{
    FastArgs a;
    a.frames = c->frames_d; a.frame_stride = c->geo.frame_stride; a.first_slot = first_slot;
    a.tiles = c->fast_tiles_d; a.geo = c->geo;
    a.barrier = 20; a.seed_score = 0;
    a.occupied = nullptr; a.cells = c->cells_d; a.n_cells = c->n_cells; a.grid_cols = c->grid_cols; a.cell_size = c->prm.cell_size;
    a.score_out = nullptr; a.nonmax_out = nullptr;
    return a;
}

}  // namespace

// host: build the tile table (level << 24 | ty << 12 | tx) for all levels
int build_fast_tiles(const LevelGeom& g, int* out /* may be null: count only */, int* level_first /* levels+1 */)
{
    int n = 0;
    for (int l = 0; l < g.levels; ++l) {
        if (level_first) level_first[l] = n;
        const int tx = (g.w[l] + FT_W - 1) / FT_W, ty = (g.h[l] + FT_H - 1) / FT_H;
        for (int y = 0; y < ty; ++y)
            for (int x = 0; x < tx; ++x) {
                if (out) out[n] = (l << 24) | (y << 12) | x;
                ++n;
            }
    }
    if (level_first) level_first[g.levels] = n;
    return n;
}

cudaError_t launch_fast_cells(dsdtm_ctx* c, int first_slot, int n, int barrier, float seed_score, bool use_occupied, cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(c->cells_d, 0, (size_t)n * c->n_cells * sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    FastArgs a = make_args(c, first_slot);
    a.barrier = barrier; a.seed_score = seed_score;
    a.occupied = use_occupied ? c->occupied_d : nullptr;
    dim3 grid(c->n_fast_tiles, n);
    fast_kernel<<<grid, NTHREADS, 0, s>>>(a);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_fast_score_map(dsdtm_ctx* c, int slot, int level, int barrier, cudaStream_t s)
{
    FastArgs a = make_args(c, slot);
    a.barrier = barrier;
    const LevelGeom& g = c->geo;
    a.score_out = c->scoremap_d;
    a.nonmax_out = c->scoremap_d + (size_t)g.w[0] * g.h[0];
    int first[DSDTM_MAX_LEVELS + 1];
    build_fast_tiles(g, nullptr, first);
    a.tiles = c->fast_tiles_d + first[level];
    dim3 grid(first[level + 1] - first[level], 1);
    fast_kernel<<<grid, NTHREADS, 0, s>>>(a);
    c->launches++;
    return cudaGetLastError();
}

}  // namespace dsdtm
