// bw_probe.cu -- what HBM bandwidth do simple streaming kernels reach on this GPU, as a function of bytes in flight per thread and
// of the store pattern? (context for the 62-65 % of the measured copy peak that pyramid / depth_convert sit at)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/bw_probe scripts/bw_probe.cu && /tmp/bw_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int U>
__global__ void __launch_bounds__(256) copy16(const uint4* __restrict__ s, uint4* __restrict__ d, size_t n)
{
    size_t i = ((size_t)blockIdx.x * U) * blockDim.x + threadIdx.x;
    uint4 v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) if (i + (size_t)k * blockDim.x < n) v[k] = __ldg(s + i + (size_t)k * blockDim.x);
#pragma unroll
    for (int k = 0; k < U; ++k) if (i + (size_t)k * blockDim.x < n) d[i + (size_t)k * blockDim.x] = v[k];
}

template <int U>
__global__ void __launch_bounds__(256) read16(const uint4* __restrict__ s, unsigned* __restrict__ d, size_t n)
{
    size_t i = ((size_t)blockIdx.x * U) * blockDim.x + threadIdx.x;
    unsigned acc = 0;
#pragma unroll
    for (int k = 0; k < U; ++k) if (i + (size_t)k * blockDim.x < n) { uint4 v = __ldg(s + i + (size_t)k * blockDim.x); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) d[0] = acc;
}

__global__ void __launch_bounds__(256) write16(uint4* __restrict__ d, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = make_uint4(1, 2, 3, 4);
}

// u16 -> f32 : (a) 16-byte load, two 16-byte stores with 32-byte lane stride (the shipped depth_convert pattern)
__global__ void __launch_bounds__(256) cvt_a(const uint4* __restrict__ s, float4* __restrict__ d, size_t n8)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    uint4 v = __ldg(s + i);
    d[2 * i] = make_float4((float)(v.x & 0xFFFF), (float)(v.x >> 16), (float)(v.y & 0xFFFF), (float)(v.y >> 16));
    d[2 * i + 1] = make_float4((float)(v.z & 0xFFFF), (float)(v.z >> 16), (float)(v.w & 0xFFFF), (float)(v.w >> 16));
}
// (b) U x (8-byte load, one fully coalesced 16-byte store)
template <int U>
__global__ void __launch_bounds__(256) cvt_b(const uint2* __restrict__ s, float4* __restrict__ d, size_t n4)
{
    size_t i = ((size_t)blockIdx.x * U) * blockDim.x + threadIdx.x;
    uint2 v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) if (i + (size_t)k * blockDim.x < n4) v[k] = __ldg(s + i + (size_t)k * blockDim.x);
#pragma unroll
    for (int k = 0; k < U; ++k) if (i + (size_t)k * blockDim.x < n4)
        d[i + (size_t)k * blockDim.x] = make_float4((float)(v[k].x & 0xFFFF), (float)(v[k].x >> 16), (float)(v[k].y & 0xFFFF), (float)(v[k].y >> 16));
}

#define TIME(name, bytes, launch)                                                        \
    do {                                                                                 \
        for (int w = 0; w < 3; ++w) { launch; }                                          \
        cudaEventRecord(e0);                                                             \
        for (int r = 0; r < 10; ++r) { launch; }                                         \
        cudaEventRecord(e1); cudaEventSynchronize(e1);                                   \
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;                           \
        printf("%-28s %8.3f ms  %8.1f GB/s  (%s)\n", name, ms, (bytes) / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError())); \
    } while (0)

int main()
{
    const size_t N = (size_t)1 << 30;            // 1 GiB source
    void *a, *b;
    cudaMalloc(&a, N); cudaMalloc(&b, 2 * N);
    cudaMemset(a, 1, N); cudaMemset(b, 0, 2 * N);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t n16 = N / 16;
    TIME("cudaMemcpy D2D", 2.0 * N, cudaMemcpyAsync(b, a, N, cudaMemcpyDeviceToDevice));
    TIME("copy16 U=1", 2.0 * N, (copy16<1><<<(unsigned)((n16 + 255) / 256), 256>>>((const uint4*)a, (uint4*)b, n16)));
    TIME("copy16 U=2", 2.0 * N, (copy16<2><<<(unsigned)((n16 / 2 + 255) / 256), 256>>>((const uint4*)a, (uint4*)b, n16)));
    TIME("copy16 U=4", 2.0 * N, (copy16<4><<<(unsigned)((n16 / 4 + 255) / 256), 256>>>((const uint4*)a, (uint4*)b, n16)));
    TIME("copy16 U=8", 2.0 * N, (copy16<8><<<(unsigned)((n16 / 8 + 255) / 256), 256>>>((const uint4*)a, (uint4*)b, n16)));
    TIME("read16 U=1", 1.0 * N, (read16<1><<<(unsigned)((n16 + 255) / 256), 256>>>((const uint4*)a, (unsigned*)b, n16)));
    TIME("read16 U=4", 1.0 * N, (read16<4><<<(unsigned)((n16 / 4 + 255) / 256), 256>>>((const uint4*)a, (unsigned*)b, n16)));
    TIME("read16 U=8", 1.0 * N, (read16<8><<<(unsigned)((n16 / 8 + 255) / 256), 256>>>((const uint4*)a, (unsigned*)b, n16)));
    TIME("write16", 1.0 * N, (write16<<<(unsigned)((n16 + 255) / 256), 256>>>((uint4*)b, n16)));
    const size_t npx = N / 2;                    // u16 pixels -> 2N bytes of float
    TIME("cvt_a (shipped pattern)", 3.0 * N, (cvt_a<<<(unsigned)((npx / 8 + 255) / 256), 256>>>((const uint4*)a, (float4*)b, npx / 8)));
    TIME("cvt_b U=1", 3.0 * N, (cvt_b<1><<<(unsigned)((npx / 4 + 255) / 256), 256>>>((const uint2*)a, (float4*)b, npx / 4)));
    TIME("cvt_b U=2", 3.0 * N, (cvt_b<2><<<(unsigned)((npx / 8 + 255) / 256), 256>>>((const uint2*)a, (float4*)b, npx / 4)));
    TIME("cvt_b U=4", 3.0 * N, (cvt_b<4><<<(unsigned)((npx / 16 + 255) / 256), 256>>>((const uint2*)a, (float4*)b, npx / 4)));
    TIME("cvt_b U=8", 3.0 * N, (cvt_b<8><<<(unsigned)((npx / 32 + 255) / 256), 256>>>((const uint2*)a, (float4*)b, npx / 4)));
    return 0;
}
