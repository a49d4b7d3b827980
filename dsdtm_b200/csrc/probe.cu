// probe.cu -- measurement aid, not on the hot path: what the FP64 pipe of THIS GPU sustains on dependent-free DFMA streams, so that
// bench.py can put the sparse-alignment kernel (fp64 Gauss-Newton, latency/issue-bound) against a measured ceiling instead of HBM
// bandwidth ("roofline.bound hbm mislabels this kernel; no measured fp64 peak exists").
#include "ctx.cuh"

namespace dsdtm {

namespace {
// 16 independent accumulator chains per thread: enough instruction-level parallelism to keep the pipe full at any occupancy.
__global__ void __launch_bounds__(256) dfma_probe_kernel(double* out, int iters, double a, double b)
{
    double x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = (double)(threadIdx.x + k) * 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += x[k];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // keeps the chains alive, never true in practice
}
}  // namespace

}  // namespace dsdtm

extern "C" int dsdtm_probe_fp64(dsdtm_ctx* c, double* tflops, double* dfma_warp_insts_per_clk_per_sm)
{
    using namespace dsdtm;
    if (!c || !tflops) return DSDTM_E_ARG;
    const int iters = 4096, threads = 256, blocks = c->sm_count * 8;
    double* out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return fail(c, DSDTM_E_CUDA, "probe alloc", cudaGetLastError());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {          // first pass warms up; best of the rest
        cudaEventRecord(e0, c->stream);
        dfma_probe_kernel<<<blocks, threads, 0, c->stream>>>(out, iters, 0.999999999, 1e-12);
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    if (err != cudaSuccess) return fail(c, DSDTM_E_CUDA, "dfma probe", err);
    const double dfma = (double)blocks * threads * (double)iters * 16.0;
    *tflops = 2.0 * dfma / (best * 1e-3) / 1e12;
    if (dfma_warp_insts_per_clk_per_sm) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
        *dfma_warp_insts_per_clk_per_sm = (dfma / 32.0) / (best * 1e-3) / ((double)khz * 1e3) / c->sm_count;   // against the NOMINAL max clock
    }
    return 0;
}
