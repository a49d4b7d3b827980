// pose_opt.cu -- Optimizer::PoseOptimization on the device: one warp per frame, see pose_opt.cuh.
#include "ctx.cuh"
#include "pose_opt.cuh"

namespace dsdtm {

// Latency-bound by construction (a dependent chain of <= 100 LM iterations over a few hundred observations); the batch form
// exists so that a sweep over independent frames fills the chip: kWarps frames per CTA, each warp in its own shared-memory slice.
template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32)
pose_opt_kernel(int n_frames, const dsdtm_ba_obs* __restrict__ obs, int obs_stride, const int* __restrict__ n_obs,
                const double* __restrict__ poses_in, int max_iters, double* __restrict__ poses_out,
                double* __restrict__ res_norm, dsdtm_ba_summary* __restrict__ summaries, int soa_stride)
{
    extern __shared__ double po_smem[];
    const int warp = threadIdx.x >> 5;
    const int frame = blockIdx.x * kWarps + warp;
    if (frame >= n_frames) return;                 // whole warps leave together; no CTA-wide barrier follows
    WarpLanes ln;
    pose_optimize(ln, n_obs[frame], obs + (size_t)frame * obs_stride, po_smem + (size_t)warp * soa_stride,
                  poses_in + 7 * (size_t)frame, max_iters, poses_out + 7 * (size_t)frame,
                  res_norm ? res_norm + (size_t)frame * obs_stride : nullptr, summaries ? summaries + frame : nullptr);
}

static const int kPoseOptSmemLimit = 200 * 1024;

cudaError_t pose_opt_init(dsdtm_ctx*)
{
    cudaError_t e = cudaFuncSetAttribute(pose_opt_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPoseOptSmemLimit);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(pose_opt_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPoseOptSmemLimit);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(pose_opt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPoseOptSmemLimit);
}

// max_obs = the largest n_obs of the batch (sizes the per-warp shared-memory slice: 48 bytes per observation)
cudaError_t launch_pose_opt(dsdtm_ctx* c, int n_frames, int obs_stride, int max_obs, int max_iters, bool want_res, bool want_sum,
                            cudaStream_t s)
{
    const int soa_stride = 6 * ((max_obs + 1) & ~1);
    const size_t per_warp = (size_t)soa_stride * sizeof(double);
    double* rn = want_res ? c->po_res_d : nullptr;
    dsdtm_ba_summary* sm = want_sum ? c->po_sum_d : nullptr;
    // a lone frame (the per-frame call) gets a one-warp CTA; batches pack 4 (or 2) frames per CTA while the slices fit
    if (n_frames >= 4 && 4 * per_warp <= (size_t)kPoseOptSmemLimit)
        pose_opt_kernel<4><<<(n_frames + 3) / 4, 128, 4 * per_warp, s>>>(n_frames, c->po_obs_d, obs_stride, c->po_nobs_d, c->po_pose_in_d,
                                                                        max_iters, c->po_pose_out_d, rn, sm, soa_stride);
    else if (n_frames >= 2 && 2 * per_warp <= (size_t)kPoseOptSmemLimit)
        pose_opt_kernel<2><<<(n_frames + 1) / 2, 64, 2 * per_warp, s>>>(n_frames, c->po_obs_d, obs_stride, c->po_nobs_d, c->po_pose_in_d,
                                                                       max_iters, c->po_pose_out_d, rn, sm, soa_stride);
    else
        pose_opt_kernel<1><<<n_frames, 32, per_warp, s>>>(n_frames, c->po_obs_d, obs_stride, c->po_nobs_d, c->po_pose_in_d, max_iters,
                                                         c->po_pose_out_d, rn, sm, soa_stride);
    c->launches++;
    return cudaGetLastError();
}

}  // namespace dsdtm
