// Host build of the product's pose-refinement routine (dsdtm_b200/csrc/pose_opt.cuh, lane policy SerialLanes) so the CPU
// suite can check the very source the kernel runs against the oracle's Ceres restatement without a GPU. Test infrastructure.
#include <vector>

#include "../../dsdtm_b200/csrc/pose_opt.cuh"

extern "C" void prod_pose_optimize(int n, const dsdtm_ba_obs* obs, const double* pose_in, int max_iters, double* pose_out,
                                   double* res_norm, dsdtm_ba_summary* summary)
{
    std::vector<double> soa(6 * (size_t)(n > 0 ? n : 1));
    dsdtm::SerialLanes ln;
    dsdtm::pose_optimize(ln, n, obs, soa.data(), pose_in, max_iters, pose_out, res_norm, summary);
}
