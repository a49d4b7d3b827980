// ref_link_stubs.cpp -- TEST INFRASTRUCTURE. Link seam for oracle/_ref/libdsdtm_ref.so: src/Tracking.cpp, src/Initializer.cpp and
// src/LocalMapping.cpp of the reference are compiled unmodified, but they name symbols of subsystems that are out of scope
// (Viewer / Pangolin, Moving_Detecter / fastMCD, Rarsac_base -- constructed but never called --, and the Ceres
// solves of Optimizer). Those get definitions here so that the library links; anything that would need their behaviour aborts.
// The pinned functions (Tracking::GetCloseKeyFrames, Tracking::UpdateLocalMap) call none of them.
#include <cstdio>
#include <cstdlib>

#include "Tracking.h"

namespace DSDTM {

static void out_of_scope(const char* what)
{
    std::fprintf(stderr, "libdsdtm_ref: %s is outside the pinned path (link stub)\n", what);
    std::abort();
}

Rarsac_base::Rarsac_base() {}
Rarsac_base::~Rarsac_base() {}
Moving_Detecter::Moving_Detecter() {}
Moving_Detecter::~Moving_Detecter() {}
cv::Mat Moving_Detecter::Mod_FastMCD(const cv::Mat, std::vector<cv::Point2f>, std::vector<cv::Point2f>) { out_of_scope("Moving_Detecter::Mod_FastMCD"); return cv::Mat(); }
void Viewer::SetCurrentCameraPose(const Sophus::SE3&) {}
void Viewer::RequestStop() {}
bool Viewer::IsStopped() { return true; }
void Viewer::Release() {}
void Optimizer::PoseOptimization(FramePtr, int) { out_of_scope("Optimizer::PoseOptimization (ceres::Solve)"); }
void Optimizer::LocalBundleAdjustment(KeyFrame*, Map*) { out_of_scope("Optimizer::LocalBundleAdjustment (ceres::Solve)"); }

}  // namespace DSDTM
