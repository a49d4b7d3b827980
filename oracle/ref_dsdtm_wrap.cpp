// ref_dsdtm_wrap.cpp -- TEST INFRASTRUCTURE. C-ABI driver over the reference's OWN classes, compiled together with the
// reference's unmodified translation units (src/Sprase_ImageAlign.cpp, Feature_alignment.cpp, Feature_detection.cpp, Camera.cpp,
// Frame.cpp, MapPoint.cpp, Keyframe.cpp, Map.cpp, Config.cpp and Thirdparty/fast/src/*.cpp, all read in place from
// /root/reference) into oracle/_ref/libdsdtm_ref.so by oracle/Makefile (target ref_dsdtm). The third-party headers those files
// include (Eigen, OpenCV, Sophus, glog, Boost, Pangolin, Ceres) are the stand-ins under tests/ref_shim/ -- see their headers for
// what is the reference's arithmetic and what is restated library arithmetic.
//
// Purpose: pin oracle/dsdtm_oracle.cpp (the restatement every CUDA kernel is checked against) to the reference's own source for
// the floating-point rows of SURVEY.md 8(a): a5-a16 and the f-1 / f-3 helpers. tests/test_ref_pin.py holds the comparisons.
// Nothing under dsdtm_b200/ links or loads this library.
#include <cstdio>
#include <cstring>
#include <vector>

#include <chrono>
#include <ctime>
#include <fstream>
#include <functional>
#include <iostream>
#include <iterator>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <utility>

// every third-party stand-in (and through them every standard header) is parsed BEFORE access control is lifted below
#include <Eigen/Dense>
#include <opencv2/opencv.hpp>
#include <sophus/se3.h>
#include <glog/logging.h>
#include <pangolin/pangolin.h>
#include <ceres/ceres.h>
#include "boost/bind.hpp"
#include "MCDWrapper.h"
#include "fast/fast.h"

// Tracking::GetCloseKeyFrames / UpdateLocalMap and the members they fill are private / protected (ref: include/Tracking.h:76-160).
// The reference's files are not touched: access control is lifted for THIS translation unit only, after the standard headers
// above have been included (access specifiers change neither layout nor name mangling).
#define private public
#define protected public
#include "Tracking.h"
#undef private
#undef protected

using namespace DSDTM;

namespace {

CameraPtr g_cam;
Map* g_map = nullptr;
std::vector<FramePtr> g_frames;
std::vector<KeyFrame*> g_kfs;
std::vector<MapPoint*> g_mps;
std::shared_ptr<Feature_detector> g_det;
std::shared_ptr<Feature_Alignment> g_fa;

Sophus::SE3 pose_from(const double p[7])  // {qw,qx,qy,qz,tx,ty,tz}, taken bit for bit (the oracle does the same)
{
    return Sophus::SE3(Sophus::SO3::fromUnitQuaternionRaw(Sophus::Quaterniond(p[0], p[1], p[2], p[3])), Eigen::Vector3d(p[4], p[5], p[6]));
}
void pose_to(const Sophus::SE3& T, double p[7])
{
    const Sophus::Quaterniond& q = T.unit_quaternion();
    p[0] = q.w(); p[1] = q.x(); p[2] = q.y(); p[3] = q.z();
    p[4] = T.translation()(0); p[5] = T.translation()(1); p[6] = T.translation()(2);
}
int find_mp(const MapPoint* m)
{
    for (size_t i = 0; i < g_mps.size(); ++i) if (g_mps[i] == m) return (int)i;
    return -1;
}
int find_kf(const KeyFrame* k)
{
    for (size_t i = 0; i < g_kfs.size(); ++i) if (g_kfs[i] == k) return (int)i;
    return -1;
}

// protected members of Sprase_ImgAlign, read through a derived class (the reference's class is not modified)
struct SAProbe : public Sprase_ImgAlign {
    SAProbe(int a, int b, int c) : Sprase_ImgAlign(a, b, c) {}
    void bind(FramePtr cur, FramePtr ref) { mCurFrame = cur; mRefFrame = ref; }
    int n() const { return (int)mRefPatch.rows(); }
    const double* patch() const { return mRefPatch.data(); }
    const double* jac() const { return mJocabianPatch.data(); }
    double pt(int r, int c) const { return mRefNormals(r, c); }
    void zero() { H.setZero(); JRes.setZero(); }
    double h(int i, int j) const { return H(i, j); }
    double b(int i) const { return JRes(i); }
};

}  // namespace

extern "C" {

// Config::setParameterFile + Camera(RGB_PinHole) (ref: src/Config.cpp:11-22, src/Camera.cpp:16-20,32-85)
int ref_init(const char* yaml_path)
{
    Config::setParameterFile(yaml_path);
    g_cam = CameraPtr(new Camera(RGB_PinHole));
    g_map = new Map();
    g_frames.clear(); g_kfs.clear(); g_mps.clear();
    g_det.reset(new Feature_detector());
    g_fa.reset(new Feature_Alignment(g_cam));
    return (g_cam->mwidth > 0 && g_cam->mheight > 0) ? 0 : -1;
}
void ref_reset_objects()
{
    g_frames.clear(); g_kfs.clear(); g_mps.clear();  // the reference never frees Feature / MapPoint / KeyFrame either
    g_map = new Map();
    g_det.reset(new Feature_detector());
    g_fa.reset(new Feature_Alignment(g_cam));
}

// ---- Frame (ref: src/Frame.cpp:27-81) ----
int ref_frame_create(const uint8_t* img, int w, int h, const float* depth, const double pose_c2w[7])
{
    cv::Mat color(h, w, CV_8UC1);
    for (int y = 0; y < h; ++y) std::memcpy(color.ptr<uchar>(y), img + (size_t)y * w, (size_t)w);
    FramePtr f;
    if (depth) {
        cv::Mat d(h, w, CV_32FC1);
        for (int y = 0; y < h; ++y) std::memcpy(d.ptr<float>(y), depth + (size_t)y * w, (size_t)w * sizeof(float));
        f = FramePtr(new Frame(g_cam, color, d, 0.0));
    } else {
        f = FramePtr(new Frame(g_cam, color, 0.0));
    }
    f->Set_Pose(pose_from(pose_c2w));
    g_frames.push_back(f);
    return (int)g_frames.size() - 1;
}
int ref_frame_levels(int fr) { return (int)g_frames[fr]->mvImg_Pyr.size(); }
void ref_frame_level(int fr, int level, uint8_t* out, int* w, int* h)
{
    const cv::Mat& m = g_frames[fr]->mvImg_Pyr[level];
    *w = m.cols; *h = m.rows;
    if (out) for (int y = 0; y < m.rows; ++y) std::memcpy(out + (size_t)y * m.cols, m.ptr<uchar>(y), (size_t)m.cols);
}
void ref_frame_set_pose(int fr, const double pose_c2w[7]) { g_frames[fr]->Set_Pose(pose_from(pose_c2w)); }
void ref_frame_get_pose(int fr, double pose_c2w[7], double center[3])
{
    pose_to(g_frames[fr]->Get_Pose(), pose_c2w);
    const Eigen::Vector3d c = g_frames[fr]->Get_CameraCnt();
    if (center) { center[0] = c(0); center[1] = c(1); center[2] = c(2); }
}
// Frame::Add_Feature(new Feature(frame, px, level), tbNormal) (ref: src/Frame.cpp:83-92, include/Feature.h:27-37)
int ref_frame_add_feature(int fr, float x, float y, int level, int with_normal)
{
    Frame* f = g_frames[fr].get();
    f->Add_Feature(new Feature(f, cv::Point2f(x, y), level), with_normal != 0);
    return (int)f->mvFeatures.size() - 1;
}
void ref_feature_set_normal(int fr, int idx, const double n[3]) { g_frames[fr]->mvFeatures[idx]->mNormal = Eigen::Vector3d(n[0], n[1], n[2]); }
// Feature::SetPose(MapPoint*) (ref: include/Feature.h:41-45)
void ref_feature_set_mappoint(int fr, int idx, int mp) { g_frames[fr]->mvFeatures[idx]->SetPose(g_mps[mp]); }
int ref_frame_feature_count(int fr) { return (int)g_frames[fr]->mvFeatures.size(); }
void ref_frame_features(int fr, float* px, int* level, double* normal, int* mp, int* initial)
{
    const Features& fs = g_frames[fr]->mvFeatures;
    for (size_t i = 0; i < fs.size(); ++i) {
        if (px) { px[2 * i] = fs[i]->mpx.x; px[2 * i + 1] = fs[i]->mpx.y; }
        if (level) level[i] = fs[i]->mlevel;
        if (normal) for (int k = 0; k < 3; ++k) normal[3 * i + k] = fs[i]->mNormal(k);
        if (mp) mp[i] = find_mp(fs[i]->Mpt);
        if (initial) initial[i] = fs[i]->mbInitial ? 1 : 0;
    }
}
int ref_frame_mappoints(int fr, int* mp, int cap)
{
    const std::vector<MapPoint*>& v = g_frames[fr]->mvMapPoints;
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) mp[i] = find_mp(v[i]);
    return (int)v.size();
}
int ref_frame_mask(int fr, uint8_t* out)
{
    const cv::Mat& m = g_frames[fr]->mImgMask;
    if (m.empty()) return 0;
    for (int y = 0; y < m.rows; ++y) std::memcpy(out + (size_t)y * m.cols, m.ptr<uchar>(y), (size_t)m.cols);
    return 1;
}
// Frame helpers of rows f-1 / f-3 (ref: src/Frame.cpp:94-157,200-224,300-323)
void ref_frame_undistort_features(int fr) { g_frames[fr]->UndistortFeatures(); }
float ref_frame_feature_depth(int fr, float x, float y) { return g_frames[fr]->Get_FeatureDetph(cv::Point2f(x, y)); }
void ref_frame_unproject(int fr, float x, float y, float d, double out[3])
{
    const Eigen::Vector3d p = g_frames[fr]->UnProject(cv::Point2f(x, y), d);
    out[0] = p(0); out[1] = p(1); out[2] = p(2);
}
int ref_frame_is_visible(int fr, const double p[3], int boundary) { return g_frames[fr]->isVisible(Eigen::Vector3d(p[0], p[1], p[2]), boundary) ? 1 : 0; }
void ref_frame_world2pixel(int fr, const double p[3], double px[2])
{
    const Eigen::Vector2d q = g_frames[fr]->World2Pixel(Eigen::Vector3d(p[0], p[1], p[2]));
    px[0] = q(0); px[1] = q(1);
}
int ref_is_in_image(float x, float y, int boundary, int level) { return g_cam->IsInImage(cv::Point2f(x, y), boundary, level) ? 1 : 0; }

// ---- KeyFrame / MapPoint (ref: src/Keyframe.cpp:10-22, src/MapPoint.cpp:19-30,45-55,133-186) ----
int ref_keyframe_create(int fr)
{
    g_kfs.push_back(new KeyFrame(g_frames[fr]));
    g_map->AddKeyFrame(g_kfs.back());
    return (int)g_kfs.size() - 1;
}
int ref_mappoint_create(const double pos[3], int kf)
{
    Eigen::Vector3d p(pos[0], pos[1], pos[2]);
    g_mps.push_back(new MapPoint(p, g_kfs[kf], g_map));
    g_map->AddMapPoint(g_mps.back());
    return (int)g_mps.size() - 1;
}
void ref_mappoint_add_observation(int mp, int kf, int feat_idx) { g_mps[mp]->Add_Observation(g_kfs[kf], (size_t)feat_idx); }
void ref_mappoint_increase_found(int mp, int n) { g_mps[mp]->IncreaseFound(n); }
int ref_mappoint_found(int mp) { return g_mps[mp]->Get_FoundNums(); }
void ref_mappoint_set_outlier(int mp, int bad) { g_mps[mp]->mbOutlier = bad != 0; }
// the feature of key frame kf at index idx becomes an observation carrier: Feature::SetPose + normal as the key frame stored it
void ref_keyframe_feature_set_mappoint(int kf, int idx, int mp) { g_kfs[kf]->mvFeatures[idx]->SetPose(g_mps[mp]); }
int ref_mappoint_closest_obs(int mp, int fr, int* kf, int* feat_idx)
{
    Feature* f = nullptr;
    KeyFrame* k = nullptr;
    const bool ok = g_mps[mp]->Get_ClosetObs(g_frames[fr].get(), f, k);
    *kf = find_kf(k);
    *feat_idx = -1;
    if (k) for (size_t i = 0; i < k->mvFeatures.size(); ++i) if (k->mvFeatures[i] == f) { *feat_idx = (int)i; break; }
    return ok ? 1 : 0;
}

// ---- Feature_detector (ref: src/Feature_detection.cpp) ----
float ref_shitomasi(const uint8_t* img, int w, int h, int u, int v)
{
    cv::Mat m(h, w, CV_8UC1, const_cast<uint8_t*>(img));
    return g_det->shiTomasiScore(m, u, v);
}
void ref_detector_set_existing(const float* px, int n)
{
    std::vector<cv::Point2f> v;
    for (int i = 0; i < n; ++i) v.push_back(cv::Point2f(px[2 * i], px[2 * i + 1]));
    g_det->Set_ExistingFeatures(v);
}
void ref_detector_set_existing_from_frame(int fr) { g_det->Set_ExistingFeatures(g_frames[fr]->mvFeatures); }
// detect() appends to frame->mvFeatures and releases the mask (ref: :69-154)
int ref_detect(int fr, double thr, int first)
{
    g_det->detect(g_frames[fr].get(), thr, first != 0);
    return (int)g_frames[fr]->mvFeatures.size();
}

// ---- Sprase_ImgAlign (ref: src/Sprase_ImageAlign.cpp) ----
int ref_sparse_align_run(int cur, int ref, int max_level, int min_level, int max_iters, double pose_cur_c2w_out[7])
{
    Sprase_ImgAlign sa(max_level, min_level, max_iters);
    const int n = sa.Run(g_frames[cur], g_frames[ref]);
    pose_to(g_frames[cur]->Get_Pose(), pose_cur_c2w_out);
    return n;
}
// one GetJocabianMat(level) + one ComputeResiduals(T, level, true) at a given pose; returns the number of staged features.
// ref_patch: n x 16, jac: (n*16) x 6 row-major, ref_pts: n x 3 (mRefNormals columns)
int ref_sparse_align_linearize(int cur, int ref, int level, const double pose_c2r[7], int cap, double* ref_patch, double* jac,
                               double* ref_pts, double H[36], double b[6], double* chi2, int* n_pts)
{
    SAProbe sa(5, 0, 1);
    sa.bind(g_frames[cur], g_frames[ref]);
    sa.GetJocabianMat(level);
    const int n = sa.n();
    if (n > cap) return -n;
    if (ref_patch) std::memcpy(ref_patch, sa.patch(), (size_t)n * 16 * sizeof(double));
    if (jac) std::memcpy(jac, sa.jac(), (size_t)n * 16 * 6 * sizeof(double));
    if (ref_pts) for (int j = 0; j < n; ++j) for (int k = 0; k < 3; ++k) ref_pts[3 * j + k] = sa.pt(k, j);
    Sophus::SE3 T = pose_from(pose_c2r);
    sa.zero();
    int np = 0;
    const double c = sa.ComputeResiduals(T, level, true, np);
    for (int i = 0; i < 6; ++i) { b[i] = sa.b(i); for (int j = 0; j < 6; ++j) H[6 * i + j] = sa.h(i, j); }
    *chi2 = c; *n_pts = np;
    return n;
}

// ---- Feature_Alignment (ref: src/Feature_alignment.cpp) ----
int ref_align2d(const uint8_t* img, int w, int h, const uint8_t patch10[100], const uint8_t patch8[64], int iters, double px[2])
{
    cv::Mat m(h, w, CV_8UC1, const_cast<uint8_t*>(img));
    uint8_t p10[100], p8[64];
    std::memcpy(p10, patch10, 100); std::memcpy(p8, patch8, 64);
    Eigen::Vector2d p(px[0], px[1]);
    const bool ok = Feature_Alignment::Align2DGaussNewton(m, p10, p8, iters, p);
    px[0] = p(0); px[1] = p(1);
    return ok ? 1 : 0;
}
int ref_best_search_level(const double A[4], int max_level)
{
    Eigen::Matrix2d M;
    M << A[0], A[1], A[2], A[3];
    return g_fa->GetBestSearchLevel(M, max_level);
}
void ref_warp_affine(const double A[4], const uint8_t* img, int w, int h, float px_x, float px_y, int ref_level, int search_level, uint8_t out[100])
{
    Eigen::Matrix2d M;
    M << A[0], A[1], A[2], A[3];
    cv::Mat m(h, w, CV_8UC1, const_cast<uint8_t*>(img));
    Feature f(nullptr, cv::Point2f(px_x, px_y), ref_level);
    g_fa->WarpAffine(M, m, &f, search_level, out);
}
void ref_solve_affine(int kf, int cur, int feat_idx, int mp, double A[4])
{
    const Eigen::Matrix2d M = g_fa->SolveAffineMatrix(g_kfs[kf], g_frames[cur], g_kfs[kf]->mvFeatures[feat_idx], g_mps[mp]);
    A[0] = M(0, 0); A[1] = M(0, 1); A[2] = M(1, 0); A[3] = M(1, 1);
}
int ref_find_match_direct(int mp, int cur, double px[2], int* level)
{
    Eigen::Vector2d p(px[0], px[1]);
    int l = 0;
    const bool ok = g_fa->FindMatchDirect(g_mps[mp], g_frames[cur], p, l);
    px[0] = p(0); px[1] = p(1); *level = l;
    return ok ? 1 : 0;
}
void ref_fa_reset_grid() { g_fa->ResetGrid(); }
int ref_fa_reproject_point(int cur, int mp) { return g_fa->ReprojectPoint(g_frames[cur], g_mps[mp]) ? 1 : 0; }
void ref_fa_search_local_points(int cur) { g_fa->SearchLocalPoints(g_frames[cur]); }

// ---- Tracking (ref: src/Tracking.cpp:14-38,257-345): the caller's half of SURVEY 8f-1 ----
static Tracking* g_tracker = nullptr;
int ref_tracking_create()
{
    g_tracker = new Tracking(g_cam, g_map, nullptr);
    return 0;
}
void ref_keyframe_add_mappoint(int kf, int idx, int mp) { g_kfs[kf]->Add_MapPoint(g_mps[mp], idx); }
// GetCloseKeyFrames(frame, list): returns the list as (key-frame id, distance) in the order the reference produced it
int ref_tracking_close_keyframes(int fr, int* kf_ids, double* dist, int cap)
{
    std::list<std::pair<KeyFrame*, double> > l;
    g_tracker->GetCloseKeyFrames(g_frames[fr].get(), l);
    int n = 0;
    for (auto it = l.begin(); it != l.end() && n < cap; ++it, ++n) { kf_ids[n] = find_kf(it->first); dist[n] = it->second; }
    return (int)l.size();
}
// UpdateLocalMap() on mCurrentFrame = frame: returns mvpLocalKeyFrames (ids, in order) and the number of local map points
int ref_tracking_update_local_map(int fr, int* local_kfs, int cap, int* n_local_points)
{
    g_tracker->mCurrentFrame = g_frames[fr];
    g_tracker->UpdateLocalMap();
    int n = 0;
    for (KeyFrame* k : g_tracker->mvpLocalKeyFrames) if (n < cap) local_kfs[n++] = find_kf(k);
    *n_local_points = (int)g_tracker->mvpLocalMapPoints.size();
    return (int)g_tracker->mvpLocalKeyFrames.size();
}
// the tracker's own Feature_Alignment instance, whose grid UpdateLocalMap has just filled (ref: src/Tracking.cpp:224)
void ref_tracking_search_local_points() { g_tracker->mFeature_Alignment->SearchLocalPoints(g_tracker->mCurrentFrame); }

// ---- SE3 stand-in, exposed so that the tests can state how far it is from the oracle's restatement ----
void ref_se3_exp(const double x[6], double out[7])
{
    Sophus::Vector6d v;
    v << x[0], x[1], x[2], x[3], x[4], x[5];
    pose_to(Sophus::SE3::exp(v), out);
}
void ref_se3_mul(const double a[7], const double b[7], double out[7]) { pose_to(pose_from(a) * pose_from(b), out); }
void ref_se3_inv(const double a[7], double out[7]) { pose_to(pose_from(a).inverse(), out); }

}  // extern "C"
