// host_shim.cpp -- TEST INFRASTRUCTURE: a C shim over the C++ host adapters (dsdtm_b200/host) so that pytest can drive
// DSDTM::Frame / Feature_detector / Sprase_ImgAlign / Feature_Alignment exactly as Tracking would.
#include <cstring>
#include <memory>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../dsdtm_b200/host/dsdtm_host.h"

using namespace DSDTM;

struct HsFrame { FramePtr f; std::vector<MapPoint*> owned; };
struct HsKf { KeyFrame* kf; };
static std::string g_err;
static std::vector<MapPoint*> g_mps;      // creation-order registry: index = map point id used by the Python side

// Tracking owns ONE Feature_Alignment for its whole life (ref: src/Tracking.cpp:31-37); so does the shim, per camera
static std::unique_ptr<Feature_Alignment> g_fa;
static void* g_fa_cam = nullptr;
static Feature_Alignment& feature_alignment(void* cam)
{
    if (!g_fa || g_fa_cam != cam) { g_fa.reset(new Feature_Alignment(*static_cast<CameraPtr*>(cam))); g_fa_cam = cam; }
    return *g_fa;
}

extern "C" {

const char* hs_last_error() { return g_err.c_str(); }
void hs_map_reset();
void hs_reset() { hs_map_reset(); g_fa.reset(); g_fa_cam = nullptr; GpuRuntime::Shutdown(); Config::Clear(); g_mps.clear(); }
void hs_config_set(const char* k, const char* v) { Config::Set(k, v); }
int hs_config_load(const char* path) { try { Config::setParameterFile(path); return 0; } catch (std::exception& e) { g_err = e.what(); return -1; } }
double hs_config_get(const char* k) { return Config::Get<double>(k); }
int hs_config_get_int(const char* k) { return Config::Get<int>(k); }

void* hs_camera_new() { return new CameraPtr(new Camera()); }

void* hs_frame_new(void* cam, const uint8_t* img, int w, int h, const double* pose7)
{
    try {
        Mat8 m(h, w, img, w);
        auto* hf = new HsFrame{ FramePtr(new Frame(*static_cast<CameraPtr*>(cam), m, 0.0)), {} };
        hf->f->Set_Pose(SE3(pose7));
        return hf;
    } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}
void hs_frame_free(void* f) { delete static_cast<HsFrame*>(f); }
void hs_frame_set_pose(void* f, const double* pose7) { static_cast<HsFrame*>(f)->f->Set_Pose(SE3(pose7)); }
void hs_frame_get_pose(void* f, double* pose7) { std::memcpy(pose7, static_cast<HsFrame*>(f)->f->Get_Pose().data(), 7 * sizeof(double)); }
int hs_frame_level(void* f, int l, uint8_t* out)
{
    const Mat8& m = static_cast<HsFrame*>(f)->f->mvImg_Pyr[l];
    for (int y = 0; y < m.rows; ++y) std::memcpy(out + (size_t)y * m.cols, m.data + (size_t)y * m.step, m.cols);
    return m.rows * m.cols;
}
int hs_frame_n_features(void* f) { return (int)static_cast<HsFrame*>(f)->f->mvFeatures.size(); }
void hs_frame_get_features(void* f, float* px, int* level, int* initial)
{
    const Features& fs = static_cast<HsFrame*>(f)->f->mvFeatures;
    for (size_t i = 0; i < fs.size(); ++i) { px[2 * i] = fs[i]->mpx.x; px[2 * i + 1] = fs[i]->mpx.y; level[i] = fs[i]->mlevel; initial[i] = fs[i]->mbInitial; }
}
void hs_frame_mask(void* f, uint8_t* out)
{
    const Mat8& m = static_cast<HsFrame*>(f)->f->mImgMask;
    if (!m.empty()) std::memcpy(out, m.data, (size_t)m.rows * m.cols);
}

// Feature_detector::detect as Initializer / CraeteKeyframe call it (ref: src/Initializer.cpp:44, src/Tracking.cpp:415-416)
int hs_detect(void* f, double thr, int use_existing)
{
    try {
        Feature_detector det;
        Frame* fr = static_cast<HsFrame*>(f)->f.get();
        if (use_existing) det.Set_ExistingFeatures(fr->mvFeatures);
        det.detect(fr, thr);
        return (int)fr->mvFeatures.size();
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// Tracking::CraeteKeyframe's RGB-D part (ref: src/Tracking.cpp:412-464): SetDepth (replaces the CV_32F mDepthImg), UndistortFeatures,
// then per non-initial feature Get_FeatureDetph + UnProject. out: n x {px.x, px.y, depth, world xyz (device), world xyz (host UnProject)}
int hs_frame_keyframe_lift(void* f, const uint16_t* depth, float depth_scale, double* out9, double* normals3)
{
    try {
        Frame* fr = static_cast<HsFrame*>(f)->f.get();
        fr->SetDepth(depth, 0, depth_scale);
        fr->UndistortFeatures();
        const auto& L = fr->Lifted();
        for (size_t i = 0; i < fr->mvFeatures.size(); ++i) {
            Feature* ft = fr->mvFeatures[i];
            double* o = out9 + 9 * i;
            o[0] = ft->mpx.x; o[1] = ft->mpx.y;
            for (int k = 0; k < 3; ++k) normals3[3 * i + k] = ft->mNormal[k];
            if (ft->mbInitial) { o[2] = -2.0; continue; }
            const float z = fr->Get_FeatureDetph(ft);
            o[2] = z;
            if (z < 0) continue;
            const Vector3d P = fr->UnProject(ft->mpx, z);
            for (int k = 0; k < 3; ++k) { o[3 + k] = L[i].point_w[k]; o[6 + k] = P[k]; }
        }
        return (int)fr->mvFeatures.size();
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// CreateInitialMapRGBD-like: bearing vectors + one MapPoint per feature with a non-zero world point
void hs_frame_attach_points(void* f, const double* pts, const uint8_t* has)
{
    HsFrame* hf = static_cast<HsFrame*>(f);
    Frame* fr = hf->f.get();
    fr->mvMapPoints.assign(fr->mvFeatures.size(), nullptr);
    for (size_t i = 0; i < fr->mvFeatures.size(); ++i) {
        Feature* ft = fr->mvFeatures[i];
        ft->mNormal = fr->mCamera->Pixel2Camera(ft->mpx, 1.0f);
        ft->mNormal.normalize();
        if (!has[i]) continue;
        MapPoint* mp = new MapPoint(Vector3d(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
        hf->owned.push_back(mp);
        g_mps.push_back(mp);
        ft->SetPose(mp);
        fr->mvMapPoints[i] = mp;
    }
}

// CraeteKeyframe's loop over the features (ref: src/Tracking.cpp:422-454): features that already carry a map point keep it;
// the NEW ones (index >= start, added by detect) get their bearing (UndistortFeatures with zero distortion) and, where
// has[i - start] is set, a map point at pts[i - start].
void hs_frame_attach_points_from(void* f, int start, const double* pts, const uint8_t* has)
{
    HsFrame* hf = static_cast<HsFrame*>(f);
    Frame* fr = hf->f.get();
    fr->mvMapPoints.resize(fr->mvFeatures.size(), nullptr);
    for (size_t i = (size_t)start; i < fr->mvFeatures.size(); ++i) {
        Feature* ft = fr->mvFeatures[i];
        ft->mNormal = fr->mCamera->Pixel2Camera(ft->mpx, 1.0f);
        ft->mNormal.normalize();
        if (!has[i - start]) continue;
        const double* p = pts + 3 * (i - start);
        MapPoint* mp = new MapPoint(Vector3d(p[0], p[1], p[2]));
        hf->owned.push_back(mp);
        g_mps.push_back(mp);
        ft->SetPose(mp);
        fr->mvMapPoints[i] = mp;
    }
}

// map point id (creation order) per feature, -1 if none
void hs_frame_feature_mp_ids(void* f, int* ids)
{
    const Features& fs = static_cast<HsFrame*>(f)->f->mvFeatures;
    for (size_t i = 0; i < fs.size(); ++i) {
        ids[i] = -1;
        for (size_t k = 0; k < g_mps.size(); ++k) if (g_mps[k] == fs[i]->Mpt) { ids[i] = (int)k; break; }
    }
}
void hs_set_use_store(int on) { Tracking::sUseStore = on != 0; }
void hs_set_speculate(int on) { Tracking::sSpeculate = on != 0; }
void hs_mappoint_increase_found(int id, int n) { if (id >= 0 && id < (int)g_mps.size()) g_mps[id]->IncreaseFound(n); }
void hs_mappoint_pose(int id, double* out3) { const Vector3d P = g_mps[id]->Get_Pose(); out3[0] = P[0]; out3[1] = P[1]; out3[2] = P[2]; }
int hs_mappoint_found(int id) { return (id >= 0 && id < (int)g_mps.size()) ? g_mps[id]->Get_FoundNums() : -1; }

int hs_sparse_align_run(int maxl, int minl, int iters, void* cur, void* ref, double* pose_out, dsdtm_iter_log* log, int cap, int* n_log)
{
    try {
        Sprase_ImgAlign sa(maxl, minl, iters);
        sa.EnableLog(log != nullptr && cap > 0);
        const int n = sa.Run(static_cast<HsFrame*>(cur)->f, static_cast<HsFrame*>(ref)->f);
        std::memcpy(pose_out, static_cast<HsFrame*>(cur)->f->Get_Pose().data(), 7 * sizeof(double));
        const auto& l = sa.LastLog();
        *n_log = (int)l.size();
        for (int i = 0; i < (int)l.size() && i < cap; ++i) log[i] = l[i];
        return n;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// Tracking::UpdateLocalMap + SearchLocalPoints with the reference's own local-map selection (ref: src/Tracking.cpp:257-345):
// a Map of all key frames, the 10 nearest close ones chosen on the device. local_out (optional): rows of mvpLocalKeyFrames.
static std::unique_ptr<Map> g_map;
static std::unique_ptr<Tracking> g_trk;
static void* g_trk_cam = nullptr;
void hs_map_reset() { g_trk.reset(); g_map.reset(); g_trk_cam = nullptr; }
int hs_map_add_keyframe(void* kf) { if (!g_map) g_map.reset(new Map()); g_map->AddKeyFrame(static_cast<HsKf*>(kf)->kf); return g_map->ReturnKeyFramesSize(); }
void hs_map_mark_moved(void* kf) { if (g_map) g_map->MarkMoved(static_cast<HsKf*>(kf)->kf); }
void hs_keyframe_set_pose(void* kf, const double* pose7) { static_cast<HsKf*>(kf)->kf->Set_Pose(SE3(pose7)); }
int hs_track_local_map(void* cam, void* cur, int* local_out, int* n_local, int* n_reprojected)
{
    try {
        if (!g_map) g_map.reset(new Map());
        if (!g_trk || g_trk_cam != cam) { g_trk.reset(new Tracking(*static_cast<CameraPtr*>(cam), g_map.get())); g_trk_cam = cam; }
        FramePtr c = static_cast<HsFrame*>(cur)->f;
        g_trk->SetCurrentFrame(c);
        g_trk->UpdateLocalMap();
        const std::vector<KeyFrame*> all = g_map->GetAllKeyFrames();
        if (n_local) *n_local = (int)g_trk->mvpLocalKeyFrames.size();
        if (local_out)
            for (size_t i = 0; i < g_trk->mvpLocalKeyFrames.size(); ++i)
                for (size_t k = 0; k < all.size(); ++k) if (all[k] == g_trk->mvpLocalKeyFrames[i]) local_out[i] = (int)k;
        if (n_reprojected) *n_reprojected = g_trk->LastReprojected();
        g_trk->mFeature_Alignment->SearchLocalPoints(c);
        return g_trk->mFeature_Alignment->LastMatches();
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// Optimizer::PoseOptimization (ref: src/Optimizer.cpp:20-101, called at src/Tracking.cpp:236). residuals: n_blocks doubles (cap entries max).
int hs_pose_optimization(void* cur, double* pose_out, dsdtm_ba_summary* summary, double* residuals, int cap)
{
    try {
        FramePtr c = static_cast<HsFrame*>(cur)->f;
        Optimizer::PoseOptimization(c, 10);                     // Tracking passes 10; the reference ignores it
        std::memcpy(pose_out, c->Get_Pose().data(), 7 * sizeof(double));
        *summary = Optimizer::LastSummary();
        const auto& r = Optimizer::LastResiduals();
        for (int i = 0; i < (int)r.size() && i < cap; ++i) residuals[i] = r[i];
        return (int)r.size();
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}
// test set-up helpers: overwrite a feature's bearing / level, mark a map point bad
void hs_frame_set_feature(void* f, int i, const double* normal3, int level)
{
    Feature* ft = static_cast<HsFrame*>(f)->f->mvFeatures[i];
    ft->mNormal = Vector3d(normal3[0], normal3[1], normal3[2]);
    ft->mlevel = level;
}
void hs_mappoint_set_bad(int id, int bad) { if (id >= 0 && id < (int)g_mps.size()) g_mps[id]->SetBad(bad != 0); }
void hs_mappoint_set_pose(int id, const double* p3) { if (id >= 0 && id < (int)g_mps.size()) g_mps[id]->Set_Pose(Vector3d(p3[0], p3[1], p3[2])); }
int hs_mappoint_is_bad(int id) { return (id >= 0 && id < (int)g_mps.size()) ? (g_mps[id]->IsBad() ? 1 : 0) : -1; }

void* hs_keyframe_new(void* f)
{
    Frame* fr = static_cast<HsFrame*>(f)->f.get();
    KeyFrame* kf = new KeyFrame(fr);
    for (size_t i = 0; i < kf->mvFeatures.size(); ++i)
        if (kf->mvFeatures[i]->Mpt) kf->mvFeatures[i]->Mpt->Add_Observation(kf, i);
    return new HsKf{ kf };
}

// UpdateLocalMap + SearchLocalPoints (ref: src/Tracking.cpp:258-313,224): reproject the map points of `kf` into `cur`, then match.
// found[i] (optional) presets MapPoint found-counts to exercise the per-cell ordering. Returns matches; new features are
// appended to cur (read back with hs_frame_get_features).
int hs_search_local_points(void* cam, void* cur, void* kf, const int* found, int* n_reprojected)
{
    try {
        Feature_Alignment& fa = feature_alignment(cam);
        KeyFrame* k = static_cast<HsKf*>(kf)->kf;
        FramePtr c = static_cast<HsFrame*>(cur)->f;
        fa.ResetGrid();
        int nr = 0;
        for (size_t i = 0; i < k->mvFeatures.size(); ++i) {
            MapPoint* mp = k->mvFeatures[i]->Mpt;
            if (!mp) continue;
            if (found) mp->IncreaseFound(found[i] - mp->Get_FoundNums());
            if (fa.ReprojectPoint(c, mp)) nr++;
        }
        if (n_reprojected) *n_reprojected = nr;
        fa.SearchLocalPoints(c);
        return fa.LastMatches();
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// UpdateLocalMap over a list of local keyframes (ref: src/Tracking.cpp:276-305): every map point is reprojected once
// (mLastProjectedFrameId), in keyframe order then feature order; then SearchLocalPoints.
int hs_search_local_points_multi(void* cam, void* cur, void** kfs, int n_kfs, int* n_reprojected)
{
    try {
        Feature_Alignment& fa = feature_alignment(cam);
        FramePtr c = static_cast<HsFrame*>(cur)->f;
        fa.ResetGrid();
        std::unordered_set<MapPoint*> seen;   // UpdateLocalMap's mvLocalMapPoints: each point once (ref: src/Tracking.cpp:258-313)
        int nr = 0;
        for (int q = 0; q < n_kfs; ++q) {
            KeyFrame* k = static_cast<HsKf*>(kfs[q])->kf;
            for (size_t i = 0; i < k->mvFeatures.size(); ++i) {
                MapPoint* mp = k->mvFeatures[i]->Mpt;
                if (!mp || mp->IsBad()) continue;
                if (!seen.insert(mp).second) continue;
                if (fa.ReprojectPoint(c, mp)) nr++;
            }
        }
        if (n_reprojected) *n_reprojected = nr;
        fa.SearchLocalPoints(c);
        return fa.LastMatches();
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

int hs_align2d_single(void* cur, int level, const uint8_t* patch10, int iters, double* px)
{
    try {
        uint8_t p10[100], p8[64];
        std::memcpy(p10, patch10, 100);
        Vector2d v(px[0], px[1]);
        const bool ok = Feature_Alignment::Align2DGaussNewton(static_cast<HsFrame*>(cur)->f, level, p10, p8, iters, v);
        px[0] = v[0]; px[1] = v[1];
        return ok ? 1 : 0;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// the reference's static signature Align2DGaussNewton(const cv::Mat&, uchar*, uchar*, int, Vector2d&) on an arbitrary image
// (ref: include/Feature_alignment.h:85, Test/test_Feature_alignment.cpp:72)
int hs_align2d_image(const uint8_t* img, int w, int h, const uint8_t* patch10, const uint8_t* patch8, int iters, double* px)
{
    try {
        uint8_t p10[100], p8[64];
        std::memcpy(p10, patch10, 100); std::memcpy(p8, patch8, 64);
        const Mat8 m(h, w, img, w);
        Vector2d v(px[0], px[1]);
        const bool ok = Feature_Alignment::Align2DGaussNewton(m, p10, p8, iters, v);
        px[0] = v[0]; px[1] = v[1];
        return ok ? 1 : 0;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// Feature_Alignment::WarpAffine + GetPatchNoBoarder for feature `idx` of a key frame (single-candidate forms, ref: :206-275)
int hs_warp_affine_single(void* cam, void* kf, int idx, const double* A4, int search_level, uint8_t* patch10, uint8_t* patch8)
{
    try {
        Feature_Alignment& fa = feature_alignment(cam);
        KeyFrame* k = static_cast<HsKf*>(kf)->kf;
        Matrix2d A;
        A(0, 0) = A4[0]; A(0, 1) = A4[1]; A(1, 0) = A4[2]; A(1, 1) = A4[3];
        fa.WarpAffine(A, k, k->mvFeatures[idx], search_level, fa.mPatch_WithBoarder);
        fa.GetPatchNoBoarder();
        std::memcpy(patch10, fa.mPatch_WithBoarder, 100);
        std::memcpy(patch8, fa.mPatch, 64);
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

void hs_circle(uint8_t* img, int w, int h, float cx, float cy, int r, int color)
{
    Mat8 m(h, w, img, w);
    circle(m, Point2f(cx, cy), r, (uchar)color);
    std::memcpy(img, m.data, (size_t)w * h);
}

}  // extern "C"
