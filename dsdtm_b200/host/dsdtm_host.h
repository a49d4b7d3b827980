// dsdtm_host.h -- C++ host side of the drop-in: the reference's class interfaces for the tracking front end
// (namespace DSDTM; same class names, method names, argument meaning and error behaviour as called from Tracking,
// ref: src/Tracking.cpp:31-37,204,224,260,299,415-416,476; src/Initializer.cpp:44; src/Frame.cpp:55), with bodies that
// marshal into the C-ABI of include/dsdtm_gpu.h. No OpenCV / Eigen / Sophus: the few value types the interfaces need
// (Point2f, Vector2d/3d, Matrix2d, SE3, an 8-bit Mat) are provided here with the semantics the hot path relies on.
// A maintainer of the reference keeps cv::Mat / Eigen / Sophus and replaces only the method bodies: see INTEGRATION.md.
//
// There is no CPU fallback: every compute method goes through dsdtm_* and throws std::runtime_error if the GPU
// context cannot be created (the reference has no error channel on these methods either; it would simply crash).
#ifndef DSDTM_HOST_H
#define DSDTM_HOST_H

#include <cmath>
#include <cstdint>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "dsdtm_gpu.h"

namespace DSDTM {

typedef unsigned char uchar;

// ------------------------------------------------------------------------------------------ value types
struct Point2f { float x, y; Point2f(float x_ = 0, float y_ = 0) : x(x_), y(y_) {} };

struct Vector2d {
    double v[2];
    Vector2d(double a = 0, double b = 0) { v[0] = a; v[1] = b; }
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    Vector2d operator-(const Vector2d& o) const { return Vector2d(v[0] - o.v[0], v[1] - o.v[1]); }
    Vector2d operator*(double s) const { return Vector2d(v[0] * s, v[1] * s); }
    Vector2d operator/(double s) const { return Vector2d(v[0] / s, v[1] / s); }
};

struct Vector3d {
    double v[3];
    Vector3d(double a = 0, double b = 0, double c = 0) { v[0] = a; v[1] = b; v[2] = c; }
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    Vector3d operator-(const Vector3d& o) const { return Vector3d(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vector3d operator+(const Vector3d& o) const { return Vector3d(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    Vector3d operator*(double s) const { return Vector3d(v[0] * s, v[1] * s, v[2] * s); }
    double dot(const Vector3d& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    void normalize() { const double n = norm(); v[0] /= n; v[1] /= n; v[2] /= n; }
    bool isZero() const { return v[0] == 0 && v[1] == 0 && v[2] == 0; }
};
inline Vector3d operator*(double s, const Vector3d& a) { return a * s; }

struct Matrix2d {
    double m[2][2];
    Matrix2d() { m[0][0] = m[1][1] = 1; m[0][1] = m[1][0] = 0; }
    double& operator()(int r, int c) { return m[r][c]; }
    double operator()(int r, int c) const { return m[r][c]; }
    double determinant() const { return m[0][0] * m[1][1] - m[1][0] * m[0][1]; }
};

// Sophus::SE3 (non-templated): unit quaternion + translation; pose7 = {qw,qx,qy,qz,tx,ty,tz}
class SE3 {
public:
    SE3();
    explicit SE3(const double pose7[7]);
    static SE3 exp(const double x[6]);
    SE3 inverse() const;
    SE3 operator*(const SE3& o) const;
    Vector3d operator*(const Vector3d& p) const;
    Vector3d translation() const { return Vector3d(p[4], p[5], p[6]); }
    const double* data() const { return p; }
private:
    double p[7];
};

// 8-bit single-channel image with cv::Mat-like sharing semantics (level 0 of a pyramid aliases the caller's image)
class Mat8 {
public:
    int rows = 0, cols = 0, step = 0;
    uchar* data = nullptr;
    Mat8() {}
    Mat8(int rows_, int cols_, uchar fill);
    Mat8(int rows_, int cols_, const uchar* src, int src_step);   // copies
    bool empty() const { return data == nullptr; }
    uchar& at(int y, int x) { return data[(size_t)y * step + x]; }
    uchar at(int y, int x) const { return data[(size_t)y * step + x]; }
    void release() { store.reset(); data = nullptr; rows = cols = step = 0; }
private:
    std::shared_ptr<std::vector<uchar>> store;
};

class GpuSlot;
// Stands in for std::vector<cv::Mat> mvImg_Pyr (ref: include/Frame.h:113). Level 0 is the caller's image; levels 1.. are built on the
// device and only the device stages read them, so their host copies are fetched from HBM on the FIRST read of mvImg_Pyr[l] (viewer,
// debugging) instead of in every Frame constructor.
class PyrLevels {
public:
    size_t size() const { return m.size(); }
    void resize(size_t n) { m.resize(n); pending.assign(n, 0); }
    Mat8& operator[](size_t l) { if (pending[l]) Fetch(l); return m[l]; }
    const Mat8& operator[](size_t l) const { if (pending[l]) Fetch(l); return m[l]; }
    void Set(size_t l, const Mat8& img) { m[l] = img; pending[l] = 0; }
    void Bind(const std::shared_ptr<GpuSlot>& gpu);             // levels 1.. now live in that slot's pyramid: host copies on demand
    bool Pending(size_t l) const { return pending[l] != 0; }
private:
    void Fetch(size_t l) const;
    mutable std::vector<Mat8> m;
    mutable std::vector<char> pending;
    std::shared_ptr<GpuSlot> gpu;
};

int cvRound(double v);                                       // round-half-to-even like OpenCV on x86
void circle(Mat8& img, Point2f center, int radius, uchar color);   // cv::circle(img, c, r, color, -1)

// ------------------------------------------------------------------------------------------ Config / Camera
// ref: include/Config.h:28-31 -- dotted-key YAML subset ("%YAML:1.0", comments, key: value)
class Config {
public:
    static void setParameterFile(const std::string& path);
    static void Set(const std::string& key, const std::string& value);
    static bool Has(const std::string& key);
    template <typename T> static T Get(const std::string& key);
    static void Clear();
private:
    static std::map<std::string, std::string>& table();
};

class Camera {   // ref: include/Camera.h:137-163, src/Camera.cpp:32-60,167-193
public:
    Camera();    // reads Camera.* from Config
    Vector2d Camera2Pixel(const Vector3d& p) const;
    Vector3d Pixel2Camera(const Point2f& px, const float& depth) const;      // float evaluation (Q8)
    Vector3d Pixel2Camera(const Vector2d& px, const float& depth) const;
    bool IsInImage(const Point2f pt, int boundary = 0, int level = 0) const;
    float mf, mfx, mfy, mcx, mcy;
    float mk1, mk2, mp1, mp2, mk3;       // ref: src/Camera.cpp:41-45 (0 when the config has no Camera.k1 .. k3)
    int mwidth, mheight;
};
typedef std::shared_ptr<Camera> CameraPtr;

// ------------------------------------------------------------------------------------------ data model (boundary types only)
class Frame;
class KeyFrame;
class MapPoint;

struct Feature {   // ref: include/Feature.h:16-51
    Frame* mframe;
    Point2f mpx;
    int mlevel;
    bool mbInitial;
    Vector3d mNormal;
    MapPoint* Mpt;
    Feature(Frame* frame, const Point2f& px, int level) : mframe(frame), mpx(px), mlevel(level), mbInitial(false), mNormal(0, 0, 0), Mpt(nullptr) {}
    void SetPose(MapPoint* mp) { Mpt = mp; mbInitial = true; }
};
typedef std::vector<Feature*> Features;

class MapPoint {   // ref: include/MapPoint.h, src/MapPoint.cpp:38-43,126-181 (the members the hot path reads)
public:
    explicit MapPoint(const Vector3d& pose) : mPose(pose) {}
    Vector3d Get_Pose() const { std::unique_lock<std::mutex> l(mMutexPos); return mPose; }
    void Set_Pose(const Vector3d& p) { { std::unique_lock<std::mutex> l(mMutexPos); mPose = p; } Touch(); }
    bool IsBad() const { return mbBad; }
    void SetBad(bool b) { mbBad = b; Touch(); Mirror(); }
    // new: row of this point in the device-resident map store (-1 = not there yet) and the change queue Map::Sync() drains
    int mStoreId = -1;
    // dense host mirrors of (found count, bad flag) by store row: SearchLocalPoints orders ~1000 candidates per frame by found count
    // and would otherwise chase a pointer per candidate into objects scattered over the heap
    struct DenseState { std::vector<int> found; std::vector<unsigned char> bad; };
    static DenseState* sDense;
    void Mirror() { if (mStoreId >= 0 && sDense && (size_t)mStoreId < sDense->found.size()) { sDense->found[mStoreId] = mnFound; sDense->bad[mStoreId] = mbBad ? 1 : 0; } }
    void Touch();
    static void DrainTouched(std::vector<MapPoint*>& out);
    int Get_FoundNums() const { return mnFound; }
    void IncreaseFound(int n = 1) { mnFound += n; Mirror(); }
    // ref: src/MapPoint.cpp:183-197 (+ SetBadFlag :91-109: the outlier flag and the observation list; the Map / KeyFrame
    // bookkeeping it also does lives outside the path)
    void EraseFound(int n = 1) { mnFound -= n; if (mnFound <= 0) { mbBad = true; mObservations.clear(); Touch(); } Mirror(); }
    void Add_Observation(KeyFrame* kf, size_t idx) { mObservations[kf] = idx; }
    bool Get_ClosetObs(const Frame* frame, Feature*& feature, KeyFrame*& kf) const;
    std::map<KeyFrame*, size_t> Get_Observations() const { return mObservations; }   // ref: src/MapPoint.cpp (copy under mMutexObs)
    // the adapters' snapshot walks the observations in place (same std::map order) instead of copying the map per candidate
    template <class F> void ForEachObservation(F&& f) const { for (const auto& o : mObservations) f(o.first, o.second); }
    unsigned long mLastProjectedFrameId = (unsigned long)-1;     // ref: include/MapPoint.h (UpdateLocalMap's per-frame dedupe)
    bool HasObservation(KeyFrame* kf, size_t idx) const { auto it = mObservations.find(kf); return it != mObservations.end() && it->second == idx; }
private:
    bool mTouched = false;
    mutable std::mutex mMutexPos;
    Vector3d mPose;
    bool mbBad = false;
    int mnFound = 1;
    std::map<KeyFrame*, size_t> mObservations;
};

class GpuSlot;   // device residency of one pyramid (RAII; shared between a Frame and the KeyFrame made from it)

class Frame {   // ref: include/Frame.h:18-136, src/Frame.cpp:48-92,167-174,286-298,318-323
public:
    Frame(CameraPtr cam, const Mat8& gray, double timestamp = 0);
    virtual ~Frame();
    void ComputeImagePyramid(const Mat8 image, PyrLevels& pyr);              // GPU: dsdtm_frame_upload_pyramid
    void Add_Feature(Feature* f, bool normal = true);
    void Add_MapPoint(MapPoint* mp) { mvMapPoints.push_back(mp); }
    void Set_Pose(const SE3& pose);
    SE3 Get_Pose() const { return mT_c2w; }
    Vector3d Get_CameraCnt() const { return mOw; }
    Vector2d World2Pixel(const Vector3d& p) const;
    void Set_Mask();
    // RGB-D keyframe path (ref: src/Frame.cpp:94-157,200-224; src/Tracking.cpp:56,412-464). SetDepth replaces the CV_32F
    // cv::Mat mDepthImg: the raw 16-bit image goes to HBM, scaling by 1/depth_scale happens at the lookups.
    void SetDepth(const uint16_t* depth, int stride_bytes, float depth_scale);
    void UndistortFeatures();                          // GPU: dsdtm_keyframe_lift over all features (also fills the cache below)
    float Get_FeatureDetph(const Feature* feature);    // value of the reference's lookup, from the cache of UndistortFeatures
    Vector3d UnProject(const Point2f px, const float d);
    const std::vector<dsdtm_lifted>& Lifted() const { return mLifted; }   // new: per-feature device results (depth, world point)

    CameraPtr mCamera;
    unsigned long mlId;                // ref: include/Frame.h:106-107
    static unsigned long mlNextId;
    double mdCloTimestamp;
    Mat8 mColorImg;
    PyrLevels mvImg_Pyr;
    Features mvFeatures;
    std::vector<MapPoint*> mvMapPoints;
    Mat8 mImgMask, mDynamicMask;
    int mPyra_levels, mMin_Dist;
    std::shared_ptr<GpuSlot> mGpu;     // new member: where the pyramid lives in HBM
    bool mHasDepth = false;
    float mDepthScale = 1.0f;
    std::vector<dsdtm_lifted> mLifted;
protected:
    SE3 mT_c2w;
    Vector3d mOw;
};
typedef std::shared_ptr<Frame> FramePtr;

class KeyFrame {   // ref: include/Keyframe.h, src/Keyframe.cpp:10-22 (copy of the frame's images, features and pose)
public:
    explicit KeyFrame(Frame* frame);
    SE3 Get_Pose() const { return mT_c2w; }
    void Set_Pose(const SE3& p);
    Vector3d Get_CameraCnt() const { return mOw; }
    PyrLevels mvImg_Pyr;
    Features mvFeatures;
    std::shared_ptr<GpuSlot> mGpu;
    // position of this key frame in the table of the snapshot being built (valid while mSnapEpoch == the snapshot's number)
    mutable unsigned long long mSnapEpoch = 0;
    mutable int mSnapIndex = -1;
    int mStoreRow = -1, mStoreSlot = -1;                        // new: row in the device-resident map store, slot recorded there
private:
    SE3 mT_c2w;
    Vector3d mOw;
};

// ------------------------------------------------------------------------------------------ the hot-path classes
struct Corner {   // ref: include/Feature_detection.h:19-33
    int x, y, level;
    float score, angle;
    Corner(int x_, int y_, float score_, int level_, float angle_) : x(x_), y(y_), level(level_), score(score_), angle(angle_) {}
    bool operator<(const Corner& c) const { return c.score < score; }
};
typedef std::vector<Corner> Corners;

class Feature_detector {   // ref: include/Feature_detection.h:36-74
public:
    Feature_detector();
    void Set_ExistingFeatures(const Features& features);
    void Set_ExistingFeatures(const std::vector<Point2f>& features);
    void detect(Frame* frame, const double detection_threshold, const bool tFirst = true);
    void ResetGrid();
    int mImg_height, mImg_width, mCell_size, mPyr_levels, mGrid_rows, mGrid_cols, mMax_fts;
    std::vector<bool> mvGrid_occupy;
};

class Sprase_ImgAlign {   // ref: include/Sprase_ImageAlign.h:20-69
public:
    Sprase_ImgAlign(int tMaxLevel, int tMinLevel, int tMaxIterators);
    void Reset();                                             // ref: src/Sprase_ImageAlign.cpp:22-27
    int Run(FramePtr tCurFrame, FramePtr tRefFrame);
    // the 2 x 6 Jacobian of the normalised projection w.r.t. the pose (ref: :169-193), row-major; host math, kept for API parity
    void GetJocabianBA(const Vector3d& tPoint, double J[12]) const;
    const SE3& Get_T_c2r() const { return mT_c2r; }           // mT_c2r of the last Run
    // last run's Gauss-Newton trace (new: the reference only prints; used by the parity tests)
    const std::vector<dsdtm_iter_log>& LastLog() const { return mLog; }
    void EnableLog(bool on) { mWantLog = on; }                // off by default: the trace costs an 18 KB read-back per frame
protected:
    int mnMaxLevel, mnMinLevel, mnMaxIterators, mnMinfts;
    SE3 mT_c2r;
    std::vector<dsdtm_iter_log> mLog;
    bool mWantLog = false;
    std::vector<dsdtm_ref_feat> mFeats;                       // reused across frames
};

static const int mHalf_PatchSize = 4;   // ref: include/Feature_alignment.h:21

class Feature_Alignment {   // ref: include/Feature_alignment.h:23-99
public:
    struct Candidate {
        MapPoint* mMpPoint;
        Vector2d mPx;
        Candidate(MapPoint* pt, Vector2d px) : mMpPoint(pt), mPx(px) {}
    };
    typedef std::list<Candidate> Cell;
    explicit Feature_Alignment(CameraPtr camera);
    ~Feature_Alignment();
    void ResetGrid();
    bool ReprojectPoint(FramePtr tFrame, MapPoint* tMPoint);
    void SearchLocalPoints(FramePtr tFrame);
    // single-candidate forms kept for API parity (each is a batch of one on the GPU)
    bool FindMatchDirect(const MapPoint* tMpPoint, const FramePtr tFrame, Vector2d& tPt, int& tLevel);
    Matrix2d SolveAffineMatrix(KeyFrame* tReferKframe, const FramePtr tCurFrame, Feature* tReferFeature, const MapPoint* tMpPoint);
    int GetBestSearchLevel(Matrix2d tAffineMat, int tMaxLevel);
    // ref: :206-259 / :261-275; single-candidate forms (a batch of one on the GPU). tPatchLarger = 10 x 10 bytes, the 8 x 8 interior
    // is written to mPatch by GetPatchNoBoarder as in the reference.
    void WarpAffine(const Matrix2d tA_c2r, KeyFrame* tReferKframe, Feature* tRefFeature, const int tSearchLevel, uchar* tPatchLarger);
    void GetPatchNoBoarder();
    static bool CellComparator(Candidate& c1, Candidate& c2);     // ref: :123-126
    uchar mPatch[2 * mHalf_PatchSize * 2 * mHalf_PatchSize];
    uchar mPatch_WithBoarder[(2 * mHalf_PatchSize + 2) * (2 * mHalf_PatchSize + 2)];
    static bool Align2DGaussNewton(const FramePtr tCurFrame, int tLevel, uchar* tPatch_WithBoarder, uchar* tPatch, int MaxIters, Vector2d& tCurPx);
    // the reference's own static signature (ref: include/Feature_alignment.h:85; Test/test_Feature_alignment.cpp:72 hands it an arbitrary
    // image): tCurImg must have the size of one pyramid level of the configured camera (the device frame pool is geometry-fixed); it is
    // uploaded into a scratch slot for the call. Throws std::invalid_argument for any other size.
    static bool Align2DGaussNewton(const Mat8& tCurImg, uchar* tPatch_WithBoarder, uchar* tPatch, int MaxIters, Vector2d& tCurPx);
    int LastMatches() const { return mLastMatches; }
    // new: Tracking::UpdateLocalMap's device path (dsdtm_store_track) hands its per-candidate records here instead of calling
    // ReprojectPoint once per map point; SearchLocalPoints then only orders the cells and replays the greedy selection
    void SetFused(const Frame* frame, std::vector<dsdtm_store_cand>&& cands, const std::vector<MapPoint*>* points, const MapPoint::DenseState* dense);
    bool HasFused(const Frame* frame) const { return mFusedFrame == frame && mFusedPoints != nullptr; }
private:
    void SearchFused(FramePtr frame);
    void MaterializeFused();            // a later ReprojectPoint call: turn the records back into the reference's cell lists
    const Frame* mFusedFrame = nullptr;
    std::vector<dsdtm_store_cand> mFused;
    const std::vector<MapPoint*>* mFusedPoints = nullptr;
    const MapPoint::DenseState* mFusedDense = nullptr;
    struct Prepared;   // one candidate after the host-side map walk
    bool Prepare(const MapPoint* mp, const FramePtr frame, const Vector2d& px, Prepared& out);
    CameraPtr mCam;
    std::vector<Cell*> mCells;
    int mCell_size, mGrid_Cols, mGrid_Rows, mMax_pts, mPyr_levels;
    int mLastMatches = 0;
};

// ref: include/Map.h:23-66 -- the part Tracking::UpdateLocalMap reads, plus the bookkeeping of the device-resident map table
// (one row per key frame, one row per entry of its mvMapPoints): a new key frame appends its rows, MarkMoved() queues the rows of
// a key frame whose pose or points a bundle adjustment changed; Sync() uploads what is pending (called by GetCloseKeyFrames).
class Map {
public:
    ~Map();                                                      // forgets the device-resident rows of this map
    void AddKeyFrame(KeyFrame* kf);                              // ref: src/Map.cpp AddKeyFrame
    std::vector<KeyFrame*> GetAllKeyFrames() const { return mvKeyFrames; }   // insertion order (the reference: std::set = address order)
    int ReturnKeyFramesSize() const { return (int)mvKeyFrames.size(); }
    void MarkMoved(KeyFrame* kf);                                // new: LocalBundleAdjustment's hook
    void Sync();                                                 // new: pending rows -> dsdtm_map_table_upload
    KeyFrame* Row(int i) const { return mvKeyFrames[i]; }
    const std::vector<MapPoint*>& StorePoints() const { return mStorePoints; }    // map-store row -> MapPoint
    const MapPoint::DenseState& Dense() const { return mDense; }
    unsigned long long Version() const { return mVersion; }      // changes whenever Sync() / SyncStore() changed a device table
    void SyncStore();                                            // new key frames / points, touched points, moved or re-uploaded key frames
private:
    std::vector<MapPoint*> mStorePoints;
    MapPoint::DenseState mDense;
    unsigned long long mVersion = 1;
    int mStoreKfs = 0;
    std::vector<KeyFrame*> mStoreMoved;
    struct Rows { int pt_begin, pt_count; };
    void Pack(KeyFrame* kf, dsdtm_map_kf& row, std::vector<double>& pts) const;
    std::vector<KeyFrame*> mvKeyFrames;
    std::vector<Rows> mRows;
    std::map<KeyFrame*, int> mIndex;
    std::vector<int> mPending;                                   // rows to (re)write
    int mUploadedKfs = 0, mPoints = 0;
};

// ref: include/Tracking.h, src/Tracking.cpp:257-345 -- ONLY the local-map selection of the tracking thread: GetCloseKeyFrames and
// UpdateLocalMap with the members they use. (The state machine around them is a caller of the path and stays the reference's.)
class Tracking {
public:
    Tracking(CameraPtr cam, Map* map);
    ~Tracking();
    void SetCurrentFrame(FramePtr f) { mCurrentFrame = f; }
    void GetCloseKeyFrames(const Frame* tFrame, std::list<std::pair<KeyFrame*, double>>& tClose_kfs) const;   // ref: :315-345
    void UpdateLocalMap();                                                                                      // ref: :257-313
    FramePtr mCurrentFrame;
    Map* mMap;
    Feature_Alignment* mFeature_Alignment;
    std::vector<KeyFrame*> mvpLocalKeyFrames;
    std::map<MapPoint*, KeyFrame*> mvpLocalMapPoints;
    int LastReprojected() const { return mLastReprojected; }
    // new: false = the literal host loop of the reference (one ReprojectPoint per map point); true (default) = dsdtm_store_track
    static bool sUseStore;
    static bool sSpeculate;                                   // Run also runs the local-map stage (one synchronisation per frame instead of two)
private:
    void UpdateLocalMapOnDevice();
    CameraPtr mCam;
    int mLastReprojected = 0;
};

// ref: include/Optimizer.h:23-44, src/Optimizer.cpp:20-101 -- the entry Tracking calls right after SearchLocalPoints
// (ref: src/Tracking.cpp:236). GPU: dsdtm_pose_optimize (the reference's ceres::Solve configuration, one launch).
class Optimizer {
public:
    static void PoseOptimization(FramePtr tCurFrame, int tIterations = 100);
    static const dsdtm_ba_summary& LastSummary();              // new: the reference only has Ceres' (commented-out) report
    static const std::vector<double>& LastResiduals();         // new: GetReprojectReidual() of the last solve
};

// ------------------------------------------------------------------------------------------ GPU runtime
// One dsdtm_ctx per process/thread of tracking, created lazily from Config (Camera.*, Gpu.Device, Gpu.MaxFrames).
class GpuRuntime {
public:
    static GpuRuntime& Instance();
    static void Shutdown();
    dsdtm_ctx* ctx() { return mCtx; }
    // uploads + builds the pyramid, returns the residency handle; levels_out (optional) receives the host copies of levels 1..
    std::shared_ptr<GpuSlot> Upload(const Mat8& level0, uint8_t* levels_out = nullptr);
    int Resident(const std::shared_ptr<GpuSlot>& s);             // slot index, re-uploading from the host copy if it was evicted
    // a slot holding ONE level image (no pyramid), for single calls on images that are not frames; released with the handle
    std::shared_ptr<GpuSlot> UploadLevel(const Mat8& img, int level);
    int levels() const { return mLevels; }
    void Release(int slot);
    // epoch = one batched call that resolves several slots before using them: detects a slot being recycled in between
    void BeginEpoch() { mEpochStart = mClock; mEpochEvicted = false; }
    bool SlotsStillValid() const { return !mEpochEvicted; }
    void SetDepthOwner(const Frame* f) { mDepthOwner = f; }      // depth slot 0 holds this frame's depth image
    // Run + UpdateLocalMap as one submission (dsdtm_track_frame_store): Tracking registers its map; Sprase_ImgAlign::Run then also runs the
    // local-map stage with the pose it has just found and parks the records here; Tracking::UpdateLocalMap takes them when frame, pose
    // and map version still match (otherwise it makes its own call). Same results either way (the device composes the pose exactly as
    // the host does; the adapter checks the bits).
    struct Speculation {
        const Frame* frame = nullptr;
        double pose[7];
        int32_t rows[16];
        int n_local = 0, n_out = 0;
        unsigned long long map_version = 0;
        std::vector<dsdtm_store_cand> cands;
    };
    void SetTrackingMap(Map* m) { mTrackingMap = m; mSpec.frame = nullptr; }
    Map* TrackingMap() const { return mTrackingMap; }
    Speculation& Spec() { return mSpec; }
    const Frame* DepthOwner() const { return mDepthOwner; }
    ~GpuRuntime();
private:
    GpuRuntime();
    int Acquire(GpuSlot* owner);
    dsdtm_ctx* mCtx = nullptr;
    int mLevels = 0, mMaxFrames = 0;
    std::vector<GpuSlot*> mOwner;      // per slot
    std::vector<unsigned long long> mStamp;
    unsigned long long mClock = 0, mEpochStart = 0;
    bool mEpochEvicted = false;
    const Frame* mDepthOwner = nullptr;
    Map* mTrackingMap = nullptr;
    Speculation mSpec;
    friend class GpuSlot;
};

class GpuSlot {
public:
    GpuSlot(const Mat8& img) : host(img) {}
    ~GpuSlot();
    Mat8 host;        // level 0 (shared buffer): lets the runtime re-upload after an eviction
    int slot = -1;
};

}  // namespace DSDTM
#endif
