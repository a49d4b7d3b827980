/*
 * dsdtm_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C++17, no third-party dependency) of the arithmetic of
 * DSDTM's tracking front end, used ONLY as the checker for the CUDA path
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference leg).
 * Nothing under dsdtm_b200/ may include, link or call this file.
 *
 * Pinning status:
 *   - FAST-10 detect/score/nonmax : PINNED against the reference's own Thirdparty/fast sources (oracle/_ref/libfast_ref.so)
 *     and the 167-corner KAT of Thirdparty/fast/test/test.cpp:20,45,52.
 *   - pyrDown, circle, undistortPoints, CLAHE : PINNED against cv2 4.13 goldens (tests/golden/).
 *   - Shi-Tomasi, grid selection, sparse alignment, SolveAffine / WarpAffine / Align2D, ReprojectPoint / SearchLocalPoints /
 *     Get_ClosetObs, GetCloseKeyFrames / UpdateLocalMap, depth lookup, UnProject : PINNED (round 2) against the reference's OWN
 *     translation units, compiled unmodified from /root/reference into oracle/_ref/libdsdtm_ref.so against the stand-in
 *     third-party headers of tests/ref_shim (recipe: oracle/Makefile target ref_dsdtm; comparisons: tests/test_ref_pin.py, bit
 *     equality; travelling golden: tests/golden/refpin.npz).
 *   - what happens INSIDE Eigen / Sophus calls (LDLT solve, reduction association, SE3::exp, quaternion product) and
 *     ceres::Solve (PoseOptimization) : RESTATED FROM THE PUBLISHED ALGORITHMS, UNPINNED -- those libraries are neither under
 *     /root/reference nor installed.
 *
 * All citations "ref:" are paths below /root/reference.
 */
#ifndef DSDTM_ORACLE_H
#define DSDTM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int   width, height;
    float fx, fy, cx, cy, f;  /* float on purpose: ref: src/Camera.cpp:34-39 reads them as float */
} orc_cam;

typedef struct {
    int   x, y, level;
    float score;
} orc_corner; /* ref: include/Feature_detection.h:19-33 (angle dropped: always 0) */

typedef struct {
    float  px[2];      /* Feature::mpx          ref: include/Feature.h:19 */
    int    level;      /* Feature::mlevel       ref: include/Feature.h:20 */
    int    initial;    /* Feature::mbInitial    ref: include/Feature.h:23 */
    double normal[3];  /* Feature::mNormal      ref: include/Feature.h:24 */
    double point_w[3]; /* Feature::Mpt->Get_Pose()  ref: src/MapPoint.cpp:38-43 */
} orc_ref_feat;

typedef struct {
    int    level, iter;
    int    n_pts;       /* visible features in this ComputeResiduals call */
    int    flags;       /* bit0: update accepted, bit1: reverted (chi2 increase), bit2: NaN, bit3: converged |x|<=eps */
    double chi2;        /* chi2New of this iteration (mean squared residual) */
    double x[6];        /* GN step */
} orc_iter_log;

/* pose = {qw,qx,qy,qz,tx,ty,tz} (Sophus::SE3: unit quaternion + translation) */

/* ---- B.1 cv::pyrDown on CV_8UC1, default border (ref: src/Frame.cpp:79) ---- */
void orc_pyrdown_u8(const uint8_t* src, int w, int h, int src_stride, uint8_t* dst /* ((w+1)/2) x ((h+1)/2), dense */);
/* ref: src/Frame.cpp:74-81; out must hold sum of level sizes; offs[levels], ws[levels], hs[levels] filled */
void orc_pyramid(const uint8_t* img, int w, int h, int levels, uint8_t* out, int* offs, int* ws, int* hs);

/* ---- FAST-10 closed forms (ref: Thirdparty/fast/src/fast_10.cpp:9-51, fast_10_score.cpp:21-31,3147, nonmax_3x3.cpp:17-112) ---- */
int  orc_fast10_detect(const uint8_t* img, int w, int h, int stride, int barrier, int16_t* xy /* 2*cap */, int cap);
void orc_fast10_score(const uint8_t* img, int stride, const int16_t* xy, int n, int* scores);
int  orc_fast_nonmax_3x3(const int16_t* xy, const int* scores, int n, int* keep);

/* ---- ref: src/Feature_detection.cpp:157-198 ---- */
float orc_shitomasi(const uint8_t* img, int w, int h, int stride, int u, int v);

/* ---- ref: src/Feature_detection.cpp:69-109 : per-cell best corner over all levels.
 * pyr = packed pyramid as produced by orc_pyramid. occupied may be NULL. cells[grid_rows*grid_cols]. */
void orc_detect_cells(const uint8_t* pyr, const int* offs, const int* ws, const int* hs, int levels,
                      int img_w, int img_h, int cell_size, const uint8_t* occupied, double thr,
                      orc_corner* cells);
/* ---- ref: src/Feature_detection.cpp:111-150 : std::sort + mask-circle selection.
 * cells is sorted in place; mask (img_w x img_h, 255 = free) is painted; returns number of features appended. */
int orc_detect_select(orc_corner* cells, int n_cells, uint8_t* mask, int img_w, int img_h, int cell_size,
                      int max_fts, int n_existing, orc_corner* out);

/* ---- B.2 cv::circle(img, center, r, color, -1) on CV_8UC1 (ref: src/Feature_detection.cpp:146, src/Feature_alignment.cpp:111, src/Frame.cpp:291) */
void orc_circle_fill(uint8_t* img, int w, int h, int stride, int cx, int cy, int r, uint8_t color);
int  orc_cvround(double v);

/* ---- ref: src/Sprase_ImageAlign.cpp:29-60,62-193,240-344 ----
 * ref_pyr/cur_pyr: packed pyramids (orc_pyramid layout). Aligns levels max_level-1 .. min_level.
 * returns n_pts of the last ComputeResiduals call (Run's return value); log may be NULL. */
int orc_sparse_align(const orc_cam* cam,
                     const uint8_t* ref_pyr, const uint8_t* cur_pyr, const int* offs, const int* ws, const int* hs,
                     const orc_ref_feat* feats, int n_feats, const double ref_center[3],
                     const double pose_c2r_in[7], int max_level, int min_level, int max_iters,
                     double pose_c2r_out[7], orc_iter_log* log, int log_cap, int* n_log);

/* ---- ref: src/Feature_alignment.cpp:160-190 ---- */
void orc_solve_affine(const orc_cam* cam, const double kf_center[3], const double ref_point_w[3],
                      const double ref_normal[3], const float ref_px[2], int ref_level,
                      const double pose_c2r[7], double A[4] /* row-major 2x2 */);
/* ---- ref: src/Feature_alignment.cpp:192-204 ---- */
int orc_best_search_level(const double A[4], int max_level);
/* ---- ref: src/Feature_alignment.cpp:206-259 ---- */
void orc_warp_affine(const double A[4], const uint8_t* ref_img, int w, int h, int stride,
                     const float ref_px[2], int ref_level, int search_level, uint8_t patch10[100]);
/* ---- ref: src/Feature_alignment.cpp:261-275 ---- */
void orc_patch_no_border(const uint8_t patch10[100], uint8_t patch8[64]);
/* ---- ref: src/Feature_alignment.cpp:318-417 ; returns converged flag; n_iters_out optional ---- */
int orc_align2d(const uint8_t* cur_img, int w, int h, int stride, const uint8_t patch10[100],
                const uint8_t patch8[64], int max_iters, double px[2], int* n_iters_out);

/* Eigen LDLT<Matrix6d>::solve restatement (App. B.4); H row-major 6x6 (lower triangle read) */
void orc_ldlt6_solve(const double H[36], const double b[6], double x[6]);
/* ---- SE3 helpers (Sophus non-templated semantics, App. B.3) ---- */
void orc_se3_exp(const double x[6], double pose[7]);
void orc_se3_mul(const double a[7], const double b[7], double out[7]);
void orc_se3_inv(const double a[7], double out[7]);
void orc_se3_act(const double a[7], const double p[3], double out[3]);
/* ref: src/Camera.cpp:173-178 + src/Frame.cpp:83-92 : float Pixel2Camera(px,1.0) widened, then normalize() */
void orc_feature_normal(const orc_cam* cam, const float px[2], double normal[3]);

/* ---- the map walk in front of FindMatchDirect ---- */
/* ref: src/Camera.cpp:187-193 Camera::IsInImage (cvRound of the float pixel, integer division of the image size) */
int orc_is_in_image(const orc_cam* cam, float x, float y, int boundary, int level);
/* ref: src/Feature_alignment.cpp:54-69 ReprojectPoint + src/Frame.cpp:318-323 World2Pixel; returns 1 when the point lands
 * inside IsInImage(px, 8); px is always written, *cell only when in the image */
int orc_reproject_point(const orc_cam* cam, const double pose_cur_c2w[7], const double point_w[3], int cell_size,
                        int grid_cols, double px[2], int* cell);
/* ref: src/MapPoint.cpp:133-174 Get_ClosetObs; kf_centers = n_obs x 3 in mObservations iteration order; *best = chosen
 * observation (0 when n_obs > 0 and no cosine is positive, -1 when n_obs == 0); returns the bool of the reference */
int orc_closest_obs(const double cur_center[3], const double point_w[3], const double* kf_centers, int n_obs, int* best);

/* ---- keyframe ingest ---- */
/* cv::undistortPoints(src, dst, K, dist, noArray(), K) as called at ref: src/Frame.cpp:121-122 : K and dist are CV_32F
 * (ref: src/Camera.cpp:53-68) widened to double, 5 fixed-point iterations (OpenCV default criteria), float in / float out.
 * dist = {k1, k2, p1, p2, k3}. PINNED against cv2 4.13 goldens (tests/golden/undistort_cv2.npz). */
void orc_undistort_points(const orc_cam* cam, const float dist[5], const float* src, int n, float* dst);
/* ref: src/Tracking.cpp:56 depthImg.convertTo(CV_32F, 1.0f/mDepthScale) on CV_16U input: float(src) * float(alpha) */
void orc_depth_convert(const uint16_t* depth, int n, float depth_scale, float* out);
/* ref: src/Frame.cpp:200-224 Get_FeatureDetph(Point2f): cvRound, centre then the 4-neighbourhood in the order
 * (-1,0) (0,-1) (1,0) (0,1), -1.0 when all are 0. Reads outside the image (undefined in the reference) count as 0. */
float orc_feature_depth(const float* depth, int w, int h, const float px[2]);
/* ref: src/Frame.cpp:152-157 UnProject: T_c2w^-1 * Pixel2Camera(px, d) (src/Camera.cpp:173-178 evaluated in float) */
void orc_unproject(const orc_cam* cam, const double pose_c2w[7], const float px[2], float d, double out[3]);

/* cv::CLAHE::apply on CV_8UC1 (cv::createCLAHE(clip_limit, Size(tiles_x, tiles_y)); ref: Test/test_Feature_detection.cpp:85-86,
 * Test/test_Euroc.cpp:64, Test/test_Optimizer.cpp:75 use (3.0, 8x8) in front of the Frame constructor). OpenCV's algorithm
 * (imgproc/src/clahe.cpp, not under /root/reference) restated for image sizes divisible by the tile grid (no border padding).
 * PINNED against cv2 4.13 goldens (tests/golden/clahe_cv2.npz). */
void orc_clahe(const uint8_t* src, int w, int h, double clip_limit, int tiles_x, int tiles_y, uint8_t* dst);

/* ---- SURVEY 8f-1 (caller side): Tracking::GetCloseKeyFrames + the ranking of UpdateLocalMap (ref: src/Tracking.cpp:315-345,261-277).
 * Key frame k owns points_w[pt_begin[k] .. + pt_count[k]) (null / zero points = {0,0,0}); kf_t = Get_Pose().translation() per key frame.
 * visible[k], dist[k] (0 when not visible); local[] = the first max_local visible key frames by distance (stable); returns their count. */
int orc_close_keyframes(const orc_cam* cam, const double pose_cur_c2w[7], const int* pt_begin, const int* pt_count,
                        const double* kf_t, int n_kfs, const double* points_w, int max_local, uint8_t* visible,
                        double* dist, int* local);

/* ---- Optimizer::PoseOptimization (ref: src/Optimizer.cpp:20-101, include/Optimizer.h:129-258) ----
 * ceres::Solve with the reference's configuration restated (trust-region Levenberg-Marquardt, DENSE_SCHUR on the single pose
 * block, CauchyLoss(1.0), PoseLocalParameterization, 100 iterations; see the .cpp). PARITY UNPINNED (Ceres is neither under
 * /root/reference nor installed; the reference holds no golden value). normals = Feature::mNormal (3 per observation),
 * levels = Feature::mlevel, points_w = MapPoint::Get_Pose(); res_norm[k] = GetReprojectReidual()[k] at the final pose. */
enum { ORC_BA_FUNCTION_TOL = 0, ORC_BA_PARAMETER_TOL = 1, ORC_BA_GRADIENT_TOL = 2, ORC_BA_NO_CONVERGENCE = 3,
       ORC_BA_FAILURE = 4, ORC_BA_MIN_RADIUS = 5, ORC_BA_NO_RESIDUALS = 6 };
typedef struct {
    int    iterations;      /* minimizer iterations after iteration 0 (summary.iterations.size() - 1) */
    int    termination;     /* ORC_BA_* */
    int    n_successful;    /* accepted steps */
    int    pad;
    double initial_cost, final_cost;
} orc_ba_summary;
int orc_pose_optimization(int n_obs, const double* normals, const int* levels, const double* points_w,
                          const double pose_in[7], int max_iters, double pose_out[7], double* res_norm,
                          orc_ba_summary* summary);

/* Batched CPU driver used only for the cpu_baseline / reference bench arm: runs pyramid(cur) +
 * sparse align + align2d for pairs [0,n) with n_threads std::threads. Layout documented in bench.py. */
int orc_pair_batch(const orc_cam* cam, int levels, const uint8_t* ref_pyrs, const uint8_t* cur_imgs, int n_pairs,
                   const orc_ref_feat* feats, int feats_per_pair, const int* n_feats,
                   const double* ref_centers, const double* poses_in,
                   int max_level, int min_level, int max_iters,
                   const uint8_t* patches10, const double* patch_px, const int* patch_level, int patches_per_pair,
                   int align_iters, int n_threads,
                   double* poses_out, int* n_tracked, double* patch_px_out, uint8_t* patch_conv);

/* per-candidate record of the batched refinement chain (layout of dsdtm_reproj in include/dsdtm_gpu.h) */
typedef struct {
    double  px_proj[2];
    double  px[2];
    int32_t cell, obs, flags, level;
} orc_reproj;
int orc_pair_batch_map(const orc_cam* cam, int levels, int cell_size, const uint8_t* ref_pyrs, const uint8_t* cur_imgs, int n_pairs,
                       const orc_ref_feat* feats, int feats_per_pair, const int* n_feats, const double* ref_centers,
                       const double* poses_ref, const double* poses_in, int max_level, int min_level, int max_iters,
                       int points_per_pair, int max_search_level, int align_iters, int n_threads,
                       double* poses_out, int* n_tracked, orc_reproj* reproj);

#ifdef __cplusplus
}
#endif
#endif
