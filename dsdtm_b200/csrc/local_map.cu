// local_map.cu -- SURVEY 8f-1: the map walk in front of FindMatchDirect, on flat snapshots of the local map.
//
//   kf_pose_kernel   : one thread per keyframe, T_c2r = T_cur * T_kf^-1 (ref: src/Feature_alignment.cpp:181) with Sophus
//                      semantics (inverse = conjugate + rotated negated translation; product normalises the quaternion).
//   local_map_kernel : one thread per map point:
//                        Frame::World2Pixel + Camera::IsInImage(px, 8) + cell index   (ref: src/Feature_alignment.cpp:54-69)
//                        MapPoint::Get_ClosetObs                                       (ref: src/MapPoint.cpp:133-174)
//                        IsInImage(ref px / 2^level, 5, level)                         (ref: src/Feature_alignment.cpp:138)
//                      and writes the dsdtm_candidate that candidate_prep_kernel (align2d.cu) consumes, so the chain
//                      reproject -> observation -> affine -> warp -> Align2D never returns to the host.
// fp64 in the reference's operation order, non-contracted (__dmul_rn / __dadd_rn), so every value that reaches the integer
// decisions (cvRound, cell index, strict '>' on the cosine) is the one the CPU computes. A few thousand points per frame:
// latency-bound by construction (one short dependent chain per thread); the point of the stage is removing the host round
// trip between sparse alignment and Align2D, not bandwidth.
#include "ctx.cuh"
#include "se3_exact.cuh"

namespace dsdtm {

namespace {

struct LmArgs {
    const dsdtm_kf_view* kfs; int n_kfs;
    const dsdtm_obs* obs;
    const dsdtm_map_point* pts; int n_pts;
    double pose_cur[7]; double cur_center[3];
    const double* pose_dev;     // optional: {pose_c2w[7], centre[3]} produced on the device (dsdtm_track_frame); overrides the two above
    float fx, fy, cx, cy;
    int width, height, cell_size, grid_cols;
    double* kf_pose;            // n_kfs x 7
    dsdtm_candidate* cand;      // n_pts
    dsdtm_reproj* out;          // n_pts
};

__global__ void __launch_bounds__(64) kf_pose_kernel(const LmArgs a)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n_kfs) return;
    double inv[7], T[7], pc[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) pc[j] = a.pose_dev ? a.pose_dev[j] : a.pose_cur[j];
    se3_inv_exact(a.kfs[k].pose_c2w, inv);
    se3_mul_exact(pc, inv, T);
    double* o = a.kf_pose + 7 * k;
#pragma unroll
    for (int j = 0; j < 7; ++j) o[j] = T[j];
}

// Camera::IsInImage (ref: src/Camera.cpp:187-193): cvRound(float) is round-half-even; integer division of the image size.
// cvRound = cvtss2si: NaN (a map point at the camera centre projects to 0/0) gives INT_MIN and fails the test, whereas
// __float2int_rn(NaN) is 0 -- hence the explicit guard (pinned against the compiled reference: test_is_in_image_and_nan).
__device__ __forceinline__ bool in_image(float x, float y, int boundary, int level, int width, int height)
{
    if (!(x == x) || !(y == y)) return false;
    const int rx = __float2int_rn(x), ry = __float2int_rn(y);
    return rx >= boundary && rx < width / (1 << level) - boundary && ry >= boundary && ry < height / (1 << level) - boundary;
}

__global__ void __launch_bounds__(128) local_map_kernel(const LmArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_pts) return;
    const dsdtm_map_point mp = a.pts[i];
    const double P0 = mp.point_w[0], P1 = mp.point_w[1], P2 = mp.point_w[2];
    // Frame::World2Pixel (ref: src/Frame.cpp:318-323): Camera2Pixel(T_c2w * P) = fx * X / Z + cx
    double pc[7], cc[3];
#pragma unroll
    for (int j = 0; j < 7; ++j) pc[j] = a.pose_dev ? a.pose_dev[j] : a.pose_cur[j];
#pragma unroll
    for (int j = 0; j < 3; ++j) cc[j] = a.pose_dev ? a.pose_dev[7 + j] : a.cur_center[j];
    double q0, q1, q2;
    qrot_exact(pc, P0, P1, P2, q0, q1, q2);
    q0 = __dadd_rn(q0, pc[4]); q1 = __dadd_rn(q1, pc[5]); q2 = __dadd_rn(q2, pc[6]);
    const double u = __dadd_rn(__ddiv_rn(__dmul_rn((double)a.fx, q0), q2), (double)a.cx);
    const double v = __dadd_rn(__ddiv_rn(__dmul_rn((double)a.fy, q1), q2), (double)a.cy);
    int flags = 0, cell = -1, best = -1;
    if (in_image(__double2float_rn(u), __double2float_rn(v), 8, 0, a.width, a.height)) {          // ref: :58
        flags |= DSDTM_LM_IN_IMAGE;
        cell = (int)__ddiv_rn(v, (double)a.cell_size) * a.grid_cols + (int)__ddiv_rn(u, (double)a.cell_size);   // ref: :60-61
    }
    // MapPoint::Get_ClosetObs: direction point -> current camera, then the observation with the largest cosine (strict >)
    double f0 = __dsub_rn(cc[0], P0), f1 = __dsub_rn(cc[1], P1), f2 = __dsub_rn(cc[2], P2);
    double n = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(f0, f0), __dmul_rn(f1, f1)), __dmul_rn(f2, f2)));
    f0 = __ddiv_rn(f0, n); f1 = __ddiv_rn(f1, n); f2 = __ddiv_rn(f2, n);
    double best_cos = 0.0;
    if (mp.obs_count > 0) best = mp.obs_begin;                                                        // min_it = begin()
    for (int j = 0; j < mp.obs_count; ++j) {
        const double* O = a.kfs[a.obs[mp.obs_begin + j].kf].center;
        double r0 = __dsub_rn(O[0], P0), r1 = __dsub_rn(O[1], P1), r2 = __dsub_rn(O[2], P2);
        const double rn = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(r0, r0), __dmul_rn(r1, r1)), __dmul_rn(r2, r2)));
        r0 = __ddiv_rn(r0, rn); r1 = __ddiv_rn(r1, rn); r2 = __ddiv_rn(r2, rn);
        const double c = __dadd_rn(__dadd_rn(__dmul_rn(r0, f0), __dmul_rn(r1, f1)), __dmul_rn(r2, f2));
        if (c > best_cos) { best_cos = c; best = mp.obs_begin + j; }
    }
    dsdtm_candidate cd;
    cd.ref_slot = -1; cd.ref_level = 0;
    cd.px[0] = u; cd.px[1] = v;
    if (best >= 0) {
        if (!(best_cos < 0.5)) flags |= DSDTM_LM_OBS_OK;                                              // ref: MapPoint.cpp:170
        const dsdtm_obs ob = a.obs[best];
        const float sc = (float)(1 << ob.level);
        if (in_image(__fdiv_rn(ob.px[0], sc), __fdiv_rn(ob.px[1], sc), 5, ob.level, a.width, a.height)) flags |= DSDTM_LM_REF_OK;
        const int all = DSDTM_LM_IN_IMAGE | DSDTM_LM_OBS_OK | DSDTM_LM_REF_OK;
        if ((flags & all) == all) {
            const dsdtm_kf_view& kf = a.kfs[ob.kf];
            cd.ref_slot = kf.slot; cd.ref_level = ob.level;
            cd.ref_px[0] = ob.px[0]; cd.ref_px[1] = ob.px[1];
#pragma unroll
            for (int k = 0; k < 3; ++k) { cd.ref_normal[k] = ob.normal[k]; cd.ref_point_w[k] = ob.point_w[k]; cd.kf_center[k] = kf.center[k]; }
#pragma unroll
            for (int k = 0; k < 7; ++k) cd.pose_c2r[k] = a.kf_pose[7 * ob.kf + k];
        }
    }
    a.cand[i] = cd;
    dsdtm_reproj r;
    r.px_proj[0] = u; r.px_proj[1] = v; r.px[0] = u; r.px[1] = v;
    r.cell = cell; r.obs = best; r.flags = flags; r.level = -1;
    a.out[i] = r;
}

// Sprase_ImgAlign::Run's last line on the device (ref: src/Sprase_ImageAlign.cpp:57 cur.Set_Pose(T_c2r * ref.Get_Pose()); src/Frame.cpp:167-174
// Set_Pose: mOw = inverse().translation()): out = {pose_c2w[7], centre[3]} for the local-map kernels of the same call.
struct ComposeArgs { const double* t_c2r; double pose_ref[7]; double* out; };
__global__ void compose_pose_kernel(const ComposeArgs a)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double T[7], Tc[7], inv[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) T[j] = a.t_c2r[j];
    se3_mul_exact(T, a.pose_ref, Tc);
    se3_inv_exact(Tc, inv);
#pragma unroll
    for (int j = 0; j < 7; ++j) a.out[j] = Tc[j];
    a.out[7] = inv[4]; a.out[8] = inv[5]; a.out[9] = inv[6];
}

// ---------------------------------------------------------------------------------------------------------------
// The refinement chain of a BATCH of independent frame pairs (bench / sweep: dsdtm_batch_run flags bit 1, dsdtm_track_batch_e2e):
// every pair's local map is its reference frame seen as ONE key frame, whose features with map points are the candidates --
// Tracking::TrackWithLocalMap right after a key frame (ref: src/Tracking.cpp:219-224,257-313). One thread per (pair, feature):
//   cur.Set_Pose(T_c2r * ref.Get_Pose())                       ref: src/Sprase_ImageAlign.cpp:57, src/Frame.cpp:167-174
//   ReprojectPoint: World2Pixel + IsInImage(px, 8) + cell      ref: src/Feature_alignment.cpp:54-69
//   Get_ClosetObs with the single observation (cos >= 0.5)     ref: src/MapPoint.cpp:133-174
//   IsInImage(ref px / 2^level, 5, level)                      ref: src/Feature_alignment.cpp:138
//   T_c2r' = cur.Get_Pose() * kf.Get_Pose().inverse()          ref: src/Feature_alignment.cpp:181
// and the dsdtm_candidate that candidate_prep_kernel consumes. Same non-contracted arithmetic as local_map_kernel / kf_pose_kernel /
// compose_pose_kernel (the pose composition is repeated per thread: ~150 flops, no extra launch, no extra pass over the poses).
struct PcArgs {
    const int* ref_slots; const dsdtm_ref_feat* feats; int feat_stride; const int* n_feats;
    const double* centers; const double* poses_ref; const double* poses_c2r;
    int pair0, n_pairs, ppp;
    float fx, fy, cx, cy;
    int width, height, cell_size, grid_cols;
    dsdtm_candidate* cand; dsdtm_reproj* out;
};

__global__ void __launch_bounds__(128) pair_candidates_kernel(const PcArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_pairs * a.ppp) return;
    const int pair = a.pair0 + t / a.ppp, j = t % a.ppp;
    const size_t i = (size_t)pair * a.ppp + j;
    dsdtm_candidate cd;
    cd.ref_slot = -1; cd.ref_level = 0; cd.px[0] = 0.0; cd.px[1] = 0.0;
    dsdtm_reproj r;
    r.px_proj[0] = r.px_proj[1] = r.px[0] = r.px[1] = 0.0; r.cell = -1; r.obs = -1; r.flags = 0; r.level = -1;
    const dsdtm_ref_feat ft = (j < a.n_feats[pair]) ? a.feats[(size_t)pair * a.feat_stride + j] : dsdtm_ref_feat{};
    if (j < a.n_feats[pair] && ft.initial) {                         // features without a map point are not in KeyFrame::GetMapPoints()
        double Tr[7], Tc2r[7], pc[7], inv[7], Tck[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) { Tr[k] = a.poses_ref[7 * pair + k]; Tc2r[k] = a.poses_c2r[7 * pair + k]; }
        se3_mul_exact(Tc2r, Tr, pc);                                 // cur pose (c2w)
        se3_inv_exact(pc, inv);                                      // cur centre = inverse().translation()
        const double cc0 = inv[4], cc1 = inv[5], cc2 = inv[6];
        se3_inv_exact(Tr, inv);
        se3_mul_exact(pc, inv, Tck);                                 // T_cur * T_kf^-1
        const double P0 = ft.point_w[0], P1 = ft.point_w[1], P2 = ft.point_w[2];
        double q0, q1, q2;
        qrot_exact(pc, P0, P1, P2, q0, q1, q2);
        q0 = __dadd_rn(q0, pc[4]); q1 = __dadd_rn(q1, pc[5]); q2 = __dadd_rn(q2, pc[6]);
        const double u = __dadd_rn(__ddiv_rn(__dmul_rn((double)a.fx, q0), q2), (double)a.cx);
        const double v = __dadd_rn(__ddiv_rn(__dmul_rn((double)a.fy, q1), q2), (double)a.cy);
        int flags = 0, cell = -1;
        if (in_image(__double2float_rn(u), __double2float_rn(v), 8, 0, a.width, a.height)) {
            flags |= DSDTM_LM_IN_IMAGE;
            cell = (int)__ddiv_rn(v, (double)a.cell_size) * a.grid_cols + (int)__ddiv_rn(u, (double)a.cell_size);
        }
        const double* O = a.centers + 3 * (size_t)pair;
        double f0 = __dsub_rn(cc0, P0), f1 = __dsub_rn(cc1, P1), f2 = __dsub_rn(cc2, P2);
        double n = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(f0, f0), __dmul_rn(f1, f1)), __dmul_rn(f2, f2)));
        f0 = __ddiv_rn(f0, n); f1 = __ddiv_rn(f1, n); f2 = __ddiv_rn(f2, n);
        double r0 = __dsub_rn(O[0], P0), r1 = __dsub_rn(O[1], P1), r2 = __dsub_rn(O[2], P2);
        const double rn = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(r0, r0), __dmul_rn(r1, r1)), __dmul_rn(r2, r2)));
        r0 = __ddiv_rn(r0, rn); r1 = __ddiv_rn(r1, rn); r2 = __ddiv_rn(r2, rn);
        const double cosv = __dadd_rn(__dadd_rn(__dmul_rn(r0, f0), __dmul_rn(r1, f1)), __dmul_rn(r2, f2));
        const double best_cos = (cosv > 0.0) ? cosv : 0.0;                                               // tMin_angle starts at 0, strict >
        if (!(best_cos < 0.5)) flags |= DSDTM_LM_OBS_OK;
        const float sc = (float)(1 << ft.level);
        if (in_image(__fdiv_rn(ft.px[0], sc), __fdiv_rn(ft.px[1], sc), 5, ft.level, a.width, a.height)) flags |= DSDTM_LM_REF_OK;
        cd.px[0] = u; cd.px[1] = v;
        const int all = DSDTM_LM_IN_IMAGE | DSDTM_LM_OBS_OK | DSDTM_LM_REF_OK;
        if ((flags & all) == all) {
            cd.ref_slot = a.ref_slots[pair]; cd.ref_level = ft.level;
            cd.ref_px[0] = ft.px[0]; cd.ref_px[1] = ft.px[1];
#pragma unroll
            for (int k = 0; k < 3; ++k) { cd.ref_normal[k] = ft.normal[k]; cd.ref_point_w[k] = ft.point_w[k]; cd.kf_center[k] = O[k]; }
#pragma unroll
            for (int k = 0; k < 7; ++k) cd.pose_c2r[k] = Tck[k];
        }
        r.px_proj[0] = u; r.px_proj[1] = v; r.px[0] = u; r.px[1] = v;
        r.cell = cell; r.obs = j; r.flags = flags;
    }
    a.cand[i] = cd;
    a.out[i] = r;
}

// After Align2D: fold (refined px * 2^level, level, converged) into the per-point records so that ONE copy returns everything.
__global__ void __launch_bounds__(128) local_map_finalize_kernel(dsdtm_reproj* __restrict__ out, const double* __restrict__ px,
                                                                 const int* __restrict__ level, const uint8_t* __restrict__ conv, int n, int i0)
{
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i0 + n) return;
    const int L = level[i];
    if (L < 0) return;                                              // not aligned: px stays the projection, level -1
    const double sc = (double)(1 << L);                             // ref: :154 tPt = tCurPx * (1 << tBestLevel) (exact)
    out[i].px[0] = __dmul_rn(px[2 * i], sc); out[i].px[1] = __dmul_rn(px[2 * i + 1], sc);
    out[i].level = L;
    if (conv[i]) out[i].flags |= DSDTM_LM_CONVERGED;
}

}  // namespace

cudaError_t launch_local_map_finalize(dsdtm_ctx* c, int n_pts, cudaStream_t s, dsdtm_reproj* out_d, int i0)
{
    local_map_finalize_kernel<<<(n_pts + 127) / 128, 128, 0, s>>>(out_d ? out_d : c->lm_reproj_d, c->patch_px_d, c->patch_level_d, c->patch_conv_d, n_pts, i0);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_pair_candidates(dsdtm_ctx* c, int pair0, int n_pairs, int feat_stride, int ppp, cudaStream_t s)
{
    PcArgs a;
    a.ref_slots = c->ref_slots_d; a.feats = c->feats_d; a.feat_stride = feat_stride; a.n_feats = c->n_feats_d;
    a.centers = c->centers_d; a.poses_ref = c->poses_ref_d; a.poses_c2r = c->poses_out_d;
    a.pair0 = pair0; a.n_pairs = n_pairs; a.ppp = ppp;
    a.fx = c->cam.fx; a.fy = c->cam.fy; a.cx = c->cam.cx; a.cy = c->cam.cy;
    a.width = c->cam.width; a.height = c->cam.height; a.cell_size = c->prm.cell_size; a.grid_cols = c->grid_cols;
    a.cand = c->cand_d; a.out = c->pair_reproj_d;
    const long long n = (long long)n_pairs * ppp;
    pair_candidates_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_compose_pose(dsdtm_ctx* c, const double* t_c2r_d, const double pose_ref_c2w[7], double* out10_d, cudaStream_t s)
{
    ComposeArgs a;
    a.t_c2r = t_c2r_d; a.out = out10_d;
    for (int k = 0; k < 7; ++k) a.pose_ref[k] = pose_ref_c2w[k];
    compose_pose_kernel<<<1, 32, 0, s>>>(a);
    c->launches++;
    return cudaGetLastError();
}

cudaError_t launch_local_map(dsdtm_ctx* c, const double pose_cur[7], const double cur_center[3], int n_kfs, int n_pts, cudaStream_t s,
                             const double* pose_dev)
{
    LmArgs a;
    a.pose_dev = pose_dev;
    a.kfs = c->lm_kfs_d; a.n_kfs = n_kfs; a.obs = c->lm_obs_d; a.pts = c->lm_pts_d; a.n_pts = n_pts;
    for (int k = 0; k < 7; ++k) a.pose_cur[k] = pose_cur ? pose_cur[k] : 0.0;
    for (int k = 0; k < 3; ++k) a.cur_center[k] = cur_center ? cur_center[k] : 0.0;
    a.fx = c->cam.fx; a.fy = c->cam.fy; a.cx = c->cam.cx; a.cy = c->cam.cy;
    a.width = c->cam.width; a.height = c->cam.height; a.cell_size = c->prm.cell_size; a.grid_cols = c->grid_cols;
    a.kf_pose = c->lm_pose_d; a.cand = c->cand_d; a.out = c->lm_reproj_d;
    if (n_kfs > 0) kf_pose_kernel<<<(n_kfs + 63) / 64, 64, 0, s>>>(a);
    local_map_kernel<<<(n_pts + 127) / 128, 128, 0, s>>>(a);
    c->launches += 2;
    return cudaGetLastError();
}

// ---- Tracking::GetCloseKeyFrames on the device-resident map table (ref: src/Tracking.cpp:315-345, src/Frame.cpp:300-311).
// One warp per key frame: the lanes test 32 of its map points at a time (same non-contracted projection as above, boundary 0) and
// the warp stops at the first chunk with a visible point -- the reference's early break, 32 points wide. The work is one read of
// the point rows of the key frames that are NOT close (24 B per point) and one chunk of those that are.
namespace {
struct CkArgs {
    const dsdtm_map_kf* kfs; const double* pts; int n_kfs;
    double pose[7];
    float fx, fy, cx, cy; int width, height;
    uint8_t* visible; double* dist;
};

__global__ void __launch_bounds__(128) close_kf_kernel(const CkArgs a)
{
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= a.n_kfs) return;
    const int lane = threadIdx.x & 31;
    const dsdtm_map_kf kf = a.kfs[k];
    bool found = false;
    // software-pipelined: the next chunk's rows are requested before this chunk's vote, so the early-exit loop is not a chain of
    // dependent global loads (ncu: 9.9 long-scoreboard stalls per issued instruction without it)
    const double* rows = a.pts + 3 * (size_t)kf.pt_begin;
    double n0 = 0.0, n1 = 0.0, n2 = 0.0;
    if (lane < kf.pt_count) { n0 = rows[3 * lane]; n1 = rows[3 * lane + 1]; n2 = rows[3 * lane + 2]; }
    for (int base = 0; base < kf.pt_count && !found; base += 32) {
        const int i = base + lane;
        const double P0 = n0, P1 = n1, P2 = n2;
        const int j = i + 32;
        if (j < kf.pt_count) { n0 = rows[3 * (size_t)j]; n1 = rows[3 * (size_t)j + 1]; n2 = rows[3 * (size_t)j + 2]; }
        bool vis = false;
        if (i < kf.pt_count) {
            if (!(P0 == 0.0 && P1 == 0.0 && P2 == 0.0)) {                     // ref: :324-328
                double q0, q1, q2;
                qrot_exact(a.pose, P0, P1, P2, q0, q1, q2);
                q0 = __dadd_rn(q0, a.pose[4]); q1 = __dadd_rn(q1, a.pose[5]); q2 = __dadd_rn(q2, a.pose[6]);
                if (!(q2 < 0.0)) {                                              // ref: src/Frame.cpp:303
                    const double u = __dadd_rn(__ddiv_rn(__dmul_rn((double)a.fx, q0), q2), (double)a.cx);
                    const double v = __dadd_rn(__ddiv_rn(__dmul_rn((double)a.fy, q1), q2), (double)a.cy);
                    vis = in_image(__double2float_rn(u), __double2float_rn(v), 0, 0, a.width, a.height);
                }
            }
        }
        found = __any_sync(0xffffffffu, vis);
    }
    if (lane == 0) {
        a.visible[k] = found ? 1 : 0;
        const double d0 = __dsub_rn(a.pose[4], kf.t[0]), d1 = __dsub_rn(a.pose[5], kf.t[1]), d2 = __dsub_rn(a.pose[6], kf.t[2]);
        a.dist[k] = found ? sqrt(__dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2))) : 0.0;   // ref: :332
    }
}
}  // namespace

cudaError_t launch_close_keyframes(dsdtm_ctx* c, const double pose_cur[7], int n_kfs, cudaStream_t s)
{
    CkArgs a;
    a.kfs = c->mt_kfs_d; a.pts = c->mt_pts_d; a.n_kfs = n_kfs;
    for (int k = 0; k < 7; ++k) a.pose[k] = pose_cur[k];
    a.fx = c->cam.fx; a.fy = c->cam.fy; a.cx = c->cam.cx; a.cy = c->cam.cy; a.width = c->cam.width; a.height = c->cam.height;
    a.visible = c->mt_vis_d; a.dist = c->mt_dist_d;
    close_kf_kernel<<<(n_kfs + 3) / 4, 128, 0, s>>>(a);
    c->launches++;
    return cudaGetLastError();
}

}  // namespace dsdtm
