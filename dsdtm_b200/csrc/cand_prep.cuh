// cand_prep.cuh -- SolveAffineMatrix + GetBestSearchLevel (ref: src/Feature_alignment.cpp:160-204) for one candidate, as a device
// function shared by candidate_prep_kernel (align2d.cu) and the map-store candidate kernel (local_map.cu), which calls it for the
// candidate it has just built instead of a separate launch. fp64 non-contracted in the reference's operation order.
#pragma once
#include "ctx.cuh"

namespace dsdtm {

struct CpArgs {
    const dsdtm_candidate* cand; int n; int cur_slot; int max_search_level;
    int i0;                      // first candidate of this launch (chunked batches)
    const int* cur_slots; int ppp;   // batched chain: candidate i belongs to pair i / ppp, whose current frame is cur_slots[pair]
    float fx, fy, cx, cy;
    double* A; float* ref_px; int* meta; int* patch_level; int* patch_slot; double* px_in;
};

__device__ __forceinline__ void cp_qrot(const double* q, double v0, double v1, double v2, double& o0, double& o1, double& o2)
{
    double uv0 = __dsub_rn(__dmul_rn(q[2], v2), __dmul_rn(q[3], v1));
    double uv1 = __dsub_rn(__dmul_rn(q[3], v0), __dmul_rn(q[1], v2));
    double uv2 = __dsub_rn(__dmul_rn(q[1], v1), __dmul_rn(q[2], v0));
    uv0 = __dadd_rn(uv0, uv0); uv1 = __dadd_rn(uv1, uv1); uv2 = __dadd_rn(uv2, uv2);
    const double c0 = __dsub_rn(__dmul_rn(q[2], uv2), __dmul_rn(q[3], uv1));
    const double c1 = __dsub_rn(__dmul_rn(q[3], uv0), __dmul_rn(q[1], uv2));
    const double c2 = __dsub_rn(__dmul_rn(q[1], uv1), __dmul_rn(q[2], uv0));
    o0 = __dadd_rn(__dadd_rn(v0, __dmul_rn(q[0], uv0)), c0);
    o1 = __dadd_rn(__dadd_rn(v1, __dmul_rn(q[0], uv1)), c1);
    o2 = __dadd_rn(__dadd_rn(v2, __dmul_rn(q[0], uv2)), c2);
}

// one candidate: SolveAffineMatrix + GetBestSearchLevel, written straight into the inputs of warp_affine_kernel / align2d_kernel
__device__ __forceinline__ void candidate_prep_one(const CpArgs& a, int i, const dsdtm_candidate& c, int cur_slot)
{
    if (c.ref_slot < 0) {                                           // rejected before FindMatchDirect's arithmetic (local_map.cu)
        a.meta[3 * i] = -1; a.meta[3 * i + 1] = 0; a.meta[3 * i + 2] = 0;
        a.patch_level[i] = -1; a.patch_slot[i] = cur_slot;
        a.px_in[2 * i] = c.px[0]; a.px_in[2 * i + 1] = c.px[1];
        return;
    }
    const double fx = (double)a.fx, fy = (double)a.fy, cx = (double)a.cx, cy = (double)a.cy;
    const int HalfLarger = 5;                                       // mHalf_PatchSize + 1
    // ref: :167  P = ||O_kf - P_w|| * mNormal
    const double d0 = __dsub_rn(c.kf_center[0], c.ref_point_w[0]), d1 = __dsub_rn(c.kf_center[1], c.ref_point_w[1]), d2 = __dsub_rn(c.kf_center[2], c.ref_point_w[2]);
    const double nrm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2)));
    const double P0 = __dmul_rn(nrm, c.ref_normal[0]), P1 = __dmul_rn(nrm, c.ref_normal[1]), P2 = __dmul_rn(nrm, c.ref_normal[2]);
    // ref: :171-172 (float px + int, evaluated in float, then widened)
    const float step = (float)(HalfLarger * (1 << c.ref_level));
    const double pxU0 = (double)__fadd_rn(c.ref_px[0], step), pxU1 = (double)c.ref_px[1];
    const double pxV0 = (double)c.ref_px[0], pxV1 = (double)__fadd_rn(c.ref_px[1], step);
    // ref: :174-179  Pixel2Camera(Vector2d, 1.0f) -> normalize -> rescale to depth P.z
    double U0 = __ddiv_rn(__dmul_rn(1.0, __dsub_rn(pxU0, cx)), fx), U1 = __ddiv_rn(__dmul_rn(1.0, __dsub_rn(pxU1, cy)), fy), U2 = 1.0;
    double V0 = __ddiv_rn(__dmul_rn(1.0, __dsub_rn(pxV0, cx)), fx), V1 = __ddiv_rn(__dmul_rn(1.0, __dsub_rn(pxV1, cy)), fy), V2 = 1.0;
    double n = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(U0, U0), __dmul_rn(U1, U1)), __dmul_rn(U2, U2)));
    U0 = __ddiv_rn(U0, n); U1 = __ddiv_rn(U1, n); U2 = __ddiv_rn(U2, n);
    n = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(V0, V0), __dmul_rn(V1, V1)), __dmul_rn(V2, V2)));
    V0 = __ddiv_rn(V0, n); V1 = __ddiv_rn(V1, n); V2 = __ddiv_rn(V2, n);
    double sc = __ddiv_rn(P2, U2); U0 = __dmul_rn(U0, sc); U1 = __dmul_rn(U1, sc); U2 = __dmul_rn(U2, sc);
    sc = __ddiv_rn(P2, V2);        V0 = __dmul_rn(V0, sc); V1 = __dmul_rn(V1, sc); V2 = __dmul_rn(V2, sc);
    // ref: :181-184  project the three points with T = T_cur * T_kf^-1 (Camera2Pixel: (fx*X)/Z + cx)
    double q0, q1, q2;
    cp_qrot(c.pose_c2r, P0, P1, P2, q0, q1, q2);
    q0 = __dadd_rn(q0, c.pose_c2r[4]); q1 = __dadd_rn(q1, c.pose_c2r[5]); q2 = __dadd_rn(q2, c.pose_c2r[6]);
    const double c0 = __dadd_rn(__ddiv_rn(__dmul_rn(fx, q0), q2), cx), c1 = __dadd_rn(__ddiv_rn(__dmul_rn(fy, q1), q2), cy);
    cp_qrot(c.pose_c2r, U0, U1, U2, q0, q1, q2);
    q0 = __dadd_rn(q0, c.pose_c2r[4]); q1 = __dadd_rn(q1, c.pose_c2r[5]); q2 = __dadd_rn(q2, c.pose_c2r[6]);
    const double cu0 = __dadd_rn(__ddiv_rn(__dmul_rn(fx, q0), q2), cx), cu1 = __dadd_rn(__ddiv_rn(__dmul_rn(fy, q1), q2), cy);
    cp_qrot(c.pose_c2r, V0, V1, V2, q0, q1, q2);
    q0 = __dadd_rn(q0, c.pose_c2r[4]); q1 = __dadd_rn(q1, c.pose_c2r[5]); q2 = __dadd_rn(q2, c.pose_c2r[6]);
    const double cv0 = __dadd_rn(__ddiv_rn(__dmul_rn(fx, q0), q2), cx), cv1 = __dadd_rn(__ddiv_rn(__dmul_rn(fy, q1), q2), cy);
    // ref: :186-187
    const double A00 = __ddiv_rn(__dsub_rn(cu0, c0), (double)HalfLarger), A10 = __ddiv_rn(__dsub_rn(cu1, c1), (double)HalfLarger);
    const double A01 = __ddiv_rn(__dsub_rn(cv0, c0), (double)HalfLarger), A11 = __ddiv_rn(__dsub_rn(cv1, c1), (double)HalfLarger);
    // ref: :192-204 GetBestSearchLevel
    int L = 0;
    double D = __dsub_rn(__dmul_rn(A00, A11), __dmul_rn(A10, A01));
    while (D > 3.0 && L < a.max_search_level) { L++; D = __dmul_rn(D, 0.25); }
    a.A[4 * i] = A00; a.A[4 * i + 1] = A01; a.A[4 * i + 2] = A10; a.A[4 * i + 3] = A11;
    a.ref_px[2 * i] = c.ref_px[0]; a.ref_px[2 * i + 1] = c.ref_px[1];
    a.meta[3 * i] = c.ref_slot; a.meta[3 * i + 1] = c.ref_level; a.meta[3 * i + 2] = L;
    a.patch_level[i] = L; a.patch_slot[i] = cur_slot;
    const double inv = 1.0 / (double)(1 << L);                      // ref: :150 tPt / (1 << level): exact power of two
    a.px_in[2 * i] = __dmul_rn(c.px[0], inv); a.px_in[2 * i + 1] = __dmul_rn(c.px[1], inv);
}


}  // namespace dsdtm
