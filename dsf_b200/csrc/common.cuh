// dsf_b200 - shared device/host helpers.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dsf_b200.h"

#define NV DSF_NV
#define NVW DSF_NVW
#define NJ DSF_NJ
#define NJOUT DSF_NJOUT
#define KP 148           // blend-shape K: 10 beta + 135 pose feature, padded to a multiple of 4
#define NP 2336          // 778*3 = 2334 padded to a multiple of 4
#define BLEND_KPAD 160    // KP rounded up to the 32-float k block of the tensor-core GEMM
#define BLEND_NPAD 2432   // NP rounded up to the 128-column tile
#define BLEND_BN_BWD 160  // KP rounded up to a legal UMMA N
#define BLEND_SPLITS 16   // maximum split-K factor of the backward contraction (workspace slots)
#define RJ_STRIDE 32     // per joint: R[9] Gr[9] Gt[3] J[3] At[3] ang[3] pad[2]
#define RJ_R 0
#define RJ_GR 9
#define RJ_GT 18
#define RJ_J 21
#define RJ_AT 24
#define RJ_ANG 27
#define VIEW DSF_VIEW_STRIDE

// workspace layout (floats per hand)
#define WS_X 0                            // GEMM operand [beta | Rs - I | 0], split into tf32 hi (BLEND_KPAD) and lo (BLEND_KPAD)
#define WS_VP (WS_X + 2 * BLEND_KPAD)
#define WS_RJ (WS_VP + NP)
#define WS_GVP (WS_RJ + NJ * RJ_STRIDE)   // skinning cotangent g_vposed, tf32 hi part ...
#define WS_GVPL (WS_GVP + NP)             // ... and lo part (operand of the backward blend GEMM, written pre-split)
#define WS_GA (WS_GVPL + NP)
#define WS_GX (WS_GA + NJ * 12)          // BLEND_SPLITS split-K partials of g_X, KP floats each
#define WS_PER_HAND (WS_GX + BLEND_SPLITS * KP)
// the MANO scratch holds whole groups of 8 hands (the tensor maps of the blend GEMM address rows in groups of 8)
#define WS_HANDS(b) (((long)(b) + 7) & ~7L)

struct DsfMano {
    float* BTh;    // (BLEND_NPAD, BLEND_KPAD) basis^T, tf32 hi part, zero padded   (forward B operand)
    float* BTl;    //                          basis^T, lo part
    float* Bh;     // (BLEND_BN_BWD, NP)       basis, hi part                         (backward B operand)
    float* Bl;     //                          basis, lo part
    float* vt;     // (NP) v_template, flat, zero padded
    float* W;      // (778,16) skin weights
    float* comp;   // (45,45)
    float* mean;   // (45)
    float* Jt;     // (16,3)   J_regressor^T v_template
    float* JS;     // (10,16,3) J_regressor^T shapedirs
    int* jr_ptr;   // (17) CSR over the 16 regressed joints
    int* jr_idx;
    float* jr_w;
    int jr_nnz;
    int* wj_ptr;   // (17) CSR of the skin weights, joint-major
    int* wj_idx;
    float* wj_w;
    int* wv_ptr;   // (779) CSR of the skin weights, vertex-major (a vertex follows a handful of joints, not 16)
    int2* wv_ent;  // (nnz) {joint, float bits of the weight}
    int* faces;    // (n_faces,3)
    unsigned int* faces_packed;   // (n_faces) i0 | i1 << 10 | i2 << 20
    unsigned short* face_order;   // (n_faces) face ids, largest rest-pose area first
    int* vf_ptr;                  // (780) CSR vertex -> incident face corners
    unsigned short* vf_ent;       // (3 n_faces) face << 2 | corner
    int n_faces;
    float* coll_mask;  // (66,66)
    int parents[NJ];
    int level[NJ];
    int maxlevel;
    int device;
};

struct ChainTopo {
    int parents[NJ];
    int level[NJ];
    int nchild[NJ];     // number of children of each joint
    int maxlevel;
};

// ---- error plumbing --------------------------------------------------------------------------
void dsf_set_error(const char* fmt, ...);
void dsf_count_launch(int n);
void dsf_reset_launch_count();

#define DSF_CHECK_CUDA(expr)                                                             \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            dsf_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,          \
                          cudaGetErrorString(_e));                                       \
            return DSF_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define DSF_CHECK_LAUNCH()                                                               \
    do {                                                                                 \
        dsf_count_launch(1);                                                             \
        DSF_CHECK_CUDA(cudaGetLastError());                                              \
    } while (0)

#define DSF_REQUIRE(cond, msg)                                                           \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            dsf_set_error("%s:%d bad argument: %s", __FILE__, __LINE__, msg);            \
            return DSF_ERR_BAD_ARG;                                                      \
        }                                                                                \
    } while (0)

// ---- small device helpers --------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum 16 per-lane values over the warp with 16 shuffles instead of 80: at every butterfly step a lane hands
// over the half of its values the partner will own and keeps the other half.  Returns the total of value
// (lane >> 1) (every total lands on a pair of lanes); deterministic order.
__device__ __forceinline__ float warp_sum16_scatter(const float* v, int lane) {
    float a[8], b[4], c[2];
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float keep = h4 ? v[8 + i] : v[i], give = h4 ? v[i] : v[8 + i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = h3 ? a[4 + i] : a[i], give = h3 ? a[i] : a[4 + i];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = h2 ? b[2 + i] : b[i], give = h2 ? b[i] : b[2 + i];
        c[i] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
    }
    const float keep = h1 ? c[1] : c[0], give = h1 ? c[0] : c[1];
    const float d = keep + __shfl_xor_sync(0xffffffffu, give, 2);
    return d + __shfl_xor_sync(0xffffffffu, d, 1);
}

// y = M(3x3 row-major) * x
__device__ __forceinline__ void mat3_vec(const float* M, const float* x, float* y) {
    y[0] = M[0] * x[0] + M[1] * x[1] + M[2] * x[2];
    y[1] = M[3] * x[0] + M[4] * x[1] + M[5] * x[2];
    y[2] = M[6] * x[0] + M[7] * x[1] + M[8] * x[2];
}
// y = M^T * x
__device__ __forceinline__ void mat3T_vec(const float* M, const float* x, float* y) {
    y[0] = M[0] * x[0] + M[3] * x[1] + M[6] * x[2];
    y[1] = M[1] * x[0] + M[4] * x[1] + M[7] * x[2];
    y[2] = M[2] * x[0] + M[5] * x[1] + M[8] * x[2];
}
// C = A * B
__device__ __forceinline__ void mat3_mul(const float* A, const float* B, float* C) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}
// C = A^T * B
__device__ __forceinline__ void mat3T_mul(const float* A, const float* B, float* C) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            C[3 * r + c] = A[r] * B[c] + A[3 + r] * B[3 + c] + A[6 + r] * B[6 + c];
}
// C = A * B^T
__device__ __forceinline__ void mat3_mulT(const float* A, const float* B, float* C) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            C[3 * r + c] = A[3 * r] * B[3 * c] + A[3 * r + 1] * B[3 * c + 1] + A[3 * r + 2] * B[3 * c + 2];
}

// Rotation from a (w,x,y,z) quaternion that is re-normalised first (mano_layer.py:697-718).
__device__ __forceinline__ void quat_to_mat(const float* q, float* R, float* qn_out, float* m_out) {
    float m = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    float w = q[0] / m, x = q[1] / m, y = q[2] / m, z = q[3] / m;
    if (qn_out) { qn_out[0] = w; qn_out[1] = x; qn_out[2] = y; qn_out[3] = z; *m_out = m; }
    float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
    float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
    R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;   R[2] = 2 * wy + 2 * xz;
    R[3] = 2 * wz + 2 * xy;   R[4] = w2 - x2 + y2 - z2; R[5] = 2 * yz - 2 * wx;
    R[6] = 2 * xz - 2 * wy;   R[7] = 2 * wx + 2 * yz;   R[8] = w2 - x2 - y2 + z2;
}

// d(loss)/d(q) for R = quat_to_mat(q), given gR
__device__ __forceinline__ void quat_to_mat_bwd(const float* q, const float* g, float* gq) {
    float qn[4], m;
    float Rtmp[9];
    quat_to_mat(q, Rtmp, qn, &m);
    float w = qn[0], x = qn[1], y = qn[2], z = qn[3];
    float gn[4];
    gn[0] = 2 * w * (g[0] + g[4] + g[8]) + 2 * (-z * g[1] + y * g[2] + z * g[3] - x * g[5] - y * g[6] + x * g[7]);
    gn[1] = 2 * x * (g[0] - g[4] - g[8]) + 2 * (y * g[1] + z * g[2] + y * g[3] - w * g[5] + z * g[6] + w * g[7]);
    gn[2] = 2 * y * (-g[0] + g[4] - g[8]) + 2 * (x * g[1] + w * g[2] + x * g[3] + z * g[5] - w * g[6] + z * g[7]);
    gn[3] = 2 * z * (-g[0] - g[4] + g[8]) + 2 * (-w * g[1] + x * g[2] + w * g[3] + y * g[5] + x * g[6] + y * g[7]);
    float d = qn[0] * gn[0] + qn[1] * gn[1] + qn[2] * gn[2] + qn[3] * gn[3];
#pragma unroll
    for (int i = 0; i < 4; ++i) gq[i] = (gn[i] - qn[i] * d) / m;
}

// batch_rodrigues (mano_layer.py:720-728): epsilon inside the norm, half-angle quaternion
__device__ __forceinline__ void rodrigues(const float* t, float* R) {
    float ex = t[0] + 1e-8f, ey = t[1] + 1e-8f, ez = t[2] + 1e-8f;
    float n = sqrtf(ex * ex + ey * ey + ez * ez);
    float half = n * 0.5f;
    float s, c;
    sincosf(half, &s, &c);
    float q[4] = {c, s * (t[0] / n), s * (t[1] / n), s * (t[2] / n)};
    quat_to_mat(q, R, nullptr, nullptr);
}

__device__ __forceinline__ void rodrigues_bwd(const float* t, const float* gR, float* gt) {
    float ex = t[0] + 1e-8f, ey = t[1] + 1e-8f, ez = t[2] + 1e-8f;
    float n = sqrtf(ex * ex + ey * ey + ez * ez);
    float half = n * 0.5f;
    float s, c;
    sincosf(half, &s, &c);
    float a[3] = {t[0] / n, t[1] / n, t[2] / n};
    float q[4] = {c, s * a[0], s * a[1], s * a[2]};
    float gq[4];
    quat_to_mat_bwd(q, gR, gq);
    float gs = gq[1] * a[0] + gq[2] * a[1] + gq[3] * a[2];
    float ga[3] = {s * gq[1], s * gq[2], s * gq[3]};
    float ghalf = -s * gq[0] + c * gs;
    float gn = 0.5f * ghalf - (ga[0] * t[0] + ga[1] * t[1] + ga[2] * t[2]) / (n * n);
    gt[0] = ga[0] / n + gn * ex / n;
    gt[1] = ga[1] / n + gn * ey / n;
    gt[2] = ga[2] / n + gn * ez / n;
}
