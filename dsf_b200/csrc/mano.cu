// dsf_b200 - MANO hand layer for sm_100a: pose/chain kernel, blend-shape GEMM, skinning kernel and
// their backward.  Replaces render_model/mano_layer.py:573-770 (see include/dsf_b200.h).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "raster.cuh"

// ------------------------------------------------------------------------------------------------
// error / launch bookkeeping
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

void dsf_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void dsf_count_launch(int n) { g_launches += n; }
void dsf_reset_launch_count() { g_launches = 0; }

extern "C" const char* dsf_last_error_string(void) { return g_err; }
extern "C" int dsf_version(void) { return 100; }
extern "C" int dsf_last_launch_count(void) { return g_launches; }

// ------------------------------------------------------------------------------------------------
// M0: constants
// ------------------------------------------------------------------------------------------------
template <typename T>
static int upload(T** dst, const T* src, size_t n) {
    DSF_CHECK_CUDA(cudaMalloc((void**)dst, n * sizeof(T)));
    DSF_CHECK_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return DSF_OK;
}

void dsf_build_collision_mask(float* m);   // coll.cu
static const int h_ring[16] = {121, 214, 215, 279, 239, 234, 92, 38, 122, 118, 117, 119, 120, 108, 79, 78};

extern "C" int dsf_mano_create(const DsfManoHost* host, DsfMano** out) {
    DSF_REQUIRE(host && out, "null host/out");
    DSF_REQUIRE(host->v_template && host->shapedirs && host->posedirs && host->j_regressor &&
                    host->hands_comp && host->hands_mean && host->weights && host->parents && host->faces,
                "null constant array");
    DSF_REQUIRE(host->n_faces > 0 && host->n_faces < 65536, "n_faces out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        dsf_set_error("no CUDA device: dsf_b200 has no CPU fallback");
        return DSF_ERR_NO_DEVICE;
    }
    DsfMano* h = (DsfMano*)calloc(1, sizeof(DsfMano));
    DSF_CHECK_CUDA(cudaGetDevice(&h->device));

    // kinematic tree levels (parents[i] < i, mano_layer.py:753-757 walks i in order)
    h->parents[0] = -1;
    h->level[0] = 0;
    h->maxlevel = 0;
    for (int i = 1; i < NJ; ++i) {
        int p = host->parents[i];
        if (p < 0 || p >= i) {
            dsf_set_error("parents[%d]=%d must be in [0,%d)", i, p, i);
            free(h);
            return DSF_ERR_BAD_ARG;
        }
        h->parents[i] = p;
        h->level[i] = h->level[p] + 1;
        if (h->level[i] > h->maxlevel) h->maxlevel = h->level[i];
    }
    for (int i = 0; i < host->n_faces * 3; ++i) {
        if (host->faces[i] < 0 || host->faces[i] >= NVW) {
            dsf_set_error("face index %d out of range", host->faces[i]);
            free(h);
            return DSF_ERR_BAD_ARG;
        }
    }

    // blend basis D (148 x 2336): rows 0-9 shapedirs, 10-144 posedirs; split into tf32 hi / lo parts
    // once, in the two operand orientations the tensor-core GEMMs read (blend_gemm.cu)
    std::vector<float> D((size_t)KP * NP, 0.f), vt(NP, 0.f);
    for (int k = 0; k < 10; ++k)
        for (int n = 0; n < NV * 3; ++n) D[(size_t)k * NP + n] = host->shapedirs[(size_t)k * NV * 3 + n];
    for (int k = 0; k < 135; ++k)
        for (int n = 0; n < NV * 3; ++n) D[(size_t)(10 + k) * NP + n] = host->posedirs[(size_t)k * NV * 3 + n];
    for (int n = 0; n < NV * 3; ++n) vt[n] = host->v_template[n];
    std::vector<float> BTh((size_t)BLEND_NPAD * BLEND_KPAD, 0.f), BTl(BTh.size(), 0.f);
    std::vector<float> Bh((size_t)BLEND_BN_BWD * NP, 0.f), Bl(Bh.size(), 0.f);
    for (int k = 0; k < KP; ++k)
        for (int n = 0; n < NP; ++n) {
            const float v = D[(size_t)k * NP + n];
            uint32_t bits;
            memcpy(&bits, &v, 4);
            bits &= 0xFFFFE000u;
            float hi;
            memcpy(&hi, &bits, 4);
            const float lo = v - hi;
            BTh[(size_t)n * BLEND_KPAD + k] = hi;
            BTl[(size_t)n * BLEND_KPAD + k] = lo;
            Bh[(size_t)k * NP + n] = hi;
            Bl[(size_t)k * NP + n] = lo;
        }

    // rest joints as an affine function of beta: J = Jt + sum_b beta_b JS[b]   (mano_layer.py:586-591)
    std::vector<float> Jt(NJ * 3), JS(10 * NJ * 3);
    for (int j = 0; j < NJ; ++j)
        for (int c = 0; c < 3; ++c) {
            double a = 0;
            for (int v = 0; v < NV; ++v) a += (double)host->j_regressor[v * NJ + j] * host->v_template[v * 3 + c];
            Jt[j * 3 + c] = (float)a;
            for (int b = 0; b < 10; ++b) {
                double s = 0;
                for (int v = 0; v < NV; ++v)
                    s += (double)host->j_regressor[v * NJ + j] * host->shapedirs[(size_t)b * NV * 3 + v * 3 + c];
                JS[(b * NJ + j) * 3 + c] = (float)s;
            }
        }
    // CSR of the regressor (joint-major), keeps every non-zero
    std::vector<int> ptr(NJ + 1, 0), idx;
    std::vector<float> w;
    for (int j = 0; j < NJ; ++j) {
        for (int v = 0; v < NV; ++v) {
            float x = host->j_regressor[v * NJ + j];
            if (x != 0.f) {
                idx.push_back(v);
                w.push_back(x);
            }
        }
        ptr[j + 1] = (int)idx.size();
    }
    h->jr_nnz = (int)idx.size();
    if (idx.empty()) { idx.push_back(0); w.push_back(0.f); }
    // CSR of the skin weights (joint-major) for the g_A reduction of the backward pass
    std::vector<int> wptr(NJ + 1, 0), widx;
    std::vector<float> wval;
    for (int j = 0; j < NJ; ++j) {
        for (int v = 0; v < NV; ++v) {
            float x = host->weights[v * NJ + j];
            if (x != 0.f) {
                widx.push_back(v);
                wval.push_back(x);
            }
        }
        wptr[j + 1] = (int)widx.size();
    }
    if (widx.empty()) { widx.push_back(0); wval.push_back(0.f); }
    // ... and vertex-major, for the skinning loops
    std::vector<int> vptr(NV + 1, 0);
    std::vector<int2> vent;
    for (int v = 0; v < NV; ++v) {
        for (int j = 0; j < NJ; ++j) {
            const float x = host->weights[v * NJ + j];
            if (x != 0.f) {
                int bits;
                memcpy(&bits, &x, 4);
                vent.push_back(make_int2(j, bits));
            }
        }
        vptr[v + 1] = (int)vent.size();
    }
    if (vent.empty()) vent.push_back(make_int2(0, 0));
    std::vector<float> mask(DSF_NSPHERE * DSF_NSPHERE);
    dsf_build_collision_mask(mask.data());

    int rc = 0;
    rc |= upload(&h->BTh, BTh.data(), BTh.size());
    rc |= upload(&h->BTl, BTl.data(), BTl.size());
    rc |= upload(&h->Bh, Bh.data(), Bh.size());
    rc |= upload(&h->Bl, Bl.data(), Bl.size());
    rc |= upload(&h->vt, vt.data(), vt.size());
    rc |= upload(&h->W, host->weights, (size_t)NV * NJ);
    rc |= upload(&h->comp, host->hands_comp, 45 * 45);
    rc |= upload(&h->mean, host->hands_mean, 45);
    rc |= upload(&h->Jt, Jt.data(), Jt.size());
    rc |= upload(&h->JS, JS.data(), JS.size());
    rc |= upload(&h->jr_ptr, ptr.data(), ptr.size());
    rc |= upload(&h->jr_idx, idx.data(), idx.size());
    rc |= upload(&h->jr_w, w.data(), w.size());
    rc |= upload(&h->wj_ptr, wptr.data(), wptr.size());
    rc |= upload(&h->wj_idx, widx.data(), widx.size());
    rc |= upload(&h->wj_w, wval.data(), wval.size());
    rc |= upload(&h->wv_ptr, vptr.data(), vptr.size());
    rc |= upload(&h->wv_ent, vent.data(), vent.size());
    rc |= upload(&h->faces, host->faces, (size_t)host->n_faces * 3);
    std::vector<unsigned int> fpk((host->n_faces + 3) & ~3, 0u);     // padded to 16 bytes for bulk copies
    for (int f = 0; f < host->n_faces; ++f)
        fpk[f] = (unsigned)host->faces[3 * f] | ((unsigned)host->faces[3 * f + 1] << 10) |
                 ((unsigned)host->faces[3 * f + 2] << 20);
    rc |= upload(&h->faces_packed, fpk.data(), fpk.size());
    {   // processing order of the rasteriser: largest triangles (rest pose) first
        std::vector<std::pair<float, int>> area(host->n_faces);
        auto vtx = [&](int v, int c) {
            if (v < NV) return host->v_template[3 * v + c];
            float a = 0.f;                                   // wrist-cap centre = mean of the ring
            for (int i = 0; i < 16; ++i) a += host->v_template[3 * h_ring[i] + c];
            return a / 16.f;
        };
        for (int f = 0; f < host->n_faces; ++f) {
            float e1[3], e2[3];
            for (int c = 0; c < 3; ++c) {
                e1[c] = vtx(host->faces[3 * f + 1], c) - vtx(host->faces[3 * f], c);
                e2[c] = vtx(host->faces[3 * f + 2], c) - vtx(host->faces[3 * f], c);
            }
            const float cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2],
                        cz = e1[0] * e2[1] - e1[1] * e2[0];
            area[f] = {-(cx * cx + cy * cy + cz * cz), f};
        }
        std::sort(area.begin(), area.end());
        std::vector<unsigned short> order(host->n_faces);
        for (int f = 0; f < host->n_faces; ++f) order[f] = (unsigned short)area[f].second;
        rc |= upload(&h->face_order, order.data(), order.size());
    }
    {   // vertex -> incident face corners (the rasteriser's gradient gather)
        std::vector<int> vptr(NVW + 1, 0);
        for (int i = 0; i < host->n_faces * 3; ++i) vptr[host->faces[i] + 1]++;
        for (int v = 0; v < NVW; ++v) vptr[v + 1] += vptr[v];
        std::vector<unsigned short> vent((size_t)host->n_faces * 3 + 1, 0);
        std::vector<int> fill(vptr.begin(), vptr.end() - 1);
        if (host->n_faces * 4 < 65536)
            for (int f = 0; f < host->n_faces; ++f)
                for (int c = 0; c < 3; ++c) vent[fill[host->faces[3 * f + c]]++] = (unsigned short)(f * 4 + c);
        rc |= upload(&h->vf_ptr, vptr.data(), vptr.size());
        rc |= upload(&h->vf_ent, vent.data(), vent.size());
    }
    rc |= upload(&h->coll_mask, mask.data(), mask.size());
    h->n_faces = host->n_faces;
    if (rc) {
        dsf_mano_free(h);
        return DSF_ERR_CUDA;
    }
    *out = h;
    return DSF_OK;
}

extern "C" int dsf_mano_free(DsfMano* h) {
    if (!h) return DSF_OK;
    void* ptrs[] = {h->BTh, h->BTl, h->Bh, h->Bl, h->vt, h->W, h->comp, h->mean, h->Jt, h->JS,
                    h->jr_ptr, h->jr_idx, h->jr_w, h->wj_ptr, h->wj_idx, h->wj_w, h->faces, h->faces_packed,
                    h->face_order, h->coll_mask, h->vf_ptr, h->vf_ent, h->wv_ptr, h->wv_ent};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    free(h);
    return DSF_OK;
}

extern "C" long dsf_mano_workspace_floats(int batch) { return WS_HANDS(batch) * WS_PER_HAND; }
extern "C" const int* dsf_mano_faces_device(const DsfMano* h, int* n_faces) {
    if (!h) return nullptr;
    if (n_faces) *n_faces = h->n_faces;
    return h->faces;
}

static ChainTopo topo_of(const DsfMano* h) {
    ChainTopo t;
    for (int i = 0; i < NJ; ++i) {
        t.parents[i] = h->parents[i];
        t.level[i] = h->level[i];
        t.nchild[i] = 0;
    }
    for (int i = 1; i < NJ; ++i)
        if (h->parents[i] >= 0 && h->parents[i] < NJ) ++t.nchild[h->parents[i]];
    t.maxlevel = h->maxlevel;
    return t;
}

// ------------------------------------------------------------------------------------------------
// K1: pose kernel - 16 lanes per hand, lane j owns joint j.
//   PCA pose -> axis angles (mano_layer.py:601), Rodrigues (:720-728), pose feature (:611),
//   rest joints from beta (:586-591), kinematic chain by tree level (:730-770).
// ------------------------------------------------------------------------------------------------
// stage the PCA basis (45 x 45) and the rest-joint regressor (10 x NJ x 3) in shared memory with cp.async: every
// copy is in flight at once and no register waits for a load (a load -> store loop per thread serialises on the
// global latency).  Both tables are cudaMalloc'ed (16-byte aligned); 4-byte copies, 128 threads: 20 per thread.
__device__ __forceinline__ void stage_pose_tables(float* s_comp, float* s_JS, const float* comp, const float* JS, int tid,
                                                  int nthr) {
#ifdef POSE_STAGE_LDG
    for (int i = tid; i < 45 * 45; i += nthr) s_comp[i] = __ldg(comp + i);
    for (int i = tid; i < 10 * NJ * 3; i += nthr) s_JS[i] = __ldg(JS + i);
#else
    // 16-byte pieces (both tables are cudaMalloc'ed, the shared arrays 16-byte aligned): 506 + 120 copies, then the
    // last float of the 2025-entry basis on its own
    for (int i = tid; i < (45 * 45) / 4; i += nthr)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(s_comp + 4 * i)), "l"(comp + 4 * i) : "memory");
    for (int i = tid; i < (10 * NJ * 3) / 4; i += nthr)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(s_JS + 4 * i)), "l"(JS + 4 * i) : "memory");
    if (tid == 0)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(s_comp + 45 * 45 - 1)), "l"(comp + 45 * 45 - 1) : "memory");
    static_assert((10 * NJ * 3) % 4 == 0 && (45 * 45) % 4 == 1, "table sizes");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

#define POSE_HPB 8   // hands per block (128 threads)

__global__ void __launch_bounds__(POSE_HPB* NJ)
mano_pose_kernel(int B, DsfManoParams p, const float* __restrict__ comp, const float* __restrict__ mean,
                 const float* __restrict__ Jt, const float* __restrict__ JS, ChainTopo topo,
                 float* __restrict__ ws) {
    __shared__ float s_theta[POSE_HPB][48];
    __shared__ float s_beta[POSE_HPB][12];
    __shared__ float s_G[POSE_HPB][NJ][15];   // Gr[9] Gt[3] J[3]
    // the PCA basis and the rest-joint regressors are read 135 + 30 times per lane: stage them once per block
    // with coalesced loads instead of walking an 8 KB table through L1 one 180-byte row per iteration
    __shared__ __align__(16) float s_comp[45 * 45 + 3];
    __shared__ __align__(16) float s_JS[10 * NJ * 3];
    stage_pose_tables(s_comp, s_JS, comp, JS, threadIdx.x, POSE_HPB * NJ);
    __syncthreads();
    const int hl = threadIdx.x / NJ;
    const int j = threadIdx.x % NJ;
    const int hand = blockIdx.x * POSE_HPB + hl;
    const bool live = hand < B;
    const int hh = live ? hand : B - 1;

    for (int k = j; k < 45; k += NJ) s_theta[hl][k] = (k < p.ncomp) ? p.theta[(size_t)hh * p.ld_theta + k] : 0.f;
    if (j < 10) s_beta[hl][j] = p.beta[(size_t)hh * p.ld_beta + j];
    __syncwarp();

    float ang[3] = {0.f, 0.f, 0.f};
    float R[9];
    if (j == 0) {
        if (p.quat_dim == 3) {
            ang[0] = p.quat[(size_t)hh * p.ld_quat];
            ang[1] = p.quat[(size_t)hh * p.ld_quat + 1];
            ang[2] = p.quat[(size_t)hh * p.ld_quat + 2];
            rodrigues(ang, R);
        } else {
            float q[4];
            for (int i = 0; i < 4; ++i) q[i] = p.quat[(size_t)hh * p.ld_quat + i];
            quat_to_mat(q, R, nullptr, nullptr);
        }
    } else {
        const int a0 = 3 * (j - 1);
        ang[0] = mean[a0]; ang[1] = mean[a0 + 1]; ang[2] = mean[a0 + 2];
        for (int k = 0; k < p.ncomp; ++k) {
            float t = s_theta[hl][k];
            ang[0] = fmaf(t, s_comp[k * 45 + a0], ang[0]);
            ang[1] = fmaf(t, s_comp[k * 45 + a0 + 1], ang[1]);
            ang[2] = fmaf(t, s_comp[k * 45 + a0 + 2], ang[2]);
        }
        rodrigues(ang, R);
    }
    float J[3] = {Jt[j * 3], Jt[j * 3 + 1], Jt[j * 3 + 2]};
#pragma unroll
    for (int b = 0; b < 10; ++b) {
        float be = s_beta[hl][b];
        J[0] = fmaf(be, s_JS[(b * NJ + j) * 3], J[0]);
        J[1] = fmaf(be, s_JS[(b * NJ + j) * 3 + 1], J[1]);
        J[2] = fmaf(be, s_JS[(b * NJ + j) * 3 + 2], J[2]);
    }
    // The kernel's outputs - the GEMM operand row and the 16 joint records of every hand - are collected in shared
    // memory and leave as whole rows (16-byte stores, consecutive lanes on consecutive addresses) at the end: written
    // from here, every store instruction would touch 32 different 128-byte lines (lane = joint, records 128 bytes
    // apart), seven times the memory transactions the data needs.
    __shared__ __align__(16) float s_X[POSE_HPB][2 * BLEND_KPAD];
    __shared__ __align__(16) float s_out[POSE_HPB][NJ * RJ_STRIDE];
    {
        // GEMM operand row X = [beta | Rs - I | 0], already split for the 3xTF32 tensor-core GEMM:
        // hi = the value with the low 13 mantissa bits cleared (exact in tf32), lo = the remainder
        float* Xh = s_X[hl];
        float* Xl = Xh + BLEND_KPAD;
        auto put = [&](int k, float v) {
            const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
            Xh[k] = hi;
            Xl[k] = v - hi;
        };
        if (j < 10) put(j, s_beta[hl][j]);
        if (j >= 1) {
#pragma unroll
            for (int e = 0; e < 9; ++e) put(10 + 9 * (j - 1) + e, R[e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f));
        }
        if (j < BLEND_KPAD - 145) { Xh[145 + j] = 0.f; Xl[145 + j] = 0.f; }
    }

    float Gr[9], Gt[3];
    if (j == 0) {
#pragma unroll
        for (int e = 0; e < 9; ++e) Gr[e] = R[e];
        Gt[0] = J[0]; Gt[1] = J[1]; Gt[2] = J[2];
    }
    s_G[hl][j][12] = J[0]; s_G[hl][j][13] = J[1]; s_G[hl][j][14] = J[2];
    if (j == 0) {
#pragma unroll
        for (int e = 0; e < 9; ++e) s_G[hl][0][e] = Gr[e];
        s_G[hl][0][9] = Gt[0]; s_G[hl][0][10] = Gt[1]; s_G[hl][0][11] = Gt[2];
    }
    __syncwarp();
    for (int lvl = 1; lvl <= topo.maxlevel; ++lvl) {
        if (topo.level[j] == lvl) {
            const float* P = s_G[hl][topo.parents[j]];
            float d[3] = {J[0] - P[12], J[1] - P[13], J[2] - P[14]};
            mat3_mul(P, R, Gr);
            mat3_vec(P, d, Gt);
            Gt[0] += P[9]; Gt[1] += P[10]; Gt[2] += P[11];
#pragma unroll
            for (int e = 0; e < 9; ++e) s_G[hl][j][e] = Gr[e];
            s_G[hl][j][9] = Gt[0]; s_G[hl][j][10] = Gt[1]; s_G[hl][j][11] = Gt[2];
        }
        __syncwarp();
    }
    {
        float* o = &s_out[hl][j * RJ_STRIDE];
        float GJ[3];
        mat3_vec(Gr, J, GJ);
#pragma unroll
        for (int e = 0; e < 9; ++e) { o[RJ_R + e] = R[e]; o[RJ_GR + e] = Gr[e]; }
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            o[RJ_GT + e] = Gt[e];
            o[RJ_J + e] = J[e];
            o[RJ_AT + e] = Gt[e] - GJ[e];   // A = G - [0 | G.J]  (mano_layer.py:765-768)
            o[RJ_ANG + e] = ang[e];
        }
        o[30] = 0.f; o[31] = 0.f;           // pad of the record
    }
    __syncthreads();
    // rows out: per hand 2 * BLEND_KPAD / 4 + NJ * RJ_STRIDE / 4 = 208 float4 (workspace rows are 16-byte aligned)
    constexpr int XQ = 2 * BLEND_KPAD / 4, RQ = NJ * RJ_STRIDE / 4;
    static_assert(RQ == POSE_HPB * NJ && XQ <= POSE_HPB * NJ, "one 16-byte piece of a hand's rows per thread");
#pragma unroll
    for (int h_ = 0; h_ < POSE_HPB; ++h_) {
        const int hd = blockIdx.x * POSE_HPB + h_;
        if (hd >= B) break;
        float* row = ws + (size_t)hd * WS_PER_HAND;
        reinterpret_cast<float4*>(row + WS_RJ)[threadIdx.x] = reinterpret_cast<const float4*>(s_out[h_])[threadIdx.x];
        if (threadIdx.x < XQ) reinterpret_cast<float4*>(row + WS_X)[threadIdx.x] = reinterpret_cast<const float4*>(s_X[h_])[threadIdx.x];
    }
}

// K2: the blend-shape contraction runs on the tensor cores, see blend_gemm.cu
int dsf_blend_forward_gemm(int M, const float* Ah, const float* Al, int lda, const float* Bh, const float* Bl, float* C,
                           int ldc, const float* bias, cudaStream_t st);
int dsf_blend_backward_gemm(int M, const float* Ah, const float* Al, int lda, const float* Bh, const float* Bl, float* C,
                            int ldc, long split_stride, cudaStream_t st);
int dsf_blend_backward_splits(int M);

// ------------------------------------------------------------------------------------------------
// K3: skinning kernel - one CTA per hand.  LBS (:619-629), joint regression (:630-633), wrist-cap
// vertex (:636-637), unit / camera scaling (:662-675).
// ------------------------------------------------------------------------------------------------
#define SKIN_T 256
static const int h_tips[5] = {333, 444, 672, 555, 744};
__constant__ int c_ring[16];
__constant__ int c_tips[5];
static bool g_const_ready[16] = {};

static int ensure_constants() {
    int dev = 0;
    DSF_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 16 && g_const_ready[dev]) return DSF_OK;
    DSF_CHECK_CUDA(cudaMemcpyToSymbol(c_ring, h_ring, sizeof(h_ring)));
    DSF_CHECK_CUDA(cudaMemcpyToSymbol(c_tips, h_tips, sizeof(h_tips)));
    if (dev < 16) g_const_ready[dev] = true;
    return DSF_OK;
}

__global__ void __launch_bounds__(SKIN_T)
mano_skin_kernel(int B, const float* __restrict__ ws, const int* __restrict__ wv_ptr, const int2* __restrict__ wv_ent,
                 const int* __restrict__ jr_ptr, const int* __restrict__ jr_idx,
                 const float* __restrict__ jr_w, const float* __restrict__ cam, int ld_cam, float unit_scale,
                 float* __restrict__ verts, float* __restrict__ joints, float* __restrict__ Rs) {
    __shared__ float sA[NJ][12];
    __shared__ float sV[NVW * 3];
    __shared__ __align__(16) float sVP[NP];                  // v_posed row, staged by one TMA bulk copy
    __shared__ __align__(8) unsigned long long s_bar;
    const int hand = blockIdx.x;
    const int tid = threadIdx.x;
    const float* wsh = ws + (size_t)hand * WS_PER_HAND;
    const bool bulk = (reinterpret_cast<uintptr_t>(wsh + WS_VP) & 15) == 0;
    if (bulk && tid == 0) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(NP * (uint32_t)sizeof(float)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(sVP)),
                     "l"(wsh + WS_VP), "r"(NP * (uint32_t)sizeof(float)), "r"(bar)
                     : "memory");
    }
    for (int i = tid; i < NJ * 12; i += SKIN_T) {
        int j = i / 12, e = i % 12;
        sA[j][e] = (e < 9) ? wsh[WS_RJ + j * RJ_STRIDE + RJ_GR + e] : wsh[WS_RJ + j * RJ_STRIDE + RJ_AT + e - 9];
    }
    if (Rs) {
        for (int i = tid; i < 15 * 9; i += SKIN_T)
            Rs[(size_t)hand * 135 + i] = wsh[WS_RJ + (1 + i / 9) * RJ_STRIDE + RJ_R + i % 9];
    }
    float s = unit_scale, tx = 0.f, ty = 0.f, tz = 0.f;
    if (cam) {
        s *= cam[(size_t)hand * ld_cam];
        tx = cam[(size_t)hand * ld_cam + 1];
        ty = cam[(size_t)hand * ld_cam + 2];
        tz = cam[(size_t)hand * ld_cam + 3];
    }
    __syncthreads();                          // sA complete; the mbarrier is initialised before anyone polls it
    const float* VP = wsh + WS_VP;
    if (bulk) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
        asm volatile(
            "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n" ::"r"(bar)
            : "memory");
        VP = sVP;
    }
    float* vo = verts + (size_t)hand * NVW * 3;
    for (int v = tid; v < NV; v += SKIN_T) {
        float x = VP[3 * v], y = VP[3 * v + 1], z = VP[3 * v + 2];
        float T[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) T[e] = 0.f;
        const int e1 = __ldg(wv_ptr + v + 1);
        for (int k = __ldg(wv_ptr + v); k < e1; ++k) {           // the joints this vertex follows, ascending
            const int2 en = __ldg(wv_ent + k);
            const float wgt = __int_as_float(en.y);
            const float* Aj = sA[en.x];
#pragma unroll
            for (int e = 0; e < 12; ++e) T[e] = fmaf(wgt, Aj[e], T[e]);
        }
        float ox = T[0] * x + T[1] * y + T[2] * z + T[9];
        float oy = T[3] * x + T[4] * y + T[5] * z + T[10];
        float oz = T[6] * x + T[7] * y + T[8] * z + T[11];
        sV[3 * v] = ox; sV[3 * v + 1] = oy; sV[3 * v + 2] = oz;
        vo[3 * v] = ox * s + tx; vo[3 * v + 1] = oy * s + ty; vo[3 * v + 2] = oz * s + tz;
    }
    __syncthreads();
    const int warp = tid / 32, lane = tid % 32;
    if (warp == 0) {   // wrist-cap centre = mean of the 16 ring vertices
        float a = 0.f, b = 0.f, c = 0.f;
        if (lane < 16) { int r = c_ring[lane]; a = sV[3 * r]; b = sV[3 * r + 1]; c = sV[3 * r + 2]; }
        a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
        if (lane == 0) {
            a *= (1.f / 16.f); b *= (1.f / 16.f); c *= (1.f / 16.f);
            vo[3 * NV] = a * s + tx; vo[3 * NV + 1] = b * s + ty; vo[3 * NV + 2] = c * s + tz;
        }
    }
    float* jo = joints + (size_t)hand * NJOUT * 3;
    for (int j = warp; j < NJ; j += SKIN_T / 32) {
        float a = 0.f, b = 0.f, c = 0.f;
        for (int e = jr_ptr[j] + lane; e < jr_ptr[j + 1]; e += 32) {
            int v = jr_idx[e];
            float w = jr_w[e];
            a = fmaf(w, sV[3 * v], a); b = fmaf(w, sV[3 * v + 1], b); c = fmaf(w, sV[3 * v + 2], c);
        }
        a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
        if (lane == 0) { jo[3 * j] = a * s + tx; jo[3 * j + 1] = b * s + ty; jo[3 * j + 2] = c * s + tz; }
    }
    if (tid < 5) {
        int v = c_tips[tid];
        jo[3 * (NJ + tid)] = sV[3 * v] * s + tx;
        jo[3 * (NJ + tid) + 1] = sV[3 * v + 1] * s + ty;
        jo[3 * (NJ + tid) + 2] = sV[3 * v + 2] * s + tz;
    }
}

// ------------------------------------------------------------------------------------------------
// Kb2: skinning backward - one CTA per hand.
//   cotangents of verts/joints -> g_vposed (ws), g_A (ws), g_cam.
// ------------------------------------------------------------------------------------------------
#define SKB_T 256
#define SKB_TILE_BYTES ((NVW * 3 * 4 + 15 + 15) & ~15)      // a 9348-byte row plus its alignment window

__global__ void __launch_bounds__(SKB_T)
mano_skin_bwd_kernel(int B, float* __restrict__ ws, const int* __restrict__ wv_ptr, const int2* __restrict__ wv_ent,
                     const int* __restrict__ jr_ptr, const int* __restrict__ jr_idx,
                     const float* __restrict__ jr_w, const int* __restrict__ wj_ptr,
                     const int* __restrict__ wj_idx, const float* __restrict__ wj_w,
                     const float* __restrict__ cam, int ld_cam,
                     float unit_scale, const float* __restrict__ verts, const float* __restrict__ joints,
                     const float* __restrict__ g_verts, const float* __restrict__ g_joints,
                     float* __restrict__ g_cam, int ld_gcam, GradTiles gt, const float* __restrict__ cube,
                     LossFold lf) {
    __shared__ float sg[NVW * 3];
    __shared__ __align__(16) float svp[NP];                  // v_posed of the hand (padded row of the workspace)
    __shared__ __align__(16) unsigned char s_tiles[2][SKB_TILE_BYTES];   // the rasteriser's two gradient shares
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ float sGr[NJ][9];
    __shared__ float sgj[NJOUT * 3];
    __shared__ float red[SKB_T / 32][12];
    const int hand = blockIdx.x, tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    float* wsh = ws + (size_t)hand * WS_PER_HAND;
    // Bulk staging (fused step at crop sizes up to 128, i.e. at most two raster tiles per hand): one thread hands the
    // hand's v_posed row and both per-tile gradient shares (28 KB) to the TMA engine before anything else happens, so
    // the copies fly while the CTA resolves the hand-level scalars - instead of every thread walking three global
    // arrays with stride-12-byte loads whose latency nothing covers.  The shares' rows are 9348 bytes apart (not a
    // multiple of 16): each copy takes the 16-byte aligned window around its row, the data starts `shift` bytes in.
    const bool bulk = gt.gv_tile != nullptr && gt.n_tiles <= 2 && ((reinterpret_cast<uintptr_t>(wsh + WS_VP) & 15) == 0);
    uint32_t t_shift[2] = {0u, 0u};
    if (bulk) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
        uint32_t t_bytes[2] = {0u, 0u};
        const char* t_src[2] = {nullptr, nullptr};
        uint32_t total = NP * (uint32_t)sizeof(float);
#pragma unroll
        for (int t = 0; t < 2; ++t)
            if (t < gt.n_tiles) {
                const uintptr_t a = reinterpret_cast<uintptr_t>(gt.gv_tile + ((size_t)hand * gt.n_tiles + t) * NVW * 3);
                t_shift[t] = (uint32_t)(a & 15);
                t_bytes[t] = (t_shift[t] + NVW * 3 * (uint32_t)sizeof(float) + 15u) & ~15u;
                t_src[t] = reinterpret_cast<const char*>(a - t_shift[t]);
                total += t_bytes[t];
            }
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(svp)),
                         "l"(wsh + WS_VP), "r"(NP * (uint32_t)sizeof(float)), "r"(bar)
                         : "memory");
#pragma unroll
            for (int t = 0; t < 2; ++t)
                if (t < gt.n_tiles)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     (uint32_t)__cvta_generic_to_shared(s_tiles[t])),
                                 "l"(t_src[t]), "r"(t_bytes[t]), "r"(bar)
                                 : "memory");
        }
    }
    float cs = 1.f, tx = 0.f, ty = 0.f, tz = 0.f;
    if (cam) {
        cs = cam[(size_t)hand * ld_cam];
        tx = cam[(size_t)hand * ld_cam + 1];
        ty = cam[(size_t)hand * ld_cam + 2];
        tz = cam[(size_t)hand * ld_cam + 3];
    }
    const float s_tot = unit_scale * cs;
    for (int i = tid; i < NJ * 9; i += SKB_T) sGr[i / 9][i % 9] = wsh[WS_RJ + (i / 9) * RJ_STRIDE + RJ_GR + i % 9];
    for (int i = tid; i < NJOUT * 3; i += SKB_T) sgj[i] = g_joints ? g_joints[(size_t)hand * NJOUT * 3 + i] : 0.f;
    // camera gradient partials: out = mano * s_tot + t
    float pt[4] = {0.f, 0.f, 0.f, 0.f};
    const float* gv = g_verts ? g_verts + (size_t)hand * NVW * 3 : nullptr;
    const float* vo = verts + (size_t)hand * NVW * 3;
    // vertex cotangent straight from the rasteriser's per-tile shares (fused step): sum the tiles, apply the
    // hand's loss normalisation gk / zhalf.  The hand-level values (normalisation, which tiles carry a gradient,
    // the loss record) are worked out once, by thread 0, not by every thread.
    __shared__ float s_gts;
    __shared__ int s_live[9];                 // [0] = number of live tiles (-1: more than 8 tiles, general path), then ids
    if (warp == 0) {
        // lane t fetches tile t's records, all at once (a scalar loop would pay one global round trip per record);
        // the sums then run over the lanes in ascending tile order, as grad_tiles_scale / the fold kernels do
        const bool has_t = (gt.gv_tile != nullptr || lf.parts != nullptr);
        const int nt = gt.gv_tile ? gt.n_tiles : lf.n_tiles;
        const float* pt_base = gt.gv_tile ? gt.parts_tile : lf.parts_tile;
        float my_sum = 0.f, my_cnt = 0.f;
        int my_flag = 0;
        for (int t0 = 0; t0 < (has_t ? nt : 0); t0 += 32) {          // one trip unless there are more than 32 tiles
            const int t = t0 + lane;
            float a = 0.f, c = 0.f;
            int f = 0;
            if (t < nt) {
                a = pt_base[((size_t)hand * nt + t) * 2];
                c = pt_base[((size_t)hand * nt + t) * 2 + 1];
                if (gt.gv_tile) f = gt.gv_flag[(size_t)hand * nt + t];
            }
            if (t0 == 0) { my_sum = a; my_cnt = c; my_flag = f; }
        }
        const float zh = (gt.gv_tile && lane == 0) ? cube[3 * hand + 2] * 0.5f : 0.f;
        if (nt <= 32 && has_t) {
            float a = 0.f, c = 0.f;
            int n_live = 0;
            for (int t = 0; t < nt; ++t) {
                a += __shfl_sync(0xffffffffu, my_sum, t);
                c += __shfl_sync(0xffffffffu, my_cnt, t);
                const int f = __shfl_sync(0xffffffffu, my_flag, t);
                if (lane == 0 && f && nt <= 8) s_live[1 + n_live] = t;
                n_live += f ? 1 : 0;
            }
            if (lane == 0) {
                s_gts = gt.gv_tile ? gt.gscale / (c + 1e-8f) / zh : 0.f;
                s_live[0] = (gt.gv_tile && nt <= 8) ? n_live : -1;
                if (lf.parts && lf.n_mesh == B) { lf.parts[2 * hand] = a; lf.parts[2 * hand + 1] = c; }
            }
        } else if (lane == 0) {                                       // general path
            s_gts = gt.gv_tile ? grad_tiles_scale(gt, hand, cube[3 * hand + 2] * 0.5f) : 0.f;
            s_live[0] = -1;
            if (lf.parts && lf.n_mesh == B) {
                float a = 0.f, c = 0.f;
                for (int t = 0; t < lf.n_tiles; ++t) {
                    a += lf.parts_tile[((size_t)hand * lf.n_tiles + t) * 2];
                    c += lf.parts_tile[((size_t)hand * lf.n_tiles + t) * 2 + 1];
                }
                lf.parts[2 * hand] = a; lf.parts[2 * hand + 1] = c;
            }
        }
    }
    __syncthreads();                          // also: the mbarrier is initialised before anyone polls it
    const float gts = s_gts;
    const int n_live_tiles = s_live[0];
    const float* tile0 = nullptr;
    const float* tile1 = nullptr;
    if (bulk) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
        asm volatile(
            "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n" ::"r"(bar)
            : "memory");
        // (select, not index: keeps t_shift in registers)
        if (n_live_tiles >= 1) tile0 = s_live[1] == 0 ? reinterpret_cast<const float*>(s_tiles[0] + t_shift[0])
                                                      : reinterpret_cast<const float*>(s_tiles[1] + t_shift[1]);
        if (n_live_tiles >= 2) tile1 = reinterpret_cast<const float*>(s_tiles[1] + t_shift[1]);
    } else {
        if (n_live_tiles >= 1) tile0 = gt.gv_tile + ((size_t)hand * gt.n_tiles + s_live[1]) * NVW * 3;
        if (n_live_tiles >= 2) tile1 = gt.gv_tile + ((size_t)hand * gt.n_tiles + s_live[2]) * NVW * 3;
    }
    for (int v = tid; v < NVW; v += SKB_T) {
        float a = gv ? gv[3 * v] : 0.f, b = gv ? gv[3 * v + 1] : 0.f, c = gv ? gv[3 * v + 2] : 0.f;
        if (gt.gv_tile) {
            if (n_live_tiles >= 0) {
                a = b = c = 0.f;
                // ascending tile order (fixed summation order); one or two live tiles is the 128 x 128 crop
                if (tile0) { a += tile0[3 * v]; b += tile0[3 * v + 1]; c += tile0[3 * v + 2]; }
                if (tile1) { a += tile1[3 * v]; b += tile1[3 * v + 1]; c += tile1[3 * v + 2]; }
                for (int t = 2; t < n_live_tiles; ++t) {
                    const float* tp = gt.gv_tile + ((size_t)hand * gt.n_tiles + s_live[1 + t]) * NVW * 3;
                    a += tp[3 * v]; b += tp[3 * v + 1]; c += tp[3 * v + 2];
                }
                a *= gts; b *= gts; c *= gts;
            } else {
                a = gts * grad_tiles_load(gt, hand, 3 * v);
                b = gts * grad_tiles_load(gt, hand, 3 * v + 1);
                c = gts * grad_tiles_load(gt, hand, 3 * v + 2);
            }
        }
        sg[3 * v] = a; sg[3 * v + 1] = b; sg[3 * v + 2] = c;
        pt[1] += a; pt[2] += b; pt[3] += c;
        pt[0] += a * (vo[3 * v] - tx) + b * (vo[3 * v + 1] - ty) + c * (vo[3 * v + 2] - tz);
    }
    __syncthreads();
    if (tid < NJOUT) {
        const float* jo = joints + (size_t)hand * NJOUT * 3 + 3 * tid;
        float a = sgj[3 * tid], b = sgj[3 * tid + 1], c = sgj[3 * tid + 2];
        pt[1] += a; pt[2] += b; pt[3] += c;
        pt[0] += a * (jo[0] - tx) + b * (jo[1] - ty) + c * (jo[2] - tz);
    }
    // wrist-cap vertex spreads onto the ring (:636)
    if (tid < 16) {
        int r = c_ring[tid];
        sg[3 * r] += sg[3 * NV] * (1.f / 16.f);
        sg[3 * r + 1] += sg[3 * NV + 1] * (1.f / 16.f);
        sg[3 * r + 2] += sg[3 * NV + 2] * (1.f / 16.f);
    }
    __syncthreads();
    if (tid < 5) {
        int v = c_tips[tid];
        sg[3 * v] += sgj[3 * (NJ + tid)];
        sg[3 * v + 1] += sgj[3 * (NJ + tid) + 1];
        sg[3 * v + 2] += sgj[3 * (NJ + tid) + 2];
    }
    __syncthreads();
    if (g_joints) {
        for (int j = warp; j < NJ; j += SKB_T / 32) {
            float a = sgj[3 * j], b = sgj[3 * j + 1], c = sgj[3 * j + 2];
            for (int e = jr_ptr[j] + lane; e < jr_ptr[j + 1]; e += 32) {
                int v = jr_idx[e];
                float w = jr_w[e];
                atomicAdd(&sg[3 * v], w * a);
                atomicAdd(&sg[3 * v + 1], w * b);
                atomicAdd(&sg[3 * v + 2], w * c);
            }
        }
    }
    __syncthreads();
    // per-vertex: g_vposed = T^T g ; sg becomes the gradient wrt the unscaled skinned vertex
    const float* VP = wsh + WS_VP;
    float* GVP = wsh + WS_GVP;
    for (int v = tid; v < NV; v += SKB_T) {
        const float g0 = sg[3 * v] * s_tot, g1 = sg[3 * v + 1] * s_tot, g2 = sg[3 * v + 2] * s_tot;
        sg[3 * v] = g0; sg[3 * v + 1] = g1; sg[3 * v + 2] = g2;
        if (!bulk) { svp[3 * v] = VP[3 * v]; svp[3 * v + 1] = VP[3 * v + 1]; svp[3 * v + 2] = VP[3 * v + 2]; }
        float T[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) T[e] = 0.f;
        const int e1 = __ldg(wv_ptr + v + 1);
        for (int k = __ldg(wv_ptr + v); k < e1; ++k) {
            const int2 en = __ldg(wv_ent + k);
            const float wgt = __int_as_float(en.y);
#pragma unroll
            for (int e = 0; e < 9; ++e) T[e] = fmaf(wgt, sGr[en.x][e], T[e]);
        }
        // written already split for the 3xTF32 tensor-core GEMM (hi = low 13 mantissa bits cleared, lo = the rest):
        // the backward contraction then streams both operands asynchronously instead of splitting in registers
        const float o[3] = {T[0] * g0 + T[3] * g1 + T[6] * g2, T[1] * g0 + T[4] * g1 + T[7] * g2,
                            T[2] * g0 + T[5] * g1 + T[8] * g2};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float hi = __uint_as_float(__float_as_uint(o[c]) & 0xFFFFE000u);
            GVP[3 * v + c] = hi;
            GVP[NP + 3 * v + c] = o[c] - hi;
        }
    }
    if (tid < NP - NV * 3) { GVP[NV * 3 + tid] = 0.f; GVP[NP + NV * 3 + tid] = 0.f; }
    __syncthreads();
    // g_A[j] = sum_v w_vj [ g (x) vp | g ] over the vertices joint j actually moves: one warp per
    // joint walks that joint's weight list (CSR) and reduces with shuffles
    float* GA = wsh + WS_GA;
    for (int j = warp; j < NJ; j += SKB_T / 32) {
        float acc[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = 0.f;
        for (int e = wj_ptr[j] + lane; e < wj_ptr[j + 1]; e += 32) {
            const int v = wj_idx[e];
            const float w = wj_w[e];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float wg = w * sg[3 * v + r];
                acc[3 * r] = fmaf(wg, svp[3 * v], acc[3 * r]);
                acc[3 * r + 1] = fmaf(wg, svp[3 * v + 1], acc[3 * r + 1]);
                acc[3 * r + 2] = fmaf(wg, svp[3 * v + 2], acc[3 * r + 2]);
                acc[9 + r] += wg;
            }
        }
        // 12 warp sums as one scatter-reduction: total e lands on lanes 2e, 2e + 1
        const float tot = warp_sum16_scatter(acc, lane);
        if (!(lane & 1) && (lane >> 1) < 12) GA[j * 12 + (lane >> 1)] = tot;
    }
    // camera gradient
    if (g_cam) {
#pragma unroll
        for (int e = 0; e < 4; ++e) pt[e] = warp_sum(pt[e]);
        if (lane == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) red[warp][e] = pt[e];
        }
        __syncthreads();
        if (tid < 4) {
            float a = 0.f;
#pragma unroll
            for (int w = 0; w < SKB_T / 32; ++w) a += red[w][tid];
            // d out / d cam.scale = unit_scale * mano = (out - t) / cam.scale
            if (tid == 0) a = a / cs;
            g_cam[(size_t)hand * ld_gcam + tid] = a;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Kb4: pose backward - 16 lanes per hand.  g_A, g_X -> g_quat, g_theta, g_beta.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(POSE_HPB* NJ)
mano_pose_bwd_kernel(int B, DsfManoParams p, DsfManoGrads g, const float* __restrict__ comp,
                     const float* __restrict__ JS, ChainTopo topo, const float* __restrict__ ws, int n_split,
                     LossFold lf) {
    __shared__ float s_acc[POSE_HPB][NJ][15];   // gGr[9] gGt[3] gJ[3]
    __shared__ float s_gang[POSE_HPB][48];
    __shared__ __align__(16) float s_comp[45 * 45 + 3];   // staged once per block, see mano_pose_kernel
    __shared__ __align__(16) float s_JS[10 * NJ * 3];
    // every per-hand input of the block (the joints' forward records and the skinning cotangent g_A: 2.75 KB per
    // hand) is staged with 16-byte cp.async copies issued together with the tables: nothing in the kernel then waits
    // on a dependent global load (the tree walk used to fetch the parent's record from the workspace at every level)
    __shared__ __align__(16) float s_rj[POSE_HPB][NJ * RJ_STRIDE];
    __shared__ __align__(16) float s_ga[POSE_HPB][NJ * 12];
    __shared__ float s_gx[POSE_HPB][KP + 4];    // split-K partials of g_X summed per hand
#pragma unroll
    for (int h_ = 0; h_ < POSE_HPB; ++h_) {
        const int hs = min(blockIdx.x * POSE_HPB + h_, B - 1);
        const float* row = ws + (size_t)hs * WS_PER_HAND;
        // the records: NJ * RJ_STRIDE / 4 = 128 pieces = one per thread; g_A: 48 pieces
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_rj[h_][4 * threadIdx.x])),
                     "l"(row + WS_RJ + 4 * threadIdx.x) : "memory");
        if (threadIdx.x < NJ * 12 / 4)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_ga[h_][4 * threadIdx.x])),
                         "l"(row + WS_GA + 4 * threadIdx.x) : "memory");
    }
    static_assert(NJ * RJ_STRIDE / 4 == POSE_HPB * NJ, "one 16-byte piece of a hand's records per thread");
    stage_pose_tables(s_comp, s_JS, comp, JS, threadIdx.x, POSE_HPB * NJ);     // commits and waits for all copies
    const int hl = threadIdx.x / NJ;
    const int j = threadIdx.x % NJ;
    const int hand = blockIdx.x * POSE_HPB + hl;
    const bool live = hand < B;
    const int hh = live ? hand : B - 1;
    const float* wsh = ws + (size_t)hh * WS_PER_HAND;
    {   // g_X = sum of the split-K partials, fixed order; the hand's 16 lanes walk the 148 columns together
        float a[(KP + NJ - 1) / NJ];
#pragma unroll
        for (int t = 0; t < (KP + NJ - 1) / NJ; ++t) a[t] = 0.f;
        for (int z = 0; z < n_split; ++z)
#pragma unroll
            for (int t = 0; t < (KP + NJ - 1) / NJ; ++t) {
                const int c = j + t * NJ;
                if (c < KP) a[t] += wsh[WS_GX + z * KP + c];
            }
#pragma unroll
        for (int t = 0; t < (KP + NJ - 1) / NJ; ++t)
            if (j + t * NJ < KP) s_gx[hl][j + t * NJ] = a[t];
    }
    __syncthreads();
    const float* rj = &s_rj[hl][j * RJ_STRIDE];
    float R[9], Gr[9], J[3], ang[3];
#pragma unroll
    for (int e = 0; e < 9; ++e) { R[e] = rj[RJ_R + e]; Gr[e] = rj[RJ_GR + e]; }
#pragma unroll
    for (int e = 0; e < 3; ++e) { J[e] = rj[RJ_J + e]; ang[e] = rj[RJ_ANG + e]; }
    const float* ga = &s_ga[hl][j * 12];
    float gAt[3] = {ga[9], ga[10], ga[11]};
    // A_r = Gr, A_t = Gt - Gr J
    float tmp[3];
    mat3T_vec(Gr, gAt, tmp);
    float* acc = s_acc[hl][j];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[3 * r + c] = ga[3 * r + c] - gAt[r] * J[c];
    acc[9] = gAt[0]; acc[10] = gAt[1]; acc[11] = gAt[2];
    acc[12] = -tmp[0]; acc[13] = -tmp[1]; acc[14] = -tmp[2];
    __syncwarp();
    float gR[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) gR[e] = 0.f;
    for (int lvl = topo.maxlevel; lvl >= 1; --lvl) {
        if (topo.level[j] == lvl) {
            const int pj = topo.parents[j];
            const float* prj = &s_rj[hl][pj * RJ_STRIDE];
            float Pg[9], Pj[3];
#pragma unroll
            for (int e = 0; e < 9; ++e) Pg[e] = prj[RJ_GR + e];
#pragma unroll
            for (int e = 0; e < 3; ++e) Pj[e] = prj[RJ_J + e];
            float gGr[9], gGt[3];
#pragma unroll
            for (int e = 0; e < 9; ++e) gGr[e] = acc[e];
            gGt[0] = acc[9]; gGt[1] = acc[10]; gGt[2] = acc[11];
            mat3T_mul(Pg, gGr, gR);                      // g_R_j = Gr_p^T g_Gr_j
            float d[3] = {J[0] - Pj[0], J[1] - Pj[1], J[2] - Pj[2]};
            float up[9];
            mat3_mulT(gGr, R, up);                       // g_Gr_p += g_Gr_j R_j^T + g_Gt_j (x) d
            float* pacc = s_acc[hl][pj];
            float gd[3];
            mat3T_vec(Pg, gGt, gd);
            if (topo.nchild[pj] == 1) {
                // the only child of its parent (every finger joint but the first): nobody else touches the parent's
                // cotangent at this level - plain adds instead of shared-memory float atomics (compare-and-swap loops)
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) pacc[3 * r + c] += up[3 * r + c] + gGt[r] * d[c];
#pragma unroll
                for (int e = 0; e < 3; ++e) { pacc[9 + e] += gGt[e]; pacc[12 + e] -= gd[e]; }
            } else {
                // several children (the five fingers at the wrist): shared-memory float atomics would be contended
                // compare-and-swap loops.  The child parks its contribution in its own slot instead - entries 0..11
                // (its consumed g_Gr, g_Gt) are free now - and the parent adds its children's slots in ascending
                // joint order below (fixed summation order; the -gd term is recomputed there from the parked g_Gt).
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) acc[3 * r + c] = up[3 * r + c] + gGt[r] * d[c];
                // acc[9..11] already hold g_Gt
            }
#pragma unroll
            for (int e = 0; e < 3; ++e) acc[12 + e] += gd[e];        // own slot
        }
        __syncwarp();
        if (topo.level[j] == lvl - 1 && topo.nchild[j] > 1) {
            for (int c = 1; c < NJ; ++c)
                if (topo.parents[c] == j) {
                    const float* slot = s_acc[hl][c];
                    float cg[3] = {slot[9], slot[10], slot[11]}, gdc[3];
#pragma unroll
                    for (int e = 0; e < 9; ++e) acc[e] += slot[e];
                    mat3T_vec(Gr, cg, gdc);                          // this joint's Gr is the child's Pg
#pragma unroll
                    for (int e = 0; e < 3; ++e) { acc[9 + e] += cg[e]; acc[12 + e] -= gdc[e]; }
                }
        }
        __syncwarp();
    }
    if (j == 0) {
#pragma unroll
        for (int e = 0; e < 9; ++e) gR[e] = acc[e];
        acc[12] += acc[9]; acc[13] += acc[10]; acc[14] += acc[11];   // Gt_0 = J_0
    }
    // pose blend shapes: pose_feature = Rs - I
    if (j >= 1) {
#pragma unroll
        for (int e = 0; e < 9; ++e) gR[e] += s_gx[hl][10 + 9 * (j - 1) + e];
    }
    __syncwarp();
    if (j == 0 && p.quat_dim == 4) {
        float q[4], gq[4];
        for (int i = 0; i < 4; ++i) q[i] = p.quat[(size_t)hh * p.ld_quat + i];
        quat_to_mat_bwd(q, gR, gq);
        if (live) for (int i = 0; i < 4; ++i) g.quat[(size_t)hh * g.ld_quat + i] = gq[i];
    } else {
        float gt[3];
        rodrigues_bwd(ang, gR, gt);
        if (j == 0) {
            if (live) for (int i = 0; i < 3; ++i) g.quat[(size_t)hh * g.ld_quat + i] = gt[i];
        } else {
            s_gang[hl][3 * (j - 1)] = gt[0];
            s_gang[hl][3 * (j - 1) + 1] = gt[1];
            s_gang[hl][3 * (j - 1) + 2] = gt[2];
        }
    }
    __syncwarp();
    if (live) {
        for (int k = j; k < p.ncomp; k += NJ) {       // angles = theta . comp[:ncomp] + mean
            float a = 0.f;
            for (int c = 0; c < 45; ++c) a = fmaf(s_comp[k * 45 + c], s_gang[hl][c], a);
            g.theta[(size_t)hh * g.ld_theta + k] = a;
        }
        if (j < 10) {                                  // direct blend-shape term + rest-joint term
            float a = s_gx[hl][j];
            for (int i = 0; i < NJ; ++i) {
                const float* gj = &s_acc[hl][i][12];
                const float* js = s_JS + (j * NJ + i) * 3;
                a += js[0] * gj[0] + js[1] * gj[1] + js[2] * gj[2];
            }
            g.beta[(size_t)hh * g.ld_beta + j] = a;
        }
    }
    // loss totals of the step (last kernel of the chain): block 0 reduces the per-tile sums in a fixed order
    if (lf.totals && blockIdx.x == 0) {
        __shared__ float s_red[POSE_HPB * NJ / 32][3];
        float sum = 0.f, cnt = 0.f, per = 0.f;
        for (int b = threadIdx.x; b < lf.n_mesh; b += POSE_HPB * NJ) {
            float a = 0.f, c = 0.f;
            for (int t = 0; t < lf.n_tiles; ++t) {
                a += lf.parts_tile[((size_t)b * lf.n_tiles + t) * 2];
                c += lf.parts_tile[((size_t)b * lf.n_tiles + t) * 2 + 1];
            }
            if (lf.n_mesh != B) { lf.parts[2 * b] = a; lf.parts[2 * b + 1] = c; }     // multi-view: parts per mesh
            sum += a; cnt += c; per += a / (c + 1e-8f);
        }
        sum = warp_sum(sum); cnt = warp_sum(cnt); per = warp_sum(per);
        if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5][0] = sum; s_red[threadIdx.x >> 5][1] = cnt; s_red[threadIdx.x >> 5][2] = per; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.f, b = 0.f, c = 0.f;
            for (int w = 0; w < POSE_HPB * NJ / 32; ++w) { a += s_red[w][0]; b += s_red[w][1]; c += s_red[w][2]; }
            lf.totals[0] = lf.weight * c / (float)lf.n_mesh;
            lf.totals[1] = a;
            lf.totals[2] = b;
            lf.totals[3] = lf.weight * c;      // un-normalised, for summing over slices / ranks
        }
    }
}

// ------------------------------------------------------------------------------------------------
// entry points
// ------------------------------------------------------------------------------------------------
static int check_params(const DsfManoParams* p) {
    DSF_REQUIRE(p && p->quat && p->theta && p->beta, "null parameter pointers");
    DSF_REQUIRE(p->quat_dim == 3 || p->quat_dim == 4, "quat_dim must be 3 or 4");
    DSF_REQUIRE(p->ncomp >= 0 && p->ncomp <= 45, "ncomp must be in [0,45]");
    return DSF_OK;
}

int dsf_mano_forward_impl(const DsfMano* h, int B, const DsfManoParams* p, float unit_scale, float* verts,
                          float* joints, float* Rs, float* ws, cudaStream_t st) {
    DSF_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "workspace must be 16-byte aligned");
    int rc = ensure_constants();
    if (rc) return rc;
    ChainTopo topo = topo_of(h);
    mano_pose_kernel<<<(B + POSE_HPB - 1) / POSE_HPB, POSE_HPB * NJ, 0, st>>>(B, *p, h->comp, h->mean, h->Jt,
                                                                              h->JS, topo, ws);
    DSF_CHECK_LAUNCH();
    // v_posed = v_template + [beta | Rs - I] . [shapedirs ; posedirs]
    rc = dsf_blend_forward_gemm(B, ws + WS_X, ws + WS_X + BLEND_KPAD, WS_PER_HAND, h->BTh, h->BTl, ws + WS_VP,
                                WS_PER_HAND, h->vt, st);
    if (rc) return rc;
    mano_skin_kernel<<<B, SKIN_T, 0, st>>>(B, ws, h->wv_ptr, h->wv_ent, h->jr_ptr, h->jr_idx, h->jr_w, p->cam, p->ld_cam,
                                           unit_scale, verts, joints, Rs);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

int dsf_mano_backward_impl(const DsfMano* h, int B, const DsfManoParams* p, float unit_scale,
                           const float* verts, const float* joints, const float* g_verts,
                           const float* g_joints, const DsfManoGrads* g, float* ws, const GradTiles* gt,
                           const float* cube, const LossFold* lf, cudaStream_t st) {
    DSF_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "workspace must be 16-byte aligned");
    int rc = ensure_constants();
    if (rc) return rc;
    ChainTopo topo = topo_of(h);
    mano_skin_bwd_kernel<<<B, SKB_T, 0, st>>>(B, ws, h->wv_ptr, h->wv_ent, h->jr_ptr, h->jr_idx, h->jr_w, h->wj_ptr, h->wj_idx,
                                              h->wj_w, p->cam, p->ld_cam,
                                              unit_scale, verts, joints, g_verts, g_joints,
                                              p->cam ? g->cam : nullptr, g->ld_cam, gt ? *gt : GradTiles{}, cube,
                                              lf ? *lf : LossFold{});
    DSF_CHECK_LAUNCH();
    // g_X = g_vposed . basis^T as BLEND_SPLITS split-K partials (summed by the pose backward kernel)
    rc = dsf_blend_backward_gemm(B, ws + WS_GVP, ws + WS_GVPL, WS_PER_HAND, h->Bh, h->Bl, ws + WS_GX, WS_PER_HAND, KP, st);
    if (rc) return rc;
    mano_pose_bwd_kernel<<<(B + POSE_HPB - 1) / POSE_HPB, POSE_HPB * NJ, 0, st>>>(B, *p, *g, h->comp, h->JS,
                                                                                  topo, ws, dsf_blend_backward_splits(B),
                                                                                  lf ? *lf : LossFold{});
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

extern "C" int dsf_mano_forward(const DsfMano* h, int batch, const DsfManoParams* p, float unit_scale,
                                float* verts, float* joints, float* Rs, float* workspace,
                                dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(h && verts && joints && workspace, "null handle/output/workspace");
    DSF_REQUIRE(batch > 0, "batch must be positive");
    int rc = check_params(p);
    if (rc) return rc;
    return dsf_mano_forward_impl(h, batch, p, unit_scale, verts, joints, Rs, workspace, (cudaStream_t)stream);
}

extern "C" int dsf_mano_backward(const DsfMano* h, int batch, const DsfManoParams* p, float unit_scale,
                                 const float* verts, const float* joints, const float* g_verts,
                                 const float* g_joints, const DsfManoGrads* g, float* workspace,
                                 dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(h && verts && joints && workspace && g, "null handle/input/workspace");
    DSF_REQUIRE(batch > 0, "batch must be positive");
    DSF_REQUIRE(g->quat && g->theta && g->beta, "null gradient outputs");
    DSF_REQUIRE(!p || !p->cam || g->cam, "cam given but g.cam is null");
    int rc = check_params(p);
    if (rc) return rc;
    return dsf_mano_backward_impl(h, batch, p, unit_scale, verts, joints, g_verts, g_joints, g, workspace,
                                  nullptr, nullptr, nullptr, (cudaStream_t)stream);
}
