// dsf_b200 - sphere-proxy self-collision term for sm_100a, forward + backward in one kernel.
// Replaces MANO_SMPL.get_sphere_radius / calculate_coll, render_model/mano_layer.py:229-317, :373-385.
// One warp per hand: 16 joint radii (mean of the 10 nearest support vertices), 66 spheres on the
// bones, masked pairwise hinge, the reference's per-row 0.1 gate, and the gradient to the joints
// (centres and radii; the mesh is a constant, every caller detaches it).
#include <math.h>

#include "common.cuh"

#define NS DSF_NSPHERE
#define NPALM 21

// The constant 66x66 pair table (mano_layer.py:239-269): 21 palm spheres (root + 4 per metacarpal)
// followed by 3 spheres for each of the 15 finger bones.
void dsf_build_collision_mask(float* m) {
    const int I = 3;
    for (int i = 0; i < NS * NS; ++i) m[i] = 0.f;
    for (int r = 0; r < NS; ++r)
        for (int c = 0; c < NS; ++c)
            if (r >= NPALM || c >= NPALM) m[r * NS + c] = 1.f;   // palm-palm pairs never collide
    for (int b = 0; b < 15; ++b) {
        const int root = b / 3 + 1;
        for (int k = 0; k < I; ++k) {
            const int row = NPALM + I * b + k;
            if (b % 3 == 0) {            // proximal bone: skip its metacarpal end and the next bone (:252-257)
                m[row * NS + root * 4] = 0.f;
                m[(root * 4) * NS + row] = 0.f;
                for (int c = NPALM + I * b; c < NPALM + I * b + I + 3 && c < NS; ++c) m[row * NS + c] = 0.f;
            } else {                     // other bones: skip the neighbours on the same finger (:259-263)
                int lo = NPALM + I * b - I;
                int hi = NPALM + I * b + 2 * I + 1;
                const int mx = NPALM + 3 * I * root;
                if (hi > mx) hi = mx;
                for (int c = lo; c < hi; ++c) m[row * NS + c] = 0.f;
            }
        }
    }
    const int th = 12 * I;               // thumb root vs palm (:265-269)
    for (int r = NPALM + th; r < NPALM + th + I + 1; ++r)
        for (int c = 0; c < NPALM; ++c) {
            m[r * NS + c] = 0.f;
            m[c * NS + r] = 0.f;
        }
}

__constant__ int c_child[15] = {2, 3, 16, 5, 6, 17, 8, 9, 18, 11, 12, 19, 14, 15, 20};
__constant__ int c_palm_child[5] = {1, 4, 7, 10, 13};
__constant__ float c_palm_t[4] = {0.2f, 0.4f, 0.6f, 0.8f};                 // linspace(0,1,6)[1:-1]
__constant__ float c_fing_t[3] = {0.f, 0.33333334f, 0.6666667f};           // linspace(0,1,4)[:-1]

#define COLL_WARPS 4

// sphere i -> (parent joint, child joint, t); radius uses the same interpolation with the palm root
// radius replaced by the clamped one.
__device__ __forceinline__ void sphere_def(int i, int* pj, int* cj, float* t) {
    if (i == 0) { *pj = 0; *cj = 0; *t = 0.f; return; }
    if (i < NPALM) {
        const int f = (i - 1) / 4, k = (i - 1) % 4;
        *pj = 0; *cj = c_palm_child[f]; *t = c_palm_t[k];
        return;
    }
    const int b = (i - NPALM) / 3, k = (i - NPALM) % 3;
    *pj = b + 1; *cj = c_child[b]; *t = c_fing_t[k];
}

// get_sphere_radius (mano_layer.py:271-317), one warp: 16 joint radii from the joints Jr and the
// mesh, 66 sphere centres from the joints Jc (calculate_coll passes the same joints for both,
// seg_pcl two different sets, :407-408).  All pointers are per-warp shared-memory scratch.
__device__ __forceinline__ float build_sphere_set(const float* Jc, const float* Jr, const float* __restrict__ Mg,
                                                  const int* __restrict__ jr_ptr, const int* __restrict__ jr_idx,
                                                  const float* __restrict__ jr_w, int lane, float* s_dist,
                                                  float* s_jr, float* s_rg, float* s_c, float* s_r, bool* pp_live) {
    // joint radii: mean of the 10 smallest distances over the regressor support (:275-280)
    for (int j = 0; j < NJ; ++j) {
        const float jx = Jr[3 * j], jy = Jr[3 * j + 1], jz = Jr[3 * j + 2];
        const int e0 = jr_ptr[j], e1 = jr_ptr[j + 1];
        for (int e = e0 + lane; e < e1; e += 32) {
            float d = INFINITY;
            if (jr_w[e] > 0.f) {
                const int v = jr_idx[e];
                const float dx = jx - Mg[3 * v], dy = jy - Mg[3 * v + 1], dz = jz - Mg[3 * v + 2];
                d = sqrtf(dx * dx + dy * dy + dz * dz + 1e-8f);
            }
            s_dist[e - e0] = d;
        }
        __syncwarp();
        float sum = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
        for (int pick = 0; pick < 10; ++pick) {
            float best = INFINITY;
            int bi = -1;
            for (int e = lane; e < e1 - e0; e += 32) {
                const float d = s_dist[e];
                if (d < best) { best = d; bi = e; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob < best || (ob == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ob; bi = oi; }
            }
            if (bi < 0) {
                sum += 100.f;            // sentinel of mano_layer.py:279 (non-support vertices)
            } else {
                sum += best;
                const int v = jr_idx[e0 + bi];
                gx += (jx - Mg[3 * v]) / best;
                gy += (jy - Mg[3 * v + 1]) / best;
                gz += (jz - Mg[3 * v + 2]) / best;
                __syncwarp();
                if (lane == 0) s_dist[bi] = INFINITY;
                __syncwarp();
            }
        }
        if (lane == 0) {
            s_jr[j] = sum * 0.1f;
            s_rg[3 * j] = gx * 0.1f; s_rg[3 * j + 1] = gy * 0.1f; s_rg[3 * j + 2] = gz * 0.1f;
        }
        __syncwarp();
    }
    if (lane < 5) s_jr[NJ + lane] = s_jr[3 * lane + 3] / 1.5f;         // fingertips (:281)
    __syncwarp();
    const float jr0 = s_jr[0] - 0.05f;
    const float pp = fminf(fmaxf(jr0, 0.01f), 0.4f);                          // palm root radius (:285)
    *pp_live = jr0 >= 0.01f && jr0 <= 0.4f;

    // 66 spheres
    for (int i = lane; i < NS; i += 32) {
        int pj, cj; float t;
        sphere_def(i, &pj, &cj, &t);
        const float rp = (pj == 0) ? pp : s_jr[pj];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float a = Jc[3 * pj + c], b = Jc[3 * cj + c];
            s_c[3 * i + c] = (i == 0) ? a : (b - a) * t + a;
        }
        s_r[i] = (i == 0) ? pp : (s_jr[cj] - rp) * t + rp;
    }
    __syncwarp();

    return pp;
}

__global__ void __launch_bounds__(COLL_WARPS * 32)
coll_kernel(int B, const float* __restrict__ joints, const float* __restrict__ mesh,
            const int* __restrict__ jr_ptr, const int* __restrict__ jr_idx, const float* __restrict__ jr_w,
            const float* __restrict__ mask, float inv_norm, float* __restrict__ per_hand,
            float* __restrict__ g_joints) {
    __shared__ float s_dist[COLL_WARPS][NV];
    __shared__ float s_J[COLL_WARPS][NJOUT * 3];
    __shared__ float s_jr[COLL_WARPS][NJOUT];
    __shared__ float s_rg[COLL_WARPS][NJ * 3];       // d radius_j / d J_j
    __shared__ float s_c[COLL_WARPS][NS * 3];
    __shared__ float s_r[COLL_WARPS][NS];
    __shared__ float s_gate[COLL_WARPS][NS];
    __shared__ float s_gJ[COLL_WARPS][NJOUT * 3];
    __shared__ float s_gjr[COLL_WARPS][NJOUT];
    const int w = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int hand = blockIdx.x * COLL_WARPS + w;
    if (hand >= B) return;
    const float* Jg = joints + (size_t)hand * NJOUT * 3;
    const float* Mg = mesh + (size_t)hand * NVW * 3;
    for (int i = lane; i < NJOUT * 3; i += 32) { s_J[w][i] = Jg[i]; s_gJ[w][i] = 0.f; }
    if (lane < NJOUT) s_gjr[w][lane] = 0.f;
    __syncwarp();

    bool pp_live;
    const float pp = build_sphere_set(s_J[w], s_J[w], Mg, jr_ptr, jr_idx, jr_w, lane, s_dist[w], s_jr[w], s_rg[w],
                                      s_c[w], s_r[w], &pp_live);
    (void)pp;

    // pass 1: row sums and the per-row gate (:378-384)
    float tot_raw = 0.f, tot_gated = 0.f;
    for (int i = lane; i < NS; i += 32) {
        const float cx = s_c[w][3 * i], cy = s_c[w][3 * i + 1], cz = s_c[w][3 * i + 2], ri = s_r[w][i];
        float row = 0.f;
        for (int j = 0; j < NS; ++j) {
            const float dx = cx - s_c[w][3 * j], dy = cy - s_c[w][3 * j + 1], dz = cz - s_c[w][3 * j + 2];
            const float dis = sqrtf(dx * dx + dy * dy + dz * dz + 1e-8f);
            const float e = fmaxf(ri + s_r[w][j] - dis, 0.f) * __ldg(mask + i * NS + j);
            row += e;
        }
        const float gate = row < 0.1f ? 1.f : 0.f;
        s_gate[w][i] = gate;
        tot_raw += row;
        tot_gated += row * gate;
    }
    tot_raw = warp_sum(tot_raw);
    tot_gated = warp_sum(tot_gated);
    if (lane == 0) { per_hand[2 * hand] = tot_raw; per_hand[2 * hand + 1] = tot_gated; }
    if (!g_joints) return;
    __syncwarp();

    // pass 2: d loss / d (centre_i, radius_i); pair (i,j) is counted once from row i (gate_i) and
    // once from row j (gate_j), the mask is symmetric
    for (int i = lane; i < NS; i += 32) {
        const float cx = s_c[w][3 * i], cy = s_c[w][3 * i + 1], cz = s_c[w][3 * i + 2], ri = s_r[w][i];
        const float gi = s_gate[w][i];
        float gr = 0.f, gcx = 0.f, gcy = 0.f, gcz = 0.f;
        for (int j = 0; j < NS; ++j) {
            const float m = __ldg(mask + i * NS + j);
            const float dx = cx - s_c[w][3 * j], dy = cy - s_c[w][3 * j + 1], dz = cz - s_c[w][3 * j + 2];
            const float dis = sqrtf(dx * dx + dy * dy + dz * dz + 1e-8f);
            if (m != 0.f && ri + s_r[w][j] - dis > 0.f) {
                const float k = inv_norm * m * (gi + s_gate[w][j]);
                gr += k;
                const float q = k / dis;
                gcx -= q * dx; gcy -= q * dy; gcz -= q * dz;
            }
        }
        // sphere -> joints / joint radii (transpose of the interpolation)
        int pj, cj; float t;
        sphere_def(i, &pj, &cj, &t);
        if (i == 0) {
            atomicAdd(&s_gJ[w][0], gcx); atomicAdd(&s_gJ[w][1], gcy); atomicAdd(&s_gJ[w][2], gcz);
            atomicAdd(&s_gjr[w][0], gr);            // index 0 accumulates d/d pp
        } else {
            atomicAdd(&s_gJ[w][3 * pj], gcx * (1.f - t));
            atomicAdd(&s_gJ[w][3 * pj + 1], gcy * (1.f - t));
            atomicAdd(&s_gJ[w][3 * pj + 2], gcz * (1.f - t));
            atomicAdd(&s_gJ[w][3 * cj], gcx * t);
            atomicAdd(&s_gJ[w][3 * cj + 1], gcy * t);
            atomicAdd(&s_gJ[w][3 * cj + 2], gcz * t);
            atomicAdd(&s_gjr[w][pj], gr * (1.f - t));
            atomicAdd(&s_gjr[w][cj], gr * t);
        }
    }
    __syncwarp();
    if (lane < 5) atomicAdd(&s_gjr[w][3 * lane + 3], s_gjr[w][NJ + lane] / 1.5f);
    __syncwarp();
    if (lane < NJ) {
        float g = s_gjr[w][lane];
        if (lane == 0 && !pp_live) g = 0.f;          // clamp passes no gradient outside [0.01,0.4]
        s_gJ[w][3 * lane] += g * s_rg[w][3 * lane];
        s_gJ[w][3 * lane + 1] += g * s_rg[w][3 * lane + 1];
        s_gJ[w][3 * lane + 2] += g * s_rg[w][3 * lane + 2];
    }
    __syncwarp();
    float* go = g_joints + (size_t)hand * NJOUT * 3;
    for (int i = lane; i < NJOUT * 3; i += 32) go[i] = s_gJ[w][i];
}

__global__ void coll_reduce_kernel(int B, const float* __restrict__ per_hand, float inv_norm,
                                   float* __restrict__ out_loss) {
    __shared__ float red[8];
    float a = 0.f;
    for (int b = threadIdx.x; b < B; b += 256) a += per_hand[2 * b + 1];
    a = warp_sum(a);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        out_loss[0] = s * inv_norm;
    }
}

int dsf_coll_impl(const DsfMano* h, int B, const float* joints, const float* mesh, float* out_loss,
                  float* per_hand, float* g_joints, cudaStream_t st) {
    const float inv_norm = 1.f / ((float)B * (float)NS);    // torch.mean over (B,66) row sums (:385)
    coll_kernel<<<(B + COLL_WARPS - 1) / COLL_WARPS, COLL_WARPS * 32, 0, st>>>(
        B, joints, mesh, h->jr_ptr, h->jr_idx, h->jr_w, h->coll_mask, inv_norm, per_hand, g_joints);
    DSF_CHECK_LAUNCH();
    coll_reduce_kernel<<<1, 256, 0, st>>>(B, per_hand, inv_norm, out_loss);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

extern "C" int dsf_coll_forward_backward(const DsfMano* h, int batch, const float* joints, const float* mesh,
                                         float* out_loss, float* per_hand, float* g_joints,
                                         dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(h && joints && mesh && out_loss && per_hand, "null argument");
    DSF_REQUIRE(batch > 0, "batch must be positive");
    return dsf_coll_impl(h, batch, joints, mesh, out_loss, per_hand, g_joints, (cudaStream_t)stream);
}


// ------------------------------------------------------------------------------------------------
// sphere set as a stand-alone product (for seg_pcl, "next" row f2): centres (B,66,3), radii (B,66)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(COLL_WARPS * 32)
sphere_set_kernel(int B, const float* __restrict__ joints_c, const float* __restrict__ joints_r,
                  const float* __restrict__ mesh, const int* __restrict__ jr_ptr, const int* __restrict__ jr_idx,
                  const float* __restrict__ jr_w, float* __restrict__ out_c, float* __restrict__ out_r) {
    __shared__ float s_dist[COLL_WARPS][NV];
    __shared__ float s_Jc[COLL_WARPS][NJOUT * 3];
    __shared__ float s_Jr[COLL_WARPS][NJOUT * 3];
    __shared__ float s_jr[COLL_WARPS][NJOUT];
    __shared__ float s_rg[COLL_WARPS][NJ * 3];
    __shared__ float s_c[COLL_WARPS][NS * 3];
    __shared__ float s_r[COLL_WARPS][NS];
    const int w = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int hand = blockIdx.x * COLL_WARPS + w;
    if (hand >= B) return;
    for (int i = lane; i < NJOUT * 3; i += 32) {
        s_Jc[w][i] = joints_c[(size_t)hand * NJOUT * 3 + i];
        s_Jr[w][i] = joints_r[(size_t)hand * NJOUT * 3 + i];
    }
    __syncwarp();
    bool pp_live;
    (void)build_sphere_set(s_Jc[w], s_Jr[w], mesh + (size_t)hand * NVW * 3, jr_ptr, jr_idx, jr_w, lane, s_dist[w],
                           s_jr[w], s_rg[w], s_c[w], s_r[w], &pp_live);
    for (int i = lane; i < NS * 3; i += 32) out_c[(size_t)hand * NS * 3 + i] = s_c[w][i];
    for (int i = lane; i < NS; i += 32) out_r[(size_t)hand * NS + i] = s_r[w][i];
}

extern "C" int dsf_sphere_set(const DsfMano* h, int batch, const float* joints_centres, const float* joints_radii,
                              const float* mesh, float* centres, float* radii, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(h && joints_centres && joints_radii && mesh && centres && radii, "null argument");
    DSF_REQUIRE(batch > 0, "batch must be positive");
    sphere_set_kernel<<<(batch + COLL_WARPS - 1) / COLL_WARPS, COLL_WARPS * 32, 0, (cudaStream_t)stream>>>(
        batch, joints_centres, joints_radii, mesh, h->jr_ptr, h->jr_idx, h->jr_w, centres, radii);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}
