// dsf_b200 - point-to-mesh-face squared distance for sm_100a (forward arg-min + backward).
// Replaces pytorch3d-0.4.0 _C.point_face_dist_forward/_backward as wrapped by metric/meshLoss.py:21-70
// (ICPLoss :347-353, JointICPLoss :377-394).  DSF always passes one face list for the whole batch, so
// no packing / first_idx tables are needed.  Two forward kernels with identical results (the exhaustive scan's
// minimum and lowest-index arg-min): point_face_fwd_all_kernel - one CTA per hand stages every face record and a
// box hierarchy once, warps pull 32-point batches and evaluate the surviving pairs with full warps (sorted points
// and faces, meshes up to 2047 faces: the MANO case) - and point_face_fwd_kernel, thread per point over chunks of
// staged records (any mesh size, unsorted input).
#include <math.h>

#include "common.cuh"

#define PF_EPS 1e-8f
#define PF_THREADS 256
#define PF_CHUNK 224     // faces staged per pass: 224 records * 96 B = 21 KB of shared memory (measured best
                         // against 448 / 160 records and 128-thread CTAs: more resident warps per SM)

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// squared distance to segment v0v1 (PointLine3DistanceForward); tt = clamped parameter, -1 if degenerate
__device__ __forceinline__ float line3(V3 p, V3 v0, V3 v1, float* tt) {
    V3 d = v1 - v0;
    float l2 = dot(d, d);
    if (l2 <= PF_EPS) {
        V3 q = p - v1;
        *tt = -1.f;
        return dot(q, q);
    }
    float t = dot(d, p - v0) / l2;
    t = fminf(fmaxf(t, 0.f), 1.f);
    V3 q = p - (v0 + d * t);
    *tt = t;
    return dot(q, q);
}

// PointTriangle3DistanceForward; branch 0 = interior (plane distance), 1/2/3 = edge v0v1 / v0v2 / v1v2
__device__ __forceinline__ float point_tri(V3 p, V3 v0, V3 v1, V3 v2, int* branch) {
    V3 e1 = v1 - v0, e2 = v2 - v0;
    V3 n = cross(e2, e1);
    float nn = sqrtf(dot(n, n));
    V3 nh = n * (1.f / (nn + PF_EPS));
    float t = dot(v0 - p, nh);
    V3 c = (p + nh * t) - v0;
    float d00 = dot(e1, e1), d01 = dot(e1, e2), d11 = dot(e2, e2), d20 = dot(c, e1), d21 = dot(c, e2);
    float den = d00 * d11 - d01 * d01 + PF_EPS;
    float w1 = (d11 * d20 - d01 * d21) / den;
    float w2 = (d00 * d21 - d01 * d20) / den;
    float w0 = 1.f - w1 - w2;
    bool inside = w0 >= 0.f && w0 <= 1.f && w1 >= 0.f && w1 <= 1.f && w2 >= 0.f && w2 <= 1.f;
    if (inside && nn > PF_EPS) {
        *branch = 0;
        return t * t;
    }
    float tt;
    float e01 = line3(p, v0, v1, &tt), e02 = line3(p, v0, v2, &tt), e12 = line3(p, v1, v2, &tt);
    float d = e01;
    *branch = 1;
    if (d > e02) { d = e02; *branch = 2; }
    if (d > e12) { d = e12; *branch = 3; }
    return d;
}

// Per-face record staged in shared memory once per chunk so that the inner (point, face) test is
// pure multiply-add work: no square root, no division (they are hoisted into the record).
#define PF_REC 24
// [0..2] v0  [3..5] e1=v1-v0  [6..8] e2=v2-v0  [9..11] unit normal  [12] d00 [13] d01 [14] d11
// [15] 1/(d00 d11 - d01^2 + eps)  [16] 1/|e1|^2  [17] 1/|e2|^2  [18] 1/|v2-v1|^2  (0 = degenerate edge)
// [19] 1 if |n| > eps  [20..22] bounding-sphere centre - v0  [23] bounding-sphere radius
__device__ __forceinline__ void build_face_record(const float* vb, const int* faces, int f, float* r) {
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    const V3 v0 = v3(vb[3 * i0], vb[3 * i0 + 1], vb[3 * i0 + 2]);
    const V3 v1 = v3(vb[3 * i1], vb[3 * i1 + 1], vb[3 * i1 + 2]);
    const V3 v2 = v3(vb[3 * i2], vb[3 * i2 + 1], vb[3 * i2 + 2]);
    const V3 e1 = v1 - v0, e2 = v2 - v0, e12 = v2 - v1;
    const V3 n = cross(e2, e1);
    const float nn = sqrtf(dot(n, n));
    const V3 nh = n * (1.f / (nn + PF_EPS));
    const float d00 = dot(e1, e1), d01 = dot(e1, e2), d11 = dot(e2, e2), l12 = dot(e12, e12);
    r[0] = v0.x; r[1] = v0.y; r[2] = v0.z;
    r[3] = e1.x; r[4] = e1.y; r[5] = e1.z;
    r[6] = e2.x; r[7] = e2.y; r[8] = e2.z;
    r[9] = nh.x; r[10] = nh.y; r[11] = nh.z;
    r[12] = d00; r[13] = d01; r[14] = d11;
    r[15] = 1.f / (d00 * d11 - d01 * d01 + PF_EPS);
    r[16] = d00 <= PF_EPS ? 0.f : 1.f / d00;
    r[17] = d11 <= PF_EPS ? 0.f : 1.f / d11;
    r[18] = l12 <= PF_EPS ? 0.f : 1.f / l12;
    r[19] = nn > PF_EPS ? 1.f : 0.f;
    // Bounding sphere for the cull of the forward scan.  The reference's inside test divides by
    // den + eps, so its barycentrics are the true ones times den / (den + eps): it accepts the plane
    // distance t^2 for projections inside the triangle SCALED about v0 by k = (den + eps) / den.  The sphere
    // (centroid, farthest vertex) therefore covers that scaled triangle, which contains the real one;
    // |p - c| - r is then a lower bound of whatever distance the full test can return.  Ill-conditioned
    // faces (den -> 0, or den lost in its rounding noise) get an unbounded radius and are never culled.
    // den = |e1|^2 |e2|^2 sin^2(angle) is only trusted when it stands clear of its own rounding noise
    // (about 1e-7 d00 d11): below a 0.6 degree corner angle the fp32 barycentrics of the reference are noise
    // themselves and the face is never culled; above it 1 % of slack on k covers the remaining error.
    const float den = d00 * d11 - d01 * d01;
    const float k = den > 1e-4f * d00 * d11 && den > 1e-30f ? 1.01f * (den + PF_EPS) / den : INFINITY;
    const V3 s1 = e1 * k, s2 = e2 * k;
    const V3 cr = (s1 + s2) * (1.f / 3.f);
    const V3 a1 = cr - s1, a2 = cr - s2;
    const float rr = sqrtf(fmaxf(dot(cr, cr), fmaxf(dot(a1, a1), dot(a2, a2))));
    const bool ok = k < 1e6f && rr < 1e18f;
    r[20] = ok ? cr.x : 0.f; r[21] = ok ? cr.y : 0.f; r[22] = ok ? cr.z : 0.f;
    r[23] = ok ? rr : 1e18f;
}

// squared distance from the point with offset a = p - origin to the segment origin + t d, t in [0,1];
// inv_l2 == 0 marks a degenerate segment: distance to its far end (PointLine3DistanceForward)
__device__ __forceinline__ float seg_d2(V3 a, V3 d, float inv_l2) {
    float t = inv_l2 == 0.f ? 1.f : fminf(fmaxf(dot(d, a) * inv_l2, 0.f), 1.f);
    const V3 q = a - d * t;
    return dot(q, q);
}

// Optional pre-pass: order the points of each hand by a 16^3 grid cell of their bounding box (counting
// sort with shared-memory integer atomics), so that the 32 points of a warp are neighbours in space and
// the per-face sphere test below rejects the same faces for the whole warp.  order (B,P) int32.
#define PS_THREADS 256
#define PS_GRID 16
__global__ void __launch_bounds__(PS_THREADS)
point_sort_kernel(int P, const float* __restrict__ points, int* __restrict__ order) {
    __shared__ int s_hist[PS_GRID * PS_GRID * PS_GRID];
    __shared__ float s_red[PS_THREADS / 32][6];
    __shared__ int s_warp[PS_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* pp = points + (size_t)b * P * 3;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < P; i += PS_THREADS)
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], pp[3 * i + k]); hi[k] = fmaxf(hi[k], pp[3 * i + k]); }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 3; ++k) { s_red[warp][k] = lo[k]; s_red[warp][3 + k] = hi[k]; }
    for (int i = tid; i < PS_GRID * PS_GRID * PS_GRID; i += PS_THREADS) s_hist[i] = 0;
    __syncthreads();
    float inv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int w = 0; w < PS_THREADS / 32; ++w) { lo[k] = fminf(lo[k], s_red[w][k]); hi[k] = fmaxf(hi[k], s_red[w][3 + k]); }
        inv[k] = hi[k] > lo[k] ? (float)PS_GRID / (hi[k] - lo[k]) : 0.f;
    }
    auto cell = [&](int i) {
        int c[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int q = (int)((pp[3 * i + k] - lo[k]) * inv[k]);
            c[k] = q < 0 ? 0 : (q > PS_GRID - 1 ? PS_GRID - 1 : q);      // NaN / inf land in a border cell
        }
        // Morton (Z-order) key: 32 consecutive points form a compact 3-D blob, not a strip
        int key = 0;
#pragma unroll
        for (int bit = 0; bit < 4; ++bit)
            key |= (((c[0] >> bit) & 1) << (3 * bit)) | (((c[1] >> bit) & 1) << (3 * bit + 1)) |
                   (((c[2] >> bit) & 1) << (3 * bit + 2));
        return key;
    };
    for (int i = tid; i < P; i += PS_THREADS) atomicAdd(&s_hist[cell(i)], 1);
    __syncthreads();
    // exclusive scan of the 4096 bins: 16 consecutive bins per thread
    const int per = PS_GRID * PS_GRID * PS_GRID / PS_THREADS, b0 = tid * per;
    int local = 0;
    for (int i = 0; i < per; ++i) local += s_hist[b0 + i];
    int incl = local;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int run = incl - local;
    for (int w = 0; w < warp; ++w) run += s_warp[w];
    for (int i = 0; i < per; ++i) { const int h = s_hist[b0 + i]; s_hist[b0 + i] = run; run += h; }
    __syncthreads();
    for (int i = tid; i < P; i += PS_THREADS) order[(size_t)b * P + atomicAdd(&s_hist[cell(i)], 1)] = i;
}

// Companion pre-pass: order the FACES of each hand the same way (Morton cell of the centroid in the mesh's own
// bounding box), so that consecutive faces are neighbours in space and groups of PF_GROUP of them fit a small
// common bounding sphere: the scan below then rejects whole groups with one test.  face_order (B,F) int32.
__global__ void __launch_bounds__(PS_THREADS)
face_sort_kernel(int V, int F, const float* __restrict__ verts, const int* __restrict__ faces, int* __restrict__ order) {
    __shared__ int s_hist[PS_GRID * PS_GRID * PS_GRID];
    __shared__ float s_red[PS_THREADS / 32][6];
    __shared__ int s_warp[PS_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* vb = verts + (size_t)b * V * 3;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < V; i += PS_THREADS)
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], vb[3 * i + k]); hi[k] = fmaxf(hi[k], vb[3 * i + k]); }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 3; ++k) { s_red[warp][k] = lo[k]; s_red[warp][3 + k] = hi[k]; }
    for (int i = tid; i < PS_GRID * PS_GRID * PS_GRID; i += PS_THREADS) s_hist[i] = 0;
    __syncthreads();
    float inv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int w = 0; w < PS_THREADS / 32; ++w) { lo[k] = fminf(lo[k], s_red[w][k]); hi[k] = fmaxf(hi[k], s_red[w][3 + k]); }
        inv[k] = hi[k] > lo[k] ? (float)PS_GRID / (hi[k] - lo[k]) : 0.f;
    }
    auto cell = [&](int f) {
        const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
        int c[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float m = (vb[3 * i0 + k] + vb[3 * i1 + k] + vb[3 * i2 + k]) * (1.f / 3.f);
            const int q = (int)((m - lo[k]) * inv[k]);
            c[k] = q < 0 ? 0 : (q > PS_GRID - 1 ? PS_GRID - 1 : q);
        }
        int key = 0;
#pragma unroll
        for (int bit = 0; bit < 4; ++bit)
            key |= (((c[0] >> bit) & 1) << (3 * bit)) | (((c[1] >> bit) & 1) << (3 * bit + 1)) |
                   (((c[2] >> bit) & 1) << (3 * bit + 2));
        return key;
    };
    for (int f = tid; f < F; f += PS_THREADS) atomicAdd(&s_hist[cell(f)], 1);
    __syncthreads();
    const int per = PS_GRID * PS_GRID * PS_GRID / PS_THREADS, b0 = tid * per;
    int local = 0;
    for (int i = 0; i < per; ++i) local += s_hist[b0 + i];
    int incl = local;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int run = incl - local;
    for (int w = 0; w < warp; ++w) run += s_warp[w];
    for (int i = 0; i < per; ++i) { const int h = s_hist[b0 + i]; s_hist[b0 + i] = run; run += h; }
    __syncthreads();
    for (int f = tid; f < F; f += PS_THREADS) order[(size_t)b * F + atomicAdd(&s_hist[cell(f)], 1)] = f;
}

#define PF_GROUP 8       // faces per group sphere (PF_CHUNK is a multiple)
#ifndef PF_SUPER
#define PF_SUPER 2       // groups per super-group box (measured: 1 -> 2.06 ms, 2 -> 1.95, 3 -> 1.94, 4 -> 1.99, 8 -> 2.16)
#endif
#define PF_MAXCH 32      // chunk reordering is used up to this many chunks (7 for the MANO mesh)

// squared distance from p to the box a = (lo.xyz, hi.x), b = (hi.y, hi.z)
__device__ __forceinline__ float box_d2(V3 p, float4 a, float4 b) {
    const float dx = fmaxf(fmaxf(a.x - p.x, p.x - a.w), 0.f);
    const float dy = fmaxf(fmaxf(a.y - p.y, p.y - b.x), 0.f);
    const float dz = fmaxf(fmaxf(a.z - p.z, p.z - b.y), 0.f);
    return dx * dx + dy * dy + dz * dz;
}

// STATS: the same scan, additionally counting per category how many (point, face) pairs were only sphere-tested,
// evaluated on the interior branch, and evaluated on the edge branch (bench.py's FP32 roofline of config C4)
template <bool STATS>
__global__ void __launch_bounds__(PF_THREADS)
point_face_fwd_kernel(int P, int V, int F, const float* __restrict__ points, const float* __restrict__ verts,
                      const int* __restrict__ faces, const int* __restrict__ order, const int* __restrict__ face_order,
                      float* __restrict__ dists, int* __restrict__ idxs, unsigned long long* __restrict__ stats) {
    unsigned int n_cull = 0, n_in = 0, n_edge = 0, n_grp = 0;
    __shared__ float4 s_gbox[PF_CHUNK / PF_GROUP][2];   // bounding box of each group of PF_GROUP staged faces
    __shared__ float4 s_sbox[(PF_CHUNK / PF_GROUP + PF_SUPER - 1) / PF_SUPER][2];   // ... of each PF_SUPER groups
    __shared__ unsigned short s_fid[PF_CHUNK];          // original face id of each staged record
    const int* fo = face_order ? face_order + (size_t)blockIdx.y * F : nullptr;
    __shared__ __align__(16) float s_rec[PF_CHUNK * PF_REC];
    const int b = blockIdx.y;
    const int slot = blockIdx.x * PF_THREADS + threadIdx.x;
    const bool live = slot < P;
    const int pi = live ? (order ? order[(size_t)b * P + slot] : slot) : 0;
    const float* pp = points + ((size_t)b * P + pi) * 3;
    const V3 p = v3(pp[0], pp[1], pp[2]);
    const float* vb = verts + (size_t)b * V * 3;
    float best = INFINITY, sb = INFINITY;          // sb = sqrt(best), refreshed when best improves
    int bi = -1;
    // Chunk order: faces and points are both in spatial (Morton) order, so the chunk of staged faces whose centroid
    // is nearest to the centroid of this CTA's 256 points almost always holds every point's nearest face.  Scanning
    // it first makes `best` tight at once and the later chunks fall to the group test; the minimum and (through the
    // lowest-index tie rule) the arg-min do not depend on the order.
    const int n_chunks = (F + PF_CHUNK - 1) / PF_CHUNK;
    __shared__ float s_cc[PF_MAXCH][3];
    __shared__ float s_pc[PF_THREADS / 32][4];
    __shared__ unsigned char s_ord[PF_MAXCH];
    const bool reorder = fo != nullptr && n_chunks > 1 && n_chunks <= PF_MAXCH;
    if (reorder) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int c = warp; c < n_chunks; c += PF_THREADS / 32) {
            const int a0 = c * PF_CHUNK, a1 = min(F, a0 + PF_CHUNK);
            float cx = 0.f, cy = 0.f, cz = 0.f;
            for (int i = a0 + lane; i < a1; i += 32) {
                const int fid = fo[i];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float* vv = vb + 3 * faces[3 * fid + k];
                    cx += vv[0]; cy += vv[1]; cz += vv[2];
                }
            }
            cx = warp_sum(cx); cy = warp_sum(cy); cz = warp_sum(cz);
            const float inv = 1.f / (3.f * (float)(a1 - a0));
            if (lane == 0) { s_cc[c][0] = cx * inv; s_cc[c][1] = cy * inv; s_cc[c][2] = cz * inv; }
        }
        const float px = warp_sum(live ? p.x : 0.f), py = warp_sum(live ? p.y : 0.f), pz = warp_sum(live ? p.z : 0.f);
        const float pn = warp_sum(live ? 1.f : 0.f);
        if (lane == 0) { s_pc[warp][0] = px; s_pc[warp][1] = py; s_pc[warp][2] = pz; s_pc[warp][3] = pn; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            for (int w = 0; w < PF_THREADS / 32; ++w)
                for (int k = 0; k < 4; ++k) c[k] += s_pc[w][k];
            const float inv = c[3] > 0.f ? 1.f / c[3] : 0.f;
            float d2[PF_MAXCH];
            for (int k = 0; k < n_chunks; ++k) {
                const float dx = s_cc[k][0] - c[0] * inv, dy = s_cc[k][1] - c[1] * inv, dz = s_cc[k][2] - c[2] * inv;
                d2[k] = dx * dx + dy * dy + dz * dz;
                // insertion sort by distance (comparisons with NaN are false: the result stays a permutation)
                int q = k;
                while (q > 0 && d2[s_ord[q - 1]] > d2[k]) { s_ord[q] = s_ord[q - 1]; --q; }
                s_ord[q] = (unsigned char)k;
            }
        }
        // visibility of s_ord: the __syncthreads() at the top of the chunk loop
    }
    for (int ci = 0; ci < n_chunks; ++ci) {
        __syncthreads();
        const int f0 = (reorder ? (int)s_ord[ci] : ci) * PF_CHUNK;
        const int nf = min(PF_CHUNK, F - f0);
        for (int i = threadIdx.x; i < nf; i += PF_THREADS) {
            const int fid = fo ? fo[f0 + i] : f0 + i;
            s_fid[i] = (unsigned short)fid;
            build_face_record(vb, faces, fid, s_rec + i * PF_REC);
        }
        __syncthreads();
        const int ng = (nf + PF_GROUP - 1) / PF_GROUP;
        const int nsg = (ng + PF_SUPER - 1) / PF_SUPER;
        if (fo) {
            // Group boxes: the axis-aligned box of the members' SCALED triangles (see build_face_record: whatever the
            // full test returns for a face is at least the distance to that triangle, hence to the box), grown by a
            // rounding margin.  A member that must never be culled (radius 1e18) makes the box infinite.  Boxes of
            // PF_GROUP faces, then of PF_SUPER groups: Morton-neighbour faces form elongated strips, which boxes fit
            // far better than spheres, and the test needs no square root of `best`.
            for (int g = threadIdx.x; g < ng; g += PF_THREADS) {
                const int a0 = g * PF_GROUP, a1 = min(nf, a0 + PF_GROUP);
                float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
                for (int i = a0; i < a1; ++i) {
                    const float* r = s_rec + i * PF_REC;
                    const float d00 = r[12], d01 = r[13], d11 = r[14];
                    const float den = d00 * d11 - d01 * d01;
                    const float k = den > 1e-4f * d00 * d11 && den > 1e-30f ? 1.01f * (den + PF_EPS) / den : INFINITY;
                    const bool ok = r[23] < 1e18f;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float s1 = r[3 + c] * k, s2 = r[6 + c] * k;
                        const float l = ok ? r[c] + fminf(0.f, fminf(s1, s2)) : -INFINITY;
                        const float h = ok ? r[c] + fmaxf(0.f, fmaxf(s1, s2)) : INFINITY;
                        lo[c] = fminf(lo[c], l); hi[c] = fmaxf(hi[c], h);
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float m = 1e-6f * fmaxf(fabsf(lo[c]), fabsf(hi[c])) + 1e-12f;
                    lo[c] -= m; hi[c] += m;
                }
                s_gbox[g][0] = make_float4(lo[0], lo[1], lo[2], hi[0]);
                s_gbox[g][1] = make_float4(hi[1], hi[2], 0.f, 0.f);
            }
            __syncthreads();
            for (int q = threadIdx.x; q < nsg; q += PF_THREADS) {
                float4 a = s_gbox[q * PF_SUPER][0], b2 = s_gbox[q * PF_SUPER][1];
                for (int g = q * PF_SUPER + 1; g < min(ng, (q + 1) * PF_SUPER); ++g) {
                    const float4 c = s_gbox[g][0], d = s_gbox[g][1];
                    a.x = fminf(a.x, c.x); a.y = fminf(a.y, c.y); a.z = fminf(a.z, c.z); a.w = fmaxf(a.w, c.w);
                    b2.x = fmaxf(b2.x, d.x); b2.y = fmaxf(b2.y, d.y);
                }
                s_sbox[q][0] = a; s_sbox[q][1] = b2;
            }
            __syncthreads();
        }
        for (int g = 0; g < ng; ++g) {
          if (fo) {
              // squared distance to the box > best: no member can win (a few 1e-4 of slack against rounding)
              const float lim = best * 1.0002f + 1e-30f;
              if (g % PF_SUPER == 0) {
                  if (STATS && live) ++n_grp;
                  if (box_d2(p, s_sbox[g / PF_SUPER][0], s_sbox[g / PF_SUPER][1]) > lim) { g += PF_SUPER - 1; continue; }
              }
              if (STATS && live) ++n_grp;
              if (box_d2(p, s_gbox[g][0], s_gbox[g][1]) > lim) continue;
          }
          const int fa = g * PF_GROUP, fb = min(nf, fa + PF_GROUP);
          for (int f = fa; f < fb; ++f) {
            const float4* r4 = reinterpret_cast<const float4*>(s_rec + f * PF_REC);
            const float4 q0 = r4[0], q5 = r4[5];
            const V3 v0 = v3(q0.x, q0.y, q0.z);
            const V3 a = p - v0;                                 // point relative to v0
            // sphere cull: the face cannot beat `best` when |p - c| >= sqrt(best) + r.  A few 1e-4 of
            // slack keep it conservative against the rounding of either side, so the surviving minimum
            // (and, with the strict < below, the arg-min) is the brute-force one.
            const V3 ac = a - v3(q5.x, q5.y, q5.z);
            const float reach = sb + q5.w;
            if (dot(ac, ac) > reach * reach * 1.0002f + 1e-30f) { if (STATS && live) ++n_cull; continue; }
            const float4 q1 = r4[1], q2 = r4[2], q3 = r4[3], q4 = r4[4];
            const V3 e1 = v3(q0.w, q1.x, q1.y), e2 = v3(q1.z, q1.w, q2.x);
            const V3 nh = v3(q2.y, q2.z, q2.w);
            const float t = -dot(a, nh);                         // signed plane distance (v0 - p) . n
            const V3 c = a + nh * t;                             // projection on the plane, relative to v0
            const float d20 = dot(c, e1), d21 = dot(c, e2);
            const float w1 = (q3.z * d20 - q3.y * d21) * q3.w;   // (d11 d20 - d01 d21) / den
            const float w2 = (q3.x * d21 - q3.y * d20) * q3.w;   // (d00 d21 - d01 d20) / den
            const float w0 = 1.f - w1 - w2;
            float d;
            if (q4.w != 0.f && w0 >= 0.f && w0 <= 1.f && w1 >= 0.f && w1 <= 1.f && w2 >= 0.f && w2 <= 1.f) {
                d = t * t;
                if (STATS && live) ++n_in;
            } else {
                if (STATS && live) ++n_edge;
                const float e01 = seg_d2(a, e1, q4.x);
                const float e02 = seg_d2(a, e2, q4.y);
                const float e12 = seg_d2(a - e1, e2 - e1, q4.z);
                d = fminf(fminf(e01, e02), e12);
            }
            // lowest face index wins ties (the staged order is a permutation when face_order is given)
            const int fid = s_fid[f];
            if (d < best || (d == best && fid < bi)) { best = d; bi = fid; sb = sqrtf(d); }
          }
        }
    }
    if (live) {
        dists[(size_t)b * P + pi] = best;
        idxs[(size_t)b * P + pi] = bi;
    }
    if (STATS) {
        unsigned int c[4] = {n_cull, n_in, n_edge, n_grp};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c[k] += __shfl_xor_sync(0xffffffffu, c[k], o);
            if ((threadIdx.x & 31) == 0 && c[k]) atomicAdd(stats + k, (unsigned long long)c[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Whole-mesh variant of the forward scan (sorted points + sorted faces, meshes whose records fit shared memory -
// the MANO mesh: 1554 faces = 149 KB).  One CTA of 1024 threads per hand stages every face record and the box
// hierarchy ONCE; its 32 warps then pull batches of 32 (spatially neighbouring) points from a shared counter - no
// chunk loop, no CTA barrier inside the scan, slow batches do not hold the others up.  Inside a warp the work is
// re-balanced twice, so that every stage runs with (nearly) full warps instead of whichever lanes happen to need it:
//   stage 1  lane per point, lock step over the super-group / group boxes; a surviving (point, group) pair is
//            appended to queue A (ballot compaction, no atomics)
//   stage 2  four pairs of queue A at a time: lane per (point, face of the group) runs the bounding-sphere test with
//            the owner's point and current bound (shuffles); survivors go to queue B
//   stage 3  whenever 32 pairs wait in queue B: lane per pair, the full point-triangle evaluation, committed as a
//            packed (distance bits << 32 | face id) key with a shared-memory atomicMin - the sequential scan's
//            "lowest face wins ties" rule, independent of the order; every lane then refreshes its bound.
// The thread-per-point scan evaluated with 7 of 32 lanes active (the evaluation ran whenever any lane needed it).
// ------------------------------------------------------------------------------------------------
#define PFA_THREADS 1024
#define PFA_WARPS (PFA_THREADS / 32)
#define PFA_QA (32 * PF_SUPER + 8)   // queue A: up to 32 lanes x PF_SUPER groups per lock step + 3 waiting
#define PFA_QB 64         // queue B: up to 32 new pairs per stage-2 step + 31 waiting

struct PfaSmem {
    float* rec;                    // F * PF_REC
    float4* gbox;                  // ng * 2
    float4* sbox;                  // nsg * 2
    unsigned long long* key;       // PFA_THREADS (32 per warp: the batch in flight)
    unsigned short* fid;           // F
    unsigned short* qa;            // warps * PFA_QA  (lane << 8 | group)
    unsigned short* qb;            // warps * PFA_QB  (lane << 11 | staged face)
    int* next;                     // batch counter
};

static size_t pfa_smem_bytes(int F) {
    const int ng = (F + PF_GROUP - 1) / PF_GROUP, nsg = (ng + PF_SUPER - 1) / PF_SUPER;
    size_t n = (size_t)F * PF_REC * 4 + (size_t)ng * 32 + (size_t)nsg * 32 + (size_t)PFA_THREADS * 8;
    n += ((size_t)F * 2 + 15) & ~(size_t)15;
    n += (size_t)PFA_WARPS * (PFA_QA + PFA_QB) * 2 + 16;
    return n;
}

__device__ __forceinline__ PfaSmem pfa_carve(unsigned char* raw, int F) {
    const int ng = (F + PF_GROUP - 1) / PF_GROUP, nsg = (ng + PF_SUPER - 1) / PF_SUPER;
    PfaSmem s;
    s.rec = reinterpret_cast<float*>(raw);
    s.gbox = reinterpret_cast<float4*>(s.rec + (size_t)F * PF_REC);
    s.sbox = s.gbox + 2 * ng;
    s.key = reinterpret_cast<unsigned long long*>(s.sbox + 2 * nsg);
    s.fid = reinterpret_cast<unsigned short*>(s.key + PFA_THREADS);
    s.qa = s.fid + (((size_t)F + 7) & ~(size_t)7);
    s.qb = s.qa + PFA_WARPS * PFA_QA;
    s.next = reinterpret_cast<int*>(s.qb + PFA_WARPS * PFA_QB);
    return s;
}

// full evaluation of (point p, staged record): the sequential scan's arithmetic, unchanged
template <bool STATS>
__device__ __forceinline__ float pfa_eval(V3 p, const float* rec, unsigned int* n_in, unsigned int* n_edge) {
    const float4* r4 = reinterpret_cast<const float4*>(rec);
    const float4 q0 = r4[0], q1 = r4[1], q2 = r4[2], q3 = r4[3], q4 = r4[4];
    const V3 a = p - v3(q0.x, q0.y, q0.z);
    const V3 e1 = v3(q0.w, q1.x, q1.y), e2 = v3(q1.z, q1.w, q2.x);
    const V3 nh = v3(q2.y, q2.z, q2.w);
    const float t = -dot(a, nh);
    const V3 c = a + nh * t;
    const float d20 = dot(c, e1), d21 = dot(c, e2);
    const float w1 = (q3.z * d20 - q3.y * d21) * q3.w;
    const float w2 = (q3.x * d21 - q3.y * d20) * q3.w;
    const float w0 = 1.f - w1 - w2;
    if (q4.w != 0.f && w0 >= 0.f && w0 <= 1.f && w1 >= 0.f && w1 <= 1.f && w2 >= 0.f && w2 <= 1.f) {
        if (STATS) ++*n_in;
        return t * t;
    }
    if (STATS) ++*n_edge;
    const float e01 = seg_d2(a, e1, q4.x);
    const float e02 = seg_d2(a, e2, q4.y);
    const float e12 = seg_d2(a - e1, e2 - e1, q4.z);
    return fminf(fminf(e01, e02), e12);
}

template <bool STATS>
__global__ void __launch_bounds__(PFA_THREADS, 1)
point_face_fwd_all_kernel(int P, int V, int F, const float* __restrict__ points, const float* __restrict__ verts,
                          const int* __restrict__ faces, const int* __restrict__ order, const int* __restrict__ face_order,
                          float* __restrict__ dists, int* __restrict__ idxs, unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(16) unsigned char pfa_raw[];
    const PfaSmem s = pfa_carve(pfa_raw, F);
    unsigned int n_cull = 0, n_in = 0, n_edge = 0, n_grp = 0;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* fo = face_order + (size_t)b * F;
    const float* vb = verts + (size_t)b * V * 3;
    const int ng = (F + PF_GROUP - 1) / PF_GROUP, nsg = (ng + PF_SUPER - 1) / PF_SUPER;
    for (int i = tid; i < F; i += PFA_THREADS) {
        const int fid = fo[i];
        s.fid[i] = (unsigned short)fid;
        build_face_record(vb, faces, fid, s.rec + (size_t)i * PF_REC);
    }
    if (tid == 0) *s.next = 0;
    __syncthreads();
    // group / super-group boxes: see point_face_fwd_kernel
    for (int g = tid; g < ng; g += PFA_THREADS) {
        const int a0 = g * PF_GROUP, a1 = min(F, a0 + PF_GROUP);
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = a0; i < a1; ++i) {
            const float* r = s.rec + (size_t)i * PF_REC;
            const float d00 = r[12], d01 = r[13], d11 = r[14];
            const float den = d00 * d11 - d01 * d01;
            const float k = den > 1e-4f * d00 * d11 && den > 1e-30f ? 1.01f * (den + PF_EPS) / den : INFINITY;
            const bool ok = r[23] < 1e18f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float s1 = r[3 + c] * k, s2 = r[6 + c] * k;
                const float l = ok ? r[c] + fminf(0.f, fminf(s1, s2)) : -INFINITY;
                const float h = ok ? r[c] + fmaxf(0.f, fmaxf(s1, s2)) : INFINITY;
                lo[c] = fminf(lo[c], l); hi[c] = fmaxf(hi[c], h);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float m = 1e-6f * fmaxf(fabsf(lo[c]), fabsf(hi[c])) + 1e-12f;
            lo[c] -= m; hi[c] += m;
        }
        s.gbox[2 * g] = make_float4(lo[0], lo[1], lo[2], hi[0]);
        s.gbox[2 * g + 1] = make_float4(hi[1], hi[2], 0.f, 0.f);
    }
    __syncthreads();
    for (int q = tid; q < nsg; q += PFA_THREADS) {
        float4 a = s.gbox[2 * q * PF_SUPER], b2 = s.gbox[2 * q * PF_SUPER + 1];
        for (int g = q * PF_SUPER + 1; g < min(ng, (q + 1) * PF_SUPER); ++g) {
            const float4 c = s.gbox[2 * g], d = s.gbox[2 * g + 1];
            a.x = fminf(a.x, c.x); a.y = fminf(a.y, c.y); a.z = fminf(a.z, c.z); a.w = fmaxf(a.w, c.w);
            b2.x = fmaxf(b2.x, d.x); b2.y = fmaxf(b2.y, d.y);
        }
        s.sbox[2 * q] = a; s.sbox[2 * q + 1] = b2;
    }
    __syncthreads();
    // ---- from here on the warps run independently (warp-level synchronisation only)
    unsigned short* qa = s.qa + warp * PFA_QA;
    unsigned short* qb = s.qb + warp * PFA_QB;
    unsigned int* wd = reinterpret_cast<unsigned int*>(s.key) + (warp << 6);      // 32 distance bit patterns ...
    unsigned int* wf = wd + 32;                                                     // ... and 32 face ids per warp
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int n_batches = (P + 31) >> 5;
    for (;;) {
        int wb = 0;
        if (lane == 0) wb = blockIdx.x + gridDim.x * atomicAdd(s.next, 1);
        wb = __shfl_sync(0xffffffffu, wb, 0);
        if (wb >= n_batches) break;
        const int slot = (wb << 5) + lane;
        const bool live = slot < P;
        const int pi = live ? order[(size_t)b * P + slot] : 0;
        const float* pp = points + ((size_t)b * P + pi) * 3;
        const V3 p = v3(pp[0], pp[1], pp[2]);
        wd[lane] = __float_as_uint(INFINITY);
        wf[lane] = 0xffffffffu;
        float best = INFINITY, sb = INFINITY;
        int nA = 0, nB = 0;                                   // queue fill levels (warp-uniform registers)
        __syncwarp();
        // stage 3: evaluate everything waiting in queue B
        auto flush_b = [&]() {
            __syncwarp();
            for (int i0 = 0; i0 < nB; i0 += 32) {
                const int i = i0 + lane;
                const unsigned int e = i < nB ? qb[i] : 0u;
                const int owner = (int)(e >> 11), f = (int)(e & 2047u);
                const V3 q = v3(__shfl_sync(0xffffffffu, p.x, owner), __shfl_sync(0xffffffffu, p.y, owner),
                                __shfl_sync(0xffffffffu, p.z, owner));
                // commit min (distance, face id) per owner with native 32-bit atomics (a 64-bit atomicMin on shared
                // memory is a compare-and-swap loop, and the pairs of one owner sit next to each other): distances
                // first; an owner whose distance improved forgets its face; then the faces of the pairs that hold
                // their owner's minimum.  NaN distances (bit pattern above +inf) never win, as in the sequential scan.
                unsigned int dbits = 0xffffffffu;
                if (i < nB) {
                    dbits = __float_as_uint(pfa_eval<STATS>(q, s.rec + (size_t)f * PF_REC, &n_in, &n_edge));
                    if (dbits < 0x7f800000u) atomicMin(wd + owner, dbits);     // +inf / NaN never win (bi stays -1)
                }
                __syncwarp();
                const unsigned int mine = wd[lane];
                if (mine < __float_as_uint(best)) { best = __uint_as_float(mine); wf[lane] = 0xffffffffu; }
                __syncwarp();
                if (i < nB && dbits < 0x7f800000u && dbits == wd[owner]) atomicMin(wf + owner, (unsigned int)s.fid[f]);
                __syncwarp();
            }
            nB = 0;
            sb = sqrtf(best);
        };
        // stage 2: `take` (<= 4) pairs from the top of queue A, lane per (pair, face of its group)
        auto sphere_step = [&](int take) {
            nA -= take;
            const int j = lane >> 3;
            const unsigned int e = j < take ? qa[nA + j] : 0u;
            const int owner = (int)(e >> 8), f = (int)(e & 255u) * PF_GROUP + (lane & 7);
            const V3 q = v3(__shfl_sync(0xffffffffu, p.x, owner), __shfl_sync(0xffffffffu, p.y, owner),
                            __shfl_sync(0xffffffffu, p.z, owner));
            const float sbo = __shfl_sync(0xffffffffu, sb, owner);
            bool pass = false;
            if (j < take && f < F) {
                const float4* r4 = reinterpret_cast<const float4*>(s.rec + (size_t)f * PF_REC);
                const float4 q0 = r4[0], q5 = r4[5];
                const V3 ac = q - v3(q0.x, q0.y, q0.z) - v3(q5.x, q5.y, q5.z);
                const float reach = sbo + q5.w;
                pass = !(dot(ac, ac) > reach * reach * 1.0002f + 1e-30f);
                if (STATS && !pass) ++n_cull;
            }
            const unsigned int mk = __ballot_sync(0xffffffffu, pass);
            if (pass) qb[nB + __popc(mk & lt_mask)] = (unsigned short)((owner << 11) | f);
            nB += __popc(mk);
            if (nB >= 32) flush_b();
        };
        // start with the super-group nearest to the batch: the first evaluations then give every lane a tight bound
        int start = 0;
        {
            const float wn = warp_sum(live ? 1.f : 0.f);
            const float inv = wn > 0.f ? 1.f / wn : 0.f;
            const V3 c = v3(warp_sum(live ? p.x : 0.f) * inv, warp_sum(live ? p.y : 0.f) * inv, warp_sum(live ? p.z : 0.f) * inv);
            float dmin = INFINITY;
            int imin = 0;
            for (int q = lane; q < nsg; q += 32) {
                const float d = box_d2(c, s.sbox[2 * q], s.sbox[2 * q + 1]);
                if (d < dmin) { dmin = d; imin = q; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float d2 = __shfl_xor_sync(0xffffffffu, dmin, o);
                const int i2 = __shfl_xor_sync(0xffffffffu, imin, o);
                if (d2 < dmin || (d2 == dmin && i2 < imin)) { dmin = d2; imin = i2; }
            }
            start = imin;
        }
        // stage 1: lock step over the boxes
        for (int k = 0; k < nsg; ++k) {
            int q = start + k;
            if (q >= nsg) q -= nsg;
            if (STATS && live) ++n_grp;
            const bool want = live && !(box_d2(p, s.sbox[2 * q], s.sbox[2 * q + 1]) > best * 1.0002f + 1e-30f);
            if (!__any_sync(0xffffffffu, want)) continue;
            const int g1 = min(ng, (q + 1) * PF_SUPER);
            for (int g = q * PF_SUPER; g < g1; ++g) {
                bool in_g = want;
                if (in_g) {
                    if (STATS) ++n_grp;
                    in_g = !(box_d2(p, s.gbox[2 * g], s.gbox[2 * g + 1]) > best * 1.0002f + 1e-30f);
                }
                const unsigned int mk = __ballot_sync(0xffffffffu, in_g);
                if (in_g) qa[nA + __popc(mk & lt_mask)] = (unsigned short)((lane << 8) | g);
                nA += __popc(mk);
            }
            __syncwarp();
            while (nA >= 4) { sphere_step(4); __syncwarp(); }
        }
        while (nA > 0) { sphere_step(min(4, nA)); __syncwarp(); }
        flush_b();
        if (live) {
            dists[(size_t)b * P + pi] = __uint_as_float(wd[lane]);
            idxs[(size_t)b * P + pi] = (int)wf[lane];
        }
        __syncwarp();
    }
    if (STATS) {
        unsigned int c[4] = {n_cull, n_in, n_edge, n_grp};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c[k] += __shfl_xor_sync(0xffffffffu, c[k], o);
            if (lane == 0 && c[k]) atomicAdd(stats + k, (unsigned long long)c[k]);
        }
    }
}

// launch the forward scan: whole-mesh kernel when the points and faces are sorted and the records fit shared memory
template <bool STATS>
static int launch_point_face_fwd(int batch, int P, int V, int F, const float* points, const float* verts, const int* faces,
                                 const int* order, const int* face_order, float* dists, int* idxs,
                                 unsigned long long* stats, cudaStream_t st) {
    const size_t smem = pfa_smem_bytes(F);
    if (order && face_order && F <= 2047 && smem <= 220 * 1024 && P >= 512) {
        static bool attr_set[16][2] = {};
        int dev = 0;
        DSF_CHECK_CUDA(cudaGetDevice(&dev));
        if (dev >= 16 || !attr_set[dev][STATS]) {
            DSF_CHECK_CUDA(cudaFuncSetAttribute(point_face_fwd_all_kernel<STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                220 * 1024));
            if (dev < 16) attr_set[dev][STATS] = true;
        }
        // one CTA per hand; small batches split every hand's point batches over two CTAs to fill the chip
        dim3 grid(batch < 148 ? 2 : 1, batch);
        point_face_fwd_all_kernel<STATS><<<grid, PFA_THREADS, smem, st>>>(P, V, F, points, verts, faces, order, face_order,
                                                                        dists, idxs, stats);
    } else {
        dim3 grid((P + PF_THREADS - 1) / PF_THREADS, batch);
        point_face_fwd_kernel<STATS><<<grid, PF_THREADS, 0, st>>>(P, V, F, points, verts, faces, order, face_order, dists,
                                                                idxs, stats);
    }
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

__device__ __forceinline__ void atomic_add3(float* dst, V3 g) {
    atomicAdd(dst, g.x); atomicAdd(dst + 1, g.y); atomicAdd(dst + 2, g.z);
}

__device__ __forceinline__ void line3_bwd(V3 p, V3 v0, V3 v1, float g, V3* gp, V3* g0, V3* g1) {
    float tt;
    (void)line3(p, v0, v1, &tt);
    if (tt < 0.f) {
        V3 q = (p - v1) * (2.f * g);
        *gp = *gp + q;
        *g1 = *g1 - q;
        return;
    }
    V3 q = (p - (v0 + (v1 - v0) * tt)) * (2.f * g);
    *gp = *gp + q;
    *g0 = *g0 - q * (1.f - tt);
    *g1 = *g1 - q * tt;
}

__global__ void __launch_bounds__(PF_THREADS)
point_face_bwd_kernel(int P, int V, const float* __restrict__ points, const float* __restrict__ verts,
                      const int* __restrict__ faces, const int* __restrict__ idxs,
                      const float* __restrict__ g_dists, float* __restrict__ g_points,
                      float* __restrict__ g_verts, const int* __restrict__ seg, const int* __restrict__ sub_ptr) {
    const int b = blockIdx.y;
    const int pi = blockIdx.x * PF_THREADS + threadIdx.x;
    if (pi >= P) return;
    const size_t o = (size_t)b * P + pi;
    int f = idxs[o];
    if (seg && f >= 0) f += sub_ptr[seg[o] - 1];          // index local to the point's own face subset
    const float g = g_dists[o];
    V3 z = v3(0.f, 0.f, 0.f), gp = z, g0 = z, g1 = z, g2 = z;
    if (f >= 0 && g != 0.f) {
        const float* vb = verts + (size_t)b * V * 3;
        const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
        const V3 p = v3(points[3 * o], points[3 * o + 1], points[3 * o + 2]);
        const V3 v0 = v3(vb[3 * i0], vb[3 * i0 + 1], vb[3 * i0 + 2]);
        const V3 v1 = v3(vb[3 * i1], vb[3 * i1 + 1], vb[3 * i1 + 2]);
        const V3 v2 = v3(vb[3 * i2], vb[3 * i2 + 1], vb[3 * i2 + 2]);
        int br;
        (void)point_tri(p, v0, v1, v2, &br);
        if (br == 0) {
            V3 e2 = v2 - v0, e1 = v1 - v0;
            V3 n = cross(e2, e1);
            float nn = sqrtf(dot(n, n));
            float s = 1.f / (nn + PF_EPS);
            V3 nh = n * s;
            V3 dv = v0 - p;
            float t = dot(dv, nh);
            float gt = 2.f * t * g;
            gp = nh * (-gt);
            g0 = nh * gt;
            V3 gnh = dv * gt;
            float proj = dot(gnh, n) * s * s / nn;
            V3 gn = gnh * s - n * proj;
            V3 ge2 = cross(e1, gn), ge1 = cross(gn, e2);
            g2 = g2 + ge2;
            g1 = g1 + ge1;
            g0 = g0 - (ge2 + ge1);
        } else if (br == 1) {
            line3_bwd(p, v0, v1, g, &gp, &g0, &g1);
        } else if (br == 2) {
            line3_bwd(p, v0, v2, g, &gp, &g0, &g2);
        } else {
            line3_bwd(p, v1, v2, g, &gp, &g1, &g2);
        }
        float* gv = g_verts + (size_t)b * V * 3;
        atomic_add3(gv + 3 * i0, g0);
        atomic_add3(gv + 3 * i1, g1);
        atomic_add3(gv + 3 * i2, g2);
    }
    if (g_points) { g_points[3 * o] = gp.x; g_points[3 * o + 1] = gp.y; g_points[3 * o + 2] = gp.z; }
}

extern "C" int dsf_point_face_forward(int batch, int P, int V, int F, const float* points, const float* verts,
                                      const int* faces, float* dists, int* idxs, int* order_ws, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(points && verts && faces && dists && idxs, "null argument");
    DSF_REQUIRE(batch > 0 && batch <= 65535 && P > 0 && V > 0 && F > 0, "sizes");
    DSF_REQUIRE(F <= 65535, "at most 65535 faces");
    int* face_order = nullptr;
    if (order_ws) {
        point_sort_kernel<<<batch, PS_THREADS, 0, (cudaStream_t)stream>>>(P, points, order_ws);
        DSF_CHECK_LAUNCH();
        face_order = order_ws + (size_t)batch * P;
        face_sort_kernel<<<batch, PS_THREADS, 0, (cudaStream_t)stream>>>(V, F, verts, faces, face_order);
        DSF_CHECK_LAUNCH();
    }
    return launch_point_face_fwd<false>(batch, P, V, F, points, verts, faces, order_ws, face_order, dists, idxs, nullptr,
                                        (cudaStream_t)stream);
}

// Work counters of the forward scan for the same inputs: stats[0] = pairs rejected by their own bounding-sphere
// test, [1] = pairs evaluated on the interior branch, [2] = on the edge branch, [3] = group-sphere tests (every
// remaining pair was rejected with its group); device, 4 x uint64, overwritten.
// Per pair the scan spends 16 flops on the sphere test (p - v0, offset to the sphere centre, squared norm, reach),
// 40 more for the plane projection + barycentrics + range checks of an interior hit (56), and 64 more when the
// three clamped edge distances are needed instead of t^2 (120); multiply, add and compare each count one.
extern "C" int dsf_point_face_stats(int batch, int P, int V, int F, const float* points, const float* verts,
                                    const int* faces, float* dists, int* idxs, int* order_ws,
                                    unsigned long long* stats, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(points && verts && faces && dists && idxs && stats, "null argument");
    DSF_REQUIRE(batch > 0 && batch <= 65535 && P > 0 && V > 0 && F > 0, "sizes");
    DSF_REQUIRE(F <= 65535, "at most 65535 faces");
    DSF_CHECK_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(unsigned long long), (cudaStream_t)stream));
    int* face_order = nullptr;
    if (order_ws) {
        point_sort_kernel<<<batch, PS_THREADS, 0, (cudaStream_t)stream>>>(P, points, order_ws);
        DSF_CHECK_LAUNCH();
        face_order = order_ws + (size_t)batch * P;
        face_sort_kernel<<<batch, PS_THREADS, 0, (cudaStream_t)stream>>>(V, F, verts, faces, face_order);
        DSF_CHECK_LAUNCH();
    }
    return launch_point_face_fwd<true>(batch, P, V, F, points, verts, faces, order_ws, face_order, dists, idxs, stats,
                                       (cudaStream_t)stream);
}

extern "C" int dsf_point_face_backward(int batch, int P, int V, int F, const float* points, const float* verts,
                                       const int* faces, const int* idxs, const float* g_dists,
                                       float* g_points, float* g_verts, dsfStream_t stream) {
    dsf_reset_launch_count();
    (void)F;
    DSF_REQUIRE(points && verts && faces && idxs && g_dists && g_verts, "null argument");
    DSF_REQUIRE(batch > 0 && batch <= 65535 && P > 0 && V > 0, "sizes");
    DSF_CHECK_CUDA(cudaMemsetAsync(g_verts, 0, (size_t)batch * V * 3 * sizeof(float), (cudaStream_t)stream));
    dim3 grid((P + PF_THREADS - 1) / PF_THREADS, batch);
    point_face_bwd_kernel<<<grid, PF_THREADS, 0, (cudaStream_t)stream>>>(P, V, points, verts, faces, idxs, g_dists,
                                                                       g_points, g_verts, nullptr, nullptr);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}


// ================================================================================================
// "next" row f2: seg_pcl (render_model/mano_layer.py:404-426) and JointICPLoss / FingerICPLoss
// (metric/meshLoss.py:356-394) without replicating the mesh 15 times.
// ================================================================================================
#define SEG_THREADS 256
#define NSPH DSF_NSPHERE

// every point -> 0 (palm) or the finger bone 1..15 whose sphere surface is nearest
__global__ void __launch_bounds__(SEG_THREADS)
seg_pcl_kernel(int P, const float* __restrict__ points, const float* __restrict__ centres,
               const float* __restrict__ radii, int* __restrict__ seg) {
    __shared__ float s_c[NSPH * 3], s_r[NSPH];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < NSPH * 3; i += SEG_THREADS) s_c[i] = centres[(size_t)b * NSPH * 3 + i];
    for (int i = threadIdx.x; i < NSPH; i += SEG_THREADS) s_r[i] = radii[(size_t)b * NSPH + i];
    __syncthreads();
    const int pi = blockIdx.x * SEG_THREADS + threadIdx.x;
    if (pi >= P) return;
    const float* pp = points + ((size_t)b * P + pi) * 3;
    const float x = pp[0], y = pp[1], z = pp[2];
    float best_palm = INFINITY, best_fing = INFINITY;
    int id = 0;
    for (int i = 0; i < NSPH; ++i) {
        const float dx = x - s_c[3 * i], dy = y - s_c[3 * i + 1], dz = z - s_c[3 * i + 2];
        const float d = fabsf(sqrtf(dx * dx + dy * dy + dz * dz + 1e-8f) - s_r[i]);
        if (i < 21) {
            best_palm = fminf(best_palm, d);
        } else if (d < best_fing) {            // first minimum, like torch.min
            best_fing = d;
            id = i - 21;
        }
    }
    seg[(size_t)b * P + pi] = best_palm < best_fing ? 0 : id / 3 + 1;
}

extern "C" int dsf_seg_pcl(int batch, int P, const float* points, const float* centres, const float* radii,
                           int* seg, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(batch > 0 && batch <= 65535 && P > 0 && points && centres && radii && seg, "null / empty argument");
    dim3 grid((P + SEG_THREADS - 1) / SEG_THREADS, batch);
    seg_pcl_kernel<<<grid, SEG_THREADS, 0, (cudaStream_t)stream>>>(P, points, centres, radii, seg);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

// grid (subset k, hand): the points labelled k+1 are compacted into shared memory and scanned
// against subset k's triangles only; every other point keeps distance 0 / index -1.
#define JI_THREADS 256
#define JI_LIST 4096

__global__ void __launch_bounds__(JI_THREADS)
joint_icp_fwd_kernel(int P, int V, const float* __restrict__ points, const float* __restrict__ verts,
                     const int* __restrict__ seg, const int* __restrict__ sub_ptr, const int* __restrict__ sub_faces,
                     float* __restrict__ dists, int* __restrict__ idxs) {
    __shared__ __align__(16) float s_rec[256 * PF_REC];
    __shared__ unsigned short s_list[JI_LIST];
    __shared__ int s_n;
    const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int f_lo = sub_ptr[k], f_hi = sub_ptr[k + 1];
    const int* faces = sub_faces + 3 * (size_t)f_lo;
    const int F = f_hi - f_lo;
    const float* vb = verts + (size_t)b * V * 3;
    const int* sg = seg + (size_t)b * P;
    for (int p0 = 0; p0 < P; p0 += JI_LIST) {
        const int np = min(JI_LIST, P - p0);
        if (tid == 0) s_n = 0;
        __syncthreads();
        for (int q0 = tid & ~31; q0 < np; q0 += JI_THREADS) {          // warp-aggregated compaction
            const int q = q0 + lane;
            const bool hit = q < np && sg[p0 + q] == k + 1;
            const unsigned int m = __ballot_sync(0xffffffffu, hit);
            int base = 0;
            if (lane == 0 && m) base = atomicAdd(&s_n, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit) s_list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)q;
        }
        __syncthreads();
        const int n = s_n;
        for (int e0 = 0; e0 < n; e0 += JI_THREADS) {
            const int e = e0 + tid;
            const int pi = e < n ? p0 + (int)s_list[e] : -1;
            V3 p = v3(0.f, 0.f, 0.f);
            if (pi >= 0) p = v3(points[((size_t)b * P + pi) * 3], points[((size_t)b * P + pi) * 3 + 1],
                                points[((size_t)b * P + pi) * 3 + 2]);
            float best = INFINITY;
            int bi = -1;
            for (int f0 = 0; f0 < F; f0 += 256) {
                const int nf = min(256, F - f0);
                __syncthreads();
                if (tid < nf) build_face_record(vb, faces, f0 + tid, s_rec + tid * PF_REC);
                __syncthreads();
                for (int f = 0; f < nf; ++f) {
                    const float4* r4 = reinterpret_cast<const float4*>(s_rec + f * PF_REC);
                    const float4 q0 = r4[0], q1 = r4[1], q2 = r4[2], q3 = r4[3], q4 = r4[4];
                    const V3 v0 = v3(q0.x, q0.y, q0.z), e1 = v3(q0.w, q1.x, q1.y), e2 = v3(q1.z, q1.w, q2.x);
                    const V3 nh = v3(q2.y, q2.z, q2.w);
                    const V3 a = p - v0;
                    const float t = -dot(a, nh);
                    const V3 c = a + nh * t;
                    const float d20 = dot(c, e1), d21 = dot(c, e2);
                    const float w1 = (q3.z * d20 - q3.y * d21) * q3.w;
                    const float w2 = (q3.x * d21 - q3.y * d20) * q3.w;
                    const float w0 = 1.f - w1 - w2;
                    float d;
                    if (q4.w != 0.f && w0 >= 0.f && w0 <= 1.f && w1 >= 0.f && w1 <= 1.f && w2 >= 0.f && w2 <= 1.f) {
                        d = t * t;
                    } else {
                        d = fminf(fminf(seg_d2(a, e1, q4.x), seg_d2(a, e2, q4.y)), seg_d2(a - e1, e2 - e1, q4.z));
                    }
                    if (d < best) { best = d; bi = f0 + f; }
                }
            }
            if (pi >= 0) {
                dists[(size_t)b * P + pi] = best;
                idxs[(size_t)b * P + pi] = bi;
            }
        }
        __syncthreads();
    }
}

extern "C" int dsf_joint_icp_forward(int batch, int P, int V, int n_subsets, const float* points, const float* verts,
                                     const int* seg, const int* subset_ptr, const int* subset_faces, float* dists,
                                     int* idxs, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(points && verts && seg && subset_ptr && subset_faces && dists && idxs, "null argument");
    DSF_REQUIRE(batch > 0 && batch <= 65535 && P > 0 && V > 0 && n_subsets > 0, "sizes");
    cudaStream_t st = (cudaStream_t)stream;
    DSF_CHECK_CUDA(cudaMemsetAsync(dists, 0, (size_t)batch * P * sizeof(float), st));
    DSF_CHECK_CUDA(cudaMemsetAsync(idxs, 0xFF, (size_t)batch * P * sizeof(int), st));
    dim3 grid(n_subsets, batch);
    joint_icp_fwd_kernel<<<grid, JI_THREADS, 0, st>>>(P, V, points, verts, seg, subset_ptr, subset_faces, dists, idxs);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

extern "C" int dsf_joint_icp_backward(int batch, int P, int V, const float* points, const float* verts,
                                      const int* seg, const int* subset_ptr, const int* subset_faces,
                                      const int* idxs, const float* g_dists, float* g_points, float* g_verts,
                                      dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(points && verts && seg && subset_ptr && subset_faces && idxs && g_dists && g_verts, "null argument");
    DSF_REQUIRE(batch > 0 && batch <= 65535 && P > 0 && V > 0, "sizes");
    DSF_CHECK_CUDA(cudaMemsetAsync(g_verts, 0, (size_t)batch * V * 3 * sizeof(float), (cudaStream_t)stream));
    dim3 grid((P + PF_THREADS - 1) / PF_THREADS, batch);
    point_face_bwd_kernel<<<grid, PF_THREADS, 0, (cudaStream_t)stream>>>(P, V, points, verts, subset_faces, idxs,
                                                                       g_dists, g_points, g_verts, seg, subset_ptr);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}
