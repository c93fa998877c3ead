// dsf_b200 - point-to-mesh-face squared distance for sm_100a (forward arg-min + backward).
// Replaces pytorch3d-0.4.0 _C.point_face_dist_forward/_backward as wrapped by metric/meshLoss.py:21-70
// (ICPLoss :347-353, JointICPLoss :377-394).  DSF always passes one face list for the whole batch, so
// no packing / first_idx tables are needed: grid = (point chunks, hands), the hand's triangles are
// staged in shared memory once per CTA and every thread scans them for its own point.
#include <math.h>

#include "common.cuh"

#define PF_EPS 1e-8f
#define PF_THREADS 128
#define PF_CHUNK 512     // faces staged per pass: 512 * 9 floats = 18 KB

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

__global__ void __launch_bounds__(PF_THREADS)
point_face_fwd_kernel(int P, int V, int F, const float* __restrict__ points, const float* __restrict__ verts,
                      const int* __restrict__ faces, float* __restrict__ dists, int* __restrict__ idxs) {
    __shared__ float s_tri[PF_CHUNK * 9];
    const int b = blockIdx.y;
    const int pi = blockIdx.x * PF_THREADS + threadIdx.x;
    const bool live = pi < P;
    const float* pp = points + ((size_t)b * P + (live ? pi : 0)) * 3;
    const V3 p = v3(pp[0], pp[1], pp[2]);
    const float* vb = verts + (size_t)b * V * 3;
    float best = INFINITY;
    int bi = -1;
    for (int f0 = 0; f0 < F; f0 += PF_CHUNK) {
        const int nf = min(PF_CHUNK, F - f0);
        __syncthreads();
        for (int i = threadIdx.x; i < nf * 3; i += PF_THREADS) {
            const int v = faces[3 * f0 + i];
            s_tri[3 * i] = vb[3 * v]; s_tri[3 * i + 1] = vb[3 * v + 1]; s_tri[3 * i + 2] = vb[3 * v + 2];
        }
        __syncthreads();
        for (int f = 0; f < nf; ++f) {
            const float* t = s_tri + 9 * f;
            int br;
            const float d = point_tri(p, v3(t[0], t[1], t[2]), v3(t[3], t[4], t[5]), v3(t[6], t[7], t[8]), &br);
            if (d < best) { best = d; bi = f0 + f; }      // strict: lowest face index wins ties
        }
    }
    if (live) {
        dists[(size_t)b * P + pi] = best;
        idxs[(size_t)b * P + pi] = bi;
    }
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
                      float* __restrict__ g_verts) {
    const int b = blockIdx.y;
    const int pi = blockIdx.x * PF_THREADS + threadIdx.x;
    if (pi >= P) return;
    const size_t o = (size_t)b * P + pi;
    const int f = idxs[o];
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
                                      const int* faces, float* dists, int* idxs, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(points && verts && faces && dists && idxs, "null argument");
    DSF_REQUIRE(batch > 0 && batch <= 65535 && P > 0 && V > 0 && F > 0, "sizes");
    dim3 grid((P + PF_THREADS - 1) / PF_THREADS, batch);
    point_face_fwd_kernel<<<grid, PF_THREADS, 0, (cudaStream_t)stream>>>(P, V, F, points, verts, faces, dists, idxs);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
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
                                                                       g_points, g_verts);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}
