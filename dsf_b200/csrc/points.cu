// dsf_b200 - rigid rotation of point sets about a per-hand centre (extra camera views).
// Replaces RotationPoints / RotationNormalPoints, render_model/mano_layer.py:874-895, whose
// torch.matmul over (B, N, 3, 1) operands becomes B*N tiny cuBLAS GEMVs (46 % of the multi-view
// config's step time when left in torch).  out = R (p - c) + c, R (B,3,3) row-major; backward gives the
// cotangents of the points and, optionally, of R and c (the caller chains R back through Rodrigues).
#include "common.cuh"

#define RP_THREADS 256

__global__ void __launch_bounds__(RP_THREADS)
rotate_points_kernel(int n, const float* __restrict__ pts, const float* __restrict__ Rm,
                     const float* __restrict__ center, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * RP_THREADS + threadIdx.x;
    if (i >= n) return;
    const float* R = Rm + 9 * (size_t)b;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (center) { cx = center[3 * b]; cy = center[3 * b + 1]; cz = center[3 * b + 2]; }
    const float* p = pts + ((size_t)b * n + i) * 3;
    const float x = p[0] - cx, y = p[1] - cy, z = p[2] - cz;
    float* o = out + ((size_t)b * n + i) * 3;
    o[0] = (R[0] * x + R[1] * y + R[2] * z) + cx;
    o[1] = (R[3] * x + R[4] * y + R[5] * z) + cy;
    o[2] = (R[6] * x + R[7] * y + R[8] * z) + cz;
}

// one CTA per hand: g_p = R^T g ; g_R = sum_n g_n (p_n - c)^T ; g_c = sum_n (g_n - R^T g_n)
__global__ void __launch_bounds__(RP_THREADS)
rotate_points_bwd_kernel(int n, const float* __restrict__ pts, const float* __restrict__ Rm,
                         const float* __restrict__ center, const float* __restrict__ g_out,
                         float* __restrict__ g_pts, float* __restrict__ g_R, float* __restrict__ g_center) {
    __shared__ float red[RP_THREADS / 32][12];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* R = Rm + 9 * (size_t)b;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (center) { cx = center[3 * b]; cy = center[3 * b + 1]; cz = center[3 * b + 2]; }
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.f;
    for (int i = tid; i < n; i += RP_THREADS) {
        const float* g = g_out + ((size_t)b * n + i) * 3;
        const float gx = g[0], gy = g[1], gz = g[2];
        const float px = R[0] * gx + R[3] * gy + R[6] * gz;
        const float py = R[1] * gx + R[4] * gy + R[7] * gz;
        const float pz = R[2] * gx + R[5] * gy + R[8] * gz;
        if (g_pts) {
            float* o = g_pts + ((size_t)b * n + i) * 3;
            o[0] = px; o[1] = py; o[2] = pz;
        }
        if (g_R || g_center) {
            const float* p = pts + ((size_t)b * n + i) * 3;
            const float x = p[0] - cx, y = p[1] - cy, z = p[2] - cz;
            acc[0] += gx * x; acc[1] += gx * y; acc[2] += gx * z;
            acc[3] += gy * x; acc[4] += gy * y; acc[5] += gy * z;
            acc[6] += gz * x; acc[7] += gz * y; acc[8] += gz * z;
            acc[9] += gx - px; acc[10] += gy - py; acc[11] += gz - pz;
        }
    }
    if (!g_R && !g_center) return;
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = warp_sum(acc[k]);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 12; ++k) red[warp][k] = acc[k];
    __syncthreads();
    if (tid < 12) {
        float t = 0.f;
        for (int w = 0; w < RP_THREADS / 32; ++w) t += red[w][tid];      // fixed order
        if (tid < 9) { if (g_R) g_R[9 * (size_t)b + tid] = t; }
        else if (g_center) g_center[3 * (size_t)b + tid - 9] = t;
    }
}

extern "C" int dsf_rotate_points(int batch, int n, const float* pts, const float* Rm, const float* center3d,
                                 float* out, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(batch > 0 && n > 0 && pts && Rm && out, "null / empty argument");
    rotate_points_kernel<<<dim3((n + RP_THREADS - 1) / RP_THREADS, batch), RP_THREADS, 0, (cudaStream_t)stream>>>(
        n, pts, Rm, center3d, out);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

extern "C" int dsf_rotate_points_backward(int batch, int n, const float* pts, const float* Rm, const float* center3d,
                                          const float* g_out, float* g_pts, float* g_R, float* g_center,
                                          dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(batch > 0 && n > 0 && pts && Rm && g_out && (g_pts || g_R || g_center), "null / empty argument");
    DSF_REQUIRE(!g_center || center3d, "g_center needs center3d");
    rotate_points_bwd_kernel<<<batch, RP_THREADS, 0, (cudaStream_t)stream>>>(n, pts, Rm, center3d, g_out, g_pts, g_R,
                                                                             g_center);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}
