// dsf_b200 - the synthetic-data generator extras of the renderer, sm_100a:
//   mask_img    random spherical occluders, render_model/mano_layer.py:1326-1340
//   synth2real  patch noise + 5x5 Gaussian smoothing, render_model/mano_layer.py:1222-1231 (+ GaussianSmoothing :808-868)
//   chamfer     the nearest-neighbour core of pytorch3d.loss.chamfer_distance as surface_loss uses it,
//               render_model/render_loss.py:37-52
// All three are HBM-bound elementwise / small-neighbourhood kernels on the rendered crops; the random numbers
// are drawn by the host wrappers with the reference's own generator calls and passed in.
#include <math.h>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// mask_img: pixel (row, col) of the R x R normalised depth crop is the point (x, y, d) with
// x = 2 (col + 0.5) / R - 1, y = 2 (row + 0.5) / R - 1 (Render.xy_mesh, :968-972); it becomes background 1.0
// when it lies inside any of the hand's n_mask spheres.  Same float32 operation order as the reference
// (squares summed left to right, sqrt, strict <).
// ------------------------------------------------------------------------------------------------
#define MK_THREADS 256
#define MK_MAX 32

__global__ void __launch_bounds__(MK_THREADS)
mask_img_kernel(int R, int n_mask, const float* __restrict__ img, const float* __restrict__ centres,
                const float* __restrict__ radii, float* __restrict__ out) {
    __shared__ float sc[MK_MAX][4];
    const int b = blockIdx.y;
    if (threadIdx.x < n_mask) {
        sc[threadIdx.x][0] = centres[((size_t)b * n_mask + threadIdx.x) * 3];
        sc[threadIdx.x][1] = centres[((size_t)b * n_mask + threadIdx.x) * 3 + 1];
        sc[threadIdx.x][2] = centres[((size_t)b * n_mask + threadIdx.x) * 3 + 2];
        sc[threadIdx.x][3] = radii[(size_t)b * n_mask + threadIdx.x];
    }
    __syncthreads();
    const float Rf = (float)R;
    for (int k = blockIdx.x * MK_THREADS + threadIdx.x; k < R * R; k += gridDim.x * MK_THREADS) {
        const int row = k / R, col = k % R;
        // numpy float64 arithmetic rounded to float32 once (:968-972)
        const float x = (float)(2.0 * ((double)col + 0.5) / (double)Rf - 1.0);
        const float y = (float)(2.0 * ((double)row + 0.5) / (double)Rf - 1.0);
        const float d = img[(size_t)b * R * R + k];
        bool hit = false;
        for (int m = 0; m < n_mask; ++m) {
            const float dx = __fsub_rn(x, sc[m][0]), dy = __fsub_rn(y, sc[m][1]), dz = __fsub_rn(d, sc[m][2]);
            const float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            hit = hit || (__fsqrt_rn(s) < sc[m][3]);
        }
        out[(size_t)b * R * R + k] = hit ? 1.f : d;
    }
}

extern "C" int dsf_mask_img(int batch, int R, const float* img, int n_mask, const float* centres,
                            const float* radii, float* out, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(batch > 0 && R > 0 && img && out, "null / empty argument");
    DSF_REQUIRE(n_mask >= 0 && n_mask <= MK_MAX && (n_mask == 0 || (centres && radii)), "n_mask must be in [0,32]");
    const int bx = (R * R + MK_THREADS * 4 - 1) / (MK_THREADS * 4);
    mask_img_kernel<<<dim3(bx, batch), MK_THREADS, 0, (cudaStream_t)stream>>>(R, n_mask, img, centres, radii, out);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

// ------------------------------------------------------------------------------------------------
// synth2real: img + nearest-upsampled patch noise on the foreground (img < bk_value), then - sigma != 0 -
// reflect-pad by 2 and a normalised 5x5 Gaussian (GaussianSmoothing.forward with kernel_size 5).  One pass:
// every output pixel gathers its 5x5 neighbourhood of the noised image with reflected indices.
// ------------------------------------------------------------------------------------------------
struct Gauss5 { float w[25]; };

__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__global__ void __launch_bounds__(MK_THREADS)
synth2real_kernel(int R, const float* __restrict__ img, const float* __restrict__ noise, int patch, float bk_value,
                  int smooth, Gauss5 g, float* __restrict__ out) {
    const int b = blockIdx.y;
    const float* im = img + (size_t)b * R * R;
    const int Rn = R / patch;
    const float* nz = noise ? noise + (size_t)b * Rn * Rn : nullptr;
    auto noised = [&](int r, int c) {
        const float v = im[r * R + c];
        if (!nz) return v;
        // upsample_nearest of the (R/patch)^2 noise; rows / columns beyond patch * (R / patch) do not exist in
        // the reference (it requires R % patch == 0)
        const float n = nz[min(r / patch, Rn - 1) * Rn + min(c / patch, Rn - 1)];
        return __fadd_rn(v, __fmul_rn(n, v < bk_value ? 1.f : 0.f));
    };
    for (int k = blockIdx.x * MK_THREADS + threadIdx.x; k < R * R; k += gridDim.x * MK_THREADS) {
        const int row = k / R, col = k % R;
        float acc;
        if (!smooth) {
            acc = noised(row, col);
        } else {
            acc = 0.f;
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int j = 0; j < 5; ++j)
                    acc = fmaf(g.w[i * 5 + j], noised(reflect_idx(row + i - 2, R), reflect_idx(col + j - 2, R)), acc);
        }
        out[(size_t)b * R * R + k] = acc;
    }
}

extern "C" int dsf_synth2real(int batch, int R, const float* img, const float* noise, int patch, float bk_value,
                              float sigma, float* out, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(batch > 0 && R >= 3 && img && out && img != out, "null / aliased argument");
    DSF_REQUIRE(!noise || (patch > 0 && R % patch == 0), "noise needs a patch size dividing R");
    Gauss5 g;
    if (sigma != 0.f) {
        // GaussianSmoothing.forward (:847-868): product of the two 1-D Gaussians in float32, normalised by its sum
        float k1[5], s = 0.f;
        for (int i = 0; i < 5; ++i) {
            const float t = ((float)i - 2.f) / sigma;
            k1[i] = 1.f / (sigma * sqrtf(2.f * 3.14159265358979323846f)) * expf(-(t * t) / 2.f);
        }
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j) { g.w[i * 5 + j] = k1[i] * k1[j]; s += g.w[i * 5 + j]; }
        for (int i = 0; i < 25; ++i) g.w[i] /= s;
    } else {
        for (int i = 0; i < 25; ++i) g.w[i] = 0.f;
    }
    const int bx = (R * R + MK_THREADS * 4 - 1) / (MK_THREADS * 4);
    synth2real_kernel<<<dim3(bx, batch), MK_THREADS, 0, (cudaStream_t)stream>>>(R, img, noise, patch, bk_value,
                                                                               sigma != 0.f, g, out);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

// ------------------------------------------------------------------------------------------------
// chamfer: for every point of x (B,P1,3) the squared distance to and index of its nearest point of y (B,P2,3),
// and vice versa (pytorch3d knn_points with K = 1, which chamfer_distance is built on).  One CTA per
// (hand, 256-point slab); the other cloud is staged through shared memory in chunks.
// Backward: d sum(w_x dist_x) + sum(w_y dist_y): both clouds receive +-2 w (x_i - y_j) of every matched pair.
// ------------------------------------------------------------------------------------------------
#define CH_THREADS 256
#define CH_CHUNK 1024

__global__ void __launch_bounds__(CH_THREADS)
chamfer_nn_kernel(int P1, int P2, const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ dist,
                  int* __restrict__ idx) {
    __shared__ float sy[CH_CHUNK * 3];
    const int b = blockIdx.y;
    const int i = blockIdx.x * CH_THREADS + threadIdx.x;
    const bool live = i < P1;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (live) {
        px = x[((size_t)b * P1 + i) * 3]; py = x[((size_t)b * P1 + i) * 3 + 1]; pz = x[((size_t)b * P1 + i) * 3 + 2];
    }
    float best = INFINITY;
    int bi = 0;
    for (int c0 = 0; c0 < P2; c0 += CH_CHUNK) {
        const int n = min(CH_CHUNK, P2 - c0);
        __syncthreads();
        for (int k = threadIdx.x; k < n * 3; k += CH_THREADS) sy[k] = y[((size_t)b * P2 + c0) * 3 + k];
        __syncthreads();
        for (int j = 0; j < n; ++j) {
            const float dx = px - sy[3 * j], dy = py - sy[3 * j + 1], dz = pz - sy[3 * j + 2];
            const float d = dx * dx + dy * dy + dz * dz;
            if (d < best) { best = d; bi = c0 + j; }       // strict: lowest index wins ties
        }
    }
    if (live) { dist[(size_t)b * P1 + i] = best; idx[(size_t)b * P1 + i] = bi; }
}

extern "C" int dsf_chamfer_forward(int batch, int P1, int P2, const float* x, const float* y, float* dist_x,
                                   int* idx_x, float* dist_y, int* idx_y, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(batch > 0 && P1 > 0 && P2 > 0 && x && y && dist_x && idx_x && dist_y && idx_y, "null / empty argument");
    DSF_REQUIRE(batch <= 65535, "batch must be <= 65535 per call");
    cudaStream_t st = (cudaStream_t)stream;
    chamfer_nn_kernel<<<dim3((P1 + CH_THREADS - 1) / CH_THREADS, batch), CH_THREADS, 0, st>>>(P1, P2, x, y, dist_x, idx_x);
    DSF_CHECK_LAUNCH();
    chamfer_nn_kernel<<<dim3((P2 + CH_THREADS - 1) / CH_THREADS, batch), CH_THREADS, 0, st>>>(P2, P1, y, x, dist_y, idx_y);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

// g_a[i] += 2 w[i] (a_i - b_idx[i]);  g_b[idx[i]] -= the same
__global__ void __launch_bounds__(CH_THREADS)
chamfer_bwd_kernel(int P1, int P2, const float* __restrict__ a, const float* __restrict__ bq, const int* __restrict__ idx,
                   const float* __restrict__ w, float* __restrict__ g_a, float* __restrict__ g_b) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * CH_THREADS + threadIdx.x;
    if (i >= P1) return;
    const int j = idx[(size_t)b * P1 + i];
    const float s = 2.f * w[(size_t)b * P1 + i];
    const float* pa = a + ((size_t)b * P1 + i) * 3;
    const float* pb = bq + ((size_t)b * P2 + j) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float g = s * (pa[c] - pb[c]);
        atomicAdd(g_a + ((size_t)b * P1 + i) * 3 + c, g);
        atomicAdd(g_b + ((size_t)b * P2 + j) * 3 + c, -g);
    }
}

extern "C" int dsf_chamfer_backward(int batch, int P1, int P2, const float* x, const float* y, const int* idx_x,
                                    const int* idx_y, const float* g_dist_x, const float* g_dist_y, float* g_x,
                                    float* g_y, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(batch > 0 && P1 > 0 && P2 > 0 && x && y && idx_x && idx_y && g_dist_x && g_dist_y && g_x && g_y,
                "null / empty argument");
    cudaStream_t st = (cudaStream_t)stream;
    DSF_CHECK_CUDA(cudaMemsetAsync(g_x, 0, (size_t)batch * P1 * 3 * sizeof(float), st));
    DSF_CHECK_CUDA(cudaMemsetAsync(g_y, 0, (size_t)batch * P2 * 3 * sizeof(float), st));
    chamfer_bwd_kernel<<<dim3((P1 + CH_THREADS - 1) / CH_THREADS, batch), CH_THREADS, 0, st>>>(P1, P2, x, y, idx_x, g_dist_x, g_x, g_y);
    DSF_CHECK_LAUNCH();
    chamfer_bwd_kernel<<<dim3((P2 + CH_THREADS - 1) / CH_THREADS, batch), CH_THREADS, 0, st>>>(P2, P1, y, x, idx_y, g_dist_y, g_y, g_x);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}
