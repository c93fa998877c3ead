// dsf_b200 - depth rasteriser for sm_100a: per-hand view set-up, forward (z-buffer by packed
// 64-bit depth|face atomicMin in shared memory, fused background fill + depth normalisation),
// backward (zbuf gradient -> camera-space vertices) and the render losses.
// Replaces pytorch3d-0.4.0 rasterize_meshes{,_backward} as used at render_model/mano_layer.py:1083
// plus Render.resize/comToBounds/Offset2Trans/warpPerspective/normalize_img (:1133-1299) and
// render_loss.py:15-21 / train_render.py:728-732.  See include/dsf_b200.h.
//
// Arithmetic that decides pix_to_face / zbuf is written with explicit round-to-nearest
// intrinsics (__fmul_rn / __fadd_rn / __fsub_rn / __fdiv_rn are never contracted into FMAs) in
// exactly the operation order of the CPU oracle (oracle/raster_oracle_impl.h), so that the two
// agree bit for bit on identical inputs.
#include <math.h>

#include "raster.cuh"

#define EPS 1e-8f

__device__ __forceinline__ float edge_rn(float px, float py, float ax, float ay, float bx, float by) {
    return __fsub_rn(__fmul_rn(__fsub_rn(px, ax), __fsub_rn(by, ay)),
                     __fmul_rn(__fsub_rn(py, ay), __fsub_rn(bx, ax)));
}

__device__ __forceinline__ float pix_to_ndc(int i, int n) {
    // pytorch3d PixToNdc of the flipped index n-1-i
    return __fadd_rn(-1.0f, __fdiv_rn(__fadd_rn(__fmul_rn(2.0f, (float)(n - 1 - i)), 1.0f), (float)n));
}

// ------------------------------------------------------------------------------------------------
// view set-up: one CTA per hand
// ------------------------------------------------------------------------------------------------
__global__ void view_setup_kernel(int mode, int B, const float* __restrict__ center3d,
                                  const float* __restrict__ cube, float fx, float fy, float px, float py,
                                  int W, int H, int R, const float* __restrict__ M_in, float* __restrict__ view,
                                  float* __restrict__ xs, float* __restrict__ ys, float* __restrict__ M_out) {
    __shared__ float sM[4];   // s_x, t_x, s_y, t_y
    __shared__ int s_lo[2], s_hi[2];
    const int b = blockIdx.x;
    const float cx = center3d[3 * b], cy = center3d[3 * b + 1], cz = center3d[3 * b + 2];
    const float sx = cube[3 * b], sy = cube[3 * b + 1], sz = cube[3 * b + 2];
    if (threadIdx.x == 0) {
        float sc, tx, ty;
        if (M_in) {
            sc = M_in[9 * b]; tx = M_in[9 * b + 2]; ty = M_in[9 * b + 5];
            sM[0] = sc; sM[1] = tx; sM[2] = M_in[9 * b + 4]; sM[3] = ty;
        } else {
            // points3DToImg (mano_layer.py:1318-1324)
            float u = __fadd_rn(__fdiv_rn(__fmul_rn(cx, fx), __fadd_rn(cz, 1e-8f)), px);
            float v = __fadd_rn(__fdiv_rn(__fmul_rn(cy, fy), cz), py);
            // comToBounds (:1133-1141)
            float ax = __fdiv_rn(__fmul_rn(u, cz), fx), ay = __fdiv_rn(__fmul_rn(v, cz), fy);
            float hx = __fdiv_rn(sx, 2.f), hy = __fdiv_rn(sy, 2.f);
            int x0 = (int)floorf(__fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(ax, hx), cz), fx), 0.5f));
            int x1 = (int)floorf(__fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(ax, hx), cz), fx), 0.5f));
            int y0 = (int)floorf(__fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(ay, hy), cz), fy), 0.5f));
            int y1 = (int)floorf(__fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(ay, hy), cz), fy), 0.5f));
            // Offset2Trans (:1143-1169)
            int wb = x1 - x0, hb = y1 - y0;
            float Rf = (float)R;
            int sz0, sz1;
            if (wb > hb) {
                sc = __fdiv_rn(Rf, (float)wb);
                sz0 = R;
                sz1 = (int)__fdiv_rn((float)(hb * R), (float)wb);
            } else {
                sc = __fdiv_rn(Rf, (float)hb);
                sz0 = (int)__fdiv_rn((float)(wb * R), (float)hb);
                sz1 = R;
            }
            float ox = floorf(__fsub_rn(__fdiv_rn(Rf, 2.f), __fdiv_rn((float)sz0, 2.f)));
            float oy = floorf(__fsub_rn(__fdiv_rn(Rf, 2.f), __fdiv_rn((float)sz1, 2.f)));
            tx = __fsub_rn(ox, __fmul_rn(sc, (float)x0));
            ty = __fsub_rn(oy, __fmul_rn(sc, (float)y0));
            sM[0] = sc; sM[1] = tx; sM[2] = sc; sM[3] = ty;
        }
        if (M_out) {
            float* m = M_out + 9 * b;
            m[0] = sM[0]; m[1] = 0.f; m[2] = sM[1];
            m[3] = 0.f; m[4] = sM[2]; m[5] = sM[3];
            m[6] = 0.f; m[7] = 0.f; m[8] = 1.f;
        }
        float* vw = view + (size_t)b * VIEW;
        float zh = __fdiv_rn(sz, 2.f);
        vw[4] = cz;
        vw[5] = zh;
        vw[6] = __fdiv_rn(__fsub_rn(__fadd_rn(cz, zh), cz), zh);   // background after normalize_img
        vw[15] = 0.f;
        vw[16] = (float)(mode != 1 ? R : (W > H ? W : H));   // side of the raster the samples index into
        vw[17] = 0.f; vw[18] = 0.f; vw[19] = 0.f;
        if (mode != 1) {
            float half = (float)R * 0.5f;
            float fxc = __fmul_rn(sM[0], fx), fyc = __fmul_rn(sM[2], fy);
            float pxc = __fadd_rn(__fmul_rn(sM[0], px), sM[1]), pyc = __fadd_rn(__fmul_rn(sM[2], py), sM[3]);
            // mode 2: sample i sits at crop coordinate i (the convention of M, JointTrans and the literal chain,
            // whose crop index c reads sensor coordinate (c - t) / s) instead of the pixel centre i + 1/2
            if (mode == 2) { pxc = __fadd_rn(pxc, 0.5f); pyc = __fadd_rn(pyc, 0.5f); }
            vw[0] = __fdiv_rn(fxc, half);
            vw[1] = __fdiv_rn(fyc, half);
            vw[2] = -__fdiv_rn(__fsub_rn(pxc, half), half);
            vw[3] = -__fdiv_rn(__fsub_rn(pyc, half), half);
            vw[7] = -half; vw[8] = ((float)R - 1.f) * 0.5f;
            vw[9] = -half; vw[10] = ((float)R - 1.f) * 0.5f;
            vw[11] = 0.f; vw[12] = (float)(R - 1); vw[13] = 0.f; vw[14] = (float)(R - 1);
            vw[15] = 1.f;      // the sample grid is affine: index = a * ndc + b exactly (a, b above)
        } else {
            float hw = (float)W * 0.5f, hh = (float)H * 0.5f;
            vw[0] = __fdiv_rn(fx, hw);
            vw[1] = __fdiv_rn(fy, hh);
            vw[2] = -__fdiv_rn(__fsub_rn(px, hw), hw);
            vw[3] = -__fdiv_rn(__fsub_rn(py, hh), hh);
            vw[7] = -sM[0] * hw; vw[8] = sM[0] * hw + sM[1];
            vw[9] = -sM[2] * hh; vw[10] = sM[2] * hh + sM[3];
        }
        s_lo[0] = s_lo[1] = R; s_hi[0] = s_hi[1] = -1;
    }
    __syncthreads();
    if (mode != 1) {
        for (int i = threadIdx.x; i < R; i += blockDim.x) {
            float v = pix_to_ndc(i, R);
            xs[(size_t)b * R + i] = v;
            ys[(size_t)b * R + i] = v;
        }
        return;
    }
    // literal chain: crop pixel c -> sensor pixel r (warpPerspective, :1244-1260) -> raster pixel q
    // (resize, :1233-1242), both nearest-neighbour grid_samples; out of range reads zero padding.
    const int S = W > H ? W : H;
    for (int i = threadIdx.x; i < 2 * R; i += blockDim.x) {
        int axis = i / R, c = i % R;
        float sc = sM[2 * axis], t = sM[2 * axis + 1];
        int n = axis == 0 ? W : H;
        float u = __fsub_rn(__fdiv_rn(__fsub_rn((float)c, t), sc), 0.5f);
        float r = rintf(u);
        float val = nanf("");
        if (r >= 0.f && r <= (float)(n - 1)) {
            float src = __fsub_rn(__fdiv_rn(__fmul_rn(__fadd_rn(__fmul_rn(2.f, r), 1.f), (float)S), (float)(2 * n)), 0.5f);
            int q = (int)rintf(src);
            q = q < 0 ? 0 : (q > S - 1 ? S - 1 : q);
            val = pix_to_ndc(q, S);
            atomicMin(&s_lo[axis], c);
            atomicMax(&s_hi[axis], c);
        }
        (axis == 0 ? xs : ys)[(size_t)b * R + c] = val;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float* vw = view + (size_t)b * VIEW;
        vw[11] = (float)s_lo[0]; vw[12] = (float)s_hi[0];
        vw[13] = (float)s_lo[1]; vw[14] = (float)s_hi[1];
    }
}

extern "C" int dsf_view_setup(int mode, int batch, const float* center3d, const float* cube,
                              const float* intr4, int W, int H, int R, const float* M_in, float* view,
                              float* xs, float* ys, float* M_out, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (direct), 1 (literal) or 2 (direct, index-aligned)");
    DSF_REQUIRE(batch > 0 && center3d && cube && intr4 && view && xs && ys, "null argument");
    DSF_REQUIRE(R >= 8 && R <= 512, "crop size R must be in [8,512]");
    DSF_REQUIRE(W > 0 && H > 0, "sensor size");
    view_setup_kernel<<<batch, 128, 0, (cudaStream_t)stream>>>(mode, batch, center3d, cube, intr4[0], intr4[1],
                                                              intr4[2], intr4[3], W, H, R, M_in, view, xs,
                                                              ys, M_out);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

// ------------------------------------------------------------------------------------------------
// shared pieces of forward / backward
// ------------------------------------------------------------------------------------------------
struct ViewRec {
    float fxn, fyn, pxn, pyn, zc, zh, bg, ax, bx, ay, by;
    int xlo, xhi, ylo, yhi;
    bool affine;
    float S;
};

__device__ __forceinline__ ViewRec load_view(const float* v) {
    ViewRec r;
    r.fxn = v[0]; r.fyn = v[1]; r.pxn = v[2]; r.pyn = v[3]; r.zc = v[4]; r.zh = v[5]; r.bg = v[6];
    r.ax = v[7]; r.bx = v[8]; r.ay = v[9]; r.by = v[10];
    r.xlo = (int)v[11]; r.xhi = (int)v[12]; r.ylo = (int)v[13]; r.yhi = (int)v[14];
    r.affine = v[15] != 0.f;
    r.S = v[16];
    return r;
}

// camera-space vertex -> (x_ndc, y_ndc, z): world->view flip R = diag(-1,-1,1) (mano_layer.py:935-938)
// then the pytorch3d screen-space calibration.  place != null applies verts*cube/2 + center first
// (mano_layer.py:1078).
__device__ __forceinline__ void project_vertex(const float* v, const float* place_scale, const float* place_off,
                                               const ViewRec& vw, float* out) {
    float x = v[0], y = v[1], z = v[2];
    if (place_scale) {
        // (v * cube) / 2 + centre; the division by two is exact, written as a multiplication
        x = __fadd_rn(__fmul_rn(__fmul_rn(x, place_scale[0]), 0.5f), place_off[0]);
        y = __fadd_rn(__fmul_rn(__fmul_rn(y, place_scale[1]), 0.5f), place_off[1]);
        z = __fadd_rn(__fmul_rn(__fmul_rn(z, place_scale[2]), 0.5f), place_off[2]);
    }
    out[0] = __fdiv_rn(__fadd_rn(__fmul_rn(vw.fxn, -x), __fmul_rn(vw.pxn, z)), z);
    out[1] = __fdiv_rn(__fadd_rn(__fmul_rn(vw.fyn, -y), __fmul_rn(vw.pyn, z)), z);
    out[2] = z;
}

// first index i in [lo,hi] with s[i] <= hiv (s non-increasing); hi+1 if none
__device__ __forceinline__ int first_le(const float* s, int lo, int hi, float a, float b, float hiv) {
    int i = (int)ceilf(fmaf(a, hiv, b));
    i = max(lo, min(hi + 1, i));
    while (i > lo && s[i - 1] <= hiv) --i;
    while (i <= hi && !(s[i] <= hiv)) ++i;
    return i;
}
// last index i in [lo,hi] with s[i] >= lov; lo-1 if none
__device__ __forceinline__ int last_ge(const float* s, int lo, int hi, float a, float b, float lov) {
    int i = (int)floorf(fmaf(a, lov, b));
    i = min(hi, max(lo - 1, i));
    while (i < hi && s[i + 1] >= lov) ++i;
    while (i >= lo && !(s[i] >= lov)) --i;
    return i;
}

struct FragEval {
    float w0, w1, w2, b0, b1, b2, pz;
    bool ok;
};

// full oracle-order evaluation of one (pixel, face) pair.  PERSP = pytorch3d's perspective_correct:
// false (the 0.4.0 RasterizationSettings default the reference runs with, mano_layer.py:946-950 passes no
// such argument) interpolates z with the screen-space barycentrics, true applies
// BarycentricPerspectiveCorrection first.
template <bool PERSP>
__device__ __forceinline__ FragEval eval_fragment_e(float e0, float e1, float e2, float z0, float z1, float z2,
                                                    float area) {
    FragEval r;
    r.w0 = __fdiv_rn(e0, area);
    r.w1 = __fdiv_rn(e1, area);
    r.w2 = __fdiv_rn(e2, area);
    if (PERSP) {
        float t0 = __fmul_rn(__fmul_rn(r.w0, z1), z2);
        float t1 = __fmul_rn(__fmul_rn(z0, r.w1), z2);
        float t2 = __fmul_rn(__fmul_rn(z0, z1), r.w2);
        float den = __fadd_rn(__fadd_rn(t0, t1), t2);
        r.b0 = __fdiv_rn(t0, den);
        r.b1 = __fdiv_rn(t1, den);
        r.b2 = __fdiv_rn(t2, den);
    } else {
        r.b0 = r.w0; r.b1 = r.w1; r.b2 = r.w2;
    }
    r.pz = __fadd_rn(__fadd_rn(__fmul_rn(r.b0, z0), __fmul_rn(r.b1, z1)), __fmul_rn(r.b2, z2));
    r.ok = !(r.pz < 0.f) && r.b0 > 0.f && r.b1 > 0.f && r.b2 > 0.f;
    return r;
}

template <bool PERSP>
__device__ __forceinline__ FragEval eval_fragment(float px, float py, float x0, float y0, float z0, float x1,
                                                  float y1, float z1, float x2, float y2, float z2, float area) {
    return eval_fragment_e<PERSP>(edge_rn(px, py, x1, y1, x2, y2), edge_rn(px, py, x2, y2, x0, y0),
                                  edge_rn(px, py, x0, y0, x1, y1), z0, z1, z2, area);
}

// Gradient of L = sum_p g(p) pz(p) over the pixels p of ONE face when depth is interpolated with the
// screen-space barycentrics (no perspective correction): pz(p) = [z0 e0 + z1 e1 + z2 e2] / (A + eps) is affine
// in p, so L depends on the pixels only through S0 = sum g and U = sum g (p - v0):
//   L = k (z0 A S0 + dz1 E1 + dz2 E2),  k = 1 / (A + eps),  A = b x a (the oracle's area without eps),
//   a = v1 - v0, b = v2 - v0, dz_i = z_i - z0, E1 = Uy bx - Ux by = sum g e1(p), E2 = Ux ay - Uy ax = sum g e2(p).
// The form is chosen for conditioning: the per-pixel chain rule of the reference (g_w_i = g z_i, then through
// w_i = e_i / area) subtracts two O(z S0) terms that agree to ~1e-4 of their size; here
// d L / d A = k^2 (z0 S0 eps - D) (using 1 - A k = eps k) and they never meet.  Verified against float64 autograd.
// out: g[0..2] = d L / d (x0, y0, z0), g[3..5] vertex 1, g[6..8] vertex 2.
template <typename T>
__device__ __forceinline__ void face_grad(T z0, T ax, T ay, T dz1, T bx, T by, T dz2, T S0, T Ux, T Uy, T* g) {
    const T A = bx * ay - by * ax;
    const T k = T(1) / (A + T(1e-8));
    const T E1 = Uy * bx - Ux * by, E2 = Ux * ay - Uy * ax;
    const T gA = k * k * (z0 * S0 * T(1e-8) - (dz1 * E1 + dz2 * E2));
    const T g_ax = -gA * by - k * dz2 * Uy, g_ay = gA * bx + k * dz2 * Ux;
    const T g_bx = gA * ay + k * dz1 * Uy, g_by = -gA * ax - k * dz1 * Ux;
    const T g_Ux = k * (dz2 * ay - dz1 * by), g_Uy = k * (dz1 * bx - dz2 * ax);
    g[3] = g_ax; g[4] = g_ay; g[5] = k * E1;
    g[6] = g_bx; g[7] = g_by; g[8] = k * E2;
    g[0] = -g_ax - g_bx - S0 * g_Ux;
    g[1] = -g_ay - g_by - S0 * g_Uy;
    g[2] = k * A * S0 - g[5] - g[8];
}

__device__ __forceinline__ float seg_dist_rn(float px, float py, float ax, float ay, float bx, float by) {
    float bax = __fsub_rn(bx, ax), bay = __fsub_rn(by, ay);
    float l2 = __fadd_rn(__fmul_rn(bax, bax), __fmul_rn(bay, bay));
    if (l2 <= EPS) {
        float dx = __fsub_rn(px, bx), dy = __fsub_rn(py, by);
        return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    }
    float t = __fdiv_rn(__fadd_rn(__fmul_rn(bax, __fsub_rn(px, ax)), __fmul_rn(bay, __fsub_rn(py, ay))), l2);
    t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
    float qx = __fsub_rn(__fadd_rn(ax, __fmul_rn(t, bax)), px);
    float qy = __fsub_rn(__fadd_rn(ay, __fmul_rn(t, bay)), py);
    return __fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy));
}

// ------------------------------------------------------------------------------------------------
// crop_hand ("next" row f1): keep only pixels whose back-projected 3-D point lies inside the box
// around the teacher skeleton - data/render_loader.py:1209-1227 with uvdImg2xyzImg (:1190-1200),
// uvd_nl2xyz_tensor (:1044-1057) and pointsImgTo3D (:336-343, flip = 1).
// ------------------------------------------------------------------------------------------------
struct CropBox {
    float lo[3], hi[3];
    float s_x, t_x, s_y, t_y;
};

// one warp: box = skeleton bounds +- offsets (skeleton = joint * cube / 2 + centre)
__device__ __forceinline__ void crop_box_warp(const CropParams& cp, int mesh, const float* center, const float* cube,
                                              int lane, CropBox* out) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int j = lane; j < cp.nj; j += 32) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __fadd_rn(__fdiv_rn(__fmul_rn(cp.joints[((size_t)mesh * cp.nj + j) * 3 + c], cube[3 * mesh + c]), 2.f),
                                      center[3 * mesh + c]);
            lo[c] = fminf(lo[c], v);
            hi[c] = fmaxf(hi[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    if (lane == 0) {
        out->lo[0] = lo[0] - cp.off_xy; out->hi[0] = hi[0] + cp.off_xy;
        out->lo[1] = lo[1] - cp.off_xy; out->hi[1] = hi[1] + cp.off_xy;
        out->lo[2] = lo[2] - cp.off_z - cp.thick; out->hi[2] = hi[2] + cp.off_z;
        const float* m = cp.M + 9 * (size_t)mesh;
        out->s_x = m[0]; out->t_x = m[2]; out->s_y = m[4]; out->t_y = m[5];
    }
}

// does crop_hand keep pixel (row, col) of an R x R normalised depth image with value val?
__device__ __forceinline__ bool crop_keep(const CropBox& b, const CropParams& cp, int row, int col, int R, float val,
                                          float zc, float zh) {
    const float rm1 = (float)R - 1.f, half = (float)R * 0.5f;
    const float gu = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, (float)col), rm1), 1.f);
    const float gv = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, (float)row), rm1), 1.f);
    const float uu = __fmul_rn(__fadd_rn(gu, 1.f), half), vv = __fmul_rn(__fadd_rn(gv, 1.f), half);
    const float d = __fadd_rn(__fmul_rn(val, zh), zc);
    const float us = __fdiv_rn(__fsub_rn(uu, b.t_x), b.s_x), vs = __fdiv_rn(__fsub_rn(vv, b.t_y), b.s_y);
    const float x = __fdiv_rn(__fmul_rn(__fsub_rn(us, cp.px), d), cp.fx);
    const float y = __fdiv_rn(__fmul_rn(__fsub_rn(vs, cp.py), d), cp.fy);
    return x > b.lo[0] && x < b.hi[0] && y > b.lo[1] && y < b.hi[1] && d > b.lo[2] && d < b.hi[2];
}

__global__ void __launch_bounds__(256)
crop_hand_kernel(int R, CropParams cp, const float* __restrict__ center, const float* __restrict__ cube,
                 const float* __restrict__ img, float* __restrict__ out, unsigned char* __restrict__ keep) {
    __shared__ CropBox box;
    const int b = blockIdx.x;
    if (threadIdx.x < 32) crop_box_warp(cp, b, center, cube, threadIdx.x, &box);
    __syncthreads();
    const float zc = center[3 * b + 2], zh = __fdiv_rn(cube[3 * b + 2], 2.f);
    for (int k = threadIdx.x; k < R * R; k += 256) {
        const float v = img[(size_t)b * R * R + k];
        const bool m = crop_keep(box, cp, k / R, k % R, R, v, zc, zh);
        out[(size_t)b * R * R + k] = m ? v : 1.f;
        if (keep) keep[(size_t)b * R * R + k] = m ? 1 : 0;
    }
}

extern "C" int dsf_crop_hand(int batch, int R, const float* img, const float* joints, int n_joints,
                             const float* center3d, const float* cube, const float* M, const float* intr4,
                             float offset_xy, float offset_z, float thickness, float* out, unsigned char* keep,
                             dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(batch > 0 && R > 1 && img && joints && n_joints > 0 && center3d && cube && M && intr4 && out,
                "null / empty argument");
    CropParams cp = {joints, M, n_joints, intr4[0], intr4[1], intr4[2], intr4[3], offset_xy, offset_z, thickness};
    crop_hand_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(R, cp, center3d, cube, img, out, keep);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

// ------------------------------------------------------------------------------------------------
// forward: grid (tiles, meshes); one CTA owns a 128 x 64 pixel tile of one mesh (half the image at
// R = 128) with its z-buffer in shared memory as packed 64-bit (depth bits << 32 | face) keys.
// Each warp pulls batches of 32 faces and keeps its lanes busy through two re-balancing steps:
//   phase A  lane per face      : cull, exact clipped pixel bbox -> number of rows
//   phase B  lane per (face,row): rows dealt out by a shuffle binary search over the prefix sums; the
//                                 row's pixel run is solved in closed form from the three edge
//                                 crossings (conservative by ~1e-4 px) and appended to a warp-private
//                                 candidate list in shared memory
//   phase C  lane per candidate : full oracle-order fragment evaluation + early z + atomicMin on the key
// A batch whose runs exceed the list is evaluated in place, so any mesh / crop size is handled.
// Epilogue (fused step): background fill + normalisation + m2d loss sums + the raster backward (integer face
// moments, closed-form face gradients, fixed-point vertex scatter), see the comments there.  Template ROWS: the
// loss target arrives as the loader's row-run packed sensor crop and is decoded where it is compared.
// ------------------------------------------------------------------------------------------------
#define RT_TW 128            // tile width  (pixels)
#define RT_TH 64             // tile height: 64 KB of keys -> two CTAs per SM
#define RT_THREADS 512
#define RT_MAXR 512
#define RT_WCANDS 512        // per-warp candidate list, 32-bit entries (16 warps x 512 x 4 B = 32 KB)
#define RT_MAXF 2047         // face id is packed into 11 bits

struct RasterSmem {
    unsigned long long* key;   // RT_TW * RT_TH
    float* vn;                 // NVW * 3 (x_ndc, y_ndc, z)
    float* xs;                 // R
    float* ys;                 // R
    unsigned int* fp;          // F packed vertex ids
    unsigned int* cands;       // per-warp candidate lists
    int* counters;             // [0] next face batch
};

__device__ __forceinline__ RasterSmem carve_smem(unsigned char* raw, int R, int F) {
    RasterSmem s;
    s.key = reinterpret_cast<unsigned long long*>(raw);
    s.vn = reinterpret_cast<float*>(s.key + RT_TW * RT_TH);
    s.xs = s.vn + 2340;
    s.ys = s.xs + R;
    s.fp = reinterpret_cast<unsigned int*>(s.ys + R);
    s.cands = s.fp + ((F + 3) & ~3);
    s.counters = reinterpret_cast<int*>(s.cands + (RT_THREADS / 32) * RT_WCANDS);
    return s;
}

static size_t raster_fwd_smem(int R, int F) {
    return (size_t)RT_TW * RT_TH * 8 + 2340 * 4 + (size_t)2 * R * 4 + (size_t)((F + 3) & ~3) * 4 +
           (size_t)(RT_THREADS / 32) * RT_WCANDS * 4 + 16;
}

// exact evaluation of pixel (i,j) against face f, commit to the z-buffer
template <bool PERSP>
__device__ __forceinline__ void eval_and_commit(const RasterSmem& s, unsigned int f, int i, int j, int tx0,
                                                int ty0, bool zmin_bounds, bool bbox_check) {
    const unsigned int pk = s.fp[f];
    const int a0 = pk & 1023, a1 = (pk >> 10) & 1023, a2 = pk >> 20;
    const float z0 = s.vn[3 * a0 + 2], z1 = s.vn[3 * a1 + 2], z2 = s.vn[3 * a2 + 2];
    unsigned long long* slot = &s.key[(j - ty0) * RT_TW + (i - tx0)];
    // early z: with perspective correction pz is a convex combination of (z0,z1,z2) up to a few ulp; without
    // it the weights sum to A / (A + eps), i.e. pz >= zmin (1 - eps / A) for a front-facing face (more for a
    // back-facing one).  zmin_bounds says eps / A is below the 1e-4 margin (phase A knows the area), so a face
    // whose nearest vertex is clearly behind the stored depth cannot win (NaN compare = empty pixel = proceed).
    const float cur = __uint_as_float((unsigned int)(*slot >> 32));
    if (zmin_bounds && fminf(z0, fminf(z1, z2)) * 0.9999f > cur) return;
    const float x0 = s.vn[3 * a0], y0 = s.vn[3 * a0 + 1];
    const float x1 = s.vn[3 * a1], y1 = s.vn[3 * a1 + 1];
    const float x2 = s.vn[3 * a2], y2 = s.vn[3 * a2 + 1];
    const float px = s.xs[i], py = s.ys[j];
    // the oracle's bounding-box rejection, exact (the candidate ranges of the direct raster are supersets)
    if (bbox_check && (px > fmaxf(x0, fmaxf(x1, x2)) || px < fminf(x0, fminf(x1, x2)) ||
                       py > fmaxf(y0, fmaxf(y1, y2)) || py < fminf(y0, fminf(y1, y2)))) return;
    const float area = __fadd_rn(edge_rn(x2, y2, x0, y0, x1, y1), EPS);
    const float e0 = edge_rn(px, py, x1, y1, x2, y2);
    const float e1 = edge_rn(px, py, x2, y2, x0, y0);
    const float e2 = edge_rn(px, py, x0, y0, x1, y1);
    // Sign pre-test, before any division.  The oracle accepts a fragment iff b0, b1, b2 > 0 with
    // b_i = fl(e_i / area) (then, PERSP, t_i = b_i z z / den).  While no quotient can underflow
    // (|e_i| >= 1e-30, |area| < 1e6) the sign of fl(e_i / area) is the sign of e_i * area, so: without
    // perspective correction a fragment whose e_i do not all carry area's sign is rejected; with it
    // (all z > 0, so sign t_i = sign w_i and b_i = t_i / den) mixed signs are rejected and three
    // negative w_i stay with the exact path (den < 0 makes all b_i positive - only possible for a
    // sliver whose area changes sign when the epsilon is added).  Everything else: exact evaluation.
    if (fabsf(e0) >= 1e-30f && fabsf(e1) >= 1e-30f && fabsf(e2) >= 1e-30f && fabsf(area) < 1e6f) {
        const bool n0 = e0 < 0.f, n1 = e1 < 0.f, n2 = e2 < 0.f;
        if (n0 != n1 || n1 != n2) return;
        if (!PERSP) {
            if (n0 != (area < 0.f)) return;
            // depth to ~1e-6 relative without the IEEE divisions: clearly behind the stored depth -> out
            const float pa = __fdividef(fmaf(e0, z0, fmaf(e1, z1, e2 * z2)), area);
            if (pa * 0.99999f > cur) return;
        }
    }
    FragEval fe = eval_fragment_e<PERSP>(e0, e1, e2, z0, z1, z2, area);
    if (!fe.ok) return;
    const float pz = fe.pz + 0.f;                               // -0 -> +0 so the bit pattern orders
    atomicMin(slot, ((unsigned long long)__float_as_uint(pz) << 32) | f);
}

// Conservative pixel runs (phase B of the forward kernel).  The three edge functions of a face are linear
// in px, e_i(px) = (px - xa_i) * dy_i - r_i with the oracle's own rounded constants, and for a face of
// orientation s = sign(area) the accepted pixels (all three oracle-order edge functions strictly of one
// sign) satisfy s * e_i > 0, i.e. px beyond / before the crossing c_i(py) = xa_i + (py - ya_i) dx_i / dy_i.
// The float evaluation of e_i can disagree with the real sign only within 2.1 u |px - xa_i| of c_i
// (u = 2^-24) and c_i is computed to a few ulp of its terms, so each bound is widened by
// 2e-6 (1 + |xa_i| + (2 + |ya_i|) |dx_i / dy_i|): about 1e-4 of a pixel at R = 128 for ordinary edges,
// never a missed pixel.  Slivers (|area| < 1e-5, where float edge signs need not be consistent with the
// orientation) and near-horizontal edges add no constraint: the whole bbox row stays a candidate.
// Phase C re-tests every candidate exactly, so the runs only need to be supersets.
__device__ __forceinline__ int warp_excl_scan(int v, int lane, int* total) {
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    *total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
}

// Fused backward of the fitting step (optional): gv_tile / gv_flag = this tile's share of d loss / d verts
// (see the epilogue).
struct FusedTail {
    float* gv_tile;                  // (n_mesh, tiles, NVW*3) un-normalised vertex gradient of this tile, or null
    int* gv_flag;                    // (n_mesh, tiles) 1 = gv_tile row written, 0 = tile carries no gradient
};

template <bool PERSP, bool ROWS>
__global__ void __launch_bounds__(RT_THREADS, 2)
raster_fwd_kernel(int R, int tiles_x, const float* __restrict__ verts, const float* __restrict__ place_scale,
                  const float* __restrict__ place_off, const unsigned int* __restrict__ faces_packed,
                  const unsigned short* __restrict__ face_order, int F,
                  const float* __restrict__ view, const float* __restrict__ xs_g, const float* __restrict__ ys_g,
                  float* __restrict__ img, int* __restrict__ p2f, float* __restrict__ zbuf,
                  float* __restrict__ bary, float* __restrict__ dists, const float* __restrict__ target,
                  float thr, float* __restrict__ parts_tile, int use_tma, CropParams crop, FusedTail tail,
                  TargetRows trows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint2 s_rowctx[ROWS ? RT_TH : 1];          // row-run target: (span record, payload offset) per tile row
    const RasterSmem s = carve_smem(smem_raw, R, F);
    // the loss target either as an fp32 plane or in the loader's row-run transport format (decoded in the epilogue)
    const bool has_target = target != nullptr || (ROWS && trows.rows != nullptr);
    const int mesh = blockIdx.y;
    const int tile = blockIdx.x;
    const int tx0 = (tile % tiles_x) * RT_TW, ty0 = (tile / tiles_x) * RT_TH;
    const int tx1 = min(R, tx0 + RT_TW) - 1, ty1 = min(R, ty0 + RT_TH) - 1;
    const int tid = threadIdx.x, lane = tid & 31;
    const ViewRec vw = load_view(view + (size_t)mesh * VIEW);

    const float* vm = verts + (size_t)mesh * NVW * 3;
    const float* ps = place_scale ? place_scale + 3 * mesh : nullptr;
    const float* po = place_off ? place_off + 3 * mesh : nullptr;
    if (target) {
        // the target tile is only needed in the epilogue: pull it into L2 now, behind the raster work
        const int rows = ty1 - ty0 + 1, lines_per_row = ((tx1 - tx0 + 1) * 4 + 127) / 128;
        for (int i = tid; i < rows * lines_per_row; i += RT_THREADS) {
            const float* pa = target + ((size_t)mesh * R + ty0 + i / lines_per_row) * R + tx0 + (i % lines_per_row) * 32;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
        }
    } else if (ROWS && trows.rows && tid < 96) {
        // row-run target: this hand's row records and packed pixels (a few KB) go to L2 now, behind the raster work;
        // one line per thread (64 lines = 8 KB of payload: a hand crop packs to ~6 KB; a longer payload is simply
        // fetched on demand)
        if (tid < 32) {
            const char* rr = reinterpret_cast<const char*>(trows.rows + (size_t)mesh * R * 2);
            if (tid * 128 < R * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(rr + tid * 128));
        } else {
            const unsigned int p0 = __ldg(trows.hand_offset + mesh), p1 = __ldg(trows.hand_offset + mesh + 1);
            const char* pp = reinterpret_cast<const char*>(trows.payload) + ((size_t)p0 * 2 & ~(size_t)127);
            if ((unsigned int)(tid - 32) * 64u < p1 - (p0 & ~63u)) asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + (tid - 32) * 128));
        }
    }
    if (use_tma) {
        // Stage this mesh's inputs with the TMA engine (cp.async.bulk, SASS UBLKCP): sample grids,
        // packed triangle list and the raw vertex block (16-byte aligned window around the 9348-byte
        // row, landed in the not-yet-used candidate lists) - one thread issues, an mbarrier collects the bytes.
        __shared__ __align__(8) unsigned long long tma_bar;
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tma_bar);
        const size_t v_off = (size_t)mesh * NVW * 3 * sizeof(float);
        const uint32_t v_shift = (uint32_t)(v_off & 15);
        const uint32_t v_bytes = (v_shift + NVW * 3 * (uint32_t)sizeof(float) + 15u) & ~15u;
        const uint32_t g_bytes = (uint32_t)R * 4u, f_bytes = (uint32_t)((F + 3) & ~3) * 4u;
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                         "r"(2 * g_bytes + f_bytes + v_bytes)
                         : "memory");
            const char* vsrc = reinterpret_cast<const char*>(verts) + (v_off - v_shift);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(s.xs)),
                         "l"(xs_g + (size_t)mesh * R), "r"(g_bytes), "r"(bar)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(s.ys)),
                         "l"(ys_g + (size_t)mesh * R), "r"(g_bytes), "r"(bar)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(s.fp)),
                         "l"(faces_packed), "r"(f_bytes), "r"(bar)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(s.cands)),
                         "l"(vsrc), "r"(v_bytes), "r"(bar)
                         : "memory");
        }
        {   // z-buffer clear overlaps the copies
            ulonglong2* k2 = reinterpret_cast<ulonglong2*>(s.key);
            for (int i = tid; i < RT_TW * RT_TH / 2; i += RT_THREADS) k2[i] = make_ulonglong2(~0ull, ~0ull);
        }
        __syncthreads();                                   // barrier initialised before anyone polls it
        asm volatile(
            "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n" ::"r"(bar)
            : "memory");
        const float* raw = reinterpret_cast<const float*>(reinterpret_cast<const char*>(s.cands) + v_shift);
        for (int v = tid; v < NVW; v += RT_THREADS) project_vertex(raw + 3 * v, ps, po, vw, s.vn + 3 * v);
    } else {
        for (int i = tid; i < R; i += RT_THREADS) {
            s.xs[i] = xs_g[(size_t)mesh * R + i];
            s.ys[i] = ys_g[(size_t)mesh * R + i];
        }
        for (int i = tid; i < F; i += RT_THREADS) s.fp[i] = faces_packed[i];
        for (int v = tid; v < NVW; v += RT_THREADS) project_vertex(vm + 3 * v, ps, po, vw, s.vn + 3 * v);
        ulonglong2* k2 = reinterpret_cast<ulonglong2*>(s.key);
        for (int i = tid; i < RT_TW * RT_TH / 2; i += RT_THREADS) k2[i] = make_ulonglong2(~0ull, ~0ull);
    }
    if (tid < 4) s.counters[tid] = 0;
    __syncthreads();

    const int cx0 = max(tx0, vw.xlo), cx1 = min(tx1, vw.xhi);
    const int cy0 = max(ty0, vw.ylo), cy1 = min(ty1, vw.yhi);
    const bool tile_live = cx0 <= cx1 && cy0 <= cy1;
    // Each warp pulls batches of 32 faces from a shared counter and runs the three phases on its
    // own private lists (warp-level synchronisation only), so no CTA barrier separates the phases
    // and a slow batch never stalls the other warps; the z-buffer is shared through atomicMin.
    const int warp = tid >> 5;
    unsigned int* my_cands = s.cands + warp * RT_WCANDS;
    int n_cands = 0;
    auto flush_cands = [&]() {
        __syncwarp();                                         // candidate writes of all lanes visible
        for (int c = lane; c < n_cands; c += 32) {
            const unsigned int e = my_cands[c];
            eval_and_commit<PERSP>(s, e & 2047u, tx0 + (int)((e >> 18) & 127u), ty0 + (int)((e >> 11) & 127u), tx0, ty0,
                                   (e >> 25) & 1u, vw.affine);
        }
        n_cands = 0;
        __syncwarp();
    };
    // Phase A (lane / face): cull, exact clipped pixel bbox, and the three edge lines of the face in the
    // form crossing(py) = m * py + c with the conservative margin of row_run() folded into c.
    // Phase B (lane / bbox row): rows are dealt out by a shuffle binary search over the prefix sums of the
    // row counts; the row's lane fetches the owner's lines by shuffle, intersects the three half-lines
    // and appends the pixel run to the warp's candidate list.  Phase C: exact evaluation (flush_cands).
    while (tile_live) {
        int fb = 0;
        if (lane == 0) fb = atomicAdd(&s.counters[0], 32);
        fb = __shfl_sync(0xffffffffu, fb, 0);
        if (fb >= F) break;
        const int f = (fb + lane < F) ? (int)__ldg(face_order + fb + lane) : F;
        int ia = 0, ib = -1, ja = 0, jb = -1;
        float e_m[3] = {0.f, 0.f, 0.f}, e_c[3] = {0.f, 0.f, 0.f};
        unsigned int e_flags = 0;          // bit i: edge i usable; bit 4 + i: edge i bounds the run from below (in x);
                                           // bit 8: zmin is a lower bound of the face's depth (early z allowed)
        if (f < F) {
            const unsigned int pk = s.fp[f];
            const int a0 = pk & 1023, a1 = (pk >> 10) & 1023, a2 = pk >> 20;
            const float x0 = s.vn[3 * a0], y0 = s.vn[3 * a0 + 1], z0 = s.vn[3 * a0 + 2];
            const float x1 = s.vn[3 * a1], y1 = s.vn[3 * a1 + 1], z1 = s.vn[3 * a1 + 2];
            const float x2 = s.vn[3 * a2], y2 = s.vn[3 * a2 + 1], z2 = s.vn[3 * a2 + 2];
            const float zmin = fminf(z0, fminf(z1, z2));
            const float farea = edge_rn(x0, y0, x1, y1, x2, y2);
            if (zmin >= EPS && !(farea <= EPS && farea >= -EPS)) {
                if (PERSP || farea < 0.f || farea > 2e-4f) e_flags |= 256u;      // eps / A < 5e-5
                const float xmin = fminf(x0, fminf(x1, x2)), xmax = fmaxf(x0, fmaxf(x1, x2));
                const float ymin = fminf(y0, fminf(y1, y2)), ymax = fmaxf(y0, fmaxf(y1, y2));
                if (vw.affine) {
                    // direct raster: index = a * ndc + b exactly, so the bbox maps to index ranges in closed form; a
                    // thousandth of a pixel of slack makes them supersets, the exact bbox comparison of the oracle is
                    // repeated per candidate (eval_and_commit)
                    const float lim = 1e6f;
                    ja = max(cy0, (int)ceilf(fminf(fmaxf(fmaf(vw.ay, ymax, vw.by) - 1e-3f, -lim), lim)));
                    jb = min(cy1, (int)floorf(fminf(fmaxf(fmaf(vw.ay, ymin, vw.by) + 1e-3f, -lim), lim)));
                    ia = max(cx0, (int)ceilf(fminf(fmaxf(fmaf(vw.ax, xmax, vw.bx) - 1e-3f, -lim), lim)));
                    ib = min(cx1, (int)floorf(fminf(fmaxf(fmaf(vw.ax, xmin, vw.bx) + 1e-3f, -lim), lim)));
                } else {
                    ja = max(cy0, first_le(s.ys, vw.ylo, vw.yhi, vw.ay, vw.by, ymax));
                    jb = min(cy1, last_ge(s.ys, vw.ylo, vw.yhi, vw.ay, vw.by, ymin));
                    if (ja <= jb) {
                        ia = max(cx0, first_le(s.xs, vw.xlo, vw.xhi, vw.ax, vw.bx, xmax));
                        ib = min(cx1, last_ge(s.xs, vw.xlo, vw.xhi, vw.ax, vw.bx, xmin));
                    }
                }
                if (ia <= ib && ja <= jb && fabsf(farea) >= 1e-5f) {     // slivers keep the whole bbox row
                    // edge i of the oracle: e_i(px) = (px - xa) * dy - (py - ya) * dx, crossing at
                    // px = xa + (py - ya) * dx / dy = m * py + (xa - ya * m).  The margin covers the oracle's
                    // rounding (2.1 u |px - xa|), the approximate division and the cancellation in m * py + c.
                    const float sg = farea > 0.f ? 1.f : -1.f;
                    const float xa[3] = {x1, x2, x0}, ya[3] = {y1, y2, y0}, xb[3] = {x2, x0, x1}, yb[3] = {y2, y0, y1};
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const float dy = __fsub_rn(yb[i], ya[i]);
                        const float m = __fdividef(__fsub_rn(xb[i], xa[i]), dy);
                        const float am = fabsf(m);
                        if (am < 1e30f) {                                  // else near-horizontal: no constraint
                            const float dl = 2e-6f * (1.f + fabsf(xa[i]) + (2.f + fabsf(ya[i])) * am);
                            const bool low = sg * dy > 0.f;
                            e_m[i] = m;
                            e_c[i] = fmaf(-ya[i], m, xa[i]) + (low ? -dl : dl);
                            e_flags |= (1u << i) | (low ? 16u << i : 0u);
                        }
                    }
                }
            }
        }
        const int n = (ib >= ia && jb >= ja) ? jb - ja + 1 : 0;
        const unsigned int own = (unsigned int)f | ((unsigned int)(ia - tx0) << 11) | ((unsigned int)(ib - tx0) << 18) |
                                 ((unsigned int)(ja - ty0) << 25);
        int total;
        const int excl = warp_excl_scan(n, lane, &total);
        const int incl = excl + n;
        for (int it0 = 0; it0 < total; it0 += 32) {
            const int t = it0 + lane;
            // owner = number of lanes whose inclusive prefix is <= t
            int ow = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int v = __shfl_sync(0xffffffffu, incl, ow + step - 1);
                if (v <= t) ow += step;
            }
            const bool live = t < total;
            ow = live ? ow : 0;
            const unsigned int o_pk = __shfl_sync(0xffffffffu, own, ow);
            const int o_ex = __shfl_sync(0xffffffffu, excl, ow);
            const unsigned int o_fl = __shfl_sync(0xffffffffu, e_flags, ow);
            float lo = -INFINITY, hi = INFINITY;
            const int j = ty0 + (int)(o_pk >> 25) + (t - o_ex);
            const float py = s.ys[live ? j : ty0];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float m = __shfl_sync(0xffffffffu, e_m[i], ow), c = __shfl_sync(0xffffffffu, e_c[i], ow);
                const float v = fmaf(m, py, c);
                if (o_fl & (1u << i)) {
                    if (o_fl & (16u << i)) lo = fmaxf(lo, v);
                    else hi = fminf(hi, v);
                }
            }
            int ka = 0, kb = -1;
            if (live) {
                const int ia_o = tx0 + (int)((o_pk >> 11) & 127u), ib_o = tx0 + (int)((o_pk >> 18) & 127u);
                if (vw.affine) {
                    // direct mode: index = ax * x + bx exactly and xs is non-increasing, so x <= hi <=> index >=
                    // ax * hi + bx; a thousandth of a pixel of slack covers the rounding of the map
                    ka = max(ia_o, (int)ceilf(fmaf(vw.ax, hi, vw.bx) - 1e-3f));
                    kb = min(ib_o, (int)floorf(fmaf(vw.ax, lo, vw.bx) + 1e-3f));
                } else {
                    ka = hi < INFINITY ? max(ia_o, first_le(s.xs, vw.xlo, vw.xhi, vw.ax, vw.bx, hi)) : ia_o;
                    kb = lo > -INFINITY ? min(ib_o, last_ge(s.xs, vw.xlo, vw.xhi, vw.ax, vw.bx, lo)) : ib_o;
                }
            }
            const unsigned int fi = o_pk & 2047u;
            const int len = max(0, kb - ka + 1);
            int c_total;
            int slot = warp_excl_scan(len, lane, &c_total);
            if (n_cands + c_total > RT_WCANDS) flush_cands();
            if (c_total > RT_WCANDS) {                           // very large faces: evaluate in place
                for (int k = ka; k <= kb; ++k) eval_and_commit<PERSP>(s, fi, k, j, tx0, ty0, (o_fl >> 8) & 1u, vw.affine);
                __syncwarp();
                continue;
            }
            slot += n_cands;
            const unsigned int base = fi | ((unsigned int)(j - ty0) << 11) | ((o_fl & 256u) << 17);
            // runs are mostly one or two pixels long: the first two stores are straight-line code
            if (len > 0) my_cands[slot] = base | ((unsigned int)(ka - tx0) << 18);
            if (len > 1) my_cands[slot + 1] = base | ((unsigned int)(ka + 1 - tx0) << 18);
            for (int k = ka + 2; k <= kb; ++k) my_cands[slot + (k - ka)] = base | ((unsigned int)(k - tx0) << 18);
            n_cands += c_total;
        }
    }
    flush_cands();                                            // ---------------- phase C (remainder)
    // Every warp prepares its own, now dead, candidate-list region for the epilogue BEFORE it waits for the others
    // (the epilogue's tables live in those regions, see below): zero fill = the face moments, the fixed-point vertex
    // sums and the gradient maxima start at zero; warp 0 also owns the words of the crop box and of the pixel-index
    // tables.  Work a warp does while the stragglers of the face loop finish is free, and the epilogue needs no
    // barrier of its own before the pixel passes.  (Scoped: nothing here stays live across the barrier.)
    {
        const bool has_t = target != nullptr || (ROWS && trows.rows != nullptr);
        const bool grad_pre = !PERSP && tail.gv_tile != nullptr && has_t;
        if (grad_pre) {
            uint4* z4 = reinterpret_cast<uint4*>(my_cands);
            for (int i = lane; i < RT_WCANDS / 4; i += 32) z4[i] = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
        }
        if (ROWS && trows.rows && warp == 1 && (R & 3) == 0) {
            // row-run target: (span record, payload offset) of every row of the tile, once per CTA.  One warp scan per
            // 128 rows of the hand: lane l owns rows 4 l .. 4 l + 3 (one 128-bit load of their records); the exclusive
            // prefix of the span lengths is each row's offset into the payload.
            const unsigned int* r32 = reinterpret_cast<const unsigned int*>(trows.rows) + (size_t)mesh * R;
            unsigned int carry = __ldg(trows.hand_offset + mesh);
            for (int c0 = 0; c0 <= ty1; c0 += 128) {
                const int q = c0 + 4 * lane;
                uint4 rc = make_uint4(0u, 0u, 0u, 0u);
                if (q + 3 < R) rc = __ldg(reinterpret_cast<const uint4*>(r32 + q));
                const unsigned int rr[4] = {rc.x, rc.y, rc.z, rc.w};
                const unsigned int e[4] = {0u, rc.x >> 16, (rc.x >> 16) + (rc.y >> 16), (rc.x >> 16) + (rc.y >> 16) + (rc.z >> 16)};
                const unsigned int tot = e[3] + (rc.w >> 16);
                unsigned int incl = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                const unsigned int excl = carry + incl - tot;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (q + i >= ty0 && q + i <= ty1) s_rowctx[q + i - ty0] = make_uint2(rr[i], excl + e[i]);
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        if (warp == 0) {
            if (has_t && crop.joints) crop_box_warp(crop, mesh, place_off, place_scale, lane, reinterpret_cast<CropBox*>(s.cands));
            if (grad_pre) {
                int* qx_ = reinterpret_cast<int*>(s.cands + 96);
                int* qy_ = qx_ + RT_TW;
                const int tw_ = tx1 - tx0 + 1, th_ = ty1 - ty0 + 1;
                const float cx = s.xs[tx0], cy = s.ys[ty0];
                const int q0x_ = (cx == cx) ? __float2int_rn(((1.f - cx) * vw.S - 1.f) * 0.5f) : 0;
                const int q0y_ = (cy == cy) ? __float2int_rn(((1.f - cy) * vw.S - 1.f) * 0.5f) : 0;
                for (int i = lane; i < tw_ + th_; i += 32) {
                    const bool isx = i < tw_;
                    const float c = isx ? s.xs[tx0 + i] : s.ys[ty0 + i - tw_];
                    const int q = (c == c) ? __float2int_rn(((1.f - c) * vw.S - 1.f) * 0.5f) : 0;
                    if (isx) qx_[i] = q - q0x_; else qy_[i - tw_] = q - q0y_;
                }
            }
        }
    }
    __syncthreads();                                          // z-buffer complete, epilogue tables ready


    // epilogue: background fill (:1084-1085) + normalize_img (:1289-1299); with a target image the
    // m2d loss partial sums of this tile (train_render.py:728-731) are produced on the way out
    const float zmax = __fadd_rn(vw.zc, vw.zh), zmin_c = __fsub_rn(vw.zc, vw.zh);
    const int tw = tx1 - tx0 + 1, th = ty1 - ty0 + 1;
    float l_sum = 0.f, l_cnt = 0.f;
    // scratch carved from the (now dead) candidate lists
    unsigned int* scratch = s.cands;
    // optional crop_hand of the rendered image before the loss (train_render.py:727): the stored
    // image stays uncropped, only the loss (and its gradient) sees the crop
    CropBox* cbox = reinterpret_cast<CropBox*>(scratch);                            // words 0 .. 15
    float* red = reinterpret_cast<float*>(scratch + 32);                            // 32 .. 95
    int* qx = reinterpret_cast<int*>(scratch + 96);                                 // raster column of each tile column, minus q0x
    int* qy = qx + RT_TW;                                                           // ... rows
    int* s_gmax = reinterpret_cast<int*>(scratch + 288);                            // float bits of max |face gradient| (xy, z)
    int* mom = reinterpret_cast<int*>(scratch + 320);                               // (F,3) face moments: sum s, sum s dqx, sum s dqy
    const int Fp = (F + 31) & ~31;
    unsigned short* flist = reinterpret_cast<unsigned short*>(mom + 3 * Fp);        // per-warp lists of touched faces (after the
                                                                                    // pixel loop; before, its group lists)
    int* sgn = reinterpret_cast<int*>(flist + Fp);                                  // (NVW,3) NDC vertex gradients, fixed point
    const bool do_crop = has_target && crop.joints;
    // Fused backward (no perspective correction): the depth of a face is affine in the sample position,
    // pz(p) = [z0 e0(p) + z1 e1(p) + z2 e2(p)] / area, so the whole zbuf cotangent of a face collapses to the
    // three moments (sum g, sum g px, sum g py) over its pixels, and d loss / d img of the m2d loss is
    // +-gk(hand) on the union mask.  The epilogue accumulates integer moments of sign(synth - real) per face -
    // px = 1 - (2 q + 1) / S with the integer raster pixel q, relative to the tile's first sample - with native
    // 32-bit shared-memory atomics (64-bit and float adds are compare-and-swap loops on this architecture, measured
    // 2x slower here); each warp then compacts its share of the face list to the touched faces and evaluates their
    // closed-form gradient (float64); the face gradients are scattered to the vertices in fixed point (scale = a
    // power of two from the tile's largest entry, again native integer atomics: exact, order-independent sums), and
    // the vertices are chained through the projection.  The per-hand factor gk / zhalf is applied by the consumer
    // (it needs the mask count of all tiles).  No pix_to_face plane, no second pass over target / img, no backward
    // kernel, bit-reproducible.
    const bool do_grad = !PERSP && tail.gv_tile != nullptr && has_target;
    const float cx0s = s.xs[tx0], cy0s = s.ys[ty0];
    const int q0x = (cx0s == cx0s) ? __float2int_rn(((1.f - cx0s) * vw.S - 1.f) * 0.5f) : 0;
    const int q0y = (cy0s == cy0s) ? __float2int_rn(((1.f - cy0s) * vw.S - 1.f) * 0.5f) : 0;
    auto add_moments = [&](int f, int n, int mi, int dqy) {
        if (f >= 0 && (n | mi) != 0) {
            if (n) { atomicAdd(&mom[3 * f], n); atomicAdd(&mom[3 * f + 2], n * dqy); }
            if (mi) atomicAdd(&mom[3 * f + 1], mi);
        }
    };
    const bool fast = !zbuf && !bary && !dists && (R & 3) == 0 && (tw & 3) == 0;
    if (fast) {
        // default outputs only: four pixels per lane, 128-bit key / target loads and image stores.  Two passes
        // per warp over its own rows: (1) groups of four background pixels (most of the image) take the
        // precomputed normalised far plane on the spot, every other group is pushed on a warp-private list;
        // (2) the list is processed with all lanes busy (a row through the hand has ~1/3 of its groups on the hand).
        const float bgval = __fdiv_rn(__fsub_rn(zmax, vw.zc), vw.zh);
        const int qw = tw >> 2;
        // the four target values of a pixel group: fp32 plane, or decoded from the row's span (outside it the value
        // target_norm gives depth 0, which is bgval)
        auto target4 = [&](size_t o, unsigned int rec, unsigned int off, int x) -> float4 {
            if (!ROWS || target) return __ldg(reinterpret_cast<const float4*>(target + o));
            float t[4] = {bgval, bgval, bgval, bgval};
            const int s0 = (int)(rec & 0xffffu), n = (int)(rec >> 16);
            if (x + 3 >= s0 && x < s0 + n) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k = x + c - s0;
                    if ((unsigned int)k < (unsigned int)n)
                        t[c] = target_norm(__ldg(trows.payload + off + k), trows.invalid, vw.zc, vw.zh, zmax, zmin_c);
                }
            }
            return make_float4(t[0], t[1], t[2], t[3]);
        };
        // general group: depth normalisation, loss terms, moments of the loss gradient
        auto process_group = [&](int ly, int lx, const float4& tg, const ulonglong2& k01, const ulonglong2& k23) {
            const size_t o = ((size_t)mesh * R + (ty0 + ly)) * R + (tx0 + lx);
            const unsigned long long kk[4] = {k01.x, k01.y, k23.x, k23.y};
            const float tt[4] = {tg.x, tg.y, tg.z, tg.w};
            float v[4];
            int ff[4];
            int run_f = -1, run_n = 0, run_mi = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                bool grad_ok = false;                 // gates of the depth normalisation: bg fill and clamp pass no gradient
                if (kk[c] == ~0ull) {
                    v[c] = bgval;
                    ff[c] = -1;
                } else {
                    const float z = __uint_as_float((unsigned int)(kk[c] >> 32));
                    float d = z <= 0.f ? 0.f : z;
                    d = (d == 0.f) ? zmax : d;
                    d = d > zmax ? zmax : d;
                    d = d < zmin_c ? zmin_c : d;
                    v[c] = __fdiv_rn(__fsub_rn(d, vw.zc), vw.zh);
                    ff[c] = (int)(unsigned int)(kk[c] & 0xffffffffu);
                    grad_ok = z > 0.f && !(z > zmax) && !(z < zmin_c);
                }
                if (has_target) {
                    const bool kept = !(do_crop && !crop_keep(*cbox, crop, ty0 + ly, tx0 + lx + c, R, v[c], vw.zc, vw.zh));
                    const float vc = kept ? v[c] : 1.f;
                    const bool m = tt[c] < thr || vc < thr;
                    if (m) { l_sum += fabsf(tt[c] - vc); l_cnt += 1.f; }
                    if (do_grad && grad_ok && kept && m) {
                        const float d = vc - tt[c];
                        const int sg = d > 0.f ? 1 : (d < 0.f ? -1 : 0);
                        if (sg) {
                            if (ff[c] != run_f) {
                                add_moments(run_f, run_n, run_mi, qy[ly]);
                                run_f = ff[c]; run_n = 0; run_mi = 0;
                            }
                            run_n += sg;
                            run_mi += sg * qx[lx + c];
                        }
                    }
                }
            }
            if (do_grad) add_moments(run_f, run_n, run_mi, qy[ly]);
            *reinterpret_cast<float4*>(img + o) = make_float4(v[0], v[1], v[2], v[3]);
            if (p2f) *reinterpret_cast<int4*>(p2f + o) = make_int4(ff[0], ff[1], ff[2], ff[3]);
        };
        // the list needs <= 128 groups per warp (64 x 128 tile: 4 rows x 32 groups) and room behind the tables
        unsigned char* glist = reinterpret_cast<unsigned char*>(flist) + warp * 128;     // flist is not in use yet
        const bool use_list = qw <= 32 && th <= 4 * (RT_THREADS / 32) && Fp * 2 >= (RT_THREADS / 32) * 128;
        int n_list = 0;
        for (int ly = warp, r = 0; ly < th; ly += RT_THREADS / 32, ++r) {
            const uint2 rctx = ROWS ? s_rowctx[ly] : make_uint2(0u, 0u);        // (span record, payload offset) of this row
            const unsigned int rec = rctx.x, off = rctx.y;
            for (int q4 = lane; q4 < ((qw + 31) & ~31); q4 += 32) {
                const bool in = q4 < qw;
                const int lx = q4 * 4;
                const size_t o = ((size_t)mesh * R + (ty0 + ly)) * R + (tx0 + lx);
                float4 tg = make_float4(1.f, 1.f, 1.f, 1.f);
                ulonglong2 k01 = make_ulonglong2(~0ull, ~0ull), k23 = k01;
                // row-run target: groups that touch the row's span are decoded in pass 2 (all lanes busy there; in
                // this row-ordered pass only the lanes over the span would work); everywhere else the target is bgval
                bool in_span = false;
                if (in) {
                    if (ROWS && !target) {
                        const int s0 = (int)(rec & 0xffffu), x = tx0 + lx;
                        in_span = has_target && x + 3 >= s0 && x < s0 + (int)(rec >> 16);
                        tg = make_float4(bgval, bgval, bgval, bgval);
                    } else if (has_target) {
                        tg = target4(o, rec, off, tx0 + lx);
                    }
                    k01 = *reinterpret_cast<const ulonglong2*>(&s.key[ly * RT_TW + lx]);
                    k23 = *reinterpret_cast<const ulonglong2*>(&s.key[ly * RT_TW + lx + 2]);
                }
                const bool all_bg = (k01.x & k01.y & k23.x & k23.y) == ~0ull && !do_crop && !in_span;
                if (in && all_bg) {
                    // four background pixels: far plane out, loss only where the target has depth
                    if (has_target) {
                        const float tt[4] = {tg.x, tg.y, tg.z, tg.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (tt[c] < thr || bgval < thr) { l_sum += fabsf(tt[c] - bgval); l_cnt += 1.f; }
                    }
                    *reinterpret_cast<float4*>(img + o) = make_float4(bgval, bgval, bgval, bgval);
                    if (p2f) *reinterpret_cast<int4*>(p2f + o) = make_int4(-1, -1, -1, -1);
                }
                const bool general = in && !all_bg;
                if (use_list) {
                    const unsigned int mk = __ballot_sync(0xffffffffu, general);
                    if (general) glist[n_list + __popc(mk & ((1u << lane) - 1u))] = (unsigned char)((r << 5) | q4);
                    n_list += __popc(mk);
                } else if (general) {
                    if (ROWS && !target && has_target) tg = target4(o, rec, off, tx0 + lx);
                    process_group(ly, lx, tg, k01, k23);
                }
            }
        }
        if (use_list) {
            __syncwarp();
            for (int k = lane; k < n_list; k += 32) {
                const int e = glist[k];
                const int ly = warp + (e >> 5) * (RT_THREADS / 32), lx = (e & 31) * 4;
                const uint2 rctx = ROWS ? s_rowctx[ly] : make_uint2(0u, 0u);
                const unsigned int rec = rctx.x, off = rctx.y;
                float4 tg = make_float4(1.f, 1.f, 1.f, 1.f);
                if (has_target) tg = target4(((size_t)mesh * R + (ty0 + ly)) * R + (tx0 + lx), rec, off, tx0 + lx);
                const ulonglong2 k01 = *reinterpret_cast<const ulonglong2*>(&s.key[ly * RT_TW + lx]);
                const ulonglong2 k23 = *reinterpret_cast<const ulonglong2*>(&s.key[ly * RT_TW + lx + 2]);
                process_group(ly, lx, tg, k01, k23);
            }
        }
    } else
    for (int ly = tid >> 5; ly < th; ly += RT_THREADS / 32) {
        // all target loads of the row are issued before any of them is consumed
        float tg[RT_TW / 32];
        if (target) {
#pragma unroll
            for (int q = 0; q < RT_TW / 32; ++q) {
                const int lx = lane + 32 * q;
                tg[q] = lx < tw ? __ldg(target + ((size_t)mesh * R + (ty0 + ly)) * R + (tx0 + lx)) : 1.f;
            }
        } else if (ROWS && trows.rows) {
            // row-run target, general path: the row's payload offset = prefix sum of the lengths of the rows above
            const unsigned int* r32 = reinterpret_cast<const unsigned int*>(trows.rows) + (size_t)mesh * R;
            unsigned int off = 0u;
            for (int q0 = 0; q0 < ty0 + ly; q0 += 32) off += q0 + lane < ty0 + ly ? __ldg(r32 + q0 + lane) >> 16 : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) off += __shfl_xor_sync(0xffffffffu, off, o);
            off += __ldg(trows.hand_offset + mesh);
            const unsigned int rec = __ldg(r32 + ty0 + ly);
            const int s0 = (int)(rec & 0xffffu), n = (int)(rec >> 16);
            const float bgt = __fdiv_rn(__fsub_rn(zmax, vw.zc), vw.zh);
#pragma unroll
            for (int q = 0; q < RT_TW / 32; ++q) {
                const int k = tx0 + lane + 32 * q - s0;
                tg[q] = (unsigned int)k < (unsigned int)n
                            ? target_norm(__ldg(trows.payload + off + k), trows.invalid, vw.zc, vw.zh, zmax, zmin_c) : bgt;
            }
        }
#pragma unroll
        for (int q = 0; q < RT_TW / 32; ++q) {
            const int lx = lane + 32 * q;
            if (lx >= tw) continue;
            const unsigned long long key = s.key[ly * RT_TW + lx];
            const int f = key == ~0ull ? -1 : (int)(unsigned int)(key & 0xffffffffu);
            const float z = f < 0 ? -1.f : __uint_as_float((unsigned int)(key >> 32));
            float d = z <= 0.f ? 0.f : z;
            d = (d == 0.f) ? zmax : d;
            d = d > zmax ? zmax : d;
            d = d < zmin_c ? zmin_c : d;
            const size_t o = ((size_t)mesh * R + (ty0 + ly)) * R + (tx0 + lx);
            const float val = __fdiv_rn(__fsub_rn(d, vw.zc), vw.zh);
            img[o] = val;
            if (p2f) p2f[o] = f;
            if (has_target) {
                const float t = tg[q];
                const bool kept = !(do_crop && !crop_keep(*cbox, crop, ty0 + ly, tx0 + lx, R, val, vw.zc, vw.zh));
                const float vc = kept ? val : 1.f;
                const bool m = t < thr || vc < thr;
                if (m) { l_sum += fabsf(t - vc); l_cnt += 1.f; }
                if (do_grad && f >= 0 && kept && m && z > 0.f && !(z > zmax) && !(z < zmin_c)) {
                    const float dd = vc - t;
                    const int sg = dd > 0.f ? 1 : (dd < 0.f ? -1 : 0);
                    add_moments(f, sg, sg * qx[lx], qy[ly]);
                }
            }
            if (zbuf) zbuf[o] = z;
            if (bary || dists) {
                float b0 = -1.f, b1 = -1.f, b2 = -1.f, dd = -1.f;
                if (f >= 0) {
                    const unsigned int pk = s.fp[f];
                    const int i0 = pk & 1023, i1 = (pk >> 10) & 1023, i2 = pk >> 20;
                    const float x0 = s.vn[3 * i0], y0 = s.vn[3 * i0 + 1], z0 = s.vn[3 * i0 + 2];
                    const float x1 = s.vn[3 * i1], y1 = s.vn[3 * i1 + 1], z1 = s.vn[3 * i1 + 2];
                    const float x2 = s.vn[3 * i2], y2 = s.vn[3 * i2 + 1], z2 = s.vn[3 * i2 + 2];
                    const float px = s.xs[tx0 + lx], py = s.ys[ty0 + ly];
                    const float area = __fadd_rn(edge_rn(x2, y2, x0, y0, x1, y1), EPS);
                    FragEval fe = eval_fragment<PERSP>(px, py, x0, y0, z0, x1, y1, z1, x2, y2, z2, area);
                    b0 = fe.b0; b1 = fe.b1; b2 = fe.b2;
                    float d01 = seg_dist_rn(px, py, x0, y0, x1, y1);
                    float d02 = seg_dist_rn(px, py, x0, y0, x2, y2);
                    float d12 = seg_dist_rn(px, py, x1, y1, x2, y2);
                    float m = d01 < d02 ? d01 : d02;
                    m = m < d12 ? m : d12;
                    dd = -m;
                }
                if (bary) { bary[3 * o] = b0; bary[3 * o + 1] = b1; bary[3 * o + 2] = b2; }
                if (dists) dists[o] = dd;
            }
        }
    }
    if (do_grad) {
        __syncthreads();                                      // moments complete, z-buffer keys dead
        // each warp owns a contiguous share of the faces: compact it to the touched ones (ballot, no atomics),
        // then one lane per touched face
        const int share = (F + RT_THREADS / 32 - 1) / (RT_THREADS / 32);
        const int f_lo = warp * share, f_hi = min(F, f_lo + share);
        unsigned short* wl = flist + f_lo;
        float* stg = reinterpret_cast<float*>(s.key) + (size_t)f_lo * 9;     // this warp's face gradients (key storage)
        int cnt = 0;
        for (int f0 = f_lo; f0 < f_hi; f0 += 32) {
            const int f = f0 + lane;
            const bool act = f < f_hi && (mom[3 * f] | mom[3 * f + 1] | mom[3 * f + 2]) != 0;
            const unsigned int mk = __ballot_sync(0xffffffffu, act);
            if (act) wl[cnt + __popc(mk & ((1u << lane) - 1u))] = (unsigned short)f;
            cnt += __popc(mk);
        }
        __syncwarp();
        float gmax_xy = 0.f, gmax_z = 0.f;
        for (int k = lane; k < cnt; k += 32) {
            const int f = wl[k];
            const int n = mom[3 * f];
            const int mi = mom[3 * f + 1] + q0x * n, mj = mom[3 * f + 2] + q0y * n;
            const unsigned int pk = s.fp[f];
            const int i0 = pk & 1023, i1 = (pk >> 10) & 1023, i2 = pk >> 20;
            // float64 for the handful of operations per touched face: sliver faces amplify rounding by
            // extent^2 / area, and the moments are exact integers worth keeping exact
            const double x0 = s.vn[3 * i0], y0 = s.vn[3 * i0 + 1], z0 = s.vn[3 * i0 + 2];
            const double ax = s.vn[3 * i1] - x0, ay = s.vn[3 * i1 + 1] - y0, dz1 = s.vn[3 * i1 + 2] - z0;
            const double bx = s.vn[3 * i2] - x0, by = s.vn[3 * i2 + 1] - y0, dz2 = s.vn[3 * i2 + 2] - z0;
            // U = sum s (p - v0) with p = 1 - (2 q + 1) / S: exact integer sums, scaled once
            const double inv_S = 1.0 / (double)vw.S, S0 = (double)n;
            const double Ux = S0 * (1.0 - inv_S - x0) - 2.0 * inv_S * (double)mi;
            const double Uy = S0 * (1.0 - inv_S - y0) - 2.0 * inv_S * (double)mj;
            double g[9];
            face_grad<double>(z0, ax, ay, dz1, bx, by, dz2, S0, Ux, Uy, g);
#pragma unroll
            for (int e = 0; e < 9; ++e) {
                const float gf = (float)g[e];
                stg[k * 9 + e] = gf;
                if (e % 3 == 2) gmax_z = fmaxf(gmax_z, fabsf(gf)); else gmax_xy = fmaxf(gmax_xy, fabsf(gf));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            gmax_xy = fmaxf(gmax_xy, __shfl_xor_sync(0xffffffffu, gmax_xy, o));
            gmax_z = fmaxf(gmax_z, __shfl_xor_sync(0xffffffffu, gmax_z, o));
        }
        if (lane == 0 && cnt > 0) {       // non-negative floats order like their bit patterns
            atomicMax(&s_gmax[0], __float_as_int(gmax_xy));
            atomicMax(&s_gmax[1], __float_as_int(gmax_z));
        }
        const int tile_has = __syncthreads_or(cnt > 0 ? 1 : 0);
        const size_t slot = (size_t)mesh * gridDim.x + tile;
        if (tid == 0) tail.gv_flag[slot] = tile_has;
        if (tile_has) {
            // fixed point: the largest entry maps to [2^26, 2^27), which leaves room for 16 incident faces per vertex
            const float m_xy = __int_as_float(s_gmax[0]), m_z = __int_as_float(s_gmax[1]);
            int e_xy = 0, e_z = 0;
            frexpf(m_xy, &e_xy);
            frexpf(m_z, &e_z);
            const float q_xy = m_xy > 0.f && m_xy < INFINITY ? ldexpf(1.f, 27 - e_xy) : 0.f;
            const float q_z = m_z > 0.f && m_z < INFINITY ? ldexpf(1.f, 27 - e_z) : 0.f;
            for (int k = lane; k < cnt; k += 32) {
                const unsigned int pk = s.fp[wl[k]];
                const int iv[3] = {(int)(pk & 1023), (int)((pk >> 10) & 1023), (int)(pk >> 20)};
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    atomicAdd(&sgn[3 * iv[c]], __float2int_rn(stg[k * 9 + 3 * c] * q_xy));
                    atomicAdd(&sgn[3 * iv[c] + 1], __float2int_rn(stg[k * 9 + 3 * c + 1] * q_xy));
                    atomicAdd(&sgn[3 * iv[c] + 2], __float2int_rn(stg[k * 9 + 3 * c + 2] * q_z));
                }
            }
            __syncthreads();
            const float r_xy = q_xy > 0.f ? 1.f / q_xy : 0.f, r_z = q_z > 0.f ? 1.f / q_z : 0.f;
            float* go = tail.gv_tile + slot * NVW * 3;
            float sxp = 1.f, syp = 1.f, szp = 1.f;
            if (ps) { sxp = ps[0] * 0.5f; syp = ps[1] * 0.5f; szp = ps[2] * 0.5f; }
            for (int v = tid; v < NVW; v += RT_THREADS) {
                const float gxn = (float)sgn[3 * v] * r_xy, gyn = (float)sgn[3 * v + 1] * r_xy, gzn = (float)sgn[3 * v + 2] * r_z;
                // x_ndc = -fxn x / z + pxn (same for y), z_ndc = z; fxn x / z = pxn - x_ndc
                const float xn = s.vn[3 * v], yn = s.vn[3 * v + 1], iz = 1.f / s.vn[3 * v + 2];
                go[3 * v] = -gxn * vw.fxn * iz * sxp;
                go[3 * v + 1] = -gyn * vw.fyn * iz * syp;
                go[3 * v + 2] = (gzn + (gxn * (vw.pxn - xn) + gyn * (vw.pyn - yn)) * iz) * szp;
            }
        }
    }
    if (!parts_tile) return;
    // fixed-order block reduction of the tile's loss sums (the consumers fold the tiles of a mesh)
    l_sum = warp_sum(l_sum);
    l_cnt = warp_sum(l_cnt);
    if (lane == 0) { red[2 * warp] = l_sum; red[2 * warp + 1] = l_cnt; }
    __syncthreads();
    if (tid == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < RT_THREADS / 32; ++w) { a += red[2 * w]; b += red[2 * w + 1]; }
        float* out = parts_tile + ((size_t)mesh * gridDim.x + tile) * 2;
        out[0] = a;
        out[1] = b;
    }
}

extern "C" int dsf_raster_tiles(int R) { return ((R + RT_TW - 1) / RT_TW) * ((R + RT_TH - 1) / RT_TH); }

int dsf_raster_forward_impl(const DsfMano* h, int n_mesh, const float* verts, const float* place_scale,
                            const float* place_off, const float* view, const float* xs, const float* ys,
                            int R, float* img, int* p2f, float* zbuf, float* bary, float* dists,
                            const float* target, float thr, float* parts_tile, const CropParams* crop,
                            int flags, const RasterFused* fused, cudaStream_t st, const TargetRows* trows) {
    if (h->n_faces > RT_MAXF) {
        dsf_set_error("rasteriser supports at most %d faces (got %d)", RT_MAXF, h->n_faces);
        return DSF_ERR_UNSUPPORTED;
    }
    const bool persp = (flags & DSF_RASTER_PERSPECTIVE_CORRECT) != 0;
    const int tiles_x = (R + RT_TW - 1) / RT_TW, tiles_y = (R + RT_TH - 1) / RT_TH;
    const size_t smem = raster_fwd_smem(R, h->n_faces);
    static bool attr_set[16] = {};
    int dev = 0;
    DSF_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 16 || !attr_set[dev]) {
        DSF_CHECK_CUDA(cudaFuncSetAttribute(raster_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)raster_fwd_smem(RT_MAXR, RT_MAXF)));
        DSF_CHECK_CUDA(cudaFuncSetAttribute(raster_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)raster_fwd_smem(RT_MAXR, RT_MAXF)));
        DSF_CHECK_CUDA(cudaFuncSetAttribute(raster_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)raster_fwd_smem(RT_MAXR, RT_MAXF)));
        if (dev < 16) attr_set[dev] = true;
    }
    FusedTail tail = {};
    if (fused && fused->gv_tile && dsf_raster_fused_grad_ok(h, flags)) {
        tail.gv_tile = fused->gv_tile;
        tail.gv_flag = fused->gv_flag;
    }
    // bulk (TMA) staging needs 16-byte aligned sources and sizes; anything else takes plain loads
    const int use_tma = (R % 4 == 0) && ((uintptr_t)verts % 16 == 0) && ((uintptr_t)xs % 16 == 0) &&
                        ((uintptr_t)ys % 16 == 0);
    dim3 grid(tiles_x * tiles_y, n_mesh);
    if (trows && (persp || target)) {
        dsf_set_error("the row-run target replaces the fp32 target and needs the non-perspective-correct rasteriser");
        return DSF_ERR_UNSUPPORTED;
    }
    auto kern = persp ? raster_fwd_kernel<true, false> : (trows ? raster_fwd_kernel<false, true> : raster_fwd_kernel<false, false>);
    kern<<<grid, RT_THREADS, smem, st>>>(R, tiles_x, verts, place_scale, place_off, h->faces_packed, h->face_order,
                                         h->n_faces, view, xs, ys, img, p2f, zbuf, bary, dists, target, thr,
                                         parts_tile, use_tma, crop ? *crop : CropParams{}, tail,
                                         trows ? *trows : TargetRows{});
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

// can the forward epilogue produce the vertex gradient itself (see raster_fwd_kernel)?  Needs the affine depth of
// the non-perspective-correct rasteriser, and its per-face / per-vertex tables must fit the candidate-list storage
// (true for the MANO mesh: 1554 faces).
bool dsf_raster_fused_grad_ok(const DsfMano* h, int flags) {
    const int Fp = (h->n_faces + 31) & ~31;
    return !(flags & (DSF_RASTER_PERSPECTIVE_CORRECT | DSF_RASTER_SEPARATE_BACKWARD)) &&
           320 + 3 * Fp + Fp / 2 + NVW * 3 <= (RT_THREADS / 32) * RT_WCANDS;
}

extern "C" int dsf_raster_forward(const DsfMano* h, int n_mesh, const float* verts_cam, const float* view,
                                  const float* xs, const float* ys, int R, float* img, int* pix_to_face,
                                  float* zbuf, float* bary, float* dists, const float* target, float thr,
                                  float* loss_parts_tile, int flags, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(h && verts_cam && view && xs && ys && img && pix_to_face, "null argument");
    DSF_REQUIRE(!target == !loss_parts_tile, "target and loss_parts_tile go together");
    DSF_REQUIRE(n_mesh > 0 && n_mesh <= 65535, "n_mesh must be in [1,65535] per call");
    DSF_REQUIRE(R >= 8 && R <= RT_MAXR, "crop size R must be in [8,512]");
    return dsf_raster_forward_impl(h, n_mesh, verts_cam, nullptr, nullptr, view, xs, ys, R, img, pix_to_face,
                                   zbuf, bary, dists, target, thr, loss_parts_tile, nullptr, flags, nullptr,
                                   (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// backward: one CTA per mesh; per foreground pixel the zbuf cotangent is pushed through the
// perspective-correct barycentrics to the three NDC vertices (shared-memory atomics), then through
// the projection to camera space.
// ------------------------------------------------------------------------------------------------
// pixels per pass (list entries are 16-bit offsets): 8192 for the 256-thread variant (16 KB list, 36 KB of
// shared memory per CTA, 6 CTAs per SM), 16384 for the 512-thread small-batch variant (one pass at R = 128)
#define RB_CHUNK_OF(threads) ((threads) == 256 ? 8192 : 16384)

// RB_THREADS: 256 keeps more CTAs (hands) in flight for large batches, 512 halves the per-hand latency
// when the batch does not fill the GPU anyway
template <int RB_THREADS, bool PERSP>
__global__ void __launch_bounds__(RB_THREADS, RB_THREADS == 256 ? 6 : 2)     // 40 registers: 6 CTAs (48 warps) per SM
raster_bwd_kernel(int R, const float* __restrict__ verts, const float* __restrict__ place_scale,
                  const float* __restrict__ place_off, const int* __restrict__ faces,
                  const float* __restrict__ view, const float* __restrict__ xs_g, const float* __restrict__ ys_g,
                  const int* __restrict__ p2f, const float* __restrict__ g_img, float* __restrict__ g_verts,
                  const float* __restrict__ target, const float* __restrict__ img,
                  const float* __restrict__ parts_tile, int n_tiles, float gscale, float thr, CropParams crop) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* svn = reinterpret_cast<float*>(smem_raw);
    float* sgn = svn + NVW * 3;
    float* sxs = sgn + NVW * 3;
    float* sys = sxs + R;
    const int mesh = blockIdx.x, tid = threadIdx.x;
    // NB float atomicAdd on shared memory is a compare-and-swap loop on this architecture (SASS
    // ATOMS.CAST.SPIN); accumulating with native global reductions (RED.E.ADD.F32 into the output
    // rows) was measured 1.9x slower (L2 atomic throughput), so the accumulator stays in shared
    // memory and the per-face warp reduction below keeps the number of atomics small.  (One 128-bit
    // CAS per vertex, ATOMS.CAS.128, instead of three 32-bit loops was measured 1.4x slower.)
    const ViewRec vw = load_view(view + (size_t)mesh * VIEW);
    for (int i = tid; i < R; i += RB_THREADS) {
        sxs[i] = xs_g[(size_t)mesh * R + i];
        sys[i] = ys_g[(size_t)mesh * R + i];
    }
    const float* vm = verts + (size_t)mesh * NVW * 3;
    const float* ps = place_scale ? place_scale + 3 * mesh : nullptr;
    const float* po = place_off ? place_off + 3 * mesh : nullptr;
    for (int v = tid; v < NVW; v += RB_THREADS) {
        project_vertex(vm + 3 * v, ps, po, vw, svn + 3 * v);
        sgn[3 * v] = 0.f; sgn[3 * v + 1] = 0.f; sgn[3 * v + 2] = 0.f;
    }
    __syncthreads();
    const float zmax = vw.zc + vw.zh, zmin_c = vw.zc - vw.zh;
    const float inv_zh = 1.f / vw.zh;
    const int* pf = p2f + (size_t)mesh * R * R;
    // cotangent of the image: given explicitly, or the m2d loss gradient recomputed on the fly:
    // d loss / d synth = gscale / (N_b + 1e-8) * sign(synth - real) on the union mask
    const float* gi = g_img ? g_img + (size_t)mesh * R * R : nullptr;
    const float* tg = target ? target + (size_t)mesh * R * R : nullptr;
    const float* im = img ? img + (size_t)mesh * R * R : nullptr;
    float n_mask = 0.f;                                    // mask count of the mesh = sum over its tiles
    if (parts_tile)
        for (int t = 0; t < n_tiles; ++t) n_mask += parts_tile[((size_t)mesh * n_tiles + t) * 2 + 1];
    const float gk = parts_tile ? gscale / (n_mask + 1e-8f) : 0.f;
    __shared__ CropBox cbox;
    const bool do_crop = !gi && crop.joints;
    if (do_crop && tid < 32) crop_box_warp(crop, mesh, place_off, place_scale, tid, &cbox);
    if (do_crop) __syncthreads();
    auto cotangent = [&](int k) -> float {
        if (gi) return gi[k];
        const float a = tg[k], c = im[k];
        // cropped-away pixels enter the loss as constant background: no gradient
        if (do_crop && !crop_keep(cbox, crop, k / R, k % R, R, c, vw.zc, vw.zh)) return 0.f;
        if (!(a < thr || c < thr)) return 0.f;
        const float d = c - a;
        return d > 0.f ? gk : (d < 0.f ? -gk : 0.f);
    };
    // Only ~20 % of the pixels are foreground.  Phase 1: the whole CTA streams pix_to_face with
    // 128-bit loads and appends the foreground pixel ids of a 16384-pixel chunk to a shared list
    // (warp-aggregated).  Phase 2: the list is consumed by full warps running the gradient body.
    unsigned short* s_list = reinterpret_cast<unsigned short*>(sys + R);
    __shared__ int s_count;
    const int lane = tid & 31;
    const int n_pix = R * R;
    const bool pf_vec = (reinterpret_cast<uintptr_t>(pf) & 15) == 0;   // odd R: planes after the first are unaligned
    constexpr int RB_CHUNK = RB_CHUNK_OF(RB_THREADS);
    for (int chunk0 = 0; chunk0 < n_pix; chunk0 += RB_CHUNK) {
      const int chunk_n = min(RB_CHUNK, n_pix - chunk0);
      if (tid == 0) s_count = 0;
      __syncthreads();
      for (int q0 = (tid & ~31) * 4; q0 < chunk_n; q0 += RB_THREADS * 4) {
          const int k0 = q0 + lane * 4;                          // 4 consecutive pixels per lane
          int4 f4 = make_int4(-1, -1, -1, -1);
          if (k0 + 3 < chunk_n && pf_vec) {
              f4 = *reinterpret_cast<const int4*>(pf + chunk0 + k0);   // plane 16-byte aligned (R*R % 4 == 0)
          } else {
              if (k0 < chunk_n) f4.x = pf[chunk0 + k0];
              if (k0 + 1 < chunk_n) f4.y = pf[chunk0 + k0 + 1];
              if (k0 + 2 < chunk_n) f4.z = pf[chunk0 + k0 + 2];
              if (k0 + 3 < chunk_n) f4.w = pf[chunk0 + k0 + 3];
          }
          const int n = (f4.x >= 0) + (f4.y >= 0) + (f4.z >= 0) + (f4.w >= 0);
          int total;
          const int excl = warp_excl_scan(n, lane, &total);
          int base = 0;
          if (lane == 0 && total > 0) base = atomicAdd(&s_count, total);
          base = __shfl_sync(0xffffffffu, base, 0);
          int slot = base + excl;
          if (f4.x >= 0) s_list[slot++] = (unsigned short)k0;
          if (f4.y >= 0) s_list[slot++] = (unsigned short)(k0 + 1);
          if (f4.z >= 0) s_list[slot++] = (unsigned short)(k0 + 2);
          if (f4.w >= 0) s_list[slot++] = (unsigned short)(k0 + 3);
      }
      __syncthreads();
      const int n_live = s_count;
      for (int eb = tid & ~31; eb < n_live; eb += RB_THREADS) {
        const int k = (eb + lane < n_live) ? chunk0 + (int)s_list[eb + lane] : -1;
        // every lane stays in the body (inactive ones carry zeros) so the warp can reduce per face
        bool act = k >= 0;
        const int kk = act ? k : 0;
        const int f = act ? pf[kk] : 0;
        const float g = act ? cotangent(kk) : 0.f;
        const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
        const float x0 = svn[3 * i0], y0 = svn[3 * i0 + 1], z0 = svn[3 * i0 + 2];
        const float x1 = svn[3 * i1], y1 = svn[3 * i1 + 1], z1 = svn[3 * i1 + 2];
        const float x2 = svn[3 * i2], y2 = svn[3 * i2 + 1], z2 = svn[3 * i2 + 2];
        const float px = sxs[kk % R], py = sys[kk / R];
        const float area = __fadd_rn(edge_rn(x2, y2, x0, y0, x1, y1), EPS);
        const float e0 = edge_rn(px, py, x1, y1, x2, y2);
        const float e1 = edge_rn(px, py, x2, y2, x0, y0);
        const float e2 = edge_rn(px, py, x0, y0, x1, y1);
        float gv[9];
        if (PERSP) {
            const float ia = 1.f / area;
            const float w0 = e0 * ia, w1 = e1 * ia, w2 = e2 * ia;
            const float t0 = w0 * z1 * z2, t1 = z0 * w1 * z2, t2 = z0 * z1 * w2;
            const float id = 1.f / (t0 + t1 + t2);
            const float b0 = t0 * id, b1 = t1 * id, b2 = t2 * id;
            const float pz = b0 * z0 + b1 * z1 + b2 * z2;
            // gates of the forward epilogue: background fill and the [zmin,zmax] clamp pass no gradient
            act = act && g != 0.f && (pz > 0.f) && !(pz > zmax) && !(pz < zmin_c);
            const float gz = act ? g * inv_zh : 0.f;
            const float gb0 = gz * z0, gb1 = gz * z1, gb2 = gz * z2;
            const float sgb = (gb0 * t0 + gb1 * t1 + gb2 * t2) * id * id;
            const float gt0 = gb0 * id - sgb, gt1 = gb1 * id - sgb, gt2 = gb2 * id - sgb;
            const float gw0 = gt0 * z1 * z2, gw1 = gt1 * z0 * z2, gw2 = gt2 * z0 * z1;
            const float ge0 = gw0 * ia, ge1 = gw1 * ia, ge2 = gw2 * ia;
            const float garea = -(gw0 * e0 + gw1 * e1 + gw2 * e2) * ia * ia;
            gv[0] = ge1 * (y2 - py) + ge2 * (py - y1) + garea * (y2 - y1);
            gv[1] = ge1 * (px - x2) + ge2 * (x1 - px) + garea * (x1 - x2);
            gv[2] = gz * b0 + gt1 * w1 * z2 + gt2 * z1 * w2;
            gv[3] = ge0 * (py - y2) + ge2 * (y0 - py) + garea * (y0 - y2);
            gv[4] = ge0 * (x2 - px) + ge2 * (px - x0) + garea * (x2 - x0);
            gv[5] = gz * b1 + gt0 * w0 * z2 + gt2 * z0 * w2;
            gv[6] = ge0 * (y1 - py) + ge1 * (py - y0) + garea * (y1 - y0);
            gv[7] = ge0 * (px - x1) + ge1 * (x0 - px) - garea * (x1 - x0);
            gv[8] = gz * b2 + gt0 * w0 * z1 + gt1 * z0 * w1;
        } else {
            // screen-space interpolation: the face's depth is affine in p, use the well-conditioned closed form
            // (face_grad) per pixel with S0 = gz, U = gz (p - v0)
            const float ia = 1.f / area;
            const float pz = (e0 * z0 + e1 * z1 + e2 * z2) * ia;
            act = act && g != 0.f && (pz > 0.f) && !(pz > zmax) && !(pz < zmin_c);
            const float gz = act ? g * inv_zh : 0.f;
            face_grad<float>(z0, x1 - x0, y1 - y0, z1 - z0, x2 - x0, y2 - y0, z2 - z0, gz, gz * (px - x0), gz * (py - y0), gv);
        }
        // Neighbouring pixels mostly hit the same face, and float atomics on shared memory are CAS
        // loops that serialise on equal addresses: sum each run of equal face ids inside the warp
        // (segmented scan) and let only the last lane of a run touch shared memory.
        const int key = act ? f : -1 - lane;
        const int key_prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = lane == 0 || key != key_prev;
        const unsigned int heads = __ballot_sync(0xffffffffu, head);
        const int run_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int e = 0; e < 9; ++e) {
                const float t = __shfl_up_sync(0xffffffffu, gv[e], o);
                if (lane - o >= run_start) gv[e] += t;
            }
        }
        const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
        if (act && tail) {
            atomicAdd(&sgn[3 * i0], gv[0]); atomicAdd(&sgn[3 * i0 + 1], gv[1]); atomicAdd(&sgn[3 * i0 + 2], gv[2]);
            atomicAdd(&sgn[3 * i1], gv[3]); atomicAdd(&sgn[3 * i1 + 1], gv[4]); atomicAdd(&sgn[3 * i1 + 2], gv[5]);
            atomicAdd(&sgn[3 * i2], gv[6]); atomicAdd(&sgn[3 * i2 + 1], gv[7]); atomicAdd(&sgn[3 * i2 + 2], gv[8]);
        }
      }
      __syncthreads();
    }
    float* go = g_verts + (size_t)mesh * NVW * 3;
    for (int v = tid; v < NVW; v += RB_THREADS) {
        float x = vm[3 * v], y = vm[3 * v + 1], z = vm[3 * v + 2];
        float sxp = 1.f, syp = 1.f, szp = 1.f;
        if (ps) {
            sxp = ps[0] * 0.5f; syp = ps[1] * 0.5f; szp = ps[2] * 0.5f;
            x = x * sxp + po[0]; y = y * syp + po[1]; z = z * szp + po[2];
        }
        const float gxn = sgn[3 * v], gyn = sgn[3 * v + 1], gzn = sgn[3 * v + 2];
        const float iz = 1.f / z;
        // x_ndc = -fxn x / z + pxn
        go[3 * v] = -gxn * vw.fxn * iz * sxp;
        go[3 * v + 1] = -gyn * vw.fyn * iz * syp;
        go[3 * v + 2] = (gzn + (gxn * vw.fxn * x + gyn * vw.fyn * y) * iz * iz) * szp;
    }
}

int dsf_raster_backward_impl(const DsfMano* h, int n_mesh, const float* verts, const float* place_scale,
                             const float* place_off, const float* view, const float* xs, const float* ys,
                             int R, const int* p2f, const float* g_img, float* g_verts, const float* target,
                             const float* img, const float* parts_tile, float gscale, float thr, const CropParams* crop,
                             int flags, cudaStream_t st) {
    const bool persp = (flags & DSF_RASTER_PERSPECTIVE_CORRECT) != 0;
    const int rb_threads = n_mesh < 2048 ? 512 : 256;
    const size_t smem = (size_t)NVW * 3 * 4 * 2 + (size_t)2 * R * 4 + (size_t)RB_CHUNK_OF(rb_threads) * 2;
    const int max_smem = (int)((size_t)NVW * 3 * 4 * 2 + (size_t)2 * RT_MAXR * 4 + (size_t)RB_CHUNK_OF(512) * 2);
    static bool attr_set[16] = {};
    int dev = 0;
    DSF_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 16 || !attr_set[dev]) {
        DSF_CHECK_CUDA(cudaFuncSetAttribute(raster_bwd_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        DSF_CHECK_CUDA(cudaFuncSetAttribute(raster_bwd_kernel<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        DSF_CHECK_CUDA(cudaFuncSetAttribute(raster_bwd_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        DSF_CHECK_CUDA(cudaFuncSetAttribute(raster_bwd_kernel<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        if (dev < 16) attr_set[dev] = true;
    }
    auto kern = rb_threads == 512 ? (persp ? raster_bwd_kernel<512, true> : raster_bwd_kernel<512, false>)
                                  : (persp ? raster_bwd_kernel<256, true> : raster_bwd_kernel<256, false>);
    kern<<<n_mesh, rb_threads, smem, st>>>(R, verts, place_scale, place_off, h->faces, view, xs, ys, p2f, g_img,
                                           g_verts, target, img, parts_tile, dsf_raster_tiles(R), gscale, thr,
                                           crop ? *crop : CropParams{});
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

extern "C" int dsf_raster_backward(const DsfMano* h, int n_mesh, const float* verts_cam, const float* view,
                                   const float* xs, const float* ys, int R, const int* pix_to_face,
                                   const float* g_img, float* g_verts_cam, int flags, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(h && verts_cam && view && xs && ys && pix_to_face && g_img && g_verts_cam, "null argument");
    DSF_REQUIRE(n_mesh > 0, "n_mesh must be positive");
    DSF_REQUIRE(R >= 8 && R <= RT_MAXR, "crop size R must be in [8,512]");
    return dsf_raster_backward_impl(h, n_mesh, verts_cam, nullptr, nullptr, view, xs, ys, R, pix_to_face,
                                    g_img, g_verts_cam, nullptr, nullptr, nullptr, 0.f, 0.f, nullptr, flags,
                                    (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// render losses
// ------------------------------------------------------------------------------------------------
#define LS_THREADS 256

__device__ __forceinline__ float2 block_sum2(float a, float b, float (*red)[2]) {
    a = warp_sum(a);
    b = warp_sum(b);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (lane == 0) { red[warp][0] = a; red[warp][1] = b; }
    __syncthreads();
    float sa = 0.f, sb = 0.f;
    for (int w = 0; w < LS_THREADS / 32; ++w) { sa += red[w][0]; sb += red[w][1]; }
    __syncthreads();
    return make_float2(sa, sb);
}

// mode 0: union mask (train_render.py:729), mode 1: intersection mask (render_loss.py:18)
__global__ void __launch_bounds__(LS_THREADS)
depth_loss_parts_kernel(int mode, int n, const float* __restrict__ real, const float* __restrict__ synth,
                        float thr, float gscale, float* __restrict__ parts, float* __restrict__ g_synth) {
    __shared__ float red[LS_THREADS / 32][2];
    const int b = blockIdx.x;
    const float* r = real + (size_t)b * n;
    const float* s = synth + (size_t)b * n;
    float sum = 0.f, cnt = 0.f;
    for (int i = threadIdx.x; i < n; i += LS_THREADS) {
        const float a = r[i], c = s[i];
        const bool m = mode == 0 ? (a < thr || c < thr) : (a < thr && c < thr);
        if (m) { sum += fabsf(a - c); cnt += 1.f; }
    }
    float2 t = block_sum2(sum, cnt, red);
    if (threadIdx.x == 0) { parts[2 * b] = t.x; parts[2 * b + 1] = t.y; }
    if (g_synth && mode == 0) {
        // loss = weight/B * sum_b  S_b / (N_b + 1e-8)
        const float k = gscale / (t.y + 1e-8f);
        float* g = g_synth + (size_t)b * n;
        for (int i = threadIdx.x; i < n; i += LS_THREADS) {
            const float a = r[i], c = s[i];
            const bool m = (a < thr || c < thr);
            const float d = c - a;
            g[i] = m ? (d > 0.f ? k : (d < 0.f ? -k : 0.f)) : 0.f;
        }
    }
}

__global__ void __launch_bounds__(LS_THREADS)
depth_loss_totals_kernel(int mode, int B, float weight, const float* __restrict__ parts,
                         float* __restrict__ totals) {
    __shared__ float red[LS_THREADS / 32][2];
    __shared__ float red2[LS_THREADS / 32][2];
    float sum = 0.f, cnt = 0.f, per = 0.f;
    for (int b = threadIdx.x; b < B; b += LS_THREADS) {
        sum += parts[2 * b];
        cnt += parts[2 * b + 1];
        per += parts[2 * b] / (parts[2 * b + 1] + 1e-8f);
    }
    float2 t = block_sum2(sum, cnt, red);
    float2 u = block_sum2(per, 0.f, red2);
    if (threadIdx.x == 0) {
        totals[0] = mode == 0 ? weight * u.x / (float)B : t.x / t.y;
        totals[1] = t.x;
        totals[2] = t.y;
        // un-normalised form for cross-rank reduction: global m2d loss = sum_r totals[3] / sum_r B_r
        totals[3] = mode == 0 ? weight * u.x : t.x;
    }
}

__global__ void __launch_bounds__(LS_THREADS)
depth_loss_grad_global_kernel(int n, const float* __restrict__ real, const float* __restrict__ synth, float thr,
                              const float* __restrict__ totals, float* __restrict__ g_synth) {
    const int b = blockIdx.x;
    const float k = 1.f / totals[2];
    for (int i = threadIdx.x; i < n; i += LS_THREADS) {
        const float a = real[(size_t)b * n + i], c = synth[(size_t)b * n + i];
        const bool m = a < thr && c < thr;
        const float d = c - a;
        g_synth[(size_t)b * n + i] = m ? (d > 0.f ? k : (d < 0.f ? -k : 0.f)) : 0.f;
    }
}

int dsf_depth_loss_impl(int mode, int B, int R, const float* real, const float* synth, float thr, float weight,
                        float* parts, float* totals, float* g_synth, cudaStream_t st) {
    const int n = R * R;
    depth_loss_parts_kernel<<<B, LS_THREADS, 0, st>>>(mode, n, real, synth, thr, weight / (float)B, parts, g_synth);
    DSF_CHECK_LAUNCH();
    depth_loss_totals_kernel<<<1, LS_THREADS, 0, st>>>(mode, B, weight, parts, totals);
    DSF_CHECK_LAUNCH();
    if (g_synth && mode == 1) {
        depth_loss_grad_global_kernel<<<B, LS_THREADS, 0, st>>>(n, real, synth, thr, totals, g_synth);
        DSF_CHECK_LAUNCH();
    }
    return DSF_OK;
}

// parts (B,2) from the per-tile sums + totals (4): one CTA, for the stand-alone dsf_raster_loss_grad
__global__ void __launch_bounds__(LS_THREADS)
fold_totals_kernel(int B, int n_tiles, float weight, const float* __restrict__ parts_tile, float* __restrict__ parts,
                   float* __restrict__ totals) {
    __shared__ float red[LS_THREADS / 32][2];
    __shared__ float red2[LS_THREADS / 32][2];
    float sum = 0.f, cnt = 0.f, per = 0.f;
    for (int b = threadIdx.x; b < B; b += LS_THREADS) {
        float a = 0.f, c = 0.f;
        for (int t = 0; t < n_tiles; ++t) {
            a += parts_tile[((size_t)b * n_tiles + t) * 2];
            c += parts_tile[((size_t)b * n_tiles + t) * 2 + 1];
        }
        parts[2 * b] = a; parts[2 * b + 1] = c;
        sum += a; cnt += c; per += a / (c + 1e-8f);
    }
    float2 t = block_sum2(sum, cnt, red);
    float2 u = block_sum2(per, 0.f, red2);
    if (threadIdx.x == 0) {
        totals[0] = weight * u.x / (float)B; totals[1] = t.x; totals[2] = t.y; totals[3] = weight * u.x;
    }
}

int dsf_fold_totals_impl(int B, int n_tiles, float weight, const float* parts_tile, float* parts, float* totals,
                         cudaStream_t st) {
    fold_totals_kernel<<<1, LS_THREADS, 0, st>>>(B, n_tiles, weight, parts_tile, parts, totals);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

// totals of a batch that was processed as n_slices separate dsf_fit_step calls (parallel streams):
// [3] (un-normalised loss), [1], [2] add up; [0] = sum [3] / norm_batch
__global__ void sum_totals_kernel(int n, const float* __restrict__ slice_totals, float norm, float* __restrict__ totals) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int i = 0; i < n; ++i) { a += slice_totals[4 * i + 1]; b += slice_totals[4 * i + 2]; c += slice_totals[4 * i + 3]; }
    totals[0] = c / norm; totals[1] = a; totals[2] = b; totals[3] = c;
}

extern "C" int dsf_sum_totals(int n_slices, const float* slice_totals, int norm_batch, float* totals,
                              dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(n_slices > 0 && slice_totals && totals && norm_batch > 0, "bad argument");
    sum_totals_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(n_slices, slice_totals, (float)norm_batch, totals);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

extern "C" int dsf_depth_loss(int mode, int batch, int R, const float* real, const float* synth, float thr,
                              float weight, float* parts, float* totals, float* g_synth, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (m2d union mask) or 1 (depth_loss)");
    DSF_REQUIRE(batch > 0 && R > 0 && real && synth && parts && totals, "null argument");
    return dsf_depth_loss_impl(mode, batch, R, real, synth, thr, weight, parts, totals, g_synth,
                               (cudaStream_t)stream);
}
