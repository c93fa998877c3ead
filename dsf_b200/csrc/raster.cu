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

#include "common.cuh"

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
        if (mode == 0) {
            float half = (float)R * 0.5f;
            float fxc = __fmul_rn(sM[0], fx), fyc = __fmul_rn(sM[2], fy);
            float pxc = __fadd_rn(__fmul_rn(sM[0], px), sM[1]), pyc = __fadd_rn(__fmul_rn(sM[2], py), sM[3]);
            vw[0] = __fdiv_rn(fxc, half);
            vw[1] = __fdiv_rn(fyc, half);
            vw[2] = -__fdiv_rn(__fsub_rn(pxc, half), half);
            vw[3] = -__fdiv_rn(__fsub_rn(pyc, half), half);
            vw[7] = -half; vw[8] = ((float)R - 1.f) * 0.5f;
            vw[9] = -half; vw[10] = ((float)R - 1.f) * 0.5f;
            vw[11] = 0.f; vw[12] = (float)(R - 1); vw[13] = 0.f; vw[14] = (float)(R - 1);
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
    if (mode == 0) {
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
    DSF_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (direct) or 1 (literal)");
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
};

__device__ __forceinline__ ViewRec load_view(const float* v) {
    ViewRec r;
    r.fxn = v[0]; r.fyn = v[1]; r.pxn = v[2]; r.pyn = v[3]; r.zc = v[4]; r.zh = v[5]; r.bg = v[6];
    r.ax = v[7]; r.bx = v[8]; r.ay = v[9]; r.by = v[10];
    r.xlo = (int)v[11]; r.xhi = (int)v[12]; r.ylo = (int)v[13]; r.yhi = (int)v[14];
    return r;
}

// camera-space vertex -> (x_ndc, y_ndc, z): world->view flip R = diag(-1,-1,1) (mano_layer.py:935-938)
// then the pytorch3d screen-space calibration.  place != null applies verts*cube/2 + center first
// (mano_layer.py:1078).
__device__ __forceinline__ void project_vertex(const float* v, const float* place_scale, const float* place_off,
                                               const ViewRec& vw, float* out) {
    float x = v[0], y = v[1], z = v[2];
    if (place_scale) {
        x = __fadd_rn(__fdiv_rn(__fmul_rn(x, place_scale[0]), 2.f), place_off[0]);
        y = __fadd_rn(__fdiv_rn(__fmul_rn(y, place_scale[1]), 2.f), place_off[1]);
        z = __fadd_rn(__fdiv_rn(__fmul_rn(z, place_scale[2]), 2.f), place_off[2]);
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

// full oracle-order evaluation of one (pixel, face) pair
__device__ __forceinline__ FragEval eval_fragment(float px, float py, float x0, float y0, float z0, float x1,
                                                  float y1, float z1, float x2, float y2, float z2, float area) {
    FragEval r;
    float e0 = edge_rn(px, py, x1, y1, x2, y2);
    float e1 = edge_rn(px, py, x2, y2, x0, y0);
    float e2 = edge_rn(px, py, x0, y0, x1, y1);
    r.w0 = __fdiv_rn(e0, area);
    r.w1 = __fdiv_rn(e1, area);
    r.w2 = __fdiv_rn(e2, area);
    float t0 = __fmul_rn(__fmul_rn(r.w0, z1), z2);
    float t1 = __fmul_rn(__fmul_rn(z0, r.w1), z2);
    float t2 = __fmul_rn(__fmul_rn(z0, z1), r.w2);
    float den = __fadd_rn(__fadd_rn(t0, t1), t2);
    r.b0 = __fdiv_rn(t0, den);
    r.b1 = __fdiv_rn(t1, den);
    r.b2 = __fdiv_rn(t2, den);
    r.pz = __fadd_rn(__fadd_rn(__fmul_rn(r.b0, z0), __fmul_rn(r.b1, z1)), __fmul_rn(r.b2, z2));
    r.ok = !(r.pz < 0.f) && r.b0 > 0.f && r.b1 > 0.f && r.b2 > 0.f;
    return r;
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
// forward: grid (tiles, meshes); each CTA owns a TILE x TILE pixel block of one mesh
// ------------------------------------------------------------------------------------------------
#define RT_TILE 64
#define RT_THREADS 256
#define RT_MAXR 512

__global__ void __launch_bounds__(RT_THREADS)
raster_fwd_kernel(int R, int tiles_x, const float* __restrict__ verts, const float* __restrict__ place_scale,
                  const float* __restrict__ place_off, const int* __restrict__ faces, int F,
                  const float* __restrict__ view, const float* __restrict__ xs_g, const float* __restrict__ ys_g,
                  float* __restrict__ img, int* __restrict__ p2f, float* __restrict__ zbuf,
                  float* __restrict__ bary, float* __restrict__ dists) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(smem_raw);
    float* svn = reinterpret_cast<float*>(skey + RT_TILE * RT_TILE);
    float* sxs = svn + NVW * 3;
    float* sys = sxs + R;
    const int mesh = blockIdx.y;
    const int tile = blockIdx.x;
    const int tx0 = (tile % tiles_x) * RT_TILE, ty0 = (tile / tiles_x) * RT_TILE;
    const int tx1 = min(R, tx0 + RT_TILE) - 1, ty1 = min(R, ty0 + RT_TILE) - 1;
    const int tid = threadIdx.x;
    const ViewRec vw = load_view(view + (size_t)mesh * VIEW);

    for (int i = tid; i < R; i += RT_THREADS) {
        sxs[i] = xs_g[(size_t)mesh * R + i];
        sys[i] = ys_g[(size_t)mesh * R + i];
    }
    const float* vm = verts + (size_t)mesh * NVW * 3;
    const float* ps = place_scale ? place_scale + 3 * mesh : nullptr;
    const float* po = place_off ? place_off + 3 * mesh : nullptr;
    for (int v = tid; v < NVW; v += RT_THREADS) project_vertex(vm + 3 * v, ps, po, vw, svn + 3 * v);
    for (int i = tid; i < RT_TILE * RT_TILE; i += RT_THREADS) skey[i] = ~0ull;
    __syncthreads();

    const int cx0 = max(tx0, vw.xlo), cx1 = min(tx1, vw.xhi);
    const int cy0 = max(ty0, vw.ylo), cy1 = min(ty1, vw.yhi);
    if (cx0 <= cx1 && cy0 <= cy1) {
        for (int f = tid; f < F; f += RT_THREADS) {
            const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
            const float x0 = svn[3 * i0], y0 = svn[3 * i0 + 1], z0 = svn[3 * i0 + 2];
            const float x1 = svn[3 * i1], y1 = svn[3 * i1 + 1], z1 = svn[3 * i1 + 2];
            const float x2 = svn[3 * i2], y2 = svn[3 * i2 + 1], z2 = svn[3 * i2 + 2];
            const float zmin = fminf(z0, fminf(z1, z2));
            if (!(zmin >= EPS)) continue;                       // behind / at the camera
            const float xmin = fminf(x0, fminf(x1, x2)), xmax = fmaxf(x0, fmaxf(x1, x2));
            const float ymin = fminf(y0, fminf(y1, y2)), ymax = fmaxf(y0, fmaxf(y1, y2));
            const int ia = max(cx0, first_le(sxs, vw.xlo, vw.xhi, vw.ax, vw.bx, xmax));
            const int ib = min(cx1, last_ge(sxs, vw.xlo, vw.xhi, vw.ax, vw.bx, xmin));
            if (ia > ib) continue;
            const int ja = max(cy0, first_le(sys, vw.ylo, vw.yhi, vw.ay, vw.by, ymax));
            const int jb = min(cy1, last_ge(sys, vw.ylo, vw.yhi, vw.ay, vw.by, ymin));
            if (ja > jb) continue;
            const float farea = edge_rn(x0, y0, x1, y1, x2, y2);
            if (farea <= EPS && farea >= -EPS) continue;        // degenerate in NDC
            const float area = __fadd_rn(edge_rn(x2, y2, x0, y0, x1, y1), EPS);
            // per-face edge deltas (same single roundings the oracle performs inline)
            const float d0y = __fsub_rn(y2, y1), d0x = __fsub_rn(x2, x1);
            const float d1y = __fsub_rn(y0, y2), d1x = __fsub_rn(x0, x2);
            const float d2y = __fsub_rn(y1, y0), d2x = __fsub_rn(x1, x0);
            for (int j = ja; j <= jb; ++j) {
                const float py = sys[j];
                const float r0 = __fmul_rn(__fsub_rn(py, y1), d0x);
                const float r1 = __fmul_rn(__fsub_rn(py, y2), d1x);
                const float r2 = __fmul_rn(__fsub_rn(py, y0), d2x);
                for (int i = ia; i <= ib; ++i) {
                    const float px = sxs[i];
                    const float e0 = __fsub_rn(__fmul_rn(__fsub_rn(px, x1), d0y), r0);
                    const float e1 = __fsub_rn(__fmul_rn(__fsub_rn(px, x2), d1y), r1);
                    const float e2 = __fsub_rn(__fmul_rn(__fsub_rn(px, x0), d2y), r2);
                    const bool pos = e0 > 0.f && e1 > 0.f && e2 > 0.f;
                    const bool neg = e0 < 0.f && e1 < 0.f && e2 < 0.f;
                    if (!(pos || neg)) continue;                // some w_i <= 0: cannot be inside
                    FragEval fe = eval_fragment(px, py, x0, y0, z0, x1, y1, z1, x2, y2, z2, area);
                    if (!fe.ok) continue;
                    const float pz = fe.pz + 0.f;               // -0 -> +0 so the bit pattern orders
                    const unsigned long long key =
                        ((unsigned long long)__float_as_uint(pz) << 32) | (unsigned int)f;
                    atomicMin(&skey[(j - ty0) * RT_TILE + (i - tx0)], key);
                }
            }
        }
    }
    __syncthreads();

    // epilogue: background fill (:1084-1085) + normalize_img (:1289-1299)
    const float zmax = __fadd_rn(vw.zc, vw.zh), zmin_c = __fsub_rn(vw.zc, vw.zh);
    const int tw = tx1 - tx0 + 1, th = ty1 - ty0 + 1;
    for (int k = tid; k < tw * th; k += RT_THREADS) {
        const int lx = k % tw, ly = k / tw;
        const unsigned long long key = skey[ly * RT_TILE + lx];
        const int f = key == ~0ull ? -1 : (int)(unsigned int)(key & 0xffffffffu);
        const float z = f < 0 ? -1.f : __uint_as_float((unsigned int)(key >> 32));
        float d = z <= 0.f ? 0.f : z;
        d = (d == 0.f) ? zmax : d;
        d = d > zmax ? zmax : d;
        d = d < zmin_c ? zmin_c : d;
        const size_t o = ((size_t)mesh * R + (ty0 + ly)) * R + (tx0 + lx);
        img[o] = __fdiv_rn(__fsub_rn(d, vw.zc), vw.zh);
        p2f[o] = f;
        if (zbuf) zbuf[o] = z;
        if (bary || dists) {
            float b0 = -1.f, b1 = -1.f, b2 = -1.f, dd = -1.f;
            if (f >= 0) {
                const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
                const float x0 = svn[3 * i0], y0 = svn[3 * i0 + 1], z0 = svn[3 * i0 + 2];
                const float x1 = svn[3 * i1], y1 = svn[3 * i1 + 1], z1 = svn[3 * i1 + 2];
                const float x2 = svn[3 * i2], y2 = svn[3 * i2 + 1], z2 = svn[3 * i2 + 2];
                const float px = sxs[tx0 + lx], py = sys[ty0 + ly];
                const float area = __fadd_rn(edge_rn(x2, y2, x0, y0, x1, y1), EPS);
                FragEval fe = eval_fragment(px, py, x0, y0, z0, x1, y1, z1, x2, y2, z2, area);
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

static size_t raster_fwd_smem(int R) {
    return (size_t)RT_TILE * RT_TILE * 8 + (size_t)NVW * 3 * 4 + (size_t)2 * R * 4;
}

int dsf_raster_forward_impl(const DsfMano* h, int n_mesh, const float* verts, const float* place_scale,
                            const float* place_off, const float* view, const float* xs, const float* ys,
                            int R, float* img, int* p2f, float* zbuf, float* bary, float* dists,
                            cudaStream_t st) {
    const int tiles_x = (R + RT_TILE - 1) / RT_TILE;
    const size_t smem = raster_fwd_smem(R);
    static bool attr_set[16] = {};
    int dev = 0;
    DSF_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 16 || !attr_set[dev]) {
        DSF_CHECK_CUDA(cudaFuncSetAttribute(raster_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)raster_fwd_smem(RT_MAXR)));
        if (dev < 16) attr_set[dev] = true;
    }
    dim3 grid(tiles_x * tiles_x, n_mesh);
    raster_fwd_kernel<<<grid, RT_THREADS, smem, st>>>(R, tiles_x, verts, place_scale, place_off, h->faces,
                                                      h->n_faces, view, xs, ys, img, p2f, zbuf, bary, dists);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

extern "C" int dsf_raster_forward(const DsfMano* h, int n_mesh, const float* verts_cam, const float* view,
                                  const float* xs, const float* ys, int R, float* img, int* pix_to_face,
                                  float* zbuf, float* bary, float* dists, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(h && verts_cam && view && xs && ys && img && pix_to_face, "null argument");
    DSF_REQUIRE(n_mesh > 0 && n_mesh <= 65535, "n_mesh must be in [1,65535] per call");
    DSF_REQUIRE(R >= 8 && R <= RT_MAXR, "crop size R must be in [8,512]");
    return dsf_raster_forward_impl(h, n_mesh, verts_cam, nullptr, nullptr, view, xs, ys, R, img, pix_to_face,
                                   zbuf, bary, dists, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// backward: one CTA per mesh; per foreground pixel the zbuf cotangent is pushed through the
// perspective-correct barycentrics to the three NDC vertices (shared-memory atomics), then through
// the projection to camera space.
// ------------------------------------------------------------------------------------------------
#define RB_THREADS 256

__global__ void __launch_bounds__(RB_THREADS)
raster_bwd_kernel(int R, const float* __restrict__ verts, const float* __restrict__ place_scale,
                  const float* __restrict__ place_off, const int* __restrict__ faces,
                  const float* __restrict__ view, const float* __restrict__ xs_g, const float* __restrict__ ys_g,
                  const int* __restrict__ p2f, const float* __restrict__ g_img, float* __restrict__ g_verts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* svn = reinterpret_cast<float*>(smem_raw);
    float* sgn = svn + NVW * 3;
    float* sxs = sgn + NVW * 3;
    float* sys = sxs + R;
    const int mesh = blockIdx.x, tid = threadIdx.x;
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
    const float* gi = g_img + (size_t)mesh * R * R;
    for (int k = tid; k < R * R; k += RB_THREADS) {
        const int f = pf[k];
        if (f < 0) continue;
        const float g = gi[k];
        if (g == 0.f) continue;
        const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
        const float x0 = svn[3 * i0], y0 = svn[3 * i0 + 1], z0 = svn[3 * i0 + 2];
        const float x1 = svn[3 * i1], y1 = svn[3 * i1 + 1], z1 = svn[3 * i1 + 2];
        const float x2 = svn[3 * i2], y2 = svn[3 * i2 + 1], z2 = svn[3 * i2 + 2];
        const float px = sxs[k % R], py = sys[k / R];
        const float area = __fadd_rn(edge_rn(x2, y2, x0, y0, x1, y1), EPS);
        const float e0 = edge_rn(px, py, x1, y1, x2, y2);
        const float e1 = edge_rn(px, py, x2, y2, x0, y0);
        const float e2 = edge_rn(px, py, x0, y0, x1, y1);
        const float ia = 1.f / area;
        const float w0 = e0 * ia, w1 = e1 * ia, w2 = e2 * ia;
        const float t0 = w0 * z1 * z2, t1 = z0 * w1 * z2, t2 = z0 * z1 * w2;
        const float den = t0 + t1 + t2;
        const float id = 1.f / den;
        const float b0 = t0 * id, b1 = t1 * id, b2 = t2 * id;
        const float pz = b0 * z0 + b1 * z1 + b2 * z2;
        // gates of the forward epilogue: background fill and the [zmin,zmax] clamp pass no gradient
        if (!(pz > 0.f) || pz > zmax || pz < zmin_c) continue;
        const float gz = g * inv_zh;
        const float gb0 = gz * z0, gb1 = gz * z1, gb2 = gz * z2;
        const float s = (gb0 * t0 + gb1 * t1 + gb2 * t2) * id * id;
        const float gt0 = gb0 * id - s, gt1 = gb1 * id - s, gt2 = gb2 * id - s;
        const float gw0 = gt0 * z1 * z2, gw1 = gt1 * z0 * z2, gw2 = gt2 * z0 * z1;
        const float gz0 = gz * b0 + gt1 * w1 * z2 + gt2 * z1 * w2;
        const float gz1 = gz * b1 + gt0 * w0 * z2 + gt2 * z0 * w2;
        const float gz2 = gz * b2 + gt0 * w0 * z1 + gt1 * z0 * w1;
        const float ge0 = gw0 * ia, ge1 = gw1 * ia, ge2 = gw2 * ia;
        const float garea = -(gw0 * e0 + gw1 * e1 + gw2 * e2) * ia * ia;
        float gx0 = ge1 * (y2 - py) + ge2 * (py - y1) + garea * (y2 - y1);
        float gy0 = ge1 * (px - x2) + ge2 * (x1 - px) + garea * (x1 - x2);
        float gx1 = ge0 * (py - y2) + ge2 * (y0 - py) + garea * (y0 - y2);
        float gy1 = ge0 * (x2 - px) + ge2 * (px - x0) + garea * (x2 - x0);
        float gx2 = ge0 * (y1 - py) + ge1 * (py - y0) + garea * (y1 - y0);
        float gy2 = ge0 * (px - x1) + ge1 * (x0 - px) - garea * (x1 - x0);
        atomicAdd(&sgn[3 * i0], gx0); atomicAdd(&sgn[3 * i0 + 1], gy0); atomicAdd(&sgn[3 * i0 + 2], gz0);
        atomicAdd(&sgn[3 * i1], gx1); atomicAdd(&sgn[3 * i1 + 1], gy1); atomicAdd(&sgn[3 * i1 + 2], gz1);
        atomicAdd(&sgn[3 * i2], gx2); atomicAdd(&sgn[3 * i2 + 1], gy2); atomicAdd(&sgn[3 * i2 + 2], gz2);
    }
    __syncthreads();
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
                             int R, const int* p2f, const float* g_img, float* g_verts, cudaStream_t st) {
    const size_t smem = (size_t)NVW * 3 * 4 * 2 + (size_t)2 * R * 4;
    raster_bwd_kernel<<<n_mesh, RB_THREADS, smem, st>>>(R, verts, place_scale, place_off, h->faces, view, xs,
                                                        ys, p2f, g_img, g_verts);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

extern "C" int dsf_raster_backward(const DsfMano* h, int n_mesh, const float* verts_cam, const float* view,
                                   const float* xs, const float* ys, int R, const int* pix_to_face,
                                   const float* g_img, float* g_verts_cam, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(h && verts_cam && view && xs && ys && pix_to_face && g_img && g_verts_cam, "null argument");
    DSF_REQUIRE(n_mesh > 0, "n_mesh must be positive");
    DSF_REQUIRE(R >= 8 && R <= RT_MAXR, "crop size R must be in [8,512]");
    return dsf_raster_backward_impl(h, n_mesh, verts_cam, nullptr, nullptr, view, xs, ys, R, pix_to_face,
                                    g_img, g_verts_cam, (cudaStream_t)stream);
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
        totals[3] = 0.f;
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

extern "C" int dsf_depth_loss(int mode, int batch, int R, const float* real, const float* synth, float thr,
                              float weight, float* parts, float* totals, float* g_synth, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (m2d union mask) or 1 (depth_loss)");
    DSF_REQUIRE(batch > 0 && R > 0 && real && synth && parts && totals, "null argument");
    return dsf_depth_loss_impl(mode, batch, R, real, synth, thr, weight, parts, totals, g_synth,
                               (cudaStream_t)stream);
}
