// dsf_b200 - depth crop -> point cloud for sm_100a ("next" row f1 of SURVEY.md section 8).
// Replaces loader.Img2pcl (data/render_loader.py:1121-1156: nearest resize, `<= 0.99` foreground
// mask, a Python loop over the batch with masked_select + multinomial) and loader.uvdImg2xyzImg
// (:1190-1200) with uvd_nl2xyznl_tensor / uvd_nl2xyz_tensor (:1044-1073), get_trans_points
// (:1113-1118) and pointsImgTo3D (:336-343).
//
// One CTA per hand, no host round trip: count the foreground pixels, choose `sample_num mod n`
// of them without replacement by taking the smallest counter-based hash keys (2-pass radix select
// in shared memory), and write [the whole foreground list repeated floor(sample_num / n) times |
// the chosen ones], both in pixel order, exactly the layout :1141-1153 produces.  The sampled
// subset is uniform and reproducible from `seed`; it is not torch.multinomial's stream (the
// consumers - ICP losses and seg_pcl - are permutation invariant).
#include <math.h>

#include "raster.cuh"

#define PCL_THREADS 512
#define PCL_WARPS (PCL_THREADS / 32)
#define PCL_BINS 4096

struct PclView {
    float mi[9];              // inverse of the crop transform M
    float cx, cy, cz;         // center3d
    float hx, hy, hz;         // cube / 2
    float fx, fy, px, py, flip;
    float half_img;           // loader.img_size / 2
};

__device__ __forceinline__ void pcl_view_load(PclView* v, int b, const float* center, const float* cube, const float* M,
                                              const float* intr4, float img_size, float flip) {
    const float* m = M + 9 * (size_t)b;
    const float a = m[0], bb = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const float c00 = e * i - f * h, c01 = f * g - d * i, c02 = d * h - e * g;
    const float det = a * c00 + bb * c01 + c * c02;
    const float r = 1.f / det;
    v->mi[0] = c00 * r; v->mi[1] = (c * h - bb * i) * r; v->mi[2] = (bb * f - c * e) * r;
    v->mi[3] = c01 * r; v->mi[4] = (a * i - c * g) * r;  v->mi[5] = (c * d - a * f) * r;
    v->mi[6] = c02 * r; v->mi[7] = (bb * g - a * h) * r; v->mi[8] = (a * e - bb * d) * r;
    v->cx = center[3 * b]; v->cy = center[3 * b + 1]; v->cz = center[3 * b + 2];
    v->hx = __fdiv_rn(cube[3 * b], 2.f); v->hy = __fdiv_rn(cube[3 * b + 1], 2.f); v->hz = __fdiv_rn(cube[3 * b + 2], 2.f);
    v->fx = intr4[0]; v->fy = intr4[1]; v->px = intr4[2]; v->py = intr4[3];
    v->flip = flip;
    v->half_img = __fdiv_rn(img_size, 2.f);
}

// normalised (u, v, d) of grid cell (row, col) -> camera-space mm (uvd_nl2xyz_tensor :1044-1057)
__device__ __forceinline__ void pcl_point(const PclView& v, int row, int col, int fs, float val, float* xyz) {
    const float fm1 = (float)fs - 1.f;
    const float gu = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, (float)col), fm1), 1.f);
    const float gv = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, (float)row), fm1), 1.f);
    const float uu = __fmul_rn(__fadd_rn(gu, 1.f), v.half_img), vv = __fmul_rn(__fadd_rn(gv, 1.f), v.half_img);
    const float d = __fadd_rn(__fmul_rn(val, v.hz), v.cz);
    const float us = __fadd_rn(__fadd_rn(__fmul_rn(v.mi[0], uu), __fmul_rn(v.mi[1], vv)), v.mi[2]);
    const float vs = __fadd_rn(__fadd_rn(__fmul_rn(v.mi[3], uu), __fmul_rn(v.mi[4], vv)), v.mi[5]);
    xyz[0] = __fdiv_rn(__fmul_rn(__fsub_rn(us, v.px), d), v.fx);
    xyz[1] = __fdiv_rn(__fmul_rn(__fmul_rn(v.flip, __fsub_rn(vs, v.py)), d), v.fy);
    xyz[2] = d;
}

__device__ __forceinline__ void pcl_normalise(const PclView& v, const float* xyz, float* out) {
    out[0] = __fdiv_rn(__fsub_rn(xyz[0], v.cx), v.hx);
    out[1] = __fdiv_rn(__fsub_rn(xyz[1], v.cy), v.hy);
    out[2] = __fdiv_rn(__fsub_rn(xyz[2], v.cz), v.hz);
}

// F.interpolate(mode="nearest"): source index = min(floor(dst * in / out), in - 1)
__device__ __forceinline__ int nearest_src(int dst, int n_in, float scale) {
    const int s = (int)floorf(__fmul_rn((float)dst, scale));
    return s < n_in - 1 ? s : n_in - 1;
}

// 24-bit draw key of (seed, hand, pixel): a 32-bit avalanche mixer; ties (about one per four hands at
// 3000 foreground pixels) are broken by pixel order
#define PCL_NOKEY 0xffffffffu
__device__ __forceinline__ unsigned pcl_key(unsigned long long seed, unsigned hand, unsigned pix) {
    unsigned x = pix + hand * 0x9E3779B9u + (unsigned)seed + (unsigned)(seed >> 32) * 0x85EBCA6Bu;
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x >> 8;
}

__device__ __forceinline__ int block_sum(int v, int* s_red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    for (int w = 0; w < PCL_WARPS; ++w) t += s_red[w];
    __syncthreads();                               // the scratch may be rewritten right after the call
    return t;
}

// CACHE: the keys of all cells fit in shared memory (feature_size <= 128), so the image is read once
// for the mask and once more for the kept cells; otherwise mask and key are recomputed in every pass.
template <bool CACHE>
__global__ void __launch_bounds__(PCL_THREADS)
img2pcl_kernel(int R_in, int fs, const float* __restrict__ img, const float* __restrict__ center,
               const float* __restrict__ cube, const float* __restrict__ M, float4 intr, float img_size, float flip,
               int sample_num, unsigned long long seed, int out_rows, float* __restrict__ pcl, int* __restrict__ count) {
    extern __shared__ __align__(16) unsigned int s_keys[];
    __shared__ PclView view;
    __shared__ int s_hist[PCL_BINS];
    __shared__ int s_red[PCL_WARPS];
    __shared__ int s_scan[3][PCL_WARPS];
    __shared__ int s_sel[2];          // chosen bin, items below it
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npix = fs * fs;
    const float* im = img + (size_t)b * R_in * R_in;
    const float scale = (float)R_in / (float)fs;
    if (tid == 0) {
        const float in4[4] = {intr.x, intr.y, intr.z, intr.w};
        pcl_view_load(&view, b, center, cube, M, in4, img_size, flip);
    }
    auto value = [&](int k) -> float {
        const int row = k / fs, col = k - row * fs;
        return R_in == fs ? im[k] : im[(size_t)nearest_src(row, R_in, scale) * R_in + nearest_src(col, R_in, scale)];
    };
    // key of cell k, PCL_NOKEY for background
    auto key_of = [&](int k) -> unsigned int {
        if (CACHE) return s_keys[k];
        return value(k) <= 0.99f ? pcl_key(seed, b, k) : PCL_NOKEY;
    };

    // every warp owns one contiguous segment of the cell sequence, so that ranks in pixel order are a
    // block-level prefix over 16 warp totals plus ballots inside the warp (no per-chunk block barrier)
    const int seg = ((npix + PCL_WARPS - 1) / PCL_WARPS + 31) & ~31;
    const int k_lo = warp * seg, k_hi = min(npix, k_lo + seg);
    int mine = 0;
    for (int k4 = k_lo + lane; k4 < k_hi; k4 += 128) {       // four independent loads in flight per lane
        float val[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) val[u] = k4 + 32 * u < k_hi ? value(k4 + 32 * u) : 1.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k4 + 32 * u;
            const bool v = val[u] <= 0.99f;
            if (CACHE && k < k_hi) s_keys[k] = v ? pcl_key(seed, b, k) : PCL_NOKEY;
            mine += v;
        }
    }
    const int n = block_sum(mine, s_red);          // also publishes `view` and the key cache
    if (tid == 0 && count) count[b] = n;
    float* out = pcl + (size_t)b * out_rows * 3;
    if (n == 0) {                                  // :1144 - an empty crop gives zeros
        for (int k = tid; k < out_rows * 3; k += PCL_THREADS) out[k] = 0.f;
        return;
    }
    const int mult = sample_num > 0 ? sample_num / n : 1;
    const int k_rem = sample_num > 0 ? sample_num - mult * n : 0;

    // k_rem-th smallest key among the foreground cells: radix select over 12 + 12 bits
    unsigned prefix = 0, prefix_mask = 0;
    int need = k_rem;                              // rank (1-based) still to be located
    if (k_rem > 0) {
        const int per = PCL_BINS / PCL_THREADS;    // consecutive bins per thread
        for (int pass = 0; pass < 2; ++pass) {
            const int shift = pass == 0 ? 12 : 0;
            for (int i = tid; i < PCL_BINS; i += PCL_THREADS) s_hist[i] = 0;
            __syncthreads();
            for (int k = tid; k < npix; k += PCL_THREADS) {
                const unsigned key = key_of(k);
                if (key != PCL_NOKEY && (key & prefix_mask) == prefix) atomicAdd(&s_hist[(key >> shift) & (PCL_BINS - 1)], 1);
            }
            __syncthreads();
            const int b0 = tid * per;
            int local = 0;
            for (int i = 0; i < per; ++i) local += s_hist[b0 + i];
            int incl = local;
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_scan[0][warp] = incl;
            __syncthreads();
            int before = incl - local;
            for (int w = 0; w < warp; ++w) before += s_scan[0][w];
            if (before < need && need <= before + local) {
                int cum = before;
                for (int i = 0; i < per; ++i) {
                    const int h = s_hist[b0 + i];
                    if (need <= cum + h) { s_sel[0] = b0 + i; s_sel[1] = cum; break; }
                    cum += h;
                }
            }
            __syncthreads();
            prefix |= (unsigned)s_sel[0] << shift;
            prefix_mask |= (unsigned)(PCL_BINS - 1) << shift;
            need -= s_sel[1];
            __syncthreads();
        }
    }
    const unsigned thr = prefix;                   // keys < thr are all taken, `need` of the keys == thr

    // ordered compaction: per-warp totals of (foreground, key < thr, key == thr) -> exclusive prefix over
    // the warps -> each warp walks its segment with ballots only
    int wv = 0, wl = 0, we = 0;
    for (int k = k_lo + lane; k < k_hi; k += 32) {
        const unsigned key = key_of(k);
        wv += key != PCL_NOKEY;
        wl += key != PCL_NOKEY && k_rem > 0 && key < thr;
        we += key != PCL_NOKEY && k_rem > 0 && key == thr;
    }
    for (int o = 16; o > 0; o >>= 1) {
        wv += __shfl_xor_sync(0xffffffffu, wv, o);
        wl += __shfl_xor_sync(0xffffffffu, wl, o);
        we += __shfl_xor_sync(0xffffffffu, we, o);
    }
    if (lane == 0) { s_scan[0][warp] = wv; s_scan[1][warp] = wl; s_scan[2][warp] = we; }
    __syncthreads();
    int base_valid = 0, base_less = 0, base_eq = 0;
    for (int w = 0; w < warp; ++w) { base_valid += s_scan[0][w]; base_less += s_scan[1][w]; base_eq += s_scan[2][w]; }
    const unsigned lower = (1u << lane) - 1u;
    for (int k4 = k_lo; k4 < k_hi; k4 += 128) {
        // keys and depths of four 32-cell chunks are fetched up front, then consumed in pixel order
        unsigned int key4[4];
        float val4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k4 + 32 * u + lane;
            key4[u] = k < k_hi ? key_of(k) : PCL_NOKEY;
            val4[u] = k < k_hi ? value(k) : 1.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = k4 + 32 * u + lane;
            const unsigned key = key4[u];
            const bool v = key != PCL_NOKEY;
            const bool less = v && k_rem > 0 && key < thr, eq = v && k_rem > 0 && key == thr;
            const unsigned mv = __ballot_sync(0xffffffffu, v), ml = __ballot_sync(0xffffffffu, less),
                           me = __ballot_sync(0xffffffffu, eq);
            if (v) {
                const int rv = base_valid + __popc(mv & lower), rl = base_less + __popc(ml & lower),
                          re = base_eq + __popc(me & lower);
                const int row = k / fs, col = k - row * fs;
                float xyz[3], q[3];
                pcl_point(view, row, col, fs, val4[u], xyz);
                pcl_normalise(view, xyz, q);
                for (int m = 0; m < mult; ++m) {
                    float* o = out + ((size_t)m * n + rv) * 3;
                    o[0] = q[0]; o[1] = q[1]; o[2] = q[2];
                }
                // selected rank in pixel order: all `less` so far plus the admitted ties so far
                const bool take = less || (eq && re < need);
                if (take) {
                    const int r = rl + (re < need ? re : need);
                    float* o = out + ((size_t)mult * n + r) * 3;
                    o[0] = q[0]; o[1] = q[1]; o[2] = q[2];
                }
            }
            base_valid += __popc(mv); base_less += __popc(ml); base_eq += __popc(me);
        }
    }
}

extern "C" int dsf_img2pcl(int batch, int R_in, int feature_size, const float* img, const float* center3d,
                           const float* cube, const float* M, const float* intr4, float img_size, float flip,
                           int sample_num, unsigned long long seed, float* pcl, int* count, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(batch > 0 && img && center3d && cube && M && intr4 && pcl, "null / empty argument");
    DSF_REQUIRE(R_in >= 2 && feature_size >= 2 && feature_size <= 1024, "feature_size must be in [2,1024]");
    DSF_REQUIRE(sample_num >= 0, "sample_num must be >= 0");
    const int out_rows = sample_num > 0 ? sample_num : feature_size * feature_size;
    const float4 in4 = make_float4(intr4[0], intr4[1], intr4[2], intr4[3]);
    if (feature_size <= 128) {
        const size_t smem = (size_t)feature_size * feature_size * sizeof(unsigned int);
        static bool attr_set[16] = {};          // the attribute is per device
        int dev = 0;
        DSF_CHECK_CUDA(cudaGetDevice(&dev));
        if (dev >= 16 || !attr_set[dev]) {
            DSF_CHECK_CUDA(cudaFuncSetAttribute(img2pcl_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 * 4));
            if (dev < 16) attr_set[dev] = true;
        }
        img2pcl_kernel<true><<<batch, PCL_THREADS, smem, (cudaStream_t)stream>>>(
            R_in, feature_size, img, center3d, cube, M, in4, img_size, flip, sample_num, seed, out_rows, pcl, count);
    } else {
        img2pcl_kernel<false><<<batch, PCL_THREADS, 0, (cudaStream_t)stream>>>(
            R_in, feature_size, img, center3d, cube, M, in4, img_size, flip, sample_num, seed, out_rows, pcl, count);
    }
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

// uvdImg2xyzImg (:1190-1200): per pixel camera-space xyz (mm) and the cube-normalised copy, both (B,3,R,R)
__global__ void __launch_bounds__(256)
uvd_img_to_xyz_kernel(int R, const float* __restrict__ img, const float* __restrict__ center,
                      const float* __restrict__ cube, const float* __restrict__ M, float4 intr, float img_size, float flip,
                      float* __restrict__ xyz_img, float* __restrict__ xyz_normal) {
    __shared__ PclView view;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) {
        const float in4[4] = {intr.x, intr.y, intr.z, intr.w};
        pcl_view_load(&view, b, center, cube, M, in4, img_size, flip);
    }
    __syncthreads();
    const int npix = R * R;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= npix) return;
    float xyz[3], q[3];
    pcl_point(view, k / R, k % R, R, img[(size_t)b * npix + k], xyz);
    pcl_normalise(view, xyz, q);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (xyz_img) xyz_img[((size_t)b * 3 + c) * npix + k] = xyz[c];
        if (xyz_normal) xyz_normal[((size_t)b * 3 + c) * npix + k] = q[c];
    }
}

extern "C" int dsf_uvd_img_to_xyz(int batch, int R, const float* img, const float* center3d, const float* cube,
                                  const float* M, const float* intr4, float img_size, float flip, float* xyz_img,
                                  float* xyz_normal, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(batch > 0 && R >= 2 && img && center3d && cube && M && intr4 && (xyz_img || xyz_normal),
                "null / empty argument");
    dim3 grid((R * R + 255) / 256, batch);
    uvd_img_to_xyz_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        R, img, center3d, cube, M, make_float4(intr4[0], intr4[1], intr4[2], intr4[3]), img_size, flip, xyz_img,
        xyz_normal);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

// ------------------------------------------------------------------------------------------------
// Sensor depth crop -> normalised target on the device.  The reference's loader crops the uint16
// millimetre depth map with INTER_NEAREST (data/render_loader.py:406,795), normalises it on the CPU
// (normalize_img :738-745) and ships 4 bytes per pixel to the GPU; here the crop travels as the
// sensor's own uint16 (half the PCIe bytes) and :738-745 runs in this kernel, same operation order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
target_from_u16_kernel(int npix, const unsigned short* __restrict__ depth, const float* __restrict__ center,
                       const float* __restrict__ cube, unsigned invalid, float* __restrict__ out) {
    const int b = blockIdx.y;
    const float cz = center[3 * b + 2], hz = __fdiv_rn(cube[3 * b + 2], 2.f);
    const float far_ = __fadd_rn(cz, hz), near_ = __fsub_rn(cz, hz);
    const unsigned short* d = depth + (size_t)b * npix;
    float* o = out + (size_t)b * npix;
    const int k = (blockIdx.x * 256 + threadIdx.x) * 8;
    if (k + 8 <= npix && ((((size_t)b * npix) & 7) == 0)) {
        const uint4 q = *reinterpret_cast<const uint4*>(d + k);
        const unsigned w[4] = {q.x, q.y, q.z, q.w};
        float r[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r[2 * i] = target_norm(w[i] & 0xffffu, invalid, cz, hz, far_, near_);
            r[2 * i + 1] = target_norm(w[i] >> 16, invalid, cz, hz, far_, near_);
        }
        *reinterpret_cast<float4*>(o + k) = make_float4(r[0], r[1], r[2], r[3]);
        *reinterpret_cast<float4*>(o + k + 4) = make_float4(r[4], r[5], r[6], r[7]);
    } else {
        for (int i = k; i < npix && i < k + 8; ++i) o[i] = target_norm(d[i], invalid, cz, hz, far_, near_);
    }
}

extern "C" int dsf_target_from_u16(int batch, int R, const unsigned short* depth_mm, const float* center3d,
                                   const float* cube, int invalid_value, float* target, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(batch > 0 && R >= 2 && depth_mm && center3d && cube && target, "null / empty argument");
    DSF_REQUIRE(invalid_value >= 0 && invalid_value <= 65535, "invalid_value must fit uint16 (0 = none)");
    DSF_REQUIRE((((size_t)depth_mm) & 15) == 0 && (((size_t)target) & 15) == 0, "buffers must be 16-byte aligned");
    const int npix = R * R;
    dim3 grid((npix + 2047) / 2048, batch);
    target_from_u16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(npix, depth_mm, center3d, cube, (unsigned)invalid_value,
                                                                     target);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

// ------------------------------------------------------------------------------------------------
// Row-run transport of the sensor crop ("data formats either side of the path"): ~80 % of a hand crop is
// background, so the loader can ship, per row, only the span between the first and the last pixel that is not
// background (depth 0 / invalid marker / at or beyond the far plane of the cube): rows (B,R,2) uint16 =
// (first column, length), hand_offset (B+1) uint32 = start of each hand's pixels in the packed uint16 payload.
// The kernel rebuilds the normalised fp32 target, bit-identical to dsf_target_from_u16 on the unpacked crop:
// inside a span the same target_norm, outside it the background value target_norm gives depth 0.
// One CTA per hand: warp-scan of the row lengths, then each warp writes whole rows (coalesced).
// ------------------------------------------------------------------------------------------------
#define RR_THREADS 256
#define RR_MAXR 512

__global__ void __launch_bounds__(RR_THREADS)
target_from_u16_rows_kernel(int R, const unsigned short* __restrict__ rows, const unsigned int* __restrict__ hand_offset,
                            const unsigned short* __restrict__ payload, const float* __restrict__ center,
                            const float* __restrict__ cube, unsigned invalid, float* __restrict__ out) {
    __shared__ unsigned int s_off[RR_MAXR + 1];
    __shared__ unsigned int s_warp[RR_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned short* rt = rows + (size_t)b * R * 2;
    // exclusive scan of the row lengths (R <= 512: two rows per thread)
    const int r0 = 2 * tid, r1 = 2 * tid + 1;
    const unsigned int l0 = r0 < R ? rt[2 * r0 + 1] : 0u, l1 = r1 < R ? rt[2 * r1 + 1] : 0u;
    unsigned int incl = l0 + l1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned int base = hand_offset[b];
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    const unsigned int ex = base + incl - (l0 + l1);
    if (r0 < R) s_off[r0] = ex;
    if (r1 < R) s_off[r1] = ex + l0;
    __syncthreads();
    const float cz = center[3 * b + 2], hz = __fdiv_rn(cube[3 * b + 2], 2.f);
    const float far_ = __fadd_rn(cz, hz), near_ = __fsub_rn(cz, hz);
    const float bg = target_norm(0u, invalid, cz, hz, far_, near_);
    float* o = out + (size_t)b * R * R;
    for (int r = warp; r < R; r += RR_THREADS / 32) {
        const int c0 = rt[2 * r], n = rt[2 * r + 1];
        const unsigned short* src = payload + s_off[r];
        for (int c = lane; c < R; c += 32) {
            const int k = c - c0;
            o[(size_t)r * R + c] = (k >= 0 && k < n) ? target_norm(src[k], invalid, cz, hz, far_, near_) : bg;
        }
    }
}

extern "C" int dsf_target_from_u16_rows(int batch, int R, const unsigned short* rows, const unsigned int* hand_offset,
                                        const unsigned short* payload, const float* center3d, const float* cube,
                                        int invalid_value, float* target, dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(batch > 0 && R >= 2 && R <= RR_MAXR && rows && hand_offset && payload && center3d && cube && target,
                "null / empty argument (R <= 512)");
    DSF_REQUIRE(invalid_value >= 0 && invalid_value <= 65535, "invalid_value must fit uint16 (0 = none)");
    target_from_u16_rows_kernel<<<batch, RR_THREADS, 0, (cudaStream_t)stream>>>(
        R, rows, hand_offset, payload, center3d, cube, (unsigned)invalid_value, target);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

// The loader's side of that format (host code, no device involved): pack (B,R,R) uint16 millimetre crops.
// rows (B,R,2), hand_offset (B+1), payload (capacity payload_cap pixels).  Returns the number of payload pixels,
// or -1 when the capacity is too small.  A pixel is background when it is 0, the invalid marker, or at / beyond the
// far plane center_z + cube_z / 2 evaluated in float32 exactly like the kernel does.
extern "C" long dsf_pack_u16_rows(int batch, int R, const unsigned short* depth_mm, const float* center3d,
                                  const float* cube, int invalid_value, unsigned short* rows,
                                  unsigned int* hand_offset, unsigned short* payload, long payload_cap) {
    if (batch <= 0 || R < 2 || R > RR_MAXR || !depth_mm || !center3d || !cube || !rows || !hand_offset || !payload)
        return -1;
    long n = 0;
    for (int b = 0; b < batch; ++b) {
        hand_offset[b] = (unsigned int)n;
        const volatile float hz = cube[3 * b + 2] / 2.f;            // float32 steps, as in target_norm
        const volatile float far_ = center3d[3 * b + 2] + hz;
        for (int r = 0; r < R; ++r) {
            const unsigned short* row = depth_mm + ((size_t)b * R + r) * R;
            int first = R, last = -1;
            for (int c = 0; c < R; ++c) {
                const unsigned u = row[c];
                const bool bgp = u == 0u || (invalid_value && u == (unsigned)invalid_value) || (float)u >= far_;
                if (!bgp) { if (first == R) first = c; last = c; }
            }
            const int len = last >= first ? last - first + 1 : 0;
            rows[((size_t)b * R + r) * 2] = (unsigned short)(len ? first : 0);
            rows[((size_t)b * R + r) * 2 + 1] = (unsigned short)len;
            if (n + len > payload_cap) return -1;
            for (int c = 0; c < len; ++c) payload[n + c] = row[first + c];
            n += len;
        }
    }
    hand_offset[batch] = (unsigned int)n;
    return n;
}
