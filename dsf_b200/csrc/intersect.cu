// dsf_b200 - intersection-volume metric for sm_100a (row I1 of SURVEY.md section 8).
// Replaces eval_coll.py:611-626 (self_intersection over the 15 watertight hand parts of
// get_part_mesh, :348-373) and util/intersect.py:102-107 (intersect_vox for an object / hand pair),
// which run trimesh on the CPU: surface-voxelise part t at `pitch` (midpoint subdivision until every
// edge <= pitch / 2, voxel = round-half-even(vertex / pitch)), count the voxel centres that lie inside
// part s (ray parity, forwards and backwards along a fixed direction), volume = count * pitch^3.
//
// Three launches per batch: (1) per hand: cap centres + part boxes in float64; (2) per (part t, hand):
// the voxel set as a bitmap over z-slabs of the part's box in shared memory, then every set voxel against every
// paired part s; (3) per hand: sum of the pair counts.  All geometry is float64 like the reference;
// the deciding arithmetic uses explicit _rn intrinsics in the oracle's operation order, so the counts
// are integers that match oracle/intersect_oracle.c exactly.
#include <math.h>

#include "common.cuh"

#define IV_THREADS 256
#define IV_MAX_PARTS 32
#ifndef IV_BITMAP_BYTES
#define IV_BITMAP_BYTES (24 * 1024)   // 196 k voxels per z-slab; small enough for 8 CTAs per SM
#endif
#define IV_MAX_LEVEL 10

struct d3 { double x, y, z; };

__device__ __forceinline__ double d_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double d_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double d_add(double a, double b) { return __dadd_rn(a, b); }

__device__ __forceinline__ d3 load3(const double* p) { d3 r = {p[0], p[1], p[2]}; return r; }

// same operation order as ray_tri() in oracle/intersect_oracle.c
__device__ __forceinline__ void ray_tri(const d3& p, const double* dir, const d3& a, const d3& b, const d3& c, int* fwd,
                                        int* bwd) {
    const double e1x = d_sub(b.x, a.x), e1y = d_sub(b.y, a.y), e1z = d_sub(b.z, a.z);
    const double e2x = d_sub(c.x, a.x), e2y = d_sub(c.y, a.y), e2z = d_sub(c.z, a.z);
    const double px = d_sub(d_mul(dir[1], e2z), d_mul(dir[2], e2y));
    const double py = d_sub(d_mul(dir[2], e2x), d_mul(dir[0], e2z));
    const double pz = d_sub(d_mul(dir[0], e2y), d_mul(dir[1], e2x));
    const double det = d_add(d_add(d_mul(e1x, px), d_mul(e1y, py)), d_mul(e1z, pz));
    if (det == 0.0) return;
    const double tx = d_sub(p.x, a.x), ty = d_sub(p.y, a.y), tz = d_sub(p.z, a.z);
    const double u = __ddiv_rn(d_add(d_add(d_mul(tx, px), d_mul(ty, py)), d_mul(tz, pz)), det);
    if (u < 0.0 || u > 1.0) return;
    const double qx = d_sub(d_mul(ty, e1z), d_mul(tz, e1y));
    const double qy = d_sub(d_mul(tz, e1x), d_mul(tx, e1z));
    const double qz = d_sub(d_mul(tx, e1y), d_mul(ty, e1x));
    const double v = __ddiv_rn(d_add(d_add(d_mul(dir[0], qx), d_mul(dir[1], qy)), d_mul(dir[2], qz)), det);
    if (v < 0.0 || d_add(u, v) > 1.0) return;
    const double t = __ddiv_rn(d_add(d_add(d_mul(e2x, qx), d_mul(e2y, qy)), d_mul(e2z, qz)), det);
    if (t > 0.0) ++*fwd;
    else if (t < 0.0) ++*bwd;
}

__constant__ double c_dir[2][3] = {{0.4395064455, 0.617598629942, 0.652231566745},
                                   {-0.617598629942, 0.652231566745, 0.4395064455}};

// (1) water mesh = verts (float32 -> float64) + cap centres; per part axis-aligned boxes
__global__ void __launch_bounds__(IV_THREADS)
ivox_prepare_kernel(int n_verts, const float* __restrict__ verts, int n_caps, const int* __restrict__ cap_ptr,
                    const int* __restrict__ cap_idx, int n_parts, const int* __restrict__ part_ptr,
                    const int* __restrict__ part_faces, double* __restrict__ wv_all, double* __restrict__ box_all) {
    const int b = blockIdx.x, tid = threadIdx.x, nw = n_verts + n_caps;
    double* wv = wv_all + (size_t)b * nw * 3;
    const float* v = verts + (size_t)b * n_verts * 3;
    for (int i = tid; i < n_verts * 3; i += IV_THREADS) wv[i] = (double)v[i];
    __syncthreads();
    for (int c = tid; c < n_caps; c += IV_THREADS) {
        double sx = 0, sy = 0, sz = 0;
        const int k0 = cap_ptr[c], k1 = cap_ptr[c + 1];
        for (int k = k0; k < k1; ++k) {
            const double* q = wv + 3 * cap_idx[k];
            sx = d_add(sx, q[0]); sy = d_add(sy, q[1]); sz = d_add(sz, q[2]);
        }
        const double n = (double)(k1 - k0);
        double* o = wv + 3 * (n_verts + c);
        o[0] = __ddiv_rn(sx, n); o[1] = __ddiv_rn(sy, n); o[2] = __ddiv_rn(sz, n);
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    for (int p = warp; p < n_parts; p += IV_THREADS / 32) {
        double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int e = 3 * part_ptr[p] + lane; e < 3 * part_ptr[p + 1]; e += 32) {
            const double* q = wv + 3 * part_faces[e];
#pragma unroll
            for (int c = 0; c < 3; ++c) { lo[c] = fmin(lo[c], q[c]); hi[c] = fmax(hi[c], q[c]); }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            for (int o = 16; o > 0; o >>= 1) {
                lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
                hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
            }
        if (lane == 0) {
            double* o = box_all + ((size_t)b * n_parts + p) * 6;
            o[0] = lo[0]; o[1] = lo[1]; o[2] = lo[2]; o[3] = hi[0]; o[4] = hi[1]; o[5] = hi[2];
        }
    }
}

// (2) one CTA per (part t, hand)
__global__ void __launch_bounds__(IV_THREADS)
ivox_count_kernel(int nw, int n_parts, const int* __restrict__ part_ptr, const int* __restrict__ part_faces,
                  const unsigned char* __restrict__ pair_mask, double pitch, const double* __restrict__ wv_all,
                  const double* __restrict__ box_all, long long* __restrict__ pair_counts,
                  long long* __restrict__ voxel_counts, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned int bitmap[];
    __shared__ int s_cnt[IV_MAX_PARTS];
    __shared__ int s_nvox, s_bad;
    const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* wv = wv_all + (size_t)b * nw * 3;
    const double* box = box_all + (size_t)b * n_parts * 6;
    long long* out = pair_counts + (size_t)b * n_parts * n_parts;
    if (tid < IV_MAX_PARTS) s_cnt[tid] = 0;
    if (tid == 0) { s_nvox = 0; s_bad = 0; }
    bool any = voxel_counts != nullptr;
    for (int s = 0; s < n_parts; ++s) any |= pair_mask[s * n_parts + t] != 0;
    if (!any) {
        for (int s = tid; s < n_parts; s += IV_THREADS) out[s * n_parts + t] = 0;
        return;
    }
    // voxel index range of this part: rounding is monotonic, so the box of the indices is the rounded box
    const double* bt = box + 6 * t;
    const long long ix0 = (long long)rint(__ddiv_rn(bt[0], pitch)), iy0 = (long long)rint(__ddiv_rn(bt[1], pitch)),
                    iz0 = (long long)rint(__ddiv_rn(bt[2], pitch));
    const long long nx = (long long)rint(__ddiv_rn(bt[3], pitch)) - ix0 + 1, ny = (long long)rint(__ddiv_rn(bt[4], pitch)) - iy0 + 1,
                    nz = (long long)rint(__ddiv_rn(bt[5], pitch)) - iz0 + 1;
    const int f0 = part_ptr[t], f1 = part_ptr[t + 1];
    // the bitmap covers a slab of z-layers at a time; parts whose box does not fit are done in several slabs
    const long long layer = nx * ny;
    const bool fits = f1 > f0 && nx > 0 && ny > 0 && nz > 0 && layer <= (long long)IV_BITMAP_BYTES * 8;
    if (!fits) {
        if (f1 > f0 && tid == 0) atomicOr(status + b, 1);      // one z-layer larger than the bitmap
        for (int s = tid; s < n_parts; s += IV_THREADS) out[s * n_parts + t] = 0;
        if (voxel_counts && tid == 0) voxel_counts[(size_t)b * n_parts + t] = 0;
        return;
    }
    const long long slab_nz = min(nz, ((long long)IV_BITMAP_BYTES * 8) / layer);
    const double max_edge = __ddiv_rn(pitch, 2.0);
    int my_vox = 0;
    for (long long z0 = 0; z0 < nz; z0 += slab_nz) {
        const long long z1 = min(nz, z0 + slab_nz);
        const int nwords = (int)((layer * (z1 - z0) + 31) / 32);
        __syncthreads();
        for (int i = tid; i < nwords; i += IV_THREADS) bitmap[i] = 0;
        __syncthreads();

        // voxelize_subdivide: a face whose longest edge is e ends up split 2^k ways, k minimal with
        // e / 2^k <= pitch / 2; its leaf vertices are the barycentric lattice points of order 2^k
        for (int f = f0 + warp; f < f1; f += IV_THREADS / 32) {
            const d3 a = load3(wv + 3 * part_faces[3 * f]), bb = load3(wv + 3 * part_faces[3 * f + 1]),
                     c = load3(wv + 3 * part_faces[3 * f + 2]);
            // skip faces that cannot reach this slab (one voxel of slack for the rounding)
            const double fz_lo = fmin(a.z, fmin(bb.z, c.z)), fz_hi = fmax(a.z, fmax(bb.z, c.z));
            if (rint(__ddiv_rn(fz_hi, pitch)) - (double)iz0 < (double)z0 || rint(__ddiv_rn(fz_lo, pitch)) - (double)iz0 >= (double)z1)
                continue;
            auto len = [](const d3& p, const d3& q) {
                const double x = d_sub(q.x, p.x), y = d_sub(q.y, p.y), z = d_sub(q.z, p.z);
                return sqrt(d_add(d_add(d_mul(x, x), d_mul(y, y)), d_mul(z, z)));
            };
            double e = fmax(len(a, bb), fmax(len(bb, c), len(c, a)));
            int k = 0;
            while (e > max_edge && k <= IV_MAX_LEVEL) { e *= 0.5; ++k; }
            if (k > IV_MAX_LEVEL) {                          // the reference raises 'max_iter exceeded'
                if (lane == 0) { atomicOr(status + b, 2); s_bad = 1; }
                continue;
            }
            const int n = 1 << k;
            const double dn = (double)n;
            for (int idx = lane; idx < (n + 1) * (n + 1); idx += 32) {
                const int i = idx / (n + 1), j = idx - i * (n + 1);
                if (i + j > n) continue;
                const double wa = (double)(n - i - j), wb = (double)i, wc = (double)j;
                const double x = __ddiv_rn(d_add(d_add(d_mul(a.x, wa), d_mul(bb.x, wb)), d_mul(c.x, wc)), dn);
                const double y = __ddiv_rn(d_add(d_add(d_mul(a.y, wa), d_mul(bb.y, wb)), d_mul(c.y, wc)), dn);
                const double z = __ddiv_rn(d_add(d_add(d_mul(a.z, wa), d_mul(bb.z, wb)), d_mul(c.z, wc)), dn);
                const long long vx = (long long)rint(__ddiv_rn(x, pitch)) - ix0, vy = (long long)rint(__ddiv_rn(y, pitch)) - iy0,
                                vz = (long long)rint(__ddiv_rn(z, pitch)) - iz0;
                if (vx < 0 || vx >= nx || vy < 0 || vy >= ny || vz < z0 || vz >= z1) continue;
                const long long lin = ((vz - z0) * ny + vy) * nx + vx;
                atomicOr(&bitmap[lin >> 5], 1u << (lin & 31));
            }
        }
        __syncthreads();

        // every occupied voxel centre against every paired part s
        for (int w = tid; w < nwords; w += IV_THREADS) {
            unsigned int bits = bitmap[w];
            my_vox += __popc(bits);
            while (bits) {
                const int bit = __ffs(bits) - 1;
                bits &= bits - 1;
                const long long lin = (long long)w * 32 + bit;
                const long long vx = lin % nx, vy = (lin / nx) % ny, vz = lin / layer + z0;
                const d3 p = {d_mul((double)(vx + ix0), pitch), d_mul((double)(vy + iy0), pitch), d_mul((double)(vz + iz0), pitch)};
                for (int s = 0; s < n_parts; ++s) {
                    if (!pair_mask[s * n_parts + t]) continue;
                    const double* bs = box + 6 * s;
                    if (p.x < bs[0] || p.x > bs[3] || p.y < bs[1] || p.y > bs[4] || p.z < bs[2] || p.z > bs[5]) continue;
                    int inside = 0;
                    for (int attempt = 0; attempt < 2; ++attempt) {
                        int fwd = 0, bwd = 0;
                        for (int f = part_ptr[s]; f < part_ptr[s + 1]; ++f)
                            ray_tri(p, c_dir[attempt], load3(wv + 3 * part_faces[3 * f]), load3(wv + 3 * part_faces[3 * f + 1]),
                                    load3(wv + 3 * part_faces[3 * f + 2]), &fwd, &bwd);
                        const int cf = fwd & 1, cb = bwd & 1;
                        if (cf == cb) { inside = cf; break; }
                        if (fwd == 0 || bwd == 0) break;
                    }
                    if (inside) atomicAdd(&s_cnt[s], 1);
                }
            }
        }
    }
    atomicAdd(&s_nvox, my_vox);
    __syncthreads();
    for (int s = tid; s < n_parts; s += IV_THREADS) out[s * n_parts + t] = s_bad ? 0 : (long long)s_cnt[s];
    if (voxel_counts && tid == 0) voxel_counts[(size_t)b * n_parts + t] = s_nvox;
}

// (3) volume = (sum of the pair counts) * pitch^3 ; -1 when a part could not be voxelised
__global__ void ivox_finish_kernel(int batch, int n_parts, const long long* __restrict__ pair_counts,
                                   const int* __restrict__ status, double pitch, double* __restrict__ volume) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    long long tot = 0;
    for (int i = 0; i < n_parts * n_parts; ++i) tot += pair_counts[(size_t)b * n_parts * n_parts + i];
    volume[b] = status[b] ? -1.0 : d_mul(d_mul(d_mul((double)tot, pitch), pitch), pitch);
}

extern "C" long dsf_intersect_workspace_bytes(int batch, int n_verts, int n_caps, int n_parts) {
    if (batch <= 0 || n_verts <= 0 || n_caps < 0 || n_parts <= 0) return -1;
    return (long)batch * ((long)(n_verts + n_caps) * 3 + (long)n_parts * 6) * (long)sizeof(double);
}

extern "C" int dsf_intersect_vox(int batch, int n_verts, const float* verts, int n_caps, const int* cap_ptr,
                                 const int* cap_idx, int n_parts, const int* part_ptr, const int* part_faces,
                                 const unsigned char* pair_mask, double pitch, long long* pair_counts,
                                 long long* voxel_counts, double* volume, int* status, void* workspace,
                                 dsfStream_t stream) {
    dsf_reset_launch_count();
    DSF_REQUIRE(batch > 0 && n_verts > 0 && verts && n_parts > 0 && part_ptr && part_faces && pair_mask && pair_counts &&
                    volume && status && workspace,
                "null / empty argument");
    DSF_REQUIRE(n_caps == 0 || (cap_ptr && cap_idx), "cap loops missing");
    DSF_REQUIRE(n_parts <= IV_MAX_PARTS, "at most 32 parts");
    DSF_REQUIRE(pitch > 0.0, "pitch must be positive");
    DSF_REQUIRE((((size_t)workspace) & 7) == 0, "workspace must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int nw = n_verts + n_caps;
    double* wv = reinterpret_cast<double*>(workspace);
    double* box = wv + (size_t)batch * nw * 3;
    static bool attr_set = false;
    if (!attr_set) {
        DSF_CHECK_CUDA(cudaFuncSetAttribute(ivox_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IV_BITMAP_BYTES));
        attr_set = true;
    }
    DSF_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int) * batch, st));
    ivox_prepare_kernel<<<batch, IV_THREADS, 0, st>>>(n_verts, verts, n_caps, cap_ptr, cap_idx, n_parts, part_ptr,
                                                      part_faces, wv, box);
    DSF_CHECK_LAUNCH();
    ivox_count_kernel<<<dim3(n_parts, batch), IV_THREADS, IV_BITMAP_BYTES, st>>>(
        nw, n_parts, part_ptr, part_faces, pair_mask, pitch, wv, box, pair_counts, voxel_counts, status);
    DSF_CHECK_LAUNCH();
    ivox_finish_kernel<<<(batch + 127) / 128, 128, 0, st>>>(batch, n_parts, pair_counts, status, pitch, volume);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}
