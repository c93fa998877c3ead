// dsf_b200 - blend-shape contraction on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   forward   v_posed(B x 2336) = [beta | Rs - I](B x 148) . [shapedirs ; posedirs](148 x 2336) + v_template
//             (render_model/mano_layer.py:586 and :613 as one GEMM)
//   backward  g_X(B x 148)      = g_vposed(B x 2336) . basis^T
//
// Vertices must match the fp32 reference to 1e-5, which single-pass TF32 (10-bit mantissa) cannot
// give, so every product is computed error-compensated ("3xTF32"):
//       a.b  ~=  a_hi.b_hi + a_hi.b_lo + a_lo.b_hi,     a_hi = a with the low 13 mantissa bits cleared
// The basis is split into hi/lo once on the host; the activations arrive pre-split from their producers (the pose
// kernel writes [beta | Rs - I], the skinning backward writes g_vposed, both as hi and lo rows of the workspace).
// Both directions run through ONE kernel, tf32x3_gemm_tma_kernel: a 128-row tile per CTA, tcgen05.mma (kind::tf32,
// M = 128, N = 128 forward / 160 backward, K = 8) issued by a single thread, fp32 accumulator in TMEM, operand tiles
// streamed by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) through a three-stage ring guarded by full / empty
// mbarriers, split-K over grid.z for the backward contraction, epilogue TMEM -> registers (+ bias) -> shared memory
// -> coalesced row segments.  tf32x3_gemm_kernel is the cp.async fallback for drivers without the tensor-map entry
// point (canonical no-swizzle K-major core-matrix layout).
#include <cuda.h>

#include "common.cuh"

#define GM 128            // rows (hands) per CTA = UMMA M
#define GKB 32            // k elements per stage (one 128-byte row of fp32)
#define G_THREADS 128

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: 8 x 16-byte core matrices; LBO = 128 B between K-adjacent core matrices,
// SBO = 1024 B between 8-row groups (cute::UMMA::SmemDescriptor, version 1 = Blackwell)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// float index of element (row r, k) inside an operand tile of GKB columns
__device__ __forceinline__ int tile_idx(int r, int k) { return ((r >> 3) * (GKB / 4) + (k >> 2)) * 32 + (r & 7) * 4 + (k & 3); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ void issue_mma(uint32_t tmem, uint32_t a_base, uint32_t b_base, uint32_t idesc, bool first) {
#pragma unroll
    for (int j = 0; j < GKB / 8; ++j) {
        const uint64_t da = make_smem_desc(a_base + j * 256), db = make_smem_desc(b_base + j * 256);
        const uint32_t acc = (first && j == 0) ? 0u : 1u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem),
            "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u)
            : "memory");
    }
}

// cp.async fallback (no tensor maps): the activation arrives already split (A = hi part, Alo = lo part, both zero
// padded to a multiple of GKB columns); every operand travels by 16-byte cp.async copies into the no-swizzle
// core-matrix layout and the k blocks run through a three-stage ring, so the loads of blocks s + 1, s + 2 are in
// flight while block s is multiplied.
template <int BN, bool PRESPLIT>
__global__ void __launch_bounds__(G_THREADS)
tf32x3_gemm_kernel(int M, int N, int K, const float* __restrict__ A, const float* __restrict__ Alo, int lda,
                   const float* __restrict__ Bh, const float* __restrict__ Bl, int ldb, float* __restrict__ C, int ldc,
                   long split_stride, const float* __restrict__ bias, int kb_per_split) {
    constexpr int NST = 3;
    constexpr int TM_COLS = BN <= 128 ? 128 : 256;
    constexpr int A_FLOATS = GM * GKB, B_FLOATS = BN * GKB;
    constexpr int STAGE_FLOATS = 2 * A_FLOATS + 2 * B_FLOATS;       // A_hi, A_lo, B_hi, B_lo
    constexpr int A_IT = GM * GKB / 4 / G_THREADS, B_IT = BN * GKB / 4 / G_THREADS;
    extern __shared__ __align__(1024) unsigned char gsm[];
    float* stage0 = reinterpret_cast<float*>(gsm);
    __shared__ __align__(8) unsigned long long mbar[3];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * BN;
    const int nkb_total = (K + GKB - 1) / GKB;
    const int kb_lo = blockIdx.z * kb_per_split, kb_hi = min(nkb_total, kb_lo + kb_per_split);
    const int n_steps = max(0, kb_hi - kb_lo);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(TM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[2])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);

    {
        static_assert(PRESPLIT, "the activation always arrives pre-split now");
        // every operand by cp.async (rows beyond M re-read row M - 1: their accumulator rows are never stored)
        auto issue_loads = [&](int step) {
            const int k0 = (kb_lo + step) * GKB;
            float* sAh = stage0 + (step % NST) * STAGE_FLOATS;
            float* sAl = sAh + A_FLOATS;
            float* sBh = sAl + A_FLOATS;
            float* sBl = sBh + B_FLOATS;
#pragma unroll
            for (int i = 0; i < A_IT; ++i) {
                const int q = i * G_THREADS + tid;
                const int r = (q >> 6) * 8 + (q & 7), k = ((q & 63) >> 3) * 4;
                const size_t go = (size_t)min(m0 + r, M - 1) * lda + k0 + k;
                cp_async16(sAh + tile_idx(r, k), A + go);
                cp_async16(sAl + tile_idx(r, k), Alo + go);
            }
#pragma unroll
            for (int i = 0; i < B_IT; ++i) {
                const int q = i * G_THREADS + tid;
                const int r = (q >> 6) * 8 + (q & 7), k = ((q & 63) >> 3) * 4;
                const size_t go = (size_t)(n0 + r) * ldb + k0 + k;
                cp_async16(sBh + tile_idx(r, k), Bh + go);
                cp_async16(sBl + tile_idx(r, k), Bl + go);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int s0 = 0; s0 < NST && s0 < n_steps; ++s0) issue_loads(s0);
        for (int step = 0; step < n_steps; ++step) {
            const int s = step % NST;
            const int newer = min(n_steps - 1 - step, NST - 1);        // groups committed after this step's
            if (newer >= 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
            else if (newer == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;");
                float* sAh = stage0 + s * STAGE_FLOATS;
                const uint32_t ah = smem_u32(sAh), al = smem_u32(sAh + A_FLOATS), bh = smem_u32(sAh + 2 * A_FLOATS),
                               bl = smem_u32(sAh + 2 * A_FLOATS + B_FLOATS);
                issue_mma(tmem, ah, bh, idesc, step == 0);      // a_hi . b_hi
                issue_mma(tmem, ah, bl, idesc, false);          // a_hi . b_lo
                issue_mma(tmem, al, bh, idesc, false);          // a_lo . b_hi
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                 smem_u32(&mbar[s]))
                             : "memory");
            }
            if (step + NST < n_steps) {                          // refill this stage once its MMAs have read it
                mbar_wait(smem_u32(&mbar[s]), (uint32_t)((step / NST) & 1));
                issue_loads(step + NST);
            }
        }
    }
    // ---- epilogue: wait for the last commit (covers every earlier MMA), TMEM -> registers -> global
    if (n_steps > 0) {
        const int last = n_steps - 1;
        mbar_wait(smem_u32(&mbar[last % NST]), (uint32_t)((last / NST) & 1));
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // TMEM -> registers (+ bias) -> shared memory (the operand stages are idle now) -> global: tcgen05.ld hands each
    // lane one ROW of the tile, so a direct store would scatter every instruction over 32 rows; through shared
    // memory each warp writes contiguous row segments.
    constexpr int PITCH = BN + 4;
    float* stile = stage0;
    {
        float* srow = stile + (warp * 32 + lane) * PITCH;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t v[16];
            if (n_steps > 0) {
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0u;
            }
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                float4 o = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                                       __uint_as_float(v[i + 3]));
                const int n = n0 + c0 + i;
                if (bias && n < N) {          // N and the column offsets are multiples of 4
                    const float4 bz = __ldg(reinterpret_cast<const float4*>(bias + n));
                    o.x += bz.x; o.y += bz.y; o.z += bz.z; o.w += bz.w;
                }
                *reinterpret_cast<float4*>(srow + c0 + i) = o;
            }
        }
    }
    __syncwarp();                                       // a warp stores exactly the 32 rows it wrote
    for (int r = 0; r < 32; ++r) {
        const int row = m0 + warp * 32 + r;
        float* crow = C + (size_t)blockIdx.z * split_stride + (size_t)row * ldc;
#pragma unroll
        for (int c = 4 * lane; c < BN; c += 128)
            if (row < M && n0 + c < N)
                *reinterpret_cast<float4*>(crow + n0 + c) = *reinterpret_cast<const float4*>(stile + (warp * 32 + r) * PITCH + c);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS));
    }
}

// ------------------------------------------------------------------------------------------------
// TMA-fed, warp-specialised variant for pre-split operands (the forward contraction): one thread streams the
// four operand tiles of every k block with cp.async.bulk.tensor (SASS UTMALDG) into a three-stage ring, one
// thread issues the tcgen05 MMAs and hands stages back through tcgen05.commit, all four warps run the epilogue.
// Operand tiles are K-major with the 128-byte swizzle: a k block is 32 floats = one 128-byte row, so each tile
// is a plain 2-D TMA box (32 x rows, CU_TENSOR_MAP_SWIZZLE_128B - full 128-byte bursts; a first version that
// gathered the no-swizzle core-matrix layout as a 4-D box of 16-byte pieces ran at a third of this rate) and the
// UMMA descriptors use LayoutType SWIZZLE_128B, SBO = 1024 B, advancing 32 bytes per K = 8 instruction.
// Rows beyond the tensor are zero-filled by the TMA unit.
// ------------------------------------------------------------------------------------------------
#define G_NST 3

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(tm), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

// K-major, SWIZZLE_128B: 8-row groups of 128-byte rows, SBO = 1024 B, LBO unused, version 1, layout type 2
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void issue_mma_sw128(uint32_t tmem, uint32_t a_base, uint32_t b_base, uint32_t idesc, bool first) {
#pragma unroll
    for (int j = 0; j < GKB / 8; ++j) {
        const uint64_t da = make_smem_desc_sw128(a_base + j * 32), db = make_smem_desc_sw128(b_base + j * 32);
        const uint32_t acc = (first && j == 0) ? 0u : 1u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem),
            "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u)
            : "memory");
    }
}

template <int BN>
__global__ void __launch_bounds__(G_THREADS)
tf32x3_gemm_tma_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                       const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, int M, int N,
                       int n_kb_total, int kb_per_split, float* __restrict__ C, int ldc, long split_stride,
                       const float* __restrict__ bias) {
    // split-K: grid.z slices of kb_per_split k blocks, partial products to C + z * split_stride (the consumer sums)
    const int kb_lo = blockIdx.z * kb_per_split;
    const int n_kb = max(0, min(n_kb_total, kb_lo + kb_per_split) - kb_lo);
    C += (size_t)blockIdx.z * split_stride;
    constexpr int TM_COLS = BN <= 128 ? 128 : 256;
    constexpr int A_FLOATS = GM * GKB, B_FLOATS = BN * GKB;
    constexpr int STAGE_FLOATS = 2 * A_FLOATS + 2 * B_FLOATS;
    constexpr uint32_t STAGE_BYTES = STAGE_FLOATS * 4;
    extern __shared__ __align__(1024) unsigned char gsm[];
    float* stage0 = reinterpret_cast<float*>(gsm);
    __shared__ __align__(8) unsigned long long full[G_NST], empty[G_NST], done;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_bias[BN];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * BN;
    // the tile's bias row goes to shared memory now, long before the epilogue wants it
    if (tid < BN / 4) {
        float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias && n0 + 4 * tid < N) bz = __ldg(reinterpret_cast<const float4*>(bias + n0) + tid);
        reinterpret_cast<float4*>(s_bias)[tid] = bz;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(TM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int i = 0; i < G_NST; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[i])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);

    if (tid == 0) {
        // ---- TMA producer
        for (int kb = 0; kb < n_kb; ++kb) {
            const int s = kb % G_NST;
            if (kb >= G_NST) mbar_wait(smem_u32(&empty[s]), (uint32_t)(((kb / G_NST) - 1) & 1));
            const uint32_t bar = smem_u32(&full[s]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(STAGE_BYTES) : "memory");
            float* sAh = stage0 + s * STAGE_FLOATS;
            tma_load_2d(smem_u32(sAh), &tmAh, (kb_lo + kb) * GKB, m0, bar);
            tma_load_2d(smem_u32(sAh + A_FLOATS), &tmAl, (kb_lo + kb) * GKB, m0, bar);
            tma_load_2d(smem_u32(sAh + 2 * A_FLOATS), &tmBh, (kb_lo + kb) * GKB, n0, bar);
            tma_load_2d(smem_u32(sAh + 2 * A_FLOATS + B_FLOATS), &tmBl, (kb_lo + kb) * GKB, n0, bar);
        }
    } else if (tid == 32) {
        // ---- MMA issuer
        for (int kb = 0; kb < n_kb; ++kb) {
            const int s = kb % G_NST;
            mbar_wait(smem_u32(&full[s]), (uint32_t)((kb / G_NST) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;");
            float* sAh = stage0 + s * STAGE_FLOATS;
            const uint32_t ah = smem_u32(sAh), al = smem_u32(sAh + A_FLOATS), bh = smem_u32(sAh + 2 * A_FLOATS),
                           bl = smem_u32(sAh + 2 * A_FLOATS + B_FLOATS);
            issue_mma_sw128(tmem, ah, bh, idesc, kb == 0);      // a_hi . b_hi
            issue_mma_sw128(tmem, ah, bl, idesc, false);        // a_hi . b_lo
            issue_mma_sw128(tmem, al, bh, idesc, false);        // a_lo . b_hi
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             smem_u32(&empty[s]))
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done))
                     : "memory");
    }
    // ---- epilogue (all warps): accumulator complete -> TMEM -> registers (+ bias) -> shared memory (the operand
    // ring is idle now) -> global.  tcgen05.ld hands every lane one ROW of the tile; storing from there would put
    // the 32 lanes of a store on 32 different rows (16 bytes each, 23 KB apart).  Through shared memory each warp
    // writes whole 512-byte row segments instead.
    mbar_wait(smem_u32(&done), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;");
    constexpr int PITCH = BN + 4;                       // floats; keeps the 16-byte row-wise stores off one bank group
    float* stile = stage0;
    {
        float* srow = stile + (warp * 32 + lane) * PITCH;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (n_kb == 0) {                           // a split-K slice beyond the last k block contributes zeros
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0u;
            }
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 bz = *reinterpret_cast<const float4*>(s_bias + c0 + i);
                *reinterpret_cast<float4*>(srow + c0 + i) =
                    make_float4(__uint_as_float(v[i]) + bz.x, __uint_as_float(v[i + 1]) + bz.y,
                                __uint_as_float(v[i + 2]) + bz.z, __uint_as_float(v[i + 3]) + bz.w);
            }
        }
    }
    __syncwarp();                                       // a warp stores exactly the 32 rows it wrote
    for (int r = 0; r < 32; ++r) {
        const int row = m0 + warp * 32 + r;
#pragma unroll
        for (int c = 4 * lane; c < BN; c += 128)
            if (row < M && n0 + c < N)                  // N and the column offsets are multiples of 4
                *reinterpret_cast<float4*>(C + (size_t)row * ldc + n0 + c) =
                    *reinterpret_cast<const float4*>(stile + (warp * 32 + r) * PITCH + c);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS));
    }
}

// tensor map of an operand matrix [rows][K] fp32 (leading dimension ld floats), 128-byte swizzle;
// box = one (box_rows x GKB) operand tile.  Returns false when the driver entry point is missing or refuses.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

bool dsf_make_operand_tmap(CUtensorMap* tm, const float* base, long rows, int K, long ld, int box_rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || (K % GKB) || (ld % 4) || box_rows > 256 || ((uintptr_t)base & 15)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {GKB, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, bool PRESPLIT>
static int launch_tf32x3(int M, int N, int K, const float* A, const float* Alo, int lda, const float* Bh, const float* Bl,
                         int ldb, float* C, int ldc, long split_stride, const float* bias, int n_split, cudaStream_t st) {
    const size_t smem = (size_t)3 * 2 * (GM + BN) * GKB * sizeof(float);
    static bool attr_set[16] = {};
    int dev = 0;
    DSF_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 16 || !attr_set[dev]) {
        DSF_CHECK_CUDA(cudaFuncSetAttribute(tf32x3_gemm_kernel<BN, PRESPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
        if (dev < 16) attr_set[dev] = true;
    }
    const int nkb = (K + GKB - 1) / GKB;
    const int kbs = (nkb + n_split - 1) / n_split;
    dim3 grid((N + BN - 1) / BN, (M + GM - 1) / GM, n_split);
    tf32x3_gemm_kernel<BN, PRESPLIT><<<grid, G_THREADS, smem, st>>>(M, N, K, A, Alo, lda, Bh, Bl, ldb, C, ldc, split_stride,
                                                                   bias, kbs);
    DSF_CHECK_LAUNCH();
    return DSF_OK;
}

// forward: C (M x 2336) = A (M x 148) . basis + bias; the activation arrives split (Ah, Al: M x 160, zero padded);
// Bh/Bl = basis^T split, (2432 x 160) zero padded
int dsf_blend_forward_gemm(int M, const float* Ah, const float* Al, int lda, const float* Bh, const float* Bl, float* C,
                           int ldc, const float* bias, cudaStream_t st) {
    // TMA path: tensor maps are cheap to encode (host only) and travel as kernel parameters.  The activation rows
    // are padded to a multiple of 8 hands by the workspace (dsf_mano_workspace_floats), as the row-group view needs.
    static int tma_state = 0;                     // 0 untried, 1 works, -1 unavailable (fall back to cp.async)
    CUtensorMap tAh, tAl, tBh, tBl;
    if (tma_state >= 0 && dsf_make_operand_tmap(&tAh, Ah, M, BLEND_KPAD, lda, GM) &&
        dsf_make_operand_tmap(&tAl, Al, M, BLEND_KPAD, lda, GM) &&
        dsf_make_operand_tmap(&tBh, Bh, BLEND_NPAD, BLEND_KPAD, BLEND_KPAD, 128) &&
        dsf_make_operand_tmap(&tBl, Bl, BLEND_NPAD, BLEND_KPAD, BLEND_KPAD, 128)) {
        tma_state = 1;
        const size_t smem = (size_t)G_NST * 2 * (GM + 128) * GKB * sizeof(float);
        static bool attr_set[16] = {};
        int dev = 0;
        DSF_CHECK_CUDA(cudaGetDevice(&dev));
        if (dev >= 16 || !attr_set[dev]) {
            DSF_CHECK_CUDA(cudaFuncSetAttribute(tf32x3_gemm_tma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem));
            if (dev < 16) attr_set[dev] = true;
        }
        dim3 grid((NP + 127) / 128, (M + GM - 1) / GM);
        tf32x3_gemm_tma_kernel<128><<<grid, G_THREADS, smem, st>>>(tAh, tAl, tBh, tBl, M, NP, BLEND_KPAD / GKB, BLEND_KPAD / GKB, C, ldc,
                                                                   0, bias);
        DSF_CHECK_LAUNCH();
        return DSF_OK;
    }
    tma_state = -1;
    return launch_tf32x3<128, true>(M, NP, BLEND_KPAD, Ah, Al, lda, Bh, Bl, BLEND_KPAD, C, ldc, 0, bias, 1, st);
}

// backward: BLEND_SPLITS partial products C_z (M x 148) = A[:, kz] (M x 2336) . basis^T[kz, :];
// Bh/Bl = basis split, (160 x 2336); the consumer sums the partials in a fixed order
int dsf_blend_backward_splits(int M) {
    // enough CTAs to cover the chip: (M / 128) row tiles x splits ~ 148 SMs, within [4, BLEND_SPLITS]
    const int m_tiles = (M + GM - 1) / GM;
    int s = 148 / m_tiles;
    return s < 4 ? 4 : (s > BLEND_SPLITS ? BLEND_SPLITS : s);
}

int dsf_blend_backward_gemm(int M, const float* Ah, const float* Al, int lda, const float* Bh, const float* Bl, float* C,
                            int ldc, long split_stride, cudaStream_t st) {
    // The cotangent arrives pre-split (mano_skin_bwd_kernel), so the backward contraction runs through the same
    // TMA-fed, warp-specialised pipeline as the forward one: 128 x 160 tiles, split-K over grid.z.
    const int n_split = dsf_blend_backward_splits(M);
    static int tma_state = 0;
    CUtensorMap tAh, tAl, tBh, tBl;
    if (tma_state >= 0 && dsf_make_operand_tmap(&tAh, Ah, M, NP, lda, GM) && dsf_make_operand_tmap(&tAl, Al, M, NP, lda, GM) &&
        dsf_make_operand_tmap(&tBh, Bh, BLEND_KPAD, NP, NP, 160) && dsf_make_operand_tmap(&tBl, Bl, BLEND_KPAD, NP, NP, 160)) {
        tma_state = 1;
        const size_t smem = (size_t)G_NST * 2 * (GM + 160) * GKB * sizeof(float);
        static bool attr_set[16] = {};
        int dev = 0;
        DSF_CHECK_CUDA(cudaGetDevice(&dev));
        if (dev >= 16 || !attr_set[dev]) {
            DSF_CHECK_CUDA(cudaFuncSetAttribute(tf32x3_gemm_tma_kernel<160>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem));
            if (dev < 16) attr_set[dev] = true;
        }
        const int nkb = NP / GKB;
        const int kbs = (nkb + n_split - 1) / n_split;
        dim3 grid(1, (M + GM - 1) / GM, n_split);
        tf32x3_gemm_tma_kernel<160><<<grid, G_THREADS, smem, st>>>(tAh, tAl, tBh, tBl, M, KP, nkb, kbs, C, ldc, split_stride,
                                                                   nullptr);
        DSF_CHECK_LAUNCH();
        return DSF_OK;
    }
    tma_state = -1;
    return launch_tf32x3<160, true>(M, KP, NP, Ah, Al, lda, Bh, Bl, NP, C, ldc, split_stride, nullptr, n_split, st);
}
