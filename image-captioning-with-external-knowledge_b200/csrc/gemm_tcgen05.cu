// bf16 GEMMs on the 5th-generation tensor cores: tcgen05.mma with TMEM accumulators, operands staged by TMA
// (cp.async.bulk.tensor, 128B swizzle) through an mbarrier ring, persistent CTAs, warp-specialised:
//   warp 0    TMA producer (one elected lane)
//   warp 1    TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma / tcgen05.commit)
//   warps 2-9 epilogue (two warps per TMEM lane quadrant, alternating 32-column chunks): tcgen05.ld -> registers ->
//             bias / ReLU+dropout / ReLU-backward / accumulate -> 16-byte global stores
//
//   gemm_tn : C[M,N] (+)= A[M,K] W[N,K]^T (+bias), both operands K-major (nn.Linear forward and dgrad).  The projections
//             here have K = 320/512, i.e. 5-8 k-blocks per tile, so the kernel is EPILOGUE/HBM-bound: two TMEM accumulator
//             stages overlap the epilogue of tile i with the main loop of tile i+1, the epilogue keeps everything in
//             registers (no local memory) and the tile width is chosen per N to avoid padding waste.
//   wgrad   : G[n,k] += sum_m dY[m,n] X[m,k], both operands MN-major (TMA boxes of the row-major activations are exactly
//             the canonical MN-major SWIZZLE_128B layout).  One work item = a 128 x K (K <= 512: the whole TMEM width)
//             output tile over a slice of the rows; partial tiles go to an fp32 workspace with coalesced 16-byte stores
//             and a second kernel reduces the slices and scatters through the packing maps into the flat gradient
//             buffer (no atomics).  L2-bound by construction (rows are streamed once per 128-wide n tile).
#include <cuda.h>

#include "common.cuh"
#include "ickb200.h"

namespace {

constexpr int BM = 128;       // UMMA M (TMEM lanes)
constexpr int BK = 64;        // reduction elements per stage = one 128-byte swizzle row of bf16
constexpr int UMMA_K = 16;    // K per tcgen05.mma for 16-bit inputs
constexpr int MAX_STAGES = 4;
constexpr int MAX_BN = 256;   // widest single tcgen05.mma N
constexpr int A_STAGE = BM * BK * 2;      // 16 KiB
constexpr int TMEM_COLS = 512;
constexpr int BAR_BYTES = 1024;           // barriers + TMEM base pointer live in the first KiB
constexpr int SMEM_DATA = 4 * (A_STAGE + MAX_BN * BK * 2);  // 192 KiB of stage buffers
constexpr int SMEM_BYTES = SMEM_DATA + BAR_BYTES + 1024 /*align slack*/;
constexpr int NEPI = 8;                   // epilogue warps
constexpr int NTHREADS = 64 + 32 * NEPI;

// ---- PTX wrappers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (CUDA error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s at 2 GHz
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D=f32, A=B=bf16
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}


// ---- shared-memory carve-up: [barriers (1 KiB)] [stage 0: A | B] [stage 1: A | B] ... ------------------------------------
struct Smem {
    uint32_t base;   // shared-space address of the 1024-aligned region
    uint32_t stage_bytes;
    uint32_t* tmem_ptr;
    __device__ __forceinline__ uint32_t full(int i) const { return base + 8 * i; }
    __device__ __forceinline__ uint32_t empty(int i) const { return base + 8 * (MAX_STAGES + i); }
    __device__ __forceinline__ uint32_t tfull(int i) const { return base + 8 * (2 * MAX_STAGES + i); }
    __device__ __forceinline__ uint32_t tempty(int i) const { return base + 8 * (2 * MAX_STAGES + 2 + i); }
    __device__ __forceinline__ uint32_t a(int stage) const { return base + BAR_BYTES + stage * stage_bytes; }
    __device__ __forceinline__ uint32_t b(int stage) const { return a(stage) + A_STAGE; }
};
__device__ __forceinline__ Smem carve(uint8_t* raw, uint32_t stage_bytes) {
    uint8_t* p = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    Smem s;
    s.base = smem_u32(p);
    s.stage_bytes = stage_bytes;
    s.tmem_ptr = (uint32_t*)(p + 8 * (2 * MAX_STAGES + 4));
    return s;
}

__device__ __forceinline__ void setup(const Smem& s, int warp, int lane, const CUtensorMap* t0, const CUtensorMap* t1) {
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(t0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(t1) : "memory");
        for (int i = 0; i < MAX_STAGES; ++i) {
            mbar_init(s.full(i), 1);
            mbar_init(s.empty(i), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(s.tfull(i), 1);
            mbar_init(s.tempty(i), NEPI);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s.tmem_ptr)), "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
}
__device__ __forceinline__ void teardown(int warp, uint32_t tmem_base) {
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---- register-resident epilogue helpers (every index is a compile-time constant after unrolling) ---------------------------
__device__ __forceinline__ void ld32_f32(const float* p, bool vec, int nv, float* v) {
    if (vec) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p + j));
            v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = j < nv ? __ldg(p + j) : 0.f;
    }
}
__device__ __forceinline__ void ld32_bf16(const bf16* p, bool vec, int nv, float* v) {
    if (vec) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) ld8(p + j, v + j);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = j < nv ? __bfloat162float(p[j]) : 0.f;
    }
}
__device__ __forceinline__ void st32_f32(float* p, bool vec, int nv, const float* v) {
    if (vec) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nv) p[j] = v[j];
    }
}
__device__ __forceinline__ void st32_bf16(bf16* p, bool vec, int nv, const float* v) {
    if (vec) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) st8(p + j, v + j);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nv) p[j] = __float2bfloat16_rn(v[j]);
    }
}

struct TnParams {
    void* C;
    const float* bias;
    const void* aux;
    int M, N, K, ldc, ldaux, epi, accumulate, c_f32, BN, n_tiles_n, n_tiles, nkb, stages;
    DropCfg drop;
};

__global__ void __launch_bounds__(NTHREADS, 1) gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmW, TnParams p) {
    extern __shared__ uint8_t smem_raw[];
    ick_resolve_seed(p.drop);
    const Smem s = carve(smem_raw, A_STAGE + p.BN * BK * 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    setup(s, warp, lane, &tmA, &tmW);
    const uint32_t tmem_base = *s.tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = (uint32_t)(BM + p.BN) * BK * 2;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const int m_idx = (tile / p.n_tiles_n) * BM, n_idx = (tile % p.n_tiles_n) * p.BN;
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(s.empty(stage), phase ^ 1);
                    mbar_expect_tx(s.full(stage), tx);
                    tma_load_2d(s.a(stage), &tmA, s.full(stage), kb * BK, m_idx);
                    tma_load_2d(s.b(stage), &tmW, s.full(stage), kb * BK, n_idx);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(p.BN, 0, 0);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                mbar_wait(s.tempty(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * MAX_BN;
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(s.full(stage), phase);
                    tc_fence_after();
                    const uint32_t a0 = s.a(stage), b0 = s.b(stage);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // K-major SWIZZLE_128B: 8-row groups 1024 B apart; a K step of 16 elements = 32 B inside the row
                        const uint64_t ad = make_desc(a0 + k * UMMA_K * 2, 16, 1024);
                        const uint64_t bd = make_desc(b0 + k * UMMA_K * 2, 16, 1024);
                        tc_mma_bf16(d_tmem, ad, bd, idesc, (kb | k) != 0);
                    }
                    tc_commit(s.empty(stage));  // frees the smem stage once these MMAs have read it
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                tc_commit(s.tfull(acc));  // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int q = warp & 3;           // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;  // which of the two warps of the quadrant: even / odd 32-column chunks
        int acc = 0;
        uint32_t acc_phase = 0;
        const bool c_al = (p.ldc % (p.c_f32 ? 4 : 8) == 0) && ((((uintptr_t)p.C) & 15) == 0);
        const bool aux_al = p.aux != nullptr && (p.ldaux % 8 == 0) && ((((uintptr_t)p.aux) & 15) == 0);
        const bool bias_al = p.bias != nullptr && ((((uintptr_t)p.bias) & 15) == 0);
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int m_idx = (tile / p.n_tiles_n) * BM, n_idx = (tile % p.n_tiles_n) * p.BN;
            mbar_wait(s.tfull(acc), acc_phase);
            tc_fence_after();
            const int row = m_idx + q * 32 + lane;
            for (int c0 = half * 32; c0 < p.BN; c0 += 64) {
                uint32_t r[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * MAX_BN + c0, r);
                const int col0 = n_idx + c0;
                if (row < p.M && col0 < p.N) {
                    const int nv = min(32, p.N - col0);
                    const bool full = nv == 32 && (col0 & 7) == 0;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if (p.bias) {
                        float t[32];
                        ld32_f32(p.bias + col0, full && bias_al, nv, t);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += t[j];
                    }
                    if (p.accumulate) {
                        float t[32];
                        if (p.c_f32) ld32_f32((const float*)p.C + (size_t)row * p.ldc + col0, false, nv, t);
                        else ld32_bf16((const bf16*)p.C + (size_t)row * p.ldc + col0, full && c_al, nv, t);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += t[j];
                    }
                    if (p.epi == 1) {
                        const uint32_t rmix = ick_rowmix(p.drop.seed, p.drop.site, (uint64_t)row);
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {  // col0 is even: (j, j+1) is one hash pair
                            float k0 = 1.f, k1 = 1.f;
                            if (p.drop.thr != 0u) {
                                const uint32_t hsh = ick_pairhash(rmix, (uint32_t)(col0 + j));
                                k0 = ick_keep_lo(hsh, p.drop.thr) ? p.drop.inv_keep : 0.f;
                                k1 = ick_keep_hi(hsh, p.drop.thr) ? p.drop.inv_keep : 0.f;
                            }
                            v[j] = fmaxf(v[j], 0.f) * k0;
                            v[j + 1] = fmaxf(v[j + 1], 0.f) * k1;
                        }
                    } else if (p.epi == 2) {
                        float t[32];
                        ld32_bf16((const bf16*)p.aux + (size_t)row * p.ldaux + col0, full && aux_al, nv, t);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = t[j] != 0.f ? v[j] * p.drop.inv_keep : 0.f;
                    }
                    if (p.c_f32) st32_f32((float*)p.C + (size_t)row * p.ldc + col0, full && c_al, nv, v);
                    else st32_bf16((bf16*)p.C + (size_t)row * p.ldc + col0, full && c_al, nv, v);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s.tempty(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    teardown(warp, tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------------
// wgrad: D[n (128 lanes), k (KT <= 512 columns)] = sum over a row slice of dY[m,n] * X[m,k].  TMA boxes are
// [64 rows of m][64 columns]: in shared memory that is the canonical MN-major SWIZZLE_128B layout (64 contiguous MN
// elements per 128-byte row, one row per reduction index), SBO = 1024 B between 8-row groups, LBO = one whole box
// (BK*128 B) between 64-wide MN blocks.  KT > 256 is issued as two tcgen05.mma per k-step (N <= 256 each).
constexpr int WG_BOX = BK * 128;  // bytes of one [64 x 64] bf16 box
struct WgParams {
    float* G;
    float* ws;  // [splits][N][Kws] partial tiles, or nullptr -> atomics straight into G
    const int* rowoff;
    const int* colmap;
    int M, N, K, KT, Kws, n_tiles_n, n_tiles_k, splits, m_per_split, stages;
};

__global__ void __launch_bounds__(NTHREADS, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY,
                                                                const __grid_constant__ CUtensorMap tmX, WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int nbx = p.KT / 64;  // X boxes per stage
    const Smem s = carve(smem_raw, A_STAGE + nbx * WG_BOX);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    setup(s, warp, lane, &tmY, &tmX);
    const uint32_t tmem_base = *s.tmem_ptr;
    const int n_work = p.n_tiles_n * p.n_tiles_k * p.splits;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = (uint32_t)(2 + nbx) * WG_BOX;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int split = w % p.splits, t = w / p.splits;
                const int n_idx = (t / p.n_tiles_k) * BM, k_idx = (t % p.n_tiles_k) * p.KT;
                const int m0 = split * p.m_per_split, m1 = min(p.M, m0 + p.m_per_split);
                for (int m = m0; m < m1; m += BK) {
                    mbar_wait(s.empty(stage), phase ^ 1);
                    mbar_expect_tx(s.full(stage), tx);
                    // m_per_split is a multiple of BK, so only the global tail block is partial (TMA zero-fills rows >= M)
                    tma_load_2d(s.a(stage), &tmY, s.full(stage), n_idx, m);
                    tma_load_2d(s.a(stage) + WG_BOX, &tmY, s.full(stage), n_idx + 64, m);
                    for (int j = 0; j < nbx; ++j) tma_load_2d(s.b(stage) + j * WG_BOX, &tmX, s.full(stage), k_idx + 64 * j, m);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const int n1 = p.KT > MAX_BN ? MAX_BN : p.KT, n2 = p.KT - n1;
            const uint32_t idesc1 = make_idesc(n1, 1, 1), idesc2 = make_idesc(n2 > 0 ? n2 : 16, 1, 1);
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                const int split = w % p.splits;
                const int m0 = split * p.m_per_split, m1 = min(p.M, m0 + p.m_per_split);
                mbar_wait(s.tempty(0), acc_phase ^ 1);  // single accumulator stage: wait until the epilogue drained it
                tc_fence_after();
                int it = 0;
                for (int m = m0; m < m1; m += BK, ++it) {
                    mbar_wait(s.full(stage), phase);
                    tc_fence_after();
                    const uint32_t a0 = s.a(stage), b0 = s.b(stage);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // a reduction step of 16 = 16 rows of 128 B
                        const uint64_t ad = make_desc(a0 + k * UMMA_K * 128, WG_BOX, 1024);
                        tc_mma_bf16(tmem_base, ad, make_desc(b0 + k * UMMA_K * 128, WG_BOX, 1024), idesc1, (it | k) != 0);
                        if (n2 > 0)
                            tc_mma_bf16(tmem_base + MAX_BN, ad, make_desc(b0 + 4 * WG_BOX + k * UMMA_K * 128, WG_BOX, 1024), idesc2,
                                        (it | k) != 0);
                    }
                    tc_commit(s.empty(stage));
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                tc_commit(s.tfull(0));
                acc_phase ^= 1;
            }
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;
        uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int split = w % p.splits, t = w / p.splits;
            const int n_idx = (t / p.n_tiles_k) * BM, k_idx = (t % p.n_tiles_k) * p.KT;
            mbar_wait(s.tfull(0), acc_phase);
            tc_fence_after();
            const int n = n_idx + q * 32 + lane;
            const int ro = (n < p.N && p.ws == nullptr) ? p.rowoff[n] : -1;
            for (int c0 = half * 32; c0 < p.KT; c0 += 64) {
                uint32_t r[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, r);
                if (p.ws != nullptr) {
                    if (n < p.N && k_idx + c0 < p.Kws) {  // Kws is a multiple of 32: whole chunks only
                        float* dst = p.ws + ((size_t)split * p.N + n) * p.Kws + k_idx + c0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(dst + j) =
                                make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                    }
                } else if (ro >= 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int k = k_idx + c0 + j;
                        if (k < p.K) {
                            const int cm = p.colmap ? p.colmap[k] : k;
                            if (cm >= 0) atomicAdd(p.G + ro + cm, __uint_as_float(r[j]));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s.tempty(0));
            acc_phase ^= 1;
        }
    }
    teardown(warp, tmem_base);
}

// sum the row-slice partials and scatter into the flat gradient buffer: G[rowoff[n] + colmap[k]] += sum_s ws[s][n][k]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ G, const int* __restrict__ rowoff,
                                                           const int* __restrict__ colmap, int N, int K, int Kws, int splits) {
    const int n = blockIdx.y;
    const int ro = rowoff[n];
    if (ro < 0) return;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
        const int cm = colmap ? colmap[k] : k;
        if (cm < 0) continue;
        float acc = 0.f;
        for (int sp = 0; sp < splits; ++sp) acc += ws[((size_t)sp * N + n) * Kws + k];
        G[ro + cm] += acc;
    }
}

// bias gradient: gflat[biasoff[n]] += sum_m dY[m,n].  HBM-bound column sum: a CTA covers 128 columns x `rpb` rows,
// 64 threads x bf16x2 per row (128 contiguous bytes per warp-pair) and 4 row groups reduced through shared memory.
__global__ void __launch_bounds__(256) bias_grad_kernel(const bf16* __restrict__ dY, float* __restrict__ G, const int* __restrict__ biasoff,
                                                        int M, int N, int ldy, int rpb) {
    __shared__ float red[4][128];
    const int tx = threadIdx.x & 63, rg = threadIdx.x >> 6;
    const int c = blockIdx.x * 128 + 2 * tx;
    const int r0 = blockIdx.y * rpb, r1 = min(M, r0 + rpb);
    float s0 = 0.f, s1 = 0.f;
    if (c + 1 < N) {
#pragma unroll 4
        for (int r = r0 + rg; r < r1; r += 4) {
            const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dY + (size_t)r * ldy + c));
            s0 += v.x;
            s1 += v.y;
        }
    } else if (c < N) {
        for (int r = r0 + rg; r < r1; r += 4) s0 += __bfloat162float(dY[(size_t)r * ldy + c]);
    }
    red[rg][2 * tx] = s0;
    red[rg][2 * tx + 1] = s1;
    __syncthreads();
    if (threadIdx.x < 128) {
        const int cc = blockIdx.x * 128 + threadIdx.x;
        if (cc < N) {
            const int bo = biasoff[cc];
            if (bo >= 0) atomicAdd(G + bo, red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x]);
        }
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeFn)p;
    }
    return fn;
}
// 2D bf16 tensor map: inner extent `inner` elements (contiguous), `rows` rows of stride ld elements; box = 64 x box_rows
int make_tmap(CUtensorMap* tm, const void* ptr, int inner, int rows, int ld, int box_rows) {
    EncodeFn enc = get_encode();
    if (!enc) {
        ick_set_error("cuTensorMapEncodeTiled entry point not available");
        return ICK_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ick_set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%d rows=%d ld=%d box_rows=%d", (int)r, ptr, inner, rows, ld, box_rows);
        return ICK_ERR_CUDA;
    }
    return ICK_OK;
}

int pick_bn(int N) {
    // widest tile whose padding waste is smallest; multiples of 32 (epilogue chunk) up to 256
    int best = 128, best_pad = 1 << 30;
    for (int bn = 256; bn >= 64; bn -= 32) {
        const int pad = (N + bn - 1) / bn * bn - N;
        if (pad < best_pad) { best_pad = pad; best = bn; }
    }
    return best;
}

int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <typename K>
int set_smem(K kernel) {
    static bool done = false;  // one static per kernel type (template instantiation)
    if (!done) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) {
            ick_set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize=%d) failed", SMEM_BYTES);
            return ICK_ERR_CUDA;
        }
        done = true;
    }
    return ICK_OK;
}

}  // namespace

extern "C" int ick_gemm_tn_tc(const void* A, const void* W, void* C, int c_dt, const float* bias, const void* aux, int M, int N, int K,
                              int lda, int ldw, int ldc, int ldaux, int epi, int accumulate, float drop_p, unsigned seed, unsigned site,
                              cudaStream_t stream) {
    ICK_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_tn_tc: bad sizes M=%d N=%d K=%d", M, N, K);
    ICK_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, "gemm_tn_tc: lda/ldw must be multiples of 8 (16-byte TMA strides)");
    ICK_REQUIRE((((uintptr_t)A) & 15) == 0 && (((uintptr_t)W) & 15) == 0, "gemm_tn_tc: operands must be 16-byte aligned");
    ICK_REQUIRE(c_dt == ICK_F32 || c_dt == ICK_BF16, "gemm_tn_tc: bad output dtype %d", c_dt);
    ICK_REQUIRE(epi >= 0 && epi <= 2 && (epi != 2 || (aux != nullptr && c_dt == ICK_BF16)), "gemm_tn_tc: bad epilogue");
    if (M == 0) return ICK_OK;
    int rc = set_smem(gemm_tn_tc_kernel);
    if (rc) return rc;
    TnParams p;
    p.C = C; p.bias = bias; p.aux = aux;
    p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.ldaux = ldaux; p.epi = epi; p.accumulate = accumulate;
    p.c_f32 = c_dt == ICK_F32;
    p.BN = pick_bn(N);
    p.n_tiles_n = (N + p.BN - 1) / p.BN;
    p.n_tiles = ((M + BM - 1) / BM) * p.n_tiles_n;
    p.nkb = (K + BK - 1) / BK;
    p.stages = MAX_STAGES;
    p.drop = make_drop(drop_p, seed, site);
    CUtensorMap tmA, tmW;
    if ((rc = make_tmap(&tmA, A, K, M, lda, BM))) return rc;
    if ((rc = make_tmap(&tmW, W, K, N, ldw, p.BN))) return rc;
    const int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
    gemm_tn_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, stream>>>(tmA, tmW, p);
    return ick_check_launch("gemm_tn_tc");
}

extern "C" int ick_wgrad_tc(const void* dY, const void* X, float* gflat, const int* rowoff, const int* colmap, const int* biasoff, int M,
                            int N, int K, int ldy, int ldx, void* workspace, long long workspace_bytes, cudaStream_t stream) {
    ICK_REQUIRE(M >= 0 && N > 0 && K > 0, "wgrad_tc: bad sizes M=%d N=%d K=%d", M, N, K);
    ICK_REQUIRE(ldy % 8 == 0 && ldx % 8 == 0, "wgrad_tc: ldy/ldx must be multiples of 8");
    ICK_REQUIRE((((uintptr_t)dY) & 15) == 0 && (((uintptr_t)X) & 15) == 0, "wgrad_tc: operands must be 16-byte aligned");
    ICK_REQUIRE(rowoff != nullptr, "wgrad_tc: rowoff is required");
    if (M == 0) return ICK_OK;
    int rc = set_smem(wgrad_tc_kernel);
    if (rc) return rc;
    WgParams p;
    p.G = gflat; p.rowoff = rowoff; p.colmap = colmap;
    p.M = M; p.N = N; p.K = K;
    const int Kpad = (K + 63) / 64 * 64;
    p.n_tiles_k = (Kpad + TMEM_COLS - 1) / TMEM_COLS;
    p.KT = ((Kpad / 64 + p.n_tiles_k - 1) / p.n_tiles_k) * 64;  // even split of the 64-wide boxes over the k tiles, <= 512
    p.n_tiles_n = (N + BM - 1) / BM;
    const int tiles = p.n_tiles_n * p.n_tiles_k;
    const int stage_bytes = A_STAGE + (p.KT / 64) * WG_BOX;
    p.stages = SMEM_DATA / stage_bytes;
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    ICK_REQUIRE(p.stages >= 2, "wgrad_tc: tile does not fit");
    int splits = (num_sms() + tiles - 1) / tiles;
    const int max_splits = (M + 8 * BK - 1) / (8 * BK);  // at least 8 k-blocks per work item
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int mps = (M + splits - 1) / splits;
    mps = (mps + BK - 1) / BK * BK;
    p.splits = (M + mps - 1) / mps;
    p.m_per_split = mps;
    p.Kws = (K + 31) / 32 * 32;
    const long long need = (long long)p.splits * N * p.Kws * 4;
    p.ws = (workspace != nullptr && workspace_bytes >= need && (((uintptr_t)workspace) & 15) == 0) ? (float*)workspace : nullptr;
    CUtensorMap tmY, tmX;
    if ((rc = make_tmap(&tmY, dY, N, M, ldy, BK))) return rc;
    if ((rc = make_tmap(&tmX, X, K, M, ldx, BK))) return rc;
    const int n_work = tiles * p.splits;
    const int grid = n_work < num_sms() ? n_work : num_sms();
    wgrad_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, stream>>>(tmY, tmX, p);
    if ((rc = ick_check_launch("wgrad_tc"))) return rc;
    if (p.ws != nullptr) {
        dim3 rgrid((K + 255) / 256, N);
        wgrad_reduce_kernel<<<rgrid, 256, 0, stream>>>(p.ws, gflat, rowoff, colmap, N, K, p.Kws, p.splits);
        if ((rc = ick_check_launch("wgrad_reduce"))) return rc;
    }
    if (biasoff) {
        const int rpb = 256;
        dim3 bgrid((N + 127) / 128, (M + rpb - 1) / rpb);
        bias_grad_kernel<<<bgrid, 256, 0, stream>>>((const bf16*)dY, gflat, biasoff, M, N, ldy, rpb);
        return ick_check_launch("bias_grad");
    }
    return ICK_OK;
}
