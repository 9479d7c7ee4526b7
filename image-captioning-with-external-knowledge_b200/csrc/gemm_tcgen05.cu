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

#ifdef ICK_TRACE
// Debug build only (-DICK_TRACE): per-role timeline of CTA 0 of the last gemm_tn_tc launch, read back by tools/gemm_trace.py.
// Each role appends to its own 1024-entry lane with a thread-local counter (plain stores, no atomics: a few cycles per event).
__device__ unsigned long long ick_trace_buf[4 * 2048];
#define ICK_TR_DECL(role) unsigned int tr_i_ = 0; const unsigned int tr_role_ = (role)
#define ICK_TR(ev, tile)                                                                              \
    do {                                                                                              \
        if (blockIdx.x == 0 && tr_i_ < 1023) {                                                        \
            ick_trace_buf[tr_role_ * 2048 + 2 * tr_i_] = ((unsigned long long)(ev) << 32) | (unsigned)(tile); \
            ick_trace_buf[tr_role_ * 2048 + 2 * tr_i_ + 1] = clock64();                               \
            ++tr_i_;                                                                                  \
            ick_trace_buf[tr_role_ * 2048 + 2 * tr_i_] = 0xFFFFFFFFFFFFFFFFull;                       \
        }                                                                                             \
    } while (0)
extern "C" int ick_debug_trace_read(unsigned long long* out, int max_entries) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, ick_trace_buf, sizeof(unsigned long long) * 4 * 2048);
    return 4 * 1024;
}
#else
#define ICK_TR_DECL(role) do {} while (0)
#define ICK_TR(ev, tile) do {} while (0)
#endif

namespace {

constexpr int BM = 128;       // UMMA M (TMEM lanes)
constexpr int BK = 64;        // reduction elements per stage = one 128-byte swizzle row of bf16
constexpr int UMMA_K = 16;    // K per tcgen05.mma for 16-bit inputs
constexpr int MAX_STAGES = 8;
constexpr int MAX_WKB = 16;   // k-blocks of a resident weight panel (K <= 1024)
constexpr int MAX_BN = 256;   // widest single tcgen05.mma N
constexpr int A_STAGE = BM * BK * 2;      // 16 KiB
constexpr int TMEM_COLS = 512;
constexpr int BAR_BYTES = 1024;           // barriers + TMEM base pointer live in the first KiB
constexpr int SMEM_BYTES = 232448;        // all 227 KiB a CTA may have
constexpr int SMEM_DATA = SMEM_BYTES - BAR_BYTES - 1024 /*align slack*/;  // 225 KiB of operand buffers
constexpr int NEPI = 8;                   // epilogue warps
constexpr int NSBOX = 2;                  // staging boxes per epilogue warp (bulk stores in flight)
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
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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

__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

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
    uint32_t panel_bytes;
    uint32_t* tmem_ptr;
    __device__ __forceinline__ uint32_t full(int i) const { return base + 8 * i; }
    __device__ __forceinline__ uint32_t empty(int i) const { return base + 8 * (MAX_STAGES + i); }
    __device__ __forceinline__ uint32_t tfull(int i) const { return base + 8 * (2 * MAX_STAGES + i); }
    __device__ __forceinline__ uint32_t tempty(int i) const { return base + 8 * (2 * MAX_STAGES + 2 + i); }
    __device__ __forceinline__ uint32_t wfull(int i) const { return base + 8 * (2 * MAX_STAGES + 4 + i); }
    __device__ __forceinline__ uint32_t panel() const { return base + BAR_BYTES; }  // resident weight panel (W-stationary mode)
    __device__ __forceinline__ uint32_t a(int stage) const { return base + BAR_BYTES + panel_bytes + stage * stage_bytes; }
    __device__ __forceinline__ uint32_t b(int stage) const { return a(stage) + A_STAGE; }
};
__device__ __forceinline__ Smem carve(uint8_t* raw, uint32_t stage_bytes, uint32_t panel_bytes = 0) {
    uint8_t* p = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    Smem s;
    s.base = smem_u32(p);
    s.stage_bytes = stage_bytes;
    s.panel_bytes = panel_bytes;
    s.tmem_ptr = (uint32_t*)(p + 8 * (2 * MAX_STAGES + 4 + MAX_WKB));
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
        for (int i = 0; i < MAX_WKB; ++i) mbar_init(s.wfull(i), 1);
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
    int wstat;    // W-stationary: CTA c keeps weight panel (c % n_tiles_n) resident and streams only A tiles
    int m_tiles;  // row tiles
    int tma_store;  // output through per-warp shared-memory staging + TMA bulk stores (needs 16-byte aligned C rows)
    // Two row groups with their own weights (ick_gemm_tn_tc_dual): row tiles >= split_tile read the second weight map, the
    // second bias and hash their dropout with the second site and a row index relative to the split.
    int split_tile;
    const float* bias2;
    uint32_t site2;
    // epi 3 (out-projection dgrad of an attention block): C = dO as usual, and the epilogue also writes the backward's row term
    // dsum[(b*H + h)*S + i] = sum_d dO[b*S+i, 32h+d] * O[b*S+i, 32h+d] (aux = O; a 32-column chunk is exactly one head), which
    // saves the separate row-dot kernel and its re-read of dO and O.  Second row group (dual launch): dsum2 / rd_S2, rows
    // counted from the split; rows in [rd_rows0, split) are padding.
    float* dsum;
    float* dsum2;
    int rd_S, rd_S2, rd_rows0, rd_H;
    DropCfg drop;
};
// Work of a CTA.  Streaming mode: tiles blockIdx.x, +gridDim.x, ... over the (m, n) grid, n fastest.  W-stationary mode:
// gridDim.x = n_tiles_n * cpp; CTA c owns panel c % n_tiles_n and row tiles c / n_tiles_n, + cpp, ... (CTAs that share a row
// tile run side by side, so the A tile is fetched from HBM once and re-read from L2).
struct TileIter {
    int m_tile, n_tile, step_m, it, end;
    bool wstat;
    int n_tiles_n;
    __device__ __forceinline__ TileIter(const TnParams& p) {
        wstat = p.wstat != 0;
        n_tiles_n = p.n_tiles_n;
        if (wstat) {
            n_tile = blockIdx.x % p.n_tiles_n;
            it = blockIdx.x / p.n_tiles_n;
            step_m = gridDim.x / p.n_tiles_n;
            end = p.m_tiles;
        } else {
            it = blockIdx.x;
            step_m = gridDim.x;
            end = p.n_tiles;
            n_tile = 0;
        }
        m_tile = 0;
    }
    __device__ __forceinline__ bool next() {  // sets m_tile / n_tile of the current work item, then advances
        if (it >= end) return false;
        if (wstat) m_tile = it;
        else { m_tile = it / n_tiles_n; n_tile = it % n_tiles_n; }
        it += step_m;
        return true;
    }
};

// KIND >= 0: the epilogue switches are compile-time constants - bits 0-1 epilogue kind, bit 2 accumulate, bit 3 fp32 output, bulk
// stores through the staging boxes; the per-role timeline showed ~200 of an epilogue chunk's ~1 150 cycles in uniform branches over
// the kinds that do not apply.  KIND = -1: everything read from the parameter block (unaligned outputs, rare combinations).
template <int KIND>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmW,
                                                                  const __grid_constant__ CUtensorMap tmW2,
                                                                  const __grid_constant__ CUtensorMap tmC, TnParams p) {
    ick_pdl_launch();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t wbox = (uint32_t)p.BN * BK * 2;  // one k-block of the weight tile
    const Smem s = p.wstat ? carve(smem_raw, A_STAGE, (uint32_t)p.nkb * wbox) : carve(smem_raw, A_STAGE + wbox);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    setup(s, warp, lane, &tmA, &tmW);
    const uint32_t tmem_base = *s.tmem_ptr;
    ick_pdl_wait();  // barriers, TMEM and tensor maps are ready: from here on the previous kernel's results are needed
    ick_resolve_seed(p.drop);

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            ICK_TR_DECL(0);
            TileIter ti(p);
            if (p.wstat && ti.it < ti.end) {  // the CTA's weight panel, once
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_expect_tx(s.wfull(kb), wbox);
                    tma_load_2d(s.panel() + kb * wbox, &tmW, s.wfull(kb), kb * BK, ti.n_tile * p.BN);
                }
            }
            const uint32_t tx = p.wstat ? (uint32_t)A_STAGE : (uint32_t)A_STAGE + wbox;
            ICK_TR(1, 0);
            while (ti.next()) {
                const int m_idx = ti.m_tile * BM, n_idx = ti.n_tile * p.BN;
                const CUtensorMap* tw = ti.m_tile >= p.split_tile ? &tmW2 : &tmW;
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(s.empty(stage), phase ^ 1);
                    ICK_TR(2, ti.m_tile * 100 + kb);
                    mbar_expect_tx(s.full(stage), tx);
                    tma_load_2d(s.a(stage), &tmA, s.full(stage), kb * BK, m_idx);
                    if (!p.wstat) tma_load_2d(s.b(stage), tw, s.full(stage), kb * BK, n_idx);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(p.BN, 0, 0);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            bool first = true;
            ICK_TR_DECL(1);
            TileIter ti(p);
            while (ti.next()) {
                mbar_wait(s.tempty(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * MAX_BN;
                ICK_TR(3, ti.m_tile);
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(s.full(stage), phase);
                    if (p.wstat && first) mbar_wait(s.wfull(kb), 0);
                    ICK_TR(4, ti.m_tile * 100 + kb);
                    tc_fence_after();
                    const uint32_t a0 = s.a(stage), b0 = p.wstat ? s.panel() + kb * wbox : s.b(stage);
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
                first = false;
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int q = warp & 3;           // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;  // which of the two warps of the quadrant: even / odd 32-column chunks
        int acc = 0;
        uint32_t acc_phase = 0;
        const int k_epi = KIND >= 0 ? (KIND & 3) : p.epi;
        const bool k_acc = KIND >= 0 ? ((KIND >> 2) & 1) != 0 : p.accumulate != 0;
        const bool k_f32 = KIND >= 0 ? ((KIND >> 3) & 1) != 0 : p.c_f32 != 0;
        const bool k_tma = KIND >= 0 ? true : p.tma_store != 0;
        const bool c_al = (p.ldc % (k_f32 ? 4 : 8) == 0) && ((((uintptr_t)p.C) & 15) == 0);
        const bool aux_al = p.aux != nullptr && (p.ldaux % 8 == 0) && ((((uintptr_t)p.aux) & 15) == 0);
        const bool bias_al = p.bias != nullptr && ((((uintptr_t)p.bias) & 15) == 0) && ((((uintptr_t)p.bias2) & 15) == 0);
        // Bulk-store path: each warp owns NSBOX staging boxes of [32 rows x 32 columns] behind the operand stages (64-byte rows
        // swizzled 64B for bf16, 128-byte rows swizzled 128B for fp32; lane r writes row r, so each 16-byte chunk column is
        // bank-conflict free); one lane issues the TMA store of a finished box and the warp goes on with its next chunk.  Stores
        // issued by the warps themselves (row per lane, or transposed through shared memory) measured 35-40% slower.
        const uint32_t sbox = k_f32 ? 4096u : 2048u;
        const uint32_t stg = s.a(0) + (uint32_t)p.stages * s.stage_bytes + (uint32_t)(warp - 2) * NSBOX * sbox;
        int sb = 0;
        ICK_TR_DECL(2);
        TileIter ti(p);
        while (ti.next()) {
            const int m_idx = ti.m_tile * BM, n_idx = ti.n_tile * p.BN;
            if (warp == 2 && lane == 0) ICK_TR(5, ti.m_tile);
            mbar_wait(s.tfull(acc), acc_phase);
            if (warp == 2 && lane == 0) ICK_TR(6, ti.m_tile);
            tc_fence_after();
            const int row = m_idx + q * 32 + lane;
            const bool second = ti.m_tile >= p.split_tile;
            const float* bias_p = second ? p.bias2 : p.bias;
            for (int c0 = half * 32; c0 < p.BN; c0 += 64) {
                float bv[32];  // bias of this chunk's columns, requested before the TMEM load so that the two latencies overlap
                {
                    const int col0 = n_idx + c0;
                    const int nv = min(32, p.N - col0);
                    if (bias_p != nullptr && nv > 0) ld32_f32(bias_p + col0, nv == 32 && (col0 & 7) == 0 && bias_al, nv, bv);
                }
                uint32_t r[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * MAX_BN + c0, r);
                if (warp == 2 && lane == 0) ICK_TR(7, ti.m_tile * 100 + c0 / 32);
                if (c0 + 64 >= p.BN) {  // this warp's last chunk of the accumulator is in registers: hand the TMEM stage back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s.tempty(acc));
                }
                const int col0 = n_idx + c0;
                if (col0 >= p.N || m_idx + q * 32 >= p.M) continue;  // warp-uniform: nothing of this chunk is inside C
                const int nv = min(32, p.N - col0);
                const bool full = nv == 32 && (col0 & 7) == 0;
                const bool live = row < p.M;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = bias_p ? __uint_as_float(r[j]) + bv[j] : __uint_as_float(r[j]);
                if (k_acc && live) {
                    float t[32];
                    if (k_f32) ld32_f32((const float*)p.C + (size_t)row * p.ldc + col0, false, nv, t);
                    else ld32_bf16((const bf16*)p.C + (size_t)row * p.ldc + col0, full && c_al, nv, t);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += t[j];
                }
                if (k_epi == 1) {
                    const uint32_t rmix = second ? ick_rowmix(p.drop.seed, p.site2, (uint64_t)(row - p.split_tile * BM))
                                                 : ick_rowmix(p.drop.seed, p.drop.site, (uint64_t)row);
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
                } else if (k_epi == 2 && live) {
                    float t[32];
                    ld32_bf16((const bf16*)p.aux + (size_t)row * p.ldaux + col0, full && aux_al, nv, t);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = t[j] != 0.f ? v[j] * p.drop.inv_keep : 0.f;
                } else if (k_epi == 3 && live) {
                    float t[32];
                    ld32_bf16((const bf16*)p.aux + (size_t)row * p.ldaux + col0, full && aux_al, nv, t);
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc = fmaf(t[j], __bfloat162float(__float2bfloat16_rn(v[j])), acc);  // dO as stored (bf16)
                    const int h = col0 >> 5;
                    if (h < p.rd_H) {
                        if (!second) {
                            if (row < p.rd_rows0) p.dsum[((size_t)(row / p.rd_S) * p.rd_H + h) * p.rd_S + row % p.rd_S] = acc;
                        } else {
                            const int r2 = row - p.split_tile * BM;
                            p.dsum2[((size_t)(r2 / p.rd_S2) * p.rd_H + h) * p.rd_S2 + r2 % p.rd_S2] = acc;
                        }
                    }
                }
                if (k_tma) {
                    const uint32_t box = stg + (uint32_t)sb * sbox;
                    if (lane == 0) bulk_wait_read<NSBOX - 1>();  // the bulk store that last read this box is done with it
                    __syncwarp();
                    if (k_f32) {
                        const uint32_t rowa = box + (uint32_t)lane * 128u;
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + (uint32_t)((c ^ (lane & 7)) << 4)), "f"(v[4 * c]),
                                         "f"(v[4 * c + 1]), "f"(v[4 * c + 2]), "f"(v[4 * c + 3])
                                         : "memory");
                    } else {
                        const uint32_t rowa = box + (uint32_t)lane * 64u;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            __nv_bfloat162 t0 = __floats2bfloat162_rn(v[8 * c], v[8 * c + 1]), t1 = __floats2bfloat162_rn(v[8 * c + 2], v[8 * c + 3]);
                            __nv_bfloat162 t2 = __floats2bfloat162_rn(v[8 * c + 4], v[8 * c + 5]), t3 = __floats2bfloat162_rn(v[8 * c + 6], v[8 * c + 7]);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + (uint32_t)((c ^ ((lane >> 1) & 3)) << 4)),
                                         "r"(*reinterpret_cast<uint32_t*>(&t0)), "r"(*reinterpret_cast<uint32_t*>(&t1)),
                                         "r"(*reinterpret_cast<uint32_t*>(&t2)), "r"(*reinterpret_cast<uint32_t*>(&t3))
                                         : "memory");
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmC, box, col0, m_idx + q * 32);  // rows >= M and columns >= N are clipped by the tensor map
                        bulk_commit();
                    }
                    sb = sb + 1 == NSBOX ? 0 : sb + 1;
                    if (warp == 2 && lane == 0) ICK_TR(8, ti.m_tile * 100 + c0 / 32);
                } else if (live) {
                    if (k_f32) st32_f32((float*)p.C + (size_t)row * p.ldc + col0, full && c_al, nv, v);
                    else st32_bf16((bf16*)p.C + (size_t)row * p.ldc + col0, full && c_al, nv, v);
                }
            }
            if (half * 32 >= p.BN) {  // (BN = 32 only) a warp without a chunk still releases the accumulator stage
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(s.tempty(acc));
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (k_tma && lane == 0) bulk_wait_read<0>();  // staging boxes must outlive their stores
        if (warp == 2 && lane == 0) ICK_TR(9, 0);
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
    const int* biasoff;  // fused bias gradient (column sums of dY) or nullptr
    float* wsb;          // [splits][N] bias partials (with ws)
    int M, N, K, KT, Kws, n_tiles_n, n_tiles_k, splits, m_per_split, stages;
};

__global__ void __launch_bounds__(NTHREADS, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY,
                                                                const __grid_constant__ CUtensorMap tmX, WgParams p) {
    ick_pdl_launch();
    extern __shared__ uint8_t smem_raw[];
    const int nbx = p.KT / 64;  // X boxes per stage
    const Smem s = carve(smem_raw, A_STAGE + nbx * WG_BOX);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Fused bias gradient: one more N=16 MMA per k-step against a constant box of ones (an all-ones operand is invariant under
    // the 128B swizzle), i.e. D[n, KT..KT+15] = sum_m dY[m,n].  The box lives behind the last stage.
    const uint32_t ones = s.a(0) + (uint32_t)p.stages * s.stage_bytes;
    if (p.biasoff != nullptr) {
        uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023) + BAR_BYTES + (size_t)p.stages * s.stage_bytes;
        for (int i = threadIdx.x; i < WG_BOX / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(base)[i] = 0x3F803F80u;  // bf16 1.0 x2
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    setup(s, warp, lane, &tmY, &tmX);
    const uint32_t tmem_base = *s.tmem_ptr;
    ick_pdl_wait();
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
            const uint32_t idesc1 = make_idesc(n1, 1, 1), idesc2 = make_idesc(n2 > 0 ? n2 : 16, 1, 1), idesc_b = make_idesc(16, 1, 1);
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
                        if (p.biasoff != nullptr)
                            tc_mma_bf16(tmem_base + p.KT, ad, make_desc(ones + k * UMMA_K * 128, WG_BOX, 1024), idesc_b, (it | k) != 0);
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
            if (p.biasoff != nullptr && half == 0) {
                uint32_t r[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + p.KT, r);
                if (n < p.N) {
                    if (p.wsb != nullptr) p.wsb[(size_t)split * p.N + n] = __uint_as_float(r[0]);
                    else if (p.biasoff[n] >= 0) atomicAdd(p.G + p.biasoff[n], __uint_as_float(r[0]));
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

// sum the row-slice partials and scatter into the flat gradient buffer: G[rowoff[n] + colmap[k]] += sum_s ws[s][n][k].
// One thread per 4 consecutive k (Kws is a multiple of 32), 16-byte loads, four splits in flight.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ G, const int* __restrict__ rowoff,
                                                           const int* __restrict__ colmap, int N, int K, int Kws, int splits,
                                                           const float* __restrict__ wsb, const int* __restrict__ biasoff) {
    ick_pdl_entry();
    const int q4 = Kws >> 2;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)N * q4) return;
    const int n = (int)(idx / q4), k = (int)(idx % q4) * 4;
    if (wsb != nullptr && k == 0 && biasoff[n] >= 0) {
        float acc = 0.f;
        for (int sp = 0; sp < splits; ++sp) acc += wsb[(size_t)sp * N + n];
        G[biasoff[n]] += acc;
    }
    const int ro = rowoff[n];
    if (ro < 0 || k >= K) return;
    const float* src = ws + (size_t)n * Kws + k;
    const size_t stride = (size_t)N * Kws;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    int sp = 0;
    for (; sp + 3 < splits; sp += 4) {
        const float4 v0 = *reinterpret_cast<const float4*>(src + (size_t)sp * stride);
        const float4 v1 = *reinterpret_cast<const float4*>(src + (size_t)(sp + 1) * stride);
        const float4 v2 = *reinterpret_cast<const float4*>(src + (size_t)(sp + 2) * stride);
        const float4 v3 = *reinterpret_cast<const float4*>(src + (size_t)(sp + 3) * stride);
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
        a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
        a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
        a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
    }
    for (; sp < splits; ++sp) {
        const float4 v0 = *reinterpret_cast<const float4*>(src + (size_t)sp * stride);
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    }
    const float r[4] = {(a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y), (a0.z + a1.z) + (a2.z + a3.z), (a0.w + a1.w) + (a2.w + a3.w)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (k + j >= K) break;
        const int cm = colmap ? colmap[k + j] : k + j;
        if (cm >= 0) G[ro + cm] += r[j];
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// Grouped wgrad: the weight gradients of several linear layers (one Transformer layer's worth: QKV / out / FFN ...) in ONE
// persistent launch.  Launched one by one, each of these small problems pays its own launch, pipeline fill, exposed epilogue
// and partial-tile reduce (15-20 us of fixed cost against 5-10 us of streaming for the decoder-sized ones); grouped, the work
// items (problem, 128-row n tile, k tile, row split) of all problems are balanced over the SMs in a single wave.  Every
// problem uses k tiles of at most WG_GKT = 320 columns, so that all stages have one size (2 dY boxes + 5 X boxes), the
// 16 bias columns always fit behind the tile in TMEM and K = 512 (FFN second layer) becomes two k tiles of 256.
constexpr int WG_MAXP = 8;
constexpr int WG_GKT = 320;
struct WgProb {
    float* ws;   // [splits][N][Kws] partial tiles
    float* wsb;  // [splits][N] bias partials (nullptr: no bias gradient)
    const int* rowoff;
    const int* colmap;
    const int* biasoff;
    int M, N, K, KT, Kws, n_tiles_k, splits, m_per_split;
    int item0;  // first work item of this problem
    int q0;     // first float4 of this problem in the grouped reduce
};
struct WgGroup {
    CUtensorMap tmY[WG_MAXP], tmX[WG_MAXP];
    WgProb pr[WG_MAXP];
    float* G;
    int nprob, n_items, stages, n_q4;
};
struct WgItem {
    int pi, split, n_idx, k_idx, k_tile, m0, m1;
};
__device__ __forceinline__ WgItem wg_item(const WgGroup& g, int w) {
    WgItem it;
    it.pi = 0;
    while (it.pi + 1 < g.nprob && w >= g.pr[it.pi + 1].item0) ++it.pi;
    const WgProb& p = g.pr[it.pi];
    const int local = w - p.item0;
    it.split = local % p.splits;
    const int t = local / p.splits;
    it.k_tile = t % p.n_tiles_k;
    it.n_idx = (t / p.n_tiles_k) * BM;
    it.k_idx = it.k_tile * p.KT;
    it.m0 = it.split * p.m_per_split;
    it.m1 = min(p.M, it.m0 + p.m_per_split);
    return it;
}

__global__ void __launch_bounds__(NTHREADS, 1) wgrad_group_tc_kernel(const __grid_constant__ WgGroup g) {
    ick_pdl_launch();
    extern __shared__ uint8_t smem_raw[];
    const Smem s = carve(smem_raw, A_STAGE + (WG_GKT / 64) * WG_BOX);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ones = s.a(0) + (uint32_t)g.stages * s.stage_bytes;  // constant box of bf16 1.0 (bias gradient operand)
    {
        uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023) + BAR_BYTES + (size_t)g.stages * s.stage_bytes;
        for (int i = threadIdx.x; i < WG_BOX / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(base)[i] = 0x3F803F80u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    setup(s, warp, lane, &g.tmY[0], &g.tmX[0]);
    const uint32_t tmem_base = *s.tmem_ptr;
    ick_pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int w = blockIdx.x; w < g.n_items; w += gridDim.x) {
                const WgItem it = wg_item(g, w);
                const WgProb& p = g.pr[it.pi];
                const int nbx = p.KT / 64;
                const uint32_t tx = (uint32_t)(2 + nbx) * WG_BOX;
                const CUtensorMap* tY = &g.tmY[it.pi];
                const CUtensorMap* tX = &g.tmX[it.pi];
                for (int m = it.m0; m < it.m1; m += BK) {
                    mbar_wait(s.empty(stage), phase ^ 1);
                    mbar_expect_tx(s.full(stage), tx);
                    tma_load_2d(s.a(stage), tY, s.full(stage), it.n_idx, m);
                    tma_load_2d(s.a(stage) + WG_BOX, tY, s.full(stage), it.n_idx + 64, m);
                    for (int j = 0; j < nbx; ++j) tma_load_2d(s.b(stage) + j * WG_BOX, tX, s.full(stage), it.k_idx + 64 * j, m);
                    if (++stage == g.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_b = make_idesc(16, 1, 1);
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int w = blockIdx.x; w < g.n_items; w += gridDim.x) {
                const WgItem it = wg_item(g, w);
                const WgProb& p = g.pr[it.pi];
                const int n1 = p.KT > MAX_BN ? MAX_BN : p.KT, n2 = p.KT - n1;
                const uint32_t idesc1 = make_idesc(n1, 1, 1), idesc2 = make_idesc(n2 > 0 ? n2 : 16, 1, 1);
                const bool bias = p.wsb != nullptr && it.k_tile == 0;
                mbar_wait(s.tempty(0), acc_phase ^ 1);
                tc_fence_after();
                int k_it = 0;
                for (int m = it.m0; m < it.m1; m += BK, ++k_it) {
                    mbar_wait(s.full(stage), phase);
                    tc_fence_after();
                    const uint32_t a0 = s.a(stage), b0 = s.b(stage);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t ad = make_desc(a0 + k * UMMA_K * 128, WG_BOX, 1024);
                        tc_mma_bf16(tmem_base, ad, make_desc(b0 + k * UMMA_K * 128, WG_BOX, 1024), idesc1, (k_it | k) != 0);
                        if (n2 > 0)
                            tc_mma_bf16(tmem_base + MAX_BN, ad, make_desc(b0 + 4 * WG_BOX + k * UMMA_K * 128, WG_BOX, 1024), idesc2,
                                        (k_it | k) != 0);
                        if (bias) tc_mma_bf16(tmem_base + p.KT, ad, make_desc(ones + k * UMMA_K * 128, WG_BOX, 1024), idesc_b, (k_it | k) != 0);
                    }
                    tc_commit(s.empty(stage));
                    if (++stage == g.stages) { stage = 0; phase ^= 1; }
                }
                tc_commit(s.tfull(0));
                acc_phase ^= 1;
            }
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;
        uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < g.n_items; w += gridDim.x) {
            const WgItem it = wg_item(g, w);
            const WgProb& p = g.pr[it.pi];
            mbar_wait(s.tfull(0), acc_phase);
            tc_fence_after();
            const int n = it.n_idx + q * 32 + lane;
            for (int c0 = half * 32; c0 < p.KT; c0 += 64) {
                uint32_t r[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, r);
                if (n < p.N && it.k_idx + c0 < p.Kws) {  // Kws is a multiple of 32: whole chunks only
                    float* dst = p.ws + ((size_t)it.split * p.N + n) * p.Kws + it.k_idx + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(dst + j) =
                            make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                }
            }
            if (p.wsb != nullptr && it.k_tile == 0 && half == 0) {
                uint32_t r[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + p.KT, r);
                if (n < p.N) p.wsb[(size_t)it.split * p.N + n] = __uint_as_float(r[0]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s.tempty(0));
            acc_phase ^= 1;
        }
    }
    teardown(warp, tmem_base);
}

// grouped reduce: one thread per 4 consecutive k of one (problem, n)
__global__ void __launch_bounds__(256) wgrad_group_reduce_kernel(const __grid_constant__ WgGroup g) {
    ick_pdl_entry();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.n_q4) return;
    int pi = 0;
    while (pi + 1 < g.nprob && idx >= g.pr[pi + 1].q0) ++pi;
    const WgProb& p = g.pr[pi];
    const int q4 = p.Kws >> 2, local = idx - p.q0;
    const int n = local / q4, k = (local % q4) * 4;
    float* G = g.G;
    if (p.wsb != nullptr && k == 0 && p.biasoff[n] >= 0) {
        float acc = 0.f;
        for (int sp = 0; sp < p.splits; ++sp) acc += p.wsb[(size_t)sp * p.N + n];
        G[p.biasoff[n]] += acc;
    }
    const int ro = p.rowoff[n];
    if (ro < 0 || k >= p.K) return;
    const float* src = p.ws + (size_t)n * p.Kws + k;
    const size_t stride = (size_t)p.N * p.Kws;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    int sp = 0;
    for (; sp + 3 < p.splits; sp += 4) {  // four partial tiles in flight per thread (the kernel is latency-bound: ~1 TB/s with two)
        const float4 v0 = *(reinterpret_cast<const float4*>(src + (size_t)sp * stride));
        const float4 v1 = *(reinterpret_cast<const float4*>(src + (size_t)(sp + 1) * stride));
        const float4 v2 = *(reinterpret_cast<const float4*>(src + (size_t)(sp + 2) * stride));
        const float4 v3 = *(reinterpret_cast<const float4*>(src + (size_t)(sp + 3) * stride));
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
        a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
        a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
        a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
    }
    for (; sp < p.splits; ++sp) {
        const float4 v0 = *(reinterpret_cast<const float4*>(src + (size_t)sp * stride));
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    }
    const float r[4] = {(a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y), (a0.z + a1.z) + (a2.z + a3.z), (a0.w + a1.w) + (a2.w + a3.w)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (k + j >= p.K) break;
        const int cm = p.colmap ? p.colmap[k + j] : k + j;
        if (cm >= 0) G[ro + cm] += r[j];
    }
}

// bias gradient: gflat[biasoff[n]] += sum_m dY[m,n].  HBM-bound column sum: a CTA covers 128 columns x `rpb` rows,
// 64 threads x bf16x2 per row (128 contiguous bytes per warp-pair) and 4 row groups reduced through shared memory.
__global__ void __launch_bounds__(256) bias_grad_kernel(const bf16* __restrict__ dY, float* __restrict__ G, const int* __restrict__ biasoff,
                                                        int M, int N, int ldy, int rpb) {
    ick_pdl_entry();
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

// ------------------------------------------------------------------------------------------------------------------------
// Linear + residual + dropout + LayerNorm in ONE kernel (the post-LN sublayer tail of every Transformer layer, G/models.py:241-244:
// x = norm(x + dropout(sublayer(x))) with the sublayer ending in out_proj / linear2):
//     s = x + dropout(A W^T + bias)        (s is stored: the LayerNorm backward needs it)
//     y = LN(s) * gamma + beta             (+ mean, rstd per row for the backward)
// A LayerNorm row is the full d_model width, so a work item is a 128-row x 320-column tile: two N = 160 tcgen05.mma per k-step
// into ONE 320-column TMEM accumulator (2 x 320 does not fit the 512 columns: single accumulator stage).  The epilogue makes two
// sweeps over the accumulator: sweep 1 forms s in registers (bias, hash dropout, residual), accumulates the row's sum and sum of
// squares, writes s back INTO TENSOR MEMORY in place (tcgen05.st) and to global (bf16, bulk stores); the two warps that share a
// lane quadrant exchange their half-row statistics through shared memory; sweep 2 re-reads s from TMEM, normalises and stores y.
// Against GEMM + add_ln_fwd this drops one launch and one write + one read of the (rows x 320) sublayer output per LayerNorm.
constexpr int LN_W = 320;                       // padded d_model = tile width
constexpr int LN_HALF = 160;                    // columns per tcgen05.mma
constexpr int LN_STAGE = A_STAGE + LN_W * BK * 2;  // 16 KiB of A + 40 KiB of W per k-block
constexpr int LN_NSBOX = 3;                     // staging boxes per epilogue warp
constexpr int LN_RED_BYTES = 2 * BM * 8;        // (sum, sumsq) per row and half
struct LnParams {
    const bf16* X;  // residual rows (nullptr: none)
    float* mean;
    float* rstd;
    const float* bias; const float* bias2;
    const float* gamma; const float* gamma2;
    const float* beta; const float* beta2;
    int M, D, ldx, nkb, stages, m_tiles, split_tile;
    float eps, inv_d;
    uint32_t site2;
    DropCfg drop;
};

__device__ __forceinline__ void ln_stage_store(uint32_t box, int lane, const float* v, const CUtensorMap* tm, int col0, int row0) {
    if (lane == 0) bulk_wait_read<LN_NSBOX - 1>();  // the bulk store that last read this box is done with it
    __syncwarp();
    const uint32_t rowa = box + (uint32_t)lane * 64u;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        __nv_bfloat162 t0 = __floats2bfloat162_rn(v[8 * c], v[8 * c + 1]), t1 = __floats2bfloat162_rn(v[8 * c + 2], v[8 * c + 3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(v[8 * c + 4], v[8 * c + 5]), t3 = __floats2bfloat162_rn(v[8 * c + 6], v[8 * c + 7]);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + (uint32_t)((c ^ ((lane >> 1) & 3)) << 4)),
                     "r"(*reinterpret_cast<uint32_t*>(&t0)), "r"(*reinterpret_cast<uint32_t*>(&t1)), "r"(*reinterpret_cast<uint32_t*>(&t2)),
                     "r"(*reinterpret_cast<uint32_t*>(&t3))
                     : "memory");
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
        tma_store_2d(tm, box, col0, row0);  // rows >= M are clipped by the tensor map
        bulk_commit();
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) gemm_ln_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                                                                  const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmS,
                                                                  const __grid_constant__ CUtensorMap tmY, LnParams p) {
    ick_pdl_launch();
    extern __shared__ uint8_t smem_raw[];
    const Smem s = carve(smem_raw, LN_STAGE);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    setup(s, warp, lane, &tmA, &tmW);
    const uint32_t tmem_base = *s.tmem_ptr;
    ick_pdl_wait();
    ick_resolve_seed(p.drop);

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
                const CUtensorMap* tw = mt >= p.split_tile ? &tmW2 : &tmW;
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(s.empty(stage), phase ^ 1);
                    mbar_expect_tx(s.full(stage), (uint32_t)LN_STAGE);
                    tma_load_2d(s.a(stage), &tmA, s.full(stage), kb * BK, mt * BM);
                    tma_load_2d(s.b(stage), tw, s.full(stage), kb * BK, 0);
                    tma_load_2d(s.b(stage) + LN_HALF * BK * 2, tw, s.full(stage), kb * BK, LN_HALF);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(LN_HALF, 0, 0);
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
                mbar_wait(s.tempty(0), acc_phase ^ 1);  // single accumulator: the epilogue must have drained the previous tile
                tc_fence_after();
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(s.full(stage), phase);
                    tc_fence_after();
                    const uint32_t a0 = s.a(stage), b0 = s.b(stage);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t ad = make_desc(a0 + k * UMMA_K * 2, 16, 1024);
                        tc_mma_bf16(tmem_base, ad, make_desc(b0 + k * UMMA_K * 2, 16, 1024), idesc, (kb | k) != 0);
                        tc_mma_bf16(tmem_base + LN_HALF, ad, make_desc(b0 + LN_HALF * BK * 2 + k * UMMA_K * 2, 16, 1024), idesc, (kb | k) != 0);
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
        const uint32_t stg = s.a(0) + (uint32_t)p.stages * s.stage_bytes + (uint32_t)(warp - 2) * LN_NSBOX * 2048u;
        uint8_t* gen_base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
        float2* red = reinterpret_cast<float2*>(gen_base + BAR_BYTES + (size_t)p.stages * LN_STAGE + NEPI * LN_NSBOX * 2048);
        int sb = 0;
        const bool x_al = p.X != nullptr && (p.ldx % 8 == 0) && ((((uintptr_t)p.X) & 15) == 0);
        // gamma / beta are views into the flat fp32 parameter buffer: any 4-byte offset
        const bool gb_al = ((((uintptr_t)p.gamma) | ((uintptr_t)p.gamma2) | ((uintptr_t)p.beta) | ((uintptr_t)p.beta2)) & 15) == 0;
        const bool dropping = p.drop.thr != 0u;
        for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
            const int m_idx = mt * BM;
            const int row = m_idx + q * 32 + lane;
            const bool live = row < p.M;
            const bool second = mt >= p.split_tile;
            const float* bias_p = second ? p.bias2 : p.bias;
            const float* gamma_p = second ? p.gamma2 : p.gamma;
            const float* beta_p = second ? p.beta2 : p.beta;
            const uint32_t rmix = second ? ick_rowmix(p.drop.seed, p.site2, (uint64_t)(row - p.split_tile * BM))
                                         : ick_rowmix(p.drop.seed, p.drop.site, (uint64_t)row);
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
            mbar_wait(s.tfull(0), acc_phase);
            tc_fence_after();
            float sum = 0.f, sq = 0.f;
            // ---- sweep 1: s = x + dropout(acc + bias) -> TMEM (in place) + global; row statistics ----
            for (int c0 = half * 32; c0 < LN_W; c0 += 64) {
                float bv[32], xv[32];
                if (bias_p != nullptr) ld32_f32(bias_p + c0, true, 32, bv);  // packed biases are LN_W wide, zero padded
                if (p.X != nullptr && live) ld32_bf16(p.X + (size_t)row * p.ldx + c0, x_al, 32, xv);
                uint32_t r[32];
                tc_ld32(trow + c0, r);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = bias_p != nullptr ? __uint_as_float(r[j]) + bv[j] : __uint_as_float(r[j]);
                if (dropping) {
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const uint32_t hsh = ick_pairhash(rmix, (uint32_t)(c0 + j));
                        v[j] *= ick_keep_lo(hsh, p.drop.thr) ? p.drop.inv_keep : 0.f;
                        v[j + 1] *= ick_keep_hi(hsh, p.drop.thr) ? p.drop.inv_keep : 0.f;
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (p.X != nullptr && live) v[j] += xv[j];
                    if (c0 + j >= p.D) v[j] = 0.f;  // pad columns [D, LN_W) stay exactly zero
                    sum += v[j];
                    sq = fmaf(v[j], v[j], sq);
                    r[j] = __float_as_uint(v[j]);
                }
                tc_st32(trow + c0, r);
                ln_stage_store(stg + (uint32_t)sb * 2048u, lane, v, &tmS, c0, m_idx + q * 32);
                sb = sb + 1 == LN_NSBOX ? 0 : sb + 1;
            }
            tc_wait_st();
            red[half * BM + q * 32 + lane] = make_float2(sum, sq);
            asm volatile("bar.sync 1, %0;" ::"n"(NEPI * 32) : "memory");
            {
                const float2 o = red[(half ^ 1) * BM + q * 32 + lane];
                sum += o.x;
                sq += o.y;
            }
            const float mean = sum * p.inv_d;
            const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, sq * p.inv_d), 0.f) + p.eps);
            const float nmr = -mean * rstd;
            if (half == 0 && live) {
                p.mean[row] = mean;
                p.rstd[row] = rstd;
            }
            // ---- sweep 2: y = LN(s) * gamma + beta ----
            for (int c0 = half * 32; c0 < LN_W; c0 += 64) {
                float gv[32], tv[32];
                const int nv = min(32, p.D - c0);  // gamma / beta hold D entries
                ld32_f32(gamma_p + c0, nv == 32 && gb_al, nv, gv);
                ld32_f32(beta_p + c0, nv == 32 && gb_al, nv, tv);
                uint32_t r[32];
                tc_ld32(trow + c0, r);
                if (c0 + 64 >= LN_W) {  // this warp's last chunk is in registers: hand the accumulator back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s.tempty(0));
                }
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = j < nv ? fmaf(fmaf(__uint_as_float(r[j]), rstd, nmr), gv[j], tv[j]) : 0.f;
                ln_stage_store(stg + (uint32_t)sb * 2048u, lane, v, &tmY, c0, m_idx + q * 32);
                sb = sb + 1 == LN_NSBOX ? 0 : sb + 1;
            }
            acc_phase ^= 1;
        }
        if (lane == 0) bulk_wait_read<0>();  // staging boxes must outlive their stores
    }
    teardown(warp, tmem_base);
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

// 2D output map for the epilogue's bulk stores: box = 32 columns x 32 rows, swizzle matching the row width (64 B / 128 B)
int make_tmap_out(CUtensorMap* tm, const void* ptr, int c_f32, int N, int M, int ldc) {
    EncodeFn enc = get_encode();
    if (!enc) {
        ick_set_error("cuTensorMapEncodeTiled entry point not available");
        return ICK_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)ldc * (c_f32 ? 4 : 2)};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(tm, c_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                     es, CU_TENSOR_MAP_INTERLEAVE_NONE, c_f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ick_set_error("cuTensorMapEncodeTiled(out) failed (%d): ptr=%p N=%d M=%d ldc=%d", (int)r, ptr, N, M, ldc);
        return ICK_ERR_CUDA;
    }
    return ICK_OK;
}

int pick_bn(int N) {
    // The kernel is bound by the operand bytes that must enter shared memory: a 128 x BN tile costs (128 + BN) * K * 2 bytes,
    // i.e. per useful output column  padded(N)/N * (1/BN + 1/128).  Multiples of 32 (epilogue chunk) up to 256.
    // (Minimising padding alone picked BN = 64 for the 10000-wide vocabulary projection: twice the operand traffic of 256.)
    int best = 128;
    double best_cost = 1e30;
    for (int bn = 256; bn >= 64; bn -= 32) {
        const int padded = (N + bn - 1) / bn * bn;
        const double cost = (double)padded / N * (1.0 / bn + 1.0 / BM);
        if (cost < best_cost - 1e-12) { best_cost = cost; best = bn; }
    }
    return best;
}

// Tile plan.  Streaming mode re-loads the weight tile with every A tile; W-stationary mode keeps a K-deep weight panel of
// BN columns resident per CTA and streams only A.  Both are scored by the bytes that must enter shared memory per useful
// output column (padding counted), and W-stationary needs >= 4 A stages next to the panel and at most 16 panels.
struct Plan {
    int wstat, BN, stages;
};
int num_sms();
Plan plan_tiles(int N, int nkb, int avail, bool allow_wstat, int m_tiles) {
    Plan best;
    best.wstat = 0;
    best.BN = pick_bn(N);
    // Few row tiles (single-step decode: M = images; small batches): the launch is one tile per CTA and bound by the latency of
    // that tile - its operand load and, above all, its epilogue (BN / 64 serial 32-column chunks per warp).  Spread the columns
    // over the idle SMs instead: the narrowest tile (>= 32) that still fits all tiles in one wave.
    if (m_tiles * ((N + best.BN - 1) / best.BN) * 2 <= num_sms()) {
        for (int bn = 32; bn < best.BN; bn += 32) {
            if (m_tiles * ((N + bn - 1) / bn) <= num_sms()) {
                best.BN = bn;
                break;
            }
        }
    }
    best.stages = avail / (A_STAGE + best.BN * BK * 2);
    if (best.stages > MAX_STAGES) best.stages = MAX_STAGES;
    const auto padded = [&](int bn) { return (N + bn - 1) / bn * bn; };
    double best_cost = (double)(BM + best.BN) / best.BN * padded(best.BN) / N;  // x K*2 bytes, common to all candidates
    if (!allow_wstat || nkb > MAX_WKB) return best;
    for (int bn = 256; bn >= 64; bn -= 32) {
        const int st = (avail - nkb * bn * BK * 2) / A_STAGE;
        if (st < 4 || padded(bn) / bn > 16) continue;
        const double cost = (double)BM / bn * padded(bn) / N;
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best.wstat = 1;
            best.BN = bn;
            best.stages = st > MAX_STAGES ? MAX_STAGES : st;
        }
    }
    return best;
}
bool use_wstat() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ICK_GEMM_WSTAT");  // measured slower than streaming on every shape of the step: opt-in only
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v != 0;
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

struct RowDot {  // epi 3 arguments (see TnParams)
    float* dsum;
    float* dsum2;
    int S, S2, rows0, H;
};
static int gemm_tn_tc_impl(const void* A, const void* W, const void* W2, void* C, int c_dt, const float* bias, const float* bias2,
                           const void* aux, int M, int M_split, int N, int K, int lda, int ldw, int ldc, int ldaux, int epi, int accumulate,
                           float drop_p, unsigned seed, unsigned site, unsigned site2, cudaStream_t stream, const RowDot* rd = nullptr) {
    ICK_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_tn_tc: bad sizes M=%d N=%d K=%d", M, N, K);
    if (W2 != nullptr) {
        ICK_REQUIRE(M_split > 0 && M_split % BM == 0 && M_split <= M, "gemm_tn_tc_dual: M_split=%d must be a positive multiple of %d <= M", M_split,
                    BM);
        ICK_REQUIRE((((uintptr_t)W2) & 15) == 0 && (bias == nullptr) == (bias2 == nullptr), "gemm_tn_tc_dual: bad second weight / bias");
    }
    ICK_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, "gemm_tn_tc: lda/ldw must be multiples of 8 (16-byte TMA strides)");
    ICK_REQUIRE((((uintptr_t)A) & 15) == 0 && (((uintptr_t)W) & 15) == 0, "gemm_tn_tc: operands must be 16-byte aligned");
    ICK_REQUIRE(c_dt == ICK_F32 || c_dt == ICK_BF16, "gemm_tn_tc: bad output dtype %d", c_dt);
    ICK_REQUIRE(epi >= 0 && epi <= 3 && (epi < 2 || (aux != nullptr && c_dt == ICK_BF16)) && ((epi == 3) == (rd != nullptr)),
                "gemm_tn_tc: bad epilogue");
    if (M == 0) return ICK_OK;
    int rc;
    TnParams p;
    p.C = C; p.bias = bias; p.aux = aux;
    p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.ldaux = ldaux; p.epi = epi; p.accumulate = accumulate;
    p.c_f32 = c_dt == ICK_F32;
    p.nkb = (K + BK - 1) / BK;
    p.m_tiles = (M + BM - 1) / BM;
    // the bulk-store epilogue needs 16-byte aligned rows of C; otherwise (e.g. the 10301-wide fp32 score rows of the geo
    // variant) the epilogue writes straight from registers, row per lane
    const int es = c_dt == ICK_F32 ? 4 : 2;
    p.tma_store = ((((uintptr_t)C) & 15) == 0 && ((size_t)ldc * es) % 16 == 0) ? 1 : 0;
    const int staging = p.tma_store ? NEPI * NSBOX * (c_dt == ICK_F32 ? 4096 : 2048) : 0;
    const Plan pl = plan_tiles(N, p.nkb, SMEM_DATA - staging, use_wstat() && W2 == nullptr, p.m_tiles);
    p.split_tile = W2 != nullptr ? M_split / BM : 0x7FFFFFFF;
    p.bias2 = W2 != nullptr ? bias2 : bias;
    p.site2 = site2;
    p.wstat = pl.wstat;
    p.BN = pl.BN;
    p.stages = pl.stages;
    ICK_REQUIRE(p.stages >= 2, "gemm_tn_tc: tile does not fit");
    p.n_tiles_n = (N + p.BN - 1) / p.BN;
    p.n_tiles = p.m_tiles * p.n_tiles_n;
    p.drop = make_drop(drop_p, seed, site);
    p.dsum = p.dsum2 = nullptr;
    p.rd_S = p.rd_S2 = 1;
    p.rd_rows0 = p.rd_H = 0;
    if (rd != nullptr) {
        p.dsum = rd->dsum; p.dsum2 = rd->dsum2;
        p.rd_S = rd->S; p.rd_S2 = rd->S2 > 0 ? rd->S2 : 1; p.rd_rows0 = rd->rows0; p.rd_H = rd->H;
    }
    CUtensorMap tmA, tmW, tmW2, tmC;
    if ((rc = make_tmap(&tmA, A, K, M, lda, BM))) return rc;
    if ((rc = make_tmap(&tmW, W, K, N, ldw, p.BN))) return rc;
    if (W2 != nullptr) {
        if ((rc = make_tmap(&tmW2, W2, K, N, ldw, p.BN))) return rc;
    } else {
        tmW2 = tmW;
    }
    if (p.tma_store) {
        if ((rc = make_tmap_out(&tmC, C, p.c_f32, N, M, ldc))) return rc;
    } else {
        tmC = tmA;  // unused
    }
    int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
    if (p.wstat) {
        int cpp = num_sms() / p.n_tiles_n;  // CTAs per weight panel
        if (cpp < 1) cpp = 1;
        if (cpp > p.m_tiles) cpp = p.m_tiles;
        grid = cpp * p.n_tiles_n;
    }
    // specialised epilogues for the combinations the step uses (bulk-store path only); everything else reads its switches at run time
    const int kind = p.tma_store ? (epi | (accumulate ? 4 : 0) | (p.c_f32 ? 8 : 0)) : -1;
    // (set_smem keeps one flag per function TYPE, and all instances share theirs: a flag per kind here)
    static bool smem_set[17] = {};
#define ICK_TN_SMEM(KK)                                                                                                      \
    if (!smem_set[(KK) + 1]) {                                                                                               \
        if (cudaFuncSetAttribute(gemm_tn_tc_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) { \
            ick_set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize=%d) failed", SMEM_BYTES);                      \
            return ICK_ERR_CUDA;                                                                                             \
        }                                                                                                                    \
        smem_set[(KK) + 1] = true;                                                                                           \
    }
#define ICK_TN_CASE(KK)                                                                                   \
    case KK:                                                                                              \
        ICK_TN_SMEM(KK)                                                                                   \
        ick_launch(gemm_tn_tc_kernel<KK>, grid, NTHREADS, SMEM_BYTES, stream)(tmA, tmW, tmW2, tmC, p);    \
        break;
    switch (kind) {
        ICK_TN_CASE(0)   // bias / plain, bf16
        ICK_TN_CASE(1)   // ReLU + dropout
        ICK_TN_CASE(2)   // ReLU / dropout backward
        ICK_TN_CASE(3)   // attention row term
        ICK_TN_CASE(4)   // accumulate into bf16
        ICK_TN_CASE(8)   // fp32 output (vocabulary scores)
        default:
            ICK_TN_SMEM(-1)
            ick_launch(gemm_tn_tc_kernel<-1>, grid, NTHREADS, SMEM_BYTES, stream)(tmA, tmW, tmW2, tmC, p);
    }
#undef ICK_TN_CASE
#undef ICK_TN_SMEM
    return ick_check_launch("gemm_tn_tc");
}

extern "C" int ick_gemm_tn_tc(const void* A, const void* W, void* C, int c_dt, const float* bias, const void* aux, int M, int N, int K,
                              int lda, int ldw, int ldc, int ldaux, int epi, int accumulate, float drop_p, unsigned seed, unsigned site,
                              cudaStream_t stream) {
    return gemm_tn_tc_impl(A, W, nullptr, C, c_dt, bias, nullptr, aux, M, 0, N, K, lda, ldw, ldc, ldaux, epi, accumulate, drop_p, seed, site, 0u,
                           stream);
}

extern "C" int ick_gemm_tn_tc_dual(const void* A, const void* W, const void* W2, void* C, int c_dt, const float* bias, const float* bias2,
                                   const void* aux, int M, int M_split, int N, int K, int lda, int ldw, int ldc, int ldaux, int epi,
                                   int accumulate, float drop_p, unsigned seed, unsigned site, unsigned site2, cudaStream_t stream) {
    ICK_REQUIRE(W2 != nullptr, "gemm_tn_tc_dual: W2 is required");
    return gemm_tn_tc_impl(A, W, W2, C, c_dt, bias, bias2, aux, M, M_split, N, K, lda, ldw, ldc, ldaux, epi, accumulate, drop_p, seed, site, site2,
                           stream);
}

extern "C" int ick_gemm_tn_tc_rowdot(const void* A, const void* W, const void* W2, void* C, const void* O, float* dsum, float* dsum2, int M,
                                     int M_split, int rows0, int N, int K, int lda, int ldw, int ldc, int ldo, int S, int S2, int H,
                                     cudaStream_t stream) {
    ICK_REQUIRE(O != nullptr && dsum != nullptr && S > 0 && H > 0 && H * 32 <= N, "gemm_tn_tc_rowdot: bad row-dot arguments");
    ICK_REQUIRE(W2 == nullptr || (dsum2 != nullptr && S2 > 0 && rows0 > 0 && rows0 <= M_split), "gemm_tn_tc_rowdot: bad second row group");
    ICK_REQUIRE(rows0 % S == 0 && (W2 == nullptr ? rows0 == M : (M - M_split) % S2 == 0), "gemm_tn_tc_rowdot: rows are not whole sequences");
    RowDot rd{dsum, dsum2, S, S2, rows0, H};
    return gemm_tn_tc_impl(A, W, W2, C, ICK_BF16, nullptr, nullptr, O, M, M_split, N, K, lda, ldw, ldc, ldo, 3, 0, 0.f, 0u, 0u, 0u, stream, &rd);
}

extern "C" int ick_gemm_add_ln_tc(const void* A, const void* W, const void* W2, const float* bias, const float* bias2, const void* X, void* S,
                                  void* Y, float* mean, float* rstd, const float* gamma, const float* beta, const float* gamma2,
                                  const float* beta2, int M, int M_split, int K, int d, int lda, int ldw, int ldx, int lds, int ldy, float eps,
                                  float drop_p, unsigned seed, unsigned site, unsigned site2, cudaStream_t stream) {
    ICK_REQUIRE(M >= 0 && K > 0 && d > 0 && d <= LN_W && d % 2 == 0, "gemm_add_ln_tc: bad sizes M=%d K=%d d=%d", M, K, d);
    ICK_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && lds % 8 == 0 && ldy % 8 == 0 && lds >= LN_W && ldy >= LN_W,
                "gemm_add_ln_tc: leading dims must be multiples of 8 and s / y rows at least %d wide", LN_W);
    ICK_REQUIRE(((((uintptr_t)A) | ((uintptr_t)W) | ((uintptr_t)S) | ((uintptr_t)Y)) & 15) == 0, "gemm_add_ln_tc: operands must be 16-byte aligned");
    ICK_REQUIRE(mean != nullptr && rstd != nullptr && gamma != nullptr && beta != nullptr, "gemm_add_ln_tc: mean / rstd / gamma / beta are required");
    ICK_REQUIRE(bias == nullptr || (((uintptr_t)bias) & 15) == 0, "gemm_add_ln_tc: bias must be 16-byte aligned (and %d wide)", LN_W);
    if (W2 != nullptr) {
        ICK_REQUIRE(M_split > 0 && M_split % BM == 0 && M_split <= M, "gemm_add_ln_tc: M_split=%d must be a positive multiple of %d <= M", M_split, BM);
        ICK_REQUIRE((((uintptr_t)W2) & 15) == 0 && (bias == nullptr) == (bias2 == nullptr) && gamma2 != nullptr && beta2 != nullptr &&
                        (bias2 == nullptr || (((uintptr_t)bias2) & 15) == 0),
                    "gemm_add_ln_tc: bad second weight / bias / LayerNorm parameters");
    }
    if (M == 0) return ICK_OK;
    int rc = set_smem(gemm_ln_tc_kernel);
    if (rc) return rc;
    LnParams p;
    p.X = (const bf16*)X; p.mean = mean; p.rstd = rstd;
    p.bias = bias; p.bias2 = W2 != nullptr ? bias2 : bias;
    p.gamma = gamma; p.gamma2 = W2 != nullptr ? gamma2 : gamma;
    p.beta = beta; p.beta2 = W2 != nullptr ? beta2 : beta;
    p.M = M; p.D = d; p.ldx = ldx;
    p.nkb = (K + BK - 1) / BK;
    p.m_tiles = (M + BM - 1) / BM;
    p.split_tile = W2 != nullptr ? M_split / BM : 0x7FFFFFFF;
    p.eps = eps; p.inv_d = 1.0f / (float)d;
    p.site2 = site2;
    p.drop = make_drop(drop_p, seed, site);
    p.stages = (SMEM_DATA - NEPI * LN_NSBOX * 2048 - LN_RED_BYTES) / LN_STAGE;
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    ICK_REQUIRE(p.stages >= 2, "gemm_add_ln_tc: stages do not fit");
    CUtensorMap tmA, tmW, tmW2, tmS, tmY;
    if ((rc = make_tmap(&tmA, A, K, M, lda, BM))) return rc;
    if ((rc = make_tmap(&tmW, W, K, LN_W, ldw, LN_HALF))) return rc;
    if (W2 != nullptr) {
        if ((rc = make_tmap(&tmW2, W2, K, LN_W, ldw, LN_HALF))) return rc;
    } else {
        tmW2 = tmW;
    }
    if ((rc = make_tmap_out(&tmS, S, 0, LN_W, M, lds))) return rc;
    if ((rc = make_tmap_out(&tmY, Y, 0, LN_W, M, ldy))) return rc;
    const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
    ick_launch(gemm_ln_tc_kernel, grid, NTHREADS, SMEM_BYTES, stream)(tmA, tmW, tmW2, tmS, tmY, p);
    return ick_check_launch("gemm_add_ln_tc");
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
    // the bias gradient rides along as 16 extra accumulator columns when they fit behind the tile (KT + 32 <= 512 TMEM columns)
    const bool fuse_bias = biasoff != nullptr && p.n_tiles_k == 1 && p.KT + 32 <= TMEM_COLS;
    p.biasoff = fuse_bias ? biasoff : nullptr;
    p.stages = (SMEM_DATA - (fuse_bias ? WG_BOX : 0)) / stage_bytes;
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    ICK_REQUIRE(p.stages >= 2, "wgrad_tc: tile does not fit");
    // Row split: every (tile, split) work item costs a pipeline fill + an epilogue, a launch runs ceil(items / SMs) waves of
    // them, and every split adds a partial tile that the reduce kernel must read back.  Pick the split count that minimises
    //   waves * (k-blocks per item * t_kb + t_fix) + splits * (partial-tile bytes / reduce bandwidth)
    // (t_kb scaled by the stage size; constants from the B200 microbenchmarks in profiles/).  The naive ceil(SMs / tiles)
    // put 152 items on 148 SMs for the 960-wide projections: two waves for one wave's worth of work.
    p.Kws = (K + 31) / 32 * 32;
    {
        const int max_splits = (M + 8 * BK - 1) / (8 * BK);  // at least 8 k-blocks per work item
        // a k-block stage of 128 + KT columns x 64 rows enters shared memory at ~42 B/clk per SM when every SM streams (the
        // L2 -> SM cap, B300_MICROARCH.md), the partial tiles are re-read from L2 at ~4 TB/s
        const double t_kb = 0.9 * (128 + p.KT) / 448.0, t_fix = 4.0;          // microseconds
        const double t_red = (double)N * p.Kws * 4 / 4.0e6;                    // per split
        double best = 1e30;
        int best_mps = (M + BK - 1) / BK * BK;
        for (int sp = 1; sp <= (max_splits > 0 ? max_splits : 1); ++sp) {
            int mps = (M + sp - 1) / sp;
            mps = (mps + BK - 1) / BK * BK;
            const int real = (M + mps - 1) / mps;
            const int waves = (tiles * real + num_sms() - 1) / num_sms();
            const double t = waves * ((mps / BK) * t_kb + t_fix) + real * t_red;
            if (t < best - 1e-9) { best = t; best_mps = mps; }
        }
        p.m_per_split = best_mps;
        p.splits = (M + best_mps - 1) / best_mps;
    }
    const long long need = (long long)p.splits * N * (p.Kws + 1) * 4;
    p.ws = (workspace != nullptr && workspace_bytes >= need && (((uintptr_t)workspace) & 15) == 0) ? (float*)workspace : nullptr;
    p.wsb = (p.ws != nullptr && fuse_bias) ? p.ws + (size_t)p.splits * N * p.Kws : nullptr;
    CUtensorMap tmY, tmX;
    if ((rc = make_tmap(&tmY, dY, N, M, ldy, BK))) return rc;
    if ((rc = make_tmap(&tmX, X, K, M, ldx, BK))) return rc;
    const int n_work = tiles * p.splits;
    const int grid = n_work < num_sms() ? n_work : num_sms();
    ick_launch(wgrad_tc_kernel, grid, NTHREADS, SMEM_BYTES, stream)(tmY, tmX, p);
    if ((rc = ick_check_launch("wgrad_tc"))) return rc;
    if (p.ws != nullptr) {
        const long long nthreads = (long long)N * (p.Kws / 4);
        const int rgrid = (int)((nthreads + 255) / 256);
        ick_launch(wgrad_reduce_kernel, rgrid, 256, 0, stream)(p.ws, gflat, rowoff, colmap, N, K, p.Kws, p.splits, p.wsb, biasoff);
        if ((rc = ick_check_launch("wgrad_reduce"))) return rc;
    }
    if (biasoff && !fuse_bias) {
        const int rpb = 256;
        dim3 bgrid((N + 127) / 128, (M + rpb - 1) / rpb);
        ick_launch(bias_grad_kernel, bgrid, 256, 0, stream)((const bf16*)dY, gflat, biasoff, M, N, ldy, rpb);
        return ick_check_launch("bias_grad");
    }
    return ICK_OK;
}

extern "C" int ick_wgrad_group_tc(int nprob, const void* const* dY, const void* const* X, const int* const* rowoff, const int* const* colmap,
                                  const int* const* biasoff, const int* M, const int* N, const int* K, const int* ldy, const int* ldx,
                                  float* gflat, void* workspace, long long workspace_bytes, cudaStream_t stream) {
    ICK_REQUIRE(nprob >= 1 && nprob <= WG_MAXP, "wgrad_group_tc: 1..%d problems, got %d", WG_MAXP, nprob);
    ICK_REQUIRE(workspace != nullptr && (((uintptr_t)workspace) & 15) == 0, "wgrad_group_tc: a 16-byte aligned workspace is required");
    int rc = set_smem(wgrad_group_tc_kernel);
    if (rc) return rc;
    WgGroup g;  // 2.8 KB: built on the host, passed by value as the kernel parameter
    g.G = gflat;
    g.nprob = nprob;
    int tiles[WG_MAXP], kbtot[WG_MAXP];
    long long units = 0;
    int tiles_total = 0, max_kb = 0;
    for (int i = 0; i < nprob; ++i) {
        ICK_REQUIRE(M[i] > 0 && N[i] > 0 && K[i] > 0, "wgrad_group_tc: bad sizes of problem %d: M=%d N=%d K=%d", i, M[i], N[i], K[i]);
        ICK_REQUIRE(ldy[i] % 8 == 0 && ldx[i] % 8 == 0, "wgrad_group_tc: ldy/ldx must be multiples of 8");
        ICK_REQUIRE((((uintptr_t)dY[i]) & 15) == 0 && (((uintptr_t)X[i]) & 15) == 0, "wgrad_group_tc: operands must be 16-byte aligned");
        ICK_REQUIRE(rowoff[i] != nullptr, "wgrad_group_tc: rowoff is required");
        WgProb& p = g.pr[i];
        p.rowoff = rowoff[i]; p.colmap = colmap[i]; p.biasoff = biasoff[i];
        p.M = M[i]; p.N = N[i]; p.K = K[i];
        const int Kpad = (K[i] + 63) / 64 * 64;
        p.n_tiles_k = (Kpad + WG_GKT - 1) / WG_GKT;
        p.KT = ((Kpad / 64 + p.n_tiles_k - 1) / p.n_tiles_k) * 64;
        p.Kws = (K[i] + 31) / 32 * 32;
        tiles[i] = ((N[i] + BM - 1) / BM) * p.n_tiles_k;
        kbtot[i] = (M[i] + BK - 1) / BK;
        units += (long long)tiles[i] * kbtot[i];
        tiles_total += tiles[i];
        if (kbtot[i] > max_kb) max_kb = kbtot[i];
        if ((rc = make_tmap(&g.tmY[i], dY[i], N[i], M[i], ldy[i], BK))) return rc;
        if ((rc = make_tmap(&g.tmX[i], X[i], K[i], M[i], ldx[i], BK))) return rc;
    }
    // one wave: the smallest k-blocks-per-item (>= 8) for which all work items fit on the SMs at once
    int kbt = (int)((units + num_sms() - 1) / num_sms());
    if (kbt < 8) kbt = 8;
    for (;; ++kbt) {
        int items = 0;
        for (int i = 0; i < nprob; ++i) items += tiles[i] * ((kbtot[i] + kbt - 1) / kbt);
        if (items <= num_sms() || kbt >= max_kb) break;
    }
    int item0 = 0, q0 = 0;
    size_t ws_off = 0;
    for (int i = 0; i < nprob; ++i) {
        WgProb& p = g.pr[i];
        p.m_per_split = kbt * BK;
        p.splits = (p.M + p.m_per_split - 1) / p.m_per_split;
        p.item0 = item0;
        item0 += tiles[i] * p.splits;
        p.q0 = q0;
        q0 += p.N * (p.Kws / 4);
        p.ws = (float*)workspace + ws_off;
        ws_off += (size_t)p.splits * p.N * p.Kws;
        p.wsb = nullptr;
        if (p.biasoff != nullptr) {
            p.wsb = (float*)workspace + ws_off;
            ws_off += ((size_t)p.splits * p.N + 3) / 4 * 4;
        }
    }
    ICK_REQUIRE((long long)ws_off * 4 <= workspace_bytes, "wgrad_group_tc: workspace too small (%lld > %lld bytes)", (long long)ws_off * 4,
                workspace_bytes);
    g.n_items = item0;
    g.n_q4 = q0;
    g.stages = (SMEM_DATA - WG_BOX) / (A_STAGE + (WG_GKT / 64) * WG_BOX);
    if (g.stages > MAX_STAGES) g.stages = MAX_STAGES;
    ICK_REQUIRE(g.stages >= 2, "wgrad_group_tc: stage does not fit");
    const int grid = g.n_items < num_sms() ? g.n_items : num_sms();
    ick_launch(wgrad_group_tc_kernel, grid, NTHREADS, SMEM_BYTES, stream)(g);
    if ((rc = ick_check_launch("wgrad_group_tc"))) return rc;
    ick_launch(wgrad_group_reduce_kernel, (g.n_q4 + 255) / 256, 256, 0, stream)(g);
    return ick_check_launch("wgrad_group_reduce");
}
