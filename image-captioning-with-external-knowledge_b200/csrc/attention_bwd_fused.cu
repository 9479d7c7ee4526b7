// Fused bf16 attention backward for head_dim <= 32: ONE persistent kernel per call produces dQ, dK and dV.
//
// The previous scheme ran two kernels that each recomputed S, P, dP and the dropout mask (dK/dV with key ownership, dQ with
// query ownership) - or, for the long self-attentions, spilled dS^T to HBM for a separate dQ GEMM (2.2x the algorithmic DRAM
// traffic).  Both are gone:
//
//   * a compute warp owns 16 keys of an (image, head) item (slabs dealt round-robin over the warps ACROSS items, so the warps
//     stay balanced over the launch) and walks the item's queries in blocks of 32: S^T = K Q^T and dP^T = V dO^T on
//     mma.sync, P^T / dS^T in registers, dV += P^T dO and dK += dS^T Q accumulated in registers (the operands Q, dO and K are
//     staged once per item by TMA, 64B-swizzled, two or more items in flight);
//   * dQ needs the sum over ALL key slabs, i.e. over warps.  Each warp drops its dS^T block - [16 keys][128 queries] bf16, written
//     in the canonical MN-major SWIZZLE_128B layout (two 64-query atoms) - into a private 4 KiB slot of shared memory, and one
//     elected thread of the issuer warp feeds it to the 5th-generation tensor core:
//         dQ[128 queries x 32] += dS[128 x 16 keys] * K[16 keys x 32]        (tcgen05.mma, M = 128, N = 32, K = 16)
//     with the B operand = the 16 key rows inside the TMA-staged K tile (MN-major, SWIZZLE_64B) and the fp32 accumulator in
//     TENSOR MEMORY, one 32-column slice per 128 queries of the item; up to four items' dQ live in TMEM at once, so the warps
//     never wait for each other;
//   * four epilogue warps read an item's dQ out of TMEM (tcgen05.ld) when its last block has been issued, scale and store it,
//     and clear the accumulator (tcgen05.st) for the slot's next tenant, so that every MMA simply accumulates;
//   * attention-probability dropout is the bit-parallel keep-word scheme of common.cuh: two helper warps hash the item's
//     (key group, query) keep words into the stage while the previous item computes - ONE hash per 32 probabilities at the
//     reference's p = 0.5 - and the compute warps test one bit per probability.
//
// Warp roles (24 warps, one CTA per SM): 0 TMA producer + per-query scalars, 1 tcgen05 issuer (+ TMEM allocation),
// 2-3 keep-word generators, 4-7 dQ epilogue, 8-23 compute.  setmaxnreg moves registers from the helper warpgroups to the
// compute warpgroups (104 per thread).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "attention_internal.h"
#include "attention_mma.cuh"

namespace {
using namespace ickattn;

constexpr int FB_NCW = 16;                 // compute warps
constexpr int FB_FIRST_CW = 8;             // first compute warp
constexpr int FB_WARPS = FB_FIRST_CW + FB_NCW;
constexpr int FB_THREADS = 32 * FB_WARPS;  // 768
constexpr int FB_MAXNI = 4;                // items whose dQ accumulators live in TMEM
constexpr int FB_MAXSTAGE = 6;
constexpr int FB_TMEM_COLS = 512;
constexpr int FB_SLOT = 4096;              // one dS^T block: [16 keys][128 queries] bf16 = two 64-query SWIZZLE_128B atoms of 2 KiB
constexpr int FB_MAXSLOT = 2;              // ring slots per compute warp
constexpr int FB_BAR_BYTES = 2048;
constexpr int FB_SMEM_MAX = 232448;

struct FArgs {
    Dims d;
    int nslabs;          // 16-key slabs of an item
    int ntq, ntk;        // 64-row tiles of Q / dO and of K
    int nq128;           // 128-query accumulator tiles of an item
    int ngroups;         // 32-key groups (keep words per query)
    int nstage, ni;      // operand stages, TMEM item slots
    int nslot;           // ring slots per compute warp (1 or 2)
    uint32_t stage_bytes, off_do, off_k, off_ls, off_ds, off_mw;
    int dbg;             // ICK_FB_DEBUG bit mask (bring-up aid): 1 no setmaxnreg, 2 no tcgen05.mma, 4 no tcgen05.ld, 8 no block math, 16 one ring slot
};

// ---- tcgen05 wrappers --------------------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_st16_zero(uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// non-blocking probe (mbarrier.try_wait may suspend the thread up to a system time limit: useless for polling many barriers)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version 1 [46,48), layout type [61,64)
// (2 = SWIZZLE_128B, 4 = SWIZZLE_64B).  MN-major operands: LBO = byte distance of the swizzle atoms along M / N, SBO = byte
// distance of the 8-row groups along K.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both MN-major, M = 128, N = 32
constexpr uint32_t FB_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

struct FSm {
    uint32_t base;  // shared-space address of the 1024-aligned region
    uint8_t* gen;   // generic pointer to the same byte
    uint32_t stage_bytes, ring_bytes;
    int nslot;
    __device__ __forceinline__ uint32_t bar(int i) const { return base + 8u * (uint32_t)i; }
    __device__ __forceinline__ uint32_t full(int s) const { return bar(s); }
    __device__ __forceinline__ uint32_t sfull(int s) const { return bar(FB_MAXSTAGE + s); }
    __device__ __forceinline__ uint32_t empty(int s) const { return bar(2 * FB_MAXSTAGE + s); }
    __device__ __forceinline__ uint32_t rfull(int w, int k) const { return bar(3 * FB_MAXSTAGE + FB_MAXSLOT * w + k); }
    __device__ __forceinline__ uint32_t rempty(int w, int k) const { return bar(3 * FB_MAXSTAGE + FB_MAXSLOT * FB_NCW + FB_MAXSLOT * w + k); }
    __device__ __forceinline__ uint32_t dqfull(int i) const { return bar(3 * FB_MAXSTAGE + 2 * FB_MAXSLOT * FB_NCW + i); }
    __device__ __forceinline__ uint32_t dqempty(int i) const { return bar(3 * FB_MAXSTAGE + 2 * FB_MAXSLOT * FB_NCW + FB_MAXNI + i); }
    __device__ __forceinline__ uint32_t* tmem_ptr() const { return reinterpret_cast<uint32_t*>(gen + 1024); }
    __device__ __forceinline__ volatile uint32_t* meta() const { return reinterpret_cast<volatile uint32_t*>(gen + 1040); }  // [warp][slot]
    __device__ __forceinline__ uint32_t slot(int w, int k) const { return base + FB_BAR_BYTES + (uint32_t)(w * nslot + k) * FB_SLOT; }
    __device__ __forceinline__ uint32_t stage(int s) const { return base + FB_BAR_BYTES + ring_bytes + (uint32_t)s * stage_bytes; }
    __device__ __forceinline__ uint8_t* stage_gen(int s) const { return gen + FB_BAR_BYTES + ring_bytes + (size_t)s * stage_bytes; }
};

__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// One 32-query block of the warp's 16-key slab: S^T and dP^T on mma.sync, P^T / dS^T in registers, dV += P^T dO, dK += dS^T Q,
// and the dS^T block into query columns [32*sub4, 32*sub4 + 32) of the warp's current ring slot (MN-major SWIZZLE_128B: 64-query
// atom sub4 >> 1 at +2048, key row r at r*128, 16-byte chunk c of the row at chunk c ^ (r & 7)).  `sub` = sub4 & 1 is also the
// half of the 64-row Q / dO tile the block reads.
//   ls / ds / mw: SHARED-space addresses of the per-query LSE (log2 domain), D and keep word of this block's queries, already
//   offset by the lane's 2*tq.  MASK: the block touches the causal diagonal of the slab.
template <bool DROP, bool MASK>
__device__ __forceinline__ void fb_block(float (*dk)[4], float (*dv)[4], uint32_t kaddr, const uint32_t (*va)[4], uint32_t qt,
                                         uint32_t gt, int sub, int q0, uint32_t ls, uint32_t ds, uint32_t mw, uint32_t mk0, uint32_t mk1,
                                         const OwnRows& r, const TileEnv& e, uint32_t slot_lane, int g) {
    const int tq = e.tq;
    float st[4][4], dpt[4][4];
    {   // the slab's K rows as A fragments, re-read from the staged tile every block (8 registers that need not stay live)
        uint32_t ka[2][4];
        ldsm_x4(ka[0][0], ka[0][1], ka[0][2], ka[0][3], kaddr);
        ldsm_x4(ka[1][0], ka[1][1], ka[1][2], ka[1][3], kaddr ^ 32u);
        mma_a_tT<4>(st, ka, qt, 4 * sub, e.lo);
    }
    mma_a_tT<4>(dpt, va, gt, 4 * sub, e.lo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 l2 = lds_f2(ls + 32 * j);
        const float2 d2 = lds_f2(ds + 32 * j);
        uint2 w2 = make_uint2(0u, 0u);
        if (DROP) w2 = lds_u2(mw + 32 * j);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const bool odd_q = (x & 1) != 0, hi_k = x >= 2;
            float p = ex2(fmaf(st[j][x], e.c, -(odd_q ? l2.y : l2.x)));
            if (MASK && (hi_k ? r.r1 : r.r0) > q0 + 8 * j + 2 * tq + (x & 1)) p = 0.f;
            const float nd = -(odd_q ? d2.y : d2.x);
            if (DROP) {
                const bool keep = ((odd_q ? w2.y : w2.x) & (hi_k ? mk1 : mk0)) != 0u;
                const float t = fmaf(dpt[j][x], e.ik, nd);
                dpt[j][x] = p * (keep ? t : nd);  // dS^T = P (drop(dP) - D)
                st[j][x] = keep ? p : 0.f;        // P^T with dropout (1/keep applied when dV is stored)
            } else {
                dpt[j][x] = p * (dpt[j][x] + nd);
                st[j][x] = p;
            }
        }
    }
    uint32_t pa[2][4];
    pack_p<4>(st, pa);
    mma_p_t<4>(dv, pa, gt, 4 * sub, e.lo);
    pack_p<4>(dpt, pa);
    const uint32_t xg = (uint32_t)(g ^ (sub << 2));
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const uint32_t a = slot_lane + ((xg ^ (uint32_t)(2 * kk + jj)) << 4);
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(pa[kk][2 * jj]) : "memory");             // key row g
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(a + 1024u), "r"(pa[kk][2 * jj + 1]) : "memory");  // key row g + 8
        }
    mma_p_t<4>(dk, pa, qt, 4 * sub, e.lo);
}
// a skipped block (causal: every query precedes every key of the slab) still owns its columns of the slot: zeros
__device__ __forceinline__ void fb_zero_block(int sub, uint32_t slot_lane, int g) {
    const uint32_t xg = (uint32_t)(g ^ (sub << 2));
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t a = slot_lane + ((xg ^ (uint32_t)c) << 4);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(0u) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(a + 1024u), "r"(0u) : "memory");
    }
}

// ---- warp roles -----------------------------------------------------------------------------------------------------------------------
struct FPtrs {
    const bf16* V;
    const float* LSE;
    const float* Dsum;
    bf16 *dQ, *dK, *dV;
    int ldv, lddq, lddk, lddv;
};

// warp 0: TMA producer + per-query scalars
__device__ __forceinline__ void fb_producer(const FSm& sm, const FArgs& a, int lane, const CUtensorMap* tmQ, const CUtensorMap* tmG,
                                            const CUtensorMap* tmK, const FPtrs& p, int n_items) {
    const Dims& d = a.d;
    int li = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
        const int s = li % a.nstage, b = item / d.H, h = item % d.H;
        mbar_wait(sm.empty(s), ((uint32_t)(li / a.nstage) & 1u) ^ 1u);
        if (lane == 0) {
            const uint32_t st = sm.stage(s), bar = sm.full(s);
            mbar_expect_tx(bar, (uint32_t)(2 * a.ntq + a.ntk) * TILE_BYTES);
            for (int t = 0; t < a.ntq; ++t) {
                tma_load_3d(st + t * TILE_BYTES, tmQ, bar, h * HD, t * TK, b);
                tma_load_3d(st + a.off_do + t * TILE_BYTES, tmG, bar, h * HD, t * TK, b);
            }
            for (int t = 0; t < a.ntk; ++t) tma_load_3d(st + a.off_k + t * TILE_BYTES, tmK, bar, h * HD, t * TK, b);
        }
        float* Ls = reinterpret_cast<float*>(sm.stage_gen(s) + a.off_ls);
        float* Ds = reinterpret_cast<float*>(sm.stage_gen(s) + a.off_ds);
        const float* L = p.LSE + ((size_t)b * d.H + h) * d.Sq;
        const float* Dg = p.Dsum + ((size_t)b * d.H + h) * d.Sq;
        for (int i = lane; i < a.ntq * TK; i += 32) {
            const bool ok = i < d.Sq;
            Ls[i] = ok ? L[i] : 0.f;
            Ds[i] = ok ? Dg[i] : 0.f;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sm.sfull(s));
    }
}

// warp 1: tcgen05 issuer, dQ[item, 128-query tile] += dS block * K slab.  Lane w < FB_NCW watches compute warp w's ring slots
// (non-blocking probes) and prepares the two descriptors of a ready block; lane 0 issues the MMAs one after the other (one thread:
// MMAs into the same accumulator stay ordered) and commits each block's slot back to its warp.  The accumulators were cleared by
// the epilogue warps, so every MMA accumulates; the only state is a per-TMEM-slot block count (four 8-bit counters in a register).
__device__ __forceinline__ void fb_issuer(const FSm& sm, const FArgs& a, int lane, uint32_t tmem_base, int my_items) {
    uint32_t nb = 0;  // lane w < FB_NCW: blocks consumed from compute warp w
    const uint32_t per_item = (uint32_t)(a.nslabs * a.nq128);
    const long long total = (long long)my_items * per_item;
    long long done = 0;
    uint32_t cnt = 0;  // lane 0: blocks issued so far for the tenant of each TMEM slot (8 bits each)
    const uint32_t ad_hi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024, version 1, SWIZZLE_128B
    const uint32_t bd_hi = (512u >> 4) | (1u << 14) | (4u << 29);   // SBO = 512, version 1, SWIZZLE_64B
    long long spin0 = clock64();
    while (done < total) {
        bool ready = false;
        uint32_t ad_lo = 0, bd_lo = 0, info = 0;
        if (lane < FB_NCW) {
            const uint32_t k = a.nslot == 2 ? (nb & 1u) : 0u, use = a.nslot == 2 ? (nb >> 1) : nb;
            ready = mbar_test(sm.rfull(lane, (int)k), use & 1u);
            if (ready) {
                const uint32_t m = sm.meta()[FB_MAXSLOT * lane + k];
                const uint32_t mli = m & 0xFFFFu, t = (m >> 16) & 15u, slab = (m >> 20) & 63u;
                const uint32_t i = mli % (uint32_t)a.ni, st = mli % (uint32_t)a.nstage;
                ad_lo = ((sm.slot(lane, (int)k) >> 4) & 0x3FFFu) | ((2048u >> 4) << 16);  // the two 64-query atoms are 2 KiB apart
                bd_lo = (((sm.stage((int)st) + a.off_k + slab * 1024u) >> 4) & 0x3FFFu) | ((512u >> 4) << 16);
                info = ((i * (uint32_t)a.nq128 + t) * 32u) | (i << 16) | (st << 20) | (k << 24);
            }
        }
        uint32_t mask = __ballot_sync(0xffffffffu, ready);
        if (mask == 0u) {
            if (clock64() - spin0 > 8000000000LL) __trap();  // a protocol bug must not hang the GPU
            __nanosleep(32);
            continue;
        }
        tc_fence_after();
        while (mask != 0u) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1u;
            const uint32_t al = __shfl_sync(0xffffffffu, ad_lo, src), bl = __shfl_sync(0xffffffffu, bd_lo, src);
            const uint32_t inf = __shfl_sync(0xffffffffu, info, src);
            if (lane == 0) {
                const uint32_t i = (inf >> 16) & 15u, st = (inf >> 20) & 15u, k = (inf >> 24) & 1u;
                if (!(a.dbg & 2))
                    tc_mma_bf16(tmem_base + (inf & 0xFFFFu), ((uint64_t)ad_hi << 32) | al, ((uint64_t)bd_hi << 32) | bl, FB_IDESC, 1u);
                tc_commit(sm.rempty(src, (int)k));
                cnt += 1u << (8 * i);
                if (((cnt >> (8 * i)) & 0xFFu) == per_item) {
                    cnt &= ~(0xFFu << (8 * i));
                    tc_commit(sm.dqfull((int)i));  // every block of the item has been accumulated
                    tc_commit(sm.empty((int)st));  // and its K tile is no longer read
                }
            }
            if (lane == src) ++nb;
            ++done;
        }
        __syncwarp();
        spin0 = clock64();
    }
}

// warps 2-3: keep words of the item, Mw[key group][query]
__device__ __forceinline__ void fb_maskgen(const FSm& sm, const FArgs& a, int warp, int lane, DropCfg drop, int n_items) {
    const Dims& d = a.d;
    ick_resolve_seed(drop);
    const uint32_t t16 = ick_attn_t16(drop.thr);
    const int nq = a.ntq * TK;
    int li = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
        const int s = li % a.nstage, b = item / d.H, h = item % d.H;
        mbar_wait(sm.empty(s), ((uint32_t)(li / a.nstage) & 1u) ^ 1u);
        uint32_t* Mw = reinterpret_cast<uint32_t*>(sm.stage_gen(s) + a.off_mw);
        for (int q = (warp - 2) * 32 + lane; q < nq; q += 64) {
            const uint32_t rm = ick_rowmix(drop.seed, drop.site, prob_row(b, d.H, h, d.Sq, q));
            for (int kg = 0; kg < a.ngroups; ++kg) Mw[kg * nq + q] = q < d.Sq ? ick_keepword(rm, (uint32_t)kg, t16) : 0u;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sm.sfull(s));
    }
}

// warps 4-7: dQ epilogue, TMEM -> registers -> scale -> bf16 rows, then clear the accumulator for the slot's next tenant
__device__ __forceinline__ void fb_epilogue(const FSm& sm, const FArgs& a, int warp, int lane, uint32_t tmem_base, const FPtrs& p, int n_items) {
    const Dims& d = a.d;
    const int qd = warp & 3;
    const uint32_t tlane = tmem_base + ((uint32_t)(qd * 32) << 16);
    // phase 0 of every dqempty barrier: the slots start out cleared
    for (int c = 0; c < a.ni * a.nq128 * 2; ++c) tc_st16_zero(tlane + 16u * (uint32_t)c);
    tc_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0)
        for (int i = 0; i < a.ni; ++i) mbar_arrive(sm.dqempty(i));
    int li = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
        const int i = li % a.ni, b = item / d.H, h = item % d.H;
        mbar_wait(sm.dqfull(i), (uint32_t)(li / a.ni) & 1u);
        tc_fence_after();
        bf16* out = p.dQ + (size_t)b * d.Sq * p.lddq + h * HD;
        for (int t = 0; t < a.nq128; ++t) {
            const int q = t * 128 + qd * 32 + lane;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t r[16];
                const uint32_t ta = tlane + (uint32_t)((i * a.nq128 + t) * 32 + 16 * c);
                if (!(a.dbg & 4)) {
                    tc_ld16(ta, r);
                } else {
#pragma unroll
                    for (int z = 0; z < 16; ++z) r[z] = 0u;
                }
                tc_st16_zero(ta);
                if (q < d.Sq) {
                    uint4 u0, u1;
                    u0.x = pack2(__uint_as_float(r[0]) * d.scale, __uint_as_float(r[1]) * d.scale);
                    u0.y = pack2(__uint_as_float(r[2]) * d.scale, __uint_as_float(r[3]) * d.scale);
                    u0.z = pack2(__uint_as_float(r[4]) * d.scale, __uint_as_float(r[5]) * d.scale);
                    u0.w = pack2(__uint_as_float(r[6]) * d.scale, __uint_as_float(r[7]) * d.scale);
                    u1.x = pack2(__uint_as_float(r[8]) * d.scale, __uint_as_float(r[9]) * d.scale);
                    u1.y = pack2(__uint_as_float(r[10]) * d.scale, __uint_as_float(r[11]) * d.scale);
                    u1.z = pack2(__uint_as_float(r[12]) * d.scale, __uint_as_float(r[13]) * d.scale);
                    u1.w = pack2(__uint_as_float(r[14]) * d.scale, __uint_as_float(r[15]) * d.scale);
                    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)q * p.lddq + 16 * c);
                    dst[0] = u0;
                    dst[1] = u1;
                }
            }
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sm.dqempty(i));
    }
}

// warps 8-23: compute
template <bool DROP>
__device__ __forceinline__ void fb_compute(const FSm& sm, const FArgs& a, int warp, int lane, const FPtrs& p, DropCfg drop, int n_items) {
    const Dims& d = a.d;
    if (DROP) ick_resolve_seed(drop);
    const int cw = warp - FB_FIRST_CW, g = lane >> 2, tq = lane & 3;
    const TileEnv env = make_env(d, drop, lane);
    // ldmatrix (non-trans) addresses of a 16-row A fragment inside a 64B-swizzled tile: matrix m = lane >> 3 holds rows
    // 8*(m & 1) + (lane & 7), 16-byte chunk 2*ks + (m >> 1)
    // (k-step 1 is chunk + 2: the same address with bit 5 flipped)
    uint32_t aoff0;
    {
        const int m = lane >> 3, row = 8 * (m & 1) + (lane & 7);
        aoff0 = (uint32_t)(row * 64 + (((m >> 1) ^ ((row >> 1) & 3)) << 4));
    }
    const int nq = a.ntq * TK;
    uint32_t nb = 0;  // blocks this warp has produced (selects the half slot and its phase)
    int li = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
        const int s = li % a.nstage, b = item / d.H, h = item % d.H;
        const uint32_t ph = (uint32_t)(li / a.nstage) & 1u;
        const uint32_t tQ = sm.stage(s), tG = tQ + a.off_do, tK = tQ + a.off_k;
        // shared-space addresses of this lane's first (query pair) scalars: LSE, D, keep words of key group 0
        const uint32_t ls0 = tQ + a.off_ls + 8u * tq, ds0 = tQ + a.off_ds + 8u * tq, mw0 = tQ + a.off_mw + 8u * tq;
        // every warp passes through every item in order, so no warp can release a stage before the producer has filled it
        mbar_wait(sm.full(s), ph);
        mbar_wait(sm.sfull(s), ph);
        bool slot_checked = false;
        for (int slab = first_slab<FB_NCW>(li, a.nslabs, cw); slab < a.nslabs; slab += FB_NCW) {
            if (!slot_checked) {
                // The item's dQ accumulators must have been read out and cleared since the TMEM slot's previous tenant.  ONLY a warp
                // that contributes blocks to the item may wait here: the item cannot complete without it, so the barrier is at most
                // one phase ahead.  (A warp without a slab could arrive after the item has come and gone - two phases later the
                // parity test reads "not yet" for ever.)
                mbar_wait(sm.dqempty(li % a.ni), (uint32_t)(li / a.ni) & 1u);
                slot_checked = true;
            }
            const OwnRows r = own_rows(16 * slab, g, drop, b, d.H, h, d.Sq, false);
            uint32_t va[2][4];
            const uint32_t kaddr = tK + (uint32_t)slab * 1024u + aoff0;
            load_own(p.V + (size_t)b * d.Sk * p.ldv + h * HD, p.ldv, r.r0, r.r1, d.Sk, d.dh, tq, va);
            float dk[4][4], dv[4][4];
            zero16(dk);
            zero16(dv);
            const uint32_t mk0 = 1u << ick_keybit((uint32_t)r.r0), mk1 = mk0 << 4;
            const uint32_t mws = mw0 + (uint32_t)((slab >> 1) * nq) * 4u;
            for (int t = 0; t < a.nq128; ++t) {
                const uint32_t k = a.nslot == 2 ? (nb & 1u) : 0u, use = a.nslot == 2 ? (nb >> 1) : nb;
                mbar_wait(sm.rempty(cw, (int)k), (use & 1u) ^ 1u);
                const uint32_t slot_lane = sm.slot(cw, (int)k) + (uint32_t)(g * 128 + tq * 4);
#pragma unroll 1
                for (int sub4 = 0; sub4 < 4; ++sub4) {
                    const int q0 = t * 128 + sub4 * SUB;
                    if (q0 >= d.Sq) break;  // columns of queries that do not exist only feed dQ rows that are never stored
                    const int sub = sub4 & 1;
                    const uint32_t qo = 4u * (uint32_t)q0, sl = slot_lane + (uint32_t)(sub4 >> 1) * 2048u;
                    const uint32_t qt = tQ + (uint32_t)(q0 >> 6) * TILE_BYTES, gt = tG + (uint32_t)(q0 >> 6) * TILE_BYTES;
                    if ((d.causal && q0 + SUB - 1 < r.wrow) || (a.dbg & 8)) {
                        fb_zero_block(sub, sl, g);  // every query precedes every key of the slab
                    } else if (d.causal && r.wrow + 15 > q0) {
                        fb_block<DROP, true>(dk, dv, kaddr, va, qt, gt, sub, q0, ls0 + qo, ds0 + qo, mws + qo, mk0, mk1, r, env, sl, g);
                    } else {
                        fb_block<DROP, false>(dk, dv, kaddr, va, qt, gt, sub, q0, ls0 + qo, ds0 + qo, mws + qo, mk0, mk1, r, env, sl, g);
                    }
                }
                fence_async_smem();  // the block is read by the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) {
                    sm.meta()[FB_MAXSLOT * cw + k] = (uint32_t)(li & 0xFFFF) | ((uint32_t)t << 16) | ((uint32_t)slab << 20);
                    mbar_arrive(sm.rfull(cw, (int)k));
                }
                ++nb;
            }
            store_slab(p.dK + (size_t)b * d.Sk * p.lddk + h * HD, p.lddk, r.r0, r.r1, d.Sk, dk, d.scale, d.scale, d.dh, tq);
            store_slab(p.dV + (size_t)b * d.Sk * p.lddv + h * HD, p.lddv, r.r0, r.r1, d.Sk, dv, env.ik, env.ik, d.dh, tq);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sm.empty(s));
    }
}

template <bool DROP>
__global__ void __launch_bounds__(FB_THREADS, 1)
    bwd_fused_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmK,
                     FPtrs p, FArgs a, DropCfg drop) {
    ick_pdl_launch();
    extern __shared__ uint8_t smem_raw[];
    const Dims& d = a.d;
    FSm sm;
    sm.gen = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    sm.base = smem_u32(sm.gen);
    sm.stage_bytes = a.stage_bytes;
    sm.nslot = a.nslot;
    sm.ring_bytes = (uint32_t)(FB_NCW * a.nslot * FB_SLOT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = d.B * d.H;
    const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    // ---- one-time setup ---------------------------------------------------------------------------------------------------------------
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
        for (int s = 0; s < a.nstage; ++s) {
            mbar_init(sm.full(s), 1);
            mbar_init(sm.sfull(s), DROP ? 3 : 1);
            mbar_init(sm.empty(s), FB_NCW + 1);
        }
        for (int w = 0; w < FB_NCW; ++w)
            for (int k = 0; k < FB_MAXSLOT; ++k) {
                mbar_init(sm.rfull(w, k), 1);
                mbar_init(sm.rempty(w, k), 1);
            }
        for (int i = 0; i < FB_MAXNI; ++i) {
            mbar_init(sm.dqfull(i), 1);
            mbar_init(sm.dqempty(i), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem_ptr())), "n"(FB_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_ptr();
    ick_pdl_wait();  // nothing above touched global memory

    // Register re-distribution per warpgroup; each role branch STARTS with its setmaxnreg so that ptxas allocates the branch against
    // the new limit: helpers 40, epilogue 56, compute 96: 4*32*40 + 4*32*56 + 16*32*96 = 61440 = the 80 x 768 registers the CTA was launched with (setmaxnreg only re-distributes the CTA's own allocation: asking for more blocks forever).
    if (warp < 4) {
        if (!(a.dbg & 1)) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0) fb_producer(sm, a, lane, &tmQ, &tmG, &tmK, p, n_items);
        else if (warp == 1) fb_issuer(sm, a, lane, tmem_base, my_items);
        else if (DROP) fb_maskgen(sm, a, warp, lane, drop, n_items);
    } else if (warp < FB_FIRST_CW) {
        if (!(a.dbg & 1)) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        fb_epilogue(sm, a, warp, lane, tmem_base, p, n_items);
    } else {
        if (!(a.dbg & 1)) asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        fb_compute<DROP>(sm, a, warp, lane, p, drop, n_items);
    }

    // ---- teardown ------------------------------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(FB_TMEM_COLS) : "memory");
    }
}

__global__ void __launch_bounds__(256) fb_rowdot_kernel(const bf16* __restrict__ O, const bf16* __restrict__ dO, float* __restrict__ Dsum, int B,
                                                        int H, int Sq, int dh, int ldo, int lddo) {
    ick_pdl_entry();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * Sq * H) return;
    const int h = (int)(idx % H);
    const long long row = idx / H;  // b * Sq + q
    const bf16* op = O + (size_t)row * ldo + h * HD;
    const bf16* gp = dO + (size_t)row * lddo + h * HD;
    float acc = 0.f;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        float x[8], y[8];
        ld8(op + 8 * v, x);
        ld8(gp + 8 * v, y);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (8 * v + i < dh) acc = fmaf(x[i], y[i], acc);
    }
    const int b = (int)(row / Sq), q = (int)(row % Sq);
    Dsum[((size_t)b * H + h) * Sq + q] = acc;
}

int fb_num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace

// Plans and launches the fused backward; returns ICK_ERR_UNSUPPORTED (without touching anything) when the shape does not fit
// its shared-memory / tensor-memory budget, so that the caller falls back to the chunked two-kernel path.
int ick_mha_bwd_fused(const void* Q, const void* K, const void* V, const void* O, const void* dO, const float* lse, float* dsum, void* dQ,
                      void* dK, void* dV, int B, int H, int Sq, int Sk, int dh, int ldq, int ldk, int ldv, int ldo, int lddo, int lddq, int lddk,
                      int lddv, int causal, DropCfg dc, cudaStream_t stream) {
    if (dh > HD || Sq > 640 || Sk > 640 || (lddq % 8) != 0) return ICK_ERR_UNSUPPORTED;
    if ((((uintptr_t)K | (uintptr_t)Q | (uintptr_t)dO | (uintptr_t)O | (uintptr_t)dQ) & 15) != 0) return ICK_ERR_UNSUPPORTED;
    FArgs a;
    {
        const char* e = getenv("ICK_FB_DEBUG");
        a.dbg = e ? atoi(e) : 0;
    }
    a.d = make_dims(B, H, Sq, Sk, dh, causal);
    a.nslabs = (Sk + 15) / 16;
    a.ntq = (Sq + TK - 1) / TK;
    a.ntk = (Sk + TK - 1) / TK;
    a.nq128 = (Sq + 127) / 128;
    a.ngroups = dc.thr != 0u ? (Sk + 31) / 32 : 0;
    a.off_do = (uint32_t)a.ntq * TILE_BYTES;
    a.off_k = 2u * a.ntq * TILE_BYTES;
    a.off_ls = a.off_k + (uint32_t)a.ntk * TILE_BYTES;
    a.off_ds = a.off_ls + (uint32_t)a.ntq * TK * 4;
    a.off_mw = a.off_ds + (uint32_t)a.ntq * TK * 4;
    a.stage_bytes = (a.off_mw + (uint32_t)a.ngroups * a.ntq * TK * 4 + 1023u) / 1024u * 1024u;
    a.ni = FB_TMEM_COLS / (a.nq128 * 32);
    if (a.ni > FB_MAXNI) a.ni = FB_MAXNI;
    if (a.ni < 2 || a.nslabs * a.nq128 > 255) return ICK_ERR_UNSUPPORTED;
    // shared memory: barriers | ring (one or two 4 KiB slots per compute warp) | operand stages.  Two stages are a must (the next
    // item loads while this one computes); a second ring slot is taken when it still fits.
    const int avail = FB_SMEM_MAX - 1024 /*alignment slack*/ - FB_BAR_BYTES;
    a.nslot = (avail - 2 * FB_NCW * FB_SLOT) / (int)a.stage_bytes >= 2 && !(a.dbg & 16) ? 2 : 1;
    int ns = (avail - a.nslot * FB_NCW * FB_SLOT) / (int)a.stage_bytes;
    if (ns < 2) return ICK_ERR_UNSUPPORTED;
    a.nstage = ns > FB_MAXSTAGE ? FB_MAXSTAGE : ns;
    // A warp can run at most nstage items ahead of the slowest one (it needs a free operand stage), and two items that share a
    // TMEM slot (li and li + ni) must never be in flight together.  Hence nstage <= ni.
    if (a.nstage > a.ni) a.nstage = a.ni;
    int rc;
    CUtensorMap tmQ, tmG, tmK;
    if ((rc = make_tmap3(&tmQ, Q, H, Sq, B, ldq))) return rc;
    if ((rc = make_tmap3(&tmG, dO, H, Sq, B, lddo))) return rc;
    if ((rc = make_tmap3(&tmK, K, H, Sk, B, ldk))) return rc;
    {
        const long long n = (long long)B * Sq * H;
        ick_launch(fb_rowdot_kernel, (int)((n + 255) / 256), 256, 0, stream)((const bf16*)O, (const bf16*)dO, dsum, B, H, Sq, dh, ldo, lddo);
        if ((rc = ick_check_launch("mha_bwd_fused(rowdot)"))) return rc;
    }
    const int smem = 1024 + FB_BAR_BYTES + FB_NCW * a.nslot * FB_SLOT + a.nstage * (int)a.stage_bytes;
    const int grid = B * H < fb_num_sms() ? B * H : fb_num_sms();
    static bool attr_done[2] = {false, false};
    const int v = dc.thr != 0u ? 1 : 0;
    if (!attr_done[v]) {
        cudaError_t e = v ? cudaFuncSetAttribute(bwd_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_MAX)
                          : cudaFuncSetAttribute(bwd_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_MAX);
        if (e != cudaSuccess) {
            ick_set_error("mha_bwd_fused: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
            return ICK_ERR_CUDA;
        }
        attr_done[v] = true;
    }
    FPtrs p;
    p.V = (const bf16*)V; p.LSE = lse; p.Dsum = dsum;
    p.dQ = (bf16*)dQ; p.dK = (bf16*)dK; p.dV = (bf16*)dV;
    p.ldv = ldv; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
    if (v) ick_launch(bwd_fused_kernel<true>, grid, FB_THREADS, smem, stream)(tmQ, tmG, tmK, p, a, dc);
    else ick_launch(bwd_fused_kernel<false>, grid, FB_THREADS, smem, stream)(tmQ, tmG, tmK, p, a, dc);
    return ick_check_launch("mha_bwd_fused");
}
