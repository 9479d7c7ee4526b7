// bf16 attention backward for head_dim <= 32 entirely on the 5th-generation tensor cores (tcgen05 + TMEM), FA4-style.
//
// One persistent CTA per SM walks its (image, head) items.  An item is cut into tiles of 128 keys x 64 queries:
//
//   issuer thread   S^T  = K_blk Q_t^T   and   dP^T = V_blk dO_t^T        tcgen05.mma  M=128 (keys)  N=64 (queries)  K=32 (head dim)
//                   operands straight from the TMA-staged, 64B-swizzled K / V / Q / dO tiles (K-major), accumulators in TMEM
//   softmax warps   two warpgroups alternate tiles; thread r of a warpgroup owns KEY row r of the tile: tcgen05.ld brings 16
//                   queries of its S^T / dP^T rows into registers, then  P = ex2(S c - lse_q),  keep bit of (query, key) from
//                   the item's keep words,  dS = P (keep ? dP / (1-p) - D_q : -D_q),  and the bf16 rows of P^T and dS^T go to
//                   two shared-memory tiles [128 keys][64 queries] in the canonical K-major SWIZZLE_128B layout (one 16-byte
//                   store per 8 queries).  The per-query scalars are the same for all threads of a warp: broadcast loads.
//   issuer thread   dV_blk += P^T dO_t     dK_blk += dS^T Q_t       M=128 (keys)     N=32  K=64 (queries)   A = the tiles (K-major),
//                   dQ_grp += dS K_blk                              M=128 (queries)  N=32  K=128 (keys)     A = the SAME dS^T tile read
//                   MN-major (its 64 queries are one atom of M, the other atom is a box of zeros), B = Q / dO / K tiles MN-major.
//                   All three accumulate in TMEM: dK / dV per key block (double-buffered), dQ for the whole item.
//   epilogue warps  tcgen05.ld of a finished dK / dV block (scale, 1/keep) and of the item's dQ -> bf16 rows in global memory.
//
// Nothing is recomputed and nothing but Q, K, V, dO, LSE, D is read from HBM and dQ, dK, dV written: the algorithmic traffic.
// Operands stream through two rings (Q/dO + per-query scalars per item; K/V + keep words per key block) filled by TMA two
// steps ahead, so that tensor core, softmax warps, epilogue and loads of consecutive tiles / key blocks / items overlap.
//
// Warp roles (16 warps): 0 TMA producer + scalars, 1 tcgen05 issuer (+ TMEM allocation), 2-3 keep-word generators, 4-7 epilogue,
// 8-11 softmax warpgroup 0, 12-15 softmax warpgroup 1.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "attention_internal.h"
#include "attention_mma.cuh"

#ifdef ICK_TB_TRACE
// Debug build only (-DICK_TB_TRACE): per-role timeline of CTA 0 of the last launch, read back by tools/attn_trace.py.
__device__ unsigned long long ick_tb_trace_buf[6 * 4096];
#define TBT_DECL(role) unsigned int tr_i_ = 0; const unsigned int tr_role_ = (role)
#define TBT(ev, tile)                                                                                      \
    do {                                                                                                   \
        if (blockIdx.x == 0 && tr_i_ < 2047) {                                                             \
            ick_tb_trace_buf[tr_role_ * 4096 + 2 * tr_i_] = ((unsigned long long)(ev) << 32) | (unsigned)(tile); \
            ick_tb_trace_buf[tr_role_ * 4096 + 2 * tr_i_ + 1] = clock64();                                 \
            ++tr_i_;                                                                                       \
            ick_tb_trace_buf[tr_role_ * 4096 + 2 * tr_i_] = 0xFFFFFFFFFFFFFFFFull;                         \
        }                                                                                                  \
    } while (0)
extern "C" int ick_debug_tb_trace_read(unsigned long long* out) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, ick_tb_trace_buf, sizeof(unsigned long long) * 6 * 4096);
    return 6 * 2048;
}
#else
#define TBT_DECL(role) do {} while (0)
#define TBT(ev, tile) do {} while (0)
#endif

namespace {
using namespace ickattn;

constexpr int TB_THREADS = 512;
constexpr int TB_KB = 128;                 // keys per block (UMMA M)
constexpr int TB_QT = 64;                  // queries per tile (UMMA N of S^T; K of dV / dK)
constexpr int TB_TILE = TB_KB * TB_QT * 2; // one P^T or dS^T tile: 16 KiB
constexpr int TB_ZBOX = 2048;              // box of zeros: 16 rows x 128 bytes
constexpr int TB_MAXKV = 4;                // K/V ring slots
constexpr int TB_BAR_BYTES = 1024;
constexpr int TB_SMEM_MAX = 232448;
// TMEM columns
constexpr uint32_t TC_S0 = 0, TC_DP0 = 64, TC_SETSTRIDE = 128;  // S^T / dP^T of set w at TC_S0 + w*128, TC_DP0 + w*128
constexpr uint32_t TC_DK0 = 256, TC_DV0 = 288, TC_DKVSTRIDE = 64;  // dK / dV buffer b at + b*64
constexpr uint32_t TC_DQ = 384;  // dQ of 128-query group g at TC_DQ + 32*g (g < 4)

struct TArgs {
    Dims d;
    int nkb;            // 128-key blocks of an item
    int ntq;            // 64-query tiles
    int nq128;          // 128-query dQ groups
    int nkv;            // K/V ring slots
    int drop;           // dropout on
    uint32_t qg_bytes, kv_bytes;          // ring slot sizes (multiples of 1024)
    uint32_t off_do, off_ls, off_ds;      // inside a Q/dO slot: Q tiles | dO tiles | lse | D
    uint32_t off_v, off_mw;               // inside a K/V slot: K block | V block | keep words [4][ntq*64]
    int dbg;
};
struct TPtrs {
    const float* LSE;
    const float* Dsum;
    bf16 *dQ, *dK, *dV;
    int lddq, lddk, lddv;
};

// ---- tcgen05 wrappers ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// descriptors as (lo, hi) words: the hi word (SBO, version, swizzle) is a constant per operand kind, the lo word (start, LBO) base + k * constant
__device__ __forceinline__ void tc_mma2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
constexpr uint32_t DH_SW64 = (512u >> 4) | (1u << 14) | (4u << 29);    // SBO 512, version 1, SWIZZLE_64B
constexpr uint32_t DH_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t dlo(uint32_t addr, uint32_t lbo) { return ((addr >> 4) & 0x3FFFu) | ((lbo >> 4) << 16); }

__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version 1 [46,48), layout [61,64) (2: 128B, 4: 64B swizzle)
__device__ __forceinline__ uint64_t sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
// kind::f16 instruction descriptor: D f32, A / B bf16, M = 128
__host__ __device__ constexpr uint32_t idesc(int n, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- shared memory: [barriers 1 KiB][zero box lo 2 KiB][P^T, dS^T tiles of warpgroup 0, 1: 64 KiB][zero box hi 2 KiB][Q/dO ring x2][K/V ring] ----
struct TSm {
    uint32_t base;
    uint8_t* gen;
    uint32_t qg_bytes, kv_bytes;
    // barriers (8 bytes each)
    __device__ __forceinline__ uint32_t bar(int i) const { return base + 8u * (uint32_t)i; }
    __device__ __forceinline__ uint32_t qg_full(int s) const { return bar(s); }            // 2
    __device__ __forceinline__ uint32_t qg_sfull(int s) const { return bar(2 + s); }       // 2
    __device__ __forceinline__ uint32_t qg_empty(int s) const { return bar(4 + s); }       // 2
    __device__ __forceinline__ uint32_t kv_full(int s) const { return bar(6 + s); }        // 4
    __device__ __forceinline__ uint32_t kv_sfull(int s) const { return bar(10 + s); }      // 4
    __device__ __forceinline__ uint32_t kv_empty(int s) const { return bar(14 + s); }      // 4
    __device__ __forceinline__ uint32_t sp_full(int w) const { return bar(18 + w); }       // 2
    __device__ __forceinline__ uint32_t sp_empty(int w) const { return bar(20 + w); }      // 2
    __device__ __forceinline__ uint32_t ps_full(int w) const { return bar(22 + w); }       // 2
    __device__ __forceinline__ uint32_t ps_empty(int w) const { return bar(24 + w); }      // 2
    __device__ __forceinline__ uint32_t dkv_full(int b) const { return bar(26 + b); }      // 2
    __device__ __forceinline__ uint32_t dkv_empty(int b) const { return bar(28 + b); }     // 2
    __device__ __forceinline__ uint32_t dq_full() const { return bar(30); }
    __device__ __forceinline__ uint32_t dq_empty() const { return bar(31); }
    __device__ __forceinline__ uint32_t* tmem_ptr() const { return reinterpret_cast<uint32_t*>(gen + 512); }
    __device__ __forceinline__ uint32_t zero_lo() const { return base + TB_BAR_BYTES; }
    __device__ __forceinline__ uint32_t ptile(int w) const { return base + TB_BAR_BYTES + TB_ZBOX + (uint32_t)w * 2 * TB_TILE; }
    __device__ __forceinline__ uint32_t dstile(int w) const { return ptile(w) + TB_TILE; }
    __device__ __forceinline__ uint32_t zero_hi() const { return base + TB_BAR_BYTES + TB_ZBOX + 4 * TB_TILE; }
    __device__ __forceinline__ uint32_t rings() const { return base + TB_BAR_BYTES + 2 * TB_ZBOX + 4 * TB_TILE; }
    __device__ __forceinline__ uint32_t qg(int s) const { return rings() + (uint32_t)s * qg_bytes; }
    __device__ __forceinline__ uint32_t kv(int s) const { return rings() + 2 * qg_bytes + (uint32_t)s * kv_bytes; }
    __device__ __forceinline__ uint8_t* gen_of(uint32_t saddr) const { return gen + (saddr - base); }
};
constexpr int TB_FIXED = TB_BAR_BYTES + 2 * TB_ZBOX + 4 * TB_TILE;  // 70656

// causal: tile (key block kb, query tile qt) contributes nothing when every query precedes every key
__device__ __forceinline__ bool tile_dead(const Dims& d, int kb, int qt) { return d.causal && qt * TB_QT + TB_QT - 1 < kb * TB_KB; }

// ---- keep words of a key block, Mw[32-key group g < 4][query]: warp 3 and (after its copies are issued) warp 0, half of the queries each ----
__device__ __forceinline__ void tb_fill_keepwords(const TSm& sm, const TArgs& a, int half, int lane, const DropCfg& drop, uint32_t t16, int s, int b, int h,
                                                  int kb) {
    const Dims& d = a.d;
    const int nq = a.ntq * TK;
    uint32_t* Mw = reinterpret_cast<uint32_t*>(sm.gen_of(sm.kv(s) + a.off_mw));
    for (int q = half * 32 + lane; q < nq; q += 64) {
        const uint32_t rm = ick_rowmix(drop.seed, drop.site, prob_row(b, d.H, h, d.Sq, q));
#pragma unroll
        for (int g = 0; g < 4; ++g) Mw[g * nq + q] = q < d.Sq ? ick_keepword(rm, (uint32_t)(4 * kb + g), t16) : 0u;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(sm.kv_sfull(s));
}
// ---- warp 0: TMA producer + per-query scalars ------------------------------------------------------------------------------------------
template <bool DROP>
__device__ __forceinline__ void tb_producer(const TSm& sm, const TArgs& a, int lane, const CUtensorMap* tmQ, const CUtensorMap* tmG,
                                            const CUtensorMap* tmK, const CUtensorMap* tmV, const TPtrs& p, DropCfg drop, int n_items) {
    const Dims& d = a.d;
    if (DROP) ick_resolve_seed(drop);
    const uint32_t t16 = ick_attn_t16(drop.thr);
    uint32_t nqg = 0, nkvu = 0;  // ring use counters
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int b = item / d.H, h = item % d.H;
        {
            const int s = (int)(nqg & 1u);
            mbar_wait(sm.qg_empty(s), ((nqg >> 1) & 1u) ^ 1u);
            const uint32_t st = sm.qg(s);
            if (lane == 0) {
                mbar_expect_tx(sm.qg_full(s), (uint32_t)(2 * a.ntq) * TILE_BYTES);
                for (int t = 0; t < a.ntq; ++t) {
                    tma_load_3d(st + t * TILE_BYTES, tmQ, sm.qg_full(s), h * HD, t * TK, b);
                    tma_load_3d(st + a.off_do + t * TILE_BYTES, tmG, sm.qg_full(s), h * HD, t * TK, b);
                }
            }
            float* Ls = reinterpret_cast<float*>(sm.gen_of(st + a.off_ls));
            float* Ds = reinterpret_cast<float*>(sm.gen_of(st + a.off_ds));
            const float* L = p.LSE + ((size_t)b * d.H + h) * d.Sq;
            const float* Dg = p.Dsum + ((size_t)b * d.H + h) * d.Sq;
            for (int i = lane; i < a.ntq * TK; i += 32) {
                const bool ok = i < d.Sq;
                Ls[i] = ok ? L[i] : 0.f;
                Ds[i] = ok ? -Dg[i] : 0.f;  // stored negated: dS = P (dP' - D) = P (dP' + (-D))
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.qg_sfull(s));
            ++nqg;
        }
        for (int kb = 0; kb < a.nkb; ++kb) {
            const int s = (int)(nkvu % (uint32_t)a.nkv);
            mbar_wait(sm.kv_empty(s), ((nkvu / (uint32_t)a.nkv) & 1u) ^ 1u);
            if (lane == 0) {
                const uint32_t st = sm.kv(s);
                mbar_expect_tx(sm.kv_full(s), 4u * TILE_BYTES);
                for (int t = 0; t < 2; ++t) {
                    tma_load_3d(st + t * TILE_BYTES, tmK, sm.kv_full(s), h * HD, (2 * kb + t) * TK, b);
                    tma_load_3d(st + a.off_v + t * TILE_BYTES, tmV, sm.kv_full(s), h * HD, (2 * kb + t) * TK, b);
                }
            }
            if (DROP) tb_fill_keepwords(sm, a, 0, lane, drop, t16, s, b, h, kb);
            ++nkvu;
        }
    }
}

__device__ __forceinline__ void tb_maskgen(const TSm& sm, const TArgs& a, int lane, DropCfg drop, int n_items) {
    const Dims& d = a.d;
    ick_resolve_seed(drop);
    const uint32_t t16 = ick_attn_t16(drop.thr);
    uint32_t nkvu = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int b = item / d.H, h = item % d.H;
        for (int kb = 0; kb < a.nkb; ++kb) {
            const int s = (int)(nkvu % (uint32_t)a.nkv);
            mbar_wait(sm.kv_empty(s), ((nkvu / (uint32_t)a.nkv) & 1u) ^ 1u);
            tb_fill_keepwords(sm, a, 1, lane, drop, t16, s, b, h, kb);
            ++nkvu;
        }
    }
}

// ---- warps 1 and 2: the two tcgen05 issuer threads --------------------------------------------------------------------------------------
// Tile sequence of the CTA (items x key blocks x live query tiles), generated incrementally
struct TileIter {
    int item, kb, qt, qt_first;
    uint32_t nqg, nkvu;
    bool done;
};
__device__ __forceinline__ void ti_settle(TileIter& it, const TArgs& a, int n_items) {
    // skip dead tiles / advance to the next key block / item until (item, kb, qt) is a live tile or the sequence ends
    while (!it.done) {
        if (it.qt < a.ntq) {
            if (!tile_dead(a.d, it.kb, it.qt)) return;
            ++it.qt;
            it.qt_first = it.qt;
            continue;
        }
        it.qt = 0;
        it.qt_first = 0;
        ++it.nkvu;
        if (++it.kb == a.nkb) {
            it.kb = 0;
            ++it.nqg;
            it.item += gridDim.x;
            if (it.item >= n_items) it.done = true;
        }
    }
}
__device__ __forceinline__ TileIter ti_begin(const TArgs& a, int n_items) {
    TileIter it;
    it.item = blockIdx.x; it.kb = 0; it.qt = 0; it.qt_first = 0; it.nqg = 0; it.nkvu = 0; it.done = it.item >= n_items;
    ti_settle(it, a, n_items);
    return it;
}

// warp 1: S^T = K_blk Q_t^T and dP^T = V_blk dO_t^T of every tile, as soon as the tile's warpgroup has read its previous tile out of
// TMEM.  A = K / V block (128 rows x 64 B, K-major SWIZZLE_64B), B = Q / dO tile (64 rows), K = 32 = two steps of 32 bytes.
__device__ __forceinline__ void tb_issuer_sp(const TSm& sm, const TArgs& a, uint32_t tmem_base, int n_items) {
    TileIter it = ti_begin(a, n_items);
    TBT_DECL(0);
    constexpr uint32_t ID = idesc(TB_QT, 0, 0);
    uint32_t n = 0, seen_qg = 0xFFFFFFFFu, seen_kv = 0xFFFFFFFFu;
    uint32_t klo = 0, vlo = 0, qbase = 0;
    while (!it.done) {
        const uint32_t qs = it.nqg & 1u, ks = it.nkvu % (uint32_t)a.nkv;
        if (seen_qg != it.nqg) {
            mbar_wait(sm.qg_full((int)qs), (it.nqg >> 1) & 1u);
            seen_qg = it.nqg;
            qbase = sm.qg((int)qs);
        }
        if (seen_kv != it.nkvu) {
            mbar_wait(sm.kv_full((int)ks), (it.nkvu / (uint32_t)a.nkv) & 1u);
            seen_kv = it.nkvu;
            klo = dlo(sm.kv((int)ks), 16u);
            vlo = dlo(sm.kv((int)ks) + a.off_v, 16u);
        }
        const int w = (int)(n & 1u);
        TBT(1, n);
        mbar_wait(sm.sp_empty(w), ((n >> 1) & 1u) ^ 1u);
        TBT(2, n);
        tc_fence_after();
        const uint32_t qlo = dlo(qbase + (uint32_t)it.qt * TILE_BYTES, 16u), glo = dlo(qbase + a.off_do + (uint32_t)it.qt * TILE_BYTES, 16u);
        const uint32_t ts = tmem_base + TC_S0 + (uint32_t)w * TC_SETSTRIDE, tp = tmem_base + TC_DP0 + (uint32_t)w * TC_SETSTRIDE;
        tc_mma2(ts, klo, DH_SW64, qlo, DH_SW64, ID, 0u);
        tc_mma2(tp, vlo, DH_SW64, glo, DH_SW64, ID, 0u);
        tc_mma2(ts, klo + 2u, DH_SW64, qlo + 2u, DH_SW64, ID, 1u);  // + 32 bytes along K
        tc_mma2(tp, vlo + 2u, DH_SW64, glo + 2u, DH_SW64, ID, 1u);
        tc_commit(sm.sp_full(w));
        TBT(3, n);
        const bool last_of_kb = it.qt == a.ntq - 1, last_of_item = last_of_kb && it.kb == a.nkb - 1;
        if (last_of_kb) tc_commit(sm.kv_empty((int)ks));     // this thread's reads of the K / V block are done when these MMAs are
        if (last_of_item) tc_commit(sm.qg_empty((int)qs));
        ++n;
        ++it.qt;
        ti_settle(it, a, n_items);
    }
}

// warp 2: the gradient MMAs of every tile once its P^T / dS^T rows are written.
//   dV += P^T dO_t, dK += dS^T Q_t : A = tile rows (keys) x 64 queries, K-major SWIZZLE_128B; B = dO / Q tile, MN-major SWIZZLE_64B (16 rows per step)
//   dQ_grp += dS K_blk             : A = the dS^T tile read MN-major (its 64 queries are one atom of M = 128, the other atom a box of zeros
//                                    LBO bytes away), B = K block MN-major; K = 128 keys = 8 steps of 16 rows
__device__ __forceinline__ void tb_issuer_grad(const TSm& sm, const TArgs& a, uint32_t tmem_base, int n_items) {
    TileIter it = ti_begin(a, n_items);
    TBT_DECL(1);
    constexpr uint32_t ID_KV = idesc(32, 0, 1), ID_Q = idesc(32, 1, 1);
    uint32_t n = 0, dkv_use = 0, dq_use = 0, seen_qg = 0xFFFFFFFFu, seen_kv = 0xFFFFFFFFu;
    uint32_t klo = 0, qbase = 0;
    const uint32_t zl4 = sm.zero_lo() >> 4, zh4 = sm.zero_hi() >> 4;
    while (!it.done) {
        const uint32_t qs = it.nqg & 1u, ks = it.nkvu % (uint32_t)a.nkv;
        if (seen_qg != it.nqg) {
            mbar_wait(sm.qg_full((int)qs), (it.nqg >> 1) & 1u);
            seen_qg = it.nqg;
            qbase = sm.qg((int)qs);
        }
        if (seen_kv != it.nkvu) {
            mbar_wait(sm.kv_full((int)ks), (it.nkvu / (uint32_t)a.nkv) & 1u);
            seen_kv = it.nkvu;
            klo = dlo(sm.kv((int)ks), 512u);
        }
        const int w = (int)(n & 1u);
        const bool first_of_kb = it.qt == it.qt_first, last_of_kb = it.qt == a.ntq - 1;
        const bool first_dq = it.kb == 0 && (it.qt & 1) == 0, last_of_item = last_of_kb && it.kb == a.nkb - 1;
        const uint32_t buf = dkv_use & 1u;
        TBT(4, n);
        mbar_wait(sm.ps_full(w), (n >> 1) & 1u);
        TBT(5, n);
        if (first_of_kb) mbar_wait(sm.dkv_empty((int)buf), ((dkv_use >> 1) & 1u) ^ 1u);
        if (first_dq) mbar_wait(sm.dq_empty(), (dq_use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t plo = dlo(sm.ptile(w), 16u), slo = dlo(sm.dstile(w), 16u);
        const uint32_t qlo = dlo(qbase + (uint32_t)it.qt * TILE_BYTES, 512u), glo = dlo(qbase + a.off_do + (uint32_t)it.qt * TILE_BYTES, 512u);
        const uint32_t tdk = tmem_base + TC_DK0 + buf * TC_DKVSTRIDE, tdv = tmem_base + TC_DV0 + buf * TC_DKVSTRIDE;
        if (!(a.dbg & 2)) {
            const uint32_t acc0 = first_of_kb ? 0u : 1u;
#pragma unroll
            for (int k = 0; k < TB_QT / 16; ++k) {  // A: + 32 bytes along K; B: + 16 rows x 64 bytes
                tc_mma2(tdv, plo + 2u * k, DH_SW128, glo + 64u * k, DH_SW64, ID_KV, k == 0 ? acc0 : 1u);
                tc_mma2(tdk, slo + 2u * k, DH_SW128, qlo + 64u * k, DH_SW64, ID_KV, k == 0 ? acc0 : 1u);
            }
        }
        if (!(a.dbg & 4)) {
            const uint32_t tdq = tmem_base + TC_DQ + 32u * (uint32_t)(it.qt >> 1);
            const uint32_t d4 = sm.dstile(w) >> 4;
            // even tile: (rows | zeros): start = rows, LBO = zero_hi - rows;  odd tile: (zeros | rows): start = zero_lo, LBO = rows - zero_lo;
            // rows advance by 16 x 128 bytes = 128 descriptor units per step
            uint32_t alo = (it.qt & 1) == 0 ? (d4 | ((zh4 - d4) << 16)) : (zl4 | ((d4 - zl4) << 16));
            const uint32_t astep = (it.qt & 1) == 0 ? (128u - (128u << 16)) : (128u << 16);
            const uint32_t acc0 = first_dq ? 0u : 1u;
#pragma unroll
            for (int k = 0; k < TB_KB / 16; ++k) tc_mma2(tdq, alo + astep * (uint32_t)k, DH_SW128, klo + 64u * k, DH_SW64, ID_Q, k == 0 ? acc0 : 1u);
        }
        tc_commit(sm.ps_empty(w));
        TBT(6, n);
        if (last_of_kb) {
            tc_commit(sm.dkv_full((int)buf));
            tc_commit(sm.kv_empty((int)ks));
            ++dkv_use;
        }
        if (last_of_item) {
            tc_commit(sm.dq_full());
            tc_commit(sm.qg_empty((int)qs));
            ++dq_use;
        }
        ++n;
        ++it.qt;
        ti_settle(it, a, n_items);
    }
}

// ---- warps 4-7: epilogue ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tb_store_row(bf16* dst, const uint32_t* r, float s) {
    uint4 u0, u1;
    u0.x = pack2(__uint_as_float(r[0]) * s, __uint_as_float(r[1]) * s);
    u0.y = pack2(__uint_as_float(r[2]) * s, __uint_as_float(r[3]) * s);
    u0.z = pack2(__uint_as_float(r[4]) * s, __uint_as_float(r[5]) * s);
    u0.w = pack2(__uint_as_float(r[6]) * s, __uint_as_float(r[7]) * s);
    u1.x = pack2(__uint_as_float(r[8]) * s, __uint_as_float(r[9]) * s);
    u1.y = pack2(__uint_as_float(r[10]) * s, __uint_as_float(r[11]) * s);
    u1.z = pack2(__uint_as_float(r[12]) * s, __uint_as_float(r[13]) * s);
    u1.w = pack2(__uint_as_float(r[14]) * s, __uint_as_float(r[15]) * s);
    reinterpret_cast<uint4*>(dst)[0] = u0;
    reinterpret_cast<uint4*>(dst)[1] = u1;
}
__device__ __forceinline__ void tb_epilogue(const TSm& sm, const TArgs& a, int warp, int lane, uint32_t tmem_base, const TPtrs& p, float inv_keep,
                                            int n_items) {
    const Dims& d = a.d;
    const int qd = warp & 3;
    const uint32_t tl = tmem_base + ((uint32_t)(qd * 32) << 16);
    uint32_t dkv_use = 0, dq_use = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int b = item / d.H, h = item % d.H;
        for (int kb = 0; kb < a.nkb; ++kb) {
            const uint32_t buf = dkv_use & 1u;
            mbar_wait(sm.dkv_full((int)buf), (dkv_use >> 1) & 1u);
            tc_fence_after();
            const int key = kb * TB_KB + qd * 32 + lane;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t rk[16], rv[16];
                tc_ld16_nowait(tl + TC_DK0 + buf * TC_DKVSTRIDE + 16u * c, rk);
                tc_ld16_nowait(tl + TC_DV0 + buf * TC_DKVSTRIDE + 16u * c, rv);
                tc_wait_ld();
                if (key < d.Sk) {
                    tb_store_row(p.dK + ((size_t)b * d.Sk + key) * p.lddk + h * HD + 16 * c, rk, d.scale);
                    tb_store_row(p.dV + ((size_t)b * d.Sk + key) * p.lddv + h * HD + 16 * c, rv, inv_keep);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.dkv_empty((int)buf));
            ++dkv_use;
        }
        mbar_wait(sm.dq_full(), dq_use & 1u);
        tc_fence_after();
        for (int g = 0; g < a.nq128; ++g) {
            const int q = g * 128 + qd * 32 + lane;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t r[16];
                tc_ld16_nowait(tl + TC_DQ + 32u * g + 16u * c, r);
                tc_wait_ld();
                if (q < d.Sq) tb_store_row(p.dQ + ((size_t)b * d.Sq + q) * p.lddq + h * HD + 16 * c, r, d.scale);
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sm.dq_empty());
        ++dq_use;
    }
}

// ---- warps 8-15: the two softmax warpgroups ---------------------------------------------------------------------------------------------------
// 16 queries of the thread's key row: registers s[16] (S^T) and dp[16] (dP^T) -> packed bf16 P^T / dS^T
template <bool DROP, bool MASK>
__device__ __forceinline__ void tb_chunk(const uint32_t* s, const uint32_t* dp, uint32_t ls, uint32_t ds, uint32_t mw, uint32_t mk, float c, float ik,
                                         int key, int q0, uint32_t* pp, uint32_t* pd) {
    float pv[16], dv[16];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const float4 l4 = lds_f4(ls + 16u * v), d4 = lds_f4(ds + 16u * v);
        uint4 w4 = make_uint4(0u, 0u, 0u, 0u);
        if (DROP) w4 = lds_u4(mw + 16u * v);
        const float lq[4] = {l4.x, l4.y, l4.z, l4.w}, nd[4] = {d4.x, d4.y, d4.z, d4.w};
        const uint32_t wq[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = 4 * v + e;
            float pr = ex2(fmaf(__uint_as_float(s[i]), c, -lq[e]));
            if (MASK && key > q0 + i) pr = 0.f;
            if (DROP) {
                const bool keep = (wq[e] & mk) != 0u;
                const float t = fmaf(__uint_as_float(dp[i]), ik, nd[e]);
                dv[i] = pr * (keep ? t : nd[e]);
                pv[i] = keep ? pr : 0.f;
            } else {
                dv[i] = pr * (__uint_as_float(dp[i]) + nd[e]);
                pv[i] = pr;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        pp[i] = pack2(pv[2 * i], pv[2 * i + 1]);
        pd[i] = pack2(dv[2 * i], dv[2 * i + 1]);
    }
}
template <bool DROP>
__device__ __forceinline__ void tb_softmax(const TSm& sm, const TArgs& a, int warp, int lane, uint32_t tmem_base, DropCfg drop, int n_items) {
    const Dims& d = a.d;
    const int w = (warp - 8) >> 2;   // warpgroup = TMEM set = tile buffer
    const int qd = warp & 3;         // TMEM lane quadrant
    const int row = qd * 32 + lane;  // key row inside the 128-key block
    const float c = d.scale_log2, ik = drop.inv_keep;
    const uint32_t tl = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)w * TC_SETSTRIDE;
    const uint32_t mk = 1u << ick_keybit((uint32_t)row);  // bit of this key inside its 32-key group (group = qd)
    // row of the K-major SWIZZLE_128B tiles: 128 bytes, 16-byte chunk ch stored at chunk ch ^ (row & 7)
    const uint32_t prow = sm.ptile(w) + (uint32_t)row * 128u, drow = sm.dstile(w) + (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    const int nq = a.ntq * TK;
    uint32_t n = 0, nqg = 0, nkvu = 0;
#ifdef ICK_TB_TRACE
    unsigned int tr_i_ = 0;
    const unsigned int tr_role_ = (lane == 0 && qd == 0) ? 2 + w : 5;  // role 5: everybody else (never read)
#endif
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t qs = nqg & 1u;
        // every warp passes every ring slot in order (also one it has no tile in): nobody can run ahead of the rings' phases
        mbar_wait(sm.qg_full((int)qs), (nqg >> 1) & 1u);
        mbar_wait(sm.qg_sfull((int)qs), (nqg >> 1) & 1u);
        const uint32_t ls0 = sm.qg((int)qs) + a.off_ls, ds0 = sm.qg((int)qs) + a.off_ds;
        for (int kb = 0; kb < a.nkb; ++kb) {
            const uint32_t ks = nkvu % (uint32_t)a.nkv, kph = (nkvu / (uint32_t)a.nkv) & 1u;
            mbar_wait(sm.kv_full((int)ks), kph);
            if (DROP) mbar_wait(sm.kv_sfull((int)ks), kph);
            const uint32_t mw0 = sm.kv((int)ks) + a.off_mw + (uint32_t)(qd * nq) * 4u;
            const int key = kb * TB_KB + row;
            // a warp whose 32 keys all lie beyond Sk has nothing to compute: its rows of the tiles are zeros
            const bool warp_live = kb * TB_KB + qd * 32 < d.Sk;
            int qt_first = 0;
            while (qt_first < a.ntq && tile_dead(d, kb, qt_first)) ++qt_first;
            for (int qt = qt_first; qt < a.ntq; ++qt, ++n) {
                if ((int)(n & 1u) != w) continue;
                const uint32_t use = n >> 1;
                TBT(7, n);
                mbar_wait(sm.sp_full(w), use & 1u);
                TBT(8, n);
                tc_fence_after();
                const bool diag = d.causal && kb * TB_KB + TB_KB - 1 > qt * TB_QT;  // some (key, query) pairs of the tile are masked
#pragma unroll 1
                for (int ch = 0; ch < TB_QT / 16; ++ch) {
                    const int q0 = qt * TB_QT + 16 * ch;
                    uint32_t pp[8], pd[8];
                    if (warp_live && q0 < d.Sq && !(a.dbg & 8)) {
                        uint32_t s[16], dp[16];
                        tc_ld16_nowait(tl + TC_S0 + 16u * ch, s);
                        tc_ld16_nowait(tl + TC_DP0 + 16u * ch, dp);
                        tc_wait_ld();
                        if (ch == TB_QT / 16 - 1) {  // S^T / dP^T of this set have been read: the issuer may overwrite them (next tile)
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(sm.sp_empty(w));
                        }
                        const uint32_t qo = 4u * (uint32_t)q0;
                        if (diag) tb_chunk<DROP, true>(s, dp, ls0 + qo, ds0 + qo, mw0 + qo, mk, c, ik, key, q0, pp, pd);
                        else tb_chunk<DROP, false>(s, dp, ls0 + qo, ds0 + qo, mw0 + qo, mk, c, ik, key, q0, pp, pd);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) pp[i] = pd[i] = 0u;
                        if (ch == TB_QT / 16 - 1) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(sm.sp_empty(w));
                        }
                    }
                    if (ch == 0) {
                        TBT(9, n);
                        mbar_wait(sm.ps_empty(w), (use & 1u) ^ 1u);  // the gradient MMAs of this warpgroup's previous tile are done
                        TBT(10, n);
                    }
                    const uint32_t c0 = ((uint32_t)(2 * ch) ^ sw) << 4, c1 = ((uint32_t)(2 * ch + 1) ^ sw) << 4;
                    sts_u4(prow + c0, pp[0], pp[1], pp[2], pp[3]);
                    sts_u4(prow + c1, pp[4], pp[5], pp[6], pp[7]);
                    sts_u4(drow + c0, pd[0], pd[1], pd[2], pd[3]);
                    sts_u4(drow + c1, pd[4], pd[5], pd[6], pd[7]);
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(sm.ps_full(w));  // P^T / dS^T rows of this warp are in place
                TBT(11, n);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.kv_empty((int)ks));
            ++nkvu;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sm.qg_empty((int)qs));
        ++nqg;
    }
}

template <bool DROP>
__global__ void __launch_bounds__(TB_THREADS, 1)
    bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, TPtrs p, TArgs a, DropCfg drop) {
    ick_pdl_launch();
    extern __shared__ uint8_t smem_raw[];
    TSm sm;
    sm.gen = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    sm.base = smem_u32(sm.gen);
    sm.qg_bytes = a.qg_bytes;
    sm.kv_bytes = a.kv_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = a.d.B * a.d.H;
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
        for (int s = 0; s < 2; ++s) {
            mbar_init(sm.qg_full(s), 1);
            mbar_init(sm.qg_sfull(s), 1);
            mbar_init(sm.qg_empty(s), 2 + 8);  // the two issuer threads' commits + the 8 softmax warps
        }
        for (int s = 0; s < TB_MAXKV; ++s) {
            mbar_init(sm.kv_full(s), 1);
            mbar_init(sm.kv_sfull(s), 2);
            mbar_init(sm.kv_empty(s), 2 + 8);
        }
        for (int w = 0; w < 2; ++w) {
            mbar_init(sm.sp_full(w), 1);
            mbar_init(sm.sp_empty(w), 4);
            mbar_init(sm.ps_full(w), 4);
            mbar_init(sm.ps_empty(w), 1);
            mbar_init(sm.dkv_full(w), 1);
            mbar_init(sm.dkv_empty(w), 4);
        }
        mbar_init(sm.dq_full(), 1);
        mbar_init(sm.dq_empty(), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        uint32_t* zl = reinterpret_cast<uint32_t*>(sm.gen_of(sm.zero_lo()));
        uint32_t* zh = reinterpret_cast<uint32_t*>(sm.gen_of(sm.zero_hi()));
        for (int i = threadIdx.x; i < TB_ZBOX / 4; i += TB_THREADS) { zl[i] = 0u; zh[i] = 0u; }
        fence_async_smem();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm.tmem_ptr())), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_ptr();
    ick_pdl_wait();

    if (warp == 0) {
        tb_producer<DROP>(sm, a, lane, &tmQ, &tmG, &tmK, &tmV, p, drop, n_items);
    } else if (warp == 1) {
        if (lane == 0) tb_issuer_sp(sm, a, tmem_base, n_items);
    } else if (warp == 2) {
        if (lane == 0) tb_issuer_grad(sm, a, tmem_base, n_items);
    } else if (warp == 3) {
        if (DROP) tb_maskgen(sm, a, lane, drop, n_items);
    } else if (warp < 8) {
        tb_epilogue(sm, a, warp, lane, tmem_base, p, drop.inv_keep, n_items);
    } else {
        tb_softmax<DROP>(sm, a, warp, lane, tmem_base, drop, n_items);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

__global__ void __launch_bounds__(256) tb_rowdot_kernel(const bf16* __restrict__ O, const bf16* __restrict__ dO, float* __restrict__ Dsum, int B, int H,
                                                        int Sq, int dh, int ldo, int lddo) {
    ick_pdl_entry();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * Sq * H) return;
    const int h = (int)(idx % H);
    const long long row = idx / H;
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

int tb_num_sms() {
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

// Plans and launches the tcgen05 backward; ICK_ERR_UNSUPPORTED (nothing launched) when the shape does not fit its shared / tensor
// memory budget (more than 512 queries, or operand rings that do not fit) - the caller then takes the mma.sync hybrid.
int ick_mha_bwd_tc(const void* Q, const void* K, const void* V, const void* O, const void* dO, const float* lse, float* dsum, void* dQ, void* dK,
                   void* dV, int B, int H, int Sq, int Sk, int dh, int ldq, int ldk, int ldv, int ldo, int lddo, int lddq, int lddk, int lddv,
                   int causal, DropCfg dc, int dsum_ready, cudaStream_t stream) {
    if (dh > HD || Sq > 512 || (lddq % 8) != 0 || (lddk % 8) != 0 || (lddv % 8) != 0) return ICK_ERR_UNSUPPORTED;
    if ((((uintptr_t)K | (uintptr_t)V | (uintptr_t)Q | (uintptr_t)dO | (uintptr_t)O | (uintptr_t)dQ | (uintptr_t)dK | (uintptr_t)dV) & 15) != 0)
        return ICK_ERR_UNSUPPORTED;
    TArgs a;
    {
        const char* e = getenv("ICK_TB_DEBUG");
        a.dbg = e ? atoi(e) : 0;
    }
    a.d = make_dims(B, H, Sq, Sk, dh, causal);
    a.nkb = (Sk + TB_KB - 1) / TB_KB;
    a.ntq = (Sq + TB_QT - 1) / TB_QT;
    a.nq128 = (Sq + 127) / 128;
    a.drop = dc.thr != 0u ? 1 : 0;
    a.off_do = (uint32_t)a.ntq * TILE_BYTES;
    a.off_ls = 2u * a.ntq * TILE_BYTES;
    a.off_ds = a.off_ls + (uint32_t)a.ntq * TK * 4;
    a.qg_bytes = (a.off_ds + (uint32_t)a.ntq * TK * 4 + 1023u) / 1024u * 1024u;
    a.off_v = 2u * TILE_BYTES;
    a.off_mw = 4u * TILE_BYTES;
    a.kv_bytes = (a.off_mw + (a.drop ? 4u * a.ntq * TK * 4 : 0u) + 1023u) / 1024u * 1024u;
    const int avail = TB_SMEM_MAX - 1024 - TB_FIXED - 2 * (int)a.qg_bytes;
    a.nkv = avail / (int)a.kv_bytes;
    if (a.nkv > TB_MAXKV) a.nkv = TB_MAXKV;
    if (a.nkv < 2) return ICK_ERR_UNSUPPORTED;
    int rc;
    CUtensorMap tmQ, tmG, tmK, tmV;
    if ((rc = make_tmap3(&tmQ, Q, H, Sq, B, ldq))) return rc;
    if ((rc = make_tmap3(&tmG, dO, H, Sq, B, lddo))) return rc;
    if ((rc = make_tmap3(&tmK, K, H, Sk, B, ldk))) return rc;
    if ((rc = make_tmap3(&tmV, V, H, Sk, B, ldv))) return rc;
    if (!dsum_ready) {
        const long long n = (long long)B * Sq * H;
        ick_launch(tb_rowdot_kernel, (int)((n + 255) / 256), 256, 0, stream)((const bf16*)O, (const bf16*)dO, dsum, B, H, Sq, dh, ldo, lddo);
        if ((rc = ick_check_launch("mha_bwd_tc(rowdot)"))) return rc;
    }
    const int smem = 1024 + TB_FIXED + 2 * (int)a.qg_bytes + a.nkv * (int)a.kv_bytes;
    const int grid = B * H < tb_num_sms() ? B * H : tb_num_sms();
    static bool attr_done[2] = {false, false};
    if (!attr_done[a.drop]) {
        cudaError_t e = a.drop ? cudaFuncSetAttribute(bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM_MAX)
                               : cudaFuncSetAttribute(bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM_MAX);
        if (e != cudaSuccess) {
            ick_set_error("mha_bwd_tc: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
            return ICK_ERR_CUDA;
        }
        attr_done[a.drop] = true;
    }
    TPtrs p;
    p.LSE = lse; p.Dsum = dsum;
    p.dQ = (bf16*)dQ; p.dK = (bf16*)dK; p.dV = (bf16*)dV;
    p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
    if (a.drop) ick_launch(bwd_tc_kernel<true>, grid, TB_THREADS, smem, stream)(tmQ, tmG, tmK, tmV, p, a, dc);
    else ick_launch(bwd_tc_kernel<false>, grid, TB_THREADS, smem, stream)(tmQ, tmG, tmK, tmV, p, a, dc);
    return ick_check_launch("mha_bwd_tc");
}
