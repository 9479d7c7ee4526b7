// bf16 attention forward for head_dim <= 32 on the 5th-generation tensor cores (tcgen05 + TMEM), the forward twin of
// attention_bwd_tc.cu.
//
// One persistent CTA per SM walks its (image, head) items.  An item is cut into blocks of 128 queries, a block walks the keys in
// tiles of 64:
//
//   issuer threads  S = Q_blk K_t^T      tcgen05.mma  M=128 (queries)  N=64 (keys)  K=32 (head dim), operands = the TMA-staged,
//                   64B-swizzled Q / K tiles (K-major), accumulator in TMEM (two buffers per warpgroup)
//   softmax warps   two warpgroups work on two different query blocks; thread r owns QUERY row r: tcgen05.ld brings the 64 scores
//                   of its row into registers (the buffer is handed back at once), row maximum, P = ex2(S c - m c), row sum,
//                   attention-probability dropout from the row's keep words (bit-parallel masks of common.cuh, hashed by the thread
//                   itself: one hash per 32 keys at the reference's p = 0.5), bf16 row of P -> a shared-memory tile [128 queries]
//                   [64 keys] (K-major SWIZZLE_128B).  The running maximum only moves when it grows by more than 2^8 (the
//                   accumulator then is rescaled in TMEM by the same threads) - exact, because the final 1/l uses the same base.
//   issuer threads  O_blk += P V_t        M=128  N=32  K=64 (keys): A = the P tile, B = the V tile MN-major, accumulator in TMEM
//   softmax warps   after the last key tile: tcgen05.ld of their O rows, x 1/(l (1-p)), bf16 rows + LSE (log2 domain) to global.
//
// Warp roles (12 warps): 0 TMA producer, 1 / 2 issuer thread of warpgroup 0 / 1 (warp 1 also allocates TMEM), 3 idle,
// 4-7 softmax warpgroup 0, 8-11 softmax warpgroup 1.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "attention_internal.h"
#include "attention_mma.cuh"

namespace {
using namespace ickattn;

constexpr int TF_THREADS = 384;
constexpr int TF_QB = 128;                    // queries per block (UMMA M)
constexpr int TF_KT = 64;                     // keys per tile
constexpr int TF_PTILE = TF_QB * TF_KT * 2;   // one P tile: 16 KiB
constexpr int TF_BAR_BYTES = 1024;
constexpr int TF_SMEM_MAX = 232448;
constexpr float TF_LAZY = 8.0f;               // the running maximum moves only when it grows by more than this (log2 domain)
// TMEM columns of warpgroup w: S buffer b at w*192 + 64*b, O at w*192 + 128
constexpr uint32_t TF_WGSTRIDE = 192, TF_O = 128;

struct FwArgs {
    Dims d;
    int nqb;         // 128-query blocks of an item
    int ntq, ntk;    // 64-row tiles of Q and of K / V
    int npb;         // P tile buffers per warpgroup (1 or 2)
    uint32_t stage_bytes, off_k, off_v;
    int dbg;
};

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
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
__host__ __device__ constexpr uint32_t idesc(int n, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- shared memory: [barriers 1 KiB][P tiles: 2 warpgroups x npb x 16 KiB][operand stage 0][operand stage 1] -----------------------------
struct FwSm {
    uint32_t base;
    uint32_t stage_bytes, ptile_bytes;
    int npb;
    __device__ __forceinline__ uint32_t bar(int i) const { return base + 8u * (uint32_t)i; }
    __device__ __forceinline__ uint32_t full(int s) const { return bar(s); }                          // 2: operands of an item landed
    __device__ __forceinline__ uint32_t empty(int s) const { return bar(2 + s); }                     // 2
    __device__ __forceinline__ uint32_t s_full(int w, int b) const { return bar(4 + 2 * w + b); }      // 4: S buffer written by the MMA
    __device__ __forceinline__ uint32_t s_empty(int w, int b) const { return bar(8 + 2 * w + b); }     // 4: ... read by the warpgroup
    __device__ __forceinline__ uint32_t p_full(int w, int b) const { return bar(12 + 2 * w + b); }     // 4: P tile written
    __device__ __forceinline__ uint32_t p_empty(int w, int b) const { return bar(16 + 2 * w + b); }    // 4: ... consumed by the MMA
    __device__ __forceinline__ uint32_t o_full(int w) const { return bar(20 + w); }                   // 2: O of a query block complete
    __device__ __forceinline__ uint32_t o_empty(int w) const { return bar(22 + w); }                  // 2: ... read out
    __device__ __forceinline__ uint32_t ptile(int w, int b) const { return base + TF_BAR_BYTES + (uint32_t)(w * npb + b) * TF_PTILE; }
    __device__ __forceinline__ uint32_t stage(int s) const { return base + TF_BAR_BYTES + ptile_bytes + (uint32_t)s * stage_bytes; }
};

// Work of the CTA as a sequence of query blocks (item-major); warpgroup = block number & 1.
struct BlkIter {
    int item, qb;
    uint32_t li;  // local item counter (operand stage = li & 1)
    bool done;
};
__device__ __forceinline__ void bi_next(BlkIter& it, const FwArgs& a, int n_items) {
    if (++it.qb == a.nqb) {
        it.qb = 0;
        ++it.li;
        it.item += gridDim.x;
        if (it.item >= n_items) it.done = true;
    }
}
// key tiles a query block has to visit (causal: up to the tile of its last query)
__device__ __forceinline__ int ntiles_of(const FwArgs& a, int qb) {
    if (!a.d.causal) return a.ntk;
    const int last_q = min(a.d.Sq, qb * TF_QB + TF_QB) - 1;
    return min(a.ntk, last_q / TF_KT + 1);
}

// ---- warp 0: TMA producer -------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tf_producer(const FwSm& sm, const FwArgs& a, const CUtensorMap* tmQ, const CUtensorMap* tmK, const CUtensorMap* tmV,
                                            int n_items) {
    const Dims& d = a.d;
    uint32_t li = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
        const int s = (int)(li & 1u), b = item / d.H, h = item % d.H;
        mbar_wait(sm.empty(s), ((li >> 1) & 1u) ^ 1u);
        const uint32_t st = sm.stage(s), bar = sm.full(s);
        mbar_expect_tx(bar, (uint32_t)(a.ntq + 2 * a.ntk) * TILE_BYTES);
        for (int t = 0; t < a.ntq; ++t) tma_load_3d(st + t * TILE_BYTES, tmQ, bar, h * HD, t * TK, b);
        for (int t = 0; t < a.ntk; ++t) {
            tma_load_3d(st + a.off_k + t * TILE_BYTES, tmK, bar, h * HD, t * TK, b);
            tma_load_3d(st + a.off_v + t * TILE_BYTES, tmV, bar, h * HD, t * TK, b);
        }
    }
}

// ---- warps 1 / 2: the issuer thread of warpgroup w ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tf_issuer(const FwSm& sm, const FwArgs& a, int w, uint32_t tmem_base, int n_items) {
    constexpr uint32_t ID_S = idesc(TF_KT, 0, 0), ID_O = idesc(32, 0, 1);
    const uint32_t tw = tmem_base + (uint32_t)w * TF_WGSTRIDE;
    BlkIter it;
    it.item = blockIdx.x; it.qb = 0; it.li = 0; it.done = it.item >= n_items;
    uint32_t n = 0;                 // query blocks seen (all warpgroups)
    uint32_t ns = 0, np = 0, nblk = 0;  // S tiles issued, P tiles consumed, blocks of this warpgroup
    uint32_t seen_li = 0xFFFFFFFFu;
    for (; !it.done; bi_next(it, a, n_items), ++n) {
        const int s = (int)(it.li & 1u);
        const bool last_blk_of_item = it.qb == a.nqb - 1;
        if ((int)(n & 1u) == w) {
            if (seen_li != it.li) {
                mbar_wait(sm.full(s), (it.li >> 1) & 1u);
                seen_li = it.li;
            }
            const uint32_t st = sm.stage(s);
            // Q block: 128 rows x 64 B K-major SWIZZLE_64B (two consecutive 64-row tiles); K tile 64 rows
            const uint32_t qlo = dlo(st + (uint32_t)it.qb * 2u * TILE_BYTES, 16u);
            const int nt = ntiles_of(a, it.qb);
            mbar_wait(sm.o_empty(w), (nblk & 1u) ^ 1u);  // the warpgroup has read the previous block's O out of TMEM
            // software pipeline: S(t+1) is issued before the P V product of tile t, so the warpgroup always finds its next scores ready
            auto issue_s = [&](int t) {
                const uint32_t b = ns & 1u;
                mbar_wait(sm.s_empty(w, (int)b), ((ns >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t klo = dlo(st + a.off_k + (uint32_t)t * TILE_BYTES, 16u);
                tc_mma2(tw + 64u * b, qlo, DH_SW64, klo, DH_SW64, ID_S, 0u);
                tc_mma2(tw + 64u * b, qlo + 2u, DH_SW64, klo + 2u, DH_SW64, ID_S, 1u);  // + 32 bytes along the head dimension
                tc_commit(sm.s_full(w, (int)b));
                ++ns;
            };
            issue_s(0);
            for (int t = 0; t < nt; ++t) {
                if (t + 1 < nt) issue_s(t + 1);
                const uint32_t pb = a.npb == 2 ? (np & 1u) : 0u, use = a.npb == 2 ? (np >> 1) : np;
                mbar_wait(sm.p_full(w, (int)pb), use & 1u);
                tc_fence_after();
                // O += P V_t : A = P tile (128 queries x 64 keys, K-major SWIZZLE_128B), B = V tile (64 keys x 32) MN-major SWIZZLE_64B
                const uint32_t plo = dlo(sm.ptile(w, (int)pb), 16u), vlo = dlo(st + a.off_v + (uint32_t)t * TILE_BYTES, 512u);
#pragma unroll
                for (int k = 0; k < TF_KT / 16; ++k) tc_mma2(tw + TF_O, plo + 2u * k, DH_SW128, vlo + 64u * k, DH_SW64, ID_O, (t | k) != 0 ? 1u : 0u);
                tc_commit(sm.p_empty(w, (int)pb));
                ++np;
            }
            tc_commit(sm.o_full(w));
            ++nblk;
        }
        // the operand stage is free when the MMAs of BOTH warpgroups' blocks of the item are done: each issuer commits once per item,
        // behind its last block of the item (a warpgroup without a block in the item commits at once)
        if (last_blk_of_item) tc_commit(sm.empty(s));
    }
}

// ---- warps 4-11: the two softmax warpgroups ------------------------------------------------------------------------------------------------------
template <bool DROP>
__device__ __forceinline__ void tf_softmax(const FwSm& sm, const FwArgs& a, int warp, int lane, uint32_t tmem_base, bf16* __restrict__ O,
                                           float* __restrict__ LSE, int ldo, DropCfg drop, int n_items) {
    const Dims& d = a.d;
    const int w = (warp - 4) >> 2, qd = warp & 3, row = qd * 32 + lane;
    const float c = d.scale_log2;
    const uint32_t tl = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)w * TF_WGSTRIDE;
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t t16 = ick_attn_t16(drop.thr);
    if (DROP) ick_resolve_seed(drop);
    BlkIter it;
    it.item = blockIdx.x; it.qb = 0; it.li = 0; it.done = it.item >= n_items;
    uint32_t n = 0, ns = 0, np = 0, nblk = 0;
    for (; !it.done; bi_next(it, a, n_items), ++n) {
        if ((int)(n & 1u) != w) continue;
        const int b = it.item / d.H, h = it.item % d.H;
        const int q = it.qb * TF_QB + row;
        const bool q_ok = q < d.Sq;
        const int nt = ntiles_of(a, it.qb);
        uint32_t rm = 0u;
        if (DROP) rm = ick_rowmix(drop.seed, drop.site, prob_row(b, d.H, h, d.Sq, q));
        float m_used = -INFINITY, l = 0.f;  // exponent base of the accumulated O / l (raw-score units), running row sum
        for (int t = 0; t < nt; ++t) {
            const uint32_t sb = ns & 1u;
            mbar_wait(sm.s_full(w, (int)sb), (ns >> 1) & 1u);
            tc_fence_after();
            uint32_t sr[64];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) tc_ld16_nowait(tl + 64u * sb + 16u * ch, sr + 16 * ch);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.s_empty(w, (int)sb));  // the scores are in registers: the buffer may take the tile after next
            ++ns;
            const int k0 = t * TF_KT;
            // masking: keys beyond Sk (last tile) and, causal, keys behind the query (tiles that cross the block's diagonal)
            const bool need_mask = k0 + TF_KT > d.Sk || (d.causal && k0 + TF_KT - 1 > it.qb * TF_QB);
            float s[64];
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                s[i] = __uint_as_float(sr[i]);
                if (need_mask && (k0 + i >= d.Sk || (d.causal && k0 + i > q))) s[i] = -INFINITY;
                mx = fmaxf(mx, s[i]);
            }
            // lazy running maximum: keep the old base unless the row maximum outgrew it by more than 2^TF_LAZY
            float alpha = 1.f;
            bool moved = false;
            if (mx > m_used + TF_LAZY / c || m_used == -INFINITY) {
                if (mx != -INFINITY) {
                    alpha = m_used == -INFINITY ? 0.f : ex2((m_used - mx) * c);
                    moved = m_used != -INFINITY;
                    m_used = mx;
                }
            }
            const uint32_t pb = a.npb == 2 ? (np & 1u) : 0u, use = a.npb == 2 ? (np >> 1) : np;
            // Rescaling O needs the previous P V product finished; so does re-using a single P buffer.  (With two P buffers the wait
            // below is for the product of the tile before the previous one.)
            if (__any_sync(0xffffffffu, moved)) {
                // all P V products issued so far for this block must have landed in O: wait for the latest one
                const uint32_t lp = np - 1u;  // (moved implies t > 0, so np > 0)
                const uint32_t lpb = a.npb == 2 ? (lp & 1u) : 0u, luse = a.npb == 2 ? (lp >> 1) : lp;
                mbar_wait(sm.p_empty(w, (int)lpb), luse & 1u);
                tc_fence_after();
                uint32_t o[32];
                tc_ld16_nowait(tl + TF_O, o);
                tc_ld16_nowait(tl + TF_O + 16u, o + 16);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tc_st16(tl + TF_O, o);
                tc_st16(tl + TF_O + 16u, o + 16);
                tc_wait_st();
                tc_fence_before();
            }
            l *= alpha;
            const float mb = m_used == -INFINITY ? 0.f : m_used * c;
            uint32_t pp[32];
            float lsum = 0.f;
#pragma unroll
            for (int g = 0; g < 2; ++g) {  // two groups of 32 keys: one keep word each
                uint32_t kw = 0xFFFFFFFFu;
                if (DROP) kw = ick_keepword(rm, (uint32_t)((k0 >> 5) + g), t16);
#pragma unroll
                for (int i = 0; i < 16; ++i) {  // pair i of the group: keys 2i, 2i+1 at bits i, 16+i
                    const float p0 = ex2(fmaf(s[32 * g + 2 * i], c, -mb)), p1 = ex2(fmaf(s[32 * g + 2 * i + 1], c, -mb));
                    lsum += p0 + p1;
                    const float d0 = (!DROP || ((kw >> i) & 1u)) ? p0 : 0.f, d1 = (!DROP || ((kw >> (16 + i)) & 1u)) ? p1 : 0.f;
                    pp[16 * g + i] = pack2(d0, d1);
                }
            }
            l += lsum;
            mbar_wait(sm.p_empty(w, (int)pb), (use & 1u) ^ 1u);  // the product that last read this P buffer is done
            const uint32_t prow = sm.ptile(w, (int)pb) + (uint32_t)row * 128u;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) sts_u4(prow + (((uint32_t)ch ^ sw) << 4), pp[4 * ch], pp[4 * ch + 1], pp[4 * ch + 2], pp[4 * ch + 3]);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.p_full(w, (int)pb));
            ++np;
        }
        // epilogue of the block: O rows x 1/(l (1 - p)), LSE = m c + log2(l)
        mbar_wait(sm.o_full(w), nblk & 1u);
        tc_fence_after();
        uint32_t o[32];
        tc_ld16_nowait(tl + TF_O, o);
        tc_ld16_nowait(tl + TF_O + 16u, o + 16);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sm.o_empty(w));
        ++nblk;
        if (q_ok) {
            const float sc = drop.inv_keep / l;
            bf16* dst = O + ((size_t)b * d.Sq + q) * ldo + h * HD;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                uint4 u;
                u.x = pack2(__uint_as_float(o[8 * v]) * sc, __uint_as_float(o[8 * v + 1]) * sc);
                u.y = pack2(__uint_as_float(o[8 * v + 2]) * sc, __uint_as_float(o[8 * v + 3]) * sc);
                u.z = pack2(__uint_as_float(o[8 * v + 4]) * sc, __uint_as_float(o[8 * v + 5]) * sc);
                u.w = pack2(__uint_as_float(o[8 * v + 6]) * sc, __uint_as_float(o[8 * v + 7]) * sc);
                reinterpret_cast<uint4*>(dst)[v] = u;
            }
            LSE[((size_t)b * d.H + h) * d.Sq + q] = m_used * c + log2f(l);
        }
    }
}

template <bool DROP>
__global__ void __launch_bounds__(TF_THREADS, 1)
    fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                  bf16* __restrict__ O, float* __restrict__ LSE, int ldo, FwArgs a, DropCfg drop) {
    ick_pdl_launch();
    extern __shared__ uint8_t smem_raw[];
    FwSm sm;
    uint8_t* gen = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    sm.base = smem_u32(gen);
    sm.stage_bytes = a.stage_bytes;
    sm.npb = a.npb;
    sm.ptile_bytes = (uint32_t)(2 * a.npb * TF_PTILE);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = a.d.B * a.d.H;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(gen + 512);
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
        for (int s = 0; s < 2; ++s) {
            mbar_init(sm.full(s), 1);
            mbar_init(sm.empty(s), 2);  // one commit per issuer thread
        }
        for (int w = 0; w < 2; ++w) {
            for (int b = 0; b < 2; ++b) {
                mbar_init(sm.s_full(w, b), 1);
                mbar_init(sm.s_empty(w, b), 4);
                mbar_init(sm.p_full(w, b), 4);
                mbar_init(sm.p_empty(w, b), 1);
            }
            mbar_init(sm.o_full(w), 1);
            mbar_init(sm.o_empty(w), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    ick_pdl_wait();

    if (warp == 0) {
        if (lane == 0) tf_producer(sm, a, &tmQ, &tmK, &tmV, n_items);
    } else if (warp == 1 || warp == 2) {
        if (lane == 0) tf_issuer(sm, a, warp - 1, tmem_base, n_items);
    } else if (warp >= 4) {
        tf_softmax<DROP>(sm, a, warp, lane, tmem_base, O, LSE, ldo, drop, n_items);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

int tf_num_sms() {
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

// Plans and launches the tcgen05 forward; ICK_ERR_UNSUPPORTED (nothing launched) when the item's operands do not fit two shared-memory
// stages - the caller then takes the mma.sync kernels.
int ick_mha_fwd_tc(const void* Q, const void* K, const void* V, void* O, float* lse, int B, int H, int Sq, int Sk, int dh, int ldq, int ldk, int ldv,
                   int ldo, int causal, DropCfg dc, cudaStream_t stream) {
    if (dh > HD || (ldo % 8) != 0) return ICK_ERR_UNSUPPORTED;
    if ((((uintptr_t)K | (uintptr_t)V | (uintptr_t)Q | (uintptr_t)O) & 15) != 0) return ICK_ERR_UNSUPPORTED;
    FwArgs a;
    {
        const char* e = getenv("ICK_TF_DEBUG");
        a.dbg = e ? atoi(e) : 0;
    }
    a.d = make_dims(B, H, Sq, Sk, dh, causal);
    a.nqb = (Sq + TF_QB - 1) / TF_QB;
    a.ntq = 2 * a.nqb;  // whole 128-query blocks are staged (rows beyond Sq are zero-filled by TMA)
    a.ntk = (Sk + TF_KT - 1) / TF_KT;
    a.off_k = (uint32_t)a.ntq * TILE_BYTES;
    a.off_v = a.off_k + (uint32_t)a.ntk * TILE_BYTES;
    a.stage_bytes = a.off_v + (uint32_t)a.ntk * TILE_BYTES;
    const int avail = TF_SMEM_MAX - 1024 - TF_BAR_BYTES - 2 * (int)a.stage_bytes;
    if (avail >= 4 * TF_PTILE) a.npb = 2;
    else if (avail >= 2 * TF_PTILE) a.npb = 1;
    else return ICK_ERR_UNSUPPORTED;
    int rc;
    CUtensorMap tmQ, tmK, tmV;
    if ((rc = make_tmap3(&tmQ, Q, H, Sq, B, ldq))) return rc;
    if ((rc = make_tmap3(&tmK, K, H, Sk, B, ldk))) return rc;
    if ((rc = make_tmap3(&tmV, V, H, Sk, B, ldv))) return rc;
    const int smem = 1024 + TF_BAR_BYTES + 2 * a.npb * TF_PTILE + 2 * (int)a.stage_bytes;
    const int grid = B * H < tf_num_sms() ? B * H : tf_num_sms();
    const int v = dc.thr != 0u ? 1 : 0;
    static bool attr_done[2] = {false, false};
    if (!attr_done[v]) {
        cudaError_t e = v ? cudaFuncSetAttribute(fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM_MAX)
                          : cudaFuncSetAttribute(fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM_MAX);
        if (e != cudaSuccess) {
            ick_set_error("mha_fwd_tc: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
            return ICK_ERR_CUDA;
        }
        attr_done[v] = true;
    }
    if (v) ick_launch(fwd_tc_kernel<true>, grid, TF_THREADS, smem, stream)(tmQ, tmK, tmV, (bf16*)O, lse, ldo, a, dc);
    else ick_launch(fwd_tc_kernel<false>, grid, TF_THREADS, smem, stream)(tmQ, tmK, tmV, (bf16*)O, lse, ldo, a, dc);
    return ick_check_launch("mha_fwd_tc");
}
