// bf16 flash attention for head_dim <= 32 (10 heads x 30 in the reference): warp-level mma.sync m16n8k16 with fp32
// accumulation, the streamed operand staged by TMA.  Same contract as the CUDA-core kernels in attention.cu (which stay the
// fp32-parity path): head-layout rows, softmax in the exp2 domain, attention-probability dropout regenerated from the
// counter hash, log2-domain LSE saved.
//
//   ownership : a warp owns 16 "own" rows (queries in fwd / dQ, keys in dK-dV) whose A fragments are read straight from
//               global memory once.
//   scheduling: persistent CTAs (one per SM, 15 compute warps + 1 TMA producer warp) pipeline whole (image, head) items
//               through 2-6 shared-memory stages; a chunked one-CTA-per-block variant covers sequences longer than 640.
//   streaming : the other operand (K,V for fwd / dQ; Q,dO for dK-dV) is brought in by TMA as 64-row x 64-byte boxes of a
//               3-D tensor map (column, position, image), 64B-swizzled, one mbarrier per box pair.  Up to 10 tiles (640
//               positions, every shape of the reference) are resident at once, so one elected thread issues all copies up
//               front and the tile loop has no block-wide barrier; rows past the end of an image are zero-filled by TMA.
//   fragments : B operands by ldmatrix (.trans for the "P x tile" products) - conflict-free under the 64B swizzle.
//   fwd : S = Q K^T, P = ex2(S*c - m) re-used from the accumulator registers as the A operand of O += P V; dropout is
//         an AND of the packed bf16 pair with a mask built from one pair hash (1/keep folded into the final 1/l).
//   dQ  : S and dP = dO V^T recomputed per tile, dS = P (drop(dP) - D), dQ += dS K.   D = rowsum(dO * O) is produced here.
//   dKV : S^T = K Q^T and dP^T = V dO^T so that P^T / dS^T come out in A-operand layout for dV += P^T dO, dK += dS^T Q.
// With 30-wide heads the kernels are bound by instruction issue (ex2, the dropout hash, scaling FMAs), not by the tensor
// pipe or HBM; see DESIGN.md for the per-element instruction budget.
#include <cuda.h>

#include "common.cuh"
#include "attention_internal.h"

#include "attention_mma.cuh"

namespace {
using namespace ickattn;

// ---- per-warp tile bodies (shared by the chunked and the persistent kernels) ------------------------------------------------
struct FwdAcc {
    float m0, m1, l0, l1;  // running max of the RAW scores, running sum
    float o[4][4];
};
__device__ __forceinline__ void fwd_init(FwdAcc& a) {
    a.m0 = a.m1 = -INFINITY;
    a.l0 = a.l1 = 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n) a.o[n][0] = a.o[n][1] = a.o[n][2] = a.o[n][3] = 0.f;
}
// one 64-key tile: S = Q K^T, online softmax, dropout, O += P V
__device__ __forceinline__ void fwd_tile(FwdAcc& a, const uint32_t (*qa)[4], uint32_t kt, uint32_t vt, int k0, const OwnRows& r,
                                         const TileEnv& e) {
    const Dims& d = e.d;
    // causal: tiles entirely beyond this warp's last query contribute nothing
    if (d.causal && k0 > r.wrow + 15) return;
    const float c = e.c;
    const int tq = e.tq;
    float s[8][4];
    mma_a_tT<8>(s, qa, kt, 0, e.lo);
    // masking is needed only on the ragged last tile and on tiles that cross this warp's causal diagonal
    if (k0 + TK > d.Sk || (d.causal && k0 + TK - 1 > r.wrow)) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int key = k0 + 8 * j + 2 * tq + (x & 1);
                const bool vis = key < d.Sk && (!d.causal || key <= (x < 2 ? r.r0 : r.r1));
                if (!vis) s[j][x] = -INFINITY;
            }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
        mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(a.m0, mx0), mn1 = fmaxf(a.m1, mx1);
    // a row with no visible key so far keeps m = -inf; use 0 as the exponent base there (all p are ex2(-inf) = 0)
    const float e0 = mn0 == -INFINITY ? 0.f : mn0 * c, e1 = mn1 == -INFINITY ? 0.f : mn1 * c;
    const float c0f = ex2(a.m0 * c - e0), c1f = ex2(a.m1 * c - e1);
    a.l0 *= c0f;
    a.l1 *= c1f;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        a.o[n][0] *= c0f; a.o[n][1] *= c0f; a.o[n][2] *= c1f; a.o[n][3] *= c1f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s[j][0] = ex2(fmaf(s[j][0], c, -e0));
        s[j][1] = ex2(fmaf(s[j][1], c, -e0));
        s[j][2] = ex2(fmaf(s[j][2], c, -e1));
        s[j][3] = ex2(fmaf(s[j][3], c, -e1));
        a.l0 += s[j][0] + s[j][1];
        a.l1 += s[j][2] + s[j][3];
    }
    uint32_t pa[4][4];
    pack_p<8>(s, pa);
    if (e.thr != 0u) {
        // keep words of the two 32-key groups of this tile for the lane's two rows; pair i = 4 * (j & 3) + tq of a group sits at
        // bits (i, 16 + i): one shift puts the pair on the two sign bits that the byte permute expands
#pragma unroll
        for (int grp = 0; grp < 2; ++grp) {
            const uint32_t w0 = ick_keepword(r.rm0, (uint32_t)(k0 >> 5) + grp, e.t16) << (3 - tq);
            const uint32_t w1 = ick_keepword(r.rm1, (uint32_t)(k0 >> 5) + grp, e.t16) << (3 - tq);
#pragma unroll
            for (int jl = 0; jl < 4; ++jl) {
                const int j = 4 * grp + jl, kk = j >> 1, jj = j & 1;
                pa[kk][2 * jj] &= keep_mask_bf16x2(w0 << (12 - 4 * jl));
                pa[kk][2 * jj + 1] &= keep_mask_bf16x2(w1 << (12 - 4 * jl));
            }
        }
    }
    mma_p_t<8>(a.o, pa, vt, 0, e.lo);
    a.m0 = mn0;
    a.m1 = mn1;
}
__device__ __forceinline__ void fwd_finish(FwdAcc& a, bf16* Ob, int ldo, float* L, const OwnRows& r, const TileEnv& e) {
    a.l0 += __shfl_xor_sync(0xffffffffu, a.l0, 1);
    a.l0 += __shfl_xor_sync(0xffffffffu, a.l0, 2);
    a.l1 += __shfl_xor_sync(0xffffffffu, a.l1, 1);
    a.l1 += __shfl_xor_sync(0xffffffffu, a.l1, 2);
    store_slab(Ob, ldo, r.r0, r.r1, e.d.Sq, a.o, e.ik / a.l0, e.ik / a.l1, e.d.dh, e.tq);
    if (e.tq == 0) {
        if (r.r0 < e.d.Sq) L[r.r0] = a.m0 * e.c + log2f(a.l0);
        if (r.r1 < e.d.Sq) L[r.r1] = a.m1 * e.c + log2f(a.l1);
    }
}

// own-row prologue of dQ: D = rowsum(dO * O) (lane pair (2r, 2r+1) handles the two 16-column halves of own row r)
__device__ __forceinline__ void dq_rowdot(const bf16* Ob, int ldo, const bf16* Gb, int lddo, float* Dsum, int wrow, int Sq, int dh, int lane,
                                          float& D0, float& D1) {
    const int rr = wrow + (lane >> 1), cb = (lane & 1) * 16, g = lane >> 2;
    float acc = 0.f;
    if (rr < Sq) {
        const bf16* op = Ob + (size_t)rr * ldo + cb;
        const bf16* gp = Gb + (size_t)rr * lddo + cb;
        float x[8], y[8];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            ld8(op + 8 * v, x);
            ld8(gp + 8 * v, y);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (cb + 8 * v + i < dh) acc = fmaf(x[i], y[i], acc);
        }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if ((lane & 1) == 0 && rr < Sq) Dsum[rr] = acc;
    D0 = __shfl_sync(0xffffffffu, acc, 2 * g);
    D1 = __shfl_sync(0xffffffffu, acc, 2 * g + 16);
}
// one 64-key tile of dQ: S and dP recomputed, dS = P (drop(dP) - D), dQ += dS K
__device__ __forceinline__ void dq_tile(float (*dq)[4], const uint32_t (*qa)[4], const uint32_t (*ga)[4], uint32_t kt, uint32_t vt, int kt0,
                                        const OwnRows& r, float lse0, float lse1, float D0, float D1, const TileEnv& e) {
    const Dims& d = e.d;
    const int tq = e.tq;
#pragma unroll 1
    for (int sub = 0; sub < TK / SUB; ++sub) {
        const int k0 = kt0 + sub * SUB;
        if (d.causal && k0 > r.wrow + 15) break;
        // keys past Sk are zero rows (TMA fill): they add nothing to dQ = dS K, so only the causal diagonal needs a mask
        const bool need_mask = d.causal && k0 + SUB - 1 > r.wrow;
        float s[4][4], dp[4][4];
        mma_a_tT<4>(s, qa, kt, 4 * sub, e.lo);
        mma_a_tT<4>(dp, ga, vt, 4 * sub, e.lo);
        uint32_t kw0 = 0u, kw1 = 0u;
        if (e.thr != 0u) {
            kw0 = ick_keepword(r.rm0, (uint32_t)(k0 >> 5), e.t16) >> tq;
            kw1 = ick_keepword(r.rm1, (uint32_t)(k0 >> 5), e.t16) >> tq;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (e.thr != 0u) {
                // the sub-step's 32 keys are one key group: bits (4j + tq, 16 + 4j + tq) of the rows' keep words (hoisted: kw0/kw1)
                dp[j][0] = ((kw0 >> (4 * j)) & 1u) ? dp[j][0] * e.ik : 0.f;
                dp[j][1] = ((kw0 >> (16 + 4 * j)) & 1u) ? dp[j][1] * e.ik : 0.f;
                dp[j][2] = ((kw1 >> (4 * j)) & 1u) ? dp[j][2] * e.ik : 0.f;
                dp[j][3] = ((kw1 >> (16 + 4 * j)) & 1u) ? dp[j][3] * e.ik : 0.f;
            }
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                float p = ex2(fmaf(s[j][x], e.c, -(x < 2 ? lse0 : lse1)));
                if (need_mask && k0 + 8 * j + 2 * tq + (x & 1) > (x < 2 ? r.r0 : r.r1)) p = 0.f;
                s[j][x] = p * (dp[j][x] - (x < 2 ? D0 : D1));
            }
        }
        uint32_t pa[2][4];
        pack_p<4>(s, pa);
        mma_p_t<4>(dq, pa, kt, 4 * sub, e.lo);
    }
}

// one 64-query tile of dK/dV for the warp's 16 keys.  ls / ds / rm: LSE, D and dropout row mix of the tile's queries (smem).
// Lanes g and g^1 own keys of the same dropout pair and see the same queries: each computes the pair hashes of ONE of its two
// key rows (even g: r0, odd g: r1) and receives the other from its partner (lane ^ 4).
// ds_out (optional): where this warp's key slab keeps dS^T for the dQ-from-dS kernel, [query sub-step][lane][8 words] in exactly
// the A-fragment order produced here (1 KiB per 16-key x 32-query block, two 16-byte stores per lane).
__device__ __forceinline__ void dkv_tile(float (*dk)[4], float (*dv)[4], const uint32_t (*ka)[4], const uint32_t (*va)[4], uint32_t qt,
                                         uint32_t gt, int qt0, const float* ls, const float* ds, const uint32_t* rm, const OwnRows& r,
                                         bool odd, const TileEnv& e, uint4* ds_out = nullptr) {
    const Dims& d = e.d;
    const int tq = e.tq;
    // the warp's 16 keys lie in one 32-key group; key r0 at bit b0 of a query's keep word, r1 = r0 + 8 four bits higher
    const uint32_t kgrp = (uint32_t)r.wrow >> 5;
    const uint32_t mk0 = 1u << ick_keybit((uint32_t)r.r0), mk1 = mk0 << 4;
    (void)odd;
#pragma unroll 1
    for (int sub = 0; sub < TK / SUB; ++sub) {
        const int q0 = qt0 + sub * SUB;                      // first query of the sub-step
        if (d.causal && q0 + SUB - 1 < r.wrow) continue;     // every query precedes every key of this warp
        // queries past Sq are zero rows of Q and dO (TMA fill) and add nothing; own keys past Sk are never stored
        const bool need_mask = d.causal && r.wrow + 15 > q0;
        const float* lsp = ls + sub * SUB + 2 * tq;
        const float* dsp = ds + sub * SUB + 2 * tq;
        const uint32_t* rmp = rm + sub * SUB + 2 * tq;
        float st[4][4], dpt[4][4];
        mma_a_tT<4>(st, ka, qt, 4 * sub, e.lo);
        mma_a_tT<4>(dpt, va, gt, 4 * sub, e.lo);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 l2 = *reinterpret_cast<const float2*>(lsp + 8 * j);
            const float2 d2 = *reinterpret_cast<const float2*>(dsp + 8 * j);
            uint32_t kwq0 = 0u, kwq1 = 0u;  // keep words of the two queries of this lane's pair
            if (e.thr != 0u) {
                const uint2 r2 = *reinterpret_cast<const uint2*>(rmp + 8 * j);
                kwq0 = ick_keepword(r2.x, kgrp, e.t16);
                kwq1 = ick_keepword(r2.y, kgrp, e.t16);
            }
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int qi = q0 + 8 * j + 2 * tq + (x & 1);
                float p = ex2(fmaf(st[j][x], e.c, -((x & 1) ? l2.y : l2.x)));
                if (need_mask && (x < 2 ? r.r0 : r.r1) > qi) p = 0.f;
                float gp = dpt[j][x];
                float pm = p;
                if (e.thr != 0u) {
                    const bool keep = (((x & 1) ? kwq1 : kwq0) & (x < 2 ? mk0 : mk1)) != 0u;
                    pm = keep ? p : 0.f;
                    gp = keep ? gp * e.ik : 0.f;
                }
                st[j][x] = pm;                                   // P^T with dropout (1/keep applied at the store) -> dV
                dpt[j][x] = p * (gp - ((x & 1) ? d2.y : d2.x));  // dS^T -> dK
            }
        }
        uint32_t pa[2][4];
        pack_p<4>(st, pa);
        mma_p_t<4>(dv, pa, gt, 4 * sub, e.lo);
        pack_p<4>(dpt, pa);
        if (ds_out != nullptr) {
            uint4* dst = ds_out + ((size_t)(q0 / SUB) * 32 + (size_t)(4 * (r.r0 - r.wrow) + tq)) * 2;
            dst[0] = make_uint4(pa[0][0], pa[0][1], pa[0][2], pa[0][3]);
            dst[1] = make_uint4(pa[1][0], pa[1][1], pa[1][2], pa[1][3]);
        }
        mma_p_t<4>(dk, pa, qt, 4 * sub, e.lo);
    }
}
// =============================================================================================================================
// Chunked kernels: one CTA per (own block, head, image), up to CH tiles resident, re-filled chunk by chunk.  Any length; used
// when an image has more than CH streamed tiles (the persistent kernels below cover every shape of the reference).
// =============================================================================================================================
__global__ void __launch_bounds__(32 * NWMAX, 2)
    fwd_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const bf16* __restrict__ Q,
               bf16* __restrict__ O, float* __restrict__ LSE, Dims d, int ldq, int ldo, int ntc, DropCfg drop) {
    ick_pdl_entry();
    extern __shared__ uint8_t smem_raw[];
    const Smem sm = carve(smem_raw, ntc);
    ick_resolve_seed(drop);
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * nw * 16;
    const int kend = d.causal ? min(d.Sk, q0 + nw * 16) : d.Sk;
    const int nt = (kend + TK - 1) / TK;
    init_bars(sm, ntc, &tmK, &tmV);
    if (threadIdx.x == 0) issue_tiles(sm, &tmK, &tmV, 0, min(nt, ntc), h, b);

    const bool active = q0 + 16 * warp < d.Sq;
    const TileEnv env = make_env(d, drop, lane);
    const OwnRows r = own_rows(q0 + 16 * warp, g, drop, b, d.H, h, d.Sq, true);
    uint32_t qa[2][4];
    load_own(Q + (size_t)b * d.Sq * ldq + h * HD, ldq, r.r0, r.r1, d.Sq, d.dh, tq, qa);
    FwdAcc acc;
    fwd_init(acc);
    for (int c0 = 0; c0 < nt; c0 += ntc) {
        const int n = min(ntc, nt - c0);
        if (c0 > 0) {
            __syncthreads();  // every warp is done with the resident tiles
            if (threadIdx.x == 0) issue_tiles(sm, &tmK, &tmV, c0, n, h, b);
        }
        const uint32_t parity = (uint32_t)(c0 / ntc) & 1u;
        for (int t = 0; t < n; ++t) {
            mbar_wait(sm.bars + 8 * t, parity);
            if (active) fwd_tile(acc, qa, sm.t0 + t * TILE_BYTES, sm.t1 + t * TILE_BYTES, (c0 + t) * TK, r, env);
        }
    }
    if (active) fwd_finish(acc, O + (size_t)b * d.Sq * ldo + h * HD, ldo, LSE + ((size_t)b * d.H + h) * d.Sq, r, env);
}

__global__ void __launch_bounds__(32 * NWMAX, 2)
    bwd_dq_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const bf16* __restrict__ Q,
                  const bf16* __restrict__ O, const bf16* __restrict__ dO, const float* __restrict__ LSE, float* __restrict__ Dsum,
                  bf16* __restrict__ dQ, Dims d, int ldq, int ldo, int lddo, int lddq, int ntc, DropCfg drop) {
    ick_pdl_entry();
    extern __shared__ uint8_t smem_raw[];
    const Smem sm = carve(smem_raw, ntc);
    ick_resolve_seed(drop);
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * nw * 16;
    const int kend = d.causal ? min(d.Sk, q0 + nw * 16) : d.Sk;
    const int nt = (kend + TK - 1) / TK;
    init_bars(sm, ntc, &tmK, &tmV);
    if (threadIdx.x == 0) issue_tiles(sm, &tmK, &tmV, 0, min(nt, ntc), h, b);

    const bool active = q0 + 16 * warp < d.Sq;
    const TileEnv env = make_env(d, drop, lane);
    const OwnRows r = own_rows(q0 + 16 * warp, g, drop, b, d.H, h, d.Sq, true);
    const bf16* Gb = dO + (size_t)b * d.Sq * lddo + h * HD;
    uint32_t qa[2][4], ga[2][4];
    load_own(Q + (size_t)b * d.Sq * ldq + h * HD, ldq, r.r0, r.r1, d.Sq, d.dh, tq, qa);
    load_own(Gb, lddo, r.r0, r.r1, d.Sq, d.dh, tq, ga);
    float D0, D1;
    dq_rowdot(O + (size_t)b * d.Sq * ldo + h * HD, ldo, Gb, lddo, Dsum + ((size_t)b * d.H + h) * d.Sq, r.wrow, d.Sq, d.dh, lane, D0, D1);
    const float* L = LSE + ((size_t)b * d.H + h) * d.Sq;
    const float lse0 = r.r0 < d.Sq ? L[r.r0] : 0.f, lse1 = r.r1 < d.Sq ? L[r.r1] : 0.f;
    float dq[4][4];
    zero16(dq);
    for (int c0 = 0; c0 < nt; c0 += ntc) {
        const int n = min(ntc, nt - c0);
        if (c0 > 0) {
            __syncthreads();
            if (threadIdx.x == 0) issue_tiles(sm, &tmK, &tmV, c0, n, h, b);
        }
        const uint32_t parity = (uint32_t)(c0 / ntc) & 1u;
        for (int t = 0; t < n; ++t) {
            mbar_wait(sm.bars + 8 * t, parity);
            if (active) dq_tile(dq, qa, ga, sm.t0 + t * TILE_BYTES, sm.t1 + t * TILE_BYTES, (c0 + t) * TK, r, lse0, lse1, D0, D1, env);
        }
    }
    if (active) store_slab(dQ + (size_t)b * d.Sq * lddq + h * HD, lddq, r.r0, r.r1, d.Sq, dq, d.scale, d.scale, d.dh, tq);
}

// fill the per-query scalars of `n` resident tiles starting at query q_first (any number of threads: tid / nthr)
__device__ __forceinline__ void fill_scalars(float* Ls, float* Ds, uint32_t* Rm, const float* L, const float* Dg, int q_first, int n, int Sq,
                                             const DropCfg& drop, uint64_t row_base, int tid, int nthr) {
    for (int i = tid; i < n * TK; i += nthr) {
        const int qi = q_first + i;
        const bool ok = qi < Sq;
        Ls[i] = ok ? L[qi] : 0.f;
        Ds[i] = ok ? Dg[qi] : 0.f;
        Rm[i] = ick_rowmix(drop.seed, drop.site, row_base + (uint64_t)qi);
    }
}

__global__ void __launch_bounds__(32 * NWMAX, 2)
    bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG, const bf16* __restrict__ K,
                   const bf16* __restrict__ V, const float* __restrict__ LSE, const float* __restrict__ Dsum, bf16* __restrict__ dK,
                   bf16* __restrict__ dV, Dims d, int ldk, int ldv, int lddk, int lddv, int ntc, DropCfg drop, uint4* __restrict__ DS) {
    ick_pdl_entry();
    extern __shared__ uint8_t smem_raw[];
    const Smem sm = carve(smem_raw, ntc);
    float* Ls = sm.scal;                        // log2-domain LSE of each resident query
    float* Ds = Ls + ntc * TK;                  // D of each resident query
    uint32_t* Rm = (uint32_t*)(Ds + ntc * TK);  // dropout row mix of each resident query
    ick_resolve_seed(drop);
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * nw * 16;
    // causal: queries before the first key of this CTA see none of its keys
    const int tbeg = d.causal ? j0 / TK : 0;
    const int nt = (d.Sq + TK - 1) / TK;
    init_bars(sm, ntc, &tmQ, &tmG);
    if (threadIdx.x == 0) issue_tiles(sm, &tmQ, &tmG, tbeg, min(nt - tbeg, ntc), h, b);

    const bool active = j0 + 16 * warp < d.Sk;
    const TileEnv env = make_env(d, drop, lane);
    const OwnRows r = own_rows(j0 + 16 * warp, g, drop, b, d.H, h, d.Sq, false);
    uint32_t ka[2][4], va[2][4];
    load_own(K + (size_t)b * d.Sk * ldk + h * HD, ldk, r.r0, r.r1, d.Sk, d.dh, tq, ka);
    load_own(V + (size_t)b * d.Sk * ldv + h * HD, ldv, r.r0, r.r1, d.Sk, d.dh, tq, va);
    const float* L = LSE + ((size_t)b * d.H + h) * d.Sq;
    const float* Dg = Dsum + ((size_t)b * d.H + h) * d.Sq;
    const bool odd = (g & 1) != 0;
    float dk[4][4], dv[4][4];
    zero16(dk);
    zero16(dv);
    const int nsub_all = 2 * nt, nslabs_all = (d.Sk + 15) / 16;
    uint4* ds_out = DS ? DS + (((size_t)b * d.H + h) * nslabs_all + (size_t)(j0 / 16 + warp)) * nsub_all * 64 : nullptr;
    for (int c0 = tbeg; c0 < nt; c0 += ntc) {
        const int n = min(ntc, nt - c0);
        if (c0 > tbeg) {
            __syncthreads();
            if (threadIdx.x == 0) issue_tiles(sm, &tmQ, &tmG, c0, n, h, b);
        }
        fill_scalars(Ls, Ds, Rm, L, Dg, c0 * TK, n, d.Sq, drop, prob_row(b, d.H, h, d.Sq, 0), threadIdx.x, blockDim.x);
        __syncthreads();
        const uint32_t parity = (uint32_t)((c0 - tbeg) / ntc) & 1u;
        for (int t = 0; t < n; ++t) {
            mbar_wait(sm.bars + 8 * t, parity);
            if (active)
                dkv_tile(dk, dv, ka, va, sm.t0 + t * TILE_BYTES, sm.t1 + t * TILE_BYTES, (c0 + t) * TK, Ls + t * TK, Ds + t * TK, Rm + t * TK, r,
                         odd, env, ds_out);
        }
    }
    if (!active) return;
    store_slab(dK + (size_t)b * d.Sk * lddk + h * HD, lddk, r.r0, r.r1, d.Sk, dk, d.scale, d.scale, d.dh, tq);
    store_slab(dV + (size_t)b * d.Sk * lddv + h * HD, lddv, r.r0, r.r1, d.Sk, dv, env.ik, env.ik, d.dh, tq);
}

// =============================================================================================================================
// Persistent kernels: one CTA per SM, PNW compute warps + one producer warp.  A work item is one (image, head): all of its
// streamed tiles (<= CH) form one pipeline stage; the producer keeps `nstage` items in flight (TMA into the stage as soon as
// every compute warp has released it), so the copies of item i+1.. overlap the math of item i and nothing but the very first
// load is exposed.  The 16-row own slabs of successive items are dealt round-robin to the compute warps (global slab number
// modulo PNW), which keeps the warps balanced to within one slab over the whole launch.
// =============================================================================================================================
constexpr int PNW_KV = 15;  // compute warps of the dK/dV kernel (128 registers per thread)
constexpr int PNW_Q = 19;   // compute warps of the forward / dQ kernels (their tile bodies fit in 96 registers: more warps to hide latency)
constexpr int PBAR = CH + 2;  // mbarriers of a stage: full[CH] (one per tile), scalars-ready, empty
struct Pipe {
    uint32_t bars, data, stage_bytes;
    uint8_t* gen;  // generic pointer to `data`
    int ntc;
    __device__ __forceinline__ uint32_t full(int s, int t) const { return bars + 8 * (s * PBAR + t); }
    __device__ __forceinline__ uint32_t sfull(int s) const { return bars + 8 * (s * PBAR + CH); }
    __device__ __forceinline__ uint32_t empty(int s) const { return bars + 8 * (s * PBAR + CH + 1); }
    __device__ __forceinline__ uint32_t t0(int s) const { return data + s * stage_bytes; }
    __device__ __forceinline__ uint32_t t1(int s) const { return t0(s) + ntc * TILE_BYTES; }
    __device__ __forceinline__ float* scal(int s) const { return (float*)(gen + (size_t)s * stage_bytes + 2 * ntc * TILE_BYTES); }
};
template <int PNW>
__device__ __forceinline__ Pipe make_pipe(uint8_t* raw, int ntc, int nstage, uint32_t stage_bytes, const CUtensorMap* a, const CUtensorMap* b) {
    uint8_t* p = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    Pipe pp;
    pp.bars = smem_u32(p);
    pp.data = pp.bars + 1024;
    pp.gen = p + 1024;
    pp.stage_bytes = stage_bytes;
    pp.ntc = ntc;
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(b) : "memory");
        for (int s = 0; s < nstage; ++s) {
            for (int t = 0; t < ntc; ++t) mbar_init(pp.full(s, t), 1);
            mbar_init(pp.sfull(s), 32);
            mbar_init(pp.empty(s), PNW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    return pp;
}
__device__ __forceinline__ void produce_item(const Pipe& pp, int s, const CUtensorMap* a, const CUtensorMap* b2, int h, int b) {
    for (int t = 0; t < pp.ntc; ++t) {
        const uint32_t bar = pp.full(s, t);
        mbar_expect_tx(bar, 2 * TILE_BYTES);
        tma_load_3d(pp.t0(s) + t * TILE_BYTES, a, bar, h * HD, t * TK, b);
        tma_load_3d(pp.t1(s) + t * TILE_BYTES, b2, bar, h * HD, t * TK, b);
    }
}
struct PArgs {
    Dims d;
    int nslabs, ntc, nstage;
    uint32_t stage_bytes;
};

template <int PNW>
__global__ void __launch_bounds__(32 * (PNW + 1), 1)
    fwd_pkernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const bf16* __restrict__ Q,
                bf16* __restrict__ O, float* __restrict__ LSE, PArgs a, int ldq, int ldo, DropCfg drop) {
    ick_pdl_entry();
    extern __shared__ uint8_t smem_raw[];
    const Dims& d = a.d;
    const Pipe pp = make_pipe<PNW>(smem_raw, a.ntc, a.nstage, a.stage_bytes, &tmK, &tmV);
    ick_resolve_seed(drop);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int n_items = d.B * d.H;
    if (warp == PNW) {
        if (lane == 0) {
            int li = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
                const int s = li % a.nstage;
                mbar_wait(pp.empty(s), ((uint32_t)(li / a.nstage) & 1u) ^ 1u);
                produce_item(pp, s, &tmK, &tmV, item % d.H, item / d.H);
            }
        }
        return;
    }
    const TileEnv env = make_env(d, drop, lane);
    int li = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
        const int s = li % a.nstage, b = item / d.H, h = item % d.H;
        const uint32_t ph = (uint32_t)(li / a.nstage) & 1u;
        for (int slab = first_slab<PNW>(li, a.nslabs, warp); slab < a.nslabs; slab += PNW) {
            const OwnRows r = own_rows(16 * slab, g, drop, b, d.H, h, d.Sq, true);
            uint32_t qa[2][4];
            load_own(Q + (size_t)b * d.Sq * ldq + h * HD, ldq, r.r0, r.r1, d.Sq, d.dh, tq, qa);
            FwdAcc acc;
            fwd_init(acc);
            const int nt = d.causal ? (min(d.Sk, 16 * slab + 16) + TK - 1) / TK : a.ntc;
            for (int t = 0; t < nt; ++t) {
                mbar_wait(pp.full(s, t), ph);
                fwd_tile(acc, qa, pp.t0(s) + t * TILE_BYTES, pp.t1(s) + t * TILE_BYTES, t * TK, r, env);
            }
            fwd_finish(acc, O + (size_t)b * d.Sq * ldo + h * HD, ldo, LSE + ((size_t)b * d.H + h) * d.Sq, r, env);
        }
        // Every warp passes through every item in order (a warp without a slab here still waits for the item's first tile), so
        // no warp can release a stage for the item that re-uses it before the producer has started that item.
        mbar_wait(pp.full(s, 0), ph);
        __syncwarp();
        if (lane == 0) mbar_arrive(pp.empty(s));
    }
}

template <int PNW>
__global__ void __launch_bounds__(32 * (PNW + 1), 1)
    bwd_dq_pkernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const bf16* __restrict__ Q,
                   const bf16* __restrict__ O, const bf16* __restrict__ dO, const float* __restrict__ LSE, float* __restrict__ Dsum,
                   bf16* __restrict__ dQ, PArgs a, int ldq, int ldo, int lddo, int lddq, DropCfg drop) {
    ick_pdl_entry();
    extern __shared__ uint8_t smem_raw[];
    const Dims& d = a.d;
    const Pipe pp = make_pipe<PNW>(smem_raw, a.ntc, a.nstage, a.stage_bytes, &tmK, &tmV);
    ick_resolve_seed(drop);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int n_items = d.B * d.H;
    if (warp == PNW) {
        if (lane == 0) {
            int li = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
                const int s = li % a.nstage;
                mbar_wait(pp.empty(s), ((uint32_t)(li / a.nstage) & 1u) ^ 1u);
                produce_item(pp, s, &tmK, &tmV, item % d.H, item / d.H);
            }
        }
        return;
    }
    const TileEnv env = make_env(d, drop, lane);
    int li = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
        const int s = li % a.nstage, b = item / d.H, h = item % d.H;
        const uint32_t ph = (uint32_t)(li / a.nstage) & 1u;
        const bf16* Gb = dO + (size_t)b * d.Sq * lddo + h * HD;
        const float* L = LSE + ((size_t)b * d.H + h) * d.Sq;
        for (int slab = first_slab<PNW>(li, a.nslabs, warp); slab < a.nslabs; slab += PNW) {
            const OwnRows r = own_rows(16 * slab, g, drop, b, d.H, h, d.Sq, true);
            uint32_t qa[2][4], ga[2][4];
            load_own(Q + (size_t)b * d.Sq * ldq + h * HD, ldq, r.r0, r.r1, d.Sq, d.dh, tq, qa);
            load_own(Gb, lddo, r.r0, r.r1, d.Sq, d.dh, tq, ga);
            float D0, D1;
            dq_rowdot(O + (size_t)b * d.Sq * ldo + h * HD, ldo, Gb, lddo, Dsum + ((size_t)b * d.H + h) * d.Sq, r.wrow, d.Sq, d.dh, lane, D0, D1);
            const float lse0 = r.r0 < d.Sq ? L[r.r0] : 0.f, lse1 = r.r1 < d.Sq ? L[r.r1] : 0.f;
            float dq[4][4];
            zero16(dq);
            const int nt = d.causal ? (min(d.Sk, 16 * slab + 16) + TK - 1) / TK : a.ntc;
            for (int t = 0; t < nt; ++t) {
                mbar_wait(pp.full(s, t), ph);
                dq_tile(dq, qa, ga, pp.t0(s) + t * TILE_BYTES, pp.t1(s) + t * TILE_BYTES, t * TK, r, lse0, lse1, D0, D1, env);
            }
            store_slab(dQ + (size_t)b * d.Sq * lddq + h * HD, lddq, r.r0, r.r1, d.Sq, dq, d.scale, d.scale, d.dh, tq);
        }
        mbar_wait(pp.full(s, 0), ph);
        __syncwarp();
        if (lane == 0) mbar_arrive(pp.empty(s));
    }
}

template <int PNW>
__global__ void __launch_bounds__(32 * (PNW + 1), 1)
    bwd_dkv_pkernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG, const bf16* __restrict__ K,
                    const bf16* __restrict__ V, const float* __restrict__ LSE, const float* __restrict__ Dsum, bf16* __restrict__ dK,
                    bf16* __restrict__ dV, PArgs a, int ldk, int ldv, int lddk, int lddv, DropCfg drop, uint4* __restrict__ DS) {
    ick_pdl_entry();
    extern __shared__ uint8_t smem_raw[];
    const Dims& d = a.d;
    const Pipe pp = make_pipe<PNW>(smem_raw, a.ntc, a.nstage, a.stage_bytes, &tmQ, &tmG);
    ick_resolve_seed(drop);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int n_items = d.B * d.H;
    if (warp == PNW) {  // the whole producer warp: lane 0 issues the copies, all lanes fill the per-query scalars of the stage
        int li = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
            const int s = li % a.nstage, b = item / d.H, h = item % d.H;
            mbar_wait(pp.empty(s), ((uint32_t)(li / a.nstage) & 1u) ^ 1u);
            if (lane == 0) produce_item(pp, s, &tmQ, &tmG, h, b);
            float* Ls = pp.scal(s);
            float* Ds = Ls + a.ntc * TK;
            uint32_t* Rm = (uint32_t*)(Ds + a.ntc * TK);
            fill_scalars(Ls, Ds, Rm, LSE + ((size_t)b * d.H + h) * d.Sq, Dsum + ((size_t)b * d.H + h) * d.Sq, 0, a.ntc, d.Sq, drop,
                         prob_row(b, d.H, h, d.Sq, 0), lane, 32);
            mbar_arrive(pp.sfull(s));  // release: this lane's scalars are visible to whoever observes the phase
        }
        return;
    }
    const TileEnv env = make_env(d, drop, lane);
    const bool odd = (g & 1) != 0;
    int li = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
        const int s = li % a.nstage, b = item / d.H, h = item % d.H;
        const uint32_t ph = (uint32_t)(li / a.nstage) & 1u;
        const float* Ls = pp.scal(s);
        const float* Ds = Ls + a.ntc * TK;
        const uint32_t* Rm = (const uint32_t*)(Ds + a.ntc * TK);
        for (int slab = first_slab<PNW>(li, a.nslabs, warp); slab < a.nslabs; slab += PNW) {
            const OwnRows r = own_rows(16 * slab, g, drop, b, d.H, h, d.Sq, false);
            uint32_t ka[2][4], va[2][4];
            load_own(K + (size_t)b * d.Sk * ldk + h * HD, ldk, r.r0, r.r1, d.Sk, d.dh, tq, ka);
            load_own(V + (size_t)b * d.Sk * ldv + h * HD, ldv, r.r0, r.r1, d.Sk, d.dh, tq, va);
            float dk[4][4], dv[4][4];
            zero16(dk);
            zero16(dv);
            uint4* ds_out = DS ? DS + ((size_t)item * a.nslabs + slab) * (size_t)(2 * a.ntc) * 64 : nullptr;
            mbar_wait(pp.sfull(s), ph);
            // causal: queries before the warp's first key see none of its keys
            for (int t = d.causal ? (16 * slab) / TK : 0; t < a.ntc; ++t) {
                mbar_wait(pp.full(s, t), ph);
                dkv_tile(dk, dv, ka, va, pp.t0(s) + t * TILE_BYTES, pp.t1(s) + t * TILE_BYTES, t * TK, Ls + t * TK, Ds + t * TK, Rm + t * TK, r, odd,
                         env, ds_out);
            }
            store_slab(dK + (size_t)b * d.Sk * lddk + h * HD, lddk, r.r0, r.r1, d.Sk, dk, d.scale, d.scale, d.dh, tq);
            store_slab(dV + (size_t)b * d.Sk * lddv + h * HD, lddv, r.r0, r.r1, d.Sk, dv, env.ik, env.ik, d.dh, tq);
        }
        mbar_wait(pp.full(s, 0), ph);
        mbar_wait(pp.sfull(s), ph);
        __syncwarp();
        if (lane == 0) mbar_arrive(pp.empty(s));
    }
}

// =============================================================================================================================
// dQ from the stored dS^T.  The dK/dV kernel has every dS^T block in registers; recomputing S, P, dP and the dropout mask a second
// time just to contract dS with K (the two-kernel scheme above) costs ~40% of the backward.  With a workspace the order becomes
//   D = rowsum(dO * O)  ->  dK/dV kernel (also stores dS^T, bf16, A-fragment order)  ->  dQ = scale * dS K  (this kernel),
// a plain batched GEMM per (image, head): one warp per 32 queries streams the key slabs, turns the stored fragments into
// A fragments of dS with movmatrix (8x8 transposes inside the warp), and accumulates 32 x 32 in registers.
// =============================================================================================================================
__device__ __forceinline__ uint32_t movm_t(uint32_t x) {
    uint32_t y;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__global__ void __launch_bounds__(256) rowdot_kernel(const bf16* __restrict__ O, const bf16* __restrict__ dO, float* __restrict__ Dsum, int B,
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

constexpr int DQ_MAXW = 20;  // warps per CTA of the dQ-from-dS kernel (one CTA per (image, head)): 640 threads x <= 102 registers
constexpr int DQ_KLD = 40;   // shared-memory row stride of the K tile in elements: 80 bytes make every ldmatrix phase conflict-free
// Warp w works on query block w % nsub and on part w / nsub of the key slabs (ksplit parts; partial 32 x 32 tiles are summed
// through shared memory), so that even a 102-query cross-attention item keeps 16 warps' worth of loads in flight.  When there
// are more query blocks than warps (Sq > 640) ksplit is 1 and the warps loop over the blocks.
__global__ void __launch_bounds__(32 * DQ_MAXW) bwd_dq_ds_kernel(const uint4* __restrict__ DS, const bf16* __restrict__ K, bf16* __restrict__ dQ,
                                                                 Dims d, int ldk, int lddq, int nslabs, int nsub, int ksplit) {
    ick_pdl_entry();
    extern __shared__ __align__(16) uint8_t dq_smem[];
    bf16* ks = reinterpret_cast<bf16*>(dq_smem);  // [16 * nslabs][DQ_KLD], rows >= Sk and columns >= dh zero
    float* red = reinterpret_cast<float*>(dq_smem + (size_t)16 * nslabs * DQ_KLD * 2);  // [warps][32 values][32 lanes] (ksplit > 1)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3, nw = blockDim.x >> 5;
    const int item = blockIdx.x, b = item / d.H, h = item % d.H;
    const bf16* Kb = K + (size_t)b * d.Sk * ldk + h * HD;
    for (int idx = threadIdx.x; idx < 16 * nslabs * 4; idx += blockDim.x) {
        const int r = idx >> 2, c = (idx & 3) * 8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < d.Sk) {
            v = *reinterpret_cast<const uint4*>(Kb + (size_t)r * ldk + c);
            if (c + 8 > d.dh) {  // zero the pad lanes of the head (dh = 30: columns 30, 31)
                bf16* e = reinterpret_cast<bf16*>(&v);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (c + i >= d.dh) e[i] = __float2bfloat16_rn(0.f);
            }
        }
        *reinterpret_cast<uint4*>(ks + (size_t)r * DQ_KLD + c) = v;
    }
    __syncthreads();
    // per-lane ldmatrix.trans address of a (16 keys x 16 d) block: row (lane&7) + 8*((lane>>3)&1), column 8*(lane>>4)
    const uint32_t ks_lane = smem_u32(ks) + (uint32_t)(((lane & 7) + 8 * ((lane >> 3) & 1)) * DQ_KLD + 8 * (lane >> 4)) * 2u;
    bf16* out = dQ + (size_t)b * d.Sq * lddq + h * HD;
    const int per = (nslabs + ksplit - 1) / ksplit;
    for (int task = warp; task < nsub * ksplit; task += nw) {
        const int sub = task % nsub, part = task / nsub;
        const int q0 = sub * SUB;
        // causal: the dK/dV kernel skipped (never wrote) the blocks whose queries all precede the slab's keys
        const int slab_lim = d.causal ? min(nslabs, (q0 + SUB - 1) / 16 + 1) : nslabs;
        const int slab_beg = part * per, slab_end = q0 < d.Sq ? min(slab_lim, slab_beg + per) : 0;
        const uint4* src = DS + (((size_t)item * nslabs) * nsub + sub) * 64 + lane * 2;
        const size_t sstride = (size_t)nsub * 64;
        float acc[2][4][4];
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[kk][j][0] = acc[kk][j][1] = acc[kk][j][2] = acc[kk][j][3] = 0.f;
#pragma unroll 4
        for (int slab = slab_beg; slab < slab_end; ++slab) {
            const uint4 v0 = src[slab * sstride], v1 = src[slab * sstride + 1];
            // B operand K[key][d] (key pairs along the reduction) by ldmatrix.trans: {b0,b1} of two d tiles per instruction
            uint32_t kb[4][2];
            const uint32_t ka = ks_lane + (uint32_t)(16 * slab * DQ_KLD) * 2u;
            ldsm_x4_trans(kb[0][0], kb[0][1], kb[1][0], kb[1][1], ka);
            ldsm_x4_trans(kb[2][0], kb[2][1], kb[3][0], kb[3][1], ka + 32u);
            const uint32_t pa[2][4] = {{v0.x, v0.y, v0.z, v0.w}, {v1.x, v1.y, v1.z, v1.w}};
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                // A fragment of dS (16 queries x 16 keys) from the stored dS^T blocks: (keys 0-7 | 8-15) x (queries 0-7 | 8-15)
                uint32_t da[4];
                da[0] = movm_t(pa[kk][0]);
                da[1] = movm_t(pa[kk][2]);
                da[2] = movm_t(pa[kk][1]);
                da[3] = movm_t(pa[kk][3]);
#pragma unroll
                for (int j = 0; j < 4; ++j) mma16816(acc[kk][j], da, kb[j][0], kb[j][1]);
            }
        }
        if (ksplit > 1) {  // (all tasks fit in one pass of the warps: the host guarantees nsub * ksplit <= warps)
            float* mine = red + (size_t)warp * 1024 + lane;
            if (part != 0) {
#pragma unroll
                for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int x = 0; x < 4; ++x) mine[((kk * 4 + j) * 4 + x) * 32] = acc[kk][j][x];
            }
            __syncthreads();
            if (part != 0) continue;
            for (int pp = 1; pp < ksplit; ++pp) {
                const float* other = red + (size_t)(warp + pp * nsub) * 1024 + lane;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int x = 0; x < 4; ++x) acc[kk][j][x] += other[((kk * 4 + j) * 4 + x) * 32];
            }
        }
        if (q0 >= d.Sq) continue;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) store_slab(out, lddq, q0 + 16 * kk + g, q0 + 16 * kk + g + 8, d.Sq, acc[kk], d.scale, d.scale, d.dh, tq);
    }
}

// own rows -> (CTAs along the own dimension, warps per CTA): slabs of 16 rows spread evenly over the fewest CTAs
void split_own(int S, int* nctas, int* nw) {
    const int slabs = (S + 15) / 16;
    *nctas = (slabs + NWMAX - 1) / NWMAX;
    *nw = (slabs + *nctas - 1) / *nctas;
}
int smem_bytes(int ntc, bool scalars) { return 1024 + 1024 + 2 * ntc * TILE_BYTES + (scalars ? 3 * ntc * TK * 4 : 0); }

constexpr int SMEM_MAX = 232448;  // 227 KiB: largest dynamic shared memory of a CTA on sm_100
// persistent-kernel plan for `ntc` streamed tiles per item: stage size and stage count (0 = does not fit, use the chunked kernel)
void plan_pipe(int ntc, bool scalars, PArgs* a) {
    a->ntc = ntc;
    a->stage_bytes = (uint32_t)((2 * ntc * TILE_BYTES + (scalars ? 3 * ntc * TK * 4 : 0) + 1023) / 1024 * 1024);
    int ns = ntc <= CH ? (SMEM_MAX - 2048) / (int)a->stage_bytes : 0;
    a->nstage = ns > 6 ? 6 : (ns < 2 ? 0 : ns);
}
int pipe_smem(const PArgs& a) { return 2048 + a.nstage * (int)a.stage_bytes; }
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
bool wide_q() {  // ICK_ATTN_QWARPS=19 selects the 19-warp forward / dQ kernels (96 registers per thread; measured 1.4% slower per step than 15 warps x 128 registers)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ICK_ATTN_QWARPS");
        v = (e && e[0] == '1' && e[1] == '9') ? 1 : 0;
    }
    return v != 0;
}
// dQ-from-dS scheme: ICK_ATTN_DS=0 never, =1 whenever a workspace is given, unset = where it measured faster on B200: long
// non-causal self-attention (the 301-slot entity / fact encoders, -8% per backward); for the 102-query decoder attentions the
// extra dS round trip through HBM costs as much as the recomputation it saves.
bool use_ds(int Sq, int Sk, int causal) {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ICK_ATTN_DS");
        v = (e && e[0] == '0') ? 0 : (e && e[0] == '1') ? 1 : 2;
    }
    if (v == 2) return !causal && Sq >= 256;
    return v != 0;
}
// ICK_ATTN_BWD = tc (default: everything on tcgen05 / TMEM, attention_bwd_tc.cu) | hybrid (mma.sync + tcgen05 dQ, attention_bwd_fused.cu) |
// split (the two-kernel backward: dQ and dK/dV each recompute the probabilities).  Each falls through to the next when a shape does not fit.
int bwd_mode() {  // (read on every call: the tests switch modes inside one process)
    const char* e = getenv("ICK_ATTN_BWD");
    int v = (e && e[0] == 's') ? 0 : (e && e[0] == 'h') ? 1 : 2;
    const char* f = getenv("ICK_ATTN_FUSED");
    if (f && f[0] == '0') v = 0;
    return v;
}
bool use_fused() { return bwd_mode() >= 1; }
// ICK_ATTN_FWD=tc selects the tcgen05 / TMEM forward (attention_fwd_tc.cu).  The default stays the mma.sync forward of this file: measured
// on B200 inside the train step 5.63 ms/step against 5.91 with the tcgen05 forward - two softmax warpgroups per SM (the TMEM budget of
// this formulation) hide less latency than the 15 compute warps here; the tcgen05 kernel wins when the keep words need many bit planes.
bool fwd_tc() {
    const char* e = getenv("ICK_ATTN_FWD");
    return e && e[0] == 't';
}
bool use_persistent() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ICK_ATTN_PERSISTENT");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

template <typename K>
int set_smem(K kernel, bool persistent = false) {
    static bool done = false;  // one static per kernel
    if (!done) {
        const int bytes = persistent ? SMEM_MAX : smem_bytes(CH, true);
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
            ick_set_error("attention: cudaFuncSetAttribute(MaxDynamicSharedMemorySize=%d) failed", bytes);
            return ICK_ERR_CUDA;
        }
        done = true;
    }
    return ICK_OK;
}

}  // namespace

int ick_mha_fwd_mma(const void* Q, const void* K, const void* V, void* O, float* lse, int B, int H, int Sq, int Sk, int dh, int ldq, int ldk,
                    int ldv, int ldo, int causal, DropCfg dc, cudaStream_t stream) {
    ICK_REQUIRE(((uintptr_t)K & 15) == 0 && ((uintptr_t)V & 15) == 0 && ((uintptr_t)Q & 3) == 0, "mha_fwd: operands must be 16-byte aligned");
    int rc;
    if (fwd_tc()) {
        rc = ick_mha_fwd_tc(Q, K, V, O, lse, B, H, Sq, Sk, dh, ldq, ldk, ldv, ldo, causal, dc, stream);
        if (rc != ICK_ERR_UNSUPPORTED) return rc;
    }
    Dims d = make_dims(B, H, Sq, Sk, dh, causal);
    CUtensorMap tmK, tmV;
    if ((rc = make_tmap3(&tmK, K, H, Sk, B, ldk))) return rc;
    if ((rc = make_tmap3(&tmV, V, H, Sk, B, ldv))) return rc;
    const int nt = (Sk + TK - 1) / TK, ntc = nt < CH ? nt : CH;
    PArgs pa;
    pa.d = d;
    pa.nslabs = (Sq + 15) / 16;
    plan_pipe(nt, false, &pa);
    if (pa.nstage && use_persistent()) {
        const int grid = B * H < num_sms() ? B * H : num_sms();
        if (wide_q()) {
            if ((rc = set_smem(fwd_pkernel<PNW_Q>, true))) return rc;
            ick_launch(fwd_pkernel<PNW_Q>, grid, 32 * (PNW_Q + 1), pipe_smem(pa), stream)(tmK, tmV, (const bf16*)Q, (bf16*)O, lse, pa, ldq, ldo, dc);
        } else {
            if ((rc = set_smem(fwd_pkernel<PNW_KV>, true))) return rc;
            ick_launch(fwd_pkernel<PNW_KV>, grid, 32 * (PNW_KV + 1), pipe_smem(pa), stream)(tmK, tmV, (const bf16*)Q, (bf16*)O, lse, pa, ldq, ldo, dc);
        }
        return ick_check_launch("mha_fwd_mma(persistent)");
    }
    if ((rc = set_smem(fwd_kernel))) return rc;
    int nctas, nw;
    split_own(Sq, &nctas, &nw);
    dim3 grid(nctas, H, B);
    ick_launch(fwd_kernel, grid, 32 * nw, smem_bytes(ntc, false), stream)(tmK, tmV, (const bf16*)Q, (bf16*)O, lse, d, ldq, ldo, ntc, dc);
    return ick_check_launch("mha_fwd_mma");
}

int ick_mha_bwd_mma(const void* Q, const void* K, const void* V, const void* O, const void* dO, const float* lse, float* dsum, void* dQ,
                    void* dK, void* dV, int B, int H, int Sq, int Sk, int dh, int ldq, int ldk, int ldv, int ldo, int lddo, int lddq, int lddk,
                    int lddv, int causal, DropCfg dc, void* workspace, long long workspace_bytes, int dsum_ready, cudaStream_t stream) {
    ICK_REQUIRE(((uintptr_t)K & 15) == 0 && ((uintptr_t)V & 15) == 0 && ((uintptr_t)Q & 15) == 0 && ((uintptr_t)dO & 15) == 0 &&
                    ((uintptr_t)O & 15) == 0,
                "mha_bwd: operands must be 16-byte aligned");
    int rc;
    if (bwd_mode() == 2) {
        rc = ick_mha_bwd_tc(Q, K, V, O, dO, lse, dsum, dQ, dK, dV, B, H, Sq, Sk, dh, ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv, causal, dc, dsum_ready, stream);
        if (rc != ICK_ERR_UNSUPPORTED) return rc;
    }
    if (use_fused()) {
        rc = ick_mha_bwd_fused(Q, K, V, O, dO, lse, dsum, dQ, dK, dV, B, H, Sq, Sk, dh, ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv, causal, dc, stream);
        if (rc != ICK_ERR_UNSUPPORTED) return rc;
    }
    Dims d = make_dims(B, H, Sq, Sk, dh, causal);
    CUtensorMap tmK, tmV, tmQ, tmG;
    if ((rc = make_tmap3(&tmK, K, H, Sk, B, ldk))) return rc;
    if ((rc = make_tmap3(&tmV, V, H, Sk, B, ldv))) return rc;
    if ((rc = make_tmap3(&tmQ, Q, H, Sq, B, ldq))) return rc;
    if ((rc = make_tmap3(&tmG, dO, H, Sq, B, lddo))) return rc;
    const int grid = B * H < num_sms() ? B * H : num_sms();
    int nctas, nw;
    int nt = (Sk + TK - 1) / TK, ntc = nt < CH ? nt : CH;
    PArgs pa;
    pa.d = d;
    // dS^T workspace: [image, head][16-key slab][32-query block][1 KiB]; with it dQ is a GEMM over the stored dS (see above)
    const int ds_nslabs = (Sk + 15) / 16, ds_nsub = 2 * ((Sq + TK - 1) / TK);
    const long long ds_need = (long long)B * H * ds_nslabs * ds_nsub * 1024;
    uint4* DS = (workspace != nullptr && workspace_bytes >= ds_need && (((uintptr_t)workspace) & 15) == 0 && use_ds(Sq, Sk, causal) &&
                 (size_t)ds_nslabs * 16 * DQ_KLD * 2 <= 128 * 1024 && (ldk % 8) == 0)
                    ? (uint4*)workspace
                    : nullptr;
    if (DS != nullptr) {
        const long long n = (long long)B * Sq * H;
        ick_launch(rowdot_kernel, (int)((n + 255) / 256), 256, 0, stream)((const bf16*)O, (const bf16*)dO, dsum, B, H, Sq, dh, ldo, lddo);
        if ((rc = ick_check_launch("mha_bwd_mma(rowdot)"))) return rc;
        nt = (Sq + TK - 1) / TK;
        ntc = nt < CH ? nt : CH;
        pa.nslabs = ds_nslabs;
        plan_pipe(nt, true, &pa);
        if (pa.nstage && use_persistent()) {
            if ((rc = set_smem(bwd_dkv_pkernel<PNW_KV>, true))) return rc;
            ick_launch(bwd_dkv_pkernel<PNW_KV>, grid, 32 * (PNW_KV + 1), pipe_smem(pa), stream)(tmQ, tmG, (const bf16*)K, (const bf16*)V, lse, dsum, (bf16*)dK,
                                                                                            (bf16*)dV, pa, ldk, ldv, lddk, lddv, dc, DS);
        } else {
            if ((rc = set_smem(bwd_dkv_kernel))) return rc;
            split_own(Sk, &nctas, &nw);
            ick_launch(bwd_dkv_kernel, dim3(nctas, H, B), 32 * nw, smem_bytes(ntc, true), stream)(tmQ, tmG, (const bf16*)K, (const bf16*)V, lse, dsum,
                                                                                              (bf16*)dK, (bf16*)dV, d, ldk, ldv, lddk, lddv, ntc, dc, DS);
        }
        if ((rc = ick_check_launch("mha_bwd_mma(dkv + dS)"))) return rc;
        if ((rc = set_smem(bwd_dq_ds_kernel, true))) return rc;
        int ksplit = 1, warps = DQ_MAXW;
        if (ds_nsub <= DQ_MAXW) {
            ksplit = DQ_MAXW / ds_nsub;
            if (ksplit > 4) ksplit = 4;
            if (ksplit > (ds_nslabs + 1) / 2) ksplit = (ds_nslabs + 1) / 2;  // at least two slabs per part
            if (ksplit < 1) ksplit = 1;
            warps = ds_nsub * ksplit;
        }
        const size_t dq_smem_bytes = (size_t)ds_nslabs * 16 * DQ_KLD * 2 + (ksplit > 1 ? (size_t)warps * 4096 : 0);
        ick_launch(bwd_dq_ds_kernel, B * H, 32 * warps, dq_smem_bytes, stream)(DS, (const bf16*)K, (bf16*)dQ, d, ldk, lddq, ds_nslabs, ds_nsub,
                                                                              ksplit);
        return ick_check_launch("mha_bwd_mma(dq from dS)");
    }
    pa.nslabs = (Sq + 15) / 16;
    plan_pipe(nt, false, &pa);
    if (pa.nstage && use_persistent()) {
        if (wide_q()) {
            if ((rc = set_smem(bwd_dq_pkernel<PNW_Q>, true))) return rc;
            ick_launch(bwd_dq_pkernel<PNW_Q>, grid, 32 * (PNW_Q + 1), pipe_smem(pa), stream)(tmK, tmV, (const bf16*)Q, (const bf16*)O, (const bf16*)dO,
                                                                                         lse, dsum, (bf16*)dQ, pa, ldq, ldo, lddo, lddq, dc);
        } else {
            if ((rc = set_smem(bwd_dq_pkernel<PNW_KV>, true))) return rc;
            ick_launch(bwd_dq_pkernel<PNW_KV>, grid, 32 * (PNW_KV + 1), pipe_smem(pa), stream)(tmK, tmV, (const bf16*)Q, (const bf16*)O, (const bf16*)dO,
                                                                                           lse, dsum, (bf16*)dQ, pa, ldq, ldo, lddo, lddq, dc);
        }
    } else {
        if ((rc = set_smem(bwd_dq_kernel))) return rc;
        split_own(Sq, &nctas, &nw);
        ick_launch(bwd_dq_kernel, dim3(nctas, H, B), 32 * nw, smem_bytes(ntc, false), stream)(tmK, tmV, (const bf16*)Q, (const bf16*)O, (const bf16*)dO, lse,
                                                                                 dsum, (bf16*)dQ, d, ldq, ldo, lddo, lddq, ntc, dc);
    }
    if ((rc = ick_check_launch("mha_bwd_mma(dq)"))) return rc;
    nt = (Sq + TK - 1) / TK;
    ntc = nt < CH ? nt : CH;
    pa.nslabs = (Sk + 15) / 16;
    plan_pipe(nt, true, &pa);
    if (pa.nstage && use_persistent()) {
        if ((rc = set_smem(bwd_dkv_pkernel<PNW_KV>, true))) return rc;
        ick_launch(bwd_dkv_pkernel<PNW_KV>, grid, 32 * (PNW_KV + 1), pipe_smem(pa), stream)(tmQ, tmG, (const bf16*)K, (const bf16*)V, lse, dsum, (bf16*)dK, (bf16*)dV,
                                                                         pa, ldk, ldv, lddk, lddv, dc, (uint4*)nullptr);
        return ick_check_launch("mha_bwd_mma(dkv, persistent)");
    }
    if ((rc = set_smem(bwd_dkv_kernel))) return rc;
    split_own(Sk, &nctas, &nw);
    ick_launch(bwd_dkv_kernel, dim3(nctas, H, B), 32 * nw, smem_bytes(ntc, true), stream)(tmQ, tmG, (const bf16*)K, (const bf16*)V, lse, dsum, (bf16*)dK,
                                                                                 (bf16*)dV, d, ldk, ldv, lddk, lddv, ntc, dc, (uint4*)nullptr);
    return ick_check_launch("mha_bwd_mma(dkv)");
}
