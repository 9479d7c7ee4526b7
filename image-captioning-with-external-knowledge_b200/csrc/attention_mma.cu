// bf16 flash attention for head_dim <= 32 (10 heads x 30 in the reference): warp-level mma.sync m16n8k16 with fp32
// accumulation, the streamed operand staged by TMA.  Same contract as the CUDA-core kernels in attention.cu (which stay the
// fp32-parity path): head-layout rows, softmax in the exp2 domain, attention-probability dropout regenerated from the
// counter hash, log2-domain LSE saved.
//
//   ownership : a warp owns 16 "own" rows (queries in fwd / dQ, keys in dK-dV) whose A fragments are read straight from
//               global memory once; a CTA is 1..8 such warps.
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

namespace {

constexpr int HD = 32;                   // padded head width: 64-byte rows
constexpr int TK = 64;                   // streamed rows per tile
constexpr int TILE_BYTES = TK * HD * 2;  // 4096
constexpr int CH = 10;                   // tiles resident at once
constexpr int NWMAX = 8;                 // warps per CTA
constexpr int SUB = 32;                  // streamed rows per register sub-step of the backward kernels

struct Dims {
    int B, H, Sq, Sk, dh;
    int causal;
    float scale, scale_log2;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ---- shared-memory carve-up: [mbarriers (1 KiB)] [tensor 0: n tiles] [tensor 1: n tiles] [per-row scalars (dKV only)] ---------
struct Smem {
    uint32_t bars, t0, t1;
    float* scal;  // generic pointer to the scalar area
};
__device__ __forceinline__ Smem carve(uint8_t* raw, int ntc) {
    uint8_t* p = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    Smem s;
    s.bars = smem_u32(p);
    s.t0 = s.bars + 1024;
    s.t1 = s.t0 + ntc * TILE_BYTES;
    s.scal = (float*)(p + 1024 + 2 * ntc * TILE_BYTES);
    return s;
}
__device__ __forceinline__ void init_bars(const Smem& s, int ntc, const CUtensorMap* a, const CUtensorMap* b) {
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(b) : "memory");
        for (int t = 0; t < ntc; ++t) mbar_init(s.bars + 8 * t, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}
// one elected thread: tiles [tile0, tile0 + n) of both streamed tensors of image b, head h
__device__ __forceinline__ void issue_tiles(const Smem& s, const CUtensorMap* a, const CUtensorMap* b2, int tile0, int n, int h, int b) {
    for (int t = 0; t < n; ++t) {
        const uint32_t bar = s.bars + 8 * t;
        mbar_expect_tx(bar, 2 * TILE_BYTES);
        tma_load_3d(s.t0 + t * TILE_BYTES, a, bar, h * HD, (tile0 + t) * TK, b);
        tma_load_3d(s.t1 + t * TILE_BYTES, b2, bar, h * HD, (tile0 + t) * TK, b);
    }
}

// Byte offset of (row r, 16-byte chunk c) inside a 64B-swizzled tile whose base is 1024-aligned: r*64 + ((c ^ ((r>>1)&3)) << 4).
// Per-lane ldmatrix offsets; rows advance in multiples of 8 (non-trans) / 16 (trans), which leaves the swizzle term unchanged.
struct LaneOff {
    uint32_t nt;     // "tile^T" B operand: matrix m = lane>>3 is chunk m of row (lane&7)
    uint32_t tr[2];  // "P x tile" B operand (.trans): row (lane&7) + 8*((lane>>3)&1), chunk 2*nd + (lane>>4)
};
__device__ __forceinline__ LaneOff lane_offsets(int lane) {
    LaneOff o;
    const int r = lane & 7, c = lane >> 3;
    o.nt = r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
    const int r2 = (lane & 7) + 8 * ((lane >> 3) & 1), hi = lane >> 4;
#pragma unroll
    for (int nd = 0; nd < 2; ++nd) o.tr[nd] = r2 * 64 + (((2 * nd + hi) ^ ((r2 >> 1) & 3)) << 4);
    return o;
}

// A-operand fragments (16 rows x 32 k) of the warp's own rows, read from global memory; rows >= rmax and columns >= dh are zero
__device__ __forceinline__ uint32_t ld_frag(const bf16* row, bool ok, int c, int dh) {
    uint32_t v = ok ? *reinterpret_cast<const uint32_t*>(row + c) : 0u;
    if (c + 1 >= dh) v = c >= dh ? 0u : (v & 0xFFFFu);
    return v;
}
__device__ __forceinline__ void load_own(const bf16* base, size_t ld, int r0, int r1, int rmax, int dh, int tq, uint32_t (*a)[4]) {
    const bf16* p0 = base + (size_t)r0 * ld;
    const bf16* p1 = base + (size_t)r1 * ld;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        a[ks][0] = ld_frag(p0, r0 < rmax, 16 * ks + 2 * tq, dh);
        a[ks][1] = ld_frag(p1, r1 < rmax, 16 * ks + 2 * tq, dh);
        a[ks][2] = ld_frag(p0, r0 < rmax, 16 * ks + 2 * tq + 8, dh);
        a[ks][3] = ld_frag(p1, r1 < rmax, 16 * ks + 2 * tq + 8, dh);
    }
}

// acc[j] (16 x 8, j = 0..NJ-1) = A(16 x 32) * T^T for rows [8*j0, 8*(j0+NJ)) of the tile at shared address `tile`
template <int NJ>
__device__ __forceinline__ void mma_a_tT(float (*acc)[4], const uint32_t (*a)[4], uint32_t tile, int j0, const LaneOff& lo) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(b0, b1, b2, b3, tile + (j0 + j) * 512 + lo.nt);
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        mma16816(acc[j], a[0], b0, b1);
        mma16816(acc[j], a[1], b2, b3);
    }
}
// out[n] (16 x 8 over d, n = 0..3) += P(16 x 8*NJ) * T[rows 8*j0 ..], P given as packed A fragments pa[NJ/2][4]
template <int NJ>
__device__ __forceinline__ void mma_p_t(float (*out)[4], const uint32_t (*pa)[4], uint32_t tile, int j0, const LaneOff& lo) {
#pragma unroll
    for (int kk = 0; kk < NJ / 2; ++kk) {
#pragma unroll
        for (int nd = 0; nd < 2; ++nd) {
            uint32_t r0, r1, r2, r3;
            ldsm_x4_trans(r0, r1, r2, r3, tile + (j0 / 2 + kk) * 1024 + lo.tr[nd]);
            mma16816(out[2 * nd], pa[kk], r0, r1);
            mma16816(out[2 * nd + 1], pa[kk], r2, r3);
        }
    }
}
// accumulator layout p[NJ][4] -> packed A fragments
template <int NJ>
__device__ __forceinline__ void pack_p(const float (*p)[4], uint32_t (*pa)[4]) {
#pragma unroll
    for (int kk = 0; kk < NJ / 2; ++kk) {
        pa[kk][0] = pack2(p[2 * kk][0], p[2 * kk][1]);
        pa[kk][1] = pack2(p[2 * kk][2], p[2 * kk][3]);
        pa[kk][2] = pack2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
        pa[kk][3] = pack2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
    }
}

// write a 16 x 32 accumulator slab (4 n-tiles) as bf16 rows; columns >= dh are written as zero
__device__ __forceinline__ void store_slab(bf16* base, size_t ld, int row_g, int row_g8, int rmax, const float (*o)[4], float s0, float s1,
                                           int dh, int tq) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const int c = 8 * n + 2 * tq;
        const float m0 = c < dh ? 1.f : 0.f, m1 = c + 1 < dh ? 1.f : 0.f;
        if (row_g < rmax) *reinterpret_cast<uint32_t*>(base + (size_t)row_g * ld + c) = pack2(o[n][0] * s0 * m0, o[n][1] * s0 * m1);
        if (row_g8 < rmax) *reinterpret_cast<uint32_t*>(base + (size_t)row_g8 * ld + c) = pack2(o[n][2] * s1 * m0, o[n][3] * s1 * m1);
    }
}

// ---- dropout of the probabilities: element (row = (b,h,query), col = key) -------------------------------------------------
__device__ __forceinline__ uint64_t prob_row(int b, int H, int h, int Sq, int qi) { return ((uint64_t)b * H + h) * (uint64_t)Sq + qi; }
// both 15-bit fields of a pair hash against thr with one add: bit 15 / bit 31 of the result = keep(even col) / keep(odd col)
__device__ __forceinline__ uint32_t keep_bits(uint32_t pairhash, uint32_t addc) { return (pairhash & 0x7FFF7FFFu) + addc; }
__device__ __forceinline__ uint32_t keep_addc(uint32_t thr) { return (0x8000u - thr) * 0x00010001u; }
// 0xFFFF / 0x0000 per half-word from the two keep bits (byte permute with sign replication)
__device__ __forceinline__ uint32_t keep_mask_bf16x2(uint32_t kb) {
    uint32_t m;
    asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(m) : "r"(kb));
    return m;
}

// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * NWMAX, 2)
    fwd_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const bf16* __restrict__ Q,
               bf16* __restrict__ O, float* __restrict__ LSE, Dims d, int ldq, int ldo, int ntc, DropCfg drop) {
    extern __shared__ uint8_t smem_raw[];
    const Smem sm = carve(smem_raw, ntc);
    ick_resolve_seed(drop);
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * nw * 16;
    const int kend = d.causal ? min(d.Sk, q0 + nw * 16) : d.Sk;
    const int nt = (kend + TK - 1) / TK;
    init_bars(sm, ntc, &tmK, &tmV);
    if (threadIdx.x == 0) issue_tiles(sm, &tmK, &tmV, 0, min(nt, ntc), h, b);

    const int qi0 = q0 + 16 * warp + g, qi1 = qi0 + 8;
    const bool active = q0 + 16 * warp < d.Sq;
    uint32_t qa[2][4];
    load_own(Q + (size_t)b * d.Sq * ldq + h * HD, ldq, qi0, qi1, d.Sq, d.dh, tq, qa);
    const LaneOff lo = lane_offsets(lane);
    const uint32_t rm0 = ick_rowmix(drop.seed, drop.site, prob_row(b, d.H, h, d.Sq, qi0));
    const uint32_t rm1 = ick_rowmix(drop.seed, drop.site, prob_row(b, d.H, h, d.Sq, qi1));
    const uint32_t addc = keep_addc(drop.thr);
    const float c = d.scale_log2;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // running max of the RAW scores, running sum
    float o[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;

    for (int c0 = 0; c0 < nt; c0 += ntc) {
        const int n = min(ntc, nt - c0);
        if (c0 > 0) {
            __syncthreads();  // every warp is done with the resident tiles
            if (threadIdx.x == 0) issue_tiles(sm, &tmK, &tmV, c0, n, h, b);
        }
        const uint32_t parity = (uint32_t)(c0 / ntc) & 1u;
        for (int t = 0; t < n; ++t) {
            mbar_wait(sm.bars + 8 * t, parity);
            if (!active) continue;
            const int k0 = (c0 + t) * TK;
            // causal: tiles entirely beyond this warp's last query contribute nothing
            if (d.causal && k0 > q0 + 16 * warp + 15) continue;
            float s[8][4];
            mma_a_tT<8>(s, qa, sm.t0 + t * TILE_BYTES, 0, lo);
            // masking is needed only on the ragged last tile and on tiles that cross this warp's causal diagonal
            if (k0 + TK > d.Sk || (d.causal && k0 + TK - 1 > q0 + 16 * warp)) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int key = k0 + 8 * j + 2 * tq + (e & 1);
                        const bool vis = key < d.Sk && (!d.causal || key <= (e < 2 ? qi0 : qi1));
                        if (!vis) s[j][e] = -INFINITY;
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
            const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
            // a row with no visible key so far keeps m = -inf; use 0 as the exponent base there (all p are ex2(-inf) = 0)
            const float e0 = mn0 == -INFINITY ? 0.f : mn0 * c, e1 = mn1 == -INFINITY ? 0.f : mn1 * c;
            const float c0f = ex2(m0 * c - e0), c1f = ex2(m1 * c - e1);
            l0 *= c0f;
            l1 *= c1f;
#pragma unroll
            for (int nn = 0; nn < 4; ++nn) {
                o[nn][0] *= c0f; o[nn][1] *= c0f; o[nn][2] *= c1f; o[nn][3] *= c1f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j][0] = ex2(fmaf(s[j][0], c, -e0));
                s[j][1] = ex2(fmaf(s[j][1], c, -e0));
                s[j][2] = ex2(fmaf(s[j][2], c, -e1));
                s[j][3] = ex2(fmaf(s[j][3], c, -e1));
                l0 += s[j][0] + s[j][1];
                l1 += s[j][2] + s[j][3];
            }
            uint32_t pa[4][4];
            pack_p<8>(s, pa);
            if (drop.thr != 0u) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const uint32_t kc = (uint32_t)(k0 + 8 * (2 * kk + jj) + 2 * tq);
                        pa[kk][2 * jj] &= keep_mask_bf16x2(keep_bits(ick_pairhash(rm0, kc), addc));
                        pa[kk][2 * jj + 1] &= keep_mask_bf16x2(keep_bits(ick_pairhash(rm1, kc), addc));
                    }
            }
            mma_p_t<8>(o, pa, sm.t1 + t * TILE_BYTES, 0, lo);
            m0 = mn0;
            m1 = mn1;
        }
    }
    if (!active) return;
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    bf16* Ob = O + (size_t)b * d.Sq * ldo + h * HD;
    store_slab(Ob, ldo, qi0, qi1, d.Sq, o, drop.inv_keep / l0, drop.inv_keep / l1, d.dh, tq);
    if (tq == 0) {
        float* L = LSE + ((size_t)b * d.H + h) * d.Sq;
        if (qi0 < d.Sq) L[qi0] = m0 * c + log2f(l0);
        if (qi1 < d.Sq) L[qi1] = m1 * c + log2f(l1);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * NWMAX, 2)
    bwd_dq_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const bf16* __restrict__ Q,
                  const bf16* __restrict__ O, const bf16* __restrict__ dO, const float* __restrict__ LSE, float* __restrict__ Dsum,
                  bf16* __restrict__ dQ, Dims d, int ldq, int ldo, int lddo, int lddq, int ntc, DropCfg drop) {
    extern __shared__ uint8_t smem_raw[];
    const Smem sm = carve(smem_raw, ntc);
    ick_resolve_seed(drop);
    const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * nw * 16;
    const int kend = d.causal ? min(d.Sk, q0 + nw * 16) : d.Sk;
    const int nt = (kend + TK - 1) / TK;
    init_bars(sm, ntc, &tmK, &tmV);
    if (threadIdx.x == 0) issue_tiles(sm, &tmK, &tmV, 0, min(nt, ntc), h, b);

    const int qi0 = q0 + 16 * warp + g, qi1 = qi0 + 8;
    const bool active = q0 + 16 * warp < d.Sq;
    const bf16* Gb = dO + (size_t)b * d.Sq * lddo + h * HD;
    uint32_t qa[2][4], ga[2][4];
    load_own(Q + (size_t)b * d.Sq * ldq + h * HD, ldq, qi0, qi1, d.Sq, d.dh, tq, qa);
    load_own(Gb, lddo, qi0, qi1, d.Sq, d.dh, tq, ga);
    // D = rowsum(dO * O): lane pair (2r, 2r+1) handles the two 16-column halves of own row r
    float D0, D1;
    {
        const int r = q0 + 16 * warp + (lane >> 1), cb = (lane & 1) * 16;
        float acc = 0.f;
        if (r < d.Sq) {
            const bf16* op = O + ((size_t)b * d.Sq + r) * ldo + h * HD + cb;
            const bf16* gp = Gb + (size_t)r * lddo + cb;
            float x[8], y[8];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                ld8(op + 8 * v, x);
                ld8(gp + 8 * v, y);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (cb + 8 * v + i < d.dh) acc = fmaf(x[i], y[i], acc);
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if ((lane & 1) == 0 && r < d.Sq) Dsum[((size_t)b * d.H + h) * d.Sq + r] = acc;
        D0 = __shfl_sync(0xffffffffu, acc, 2 * g);
        D1 = __shfl_sync(0xffffffffu, acc, 2 * g + 16);
    }
    const LaneOff lo = lane_offsets(lane);
    const uint32_t rm0 = ick_rowmix(drop.seed, drop.site, prob_row(b, d.H, h, d.Sq, qi0));
    const uint32_t rm1 = ick_rowmix(drop.seed, drop.site, prob_row(b, d.H, h, d.Sq, qi1));
    const uint32_t addc = keep_addc(drop.thr);
    const float* L = LSE + ((size_t)b * d.H + h) * d.Sq;
    const float lse0 = qi0 < d.Sq ? L[qi0] : 0.f, lse1 = qi1 < d.Sq ? L[qi1] : 0.f;
    const float c = d.scale_log2, ik = drop.inv_keep;
    float dq[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;

    for (int c0 = 0; c0 < nt; c0 += ntc) {
        const int n = min(ntc, nt - c0);
        if (c0 > 0) {
            __syncthreads();
            if (threadIdx.x == 0) issue_tiles(sm, &tmK, &tmV, c0, n, h, b);
        }
        const uint32_t parity = (uint32_t)(c0 / ntc) & 1u;
        for (int t = 0; t < n; ++t) {
            mbar_wait(sm.bars + 8 * t, parity);
            if (!active) continue;
            const uint32_t kt = sm.t0 + t * TILE_BYTES, vt = sm.t1 + t * TILE_BYTES;
#pragma unroll 1
            for (int sub = 0; sub < TK / SUB; ++sub) {
                const int k0 = (c0 + t) * TK + sub * SUB;
                if (d.causal && k0 > q0 + 16 * warp + 15) break;
                // keys past Sk are zero rows (TMA fill): they add nothing to dQ = dS K, so only the causal diagonal needs a mask
                const bool need_mask = d.causal && k0 + SUB - 1 > q0 + 16 * warp;
                float s[4][4], dp[4][4];
                mma_a_tT<4>(s, qa, kt, 4 * sub, lo);
                mma_a_tT<4>(dp, ga, vt, 4 * sub, lo);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (drop.thr != 0u) {
                        const uint32_t kc = (uint32_t)(k0 + 8 * j + 2 * tq);
                        const uint32_t kb0 = keep_bits(ick_pairhash(rm0, kc), addc), kb1 = keep_bits(ick_pairhash(rm1, kc), addc);
                        dp[j][0] = (kb0 & 0x8000u) ? dp[j][0] * ik : 0.f;
                        dp[j][1] = (kb0 & 0x80000000u) ? dp[j][1] * ik : 0.f;
                        dp[j][2] = (kb1 & 0x8000u) ? dp[j][2] * ik : 0.f;
                        dp[j][3] = (kb1 & 0x80000000u) ? dp[j][3] * ik : 0.f;
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float p = ex2(fmaf(s[j][e], c, -(e < 2 ? lse0 : lse1)));
                        if (need_mask && k0 + 8 * j + 2 * tq + (e & 1) > (e < 2 ? qi0 : qi1)) p = 0.f;
                        s[j][e] = p * (dp[j][e] - (e < 2 ? D0 : D1));
                    }
                }
                uint32_t pa[2][4];
                pack_p<4>(s, pa);
                mma_p_t<4>(dq, pa, kt, 4 * sub, lo);
            }
        }
    }
    if (!active) return;
    bf16* dQb = dQ + (size_t)b * d.Sq * lddq + h * HD;
    store_slab(dQb, lddq, qi0, qi1, d.Sq, dq, d.scale, d.scale, d.dh, tq);
}

// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * NWMAX, 2)
    bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG, const bf16* __restrict__ K,
                   const bf16* __restrict__ V, const float* __restrict__ LSE, const float* __restrict__ Dsum, bf16* __restrict__ dK,
                   bf16* __restrict__ dV, Dims d, int ldk, int ldv, int lddk, int lddv, int ntc, DropCfg drop) {
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

    const int kj0 = j0 + 16 * warp + g, kj1 = kj0 + 8;
    const bool active = j0 + 16 * warp < d.Sk;
    uint32_t ka[2][4], va[2][4];
    load_own(K + (size_t)b * d.Sk * ldk + h * HD, ldk, kj0, kj1, d.Sk, d.dh, tq, ka);
    load_own(V + (size_t)b * d.Sk * ldv + h * HD, ldv, kj0, kj1, d.Sk, d.dh, tq, va);
    const LaneOff lo = lane_offsets(lane);
    const float* L = LSE + ((size_t)b * d.H + h) * d.Sq;
    const float* Dg = Dsum + ((size_t)b * d.H + h) * d.Sq;
    // Lanes g and g^1 own keys of the same dropout pair and see the same queries: each computes the pair hashes of ONE of its
    // two key rows (even g: kj0, odd g: kj1) and receives the other from its partner (lane ^ 4).
    const bool odd = (g & 1) != 0;
    const uint32_t mykey = (uint32_t)(odd ? kj1 : kj0);
    const uint32_t bit0 = odd ? 0x80000000u : 0x8000u;  // keep bit of this lane's key parity
    const uint32_t addc = keep_addc(drop.thr);
    const float c = d.scale_log2, ik = drop.inv_keep;
    float dk[4][4], dv[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
        dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
    }
    for (int c0 = tbeg; c0 < nt; c0 += ntc) {
        const int n = min(ntc, nt - c0);
        if (c0 > tbeg) {
            __syncthreads();
            if (threadIdx.x == 0) issue_tiles(sm, &tmQ, &tmG, c0, n, h, b);
        }
        for (int i = threadIdx.x; i < n * TK; i += blockDim.x) {
            const int qi = c0 * TK + i;
            const bool ok = qi < d.Sq;
            Ls[i] = ok ? L[qi] : 0.f;
            Ds[i] = ok ? Dg[qi] : 0.f;
            Rm[i] = ick_rowmix(drop.seed, drop.site, prob_row(b, d.H, h, d.Sq, qi));
        }
        __syncthreads();
        const uint32_t parity = (uint32_t)((c0 - tbeg) / ntc) & 1u;
        for (int t = 0; t < n; ++t) {
            mbar_wait(sm.bars + 8 * t, parity);
            if (!active) continue;
            const uint32_t qt = sm.t0 + t * TILE_BYTES, gt = sm.t1 + t * TILE_BYTES;
#pragma unroll 1
            for (int sub = 0; sub < TK / SUB; ++sub) {
                const int q0 = (c0 + t) * TK + sub * SUB;  // first query of the sub-step
                if (d.causal && q0 + SUB - 1 < j0 + 16 * warp) continue;  // every query precedes every key of this warp
                // queries past Sq are zero rows of Q and dO (TMA fill) and add nothing; own keys past Sk are never stored
                const bool need_mask = d.causal && j0 + 16 * warp + 15 > q0;
                const float* ls = Ls + t * TK + sub * SUB + 2 * tq;
                const float* ds = Ds + t * TK + sub * SUB + 2 * tq;
                const uint32_t* rm = Rm + t * TK + sub * SUB + 2 * tq;
                float st[4][4], dpt[4][4];
                mma_a_tT<4>(st, ka, qt, 4 * sub, lo);
                mma_a_tT<4>(dpt, va, gt, 4 * sub, lo);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 l2 = *reinterpret_cast<const float2*>(ls + 8 * j);
                    const float2 d2 = *reinterpret_cast<const float2*>(ds + 8 * j);
                    uint32_t hq0a = 0, hq0b = 0, hq1a = 0, hq1b = 0;  // keep bits of (query 0/1 of the pair, key row a = kj0 / b = kj1)
                    if (drop.thr != 0u) {
                        const uint2 r2 = *reinterpret_cast<const uint2*>(rm + 8 * j);
                        const uint32_t m0 = keep_bits(ick_pairhash(r2.x, mykey), addc), m1 = keep_bits(ick_pairhash(r2.y, mykey), addc);
                        const uint32_t o0 = __shfl_xor_sync(0xffffffffu, m0, 4), o1 = __shfl_xor_sync(0xffffffffu, m1, 4);
                        hq0a = odd ? o0 : m0; hq0b = odd ? m0 : o0;
                        hq1a = odd ? o1 : m1; hq1b = odd ? m1 : o1;
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int qi = q0 + 8 * j + 2 * tq + (e & 1);
                        float p = ex2(fmaf(st[j][e], c, -((e & 1) ? l2.y : l2.x)));
                        if (need_mask && (e < 2 ? kj0 : kj1) > qi) p = 0.f;
                        float gp = dpt[j][e];
                        float pm = p;
                        if (drop.thr != 0u) {
                            const uint32_t kb = (e == 0) ? hq0a : (e == 1) ? hq1a : (e == 2) ? hq0b : hq1b;
                            const bool keep = (kb & bit0) != 0u;
                            pm = keep ? p : 0.f;
                            gp = keep ? gp * ik : 0.f;
                        }
                        st[j][e] = pm;                                   // P^T with dropout (1/keep applied at the store) -> dV
                        dpt[j][e] = p * (gp - ((e & 1) ? d2.y : d2.x));  // dS^T -> dK
                    }
                }
                uint32_t pa[2][4];
                pack_p<4>(st, pa);
                mma_p_t<4>(dv, pa, gt, 4 * sub, lo);
                pack_p<4>(dpt, pa);
                mma_p_t<4>(dk, pa, qt, 4 * sub, lo);
            }
        }
    }
    if (!active) return;
    bf16* dKb = dK + (size_t)b * d.Sk * lddk + h * HD;
    bf16* dVb = dV + (size_t)b * d.Sk * lddv + h * HD;
    store_slab(dKb, lddk, kj0, kj1, d.Sk, dk, d.scale, d.scale, d.dh, tq);
    store_slab(dVb, lddv, kj0, kj1, d.Sk, dv, ik, ik, d.dh, tq);
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
// 3-D bf16 map over a (B, S, ld) head-layout tensor: (column < H*32, position < S, image < B); box = 32 columns x 64 positions
int make_tmap3(CUtensorMap* tm, const void* ptr, int H, int S, int B, int ld) {
    EncodeFn enc = get_encode();
    if (!enc) {
        ick_set_error("cuTensorMapEncodeTiled entry point not available");
        return ICK_ERR_CUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)H * HD, (cuuint64_t)S, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)S * ld * 2};
    cuuint32_t box[3] = {HD, TK, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ick_set_error("attention: cuTensorMapEncodeTiled failed (%d): ptr=%p H=%d S=%d B=%d ld=%d", (int)r, ptr, H, S, B, ld);
        return ICK_ERR_CUDA;
    }
    return ICK_OK;
}

Dims make_dims(int B, int H, int Sq, int Sk, int dh, int causal) {
    Dims d;
    d.B = B; d.H = H; d.Sq = Sq; d.Sk = Sk; d.dh = dh;
    d.causal = causal;
    d.scale = 1.0f / sqrtf((float)dh);
    d.scale_log2 = d.scale * 1.4426950408889634f;
    return d;
}

// own rows -> (CTAs along the own dimension, warps per CTA): slabs of 16 rows spread evenly over the fewest CTAs
void split_own(int S, int* nctas, int* nw) {
    const int slabs = (S + 15) / 16;
    *nctas = (slabs + NWMAX - 1) / NWMAX;
    *nw = (slabs + *nctas - 1) / *nctas;
}
int smem_bytes(int ntc, bool scalars) { return 1024 + 1024 + 2 * ntc * TILE_BYTES + (scalars ? 3 * ntc * TK * 4 : 0); }

template <typename K>
int set_smem(K kernel) {
    static bool done = false;  // one static per kernel
    if (!done) {
        const int bytes = smem_bytes(CH, true);
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
    int rc = set_smem(fwd_kernel);
    if (rc) return rc;
    Dims d = make_dims(B, H, Sq, Sk, dh, causal);
    CUtensorMap tmK, tmV;
    if ((rc = make_tmap3(&tmK, K, H, Sk, B, ldk))) return rc;
    if ((rc = make_tmap3(&tmV, V, H, Sk, B, ldv))) return rc;
    int nctas, nw;
    split_own(Sq, &nctas, &nw);
    const int nt = (Sk + TK - 1) / TK, ntc = nt < CH ? nt : CH;
    dim3 grid(nctas, H, B);
    fwd_kernel<<<grid, 32 * nw, smem_bytes(ntc, false), stream>>>(tmK, tmV, (const bf16*)Q, (bf16*)O, lse, d, ldq, ldo, ntc, dc);
    return ick_check_launch("mha_fwd_mma");
}

int ick_mha_bwd_mma(const void* Q, const void* K, const void* V, const void* O, const void* dO, const float* lse, float* dsum, void* dQ,
                    void* dK, void* dV, int B, int H, int Sq, int Sk, int dh, int ldq, int ldk, int ldv, int ldo, int lddo, int lddq, int lddk,
                    int lddv, int causal, DropCfg dc, cudaStream_t stream) {
    ICK_REQUIRE(((uintptr_t)K & 15) == 0 && ((uintptr_t)V & 15) == 0 && ((uintptr_t)Q & 15) == 0 && ((uintptr_t)dO & 15) == 0 &&
                    ((uintptr_t)O & 15) == 0,
                "mha_bwd: operands must be 16-byte aligned");
    int rc = set_smem(bwd_dq_kernel);
    if (rc) return rc;
    if ((rc = set_smem(bwd_dkv_kernel))) return rc;
    Dims d = make_dims(B, H, Sq, Sk, dh, causal);
    CUtensorMap tmK, tmV, tmQ, tmG;
    if ((rc = make_tmap3(&tmK, K, H, Sk, B, ldk))) return rc;
    if ((rc = make_tmap3(&tmV, V, H, Sk, B, ldv))) return rc;
    if ((rc = make_tmap3(&tmQ, Q, H, Sq, B, ldq))) return rc;
    if ((rc = make_tmap3(&tmG, dO, H, Sq, B, lddo))) return rc;
    int nctas, nw;
    split_own(Sq, &nctas, &nw);
    int nt = (Sk + TK - 1) / TK, ntc = nt < CH ? nt : CH;
    bwd_dq_kernel<<<dim3(nctas, H, B), 32 * nw, smem_bytes(ntc, false), stream>>>(tmK, tmV, (const bf16*)Q, (const bf16*)O, (const bf16*)dO, lse,
                                                                                 dsum, (bf16*)dQ, d, ldq, ldo, lddo, lddq, ntc, dc);
    if ((rc = ick_check_launch("mha_bwd_mma(dq)"))) return rc;
    split_own(Sk, &nctas, &nw);
    nt = (Sq + TK - 1) / TK;
    ntc = nt < CH ? nt : CH;
    bwd_dkv_kernel<<<dim3(nctas, H, B), 32 * nw, smem_bytes(ntc, true), stream>>>(tmQ, tmG, (const bf16*)K, (const bf16*)V, lse, dsum, (bf16*)dK,
                                                                                 (bf16*)dV, d, ldk, ldv, lddk, lddv, ntc, dc);
    return ick_check_launch("mha_bwd_mma(dkv)");
}
