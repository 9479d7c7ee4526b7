// bf16 flash attention on the tensor cores (warp-level mma.sync m16n8k16, fp32 accumulate) for head_dim <= 32.
// Same contract as the CUDA-core kernels in attention.cu (which stay the fp32-parity path): head layout rows, online
// softmax in the exp2 domain, attention-probability dropout regenerated from the counter hash, log2-domain LSE saved.
//
//   fwd : one warp owns 16 queries; S = Q K^T (2 k-steps of 16 over d), P = exp2(S*c - m) re-used straight from the
//         accumulator registers as the A operand of O += P V (V fragments via ldmatrix.trans).
//   dQ  : same ownership; S and dP = dO V^T recomputed per key tile, dS = P (dP*mask - D), dQ += dS K.
//   dKV : one warp owns 16 keys; S^T = K Q^T and dP^T = V dO^T so that P^T / dS^T come out in A-operand layout for
//         dV += P^T dO and dK += dS^T Q.
// Shared-memory tiles are [64 rows][40 bf16] (80-byte rows): conflict-free for both the 32-bit fragment loads and
// ldmatrix.  Heads have only 30-32 useful columns, so these kernels are bound by exp2/FMA issue and tile loads, not by
// the tensor pipe; a tcgen05/TMEM version would not change that bound.
#include "common.cuh"
#include "attention_internal.h"

namespace {

constexpr int HD = 32;
constexpr int TQ = 64;   // rows (queries or keys) owned by a CTA: 4 warps x 16
constexpr int TK = 64;   // rows of the streamed operand per shared-memory tile
constexpr int LDS = 40;  // bf16 elements per shared-memory row (80 B)

struct Dims {
    int B, H, Sq, Sk, dh;
    int ldq, ldk, ldv, ldo;
    int causal;
    float scale, scale_log2;
};

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const bf16* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t lds32(const bf16* p) { return *reinterpret_cast<const uint32_t*>(p); }

// rows [r0, r0+64) x 32 columns of a (rows, ld) bf16 matrix -> smem tile; rows >= rmax and columns >= dh are zeroed
// (synchronous; used for the operand that becomes A fragments, whose pad lanes MUST be zero)
__device__ __forceinline__ void load_tile(const bf16* __restrict__ base, size_t ld, int r0, int rmax, int dh, bf16 (*dst)[LDS]) {
    for (int idx = threadIdx.x; idx < TK * 4; idx += blockDim.x) {
        const int r = idx >> 2, c = (idx & 3) * 8;
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (r0 + r < rmax) u = *reinterpret_cast<const uint4*>(base + (size_t)(r0 + r) * ld + c);
        if (c + 8 > dh) {
            bf16* e = reinterpret_cast<bf16*>(&u);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (c + i >= dh) e[i] = __float2bfloat16_rn(0.f);
        }
        *reinterpret_cast<uint4*>(&dst[r][c]) = u;
    }
}

// asynchronous variant for the STREAMED operand (cp.async, zero-fill for rows >= rmax).  Its pad lanes only ever meet
// zero A-fragment lanes or output columns that are masked at the store, so they need no zeroing.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void load_tile_async(const bf16* __restrict__ base, size_t ld, int r0, int rmax, bf16 (*dst)[LDS]) {
    for (int idx = threadIdx.x; idx < TK * 4; idx += blockDim.x) {
        const int r = idx >> 2, c = (idx & 3) * 8;
        const bool ok = r0 + r < rmax;
        cp_async16((uint32_t)__cvta_generic_to_shared(&dst[r][c]), base + (size_t)(ok ? r0 + r : rmax - 1) * ld + c, ok);
    }
}

// A-operand fragments (16 rows x 32 k) of the warp's 16-row slab of a smem tile
__device__ __forceinline__ void load_a_frags(const bf16 (*t)[LDS], int row0, int g, int tq, uint32_t (*a)[4]) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        a[ks][0] = lds32(&t[row0 + g][16 * ks + 2 * tq]);
        a[ks][1] = lds32(&t[row0 + g + 8][16 * ks + 2 * tq]);
        a[ks][2] = lds32(&t[row0 + g][16 * ks + 2 * tq + 8]);
        a[ks][3] = lds32(&t[row0 + g + 8][16 * ks + 2 * tq + 8]);
    }
}

// acc[j] (16 x 8 tiles, j = 0..7) = A(16 x 32) * T^T where T is a [64][32] smem tile (B[k=d][n=row of T])
__device__ __forceinline__ void mma_a_tT(float (*acc)[4], const uint32_t (*a)[4], const bf16 (*t)[LDS], int g, int tq) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const uint32_t b0 = lds32(&t[8 * j + g][16 * ks + 2 * tq]);
            const uint32_t b1 = lds32(&t[8 * j + g][16 * ks + 2 * tq + 8]);
            mma16816(acc[j], a[ks], b0, b1);
        }
    }
}

// out[n] (16 x 8 tiles over d, n = 0..3) += P(16 x 64, accumulator layout p[8][4]) * T, T = [64][32] smem tile (B[k=row of T][n=d])
__device__ __forceinline__ void mma_p_t(float (*out)[4], const float (*p)[4], const bf16 (*t)[LDS], int lane) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4];
        a[0] = pack2(p[2 * kk][0], p[2 * kk][1]);
        a[1] = pack2(p[2 * kk][2], p[2 * kk][3]);
        a[2] = pack2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
        a[3] = pack2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
        for (int nd = 0; nd < 2; ++nd) {
            uint32_t r0, r1, r2, r3;
            ldmatrix_x4_trans(r0, r1, r2, r3, &t[16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8][16 * nd + (lane >> 4) * 8]);
            mma16816(out[2 * nd], a, r0, r1);
            mma16816(out[2 * nd + 1], a, r2, r3);
        }
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

// dropout multiplier of probability (row base index + key); the 64-bit row base is computed once per tile row
__device__ __forceinline__ float drop_at(const DropCfg& drop, uint64_t rowbase, int key) {
    return ick_hash(drop.seed, drop.site, rowbase + (uint64_t)key) >= drop.thr ? drop.inv_keep : 0.f;
}

// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) fwd_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ K, const bf16* __restrict__ V,
                                                  bf16* __restrict__ O, float* __restrict__ LSE, Dims d, DropCfg drop) {
    __shared__ __align__(16) bf16 Qs[TQ][LDS];
    __shared__ __align__(16) bf16 Ks[2][TK][LDS];
    __shared__ __align__(16) bf16 Vs[2][TK][LDS];
    ick_resolve_seed(drop);
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TQ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const bf16* Qb = Q + (size_t)b * d.Sq * d.ldq + h * HD;
    const bf16* Kb = K + (size_t)b * d.Sk * d.ldk + h * HD;
    const bf16* Vb = V + (size_t)b * d.Sk * d.ldv + h * HD;
    const int kend = d.causal ? min(d.Sk, q0 + TQ) : d.Sk;
    const int nt = (kend + TK - 1) / TK;
    load_tile_async(Kb, d.ldk, 0, d.Sk, Ks[0]);
    load_tile_async(Vb, d.ldv, 0, d.Sk, Vs[0]);
    cp_commit();
    load_tile(Qb, d.ldq, q0, d.Sq, d.dh, Qs);
    __syncthreads();
    uint32_t qa[2][4];
    load_a_frags(Qs, 16 * warp, g, tq, qa);
    const int qi0 = q0 + 16 * warp + g, qi1 = qi0 + 8;
    const uint64_t rb0 = (((uint64_t)b * d.H + h) * (uint64_t)d.Sq + qi0) * (uint64_t)d.Sk, rb1 = rb0 + 8ull * (uint64_t)d.Sk;
    const float c = d.scale_log2;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // running max of the RAW scores, running sum
    float o[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;

    for (int it = 0; it < nt; ++it) {
        const int k0 = it * TK, buf = it & 1;
        cp_wait_all();
        __syncthreads();  // tile `it` has landed for everyone, and everyone is done with tile it-1 (buffer buf^1)
        if (it + 1 < nt) {
            load_tile_async(Kb, d.ldk, k0 + TK, d.Sk, Ks[buf ^ 1]);
            load_tile_async(Vb, d.ldv, k0 + TK, d.Sk, Vs[buf ^ 1]);
            cp_commit();
        }
        float s[8][4];
        mma_a_tT(s, qa, Ks[buf], g, tq);
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
        // a row with no visible key so far keeps m = -inf; use 0 as the exponent base there (all p are exp2(-inf) = 0)
        const float e0 = mn0 == -INFINITY ? 0.f : mn0 * c, e1 = mn1 == -INFINITY ? 0.f : mn1 * c;
        const float c0 = exp2f(m0 * c - e0), c1 = exp2f(m1 * c - e1);
        l0 *= c0;
        l1 *= c1;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            o[n][0] *= c0; o[n][1] *= c0; o[n][2] *= c1; o[n][3] *= c1;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p = exp2f(fmaf(s[j][e], c, -(e < 2 ? e0 : e1)));
                if (e < 2) l0 += p; else l1 += p;
                s[j][e] = p;
            }
        if (drop.thr != 0u) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) s[j][e] *= drop_at(drop, e < 2 ? rb0 : rb1, k0 + 8 * j + 2 * tq + (e & 1));
        }
        mma_p_t(o, s, Vs[buf], lane);
        m0 = mn0;
        m1 = mn1;
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    bf16* Ob = O + (size_t)b * d.Sq * d.ldo + h * HD;
    store_slab(Ob, d.ldo, qi0, qi1, d.Sq, o, 1.f / l0, 1.f / l1, d.dh, tq);
    if (tq == 0) {
        float* L = LSE + ((size_t)b * d.H + h) * d.Sq;
        if (qi0 < d.Sq) L[qi0] = m0 * c + log2f(l0);
        if (qi1 < d.Sq) L[qi1] = m1 * c + log2f(l1);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bwd_dq_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ K, const bf16* __restrict__ V,
                                                     const bf16* __restrict__ O, const bf16* __restrict__ dO, const float* __restrict__ LSE,
                                                     float* __restrict__ Dsum, bf16* __restrict__ dQ, Dims d, int lddo, int lddq, DropCfg drop) {
    __shared__ __align__(16) bf16 Qs[TQ][LDS];
    __shared__ __align__(16) bf16 Gs[TQ][LDS];
    __shared__ __align__(16) bf16 Ks[2][TK][LDS];
    __shared__ __align__(16) bf16 Vs[2][TK][LDS];
    __shared__ float Ds[TQ];
    ick_resolve_seed(drop);
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TQ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const bf16* Qb = Q + (size_t)b * d.Sq * d.ldq + h * HD;
    const bf16* Kb = K + (size_t)b * d.Sk * d.ldk + h * HD;
    const bf16* Vb = V + (size_t)b * d.Sk * d.ldv + h * HD;
    const bf16* Gb = dO + (size_t)b * d.Sq * lddo + h * HD;
    const bf16* Ob = O + (size_t)b * d.Sq * d.ldo + h * HD;
    const int kend = d.causal ? min(d.Sk, q0 + TQ) : d.Sk;
    const int nt = (kend + TK - 1) / TK;
    load_tile_async(Kb, d.ldk, 0, d.Sk, Ks[0]);
    load_tile_async(Vb, d.ldv, 0, d.Sk, Vs[0]);
    cp_commit();
    load_tile(Qb, d.ldq, q0, d.Sq, d.dh, Qs);
    load_tile(Gb, lddo, q0, d.Sq, d.dh, Gs);
    load_tile(Ob, d.ldo, q0, d.Sq, d.dh, Ks[1]);  // O staged in the second K buffer for the row dot products
    __syncthreads();
    if (threadIdx.x < TQ) {
        float acc = 0.f;
#pragma unroll
        for (int cc = 0; cc < HD; ++cc) acc = fmaf(__bfloat162float(Gs[threadIdx.x][cc]), __bfloat162float(Ks[1][threadIdx.x][cc]), acc);
        Ds[threadIdx.x] = acc;
        if (q0 + threadIdx.x < d.Sq) Dsum[((size_t)b * d.H + h) * d.Sq + q0 + threadIdx.x] = acc;
    }
    uint32_t qa[2][4], ga[2][4];
    load_a_frags(Qs, 16 * warp, g, tq, qa);
    load_a_frags(Gs, 16 * warp, g, tq, ga);
    __syncthreads();  // Ds visible; the staged O may be overwritten from here on
    const int qi0 = q0 + 16 * warp + g, qi1 = qi0 + 8;
    const uint64_t rb0 = (((uint64_t)b * d.H + h) * (uint64_t)d.Sq + qi0) * (uint64_t)d.Sk, rb1 = rb0 + 8ull * (uint64_t)d.Sk;
    const float* L = LSE + ((size_t)b * d.H + h) * d.Sq;
    const float lse0 = qi0 < d.Sq ? L[qi0] : 0.f, lse1 = qi1 < d.Sq ? L[qi1] : 0.f;
    const float D0 = Ds[16 * warp + g], D1 = Ds[16 * warp + g + 8];
    const float c = d.scale_log2;
    float dq[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;

    for (int it = 0; it < nt; ++it) {
        const int k0 = it * TK, buf = it & 1;
        cp_wait_all();
        __syncthreads();
        if (it + 1 < nt) {
            load_tile_async(Kb, d.ldk, k0 + TK, d.Sk, Ks[buf ^ 1]);
            load_tile_async(Vb, d.ldv, k0 + TK, d.Sk, Vs[buf ^ 1]);
            cp_commit();
        }
        float s[8][4], dp[8][4];
        mma_a_tT(s, qa, Ks[buf], g, tq);
        mma_a_tT(dp, ga, Vs[buf], g, tq);
        const bool need_mask = k0 + TK > d.Sk || q0 + TQ > d.Sq || (d.causal && k0 + TK - 1 > q0 + 16 * warp);
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = k0 + 8 * j + 2 * tq + (e & 1);
                float p = exp2f(fmaf(s[j][e], c, -(e < 2 ? lse0 : lse1)));
                if (need_mask) {
                    const int qi = e < 2 ? qi0 : qi1;
                    if (!(key < d.Sk && qi < d.Sq && (!d.causal || key <= qi))) p = 0.f;
                }
                float g_ = dp[j][e];
                if (drop.thr != 0u) g_ *= drop_at(drop, e < 2 ? rb0 : rb1, key);
                s[j][e] = p * (g_ - (e < 2 ? D0 : D1));
            }
        mma_p_t(dq, s, Ks[buf], lane);
    }
    bf16* dQb = dQ + (size_t)b * d.Sq * lddq + h * HD;
    store_slab(dQb, lddq, qi0, qi1, d.Sq, dq, d.scale, d.scale, d.dh, tq);
}

// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bwd_dkv_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ K, const bf16* __restrict__ V,
                                                      const bf16* __restrict__ dO, const float* __restrict__ LSE, const float* __restrict__ Dsum,
                                                      bf16* __restrict__ dK, bf16* __restrict__ dV, Dims d, int lddo, int lddk, int lddv,
                                                      DropCfg drop) {
    __shared__ __align__(16) bf16 Ks[TQ][LDS];
    __shared__ __align__(16) bf16 Vs[TQ][LDS];
    __shared__ __align__(16) bf16 Qs[2][TK][LDS];
    __shared__ __align__(16) bf16 Gs[2][TK][LDS];
    __shared__ float Ls[2][TK], Ds[2][TK];
    ick_resolve_seed(drop);
    const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * TQ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const bf16* Qb = Q + (size_t)b * d.Sq * d.ldq + h * HD;
    const bf16* Kb = K + (size_t)b * d.Sk * d.ldk + h * HD;
    const bf16* Vb = V + (size_t)b * d.Sk * d.ldv + h * HD;
    const bf16* Gb = dO + (size_t)b * d.Sq * lddo + h * HD;
    const float* L = LSE + ((size_t)b * d.H + h) * d.Sq;
    const float* Dg = Dsum + ((size_t)b * d.H + h) * d.Sq;
    // causal: queries before the first key of this CTA see none of its keys
    const int qbeg = d.causal ? (j0 / TK) * TK : 0;
    const int nt = (d.Sq - qbeg + TK - 1) / TK;

    auto issue = [&](int q0, int buf) {
        load_tile_async(Qb, d.ldq, q0, d.Sq, Qs[buf]);
        load_tile_async(Gb, lddo, q0, d.Sq, Gs[buf]);
        if (threadIdx.x < TK) {
            const int i = q0 + threadIdx.x;
            const bool ok = i < d.Sq;
            cp_async4((uint32_t)__cvta_generic_to_shared(&Ls[buf][threadIdx.x]), L + (ok ? i : 0), ok);
            cp_async4((uint32_t)__cvta_generic_to_shared(&Ds[buf][threadIdx.x]), Dg + (ok ? i : 0), ok);
        }
        cp_commit();
    };
    issue(qbeg, 0);
    load_tile(Kb, d.ldk, j0, d.Sk, d.dh, Ks);
    load_tile(Vb, d.ldv, j0, d.Sk, d.dh, Vs);
    __syncthreads();
    uint32_t ka[2][4], va[2][4];
    load_a_frags(Ks, 16 * warp, g, tq, ka);
    load_a_frags(Vs, 16 * warp, g, tq, va);
    const int kj0 = j0 + 16 * warp + g, kj1 = kj0 + 8;
    const uint64_t hb = ((uint64_t)b * d.H + h) * (uint64_t)d.Sq;
    const float c = d.scale_log2;
    float dk[4][4], dv[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
        dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
    }
    for (int it = 0; it < nt; ++it) {
        const int q0 = qbeg + it * TK, buf = it & 1;
        cp_wait_all();
        __syncthreads();
        if (it + 1 < nt) issue(q0 + TK, buf ^ 1);
        float st[8][4], dpt[8][4];
        mma_a_tT(st, ka, Qs[buf], g, tq);
        mma_a_tT(dpt, va, Gs[buf], g, tq);
        const bool need_mask = q0 + TK > d.Sq || j0 + TQ > d.Sk || (d.causal && j0 + 16 * warp + 15 > q0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int qc = 8 * j + 2 * tq + (e & 1);
                const int qi = q0 + qc;
                const int key = e < 2 ? kj0 : kj1;
                float p = exp2f(fmaf(st[j][e], c, -Ls[buf][qc]));
                if (need_mask) {
                    if (!(key < d.Sk && qi < d.Sq && (!d.causal || key <= qi))) p = 0.f;
                }
                float mul = 1.f;
                if (drop.thr != 0u) mul = drop_at(drop, (hb + (uint64_t)qi) * (uint64_t)d.Sk, key);
                st[j][e] = p * mul;                                // P^T with dropout -> dV
                dpt[j][e] = p * (dpt[j][e] * mul - Ds[buf][qc]);   // dS^T              -> dK
            }
        mma_p_t(dv, st, Gs[buf], lane);
        mma_p_t(dk, dpt, Qs[buf], lane);
    }
    bf16* dKb = dK + (size_t)b * d.Sk * lddk + h * HD;
    bf16* dVb = dV + (size_t)b * d.Sk * lddv + h * HD;
    store_slab(dKb, lddk, kj0, kj1, d.Sk, dk, d.scale, d.scale, d.dh, tq);
    store_slab(dVb, lddv, kj0, kj1, d.Sk, dv, 1.f, 1.f, d.dh, tq);
}

Dims make_dims(int B, int H, int Sq, int Sk, int dh, int ldq, int ldk, int ldv, int ldo, int causal) {
    Dims d;
    d.B = B; d.H = H; d.Sq = Sq; d.Sk = Sk; d.dh = dh;
    d.ldq = ldq; d.ldk = ldk; d.ldv = ldv; d.ldo = ldo;
    d.causal = causal;
    d.scale = 1.0f / sqrtf((float)dh);
    d.scale_log2 = d.scale * 1.4426950408889634f;
    return d;
}

}  // namespace

int ick_mha_fwd_mma(const void* Q, const void* K, const void* V, void* O, float* lse, int B, int H, int Sq, int Sk, int dh, int ldq, int ldk,
                    int ldv, int ldo, int causal, DropCfg dc, cudaStream_t stream) {
    Dims d = make_dims(B, H, Sq, Sk, dh, ldq, ldk, ldv, ldo, causal);
    dim3 grid((Sq + TQ - 1) / TQ, H, B);
    fwd_kernel<<<grid, 128, 0, stream>>>((const bf16*)Q, (const bf16*)K, (const bf16*)V, (bf16*)O, lse, d, dc);
    return ick_check_launch("mha_fwd_mma");
}

int ick_mha_bwd_mma(const void* Q, const void* K, const void* V, const void* O, const void* dO, const float* lse, float* dsum, void* dQ,
                    void* dK, void* dV, int B, int H, int Sq, int Sk, int dh, int ldq, int ldk, int ldv, int ldo, int lddo, int lddq, int lddk,
                    int lddv, int causal, DropCfg dc, cudaStream_t stream) {
    Dims d = make_dims(B, H, Sq, Sk, dh, ldq, ldk, ldv, ldo, causal);
    dim3 gq((Sq + TQ - 1) / TQ, H, B), gk((Sk + TQ - 1) / TQ, H, B);
    bwd_dq_kernel<<<gq, 128, 0, stream>>>((const bf16*)Q, (const bf16*)K, (const bf16*)V, (const bf16*)O, (const bf16*)dO, lse, dsum, (bf16*)dQ,
                                          d, lddo, lddq, dc);
    bwd_dkv_kernel<<<gk, 128, 0, stream>>>((const bf16*)Q, (const bf16*)K, (const bf16*)V, (const bf16*)dO, lse, dsum, (bf16*)dK, (bf16*)dV, d,
                                           lddo, lddk, lddv, dc);
    return ick_check_launch("mha_bwd_mma");
}
