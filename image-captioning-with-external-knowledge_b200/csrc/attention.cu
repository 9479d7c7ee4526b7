// Multi-head scaled-dot-product attention, forward and backward, for head_dim <= 32 (the reference uses 10 heads of
// 30, nn.TransformerEncoderLayer/DecoderLayer(300, 10, ...), G/models.py:241-244).  Four call sites: entity and fact
// self-attention (no mask), decoder causal self-attention, decoder cross-attention over [pixels; entities; facts].
//
// Layout: Q/K/V/O rows are (batch, position); head h occupies columns [h*32, h*32+32) ("head layout": the 30 real
// columns followed by 2 zero pads, produced by the packed projection weights), so a head row is one aligned
// 64 B (bf16) / 128 B (fp32) vector.  Flash-style: one thread owns one query (fwd, dQ) or one key (dK/dV), the other
// operand streams through shared memory in tiles and is read as warp-wide broadcasts; the softmax is online
// (running max / sum in the exp2 domain) so the score matrix never exists in memory.  Attention-probability dropout
// (torch applies it after the softmax) is regenerated from the counter hash in the backward pass.
#include <cstdlib>

#include "attention_internal.h"
#include "common.cuh"
#include "ickb200.h"

namespace {

constexpr int HD = 32;    // padded head dim
constexpr int NT = 128;   // threads per CTA = queries (or keys) per CTA
constexpr int KT = 64;    // keys per shared-memory tile (fwd / dQ)
constexpr int QT = 32;    // queries per shared-memory tile (dK/dV)

template <typename T>
__device__ __forceinline__ void load_row32(const T* p, float* v) {
#pragma unroll
    for (int c = 0; c < HD; c += 8) ld8(p + c, v + c);
}
template <typename T>
__device__ __forceinline__ void store_row32(T* p, const float* v) {
#pragma unroll
    for (int c = 0; c < HD; c += 8) st8(p + c, v + c);
}

// cooperative tile load: rows [r0, r0+nrows) of a (rows, ld) matrix, 32 columns starting at col0, into smem as floats
template <typename T>
__device__ __forceinline__ void load_tile(const T* base, size_t ld, int r0, int rmax, int nrows, float (*dst)[HD]) {
    for (int idx = threadIdx.x; idx < nrows * 4; idx += NT) {
        const int r = idx >> 2, c = (idx & 3) * 8;
        float v[8];
        if (r0 + r < rmax) ld8(base + (size_t)(r0 + r) * ld + c, v);
        else
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        *reinterpret_cast<float4*>(&dst[r][c]) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(&dst[r][c + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    }
}

__device__ __forceinline__ float dot32(const float* a, const float* b_smem) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < HD; c += 4) {
        const float4 b = *reinterpret_cast<const float4*>(b_smem + c);
        s = fmaf(a[c], b.x, s);
        s = fmaf(a[c + 1], b.y, s);
        s = fmaf(a[c + 2], b.z, s);
        s = fmaf(a[c + 3], b.w, s);
    }
    return s;
}
__device__ __forceinline__ void axpy32(float* acc, float a, const float* x_smem) {
#pragma unroll
    for (int c = 0; c < HD; c += 4) {
        const float4 x = *reinterpret_cast<const float4*>(x_smem + c);
        acc[c] = fmaf(a, x.x, acc[c]);
        acc[c + 1] = fmaf(a, x.y, acc[c + 1]);
        acc[c + 2] = fmaf(a, x.z, acc[c + 2]);
        acc[c + 3] = fmaf(a, x.w, acc[c + 3]);
    }
}

struct AttnDims {
    int B, H, Sq, Sk, dh;
    int ldq, ldk, ldv, ldo;
    int causal;
    float scale_log2;  // (1/sqrt(dh)) * log2(e)
    float scale;       // 1/sqrt(dh)
};

// dropout row of query (b,h,i): probabilities are addressed as (row, key)
__device__ __forceinline__ uint64_t prow(const AttnDims& d, int b, int h, int i) {
    return ((uint64_t)b * d.H + h) * (uint64_t)d.Sq + i;
}

// Attention-probability dropout (common.cuh: bit-parallel keep words): multiplier of probability (row mix rmix, key j); the keep
// word of the key's group of 32 is cached across consecutive keys of a row.
struct KeepCache {
    uint32_t kg = 0xFFFFFFFFu, word = 0u;
    __device__ __forceinline__ float mul(const DropCfg& drop, uint32_t rmix, uint32_t key) {
        if (drop.thr == 0u) return 1.0f;
        if ((key >> 5) != kg) {
            kg = key >> 5;
            word = ick_keepword(rmix, kg, ick_attn_t16(drop.thr));
        }
        return ((word >> ick_keybit(key)) & 1u) ? drop.inv_keep : 0.0f;
    }
};

// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT) mha_fwd_kernel(const T* __restrict__ Q, const T* __restrict__ K,
                                                     const T* __restrict__ V, T* __restrict__ O, float* __restrict__ LSE,
                                                     AttnDims d, DropCfg drop) {
    ick_pdl_entry();
    __shared__ __align__(16) float Ks[KT][HD];
    __shared__ __align__(16) float Vs[KT][HD];
    ick_resolve_seed(drop);
    const int b = blockIdx.z, h = blockIdx.y;
    const int i = blockIdx.x * NT + threadIdx.x;
    const bool active = i < d.Sq;
    const T* Kb = K + (size_t)b * d.Sk * d.ldk + h * HD;
    const T* Vb = V + (size_t)b * d.Sk * d.ldv + h * HD;

    float q[HD], acc[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) { q[c] = 0.f; acc[c] = 0.f; }
    if (active) load_row32(Q + ((size_t)b * d.Sq + i) * d.ldq + h * HD, q);
#pragma unroll
    for (int c = 0; c < HD; ++c) q[c] = (c < d.dh) ? q[c] * d.scale_log2 : 0.f;  // pad lanes never contribute
    float m = -INFINITY, l = 0.f;
    const uint32_t rmix = ick_rowmix(drop.seed, drop.site, prow(d, b, h, i));
    KeepCache kc;

    // causal: no key beyond the last query of this CTA is visible
    const int kmax = d.causal ? min(d.Sk, (int)(blockIdx.x * NT + NT)) : d.Sk;
    for (int k0 = 0; k0 < kmax; k0 += KT) {
        __syncthreads();
        load_tile(Kb, d.ldk, k0, d.Sk, KT, Ks);
        load_tile(Vb, d.ldv, k0, d.Sk, KT, Vs);
        __syncthreads();
        const int nk = min(KT, d.Sk - k0);
        if (!active) continue;
        for (int j0 = 0; j0 < nk; j0 += 8) {
            float s[8];
            float cmax = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = k0 + j0 + jj;
                const bool vis = (j0 + jj < nk) && (!d.causal || j <= i);
                s[jj] = vis ? dot32(q, Ks[j0 + jj]) : -INFINITY;
                cmax = fmaxf(cmax, s[jj]);
            }
            if (cmax == -INFINITY) continue;  // whole chunk masked
            const float mnew = fmaxf(m, cmax);
            const float corr = exp2f(m - mnew);  // m = -inf on the first visible chunk -> 0
            l *= corr;
#pragma unroll
            for (int c = 0; c < HD; ++c) acc[c] *= corr;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float p = exp2f(s[jj] - mnew);  // masked -> exp2(-inf) = 0
                l += p;
                const float pd = p * kc.mul(drop, rmix, (uint32_t)(k0 + j0 + jj));
                axpy32(acc, pd, Vs[j0 + jj]);
            }
            m = mnew;
        }
    }
    if (active) {
        const float inv = 1.f / l;
#pragma unroll
        for (int c = 0; c < HD; ++c) acc[c] = (c < d.dh) ? acc[c] * inv : 0.f;
        store_row32(O + ((size_t)b * d.Sq + i) * d.ldo + h * HD, acc);
        LSE[((size_t)b * d.H + h) * d.Sq + i] = m + log2f(l);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// dQ: one thread per query.  Also writes Dsum[b,h,i] = sum_c dO*O for the dK/dV kernel.
template <typename T>
__global__ void __launch_bounds__(NT) mha_bwd_dq_kernel(const T* __restrict__ Q, const T* __restrict__ K,
                                                        const T* __restrict__ V, const T* __restrict__ O,
                                                        const T* __restrict__ dO, const float* __restrict__ LSE,
                                                        float* __restrict__ Dsum, T* __restrict__ dQ, AttnDims d,
                                                        int lddo, int lddq, DropCfg drop) {
    ick_pdl_entry();
    __shared__ __align__(16) float Ks[KT][HD];
    __shared__ __align__(16) float Vs[KT][HD];
    ick_resolve_seed(drop);
    const int b = blockIdx.z, h = blockIdx.y;
    const int i = blockIdx.x * NT + threadIdx.x;
    const bool active = i < d.Sq;
    const T* Kb = K + (size_t)b * d.Sk * d.ldk + h * HD;
    const T* Vb = V + (size_t)b * d.Sk * d.ldv + h * HD;

    float q[HD], go[HD], dq[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) { q[c] = 0.f; go[c] = 0.f; dq[c] = 0.f; }
    float lse = 0.f, Di = 0.f;
    if (active) {
        load_row32(Q + ((size_t)b * d.Sq + i) * d.ldq + h * HD, q);
        load_row32(dO + ((size_t)b * d.Sq + i) * lddo + h * HD, go);
        float o[HD];
        load_row32(O + ((size_t)b * d.Sq + i) * d.ldo + h * HD, o);
#pragma unroll
        for (int c = 0; c < HD; ++c) Di = fmaf(go[c], o[c], Di);
        lse = LSE[((size_t)b * d.H + h) * d.Sq + i];
        Dsum[((size_t)b * d.H + h) * d.Sq + i] = Di;
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) {
        q[c] = (c < d.dh) ? q[c] * d.scale_log2 : 0.f;
        go[c] = (c < d.dh) ? go[c] : 0.f;
    }

    const uint32_t rmix = ick_rowmix(drop.seed, drop.site, prow(d, b, h, i));
    KeepCache kc;
    const int kmax = d.causal ? min(d.Sk, (int)(blockIdx.x * NT + NT)) : d.Sk;
    for (int k0 = 0; k0 < kmax; k0 += KT) {
        __syncthreads();
        load_tile(Kb, d.ldk, k0, d.Sk, KT, Ks);
        load_tile(Vb, d.ldv, k0, d.Sk, KT, Vs);
        __syncthreads();
        const int nk = min(KT, d.Sk - k0);
        if (!active) continue;
        for (int jj = 0; jj < nk; ++jj) {
            const int j = k0 + jj;
            if (d.causal && j > i) break;
            const float p = exp2f(dot32(q, Ks[jj]) - lse);
            const float dp = dot32(go, Vs[jj]) * kc.mul(drop, rmix, (uint32_t)j);
            const float ds = p * (dp - Di);
            axpy32(dq, ds, Ks[jj]);
        }
    }
    if (active) {
#pragma unroll
        for (int c = 0; c < HD; ++c) dq[c] = (c < d.dh) ? dq[c] * d.scale : 0.f;
        store_row32(dQ + ((size_t)b * d.Sq + i) * lddq + h * HD, dq);
    }
}

// dK, dV: one thread per key; queries stream through shared memory.
template <typename T>
__global__ void __launch_bounds__(NT) mha_bwd_dkv_kernel(const T* __restrict__ Q, const T* __restrict__ K,
                                                         const T* __restrict__ V, const T* __restrict__ dO,
                                                         const float* __restrict__ LSE, const float* __restrict__ Dsum,
                                                         T* __restrict__ dK, T* __restrict__ dV, AttnDims d, int lddo,
                                                         int lddk, int lddv, DropCfg drop) {
    ick_pdl_entry();
    __shared__ __align__(16) float Qs[QT][HD];
    __shared__ __align__(16) float Gs[QT][HD];
    __shared__ float Ls[QT], Ds[QT];
    ick_resolve_seed(drop);
    const int b = blockIdx.z, h = blockIdx.y;
    const int j = blockIdx.x * NT + threadIdx.x;
    const bool active = j < d.Sk;
    const T* Qb = Q + (size_t)b * d.Sq * d.ldq + h * HD;
    const T* Gb = dO + (size_t)b * d.Sq * lddo + h * HD;
    const float* Lb = LSE + ((size_t)b * d.H + h) * d.Sq;
    const float* Db = Dsum + ((size_t)b * d.H + h) * d.Sq;

    float k[HD], v[HD], dk[HD], dv[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) { k[c] = 0.f; v[c] = 0.f; dk[c] = 0.f; dv[c] = 0.f; }
    if (active) {
        load_row32(K + ((size_t)b * d.Sk + j) * d.ldk + h * HD, k);
        load_row32(V + ((size_t)b * d.Sk + j) * d.ldv + h * HD, v);
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) {
        k[c] = (c < d.dh) ? k[c] * d.scale_log2 : 0.f;
        v[c] = (c < d.dh) ? v[c] : 0.f;
    }

    // causal: queries before the first key of this CTA see none of its keys
    const int qbeg = d.causal ? (int)(blockIdx.x * NT) / QT * QT : 0;
    for (int q0 = qbeg; q0 < d.Sq; q0 += QT) {
        __syncthreads();
        load_tile(Qb, d.ldq, q0, d.Sq, QT, Qs);
        load_tile(Gb, lddo, q0, d.Sq, QT, Gs);
        if (threadIdx.x < QT) {
            const int i = q0 + threadIdx.x;
            Ls[threadIdx.x] = i < d.Sq ? Lb[i] : 0.f;
            Ds[threadIdx.x] = i < d.Sq ? Db[i] : 0.f;
        }
        __syncthreads();
        const int nq = min(QT, d.Sq - q0);
        if (!active) continue;
        for (int ii = 0; ii < nq; ++ii) {
            const int i = q0 + ii;
            if (d.causal && j > i) continue;
            const float p = exp2f(dot32(k, Qs[ii]) - Ls[ii]);
            KeepCache kc;  // one key per thread, a new row per query: nothing to cache (fp32 parity path)
            const float mul = kc.mul(drop, ick_rowmix(drop.seed, drop.site, prow(d, b, h, i)), (uint32_t)j);
            axpy32(dv, p * mul, Gs[ii]);
            const float dp = dot32(v, Gs[ii]) * mul;
            const float ds = p * (dp - Ds[ii]);
            axpy32(dk, ds, Qs[ii]);
        }
    }
    if (active) {
#pragma unroll
        for (int c = 0; c < HD; ++c) {
            dk[c] = (c < d.dh) ? dk[c] * d.scale : 0.f;
            dv[c] = (c < d.dh) ? dv[c] : 0.f;
        }
        store_row32(dK + ((size_t)b * d.Sk + j) * lddk + h * HD, dk);
        store_row32(dV + ((size_t)b * d.Sk + j) * lddv + h * HD, dv);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Incremental (KV-cached) decode attention: one query per (batch, head); a warp per (b,h), lanes split the keys.
// K/V caches hold `klen[b]`-many valid rows (or `klen_all` if klen is null).
template <typename T>
__global__ void __launch_bounds__(128) mha_decode_kernel(const T* __restrict__ Q, const T* __restrict__ K,
                                                         const T* __restrict__ V, T* __restrict__ O, int B, int H, int dh,
                                                         int ldq, int ldk, int ldv, int ldo, long long kbatch_stride,
                                                         long long vbatch_stride, int klen, float scale_log2) {
    ick_pdl_entry();
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= B * H) return;
    const int b = w / H, h = w % H;
    float q[HD];
    load_row32(Q + (size_t)b * ldq + h * HD, q);
#pragma unroll
    for (int c = 0; c < HD; ++c) q[c] = (c < dh) ? q[c] * scale_log2 : 0.f;
    const T* Kb = K + (size_t)b * kbatch_stride + h * HD;
    const T* Vb = V + (size_t)b * vbatch_stride + h * HD;
    float m = -INFINITY, l = 0.f, acc[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = 0.f;
    for (int j = lane; j < klen; j += 32) {
        float kv[HD];
        load_row32(Kb + (size_t)j * ldk, kv);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < HD; ++c) s = fmaf(q[c], kv[c], s);
        const float mnew = fmaxf(m, s);
        const float corr = exp2f(m - mnew);
        const float p = exp2f(s - mnew);
        l = l * corr + p;
        load_row32(Vb + (size_t)j * ldv, kv);
#pragma unroll
        for (int c = 0; c < HD; ++c) acc[c] = fmaf(p, kv[c], acc[c] * corr);
        m = mnew;
    }
    // merge the 32 partial softmaxes
    const float mall = warp_max(m);
    const float f = (m == -INFINITY) ? 0.f : exp2f(m - mall);
    l = warp_sum(l * f);
#pragma unroll
    for (int c = 0; c < HD; ++c) acc[c] = warp_sum(acc[c] * f);
    if (lane == 0) {
        const float inv = 1.f / l;
#pragma unroll
        for (int c = 0; c < HD; ++c) acc[c] = (c < dh) ? acc[c] * inv : 0.f;
        store_row32(O + (size_t)b * ldo + h * HD, acc);
    }
}

// bf16 decode attention over a cache whose K and V blocks are adjacent (K columns [0, H*32), V columns [H*32, 2*H*32) of the
// same row - both the per-layer memory K/V projection and the self-attention q|k|v cache are laid out like that).  One CTA per
// image; a thread owns (key group, head, quarter of the head): 40 consecutive threads read the contiguous 640-byte K block and
// the 640-byte V block of one cached position with 16-byte loads, eight positions per iteration.  The warp-per-(image, head)
// kernel above reads 64-byte rows 3840 bytes apart, which is what limits it to ~3 TB/s.
#ifndef ICK_DR_KG
#define ICK_DR_KG 8
#endif
constexpr int DR_KG = ICK_DR_KG;  // cached positions per iteration
#ifndef ICK_DR_UNROLL
#define ICK_DR_UNROLL 2  // K/V row pairs a thread keeps in flight (A/B builds: -DICK_DR_UNROLL=4)
#endif
constexpr int DR_UNROLL = ICK_DR_UNROLL;
__global__ void __launch_bounds__(DR_KG * 40) mha_decode_rows_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ KV, bf16* __restrict__ O,
                                                                     int H, int dh, int ldq, int ldkv, int ldo, long long batch_stride,
                                                                     int klen, float scale_log2) {
    ick_pdl_entry();
    extern __shared__ float dr_red[];  // [DR_KG][H*4][10]: m, l, acc[8] of every (key group, head, quarter)
    const int slots = H * 4;           // threads per cached position
    const int kg = threadIdx.x / slots, sl = threadIdx.x % slots, h = sl >> 2, qd = sl & 3;
    const int b = blockIdx.x;
    float q[8];
    ld8(Q + (size_t)b * ldq + h * HD + qd * 8, q);
#pragma unroll
    for (int c = 0; c < 8; ++c) q[c] = (qd * 8 + c < dh) ? q[c] * scale_log2 : 0.f;  // pad lanes never contribute
    const bf16* base = KV + (size_t)b * batch_stride + h * HD + qd * 8;
    const int voff = H * HD;
    const unsigned qmask = 0xFu << ((threadIdx.x & 31) & ~3);
    float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll DR_UNROLL
    for (int j = kg; j < klen; j += DR_KG) {
        float kx[8], vx[8];
        ld8(base + (size_t)j * ldkv, kx);
        ld8(base + (size_t)j * ldkv + voff, vx);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) s = fmaf(q[c], kx[c], s);
        s += __shfl_xor_sync(qmask, s, 1);  // the four quarter threads of a (position, head) iterate together; other lanes of the
        s += __shfl_xor_sync(qmask, s, 2);  // warp may belong to another key group with a different trip count
        const float mnew = fmaxf(m, s);
        const float corr = exp2f(m - mnew);
        const float p = exp2f(s - mnew);
        l = l * corr + p;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fmaf(p, vx[c], acc[c] * corr);
        m = mnew;
    }
    float* mine = dr_red + ((size_t)kg * slots + sl) * 10;
    mine[0] = m;
    mine[1] = l;
#pragma unroll
    for (int c = 0; c < 8; ++c) mine[2 + c] = acc[c];
    __syncthreads();
    if (kg != 0) return;
    float mall = m;
    for (int g = 1; g < DR_KG; ++g) mall = fmaxf(mall, dr_red[((size_t)g * slots + sl) * 10]);
    float lsum = 0.f, out[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) out[c] = 0.f;
    for (int g = 0; g < DR_KG; ++g) {
        const float* r = dr_red + ((size_t)g * slots + sl) * 10;
        const float f = r[0] == -INFINITY ? 0.f : exp2f(r[0] - mall);
        lsum = fmaf(r[1], f, lsum);
#pragma unroll
        for (int c = 0; c < 8; ++c) out[c] = fmaf(r[2 + c], f, out[c]);
    }
    const float inv = 1.f / lsum;
#pragma unroll
    for (int c = 0; c < 8; ++c) out[c] = (qd * 8 + c < dh) ? out[c] * inv : 0.f;
    st8(O + (size_t)b * ldo + h * HD + qd * 8, out);
}

// Beam-search decode attention (extension: the reference decodes greedily): the G beams of an image are
// consecutive query rows.  Same thread layout as mha_decode_rows_kernel - a thread owns (key group, head, quarter of the
// head) - but every cached K / V row that is loaded serves all G queries of the CTA.
//   anc == nullptr (cross-attention): one CTA per image, G = beams per image; the beams share the image's memory K/V, which
//     is therefore streamed once per image instead of once per beam.
//   anc != nullptr (self-attention): one CTA per beam row r (G = 1).  The cache is position-major, row (j, r') of it holds
//     the K/V that beam slot r' produced at step j, and anc[r*anc_ld + j] names the slot of row r's ancestor at step j:
//     re-ordering the beams only rewrites the small ancestor table, the cached rows never move.
template <typename T, int G>
__global__ void __launch_bounds__(DR_KG * 40) mha_decode_beam_kernel(const T* __restrict__ Q, const T* __restrict__ K, const T* __restrict__ V,
                                                                     T* __restrict__ O, int H, int dh, int ldq, int ldk, int ldv, int ldo,
                                                                     long long kimg_stride, long long vimg_stride, int klen, float scale_log2,
                                                                     const int* __restrict__ anc, int anc_ld, int group,
                                                                     long long kpos_stride, long long vpos_stride) {
    ick_pdl_entry();
    extern __shared__ float dr_red[];  // [DR_KG][H*4][10]: m, l, acc[8] of every (key group, head, quarter), one query at a time
    const int slots = H * 4;
    const int kg = threadIdx.x / slots, sl = threadIdx.x % slots, h = sl >> 2, qd = sl & 3;
    const int r0 = blockIdx.x * G;  // first query row of this CTA
    const int col = h * HD + qd * 8;
    float q[G][8];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        ld8(Q + (size_t)(r0 + g) * ldq + col, q[g]);
#pragma unroll
        for (int c = 0; c < 8; ++c) q[g][c] = (qd * 8 + c < dh) ? q[g][c] * scale_log2 : 0.f;
    }
    const int img = anc ? r0 / group : blockIdx.x;
    const T* kbase = K + (anc ? (size_t)img * group * ldk : (size_t)img * kimg_stride) + col;
    const T* vbase = V + (anc ? (size_t)img * group * ldv : (size_t)img * vimg_stride) + col;
    const int* arow = anc ? anc + (size_t)r0 * anc_ld : nullptr;
    const unsigned qmask = 0xFu << ((threadIdx.x & 31) & ~3);
    float m[G], l[G], acc[G][8];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        m[g] = -INFINITY;
        l[g] = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[g][c] = 0.f;
    }
    for (int j = kg; j < klen; j += DR_KG) {
        float kx[8], vx[8];
        if (anc) {
            const int slot = arow[j];
            ld8(kbase + (size_t)j * kpos_stride + (size_t)slot * ldk, kx);
            ld8(vbase + (size_t)j * vpos_stride + (size_t)slot * ldv, vx);
        } else {
            ld8(kbase + (size_t)j * ldk, kx);
            ld8(vbase + (size_t)j * ldv, vx);
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) s = fmaf(q[g][c], kx[c], s);
            s += __shfl_xor_sync(qmask, s, 1);
            s += __shfl_xor_sync(qmask, s, 2);
            const float mnew = fmaxf(m[g], s);
            const float corr = exp2f(m[g] - mnew);
            const float p = exp2f(s - mnew);
            l[g] = l[g] * corr + p;
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[g][c] = fmaf(p, vx[c], acc[g][c] * corr);
            m[g] = mnew;
        }
    }
    float* mine = dr_red + ((size_t)kg * slots + sl) * 10;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (g) __syncthreads();
        mine[0] = m[g];
        mine[1] = l[g];
#pragma unroll
        for (int c = 0; c < 8; ++c) mine[2 + c] = acc[g][c];
        __syncthreads();
        if (kg == 0) {
            float mall = m[g];
            for (int k2 = 1; k2 < DR_KG; ++k2) mall = fmaxf(mall, dr_red[((size_t)k2 * slots + sl) * 10]);
            float lsum = 0.f, out[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) out[c] = 0.f;
            for (int k2 = 0; k2 < DR_KG; ++k2) {
                const float* r = dr_red + ((size_t)k2 * slots + sl) * 10;
                const float f = r[0] == -INFINITY ? 0.f : exp2f(r[0] - mall);
                lsum = fmaf(r[1], f, lsum);
#pragma unroll
                for (int c = 0; c < 8; ++c) out[c] = fmaf(r[2 + c], f, out[c]);
            }
            const float inv = 1.f / lsum;
#pragma unroll
            for (int c = 0; c < 8; ++c) out[c] = (qd * 8 + c < dh) ? out[c] * inv : 0.f;
            st8(O + (size_t)(r0 + g) * ldo + col, out);
        }
    }
}

template <typename T>
int launch_decode_beam(const T* Q, const T* K, const T* V, T* O, int rows, int group, int H, int dh, int ldq, int ldk, int ldv, int ldo,
                       long long kimg_stride, long long vimg_stride, int klen, const int* anc, int anc_ld, long long kpos_stride,
                       long long vpos_stride, cudaStream_t stream) {
    const float sl2 = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
    const size_t smem = (size_t)DR_KG * H * 4 * 10 * sizeof(float);
    const int nt = DR_KG * H * 4;
#define ICK_BEAM_CASE(GG)                                                                                                              \
    case GG:                                                                                                                           \
        ick_launch(mha_decode_beam_kernel<T, GG>, rows / GG, nt, smem, stream)(Q, K, V, O, H, dh, ldq, ldk, ldv, ldo, kimg_stride, vimg_stride, \
                                                                                klen, sl2, anc, anc_ld, group, kpos_stride, vpos_stride);  \
        break;
    switch (anc ? 1 : group) {
        ICK_BEAM_CASE(1)
        ICK_BEAM_CASE(2)
        ICK_BEAM_CASE(3)
        ICK_BEAM_CASE(4)
        ICK_BEAM_CASE(5)
        ICK_BEAM_CASE(6)
        ICK_BEAM_CASE(7)
        ICK_BEAM_CASE(8)
        default:
            ick_set_error("mha_decode_beam: group %d not in [1, 8]", group);
            return ICK_ERR_UNSUPPORTED;
    }
#undef ICK_BEAM_CASE
    return ick_check_launch("mha_decode_beam");
}

// bf16 runs on the tensor-core kernels (attention_mma.cu); ICKB200_ATTN_SIMT=1 forces the CUDA-core kernels (A/B testing)
bool use_mma() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ICKB200_ATTN_SIMT");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

bool decode_rows() {  // ICK_DECODE_ROWS=0: always the warp-per-(image, head) decode kernel (A/B aid)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ICK_DECODE_ROWS");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

int check_dims(const char* what, int B, int H, int Sq, int Sk, int dh, int ldq, int ldk, int ldv, int ldo) {
    ICK_REQUIRE(B > 0 && H > 0 && Sq > 0 && Sk > 0, "%s: bad sizes B=%d H=%d Sq=%d Sk=%d", what, B, H, Sq, Sk);
    ICK_REQUIRE(dh > 0 && dh <= HD, "%s: head_dim %d not in (0, 32]", what, dh);
    ICK_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "%s: leading dims must be multiples of 8", what);
    ICK_REQUIRE(ldq >= H * HD && ldk >= H * HD && ldv >= H * HD && ldo >= H * HD, "%s: leading dims smaller than H*32", what);
    return ICK_OK;
}

AttnDims make_dims(int B, int H, int Sq, int Sk, int dh, int ldq, int ldk, int ldv, int ldo, int causal) {
    AttnDims d;
    d.B = B; d.H = H; d.Sq = Sq; d.Sk = Sk; d.dh = dh;
    d.ldq = ldq; d.ldk = ldk; d.ldv = ldv; d.ldo = ldo;
    d.causal = causal;
    d.scale = 1.0f / sqrtf((float)dh);
    d.scale_log2 = d.scale * 1.4426950408889634f;
    return d;
}

}  // namespace

extern "C" int ick_mha_fwd(const void* Q, const void* K, const void* V, void* O, float* lse, int dt, int B, int H, int Sq,
                           int Sk, int dh, int ldq, int ldk, int ldv, int ldo, int causal, float drop_p, unsigned seed,
                           unsigned site, cudaStream_t stream) {
    int rc = check_dims("mha_fwd", B, H, Sq, Sk, dh, ldq, ldk, ldv, ldo);
    if (rc) return rc;
    ICK_REQUIRE(!causal || Sq == Sk, "mha_fwd: causal needs Sq == Sk");
    AttnDims d = make_dims(B, H, Sq, Sk, dh, ldq, ldk, ldv, ldo, causal);
    DropCfg dc = make_drop(drop_p, seed, site);
    dim3 grid((Sq + NT - 1) / NT, H, B);
    if (dt == ICK_BF16 && use_mma()) return ick_mha_fwd_mma(Q, K, V, O, lse, B, H, Sq, Sk, dh, ldq, ldk, ldv, ldo, causal, dc, stream);
    if (dt == ICK_F32)
        ick_launch(mha_fwd_kernel<float>, grid, NT, 0, stream)((const float*)Q, (const float*)K, (const float*)V, (float*)O, lse, d, dc);
    else if (dt == ICK_BF16)
        ick_launch(mha_fwd_kernel<bf16>, grid, NT, 0, stream)((const bf16*)Q, (const bf16*)K, (const bf16*)V, (bf16*)O, lse, d, dc);
    else {
        ick_set_error("mha_fwd: bad dtype %d", dt);
        return ICK_ERR_UNSUPPORTED;
    }
    return ick_check_launch("mha_fwd");
}

extern "C" int ick_mha_bwd(const void* Q, const void* K, const void* V, const void* O, const void* dO, const float* lse,
                           float* dsum, void* dQ, void* dK, void* dV, int dt, int B, int H, int Sq, int Sk, int dh, int ldq,
                           int ldk, int ldv, int ldo, int lddo, int lddq, int lddk, int lddv, int causal, float drop_p,
                           unsigned seed, unsigned site, void* workspace, long long workspace_bytes, int dsum_ready,
                           cudaStream_t stream) {
    int rc = check_dims("mha_bwd", B, H, Sq, Sk, dh, ldq, ldk, ldv, ldo);
    if (rc) return rc;
    ICK_REQUIRE(lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0, "mha_bwd: grad leading dims must be multiples of 8");
    ICK_REQUIRE(!causal || Sq == Sk, "mha_bwd: causal needs Sq == Sk");
    AttnDims d = make_dims(B, H, Sq, Sk, dh, ldq, ldk, ldv, ldo, causal);
    DropCfg dc = make_drop(drop_p, seed, site);
    dim3 gq((Sq + NT - 1) / NT, H, B), gk((Sk + NT - 1) / NT, H, B);
    if (dt == ICK_BF16 && use_mma())
        return ick_mha_bwd_mma(Q, K, V, O, dO, lse, dsum, dQ, dK, dV, B, H, Sq, Sk, dh, ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv, causal, dc,
                               workspace, workspace_bytes, dsum_ready, stream);
    if (dt == ICK_F32) {
        ick_launch(mha_bwd_dq_kernel<float>, gq, NT, 0, stream)((const float*)Q, (const float*)K, (const float*)V, (const float*)O,
                                                        (const float*)dO, lse, dsum, (float*)dQ, d, lddo, lddq, dc);
        ick_launch(mha_bwd_dkv_kernel<float>, gk, NT, 0, stream)((const float*)Q, (const float*)K, (const float*)V, (const float*)dO,
                                                         lse, dsum, (float*)dK, (float*)dV, d, lddo, lddk, lddv, dc);
    } else if (dt == ICK_BF16) {
        ick_launch(mha_bwd_dq_kernel<bf16>, gq, NT, 0, stream)((const bf16*)Q, (const bf16*)K, (const bf16*)V, (const bf16*)O,
                                                       (const bf16*)dO, lse, dsum, (bf16*)dQ, d, lddo, lddq, dc);
        ick_launch(mha_bwd_dkv_kernel<bf16>, gk, NT, 0, stream)((const bf16*)Q, (const bf16*)K, (const bf16*)V, (const bf16*)dO, lse,
                                                        dsum, (bf16*)dK, (bf16*)dV, d, lddo, lddk, lddv, dc);
    } else {
        ick_set_error("mha_bwd: bad dtype %d", dt);
        return ICK_ERR_UNSUPPORTED;
    }
    return ick_check_launch("mha_bwd");
}

extern "C" int ick_mha_decode(const void* Q, const void* K, const void* V, void* O, int dt, int B, int H, int dh, int ldq,
                              int ldk, int ldv, int ldo, long long kbatch_stride, long long vbatch_stride, int klen,
                              cudaStream_t stream) {
    ICK_REQUIRE(B > 0 && H > 0 && klen > 0 && dh > 0 && dh <= HD, "mha_decode: bad sizes");
    ICK_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && kbatch_stride % 8 == 0 && vbatch_stride % 8 == 0,
                "mha_decode: strides must be multiples of 8");
    const float sl2 = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
    if (dt == ICK_BF16 && (const bf16*)V == (const bf16*)K + H * HD && ldk == ldv && kbatch_stride == vbatch_stride && klen >= 64) {
        // long cached sequences with contiguous K|V rows (the cross-attention over the memory): TMA-staged streaming kernel
        const int rc = ick_mha_decode_tma(Q, K, O, B, H, dh, ldq, ldk, ldo, kbatch_stride, klen, stream);
        if (rc != ICK_ERR_UNSUPPORTED) return rc;
    }
    if (dt == ICK_BF16 && (const bf16*)V == (const bf16*)K + H * HD && ldk == ldv && kbatch_stride == vbatch_stride && H <= 10 &&
        ((((uintptr_t)K) | ((uintptr_t)Q) | ((uintptr_t)O)) & 15) == 0 && decode_rows()) {
        const size_t smem = (size_t)DR_KG * H * 4 * 10 * sizeof(float);
        ick_launch(mha_decode_rows_kernel, B, DR_KG * H * 4, smem, stream)((const bf16*)Q, (const bf16*)K, (bf16*)O, H, dh, ldq, ldk, ldo, kbatch_stride,
                                                                        klen, sl2);
        return ick_check_launch("mha_decode(rows)");
    }
    const int warps = B * H;
    dim3 grid((warps * 32 + 127) / 128);
    if (dt == ICK_F32)
        ick_launch(mha_decode_kernel<float>, grid, 128, 0, stream)((const float*)Q, (const float*)K, (const float*)V, (float*)O, B, H, dh,
                                                           ldq, ldk, ldv, ldo, kbatch_stride, vbatch_stride, klen, sl2);
    else if (dt == ICK_BF16)
        ick_launch(mha_decode_kernel<bf16>, grid, 128, 0, stream)((const bf16*)Q, (const bf16*)K, (const bf16*)V, (bf16*)O, B, H, dh, ldq,
                                                          ldk, ldv, ldo, kbatch_stride, vbatch_stride, klen, sl2);
    else {
        ick_set_error("mha_decode: bad dtype %d", dt);
        return ICK_ERR_UNSUPPORTED;
    }
    return ick_check_launch("mha_decode");
}

extern "C" int ick_mha_decode_beam(const void* Q, const void* K, const void* V, void* O, int dt, int rows, int group, int H, int dh,
                                   int ldq, int ldk, int ldv, int ldo, long long kimg_stride, long long vimg_stride, int klen,
                                   const int* anc, int anc_ld, long long kpos_stride, long long vpos_stride, cudaStream_t stream) {
    ICK_REQUIRE(rows > 0 && group >= 1 && group <= 8 && rows % group == 0, "mha_decode_beam: rows=%d group=%d", rows, group);
    ICK_REQUIRE(H > 0 && H <= 10 && klen > 0 && dh > 0 && dh <= HD, "mha_decode_beam: bad sizes H=%d klen=%d dh=%d", H, klen, dh);
    ICK_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && kimg_stride % 8 == 0 && vimg_stride % 8 == 0 &&
                    kpos_stride % 8 == 0 && vpos_stride % 8 == 0,
                "mha_decode_beam: strides must be multiples of 8");
    ICK_REQUIRE(anc == nullptr || anc_ld >= klen, "mha_decode_beam: ancestor table narrower than klen");
    const uintptr_t al = dt == ICK_F32 ? 31 : 15;
    ICK_REQUIRE(((((uintptr_t)Q) | ((uintptr_t)K) | ((uintptr_t)V) | ((uintptr_t)O)) & al) == 0, "mha_decode_beam: misaligned operand");
    if (dt == ICK_BF16 && anc == nullptr && (const bf16*)V == (const bf16*)K + H * HD && ldk == ldv && kimg_stride == vimg_stride) {
        // shared keys with contiguous K|V rows (the cross-attention over the memory): tensor-core kernel fed by whole-row TMA copies
        const int rc = ick_mha_decode_tma_mma(Q, K, O, rows / group, group, H, dh, ldq, ldk, ldo, kimg_stride, klen, stream);
        if (rc != ICK_ERR_UNSUPPORTED) return rc;
    }
    if (dt == ICK_F32)
        return launch_decode_beam((const float*)Q, (const float*)K, (const float*)V, (float*)O, rows, group, H, dh, ldq, ldk, ldv, ldo,
                                  kimg_stride, vimg_stride, klen, anc, anc_ld, kpos_stride, vpos_stride, stream);
    if (dt == ICK_BF16)
        return launch_decode_beam((const bf16*)Q, (const bf16*)K, (const bf16*)V, (bf16*)O, rows, group, H, dh, ldq, ldk, ldv, ldo,
                                  kimg_stride, vimg_stride, klen, anc, anc_ld, kpos_stride, vpos_stride, stream);
    ick_set_error("mha_decode_beam: bad dtype %d", dt);
    return ICK_ERR_UNSUPPORTED;
}
