// Context-preparation and pointer-head kernels of the caption decoder (all HBM/latency-bound integer + gather work):
//   entity encoder      G/models.py:82-104, K/models.py:82-133, N/models.py:79-134
//   fact encoder        K/models.py:170-188
//   caption embedder    G/models.py:143-181, K/models.py:209-259  (+ sqrt(d) scale and positional table, G:355-357)
//   pixel hand-off      encoder_out.permute(2,0,1) + cat, G/models.py:347-349
//   context indicators  K/models.py:380-418 as a first-mention scan (no B*T host syncs) + predicate gate K:436-437
//   pointer heads       fc_entity / fc_fact as a bilinear form, K/models.py:440-452 (no (T,B,E,300) tensor)
// Row layout everywhere: (batch, position) rows of `ld` elements, logical width D, pad columns [D, ld) zero.
#include <cstdlib>

#include <cuda_fp16.h>

#include "common.cuh"
#include "pointer_internal.h"
#include "ickb200.h"

namespace {

constexpr int FIRST_NONE = 1 << 29;  // "never mentioned" sentinel for first_t / tmin

__device__ __forceinline__ float east_dist(float az) {
    // G/models.py:106-115
    return (az >= -90.f ? fabsf(90.f - az) : 90.f + fabsf(az + 180.f)) / 180.f;
}

// number of facts whose subject is entity e (0 for the last slot <unk_ent>), computed by one warp
__device__ __forceinline__ float fact_count_warp(const long long* facts_b, int F, int e, int E, int lane) {
    if (e == E - 1) return 0.f;
    int c = 0;
    for (int f = lane; f < F; f += 32) c += (facts_b[3 * f + 1] == (long long)e);
    return warp_sum((float)c);
}

// hand-made feature columns + type embedding (before the N-variant name multiply)
__device__ __forceinline__ float ent_base(int variant, const float* er, float cnt, const float* type_row, int nf, int c) {
    if (c >= nf) return type_row[c - nf];
    if (variant == 2) {  // N
        switch (c) {
            case 0: return er[1];
            case 1: return er[2];
            case 2: return er[3];
            case 3: return cnt;
            default: return cnt > 0.f ? 1.f : 0.f;
        }
    }
    switch (c) {
        case 0: return er[1];
        case 1: return fabsf(er[2]) / 180.f;
        case 2: return east_dist(er[2]);
        case 3: return er[3];
        case 4: return cnt;
        default: return cnt > 0.f ? 1.f : 0.f;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) entity_encode_fwd_kernel(const float* __restrict__ ent, const long long* __restrict__ facts,
                                                                const float* __restrict__ type_emb, const T* __restrict__ wemb,
                                                                T* __restrict__ out, int variant, int B, int E, int C, int F,
                                                                int D, int ld, int ldw, int ntypes, int V) {
    ick_pdl_entry();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= B * E) return;
    const int b = row / E, e = row % E;
    const float* er = ent + (size_t)row * C;
    const int nf = variant == 0 ? 4 : (variant == 1 ? 6 : 5);
    float cnt = 0.f;
    if (variant != 0) cnt = fact_count_warp(facts + (size_t)b * F * 3, F, e, E, lane);
    int ty = (int)er[4];
    ty = min(max(ty, 0), ntypes - 1);
    const float* trow = type_emb + (size_t)ty * (D - nf);
    int nm[5] = {0, 0, 0, 0, 0};
    if (variant == 2)
#pragma unroll
        for (int k = 0; k < 5; ++k) nm[k] = min(max((int)er[5 + k], 0), V - 1);
    T* o = out + (size_t)row * ld;
    for (int c = lane; c < ld; c += 32) {
        float v = 0.f;
        if (c < D) {
            v = ent_base(variant, er, cnt, trow, nf, c);
            if (variant == 2) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 5; ++k) a += to_f(wemb[(size_t)nm[k] * ldw + c]);
                v *= a / 5.f;
            }
        }
        o[c] = from_f<T>(v);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) entity_encode_bwd_kernel(const float* __restrict__ dEnt, const float* __restrict__ ent,
                                                                const long long* __restrict__ facts,
                                                                const float* __restrict__ type_emb, const T* __restrict__ wemb,
                                                                float* __restrict__ gflat, int type_off, int word_off, int variant,
                                                                int B, int E, int C, int F, int D, int ld, int ldw, int ntypes,
                                                                int V) {
    ick_pdl_entry();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= B * E) return;
    const int b = row / E, e = row % E;
    const float* er = ent + (size_t)row * C;
    const int nf = variant == 0 ? 4 : (variant == 1 ? 6 : 5);
    int ty = (int)er[4];
    ty = min(max(ty, 0), ntypes - 1);
    const float* g = dEnt + (size_t)row * ld;
    float* dtype_row = gflat + type_off + (size_t)ty * (D - nf);
    if (variant != 2) {
        for (int c = nf + lane; c < D; c += 32) atomicAdd(dtype_row + (c - nf), g[c]);
        return;
    }
    const float cnt = fact_count_warp(facts + (size_t)b * F * 3, F, e, E, lane);
    const float* trow = type_emb + (size_t)ty * (D - nf);
    int nm[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) nm[k] = min(max((int)er[5 + k], 0), V - 1);
    for (int c = lane; c < D; c += 32) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 5; ++k) a += to_f(wemb[(size_t)nm[k] * ldw + c]);
        a /= 5.f;
        const float base = ent_base(2, er, cnt, trow, nf, c);
        const float gv = g[c];
        if (c >= nf) atomicAdd(dtype_row + (c - nf), gv * a);
        const float dn = gv * base / 5.f;
#pragma unroll
        for (int k = 0; k < 5; ++k) atomicAdd(gflat + word_off + (size_t)nm[k] * D + c, dn);
    }
}

// ---- fact encoder ----------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fact_encode_fwd_kernel(const long long* __restrict__ facts, const T* __restrict__ entenc,
                                                              const float* __restrict__ pred_emb, T* __restrict__ out, int B, int E,
                                                              int F, int D, int ld, int NP) {
    ick_pdl_entry();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= B * F) return;
    const int b = row / F;
    const int subj = min(max((int)facts[(size_t)row * 3 + 1], 0), E - 1);
    const int pred = min(max((int)facts[(size_t)row * 3 + 2], 0), NP - 1);
    const T* er = entenc + ((size_t)b * E + subj) * ld;
    const float* pr = pred_emb + (size_t)pred * D;
    T* o = out + (size_t)row * ld;
    for (int c = lane; c < ld; c += 32) o[c] = from_f<T>(c < D ? to_f(er[c]) + pr[c] : 0.f);
}

__global__ void __launch_bounds__(256) fact_encode_bwd_kernel(const float* __restrict__ dFact, const long long* __restrict__ facts,
                                                              float* __restrict__ dEnt, float* __restrict__ gflat, int pred_off,
                                                              int B, int E, int F, int D, int ld, int NP) {
    ick_pdl_entry();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= B * F) return;
    const int b = row / F;
    const int subj = min(max((int)facts[(size_t)row * 3 + 1], 0), E - 1);
    const int pred = min(max((int)facts[(size_t)row * 3 + 2], 0), NP - 1);
    const float* g = dFact + (size_t)row * ld;
    float* de = dEnt + ((size_t)b * E + subj) * ld;
    float* dp = gflat + pred_off + (size_t)pred * D;
    for (int c = lane; c < D; c += 32) {
        const float v = g[c];
        atomicAdd(de + c, v);
        atomicAdd(dp + c, v);
    }
}

// ---- caption embedder ------------------------------------------------------------------------------------------------
struct TokSel {
    int kind;  // 0 word, 1 entity, 2 fact
    int idx;
};
__device__ __forceinline__ TokSel select_token(long long tok, long long mask, int V, int E, int F, int pad) {
    TokSel s;
    if (mask == 1) {
        long long e = tok - V;
        s.kind = 1;
        s.idx = (e < 0 || e >= E) ? E - 1 : (int)e;  // out-of-range -> <unk_ent> (last slot), G/models.py:160
    } else if (mask == 2 && F > 0) {
        long long f = tok - V - E;
        s.kind = 2;
        s.idx = (f < 0 || f >= F) ? F - 1 : (int)f;  // K/models.py:232
    } else {
        s.kind = 0;
        s.idx = (tok >= V || tok < 0) ? pad : (int)tok;  // pointer ids embed as <pad>, G/models.py:165-166
    }
    return s;
}

template <typename T>
__global__ void __launch_bounds__(256) caption_embed_fwd_kernel(const long long* __restrict__ caps, const long long* __restrict__ masks,
                                                                const T* __restrict__ wemb, const T* __restrict__ entenc,
                                                                const T* __restrict__ factenc, const float* __restrict__ pe,
                                                                T* __restrict__ out, int B, int Tstride, int t0, int Tn, int V, int E,
                                                                int F, int D, int ld, int ldw, int pad, float scale, int group,
                                                                DropCfg drop) {
    ick_pdl_entry();
    ick_resolve_seed(drop);
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= B * Tn) return;
    const int b = row / Tn, t = t0 + row % Tn;
    const int bc = b / group;  // `group` consecutive caption rows (the beams of one image) share one entity / fact context
    const TokSel s = select_token(caps[(size_t)b * Tstride + t], masks[(size_t)b * Tstride + t], V, E, F, pad);
    const T* src = s.kind == 0 ? wemb + (size_t)s.idx * ldw
                 : s.kind == 1 ? entenc + ((size_t)bc * E + s.idx) * ld
                               : factenc + ((size_t)bc * F + s.idx) * ld;
    const float* per = pe + (size_t)t * D;
    T* o = out + (size_t)row * ld;
    const uint32_t rmix = ick_rowmix(drop.seed, drop.site, (uint64_t)row);
    for (int c = lane; c < ld; c += 32) {
        float v = 0.f;
        if (c < D) v = (to_f(src[c]) * scale + per[c]) * ick_drop_mul(drop, rmix, (uint32_t)c);
        o[c] = from_f<T>(v);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) caption_embed_bwd_kernel(const T* __restrict__ dX, const long long* __restrict__ caps,
                                                                const long long* __restrict__ masks, float* __restrict__ dEnt,
                                                                float* __restrict__ dFact, float* __restrict__ gflat, int word_off,
                                                                int B, int T_, int V, int E, int F, int D, int ld, int pad, float scale,
                                                                DropCfg drop) {
    ick_pdl_entry();
    ick_resolve_seed(drop);
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= B * T_) return;
    const int b = row / T_;
    const TokSel s = select_token(caps[row], masks[row], V, E, F, pad);
    float* dst = s.kind == 0 ? gflat + word_off + (size_t)s.idx * D
               : s.kind == 1 ? dEnt + ((size_t)b * E + s.idx) * ld
                             : dFact + ((size_t)b * F + s.idx) * ld;
    const T* g = dX + (size_t)row * ld;
    const uint32_t rmix = ick_rowmix(drop.seed, drop.site, (uint64_t)row);
    // Rows behind the end of a caption carry an exactly-zero gradient (their targets are <pad>, ignored by the loss, and the
    // causal decoder lets nothing flow back into them) and they all hit the same <pad> embedding row: skip them instead of
    // serialising thousands of atomic adds of 0.0 on one address.
    constexpr int MAXC = 16;  // columns per lane: supports D <= 512
    float v[MAXC];
    bool any = false;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        v[i] = c < D ? to_f(g[c]) : 0.f;
        any |= v[i] != 0.f;
    }
    if (!__any_sync(0xffffffffu, any)) return;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = lane + 32 * i;
        if (c < D) atomicAdd(dst + c, v[i] * scale * ick_drop_mul(drop, rmix, (uint32_t)c));
    }
}

// ---- pixels: (B, D, P) fp32 channel-major  <->  memory rows (b*M + p, c) ------------------------------------------------
// 64 channels x 32 pixels per CTA: a warp reads 32 consecutive pixels of a channel (128 bytes) and writes 64 consecutive channels of a
// pixel as element pairs (128 bytes in bf16; the 32 x 32 version wrote 64-byte rows of single elements); ld must be even.
template <typename T>
__global__ void __launch_bounds__(256) pixels_fwd_kernel(const float* __restrict__ enc, T* __restrict__ mem, int D, int P, int M, int ld) {
    ick_pdl_entry();
    __shared__ float tile[64][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int i = ty; i < 64; i += 8) {
        const int c = c0 + i, p = p0 + tx;
        tile[i][tx] = (c < D && p < P) ? __ldg(enc + ((size_t)b * D + c) * P + p) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = ty; i < 32; i += 8) {
        const int p = p0 + i, c = c0 + 2 * tx;
        if (p < P && c < ld) st2(mem + ((size_t)b * M + p) * ld + c, c < D ? tile[2 * tx][i] : 0.f, c + 1 < D ? tile[2 * tx + 1][i] : 0.f);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) pixels_bwd_kernel(const T* __restrict__ dmem, float* __restrict__ denc, int D, int P, int M, int ld) {
    ick_pdl_entry();
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int p = p0 + i, c = c0 + tx;
        tile[i][tx] = (p < P && c < D) ? to_f(dmem[((size_t)b * M + p) * ld + c]) : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, p = p0 + tx;
        if (c < D && p < P) denc[((size_t)b * D + c) * P + p] = tile[tx][i];
    }
}

// ---- encoder hand-off: AdaptiveAvgPool2d + layout change for the 1x1 convolution --------------------------------------------------
// Encoder.forward (G/models.py:42-46) pools the ResNet trunk output (B, C, Hin, Win) to (B, C, Hout, Wout) and applies a 1x1
// convolution, i.e. a GEMM over channels.  This kernel writes the pooled features directly as the K-major A operand of that
// GEMM: rows[(b*Hout + oy)*Wout + ox][c] = mean of x[b, c, window(oy), window(ox)], window(o) = [floor(o*In/Out),
// ceil((o+1)*In/Out)) (torch's adaptive pooling).  One CTA per (image, 64 channels): the 64 input planes go through shared
// memory (coalesced plane reads), each output row segment is 64 contiguous channels.
constexpr int PR_C = 64;
template <typename T>
__global__ void __launch_bounds__(256) pool_rows_kernel(const float* __restrict__ x, T* __restrict__ rows, int C, int Hin, int Win, int Hout,
                                                        int Wout, int ldo) {
    ick_pdl_entry();
    extern __shared__ float planes[];  // [PR_C][Hin*Win + 1], then the pooling windows of the Hout*Wout outputs
    const int b = blockIdx.y, c0 = blockIdx.x * PR_C;
    const int npix = Hin * Win, pst = npix + 1, nout = Hout * Wout;
    int4* win = reinterpret_cast<int4*>(planes + ((PR_C * pst + 3) & ~3));  // (first pixel, window width, window height, -)
    float* inv = reinterpret_cast<float*>(win + nout);
    for (int o = threadIdx.x; o < nout; o += blockDim.x) {  // the integer divisions of the window bounds, once per CTA
        const int oy = o / Wout, ox = o % Wout;
        const int y0 = (oy * Hin) / Hout, y1 = ((oy + 1) * Hin + Hout - 1) / Hout;
        const int x0 = (ox * Win) / Wout, x1 = ((ox + 1) * Win + Wout - 1) / Wout;
        win[o] = make_int4(y0 * Win + x0, x1 - x0, y1 - y0, 0);
        inv[o] = 1.0f / (float)((y1 - y0) * (x1 - x0));
    }
    for (int idx = threadIdx.x; idx < PR_C * npix; idx += blockDim.x) {
        const int c = idx / npix, p = idx - c * npix;
        planes[c * pst + p] = c0 + c < C ? x[((size_t)b * C + c0 + c) * npix + p] : 0.f;
    }
    __syncthreads();
    const int c = threadIdx.x % PR_C;
    if (c0 + c >= C) return;
    const float* pl = planes + c * pst;
    for (int o = threadIdx.x / PR_C; o < nout; o += blockDim.x / PR_C) {
        const int4 w = win[o];
        float acc = 0.f;
        for (int yy = 0; yy < w.z; ++yy)
            for (int xx = 0; xx < w.y; ++xx) acc += pl[w.x + yy * Win + xx];
        rows[((size_t)b * nout + o) * ldo + c0 + c] = from_f<T>(acc * inv[o]);
    }
}

// Backward of pool_rows: d rows (B*Hout*Wout, ldo) -> d x (B, C, Hin, Win) fp32.  An input pixel receives, from every output
// cell whose window covers it, that cell's gradient divided by the window size (AdaptiveAvgPool2d backward).  One CTA per
// (64-channel slab, image): the slab of the row gradients is staged in shared memory by coalesced reads along the channel
// axis, then every (channel, pixel) sums over the (contiguous) range of output rows / columns that cover it.
template <typename T>
__global__ void __launch_bounds__(256) pool_rows_bwd_kernel(const T* __restrict__ drows, float* __restrict__ dx, int C, int Hin, int Win,
                                                            int Hout, int Wout, int ldo) {
    ick_pdl_entry();
    extern __shared__ float planes[];  // [nout][PR_C + 1] gradients, then per input row / column: (first cell, last cell) covering it
    const int b = blockIdx.y, c0 = blockIdx.x * PR_C;
    const int npix = Hin * Win, nout = Hout * Wout, cst = PR_C + 1;
    int2* yr = reinterpret_cast<int2*>(planes + ((nout * cst + 1) & ~1));
    int2* xr = yr + Hin;
    for (int i = threadIdx.x; i < Hin + Win; i += blockDim.x) {
        const bool isy = i < Hin;
        const int v = isy ? i : i - Hin, nin = isy ? Hin : Win, no = isy ? Hout : Wout;
        int lo = no, hi = -1;
        for (int o = 0; o < no; ++o) {
            const int a0 = (o * nin) / no, a1 = ((o + 1) * nin + no - 1) / no;
            if (a0 <= v && v < a1) { lo = o < lo ? o : lo; hi = o; }
        }
        (isy ? yr : xr)[v] = make_int2(lo, hi);
    }
    for (int idx = threadIdx.x; idx < nout * PR_C; idx += blockDim.x) {
        const int o = idx / PR_C, c = idx - o * PR_C;
        planes[o * cst + c] = c0 + c < C ? to_f(drows[((size_t)b * nout + o) * ldo + c0 + c]) : 0.f;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < PR_C * npix; idx += blockDim.x) {
        const int c = idx / npix, p = idx - c * npix;
        if (c0 + c >= C) break;
        const int y = p / Win, x = p - y * Win;
        const int2 ry = yr[y], rx = xr[x];
        float acc = 0.f;
        for (int oy = ry.x; oy <= ry.y; ++oy) {
            const int hy = ((oy + 1) * Hin + Hout - 1) / Hout - (oy * Hin) / Hout;
            for (int ox = rx.x; ox <= rx.y; ++ox) {
                const int wx = ((ox + 1) * Win + Wout - 1) / Wout - (ox * Win) / Wout;
                acc += planes[(oy * Wout + ox) * cst + c] / (float)(hy * wx);
            }
        }
        dx[((size_t)b * C + c0 + c) * npix + p] = acc;
    }
}

// ---- image preparation (the input pipeline's device half) ------------------------------------------------------
// The HDF5 files hold images as fp16 (N, 3, H, W) with values in [0, 255] (G/create_input_files.py:99-101, :334-337);
// CaptionDataset.__getitem__ divides by 255 (numpy fp16 array / python float -> fp16 result), converts to fp32
// (G/datasets.py:44) and train.py's transform normalises per channel with the ImageNet mean / std (G/train.py:139-141,
// torchvision Normalize: sub then div in fp32).  Here the raw fp16 batch is copied to the device as stored and one kernel does
//   y = (float(half(float(raw) / 255)) - mean[c]) / std[c]
// - the same roundings, so the fp32 output is bit-identical - written as fp32 / bf16 in NCHW or NHWC (channels-last, the
// layout the cuDNN trunk prefers).  HBM-bound: 2 bytes in, 2-4 bytes out per element; a thread handles 8 pixels of all 3 planes.
template <typename T>
__global__ void __launch_bounds__(256) image_prep_kernel(const __half* __restrict__ raw, T* __restrict__ out, long long npix8, int HW,
                                                         float m0, float m1, float m2, float s0, float s1, float s2, int nhwc) {
    ick_pdl_entry();
    const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
    const int hw8 = HW / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix8; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / hw8;
        const int p = (int)(i % hw8) * 8;
        float v[3][8];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint4 u = *reinterpret_cast<const uint4*>(raw + ((size_t)n * 3 + c) * HW + p);
            const __half* hh = reinterpret_cast<const __half*>(&u);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[c][k] = (__half2float(__float2half_rn(__half2float(hh[k]) / 255.0f)) - mean[c]) / sd[c];
        }
        if (!nhwc) {
#pragma unroll
            for (int c = 0; c < 3; ++c) st8(out + ((size_t)n * 3 + c) * HW + p, v[c]);
        } else {
            float w[24];
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int c = 0; c < 3; ++c) w[3 * k + c] = v[c][k];
            T* o = out + ((size_t)n * HW + p) * 3;
#pragma unroll
            for (int q = 0; q < 3; ++q) st8(o + 8 * q, w + 8 * q);
        }
    }
}

// ---- context indicators ----------------------------------------------------------------------------------------------
// first_t[b,f]: first caption position holding an entity token (value in [V, V+E)) whose slot is the subject of fact f.
// tmin[b,f]   : for the representative (lowest-index) fact of each distinct predicate, the earliest first_t over all
//               facts sharing that predicate; FIRST_NONE for the others (the predicate indicator is a SET of predicates).
__global__ void __launch_bounds__(256) fact_first_mention_kernel(const long long* __restrict__ caps, const long long* __restrict__ facts,
                                                                 int* __restrict__ first_t, int* __restrict__ tmin, int T_, int F,
                                                                 int V, int E, int group, int NP) {
    ick_pdl_entry();
    extern __shared__ int sm[];
    int* sfirst = sm;
    int* spred = sm + F;
    const int b = blockIdx.x;
    const long long* cb = caps + (size_t)b * T_;
    const long long* fb = facts + (size_t)(b / group) * F * 3;  // the `group` beams of an image share its fact list
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        const long long subj = fb[3 * f + 1];
        int first = FIRST_NONE;
        for (int t = 0; t < T_; ++t) {
            const long long tok = cb[t];
            if (tok >= V && tok < (long long)V + E && tok - V == subj) { first = t; break; }
        }
        sfirst[f] = first;
        spred[f] = (int)fb[3 * f + 2];
        first_t[(size_t)b * F + f] = first;
    }
    if (NP > 0) {  // predicate ids are known to lie in [0, NP): two shared tables replace the O(F^2) duplicate search
        int* ptm = sm + 2 * F;   // earliest first mention per predicate
        int* prep = ptm + NP;    // lowest fact index per predicate (its representative)
        for (int p = threadIdx.x; p < NP; p += blockDim.x) {
            ptm[p] = FIRST_NONE;
            prep[p] = 0x7fffffff;
        }
        __syncthreads();
        for (int f = threadIdx.x; f < F; f += blockDim.x) {
            const int p = min(max(spred[f], 0), NP - 1);
            atomicMin(&ptm[p], sfirst[f]);
            atomicMin(&prep[p], f);
        }
        __syncthreads();
        for (int f = threadIdx.x; f < F; f += blockDim.x) {
            const int p = min(max(spred[f], 0), NP - 1);
            tmin[(size_t)b * F + f] = prep[p] == f ? ptm[p] : FIRST_NONE;
        }
        return;
    }
    __syncthreads();
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        const int p = spred[f];
        int tm = FIRST_NONE;
        bool rep = true;
        for (int g = 0; g < F; ++g)
            if (spred[g] == p) {
                tm = min(tm, sfirst[g]);
                if (g < f) rep = false;
            }
        tmin[(size_t)b * F + f] = rep ? tm : FIRST_NONE;
    }
}

// gate[b,t,:] = bias + sum_{f representative, tmin[b,f] < t + lag} WpT[pred_f, :]   (= fc_predicate(predicate_indicator))
// and, fused, hg = h * gate (the input of fc_vocab, K/models.py:437).  lag = 0 teacher-forced, 1 in predict mode.
template <typename T>
__global__ void __launch_bounds__(128) pred_gate_fwd_kernel(const int* __restrict__ tmin, const long long* __restrict__ facts,
                                                            const float* __restrict__ WpT, const float* __restrict__ bias,
                                                            const T* __restrict__ h, T* __restrict__ gate, T* __restrict__ hg, int Tn,
                                                            int t0, int F, int D, int ld, int ldp, int NP, int lag, int group) {
    ick_pdl_entry();
    extern __shared__ int sact[];  // predicate ids of the active facts
    __shared__ int nact;
    const int b = blockIdx.y, tt = blockIdx.x, t = t0 + tt;
    if (threadIdx.x == 0) nact = 0;
    __syncthreads();
    for (int f = threadIdx.x; f < F; f += blockDim.x)
        if (tmin[(size_t)b * F + f] < t + lag) {
            const int p = min(max((int)facts[((size_t)(b / group) * F + f) * 3 + 2], 0), NP - 1);
            sact[atomicAdd(&nact, 1)] = p;
        }
    __syncthreads();
    const size_t row = (size_t)b * Tn + tt;
    for (int c = threadIdx.x; c < ld; c += blockDim.x) {
        float g = 0.f;
        if (c < D) {
            g = bias[c];
            for (int k = 0; k < nact; ++k) g += WpT[(size_t)sact[k] * ldp + c];
        }
        gate[row * ld + c] = from_f<T>(g);
        if (hg) hg[row * ld + c] = from_f<T>(c < D ? to_f(h[row * ld + c]) * g : 0.f);
    }
}

// dG = dHG * h ; dH = dHG * gate
template <typename T>
__global__ void __launch_bounds__(256) gate_mul_bwd_kernel(const T* __restrict__ dHG, const T* __restrict__ h, const T* __restrict__ gate,
                                                           T* __restrict__ dG, T* __restrict__ dH, size_t n) {
    ick_pdl_entry();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float g = to_f(dHG[i]);
        dG[i] = from_f<T>(g * to_f(h[i]));
        dH[i] = from_f<T>(g * to_f(gate[i]));
    }
}

// dWp[c, pred_f] += sum_{t: tmin < t + lag} dG[b,t,c]  (fc_predicate.weight is (D, NP)); one CTA per (b, f)
template <typename T>
__global__ void __launch_bounds__(128) pred_gate_bwd_kernel(const T* __restrict__ dG, const int* __restrict__ tmin,
                                                            const long long* __restrict__ facts, float* __restrict__ gflat, int wp_off,
                                                            int T_, int F, int D, int ld, int NP, int lag) {
    ick_pdl_entry();
    const int b = blockIdx.y, f = blockIdx.x;
    const int tm = tmin[(size_t)b * F + f];
    if (tm >= FIRST_NONE) return;
    const int p = min(max((int)facts[((size_t)b * F + f) * 3 + 2], 0), NP - 1);
    const int tbeg = max(0, tm + 1 - lag);
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // four independent chains: the loop is load-latency bound
        const T* col = dG + (size_t)b * T_ * ld + c;
        int t = tbeg;
        for (; t + 3 < T_; t += 4) {
            s0 += to_f(col[(size_t)t * ld]);
            s1 += to_f(col[(size_t)(t + 1) * ld]);
            s2 += to_f(col[(size_t)(t + 2) * ld]);
            s3 += to_f(col[(size_t)(t + 3) * ld]);
        }
        for (; t < T_; ++t) s0 += to_f(col[(size_t)t * ld]);
        atomicAdd(gflat + wp_off + (size_t)c * NP + p, (s0 + s1) + (s2 + s3));
    }
}

// ---- pointer heads -----------------------------------------------------------------------------------------------------
// scores[b,t,col0+s] = bias + mask(b,t,s) * sum_d h[b,t,d] * w[d] * ctx[b,s,d];   mask = first_t[b,s] < t + lag (facts only)
constexpr int PT_T = 16;  // time steps per CTA (teacher-forced forward); the single-step decode instantiates PT = 1
template <typename T, int PT>
__global__ void __launch_bounds__(128) pointer_fwd_kernel(const T* __restrict__ h, const T* __restrict__ ctx, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const int* __restrict__ first_t,
                                                          float* __restrict__ scores, int Tn, int t0, int S, int D, int ld, int lds,
                                                          int col0, int lag, int beams) {
    // beams != 0: the Tn rows of "image" b are the beams of one image at the SAME time step t0 - they share the ctx rows,
    // each has its own first_t row (its own caption history) and time does not advance along the rows.
    ick_pdl_entry();
    extern __shared__ __align__(16) float hw[];  // [PT][Dp]
    const int Dp = (D + 7) & ~7;
    const int b = blockIdx.z, tt0 = blockIdx.y * PT;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    for (int idx = threadIdx.x; idx < PT * Dp; idx += blockDim.x) {
        const int t = idx / Dp, d = idx % Dp;
        hw[idx] = (tt0 + t < Tn && d < D) ? to_f(h[((size_t)b * Tn + tt0 + t) * ld + d]) * w[d] : 0.f;
    }
    __syncthreads();
    if (s >= S) return;
    float acc[PT];
#pragma unroll
    for (int t = 0; t < PT; ++t) acc[t] = 0.f;
    const T* cr = ctx + ((size_t)b * S + s) * ld;
#pragma unroll 4
    for (int d = 0; d < Dp; d += 8) {
        float c[8];
        ld8(cr + d, c);  // pad columns of ctx rows are zero and hw is zero there too
#pragma unroll
        for (int t = 0; t < PT; ++t) {
            const float4 a0 = *reinterpret_cast<const float4*>(&hw[t * Dp + d]);
            const float4 a1 = *reinterpret_cast<const float4*>(&hw[t * Dp + d + 4]);
            acc[t] += a0.x * c[0] + a0.y * c[1] + a0.z * c[2] + a0.w * c[3] + a1.x * c[4] + a1.y * c[5] + a1.z * c[6] + a1.w * c[7];
        }
    }
    const float bv = bias[0];
    const int ft = (first_t && !beams) ? first_t[(size_t)b * S + s] : -1;
#pragma unroll
    for (int t = 0; t < PT; ++t) {
        if (tt0 + t >= Tn) break;
        float m;
        if (beams) m = (!first_t || first_t[((size_t)b * Tn + tt0 + t) * S + s] < t0 + lag) ? 1.f : 0.f;
        else m = (ft < t0 + tt0 + t + lag) ? 1.f : 0.f;
        scores[((size_t)b * Tn + tt0 + t) * lds + col0 + s] = acc[t] * m + bv;
    }
}

constexpr int PB_D = 32;  // feature columns per CTA in the backward kernels
// dCtx[b,s,d] += w[d] * sum_t m * dS[b,t,col0+s] * h[b,t,d];   dbias += sum m-independent dS   (thread per slot)
template <typename T>
__global__ void __launch_bounds__(128) pointer_bwd_ctx_kernel(const T* __restrict__ dS, const T* __restrict__ h, const float* __restrict__ w,
                                                              const int* __restrict__ first_t, float* __restrict__ dCtx,
                                                              float* __restrict__ gflat, int bias_off, int T_, int S, int D, int ld,
                                                              int ldds, int col0, int lag) {
    ick_pdl_entry();
    __shared__ __align__(16) float hs[32][PB_D];
    __shared__ float red[4];
    const int b = blockIdx.z, d0 = blockIdx.y * PB_D;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = s < S;
    const int ft = (first_t && active) ? first_t[(size_t)b * S + s] : -1;
    float acc[PB_D];
#pragma unroll
    for (int i = 0; i < PB_D; ++i) acc[i] = 0.f;
    float bsum = 0.f;
    for (int tb = 0; tb < T_; tb += 32) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < 32 * PB_D; idx += blockDim.x) {
            const int t = idx / PB_D, d = idx % PB_D;
            hs[t][d] = (tb + t < T_ && d0 + d < D) ? to_f(h[((size_t)b * T_ + tb + t) * ld + d0 + d]) : 0.f;
        }
        __syncthreads();
        if (!active) continue;
        const int nt = min(32, T_ - tb);
        for (int t = 0; t < nt; ++t) {
            const float g = to_f(dS[((size_t)b * T_ + tb + t) * ldds + col0 + s]);
            bsum += g;
            if (ft < tb + t + lag) {
#pragma unroll
                for (int i = 0; i < PB_D; i += 4) {
                    const float4 hv = *reinterpret_cast<const float4*>(&hs[t][i]);
                    acc[i] = fmaf(g, hv.x, acc[i]);
                    acc[i + 1] = fmaf(g, hv.y, acc[i + 1]);
                    acc[i + 2] = fmaf(g, hv.z, acc[i + 2]);
                    acc[i + 3] = fmaf(g, hv.w, acc[i + 3]);
                }
            }
        }
    }
    if (active) {
        float* o = dCtx + ((size_t)b * S + s) * ld + d0;
#pragma unroll
        for (int i = 0; i < PB_D; ++i)
            if (d0 + i < D) o[i] += acc[i] * w[d0 + i];
    }
    if (blockIdx.y == 0) {
        bsum = warp_sum(active ? bsum : 0.f);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = bsum;
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(gflat + bias_off, red[0] + red[1] + red[2] + red[3]);
    }
}

// G[t,d] = sum_s m * dS[b,t,col0+s] * ctx[b,s,d];  dH[b,t,d] += w[d]*G;  dw[d] += sum_{b,t} h[b,t,d]*G   (thread per step)
template <typename T>
__global__ void __launch_bounds__(128) pointer_bwd_h_kernel(const T* __restrict__ dS, const T* __restrict__ h, const T* __restrict__ ctx,
                                                            const float* __restrict__ w, const int* __restrict__ first_t,
                                                            T* __restrict__ dH, float* __restrict__ gflat, int w_off, int T_, int S, int D,
                                                            int ld, int ldds, int col0, int lag) {
    ick_pdl_entry();
    __shared__ __align__(16) float cs[32][PB_D];
    __shared__ int sft[32];
    __shared__ float red[4][PB_D];
    const int b = blockIdx.z, d0 = blockIdx.y * PB_D;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = t < T_;
    float acc[PB_D];
#pragma unroll
    for (int i = 0; i < PB_D; ++i) acc[i] = 0.f;
    for (int sb = 0; sb < S; sb += 32) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < 32 * PB_D; idx += blockDim.x) {
            const int s = idx / PB_D, d = idx % PB_D;
            cs[s][d] = (sb + s < S && d0 + d < D) ? to_f(ctx[((size_t)b * S + sb + s) * ld + d0 + d]) : 0.f;
        }
        if (threadIdx.x < 32) sft[threadIdx.x] = (first_t && sb + threadIdx.x < S) ? first_t[(size_t)b * S + sb + threadIdx.x] : -1;
        __syncthreads();
        if (!active) continue;
        const int ns = min(32, S - sb);
        for (int s = 0; s < ns; ++s) {
            if (!(sft[s] < t + lag)) continue;
            const float g = to_f(dS[((size_t)b * T_ + t) * ldds + col0 + sb + s]);
#pragma unroll
            for (int i = 0; i < PB_D; i += 4) {
                const float4 cv = *reinterpret_cast<const float4*>(&cs[s][i]);
                acc[i] = fmaf(g, cv.x, acc[i]);
                acc[i + 1] = fmaf(g, cv.y, acc[i + 1]);
                acc[i + 2] = fmaf(g, cv.z, acc[i + 2]);
                acc[i + 3] = fmaf(g, cv.w, acc[i + 3]);
            }
        }
    }
    // dH and the per-column dw partial
    float dwp[PB_D];
#pragma unroll
    for (int i = 0; i < PB_D; ++i) dwp[i] = 0.f;
    if (active) {
        const size_t r = ((size_t)b * T_ + t) * ld + d0;
#pragma unroll
        for (int i = 0; i < PB_D; ++i)
            if (d0 + i < D) {
                const float hv = to_f(h[r + i]);
                dwp[i] = hv * acc[i];
                dH[r + i] = from_f<T>(to_f(dH[r + i]) + acc[i] * w[d0 + i]);
            }
    }
#pragma unroll
    for (int i = 0; i < PB_D; ++i) {
        const float v = warp_sum(dwp[i]);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < PB_D && d0 + threadIdx.x < D)
        atomicAdd(gflat + w_off + d0 + threadIdx.x,
                  red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x]);
}

inline int rows_grid(long long rows) { return (int)((rows * 32 + 255) / 256); }

}  // namespace

#define ICK_BAD_DT(name, dt)                       \
    do {                                           \
        ick_set_error(name ": bad dtype %d", (dt)); \
        return ICK_ERR_UNSUPPORTED;                \
    } while (0)

extern "C" int ick_entity_encode_fwd(const float* entities, const long long* facts, const float* type_emb, const void* word_emb,
                                     void* out, int dt, int variant, int B, int E, int C, int F, int D, int ld, int ldw, int ntypes,
                                     int V, cudaStream_t stream) {
    ICK_REQUIRE(variant >= 0 && variant <= 2, "entity_encode_fwd: bad variant %d", variant);
    ICK_REQUIRE(variant == 0 || facts != nullptr, "entity_encode_fwd: variant needs facts");
    ICK_REQUIRE(variant != 2 || (word_emb != nullptr && C >= 10), "entity_encode_fwd: news variant needs word embeddings and 10 columns");
    ICK_REQUIRE(C >= 5 && D <= ld && ntypes > 0, "entity_encode_fwd: bad sizes");
    if (B * E == 0) return ICK_OK;
    const int grid = rows_grid((long long)B * E);
    if (dt == ICK_F32)
        ick_launch(entity_encode_fwd_kernel<float>, grid, 256, 0, stream)(entities, facts, type_emb, (const float*)word_emb, (float*)out, variant, B,
                                                                   E, C, F, D, ld, ldw, ntypes, V);
    else if (dt == ICK_BF16)
        ick_launch(entity_encode_fwd_kernel<bf16>, grid, 256, 0, stream)(entities, facts, type_emb, (const bf16*)word_emb, (bf16*)out, variant, B, E,
                                                                  C, F, D, ld, ldw, ntypes, V);
    else ICK_BAD_DT("entity_encode_fwd", dt);
    return ick_check_launch("entity_encode_fwd");
}

extern "C" int ick_entity_encode_bwd(const float* dEnt, const float* entities, const long long* facts, const float* type_emb,
                                     const void* word_emb, float* gflat, int type_off, int word_off, int dt, int variant, int B, int E,
                                     int C, int F, int D, int ld, int ldw, int ntypes, int V, cudaStream_t stream) {
    ICK_REQUIRE(variant >= 0 && variant <= 2, "entity_encode_bwd: bad variant %d", variant);
    ICK_REQUIRE(variant != 2 || (word_emb != nullptr && facts != nullptr && C >= 10), "entity_encode_bwd: news variant inputs missing");
    if (B * E == 0) return ICK_OK;
    const int grid = rows_grid((long long)B * E);
    if (dt == ICK_F32)
        ick_launch(entity_encode_bwd_kernel<float>, grid, 256, 0, stream)(dEnt, entities, facts, type_emb, (const float*)word_emb, gflat, type_off,
                                                                   word_off, variant, B, E, C, F, D, ld, ldw, ntypes, V);
    else if (dt == ICK_BF16)
        ick_launch(entity_encode_bwd_kernel<bf16>, grid, 256, 0, stream)(dEnt, entities, facts, type_emb, (const bf16*)word_emb, gflat, type_off,
                                                                  word_off, variant, B, E, C, F, D, ld, ldw, ntypes, V);
    else ICK_BAD_DT("entity_encode_bwd", dt);
    return ick_check_launch("entity_encode_bwd");
}

extern "C" int ick_fact_encode_fwd(const long long* facts, const void* ent_enc, const float* pred_emb, void* out, int dt, int B, int E,
                                   int F, int D, int ld, int NP, cudaStream_t stream) {
    ICK_REQUIRE(D <= ld && NP > 0 && E > 0, "fact_encode_fwd: bad sizes");
    if (B * F == 0) return ICK_OK;
    const int grid = rows_grid((long long)B * F);
    if (dt == ICK_F32)
        ick_launch(fact_encode_fwd_kernel<float>, grid, 256, 0, stream)(facts, (const float*)ent_enc, pred_emb, (float*)out, B, E, F, D, ld, NP);
    else if (dt == ICK_BF16)
        ick_launch(fact_encode_fwd_kernel<bf16>, grid, 256, 0, stream)(facts, (const bf16*)ent_enc, pred_emb, (bf16*)out, B, E, F, D, ld, NP);
    else ICK_BAD_DT("fact_encode_fwd", dt);
    return ick_check_launch("fact_encode_fwd");
}

extern "C" int ick_fact_encode_bwd(const float* dFact, const long long* facts, float* dEnt, float* gflat, int pred_off, int B, int E,
                                   int F, int D, int ld, int NP, cudaStream_t stream) {
    if (B * F == 0) return ICK_OK;
    ick_launch(fact_encode_bwd_kernel, rows_grid((long long)B * F), 256, 0, stream)(dFact, facts, dEnt, gflat, pred_off, B, E, F, D, ld, NP);
    return ick_check_launch("fact_encode_bwd");
}

extern "C" int ick_caption_embed_fwd(const long long* captions, const long long* masks, const void* word_emb, const void* ent_enc,
                                     const void* fact_enc, const float* pe, void* out, int dt, int B, int Tstride, int t0, int Tn, int V,
                                     int E, int F, int D, int ld, int ldw, int pad, float scale, int group, float drop_p,
                                     unsigned seed, unsigned site, cudaStream_t stream) {
    ICK_REQUIRE(t0 >= 0 && t0 + Tn <= Tstride && D <= ld, "caption_embed_fwd: bad sizes");
    ICK_REQUIRE(group >= 1 && B % group == 0, "caption_embed_fwd: B=%d is not a multiple of group=%d", B, group);
    ICK_REQUIRE(F == 0 || fact_enc != nullptr, "caption_embed_fwd: facts expected");
    if (B * Tn == 0) return ICK_OK;
    DropCfg dc = make_drop(drop_p, seed, site);
    const int grid = rows_grid((long long)B * Tn);
    if (dt == ICK_F32)
        ick_launch(caption_embed_fwd_kernel<float>, grid, 256, 0, stream)(captions, masks, (const float*)word_emb, (const float*)ent_enc,
                                                                   (const float*)fact_enc, pe, (float*)out, B, Tstride, t0, Tn, V, E, F, D,
                                                                   ld, ldw, pad, scale, group, dc);
    else if (dt == ICK_BF16)
        ick_launch(caption_embed_fwd_kernel<bf16>, grid, 256, 0, stream)(captions, masks, (const bf16*)word_emb, (const bf16*)ent_enc,
                                                                  (const bf16*)fact_enc, pe, (bf16*)out, B, Tstride, t0, Tn, V, E, F, D, ld,
                                                                  ldw, pad, scale, group, dc);
    else ICK_BAD_DT("caption_embed_fwd", dt);
    return ick_check_launch("caption_embed_fwd");
}

extern "C" int ick_caption_embed_bwd(const void* dX, const long long* captions, const long long* masks, float* dEnt, float* dFact,
                                     float* gflat, int word_off, int dt, int B, int T, int V, int E, int F, int D, int ld, int pad,
                                     float scale, float drop_p, unsigned seed, unsigned site, cudaStream_t stream) {
    ICK_REQUIRE(F == 0 || dFact != nullptr, "caption_embed_bwd: dFact expected");
    ICK_REQUIRE(D <= 512, "caption_embed_bwd: D=%d > 512", D);
    if (B * T == 0) return ICK_OK;
    DropCfg dc = make_drop(drop_p, seed, site);
    const int grid = rows_grid((long long)B * T);
    if (dt == ICK_F32)
        ick_launch(caption_embed_bwd_kernel<float>, grid, 256, 0, stream)((const float*)dX, captions, masks, dEnt, dFact, gflat, word_off, B, T, V,
                                                                   E, F, D, ld, pad, scale, dc);
    else if (dt == ICK_BF16)
        ick_launch(caption_embed_bwd_kernel<bf16>, grid, 256, 0, stream)((const bf16*)dX, captions, masks, dEnt, dFact, gflat, word_off, B, T, V, E,
                                                                  F, D, ld, pad, scale, dc);
    else ICK_BAD_DT("caption_embed_bwd", dt);
    return ick_check_launch("caption_embed_bwd");
}

extern "C" int ick_pixels_fwd(const float* encoder_out, void* memory, int dt, int B, int D, int P, int M, int ld, cudaStream_t stream) {
    ICK_REQUIRE(P <= M && D <= ld, "pixels_fwd: bad sizes");
    if (B * P == 0) return ICK_OK;
    ICK_REQUIRE(ld % 2 == 0, "pixels_fwd: ld must be even");
    dim3 grid((P + 31) / 32, (ld + 63) / 64, B);
    if (dt == ICK_F32) ick_launch(pixels_fwd_kernel<float>, grid, 256, 0, stream)(encoder_out, (float*)memory, D, P, M, ld);
    else if (dt == ICK_BF16) ick_launch(pixels_fwd_kernel<bf16>, grid, 256, 0, stream)(encoder_out, (bf16*)memory, D, P, M, ld);
    else ICK_BAD_DT("pixels_fwd", dt);
    return ick_check_launch("pixels_fwd");
}

extern "C" int ick_pixels_bwd(const void* dmemory, float* d_encoder_out, int dt, int B, int D, int P, int M, int ld, cudaStream_t stream) {
    ICK_REQUIRE(P <= M && D <= ld, "pixels_bwd: bad sizes");
    if (B * P == 0) return ICK_OK;
    dim3 grid((P + 31) / 32, (D + 31) / 32, B);
    if (dt == ICK_F32) ick_launch(pixels_bwd_kernel<float>, grid, 256, 0, stream)((const float*)dmemory, d_encoder_out, D, P, M, ld);
    else if (dt == ICK_BF16) ick_launch(pixels_bwd_kernel<bf16>, grid, 256, 0, stream)((const bf16*)dmemory, d_encoder_out, D, P, M, ld);
    else ICK_BAD_DT("pixels_bwd", dt);
    return ick_check_launch("pixels_bwd");
}

extern "C" int ick_fact_first_mention(const long long* captions, const long long* facts, int* first_t, int* tmin, int B, int T, int F,
                                      int V, int E, int group, int NP, cudaStream_t stream) {
    ICK_REQUIRE(F > 0 && F <= 4096, "fact_first_mention: F=%d out of range", F);
    ICK_REQUIRE(group >= 1 && B % group == 0, "fact_first_mention: B=%d is not a multiple of group=%d", B, group);
    if (B == 0) return ICK_OK;
    if (NP < 0 || (size_t)(2 * F + 2 * NP) * sizeof(int) > 46 * 1024) NP = 0;  // tables do not fit: pairwise search
    ick_launch(fact_first_mention_kernel, B, 256, (size_t)(2 * F + 2 * NP) * sizeof(int), stream)(captions, facts, first_t, tmin, T, F, V, E, group,
                                                                                                 NP);
    return ick_check_launch("fact_first_mention");
}

extern "C" int ick_pred_gate_fwd(const int* tmin, const long long* facts, const float* WpT, const float* bias, const void* h, void* gate,
                                 void* hg, int dt, int B, int Tn, int t0, int F, int D, int ld, int ldp, int NP, int lag, int group,
                                 cudaStream_t stream) {
    ICK_REQUIRE(F > 0 && F <= 4096 && D <= ld && D <= ldp, "pred_gate_fwd: bad sizes");
    ICK_REQUIRE(group >= 1 && B % group == 0, "pred_gate_fwd: B=%d is not a multiple of group=%d", B, group);
    ICK_REQUIRE(hg == nullptr || h != nullptr, "pred_gate_fwd: hg needs h");
    if (B * Tn == 0) return ICK_OK;
    dim3 grid(Tn, B);
    if (dt == ICK_F32)
        ick_launch(pred_gate_fwd_kernel<float>, grid, 128, F * sizeof(int), stream)(tmin, facts, WpT, bias, (const float*)h, (float*)gate, (float*)hg,
                                                                            Tn, t0, F, D, ld, ldp, NP, lag, group);
    else if (dt == ICK_BF16)
        ick_launch(pred_gate_fwd_kernel<bf16>, grid, 128, F * sizeof(int), stream)(tmin, facts, WpT, bias, (const bf16*)h, (bf16*)gate, (bf16*)hg, Tn,
                                                                           t0, F, D, ld, ldp, NP, lag, group);
    else ICK_BAD_DT("pred_gate_fwd", dt);
    return ick_check_launch("pred_gate_fwd");
}

extern "C" int ick_gate_mul_bwd(const void* dHG, const void* h, const void* gate, void* dG, void* dH, int dt, long long n,
                                cudaStream_t stream) {
    if (n == 0) return ICK_OK;
    const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    if (dt == ICK_F32)
        ick_launch(gate_mul_bwd_kernel<float>, grid, 256, 0, stream)((const float*)dHG, (const float*)h, (const float*)gate, (float*)dG, (float*)dH, (size_t)n);
    else if (dt == ICK_BF16)
        ick_launch(gate_mul_bwd_kernel<bf16>, grid, 256, 0, stream)((const bf16*)dHG, (const bf16*)h, (const bf16*)gate, (bf16*)dG, (bf16*)dH, (size_t)n);
    else ICK_BAD_DT("gate_mul_bwd", dt);
    return ick_check_launch("gate_mul_bwd");
}

extern "C" int ick_pred_gate_bwd(const void* dG, const int* tmin, const long long* facts, float* gflat, int wp_off, int dt, int B, int T,
                                 int F, int D, int ld, int NP, int lag, cudaStream_t stream) {
    if (B * F == 0) return ICK_OK;
    dim3 grid(F, B);
    if (dt == ICK_F32)
        ick_launch(pred_gate_bwd_kernel<float>, grid, 128, 0, stream)((const float*)dG, tmin, facts, gflat, wp_off, T, F, D, ld, NP, lag);
    else if (dt == ICK_BF16)
        ick_launch(pred_gate_bwd_kernel<bf16>, grid, 128, 0, stream)((const bf16*)dG, tmin, facts, gflat, wp_off, T, F, D, ld, NP, lag);
    else ICK_BAD_DT("pred_gate_bwd", dt);
    return ick_check_launch("pred_gate_bwd");
}

static bool ptr_decode_mma() {  // ICK_PTR_DECODE_MMA=0: single-step pointer scores on the CUDA-core kernel (A/B aid)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ICK_PTR_DECODE_MMA");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

extern "C" int ick_pointer_fwd(const void* h, const void* ctx, const float* w, const float* bias, const int* first_t, float* scores,
                               int dt, int B, int Tn, int t0, int S, int D, int ld, int ldscores, int col0, int lag, int group,
                               cudaStream_t stream) {
    ICK_REQUIRE(D <= ld && ld % 8 == 0 && ((D + 7) & ~7) <= ld, "pointer_fwd: bad sizes D=%d ld=%d", D, ld);
    ICK_REQUIRE(group >= 1 && B % group == 0, "pointer_fwd: B=%d is not a multiple of group=%d", B, group);
    if (B * Tn * S == 0) return ICK_OK;
    if (group > 1) {  // beam decode: `group` consecutive rows are the beams of one image (one time step each)
        ICK_REQUIRE(Tn == 1 && group <= 8, "pointer_fwd: group=%d needs Tn == 1 and group <= 8", group);
        if (dt == ICK_BF16) {  // per-image (group x S x D) GEMM on the tensor cores: the ctx rows are staged with coalesced 16-byte loads
            const int rc = ick_pointer_fwd_mma(h, ctx, w, bias, first_t, scores, B / group, group, t0, S, D, ld, ldscores, col0, lag, 1, stream);
            if (rc != ICK_ERR_UNSUPPORTED) return rc;
        }
        const int Dp8 = (D + 7) & ~7;
        const size_t smemb = (size_t)8 * Dp8 * sizeof(float);
        dim3 gridb((S + 127) / 128, 1, B / group);
        if (dt == ICK_F32)
            ick_launch(pointer_fwd_kernel<float, 8>, gridb, 128, smemb, stream)((const float*)h, (const float*)ctx, w, bias, first_t, scores, group,
                                                                           t0, S, D, ld, ldscores, col0, lag, 1);
        else if (dt == ICK_BF16)
            ick_launch(pointer_fwd_kernel<bf16, 8>, gridb, 128, smemb, stream)((const bf16*)h, (const bf16*)ctx, w, bias, first_t, scores, group, t0,
                                                                          S, D, ld, ldscores, col0, lag, 1);
        else ICK_BAD_DT("pointer_fwd", dt);
        return ick_check_launch("pointer_fwd");
    }
    if (dt == ICK_BF16 && Tn >= 16) {  // tensor-core path (teacher-forced forward); single-step decode stays on the CUDA cores
        const int rc = ick_pointer_fwd_mma(h, ctx, w, bias, first_t, scores, B, Tn, t0, S, D, ld, ldscores, col0, lag, 0, stream);
        if (rc != ICK_ERR_UNSUPPORTED) return rc;
    }
    const int Dp = (D + 7) & ~7;
    if (Tn == 1 && dt == ICK_BF16 && ptr_decode_mma()) {  // greedy decode, bf16: same tensor-core kernel, ctx rows staged coalesced
        const int rc = ick_pointer_fwd_mma(h, ctx, w, bias, first_t, scores, B, Tn, t0, S, D, ld, ldscores, col0, lag, 0, stream);
        if (rc != ICK_ERR_UNSUPPORTED) return rc;
    }
    if (Tn == 1) {  // greedy decode: one time step per image - no 16-step register tile, the ctx rows stream through
        const size_t smem1 = (size_t)Dp * sizeof(float);
        dim3 grid1((S + 127) / 128, 1, B);
        if (dt == ICK_F32)
            ick_launch(pointer_fwd_kernel<float, 1>, grid1, 128, smem1, stream)((const float*)h, (const float*)ctx, w, bias, first_t, scores, Tn, t0, S,
                                                                          D, ld, ldscores, col0, lag, 0);
        else if (dt == ICK_BF16)
            ick_launch(pointer_fwd_kernel<bf16, 1>, grid1, 128, smem1, stream)((const bf16*)h, (const bf16*)ctx, w, bias, first_t, scores, Tn, t0, S, D,
                                                                         ld, ldscores, col0, lag, 0);
        else ICK_BAD_DT("pointer_fwd", dt);
        return ick_check_launch("pointer_fwd");
    }
    const size_t smem = (size_t)PT_T * Dp * sizeof(float);
    dim3 grid((S + 127) / 128, (Tn + PT_T - 1) / PT_T, B);
    if (dt == ICK_F32)
        ick_launch(pointer_fwd_kernel<float, PT_T>, grid, 128, smem, stream)((const float*)h, (const float*)ctx, w, bias, first_t, scores, Tn, t0, S, D,
                                                                       ld, ldscores, col0, lag, 0);
    else if (dt == ICK_BF16)
        ick_launch(pointer_fwd_kernel<bf16, PT_T>, grid, 128, smem, stream)((const bf16*)h, (const bf16*)ctx, w, bias, first_t, scores, Tn, t0, S, D,
                                                                      ld, ldscores, col0, lag, 0);
    else ICK_BAD_DT("pointer_fwd", dt);
    return ick_check_launch("pointer_fwd");
}

extern "C" int ick_pointer_bwd(const void* dS, const void* h, const void* ctx, const float* w, const int* first_t, float* dCtx, void* dH,
                               float* gflat, int w_off, int bias_off, int dt, int B, int T, int S, int D, int ld, int ldds, int col0,
                               int lag, cudaStream_t stream) {
    ICK_REQUIRE(D <= ld, "pointer_bwd: bad sizes");
    if (B * T * S == 0) return ICK_OK;
    if (dt == ICK_BF16 && T >= 16) {
        const int rc = ick_pointer_bwd_mma(dS, h, ctx, w, first_t, dCtx, dH, gflat, w_off, bias_off, B, T, S, D, ld, ldds, col0, lag, stream);
        if (rc != ICK_ERR_UNSUPPORTED) return rc;
    }
    dim3 g1((S + 127) / 128, (D + PB_D - 1) / PB_D, B), g2((T + 127) / 128, (D + PB_D - 1) / PB_D, B);
    if (dt == ICK_F32) {
        ick_launch(pointer_bwd_ctx_kernel<float>, g1, 128, 0, stream)((const float*)dS, (const float*)h, w, first_t, dCtx, gflat, bias_off, T, S, D,
                                                              ld, ldds, col0, lag);
        ick_launch(pointer_bwd_h_kernel<float>, g2, 128, 0, stream)((const float*)dS, (const float*)h, (const float*)ctx, w, first_t, (float*)dH,
                                                            gflat, w_off, T, S, D, ld, ldds, col0, lag);
    } else if (dt == ICK_BF16) {
        ick_launch(pointer_bwd_ctx_kernel<bf16>, g1, 128, 0, stream)((const bf16*)dS, (const bf16*)h, w, first_t, dCtx, gflat, bias_off, T, S, D, ld,
                                                             ldds, col0, lag);
        ick_launch(pointer_bwd_h_kernel<bf16>, g2, 128, 0, stream)((const bf16*)dS, (const bf16*)h, (const bf16*)ctx, w, first_t, (bf16*)dH, gflat,
                                                           w_off, T, S, D, ld, ldds, col0, lag);
    } else ICK_BAD_DT("pointer_bwd", dt);
    return ick_check_launch("pointer_bwd");
}

extern "C" int ick_pool_rows_fwd(const float* x, void* rows, int dt, int B, int C, int Hin, int Win, int Hout, int Wout, int ldo,
                                 cudaStream_t stream) {
    ICK_REQUIRE(B >= 0 && C > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && ldo >= C, "pool_rows_fwd: bad sizes");
    const size_t smem = (size_t)((PR_C * (Hin * Win + 1) + 3) & ~3) * sizeof(float) + (size_t)Hout * Wout * (sizeof(int4) + sizeof(float));
    ICK_REQUIRE(smem <= 48 * 1024, "pool_rows_fwd: input planes of %d x %d do not fit the staging buffer", Hin, Win);
    if (B == 0) return ICK_OK;
    dim3 grid((C + PR_C - 1) / PR_C, B);
    if (dt == ICK_F32) ick_launch(pool_rows_kernel<float>, grid, 256, smem, stream)(x, (float*)rows, C, Hin, Win, Hout, Wout, ldo);
    else if (dt == ICK_BF16) ick_launch(pool_rows_kernel<bf16>, grid, 256, smem, stream)(x, (bf16*)rows, C, Hin, Win, Hout, Wout, ldo);
    else ICK_BAD_DT("pool_rows_fwd", dt);
    return ick_check_launch("pool_rows_fwd");
}

extern "C" int ick_pool_rows_bwd(const void* drows, float* dx, int dt, int B, int C, int Hin, int Win, int Hout, int Wout, int ldo,
                                 cudaStream_t stream) {
    ICK_REQUIRE(B >= 0 && C > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && ldo >= C, "pool_rows_bwd: bad sizes");
    const size_t smem = (size_t)((Hout * Wout * (PR_C + 1) + 1) & ~1) * sizeof(float) + (size_t)(Hin + Win) * sizeof(int2);
    ICK_REQUIRE(smem <= 100 * 1024, "pool_rows_bwd: %d x %d output cells do not fit the staging buffer", Hout, Wout);
    if (B == 0) return ICK_OK;
    dim3 grid((C + PR_C - 1) / PR_C, B);
    if (dt == ICK_F32) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(pool_rows_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ick_launch(pool_rows_bwd_kernel<float>, grid, 256, smem, stream)((const float*)drows, dx, C, Hin, Win, Hout, Wout, ldo);
    } else if (dt == ICK_BF16) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(pool_rows_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ick_launch(pool_rows_bwd_kernel<bf16>, grid, 256, smem, stream)((const bf16*)drows, dx, C, Hin, Win, Hout, Wout, ldo);
    } else ICK_BAD_DT("pool_rows_bwd", dt);
    return ick_check_launch("pool_rows_bwd");
}

extern "C" int ick_image_prep(const void* raw_f16, void* out, int dt, long long N, int C, int HW, const float* mean, const float* stdev,
                              int channels_last, cudaStream_t stream) {
    ICK_REQUIRE(C == 3 && HW > 0 && HW % 8 == 0 && N >= 0, "image_prep: needs 3 channels and H*W a multiple of 8 (C=%d, HW=%d)", C, HW);
    ICK_REQUIRE((((uintptr_t)raw_f16 | (uintptr_t)out) & 31) == 0, "image_prep: buffers must be 32-byte aligned");
    if (N == 0) return ICK_OK;
    const long long npix8 = N * (HW / 8);
    const long long want = (npix8 + 255) / 256;
    const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
    if (dt == ICK_F32)
        ick_launch(image_prep_kernel<float>, grid, 256, 0, stream)((const __half*)raw_f16, (float*)out, npix8, HW, mean[0], mean[1], mean[2], stdev[0],
                                                           stdev[1], stdev[2], channels_last);
    else if (dt == ICK_BF16)
        ick_launch(image_prep_kernel<bf16>, grid, 256, 0, stream)((const __half*)raw_f16, (bf16*)out, npix8, HW, mean[0], mean[1], mean[2], stdev[0],
                                                          stdev[1], stdev[2], channels_last);
    else ICK_BAD_DT("image_prep", dt);
    return ick_check_launch("image_prep");
}
