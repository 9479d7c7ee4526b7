// Loss, optimizer, packing and decode-selection kernels.
//   masked cross-entropy  = pack_padded_sequence + CrossEntropyLoss(ignore_index=<pad>)      G/train.py:275-281
//   clamp + Adam + repack = ut.clip_gradient (G/utils.py:75-85) + torch.optim.Adam step      G/train.py:287-292
//   greedy select         = argmax / top-2 / repetition clean-up state machine of predict()  G/models.py:409-442
#include <stdlib.h>

#include "common.cuh"
#include "ickb200.h"

namespace {

__device__ __forceinline__ void online_merge(float& m, float& l, float m2, float l2) {
    const float mn = fmaxf(m, m2);
    if (mn == -INFINITY) return;
    l = l * __expf(m - mn) + l2 * __expf(m2 - mn);
    m = mn;
}

// One CTA per (b,t) row.  Valid rows: t < decode_len[b] and target = captions[b,t+1] != pad.
// loss_acc[0] += sum(lse - s[target]); loss_acc[1] += #valid.  dS (optional) = softmax - onehot (UNSCALED: the 1/N of
// the mean is folded into the optimizer step, which also makes multi-GPU normalisation exact); zeros on invalid rows.
template <typename TD>
__global__ void __launch_bounds__(256) ce_kernel(const float* __restrict__ scores, const long long* __restrict__ caps,
                                                 const int* __restrict__ decode_len, float* __restrict__ loss_acc, TD* __restrict__ dS,
                                                 int T_, int W, int lds, int ldd, int pad) {
    ick_pdl_entry();
    __shared__ float sm[8], sl[8];
    __shared__ float s_lse;
    const int row = blockIdx.x, b = row / T_, t = row % T_;
    const float* s = scores + (size_t)row * lds;
    TD* d = dS ? dS + (size_t)row * ldd : nullptr;
    long long target = pad;
    if (t < decode_len[b] && t + 1 < T_) target = caps[(size_t)b * T_ + t + 1];
    const bool valid = target != pad && target >= 0 && target < W;
    if (!valid) {
        if (d)
            for (int c = threadIdx.x; c < ldd; c += blockDim.x) d[c] = from_f<TD>(0.f);
        return;
    }
    float m = -INFINITY, l = 0.f;
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        const float x = s[c];
        const float mn = fmaxf(m, x);
        l = l * __expf(m - mn) + __expf(x - mn);
        m = mn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
        online_merge(m, l, m2, l2);
    }
    if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; sl[threadIdx.x >> 5] = l; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float mm = sm[0], ll = sl[0];
        for (int w = 1; w < 8; ++w) online_merge(mm, ll, sm[w], sl[w]);
        const float lse = mm + logf(ll);
        s_lse = lse;
        atomicAdd(loss_acc, lse - s[target]);
        atomicAdd(loss_acc + 1, 1.f);
    }
    __syncthreads();
    if (d) {
        const float lse = s_lse;
        for (int c = threadIdx.x; c < ldd; c += blockDim.x) {
            float g = 0.f;
            if (c < W) g = __expf(s[c] - lse) - (c == (int)target ? 1.f : 0.f);
            d[c] = from_f<TD>(g);
        }
    }
}

// Register-resident variant for rows of at most CE_NT * 2 * CE_NPP columns with even W / lds / ldd (the K and N score widths): a thread
// owns CE_NPP column PAIRS (8-byte loads, all in flight at once), the row maximum and the sum of exponentials are two passes over
// registers with one block barrier each, and the gradient row is written from the same registers as bf16x2 / float2 - no second read
// of the scores.  The scalar kernel above runs a dependent online-softmax chain per element (two exponentials each) and re-reads the
// row for the gradient: 148 us for 13 056 x 10 352 scores, 17 instructions per element (ncu: issue slots 47 %, DRAM 32 %).
constexpr int CE_NT = 512, CE_NPP = 11;
template <typename TD>
__global__ void __launch_bounds__(CE_NT) ce_reg_kernel(const float* __restrict__ scores, const long long* __restrict__ caps,
                                                       const int* __restrict__ decode_len, float* __restrict__ loss_acc, TD* __restrict__ dS,
                                                       int T_, int W, int lds, int ldd, int pad) {
    ick_pdl_entry();
    constexpr int NW = CE_NT / 32;
    __shared__ float red[NW];
    __shared__ float s_bc;
    const int row = blockIdx.x, b = row / T_, t = row % T_, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* s = scores + (size_t)row * lds;
    TD* d = dS ? dS + (size_t)row * ldd : nullptr;
    long long target = pad;
    if (t < decode_len[b] && t + 1 < T_) target = caps[(size_t)b * T_ + t + 1];
    const bool valid = target != pad && target >= 0 && target < W;
    if (!valid) {
        if (d)
            for (int c = 2 * tid; c < ldd; c += 2 * CE_NT) st2(d + c, 0.f, 0.f);
        return;
    }
    float2 x[CE_NPP];
#pragma unroll
    for (int u = 0; u < CE_NPP; ++u) {
        const int c = 2 * (tid + u * CE_NT);
        x[u] = c < W ? __ldg(reinterpret_cast<const float2*>(s + c)) : make_float2(-INFINITY, -INFINITY);
    }
    float m = -INFINITY;
#pragma unroll
    for (int u = 0; u < CE_NPP; ++u) m = fmaxf(m, fmaxf(x[u].x, x[u].y));
    m = warp_max(m);
    if (lane == 0) red[wid] = m;
    __syncthreads();
    m = lane < NW ? red[lane] : -INFINITY;
    m = warp_max(m);  // every warp folds the NW maxima itself
    float l = 0.f;
#pragma unroll
    for (int u = 0; u < CE_NPP; ++u) {
        x[u].x = __expf(x[u].x - m);  // exp(-inf) = 0 for the slots past the row
        x[u].y = __expf(x[u].y - m);
        l += x[u].x + x[u].y;
    }
    l = warp_sum(l);
    __syncthreads();  // everybody has read red[]
    if (lane == 0) red[wid] = l;
    __syncthreads();
    l = lane < NW ? red[lane] : 0.f;
    l = warp_sum(l);
    if (tid == 0) {
        atomicAdd(loss_acc, (m + logf(l)) - s[target]);
        atomicAdd(loss_acc + 1, 1.f);
    }
    if (d) {
        const float inv = 1.f / l;
        const int tg = (int)target;
#pragma unroll
        for (int u = 0; u < CE_NPP; ++u) {
            const int c = 2 * (tid + u * CE_NT);
            if (c < ldd) {
                float g0 = c < W ? x[u].x * inv : 0.f, g1 = c + 1 < W ? x[u].y * inv : 0.f;
                if (c == tg) g0 -= 1.f;
                if (c + 1 == tg) g1 -= 1.f;
                st2(d + c, g0, g1);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                   float bc1, float bc2_sqrt, float clip, const float* __restrict__ count,
                                                   float grad_scale, const int* __restrict__ dstA, const int* __restrict__ dstB,
                                                   const int* __restrict__ dstC, T* __restrict__ packT, float* __restrict__ packF,
                                                   int update, const int* __restrict__ step_dev, const float* __restrict__ lr_dev) {
    ick_pdl_entry();
    float gs = grad_scale;
    if (count) gs /= fmaxf(count[0], 1.f);
    if (step_dev) {  // bias corrections from a device-resident step counter (CUDA-graph replay)
        const float t = (float)step_dev[0];
        bc1 = 1.f - powf(b1, t);
        bc2_sqrt = sqrtf(1.f - powf(b2, t));
    }
    if (lr_dev) lr = lr_dev[0];
    const float step = lr / bc1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float pv = p[i];
        if (update) {
            float gv = g[i] * gs;
            if (clip > 0.f) gv = fminf(fmaxf(gv, -clip), clip);
            const float mv = b1 * m[i] + (1.f - b1) * gv;
            const float vv = b2 * v[i] + (1.f - b2) * gv * gv;
            m[i] = mv;
            v[i] = vv;
            pv -= step * mv / (sqrtf(vv) / bc2_sqrt + eps);
            p[i] = pv;
        }
        if (dstA) { const int a = dstA[i]; if (a >= 0) packT[a] = from_f<T>(pv); }
        if (dstB) { const int a = dstB[i]; if (a >= 0) packT[a] = from_f<T>(pv); }
        if (dstC) { const int a = dstC[i]; if (a >= 0) packF[a] = pv; }
    }
}

// Same update, four consecutive parameters per thread: 16-byte loads / stores of p, g, m, v and of the three index arrays (all
// buffers 16-byte aligned, checked by the launcher), every load of a thread's group issued before the first use.  The scalar
// kernel ran at 3.4 TB/s of DRAM traffic on 15.2 M parameters (profiles/r02c_launches.csv).
template <typename T>
__global__ void __launch_bounds__(256) adam_kernel_v4(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                      float* __restrict__ v, long long n4, float lr, float b1, float b2, float eps,
                                                      float bc1, float bc2_sqrt, float clip, const float* __restrict__ count,
                                                      float grad_scale, const int* __restrict__ dstA, const int* __restrict__ dstB,
                                                      const int* __restrict__ dstC, T* __restrict__ packT, float* __restrict__ packF,
                                                      int update, const int* __restrict__ step_dev, const float* __restrict__ lr_dev) {
    ick_pdl_entry();
    float gs = grad_scale;
    if (count) gs /= fmaxf(count[0], 1.f);
    if (step_dev) {
        const float t = (float)step_dev[0];
        bc1 = 1.f - powf(b1, t);
        bc2_sqrt = sqrtf(1.f - powf(b2, t));
    }
    if (lr_dev) lr = lr_dev[0];
    const float step = lr / bc1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pv4 = reinterpret_cast<const float4*>(p)[i];
        int4 a4 = make_int4(-1, -1, -1, -1), b4 = a4, c4 = a4;
        if (dstA) a4 = __ldg(reinterpret_cast<const int4*>(dstA) + i);
        if (dstB) b4 = __ldg(reinterpret_cast<const int4*>(dstB) + i);
        if (dstC) c4 = __ldg(reinterpret_cast<const int4*>(dstC) + i);
        float pv[4] = {pv4.x, pv4.y, pv4.z, pv4.w};
        if (update) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
            const float4 m4 = reinterpret_cast<const float4*>(m)[i];
            const float4 v4 = reinterpret_cast<const float4*>(v)[i];
            const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
            float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float gv = gg[j] * gs;
                if (clip > 0.f) gv = fminf(fmaxf(gv, -clip), clip);
                mm[j] = b1 * mm[j] + (1.f - b1) * gv;
                vv[j] = b2 * vv[j] + (1.f - b2) * gv * gv;
                pv[j] -= step * mm[j] / (sqrtf(vv[j]) / bc2_sqrt + eps);
            }
            reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
            reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
            reinterpret_cast<float4*>(p)[i] = make_float4(pv[0], pv[1], pv[2], pv[3]);
        }
        const int aa[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w}, cc[4] = {c4.x, c4.y, c4.z, c4.w};
        // a parameter row is contiguous in the K-major copy: four consecutive destinations, 8-byte aligned, become one store
        if (aa[0] >= 0 && (aa[0] & 3) == 0 && aa[1] == aa[0] + 1 && aa[2] == aa[0] + 2 && aa[3] == aa[0] + 3 && sizeof(T) == 2) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(pv[0], pv[1]), hi = __floats2bfloat162_rn(pv[2], pv[3]);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&lo);
            u.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(packT + aa[0]) = u;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (aa[j] >= 0) packT[aa[j]] = from_f<T>(pv[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (bb[j] >= 0) packT[bb[j]] = from_f<T>(pv[j]);
            if (cc[j] >= 0) packF[cc[j]] = pv[j];
        }
    }
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) cast2d_kernel(const TS* __restrict__ src, TD* __restrict__ dst, long long rows, int cols, int lds,
                                                     int ldd) {
    ick_pdl_entry();
    const long long total = rows * ldd;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / ldd;
        const int c = (int)(i % ldd);
        dst[i] = from_f<TD>(c < cols ? to_f(src[r * lds + c]) : 0.f);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) accum_f32_kernel(const T* __restrict__ src, float* __restrict__ dst, long long n) {
    ick_pdl_entry();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] += to_f(src[i]);
}
// eight elements per thread (16-byte source loads for bf16, two 16-byte read-modify-writes of the fp32 destination)
template <typename T>
__global__ void __launch_bounds__(256) accum_f32_v8_kernel(const T* __restrict__ src, float* __restrict__ dst, long long n8) {
    ick_pdl_entry();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float a[8], b[8];
        ld8(src + 8 * i, a);
        ld8(dst + 8 * i, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] += a[j];
        st8(dst + 8 * i, b);
    }
}

// out[c] += sum_r x[r, c]; grid-stride over row blocks, columns across threads
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, float* __restrict__ out, long long rows, int cols, int ld,
                                                     int rows_per_block) {
    ick_pdl_entry();
    const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // four rows in flight (one dependent add chain per column was latency-bound)
        long long r = r0;
        for (; r + 3 < r1; r += 4) {
            s0 += to_f(x[r * ld + c]);
            s1 += to_f(x[(r + 1) * ld + c]);
            s2 += to_f(x[(r + 2) * ld + c]);
            s3 += to_f(x[(r + 3) * ld + c]);
        }
        for (; r < r1; ++r) s0 += to_f(x[r * ld + c]);
        atomicAdd(out + c, (s0 + s1) + (s2 + s3));
    }
}

// ---- greedy decode selection ------------------------------------------------------------------------------------------
struct Top2 {
    float v1, v2;
    int i1, i2;
};
__device__ __forceinline__ void top2_push(Top2& t, float v, int i) {
    // ties keep the lower index first (torch argmax returns the first maximal index)
    if (v > t.v1 || (v == t.v1 && i < t.i1)) {
        t.v2 = t.v1; t.i2 = t.i1; t.v1 = v; t.i1 = i;
    } else if (v > t.v2 || (v == t.v2 && i < t.i2)) {
        t.v2 = v; t.i2 = i;
    }
}
__device__ __forceinline__ void top2_merge(Top2& a, const Top2& b) {
    top2_push(a, b.v1, b.i1);
    top2_push(a, b.v2, b.i2);
}

__global__ void __launch_bounds__(256) greedy_select_kernel(const float* __restrict__ scores, int W, int lds, long long* __restrict__ output,
                                                            int* __restrict__ second, long long* __restrict__ captions,
                                                            long long* __restrict__ masks, int* __restrict__ done, float* __restrict__ margins,
                                                            int step, int Tmax, int V, int E, int has_facts, int end_tok) {
    ick_pdl_entry();
    __shared__ Top2 st[8];
    const int b = blockIdx.x;
    if (done[b]) return;
    const float* s = scores + (size_t)b * lds;
    Top2 t{-INFINITY, -INFINITY, 0x7fffffff, 0x7fffffff};
    for (int c0 = threadIdx.x; c0 < W; c0 += 256 * 8) {  // 8 loads in flight per thread (one per iteration is DRAM-latency-bound)
        float xs[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) xs[u] = c0 + 256 * u < W ? s[c0 + 256 * u] : -INFINITY;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (c0 + 256 * u < W) top2_push(t, xs[u], c0 + 256 * u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Top2 u;
        u.v1 = __shfl_xor_sync(0xffffffffu, t.v1, o);
        u.v2 = __shfl_xor_sync(0xffffffffu, t.v2, o);
        u.i1 = __shfl_xor_sync(0xffffffffu, t.i1, o);
        u.i2 = __shfl_xor_sync(0xffffffffu, t.i2, o);
        top2_merge(t, u);
    }
    if ((threadIdx.x & 31) == 0) st[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int w = 1; w < 8; ++w) top2_merge(t, st[w]);
    long long* out = output + (size_t)b * Tmax;
    int* sec = second + (size_t)b * Tmax;
    out[step] = t.i1;
    if (margins) margins[(size_t)b * Tmax + step] = t.v1 - t.v2;
    if (t.i1 == end_tok) {  // <end> is checked before the clean-up, G/models.py:414-416
        done[b] = 1;
        return;
    }
    sec[step] = t.i2;
    // repetition clean-up: repeat lengths 1,2,3 (dupl_idx 0,2,4), shortest first, first match wins (G/models.py:421-435)
    for (int dupl = 0; dupl <= 4; dupl += 2) {
        if (step > dupl) {
            const int n = (dupl + 2) / 2;
            bool same = true;
            for (int k = 0; k < n; ++k) same = same && (out[step - k] == out[step - n - k]);
            if (same) {
                const int nrw = dupl == 0 ? 1 : dupl;
                for (int k = 0; k < nrw; ++k) out[step - k] = sec[step - k];
                break;
            }
        }
    }
    if (step < Tmax - 1) {
        const long long o = out[step];
        captions[(size_t)b * Tmax + step + 1] = o;
        masks[(size_t)b * Tmax + step + 1] = (has_facts && o >= (long long)V + E) ? 2 : (o >= V ? 1 : 0);
    }
}

// ---- beam search step (extension - the reference decodes greedily) -------------------------------------------
// One CTA per image.  Rows [img*G, img*G + nrows) of `scores` are the live beams (nrows = 1 at step 0, else k = ksel[img]).
// candidate(j, c) = cum[j] + log_softmax(scores[j])[c]; the k best candidates over the live rows are taken in descending
// order (ties: lower row, then lower column).  A candidate whose token is <end> completes a caption - it replaces `result`
// if its score beats the best completed one - and lowers k; the others become the new beams 0..k'-1 in that order: token and
// mask history and the ancestor table of the parent are copied from the *_in to the *_out buffers and extended by the new token.
template <int G>
__device__ __forceinline__ void beam_insert(float (&v)[G], int (&ix)[G], float x, int i) {
    if (!(x > v[G - 1] || (x == v[G - 1] && i < ix[G - 1]))) return;
    v[G - 1] = x;
    ix[G - 1] = i;
#pragma unroll
    for (int k = G - 1; k > 0; --k) {
        const bool up = v[k] > v[k - 1] || (v[k] == v[k - 1] && ix[k] < ix[k - 1]);
        if (up) {
            const float tv = v[k]; v[k] = v[k - 1]; v[k - 1] = tv;
            const int ti = ix[k]; ix[k] = ix[k - 1]; ix[k - 1] = ti;
        }
    }
}

// stage 1, one CTA per (image, beam slot): log-sum-exp of the row and its G best columns, as candidates cum + log_softmax
template <int G>
__global__ void __launch_bounds__(256) beam_row_topk_kernel(const float* __restrict__ scores, int W, int lds, const float* __restrict__ cum,
                                                            const int* __restrict__ ksel, float* __restrict__ cand_v, int* __restrict__ cand_i,
                                                            int step) {
    ick_pdl_entry();
    __shared__ float red_m[8], red_l[8], red_v[8];
    __shared__ int red_i[8];
    __shared__ float s_logz;
    __shared__ int s_win;
    const int row = blockIdx.x, img = row / G, j = row % G, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k = ksel[img];
    if (k <= 0 || j >= (step == 0 ? 1 : k)) return;  // not a live beam
    const float* s = scores + (size_t)row * lds;
    float rm = -INFINITY, rl = 0.f, rv[G];
    int ri[G];
#pragma unroll
    for (int a = 0; a < G; ++a) { rv[a] = -INFINITY; ri[a] = 0x7fffffff; }
    // 8 independent loads per thread in flight: the scan itself is a dependent chain (running max, sorted insert), and with one
    // load per iteration the kernel is bound by DRAM latency (measured 0.9 TB/s), not bandwidth
    for (int c0 = tid; c0 < W; c0 += 256 * 8) {
        float xs[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) xs[u] = c0 + 256 * u < W ? s[c0 + 256 * u] : -INFINITY;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = c0 + 256 * u;
            if (c >= W) break;
            const float x = xs[u];
            if (x > rm) {
                rl = rl * expf(rm - x) + 1.f;
                rm = x;
            } else {
                rl += expf(x - rm);
            }
            beam_insert<G>(rv, ri, x, c);
        }
    }
    const float wm = warp_max(rm);
    const float wl = warp_sum(rm == -INFINITY ? 0.f : rl * expf(rm - wm));
    if (lane == 0) { red_m[wid] = wm; red_l[wid] = wl; }
    __syncthreads();
    if (tid == 0) {
        float M = red_m[0];
        for (int w = 1; w < 8; ++w) M = fmaxf(M, red_m[w]);
        float L = 0.f;
        for (int w = 0; w < 8; ++w) L += red_m[w] == -INFINITY ? 0.f : red_l[w] * expf(red_m[w] - M);
        s_logz = M + logf(L);
    }
    __syncthreads();
    const float logz = s_logz, cj = cum[row];
    // the G best of the row, one per round: every thread offers the head of its own sorted list
    for (int r = 0; r < G; ++r) {
        float v = rv[0];
        int i = ri[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
            if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; }
        }
        if (lane == 0) { red_v[wid] = v; red_i[wid] = i; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; ++w)
                if (red_v[w] > v || (red_v[w] == v && red_i[w] < i)) { v = red_v[w]; i = red_i[w]; }
            s_win = i;
            cand_v[(size_t)row * G + r] = i == 0x7fffffff ? -INFINITY : cj + (v - logz);
            cand_i[(size_t)row * G + r] = i == 0x7fffffff ? i : j * W + i;
        }
        __syncthreads();
        if (ri[0] == s_win && s_win != 0x7fffffff) {  // the winner pops its head
#pragma unroll
            for (int a = 0; a + 1 < G; ++a) { rv[a] = rv[a + 1]; ri[a] = ri[a + 1]; }
            rv[G - 1] = -INFINITY;
            ri[G - 1] = 0x7fffffff;
        }
    }
}

// stage 1, register-resident variant for rows of at most TK_NT * TK_NPT columns (every shape of the reference): the whole row sits in
// registers (all loads independent and in flight at once).  Every warp reduces its own columns: (max, sum of exponentials at that max)
// and its G best columns by G rounds of a branch-free thread-local argmax + a warp argmax (the winning lane blanks its element) - no
// block barrier inside the rounds; after ONE barrier warp 0 folds the per-warp pairs into log Z and takes the G best of the
// NW * G warp candidates.  (The first version ran the G rounds block-wide: 14 barriers with serial thread-0 folds per row, 100 us
// per launch at 3125 rows x 10352 columns = 1.3 TB/s.)  The scan kernel above keeps a sorted top-G list per thread, and with 40
// elements per thread some lane of a warp inserts at almost every element: the warp runs the divergent insertion path for the whole
// row (134 us per launch).  Same total order (value descending, column ascending), hence the same candidates.
constexpr int TK_NT = 512, TK_NPT = 24;
template <int G>
__global__ void __launch_bounds__(TK_NT) beam_row_topk_reg_kernel(const float* __restrict__ scores, int W, int lds, const float* __restrict__ cum,
                                                                  const int* __restrict__ ksel, float* __restrict__ cand_v,
                                                                  int* __restrict__ cand_i, int step) {
    ick_pdl_entry();
    constexpr int NW = TK_NT / 32;
    __shared__ float red_m[NW], red_l[NW];
    __shared__ float wv[NW * G];
    __shared__ int wi[NW * G];
    const int row = blockIdx.x, img = row / G, j = row % G, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int k = ksel[img];
    if (k <= 0 || j >= (step == 0 ? 1 : k)) return;  // not a live beam
    const float* s = scores + (size_t)row * lds;
    float xs[TK_NPT];
#pragma unroll
    for (int u = 0; u < TK_NPT; ++u) {
        const int c = tid + u * TK_NT;
        xs[u] = c < W ? s[c] : -INFINITY;
    }
    // log-sum-exp: per-warp (max, sum at that max), one barrier, then every warp folds the NW pairs itself (no serial thread-0 loop)
    float tm = xs[0];
#pragma unroll
    for (int u = 1; u < TK_NPT; ++u) tm = fmaxf(tm, xs[u]);
    tm = warp_max(tm);
    float ts = 0.f;
    if (tm != -INFINITY) {
#pragma unroll
        for (int u = 0; u < TK_NPT; ++u) ts += __expf(xs[u] - tm);  // ex2.approx (the kernel is issue-bound); exp(-inf) = 0 past the row
    }
    ts = warp_sum(ts);
    if (lane == 0) { red_m[wid] = tm; red_l[wid] = ts; }
    // the G best of this warp's columns, in order (value descending, column ascending): thread-local argmax + warp argmax per round,
    // the winning lane blanks its element.  No block barrier inside the rounds.
    for (int r = 0; r < G; ++r) {
        float bv = xs[0];
        int bu = 0;
#pragma unroll
        for (int u = 1; u < TK_NPT; ++u)
            if (xs[u] > bv) { bv = xs[u]; bu = u; }
        const int bi = bv == -INFINITY ? 0x7fffffff : tid + bu * TK_NT;
        float v = bv;
        int i = bi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
            if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; }
        }
        if (lane == 0) { wv[wid * G + r] = v; wi[wid * G + r] = i; }
        if (bi == i && bi != 0x7fffffff) {
#pragma unroll
            for (int u = 0; u < TK_NPT; ++u)
                if (u == bu) xs[u] = -INFINITY;
        }
    }
    __syncthreads();
    if (wid != 0) return;
    // warp 0: log Z from the NW pairs, then the G best of the NW * G warp candidates (every global winner is among them)
    float m = lane < NW ? red_m[lane] : -INFINITY;
    const float M = warp_max(m);
    float l = (lane < NW && m != -INFINITY) ? red_l[lane] * expf(m - M) : 0.f;
    l = warp_sum(l);
    const float logz = M + logf(l), cj = cum[row];
    constexpr int CPL = (NW * G + 31) / 32;  // candidates per lane
    float cv[CPL];
    int ci[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const int idx = lane + 32 * q;
        cv[q] = idx < NW * G ? wv[idx] : -INFINITY;
        ci[q] = idx < NW * G ? wi[idx] : 0x7fffffff;
    }
    for (int r = 0; r < G; ++r) {
        float v = cv[0];
        int i = ci[0];
#pragma unroll
        for (int q = 1; q < CPL; ++q)
            if (cv[q] > v || (cv[q] == v && ci[q] < i)) { v = cv[q]; i = ci[q]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
            if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; }
        }
        if (lane == 0) {
            cand_v[(size_t)row * G + r] = i == 0x7fffffff ? -INFINITY : cj + (v - logz);
            cand_i[(size_t)row * G + r] = i == 0x7fffffff ? i : j * W + i;
        }
#pragma unroll
        for (int q = 0; q < CPL; ++q)
            if (ci[q] == i && i != 0x7fffffff) { cv[q] = -INFINITY; ci[q] = 0x7fffffff; }
    }
}

// stage 2, one CTA per image: the k best of the live rows' candidate lists (each sorted), then the beam bookkeeping
template <int G>
__global__ void __launch_bounds__(64) beam_select_kernel(const float* __restrict__ cand_v, const int* __restrict__ cand_i, int W,
                                                         float* __restrict__ cum, int* __restrict__ ksel, const long long* __restrict__ tok_in,
                                                         const long long* __restrict__ mask_in, long long* __restrict__ tok_out,
                                                         long long* __restrict__ mask_out, const int* __restrict__ anc_in,
                                                         int* __restrict__ anc_out, float* __restrict__ best, long long* __restrict__ result,
                                                         int step, int Tmax, int V, int E, int has_facts, int end_tok, int pad_tok) {
    ick_pdl_entry();
    __shared__ float sel_v[G];
    __shared__ int sel_i[G], sel_slot[G];
    __shared__ int s_bestr, s_k;
    const int img = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) s_k = ksel[img];  // thread 0 rewrites ksel[img] below: everybody takes k from this one read
    __syncthreads();
    const int k = s_k;
    if (k <= 0) return;
    const int nrows = step == 0 ? 1 : k;
    if (tid == 0) {
        int head[G];
#pragma unroll
        for (int j = 0; j < G; ++j) head[j] = 0;
        for (int r = 0; r < k; ++r) {  // merge of nrows sorted lists; ties: lower (row, column)
            float bv = -INFINITY;
            int bi = 0x7fffffff, bj = -1;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                if (j >= nrows || head[j] >= G) continue;
                const float v = cand_v[(size_t)(img * G + j) * G + head[j]];
                const int i = cand_i[(size_t)(img * G + j) * G + head[j]];
                if (bj < 0 || v > bv || (v == bv && i < bi)) { bv = v; bi = i; bj = j; }
            }
            sel_v[r] = bv;
            sel_i[r] = bi;
#pragma unroll
            for (int j = 0; j < G; ++j)
                if (j == bj) ++head[j];
        }
        int nalive = 0, bestr = -1;
        float bv = best[img];
        for (int r = 0; r < k; ++r) {
            const int c = sel_i[r] % W;
            if (c == end_tok) {
                sel_slot[r] = -1;
                if (sel_v[r] > bv) { bv = sel_v[r]; bestr = r; }
            } else {
                sel_slot[r] = nalive++;
            }
        }
        if (step == Tmax - 1 && bv == -INFINITY && nalive > 0)  // nothing ever completed: the best live beam is the caption
            for (int r = 0; r < k; ++r)
                if (sel_slot[r] == 0) { bv = sel_v[r]; bestr = r; }
        best[img] = bv;
        s_bestr = bestr;
        ksel[img] = nalive;
    }
    __syncthreads();
    const int hist = step + 1;  // positions 0..step of the parent's history
    for (int e = tid; e < k * hist; e += 64) {
        const int r = e / hist, t = e % hist, slot = sel_slot[r];
        if (slot < 0) continue;
        const size_t src = (size_t)(img * G + sel_i[r] / W) * Tmax + t, dst = (size_t)(img * G + slot) * Tmax + t;
        tok_out[dst] = tok_in[src];
        mask_out[dst] = mask_in[src];
        anc_out[dst] = anc_in[src];
    }
    if (tid < k && sel_slot[tid] >= 0) {
        const int slot = sel_slot[tid];
        cum[img * G + slot] = sel_v[tid];
        if (step + 1 < Tmax) {
            const long long c = sel_i[tid] % W;
            const size_t dst = (size_t)(img * G + slot) * Tmax + step + 1;
            tok_out[dst] = c;
            mask_out[dst] = (has_facts && c >= (long long)V + E) ? 2 : (c >= V ? 1 : 0);
            anc_out[dst] = slot;
        }
    }
    const int bestr = s_bestr;
    if (bestr >= 0) {
        const size_t src = (size_t)(img * G + sel_i[bestr] / W) * Tmax;
        for (int t = tid; t < Tmax; t += 64)
            result[(size_t)img * Tmax + t] = t < step ? tok_in[src + t + 1] : (t == step ? (long long)(sel_i[bestr] % W) : (long long)pad_tok);
    }
}

inline int ew_grid(long long n) {
    long long g = (n + 255) / 256;
    return (int)(g < 148 * 16 ? (g < 1 ? 1 : g) : 148 * 16);
}

}  // namespace

extern "C" int ick_ce_fwd_bwd(const float* scores, const long long* captions_sorted, const int* decode_len, float* loss_acc,
                              void* dscores, int dt, int B, int T, int W, int lds, int ldd, int pad, cudaStream_t stream) {
    ICK_REQUIRE(B >= 0 && T > 0 && W > 0 && lds >= W, "ce: bad sizes");
    ICK_REQUIRE(dscores == nullptr || ldd >= W, "ce: ldd < W");
    if (B == 0) return ICK_OK;
    const bool reg_rows = W <= CE_NT * 2 * CE_NPP && ldd <= CE_NT * 2 * CE_NPP && ((W | lds | ldd) & 1) == 0 && (((uintptr_t)scores) & 7) == 0 &&
                          (dscores == nullptr || (((uintptr_t)dscores) & 7) == 0);
    if (reg_rows && (dscores == nullptr || dt == ICK_F32))
        ick_launch(ce_reg_kernel<float>, B * T, CE_NT, 0, stream)(scores, captions_sorted, decode_len, loss_acc, (float*)dscores, T, W, lds, ldd, pad);
    else if (reg_rows && dt == ICK_BF16)
        ick_launch(ce_reg_kernel<bf16>, B * T, CE_NT, 0, stream)(scores, captions_sorted, decode_len, loss_acc, (bf16*)dscores, T, W, lds, ldd, pad);
    else if (dscores == nullptr || dt == ICK_F32)
        ick_launch(ce_kernel<float>, B * T, 256, 0, stream)(scores, captions_sorted, decode_len, loss_acc, (float*)dscores, T, W, lds, ldd, pad);
    else if (dt == ICK_BF16)
        ick_launch(ce_kernel<bf16>, B * T, 256, 0, stream)(scores, captions_sorted, decode_len, loss_acc, (bf16*)dscores, T, W, lds, ldd, pad);
    else {
        ick_set_error("ce: bad dtype %d", dt);
        return ICK_ERR_UNSUPPORTED;
    }
    return ick_check_launch("ce_fwd_bwd");
}

extern "C" int ick_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                             float bias_corr1, float bias_corr2, float clip, const float* count, float grad_scale, const int* dstA,
                             const int* dstB, const int* dstC, void* packT, int dt, float* packF, int update, const int* step_dev,
                             const float* lr_dev, cudaStream_t stream) {
    ICK_REQUIRE(n >= 0, "adam: bad n");
    ICK_REQUIRE(!update || (g && m && v), "adam: update needs g, m, v");
    ICK_REQUIRE((!dstA && !dstB) || packT, "adam: packT missing");
    ICK_REQUIRE(!dstC || packF, "adam: packF missing");
    if (n == 0) return ICK_OK;
    const float bc2s = sqrtf(bias_corr2);
    const uintptr_t al = (uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)dstA | (uintptr_t)dstB | (uintptr_t)dstC | (uintptr_t)packT;
    if ((al & 15) == 0 && n >= 4 && (dt == ICK_F32 || dt == ICK_BF16)) {
        const long long n4 = n / 4, tail = n - 4 * n4;
        if (dt == ICK_F32)
            ick_launch(adam_kernel_v4<float>, ew_grid(n4), 256, 0, stream)(p, g, m, v, n4, lr, beta1, beta2, eps, bias_corr1, bc2s, clip, count,
                                                                        grad_scale, dstA, dstB, dstC, (float*)packT, packF, update, step_dev, lr_dev);
        else
            ick_launch(adam_kernel_v4<bf16>, ew_grid(n4), 256, 0, stream)(p, g, m, v, n4, lr, beta1, beta2, eps, bias_corr1, bc2s, clip, count,
                                                                       grad_scale, dstA, dstB, dstC, (bf16*)packT, packF, update, step_dev, lr_dev);
        if (tail > 0) {  // the last 1-3 parameters through the scalar kernel
            const long long o = 4 * n4;
            const int* a = dstA ? dstA + o : nullptr;
            const int* b = dstB ? dstB + o : nullptr;
            const int* c = dstC ? dstC + o : nullptr;
            if (dt == ICK_F32)
                ick_launch(adam_kernel<float>, 1, 32, 0, stream)(p + o, g ? g + o : g, m ? m + o : m, v ? v + o : v, tail, lr, beta1, beta2, eps,
                                                                 bias_corr1, bc2s, clip, count, grad_scale, a, b, c, (float*)packT, packF, update,
                                                                 step_dev, lr_dev);
            else
                ick_launch(adam_kernel<bf16>, 1, 32, 0, stream)(p + o, g ? g + o : g, m ? m + o : m, v ? v + o : v, tail, lr, beta1, beta2, eps,
                                                                bias_corr1, bc2s, clip, count, grad_scale, a, b, c, (bf16*)packT, packF, update,
                                                                step_dev, lr_dev);
        }
        return ick_check_launch("adam_step");
    }
    if (dt == ICK_F32)
        ick_launch(adam_kernel<float>, ew_grid(n), 256, 0, stream)(p, g, m, v, n, lr, beta1, beta2, eps, bias_corr1, bc2s, clip, count, grad_scale,
                                                           dstA, dstB, dstC, (float*)packT, packF, update, step_dev, lr_dev);
    else if (dt == ICK_BF16)
        ick_launch(adam_kernel<bf16>, ew_grid(n), 256, 0, stream)(p, g, m, v, n, lr, beta1, beta2, eps, bias_corr1, bc2s, clip, count, grad_scale,
                                                          dstA, dstB, dstC, (bf16*)packT, packF, update, step_dev, lr_dev);
    else {
        ick_set_error("adam: bad dtype %d", dt);
        return ICK_ERR_UNSUPPORTED;
    }
    return ick_check_launch("adam_step");
}

extern "C" int ick_cast2d(const void* src, int src_dt, void* dst, int dst_dt, long long rows, int cols, int lds, int ldd,
                          cudaStream_t stream) {
    ICK_REQUIRE(rows >= 0 && cols >= 0 && lds >= cols, "cast2d: bad sizes");
    if (rows * ldd == 0) return ICK_OK;
    const int grid = ew_grid(rows * ldd);
    if (src_dt == ICK_F32 && dst_dt == ICK_BF16)
        ick_launch(cast2d_kernel<float, bf16>, grid, 256, 0, stream)((const float*)src, (bf16*)dst, rows, cols, lds, ldd);
    else if (src_dt == ICK_F32 && dst_dt == ICK_F32)
        ick_launch(cast2d_kernel<float, float>, grid, 256, 0, stream)((const float*)src, (float*)dst, rows, cols, lds, ldd);
    else if (src_dt == ICK_BF16 && dst_dt == ICK_F32)
        ick_launch(cast2d_kernel<bf16, float>, grid, 256, 0, stream)((const bf16*)src, (float*)dst, rows, cols, lds, ldd);
    else if (src_dt == ICK_BF16 && dst_dt == ICK_BF16)
        ick_launch(cast2d_kernel<bf16, bf16>, grid, 256, 0, stream)((const bf16*)src, (bf16*)dst, rows, cols, lds, ldd);
    else {
        ick_set_error("cast2d: bad dtypes %d -> %d", src_dt, dst_dt);
        return ICK_ERR_UNSUPPORTED;
    }
    return ick_check_launch("cast2d");
}

extern "C" int ick_accum_f32(const void* src, int dt, float* dst, long long n, cudaStream_t stream) {
    if (n == 0) return ICK_OK;
    if (n % 8 == 0 && ((((uintptr_t)src) | ((uintptr_t)dst)) & 31) == 0 && (dt == ICK_F32 || dt == ICK_BF16)) {
        if (dt == ICK_F32) ick_launch(accum_f32_v8_kernel<float>, ew_grid(n / 8), 256, 0, stream)((const float*)src, dst, n / 8);
        else ick_launch(accum_f32_v8_kernel<bf16>, ew_grid(n / 8), 256, 0, stream)((const bf16*)src, dst, n / 8);
        return ick_check_launch("accum_f32");
    }
    if (dt == ICK_F32) ick_launch(accum_f32_kernel<float>, ew_grid(n), 256, 0, stream)((const float*)src, dst, n);
    else if (dt == ICK_BF16) ick_launch(accum_f32_kernel<bf16>, ew_grid(n), 256, 0, stream)((const bf16*)src, dst, n);
    else {
        ick_set_error("accum_f32: bad dtype %d", dt);
        return ICK_ERR_UNSUPPORTED;
    }
    return ick_check_launch("accum_f32");
}

extern "C" int ick_colsum(const void* x, int dt, float* out, long long rows, int cols, int ld, cudaStream_t stream) {
    if (rows == 0 || cols == 0) return ICK_OK;
    const int rpb = 32;  // 13 056 rows -> 408 blocks
    const int grid = (int)((rows + rpb - 1) / rpb);
    if (dt == ICK_F32) ick_launch(colsum_kernel<float>, grid, 256, 0, stream)((const float*)x, out, rows, cols, ld, rpb);
    else if (dt == ICK_BF16) ick_launch(colsum_kernel<bf16>, grid, 256, 0, stream)((const bf16*)x, out, rows, cols, ld, rpb);
    else {
        ick_set_error("colsum: bad dtype %d", dt);
        return ICK_ERR_UNSUPPORTED;
    }
    return ick_check_launch("colsum");
}

extern "C" int ick_greedy_select(const float* scores, int W, int lds, long long* output, int* second, long long* captions,
                                 long long* masks, int* done, float* margins, int B, int step, int Tmax, int V, int E, int has_facts,
                                 int end_tok, cudaStream_t stream) {
    ICK_REQUIRE(B >= 0 && step >= 0 && step < Tmax && W >= 2, "greedy_select: bad sizes");
    if (B == 0) return ICK_OK;
    ick_launch(greedy_select_kernel, B, 256, 0, stream)(scores, W, lds, output, second, captions, masks, done, margins, step, Tmax, V, E, has_facts,
                                                end_tok);
    return ick_check_launch("greedy_select");
}

extern "C" int ick_beam_select(const float* scores, int W, int lds, float* cum, int* ksel, const long long* tok_in,
                               const long long* mask_in, long long* tok_out, long long* mask_out, const int* anc_in, int* anc_out,
                               float* best, long long* result, int images, int group, int step, int Tmax, int V, int E, int has_facts,
                               int end_tok, int pad_tok, void* workspace, long long workspace_bytes, cudaStream_t stream) {
    ICK_REQUIRE(images >= 0 && step >= 0 && step < Tmax && W >= 2, "beam_select: bad sizes");
    ICK_REQUIRE(group >= 1 && group <= 8 && (long long)group * W < 0x7fffffffLL, "beam_select: group=%d out of range", group);
    ICK_REQUIRE(tok_in != tok_out && mask_in != mask_out && anc_in != anc_out, "beam_select: histories must be double-buffered");
    const long long ncand = (long long)images * group * group;
    ICK_REQUIRE(workspace != nullptr && workspace_bytes >= ncand * 8, "beam_select: workspace of %lld bytes needed", ncand * 8);
    if (images == 0) return ICK_OK;
    float* cand_v = (float*)workspace;
    int* cand_i = (int*)(cand_v + ncand);
    static int scan_only = -1;  // ICK_BEAM_TOPK=scan: always the per-thread sorted-list kernel (A/B aid)
    if (scan_only < 0) {
        const char* e = getenv("ICK_BEAM_TOPK");
        scan_only = (e && e[0] == 's') ? 1 : 0;
    }
    const bool reg_rows = !scan_only && W <= TK_NT * TK_NPT;
#define ICK_BEAM_SEL(GG)                                                                                                               \
    case GG:                                                                                                                           \
        if (reg_rows)                                                                                                                  \
            ick_launch(beam_row_topk_reg_kernel<GG>, images * GG, TK_NT, 0, stream)(scores, W, lds, cum, ksel, cand_v, cand_i, step);        \
        else                                                                                                                           \
            ick_launch(beam_row_topk_kernel<GG>, images * GG, 256, 0, stream)(scores, W, lds, cum, ksel, cand_v, cand_i, step);          \
        ick_launch(beam_select_kernel<GG>, images, 64, 0, stream)(cand_v, cand_i, W, cum, ksel, tok_in, mask_in, tok_out, mask_out, anc_in, \
                                                                   anc_out, best, result, step, Tmax, V, E, has_facts, end_tok, pad_tok); \
        break;
    switch (group) {
        ICK_BEAM_SEL(1)
        ICK_BEAM_SEL(2)
        ICK_BEAM_SEL(3)
        ICK_BEAM_SEL(4)
        ICK_BEAM_SEL(5)
        ICK_BEAM_SEL(6)
        ICK_BEAM_SEL(7)
        ICK_BEAM_SEL(8)
    }
#undef ICK_BEAM_SEL
    return ick_check_launch("beam_select");
}

// ---- error plumbing ---------------------------------------------------------------------------------------------------
#include <cstdarg>
#include <cstdio>
static thread_local char g_err[512] = "";
void ick_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int ick_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        ick_set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return ICK_ERR_CUDA;
    }
    return ICK_OK;
}
extern "C" const char* ick_last_error(void) { return g_err; }
static const uint32_t* g_seed_source = nullptr;
const uint32_t* ick_seed_source() { return g_seed_source; }

bool ick_pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ICK_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}
extern "C" int ick_set_seed_source(const unsigned* seed_dev) {
    g_seed_source = seed_dev;
    return ICK_OK;
}
extern "C" int ick_abi_version(void) { return ICK_ABI_VERSION; }

long long ick_launch_counter = 0;
extern "C" int ick_launch_count(long long* out) {
    ICK_REQUIRE(out != nullptr, "launch_count: null output");
    *out = ick_launch_counter;
    return ICK_OK;
}
