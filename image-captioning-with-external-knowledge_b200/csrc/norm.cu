// Fused residual-add + dropout + LayerNorm (post-LN Transformer sublayer tail), forward and backward.
//   s = x + dropout(sub);  y = (s - mean) * rstd * gamma + beta          (torch LayerNorm: biased variance, eps 1e-5)
// `sub` is overwritten with s (needed by the backward pass, saves one tensor); y can be written with a per-batch row
// remap so that the last encoder layer writes straight into the decoder's memory buffer [pixels; entities; facts]
// (the reference concatenates, G/models.py:349 / K/models.py:497-499).  One warp per row, HBM-bound.
#include "common.cuh"
#include "ickb200.h"

namespace {

constexpr int MAXP = 8;  // pairs per lane: supports d <= 512

struct RowMap {
    int s_in, s_out, off;  // out_row = (r / s_in) * s_out + off + r % s_in ; s_in == 0 -> identity
    __device__ __forceinline__ size_t map(int r) const {
        return s_in == 0 ? (size_t)r : (size_t)(r / s_in) * s_out + off + (r % s_in);
    }
};


// ---- specialised kernels for the model width (compile-time D / LD: no per-pair predicates, address or loop arithmetic) ----
// Same arithmetic contract as the generic kernels below; the profile of the generic backward kernel showed it issue-bound
// (IMAD / ISETP / BRA were a third of all instructions), so everything that depends on d is folded at compile time, the
// normalisation is written as FMAs and the per-row scalars are hoisted.
template <int NP, int I>
__device__ __forceinline__ bool pair_valid(int lane) {
    if constexpr (32 * I + 31 < NP) return true;
    else if constexpr (32 * I >= NP) return false;
    else return lane + 32 * I < NP;
}

// Up to two row groups with their own LayerNorm parameters, dropout site and output row map in one launch (the entity and the
// fact encoder stacks run in lockstep on one row-concatenated buffer): blocks [0, nb0) take group 0, the rest group 1, so a
// block (and its dgamma/dbeta partials in the backward) belongs to exactly one parameter set.  Rows and dropout row indices
// are group-local; tensors are indexed by row_start + r.
struct LnGroup {
    const float* gamma;
    const float* beta;   // forward only
    float* dgamma;       // backward only
    float* dbeta;
    RowMap map;          // forward: output row map; backward: dy row map
    int rows, row_start;
    uint32_t site;
};
struct LnGroups {
    LnGroup g[2];
    int nb0;
};

template <typename T, int D, int LD, int MINB>
__global__ void __launch_bounds__(256, MINB) add_ln_fwd_fast_kernel(const T* __restrict__ X, T* __restrict__ SUB, T* __restrict__ Y,
                                                                    float* __restrict__ MEAN, float* __restrict__ RSTD, float eps,
                                                                    LnGroups gs, DropCfg drop) {
    ick_pdl_entry();
    ick_resolve_seed(drop);
    constexpr int NP = D / 2, NPL = LD / 2, NI = (NPL + 31) / 32;
    const bool second = (int)blockIdx.x >= gs.nb0;
    const LnGroup& G = gs.g[second ? 1 : 0];
    const int warp = (((int)blockIdx.x - (second ? gs.nb0 : 0)) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int nwarps = ((second ? (int)gridDim.x - gs.nb0 : gs.nb0) * blockDim.x) >> 5;
    const int rows = G.rows;
    const bool has_x = X != nullptr;
    const bool dropping = drop.thr != 0u;
    const float* gamma = G.gamma;
    const float* beta = G.beta;
    const RowMap ymap = G.map;
    float2 gm[NI], bt[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int p = lane + 32 * i;
        const bool ok = p < NP;
        gm[i] = ok ? *reinterpret_cast<const float2*>(gamma + 2 * p) : make_float2(0.f, 0.f);
        bt[i] = ok ? *reinterpret_cast<const float2*>(beta + 2 * p) : make_float2(0.f, 0.f);
    }
    for (int r = warp; r < rows; r += nwarps) {
        const size_t gr = (size_t)(G.row_start + r);
        const T* x = X + gr * LD;
        T* s = SUB + gr * LD;
        float2 v[NI], xr[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) {  // all loads of the row first (memory-level parallelism)
            const int p = lane + 32 * i;
            v[i] = make_float2(0.f, 0.f);
            xr[i] = make_float2(0.f, 0.f);
            if (p < NP) {
                v[i] = ld2(s + 2 * p);
                if (has_x) xr[i] = ld2(x + 2 * p);
            }
        }
        float sum = 0.f;
        if (dropping) {
            const uint32_t rmix = ick_rowmix(drop.seed, G.site, (uint64_t)r);
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const int p = lane + 32 * i;
                const uint32_t hsh = ick_pairhash_idx(rmix, (uint32_t)p);
                v[i].x *= ick_keep_lo(hsh, drop.thr) ? drop.inv_keep : 0.f;
                v[i].y *= ick_keep_hi(hsh, drop.thr) ? drop.inv_keep : 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int p = lane + 32 * i;
            v[i].x += xr[i].x;
            v[i].y += xr[i].y;
            if (p < NP) st2(s + 2 * p, v[i].x, v[i].y);
            sum += v[i].x + v[i].y;
        }
        const float mean = warp_sum(sum) * (1.0f / (float)D);
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int p = lane + 32 * i;
            const float a = v[i].x - mean, b = v[i].y - mean;
            if (p < NP) var += a * a + b * b;
        }
        const float rstd = rsqrtf(warp_sum(var) * (1.0f / (float)D) + eps);
        const float nmr = -mean * rstd;
        // a row map addresses the output buffer from its base; without one the output shares the input's row numbering
        T* y = Y + (ymap.s_in == 0 ? gr : ymap.map(r)) * LD;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int p = lane + 32 * i;
            if (p < NP) {
                st2(y + 2 * p, fmaf(fmaf(v[i].x, rstd, nmr), gm[i].x, bt[i].x), fmaf(fmaf(v[i].y, rstd, nmr), gm[i].y, bt[i].y));
            } else if (p < NPL) {
                st2(y + 2 * p, 0.f, 0.f);  // zero the pad columns [D, LD)
            }
        }
        if (lane == 0) {
            MEAN[gr] = mean;
            RSTD[gr] = rstd;
        }
    }
}

template <typename T, int D, int LD, int MINB>
__global__ void __launch_bounds__(256, MINB) add_ln_bwd_fast_kernel(const T* __restrict__ DY, const T* __restrict__ S,
                                                                    const float* __restrict__ MEAN, const float* __restrict__ RSTD, T* DRES,
                                                                    T* __restrict__ DSUB, int acc_res, LnGroups gs, DropCfg drop) {
    ick_pdl_entry();
    constexpr int NP = D / 2, NPL = LD / 2, NI = (NPL + 31) / 32;
    __shared__ float sg[2 * NI * 32], sb[2 * NI * 32];
    ick_resolve_seed(drop);
    const bool second = (int)blockIdx.x >= gs.nb0;
    const LnGroup& G = gs.g[second ? 1 : 0];
    const int warp = (((int)blockIdx.x - (second ? gs.nb0 : 0)) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int nwarps = ((second ? (int)gridDim.x - gs.nb0 : gs.nb0) * blockDim.x) >> 5;
    const int rows = G.rows;
    const float* gamma = G.gamma;
    float* dgamma = G.dgamma;
    float* dbeta = G.dbeta;
    const RowMap dymap = G.map;
    const bool dropping = drop.thr != 0u;
    const bool has_res = DRES != nullptr, has_sub = DSUB != nullptr;
    for (int i = threadIdx.x; i < 2 * NI * 32; i += blockDim.x) { sg[i] = 0.f; sb[i] = 0.f; }
    __syncthreads();
    float2 gm[NI], ag[NI], ab[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int p = lane + 32 * i;
        gm[i] = p < NP ? *reinterpret_cast<const float2*>(gamma + 2 * p) : make_float2(0.f, 0.f);
        ag[i] = make_float2(0.f, 0.f);
        ab[i] = make_float2(0.f, 0.f);
    }
    constexpr float inv_d = 1.0f / (float)D;
    for (int r = warp; r < rows; r += nwarps) {
        const size_t gr = (size_t)(G.row_start + r);
        const T* dy = DY + (dymap.s_in == 0 ? gr : dymap.map(r)) * LD;
        const T* s = S + gr * LD;
        const float mean = MEAN[gr], rstd = RSTD[gr];
        const float nmr = -mean * rstd;
        float2 g[NI], xh[NI];
        float sum_g = 0.f, sum_gx = 0.f;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int p = lane + 32 * i;
            g[i] = make_float2(0.f, 0.f);
            xh[i] = make_float2(0.f, 0.f);
            if (p < NP) {
                g[i] = ld2(dy + 2 * p);
                xh[i] = ld2(s + 2 * p);
            }
        }
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int p = lane + 32 * i;
            const bool ok = p < NP;
            xh[i].x = ok ? fmaf(xh[i].x, rstd, nmr) : 0.f;
            xh[i].y = ok ? fmaf(xh[i].y, rstd, nmr) : 0.f;
            ag[i].x = fmaf(g[i].x, xh[i].x, ag[i].x);
            ag[i].y = fmaf(g[i].y, xh[i].y, ag[i].y);
            ab[i].x += g[i].x;
            ab[i].y += g[i].y;
            g[i].x *= gm[i].x;
            g[i].y *= gm[i].y;
            sum_g += g[i].x + g[i].y;
            sum_gx = fmaf(g[i].x, xh[i].x, fmaf(g[i].y, xh[i].y, sum_gx));
        }
        const float c1 = warp_sum(sum_g) * inv_d * rstd;
        const float c2 = warp_sum(sum_gx) * inv_d * rstd;
        T* dres = DRES + gr * LD;
        T* dsub = DSUB + gr * LD;
        const uint32_t rmix = dropping ? ick_rowmix(drop.seed, G.site, (uint64_t)r) : 0u;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int p = lane + 32 * i;
            if (p < NP) {
                float a = fmaf(-xh[i].x, c2, fmaf(g[i].x, rstd, -c1));
                float b = fmaf(-xh[i].y, c2, fmaf(g[i].y, rstd, -c1));
                if (has_sub) {
                    float k0 = 1.f, k1 = 1.f;
                    if (dropping) {
                        const uint32_t hsh = ick_pairhash_idx(rmix, (uint32_t)p);
                        k0 = ick_keep_lo(hsh, drop.thr) ? drop.inv_keep : 0.f;
                        k1 = ick_keep_hi(hsh, drop.thr) ? drop.inv_keep : 0.f;
                    }
                    st2(dsub + 2 * p, a * k0, b * k1);
                }
                if (has_res) {
                    if (acc_res) {
                        const float2 o = ld2(dres + 2 * p);
                        a += o.x;
                        b += o.y;
                    }
                    st2(dres + 2 * p, a, b);
                }
            } else if (p < NPL) {
                if (has_sub) st2(dsub + 2 * p, 0.f, 0.f);
                if (has_res && !acc_res) st2(dres + 2 * p, 0.f, 0.f);
            }
        }
    }
    // block reduction of the per-lane dgamma/dbeta partials, then one atomic per column per CTA
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int p = lane + 32 * i;
        if (p < NP) {
            atomicAdd(&sg[2 * p], ag[i].x);
            atomicAdd(&sg[2 * p + 1], ag[i].y);
            atomicAdd(&sb[2 * p], ab[i].x);
            atomicAdd(&sb[2 * p + 1], ab[i].y);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        if (dgamma) atomicAdd(dgamma + c, sg[c]);
        if (dbeta) atomicAdd(dbeta + c, sb[c]);
    }
}

constexpr int FAST_D = 300, FAST_LD = 320;  // the model width of all three reference variants (emb_dim 300, G/train.py:28)

template <typename T>
__global__ void __launch_bounds__(256) add_ln_fwd_kernel(const T* __restrict__ X, T* __restrict__ SUB,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         T* __restrict__ Y, float* __restrict__ MEAN, float* __restrict__ RSTD,
                                                         int rows, int d, int ldx, int lds, int ldy, float eps, RowMap ymap,
                                                         DropCfg drop) {
    ick_pdl_entry();
    ick_resolve_seed(drop);
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int npairs = d >> 1;
    for (int r = warp; r < rows; r += nwarps) {
        const T* x = X ? X + (size_t)r * ldx : nullptr;
        T* s = SUB + (size_t)r * lds;
        float v[2 * MAXP];
        float sum = 0.f;
        const uint32_t rmix = ick_rowmix(drop.seed, drop.site, (uint64_t)r);
#pragma unroll
        for (int i = 0; i < MAXP; ++i) {
            const int p = lane + 32 * i;
            float a = 0.f, b = 0.f;
            if (p < npairs) {
                const float2 u = ld2(s + 2 * p);
                a = u.x;
                b = u.y;
                if (drop.thr != 0u) {
                    const uint32_t hsh = ick_pairhash(rmix, (uint32_t)(2 * p));
                    a *= ick_keep_lo(hsh, drop.thr) ? drop.inv_keep : 0.f;
                    b *= ick_keep_hi(hsh, drop.thr) ? drop.inv_keep : 0.f;
                }
                if (x) {
                    const float2 xr = ld2(x + 2 * p);
                    a += xr.x;
                    b += xr.y;
                }
                st2(s + 2 * p, a, b);
            }
            v[2 * i] = a;
            v[2 * i + 1] = b;
            sum += a + b;
        }
        const float mean = warp_sum(sum) / (float)d;
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < MAXP; ++i) {
            const int p = lane + 32 * i;
            if (p < npairs) {
                const float a = v[2 * i] - mean, b = v[2 * i + 1] - mean;
                var += a * a + b * b;
            }
        }
        const float rstd = rsqrtf(warp_sum(var) / (float)d + eps);
        T* y = Y + ymap.map(r) * ldy;
#pragma unroll
        for (int i = 0; i < MAXP; ++i) {
            const int p = lane + 32 * i;
            if (p < npairs) {
                const float2 g = *reinterpret_cast<const float2*>(gamma + 2 * p);
                const float2 bb = *reinterpret_cast<const float2*>(beta + 2 * p);
                st2(y + 2 * p, (v[2 * i] - mean) * rstd * g.x + bb.x, (v[2 * i + 1] - mean) * rstd * g.y + bb.y);
            } else if (2 * p < ldy) {
                st2(y + 2 * p, 0.f, 0.f);  // zero the pad columns [d, ld)
            }
        }
        if (lane == 0) {
            MEAN[r] = mean;
            RSTD[r] = rstd;
        }
    }
}

// dy rows may be remapped (dymap) when the gradient arrives through the memory buffer.
// Outputs: DRES (= ds, or += ds when acc_res) and DSUB (= ds * dropmask); dgamma/dbeta accumulate atomically.
template <typename T>
__global__ void __launch_bounds__(256) add_ln_bwd_kernel(const T* __restrict__ DY, const T* __restrict__ S,
                                                         const float* __restrict__ MEAN, const float* __restrict__ RSTD,
                                                         const float* __restrict__ gamma, T* DRES, T* __restrict__ DSUB,
                                                         float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int d,
                                                         int lddy, int lds, int ldres, int ldsub, RowMap dymap, int acc_res,
                                                         DropCfg drop) {
    ick_pdl_entry();
    __shared__ float sg[2 * MAXP * 32], sb[2 * MAXP * 32];
    ick_resolve_seed(drop);
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int npairs = d >> 1;
    for (int i = threadIdx.x; i < 2 * MAXP * 32; i += blockDim.x) { sg[i] = 0.f; sb[i] = 0.f; }
    __syncthreads();

    float gam[2 * MAXP], ag[2 * MAXP], ab[2 * MAXP];
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
        const int p = lane + 32 * i;
        gam[2 * i] = p < npairs ? gamma[2 * p] : 0.f;
        gam[2 * i + 1] = p < npairs ? gamma[2 * p + 1] : 0.f;
        ag[2 * i] = ag[2 * i + 1] = ab[2 * i] = ab[2 * i + 1] = 0.f;
    }
    for (int r = warp; r < rows; r += nwarps) {
        const T* dy = DY + dymap.map(r) * lddy;
        const T* s = S + (size_t)r * lds;
        const float mean = MEAN[r], rstd = RSTD[r];
        const uint32_t rmix = ick_rowmix(drop.seed, drop.site, (uint64_t)r);
        float g[2 * MAXP], xh[2 * MAXP];
        float sum_g = 0.f, sum_gx = 0.f;
#pragma unroll
        for (int i = 0; i < MAXP; ++i) {
            const int p = lane + 32 * i;
            float2 dyv = make_float2(0.f, 0.f), sv = make_float2(mean, mean);
            if (p < npairs) {
                dyv = ld2(dy + 2 * p);
                sv = ld2(s + 2 * p);
            }
            xh[2 * i] = (sv.x - mean) * rstd;
            xh[2 * i + 1] = (sv.y - mean) * rstd;
            g[2 * i] = dyv.x * gam[2 * i];
            g[2 * i + 1] = dyv.y * gam[2 * i + 1];
            ag[2 * i] += dyv.x * xh[2 * i];
            ag[2 * i + 1] += dyv.y * xh[2 * i + 1];
            ab[2 * i] += dyv.x;
            ab[2 * i + 1] += dyv.y;
            sum_g += g[2 * i] + g[2 * i + 1];
            sum_gx += g[2 * i] * xh[2 * i] + g[2 * i + 1] * xh[2 * i + 1];
        }
        const float mg = warp_sum(sum_g) / (float)d;
        const float mgx = warp_sum(sum_gx) / (float)d;
        T* dres = DRES ? DRES + (size_t)r * ldres : nullptr;
        T* dsub = DSUB ? DSUB + (size_t)r * ldsub : nullptr;
#pragma unroll
        for (int i = 0; i < MAXP; ++i) {
            const int p = lane + 32 * i;
            if (p < npairs) {
                float a = rstd * (g[2 * i] - mg - xh[2 * i] * mgx);
                float b = rstd * (g[2 * i + 1] - mg - xh[2 * i + 1] * mgx);
                if (dsub) {
                    float k0 = 1.f, k1 = 1.f;
                    if (drop.thr != 0u) {
                        const uint32_t hsh = ick_pairhash(rmix, (uint32_t)(2 * p));
                        k0 = ick_keep_lo(hsh, drop.thr) ? drop.inv_keep : 0.f;
                        k1 = ick_keep_hi(hsh, drop.thr) ? drop.inv_keep : 0.f;
                    }
                    st2(dsub + 2 * p, a * k0, b * k1);
                }
                if (dres) {
                    if (acc_res) {
                        const float2 o = ld2(dres + 2 * p);
                        a += o.x;
                        b += o.y;
                    }
                    st2(dres + 2 * p, a, b);
                }
            } else {
                if (dsub && 2 * p < ldsub) st2(dsub + 2 * p, 0.f, 0.f);
                if (dres && !acc_res && 2 * p < ldres) st2(dres + 2 * p, 0.f, 0.f);
            }
        }
    }
    // block reduction of the per-lane dgamma/dbeta partials, then one atomic per column per CTA
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
        const int p = lane + 32 * i;
        if (p < npairs) {
            atomicAdd(&sg[2 * p], ag[2 * i]);
            atomicAdd(&sg[2 * p + 1], ag[2 * i + 1]);
            atomicAdd(&sb[2 * p], ab[2 * i]);
            atomicAdd(&sb[2 * p + 1], ab[2 * i + 1]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        if (dgamma) atomicAdd(dgamma + c, sg[c]);
        if (dbeta) atomicAdd(dbeta + c, sb[c]);
    }
}

int launch_ln_fwd_fast(const void* x, void* sub, void* y, float* mean, float* rstd, int dt, float eps, const LnGroups& gs, int blocks,
                       const DropCfg& dc, cudaStream_t stream) {
    if (dt == ICK_F32)
        ick_launch(add_ln_fwd_fast_kernel<float, FAST_D, FAST_LD, 4>, blocks, 256, 0, stream)((const float*)x, (float*)sub, (float*)y, mean, rstd, eps,
                                                                                           gs, dc);
    else
        ick_launch(add_ln_fwd_fast_kernel<bf16, FAST_D, FAST_LD, 4>, blocks, 256, 0, stream)((const bf16*)x, (bf16*)sub, (bf16*)y, mean, rstd, eps, gs,
                                                                                          dc);
    return ick_check_launch("add_ln_fwd");
}
int launch_ln_bwd_fast(const void* dy, const void* s, const float* mean, const float* rstd, void* dres, void* dsub, int dt, int acc_res,
                       const LnGroups& gs, int blocks, const DropCfg& dc, cudaStream_t stream) {
    if (dt == ICK_F32)
        ick_launch(add_ln_bwd_fast_kernel<float, FAST_D, FAST_LD, 3>, blocks, 256, 0, stream)((const float*)dy, (const float*)s, mean, rstd,
                                                                                           (float*)dres, (float*)dsub, acc_res, gs, dc);
    else
        ick_launch(add_ln_bwd_fast_kernel<bf16, FAST_D, FAST_LD, 3>, blocks, 256, 0, stream)((const bf16*)dy, (const bf16*)s, mean, rstd, (bf16*)dres,
                                                                                          (bf16*)dsub, acc_res, gs, dc);
    return ick_check_launch("add_ln_bwd");
}
// blocks of a two-group launch: proportional to the rows, at least one block per non-empty group
void split_blocks(int rows0, int rows1, int max_blocks, int* nb0, int* nb) {
    int total = min((rows0 + rows1 + 7) / 8, max_blocks);
    if (total < 2) total = 2;
    int b0 = (int)((long long)total * rows0 / (rows0 + rows1));
    if (b0 < 1) b0 = 1;
    if (b0 > total - 1) b0 = total - 1;
    *nb0 = b0;
    *nb = total;
}

}  // namespace

extern "C" int ick_add_ln_fwd_dual(const void* x, void* sub, void* y, float* mean, float* rstd, int dt, int d, int ld, float eps, int rows0,
                                   int rows1, int row1_start, const float* gamma0, const float* beta0, const float* gamma1, const float* beta1,
                                   int map0_s_in, int map0_s_out, int map0_off, int map1_s_in, int map1_s_out, int map1_off, float drop_p,
                                   unsigned seed, unsigned site0, unsigned site1, cudaStream_t stream) {
    ICK_REQUIRE(d == FAST_D && ld == FAST_LD && (dt == ICK_F32 || dt == ICK_BF16), "add_ln_fwd_dual: only d=%d, ld=%d rows are supported", FAST_D,
                FAST_LD);
    ICK_REQUIRE(rows0 > 0 && rows1 > 0 && row1_start >= rows0, "add_ln_fwd_dual: bad row groups");
    DropCfg dc = make_drop(drop_p, seed, site0);
    LnGroups gs = {};
    gs.g[0].gamma = gamma0; gs.g[0].beta = beta0; gs.g[0].map = RowMap{map0_s_in, map0_s_out, map0_off}; gs.g[0].rows = rows0; gs.g[0].row_start = 0;
    gs.g[0].site = site0;
    gs.g[1].gamma = gamma1; gs.g[1].beta = beta1; gs.g[1].map = RowMap{map1_s_in, map1_s_out, map1_off}; gs.g[1].rows = rows1;
    gs.g[1].row_start = row1_start; gs.g[1].site = site1;
    int nb;
    split_blocks(rows0, rows1, 148 * 4, &gs.nb0, &nb);
    return launch_ln_fwd_fast(x, sub, y, mean, rstd, dt, eps, gs, nb, dc, stream);
}

extern "C" int ick_add_ln_bwd_dual(const void* dy, const void* s, const float* mean, const float* rstd, void* dres, void* dsub, int dt, int d,
                                   int ld, int rows0, int rows1, int row1_start, const float* gamma0, const float* gamma1, float* dgamma0,
                                   float* dbeta0, float* dgamma1, float* dbeta1, int map0_s_in, int map0_s_out, int map0_off, int map1_s_in,
                                   int map1_s_out, int map1_off, int acc_res, float drop_p, unsigned seed, unsigned site0, unsigned site1,
                                   cudaStream_t stream) {
    ICK_REQUIRE(d == FAST_D && ld == FAST_LD && (dt == ICK_F32 || dt == ICK_BF16), "add_ln_bwd_dual: only d=%d, ld=%d rows are supported", FAST_D,
                FAST_LD);
    ICK_REQUIRE(rows0 > 0 && rows1 > 0 && row1_start >= rows0, "add_ln_bwd_dual: bad row groups");
    DropCfg dc = make_drop(drop_p, seed, site0);
    LnGroups gs = {};
    gs.g[0].gamma = gamma0; gs.g[0].dgamma = dgamma0; gs.g[0].dbeta = dbeta0; gs.g[0].map = RowMap{map0_s_in, map0_s_out, map0_off};
    gs.g[0].rows = rows0; gs.g[0].row_start = 0; gs.g[0].site = site0;
    gs.g[1].gamma = gamma1; gs.g[1].dgamma = dgamma1; gs.g[1].dbeta = dbeta1; gs.g[1].map = RowMap{map1_s_in, map1_s_out, map1_off};
    gs.g[1].rows = rows1; gs.g[1].row_start = row1_start; gs.g[1].site = site1;
    int nb;
    split_blocks(rows0, rows1, 148 * 3, &gs.nb0, &nb);
    return launch_ln_bwd_fast(dy, s, mean, rstd, dres, dsub, dt, acc_res, gs, nb, dc, stream);
}

extern "C" int ick_add_ln_fwd(const void* x, void* sub, const float* gamma, const float* beta, void* y, float* mean,
                              float* rstd, int dt, int rows, int d, int ldx, int lds, int ldy, float eps, int map_s_in,
                              int map_s_out, int map_off, float drop_p, unsigned seed, unsigned site, cudaStream_t stream) {
    ICK_REQUIRE(rows >= 0 && d > 0 && d % 2 == 0 && d <= 2 * MAXP * 32, "add_ln_fwd: d=%d must be even and <= 512", d);
    ICK_REQUIRE(ldx % 2 == 0 && lds % 2 == 0 && ldy % 2 == 0 && lds >= d && ldy >= d, "add_ln_fwd: bad leading dims");
    ICK_REQUIRE(ldy <= 2 * MAXP * 32, "add_ln_fwd: ldy too large");
    if (rows == 0) return ICK_OK;
    RowMap m{map_s_in, map_s_out, map_off};
    DropCfg dc = make_drop(drop_p, seed, site);
    if (d == FAST_D && ldx == FAST_LD && lds == FAST_LD && ldy == FAST_LD && (dt == ICK_F32 || dt == ICK_BF16)) {
        LnGroups gs = {};
        gs.g[0].gamma = gamma; gs.g[0].beta = beta; gs.g[0].map = m; gs.g[0].rows = rows; gs.g[0].row_start = 0; gs.g[0].site = dc.site;
        gs.nb0 = min((rows + 7) / 8, 148 * 4);
        return launch_ln_fwd_fast(x, sub, y, mean, rstd, dt, eps, gs, gs.nb0, dc, stream);
    }
    const int blocks = min((rows + 7) / 8, 148 * 8);
    if (dt == ICK_F32)
        ick_launch(add_ln_fwd_kernel<float>, blocks, 256, 0, stream)((const float*)x, (float*)sub, gamma, beta, (float*)y, mean, rstd, rows,
                                                             d, ldx, lds, ldy, eps, m, dc);
    else if (dt == ICK_BF16)
        ick_launch(add_ln_fwd_kernel<bf16>, blocks, 256, 0, stream)((const bf16*)x, (bf16*)sub, gamma, beta, (bf16*)y, mean, rstd, rows, d,
                                                            ldx, lds, ldy, eps, m, dc);
    else {
        ick_set_error("add_ln_fwd: bad dtype %d", dt);
        return ICK_ERR_UNSUPPORTED;
    }
    return ick_check_launch("add_ln_fwd");
}

extern "C" int ick_add_ln_bwd(const void* dy, const void* s, const float* mean, const float* rstd, const float* gamma,
                              void* dres, void* dsub, float* dgamma, float* dbeta, int dt, int rows, int d, int lddy, int lds,
                              int ldres, int ldsub, int map_s_in, int map_s_out, int map_off, int acc_res, float drop_p,
                              unsigned seed, unsigned site, cudaStream_t stream) {
    ICK_REQUIRE(rows >= 0 && d > 0 && d % 2 == 0 && d <= 2 * MAXP * 32, "add_ln_bwd: d=%d must be even and <= 512", d);
    ICK_REQUIRE(lddy % 2 == 0 && lds % 2 == 0 && ldres % 2 == 0 && ldsub % 2 == 0, "add_ln_bwd: bad leading dims");
    ICK_REQUIRE(ldres <= 2 * MAXP * 32 && ldsub <= 2 * MAXP * 32, "add_ln_bwd: leading dims too large");
    if (rows == 0) return ICK_OK;
    RowMap m{map_s_in, map_s_out, map_off};
    DropCfg dc = make_drop(drop_p, seed, site);
    if (d == FAST_D && lddy == FAST_LD && lds == FAST_LD && (dres == nullptr || ldres == FAST_LD) && (dsub == nullptr || ldsub == FAST_LD) &&
        (dt == ICK_F32 || dt == ICK_BF16)) {
        LnGroups gs = {};
        gs.g[0].gamma = gamma; gs.g[0].dgamma = dgamma; gs.g[0].dbeta = dbeta; gs.g[0].map = m; gs.g[0].rows = rows; gs.g[0].row_start = 0;
        gs.g[0].site = dc.site;
        gs.nb0 = min((rows + 7) / 8, 148 * 3);
        return launch_ln_bwd_fast(dy, s, mean, rstd, dres, dsub, dt, acc_res, gs, gs.nb0, dc, stream);
    }
    const int blocks = min((rows + 7) / 8, 148 * 2);
    if (dt == ICK_F32)
        ick_launch(add_ln_bwd_kernel<float>, blocks, 256, 0, stream)((const float*)dy, (const float*)s, mean, rstd, gamma, (float*)dres,
                                                             (float*)dsub, dgamma, dbeta, rows, d, lddy, lds, ldres, ldsub, m,
                                                             acc_res, dc);
    else if (dt == ICK_BF16)
        ick_launch(add_ln_bwd_kernel<bf16>, blocks, 256, 0, stream)((const bf16*)dy, (const bf16*)s, mean, rstd, gamma, (bf16*)dres,
                                                            (bf16*)dsub, dgamma, dbeta, rows, d, lddy, lds, ldres, ldsub, m, acc_res,
                                                            dc);
    else {
        ick_set_error("add_ln_bwd: bad dtype %d", dt);
        return ICK_ERR_UNSUPPORTED;
    }
    return ick_check_launch("add_ln_bwd");
}
