// Pointer heads (fc_entity / fc_fact over h * ctx, get_scores K/models.py:440-452) as per-image bf16 GEMMs on the tensor
// cores (mma.sync m16n8k16, fp32 accumulate):
//   fwd : scores[b,t,col0+s] = bias + mask(t,s) * sum_d (h[b,t,d] w[d]) ctx[b,s,d]           one CTA per (64 slots, image)
//   bwd : G  = (mask*dS)   ctx   -> dH[b,t,:]   += w * G,   dw += sum_{b,t} h * G              one CTA per (64 features, image)
//         C2 = (mask*dS)^T h     -> dCtx[b,s,:] += w * C2,  dbias += sum dS (unmasked)
// mask(t,s) = first_t[b,s] < t + lag for the fact head (the indicator multiplies fc_fact's INPUT, so the bias survives), 1 for
// the entity head.  Operand tiles are staged in shared memory with row strides that are odd multiples of 16 bytes, which makes
// every ldmatrix conflict-free; (mask*dS)^T is never materialised - its A fragments come from ldmatrix.trans of the dS tile.
#include <stdlib.h>

#include "mma.cuh"
#include "pointer_internal.h"

namespace {

constexpr int NT = 256;   // threads per CTA (8 warps)
constexpr int CW = 64;    // slots per CTA of the forward kernel
constexpr int CWB = 32;   // feature columns per CTA of the backward kernel: 103 KB of tiles for the 301-slot head -> two CTAs per SM
constexpr int CLD = CWB + 8;
constexpr int SMEM_MAX = 232448;

__device__ __forceinline__ uint4 zero4() { return make_uint4(0u, 0u, 0u, 0u); }

// A fragment (16 rows x 16 k) of a row-major [rows][ld] tile: rows r0.., k columns k0..
__device__ __forceinline__ void a_frag(uint32_t* a, const bf16* tile, int ld, int r0, int k0, int lane) {
    ick_ldsm_x4(a[0], a[1], a[2], a[3], ick_smem_u32(tile + (size_t)(r0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * ld + k0 + 8 * (lane >> 4)));
}
// A fragment of the TRANSPOSE of a row-major tile: A[m][k] = tile[k0 + k][m0 + m]
__device__ __forceinline__ void a_frag_t(uint32_t* a, const bf16* tile, int ld, int m0, int k0, int lane) {
    const int mi = lane >> 3;
    ick_ldsm_x4_trans(a[0], a[1], a[2], a[3], ick_smem_u32(tile + (size_t)(k0 + (lane & 7) + 8 * (mi >> 1)) * ld + m0 + 8 * (mi & 1)));
}
// B fragments of two n-tiles (16 n) from a tile stored [n][k] (k contiguous): {b0,b1} of n-tile 0, {b0,b1} of n-tile 1
__device__ __forceinline__ void b_frag_nk(uint32_t* r, const bf16* tile, int ld, int n0, int k0, int lane) {
    ick_ldsm_x4(r[0], r[1], r[2], r[3], ick_smem_u32(tile + (size_t)(n0 + (lane & 7) + 8 * (lane >> 4)) * ld + k0 + 8 * ((lane >> 3) & 1)));
}
// same from a tile stored [k][n] (n contiguous)
__device__ __forceinline__ void b_frag_kn(uint32_t* r, const bf16* tile, int ld, int n0, int k0, int lane) {
    ick_ldsm_x4_trans(r[0], r[1], r[2], r[3], ick_smem_u32(tile + (size_t)(k0 + (lane & 7) + 8 * ((lane >> 3) & 1)) * ld + n0 + 8 * (lane >> 4)));
}

__global__ void __launch_bounds__(NT) pointer_fwd_mma_kernel(const bf16* __restrict__ h, const bf16* __restrict__ ctx, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const int* __restrict__ first_t,
                                                             float* __restrict__ scores, int Tn, int t0, int S, int D, int ld, int lds,
                                                             int col0, int lag, int Tp, int KP, int beams) {
    // beams != 0 (beam-search decode): the Tn rows of image b are its beams at the same time step t0 - one first_t row each
    ick_pdl_entry();
    extern __shared__ __align__(16) uint8_t smem[];
    const int PLD = KP + 8;
    bf16* hw = reinterpret_cast<bf16*>(smem);  // [Tp][PLD]  h * w
    bf16* cs = hw + (size_t)Tp * PLD;           // [CW][PLD]  ctx rows of this slot tile
    const int b = blockIdx.y, s0 = blockIdx.x * CW;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int kc = KP / 8;
    for (int idx = threadIdx.x; idx < Tp * kc; idx += NT) {
        const int t = idx / kc, c = (idx % kc) * 8;
        uint4 out = zero4();
        if (t < Tn) {
            float x[8];
            ld8(h + ((size_t)b * Tn + t) * ld + c, x);
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = c + i < D ? x[i] * __ldg(w + c + i) : 0.f;
            st8(reinterpret_cast<bf16*>(&out), x);
        }
        *reinterpret_cast<uint4*>(hw + (size_t)t * PLD + c) = out;
    }
    // ctx rows of the slot tile: four 16-byte chunks per thread requested before the first shared-memory store (one load per trip
    // left the staging latency-bound: 2 TB/s on the 64-row tiles of a decode step)
    for (int idx0 = threadIdx.x; idx0 < CW * kc; idx0 += 4 * NT) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * NT, s = idx / kc, c = (idx % kc) * 8;
            v[u] = (idx < CW * kc && s0 + s < S) ? __ldg(reinterpret_cast<const uint4*>(ctx + ((size_t)b * S + s0 + s) * ld + c)) : zero4();
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * NT, s = idx / kc, c = (idx % kc) * 8;
            if (idx < CW * kc) *reinterpret_cast<uint4*>(cs + (size_t)s * PLD + c) = v[u];
        }
    }
    __syncthreads();
    const float bv = bias[0];
    for (int mt = warp; mt < Tp / 16; mt += NT / 32) {
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        for (int k0 = 0; k0 < KP; k0 += 16) {
            uint32_t a[4];
            a_frag(a, hw, PLD, 16 * mt, k0, lane);
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t r[4];
                b_frag_nk(r, cs, PLD, 16 * np, k0, lane);
                ick_mma16816(acc[2 * np], a, r[0], r[1]);
                ick_mma16816(acc[2 * np + 1], a, r[2], r[3]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int t = 16 * mt + g + (x >> 1) * 8, s = s0 + 8 * j + 2 * tq + (x & 1);
                if (t < Tn && s < S) {
                    const bool on = first_t == nullptr || (beams ? first_t[((size_t)b * Tn + t) * S + s] < t0 + lag
                                                                 : first_t[(size_t)b * S + s] < t0 + t + lag);
                    scores[((size_t)b * Tn + t) * lds + col0 + s] = (on ? acc[j][x] : 0.f) + bv;
                }
            }
    }
}

__global__ void __launch_bounds__(NT) pointer_bwd_mma_kernel(const bf16* __restrict__ dS, const bf16* __restrict__ h, const bf16* __restrict__ ctx,
                                                             const float* __restrict__ w, const int* __restrict__ first_t,
                                                             float* __restrict__ dCtx, bf16* __restrict__ dH, float* __restrict__ gflat,
                                                             int w_off, int bias_off, int T, int S, int D, int ld, int ldds, int col0, int lag,
                                                             int Tp, int Sp) {
    ick_pdl_entry();
    extern __shared__ __align__(16) uint8_t smem[];
    const int DLD = Sp + 8;
    bf16* ds = reinterpret_cast<bf16*>(smem);  // [Tp][DLD]  mask * dS of this image
    bf16* cs = ds + (size_t)Tp * DLD;           // [Sp][CLD]  ctx columns d0..d0+CWB-1
    bf16* hs = cs + (size_t)Sp * CLD;           // [Tp][CLD]  h   columns d0..d0+CWB-1
    float* red = reinterpret_cast<float*>(hs + (size_t)Tp * CLD);  // [8][CWB] per-warp dw partials
    int* ft = reinterpret_cast<int*>(red + 8 * CWB);                // [Sp]
    const int d0 = blockIdx.x * CWB, b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    for (int s = threadIdx.x; s < Sp; s += NT) ft[s] = (first_t != nullptr && s < S) ? first_t[(size_t)b * S + s] : -0x40000000;
    // ctx / h column slices: up to four 16-byte loads requested per thread before the first shared-memory store
    for (int idx0 = threadIdx.x; idx0 < Sp * (CWB / 8); idx0 += 4 * NT) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * NT, s = idx / (CWB / 8), c = (idx % (CWB / 8)) * 8;
            v[u] = (idx < Sp * (CWB / 8) && s < S && d0 + c < ld) ? __ldg(reinterpret_cast<const uint4*>(ctx + ((size_t)b * S + s) * ld + d0 + c)) : zero4();
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * NT, s = idx / (CWB / 8), c = (idx % (CWB / 8)) * 8;
            if (idx < Sp * (CWB / 8)) *reinterpret_cast<uint4*>(cs + (size_t)s * CLD + c) = v[u];
        }
    }
    for (int idx0 = threadIdx.x; idx0 < Tp * (CWB / 8); idx0 += 2 * NT) {
        uint4 v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int idx = idx0 + u * NT, t = idx / (CWB / 8), c = (idx % (CWB / 8)) * 8;
            v[u] = (idx < Tp * (CWB / 8) && t < T && d0 + c < ld) ? __ldg(reinterpret_cast<const uint4*>(h + ((size_t)b * T + t) * ld + d0 + c)) : zero4();
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int idx = idx0 + u * NT, t = idx / (CWB / 8), c = (idx % (CWB / 8)) * 8;
            if (idx < Tp * (CWB / 8)) *reinterpret_cast<uint4*>(hs + (size_t)t * CLD + c) = v[u];
        }
    }
    __syncthreads();  // ft visible
    // dS tile (+ mask); the slice may start at an odd column (col0 = V + E), so pairs are used only when 4-byte aligned
    float bsum = 0.f;
    const bool pair_ok = ((col0 | ldds) & 1) == 0;
    const int sp2 = Sp / 2;
    // Unmasked head whose slice starts on a 16-byte boundary (the entity head: col0 = V): 8 gradients per load, four loads in
    // flight per thread.  The tail chunk is cut at S (the columns behind it belong to the next head).
    const bool vec_ok = first_t == nullptr && ((col0 | ldds) & 7) == 0 && ((reinterpret_cast<uintptr_t>(dS) & 15) == 0);
    if (vec_ok) {
        // EIGHT 16-byte loads requested per thread before the first is used (ncu source view, profiles/README.md r03b: a third of the
        // kernel's stall samples sat on the first use of each single load of this loop); the bias-gradient sum only where it is used
        const int sp8 = Sp / 8;
        const bool want_bsum = blockIdx.x == 0;
        for (int idx0 = threadIdx.x; idx0 < Tp * sp8; idx0 += 8 * NT) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = idx0 + u * NT, t = idx / sp8, s8 = (idx % sp8) * 8;
                v[u] = (idx < Tp * sp8 && t < T && s8 < S) ? __ldg(reinterpret_cast<const uint4*>(dS + ((size_t)b * T + t) * ldds + col0 + s8)) : zero4();
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = idx0 + u * NT, t = idx / sp8, s8 = (idx % sp8) * 8;
                if (idx >= Tp * sp8) break;
                if (s8 + 8 > S || want_bsum) {  // ragged tail chunk of a row (the columns behind S belong to the next head), or the sum
                    __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&v[u]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float2 f = __bfloat1622float2(hp[i]);
                        if (s8 + 2 * i >= S) f.x = 0.f;
                        if (s8 + 2 * i + 1 >= S) f.y = 0.f;
                        if (s8 + 8 > S) hp[i] = __floats2bfloat162_rn(f.x, f.y);
                        bsum += f.x + f.y;
                    }
                }
                *reinterpret_cast<uint4*>(ds + (size_t)t * DLD + s8) = v[u];
            }
        }
    } else
    for (int idx = threadIdx.x; idx < Tp * sp2; idx += NT) {
        const int t = idx / sp2, s = (idx % sp2) * 2;
        float v0 = 0.f, v1 = 0.f;
        if (t < T) {
            const bf16* src = dS + ((size_t)b * T + t) * ldds + col0 + s;
            if (pair_ok && s + 1 < S) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(src));
                v0 = f.x;
                v1 = f.y;
            } else {
                if (s < S) v0 = __bfloat162float(src[0]);
                if (s + 1 < S) v1 = __bfloat162float(src[1]);
            }
            bsum += v0 + v1;
            if (!(ft[s] < t + lag)) v0 = 0.f;
            if (!(ft[s + 1] < t + lag)) v1 = 0.f;
        }
        *reinterpret_cast<uint32_t*>(ds + (size_t)t * DLD + s) = ick_pack2(v0, v1);
    }
    if (blockIdx.x == 0) {  // dbias = sum of the UNMASKED dS
        bsum = warp_sum(bsum);
        if (lane == 0) atomicAdd(gflat + bias_off, bsum);
    }
    __syncthreads();

    // ---- G = ds * cs  (T x 64), K = slots ---------------------------------------------------------------------------------
    float dwp[CWB / 8][2];
#pragma unroll
    for (int j = 0; j < CWB / 8; ++j) dwp[j][0] = dwp[j][1] = 0.f;
    for (int mt = warp; mt < Tp / 16; mt += NT / 32) {
        float acc[CWB / 8][4];
#pragma unroll
        for (int j = 0; j < CWB / 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        for (int k0 = 0; k0 < Sp; k0 += 16) {
            uint32_t a[4];
            a_frag(a, ds, DLD, 16 * mt, k0, lane);
#pragma unroll
            for (int np = 0; np < CWB / 16; ++np) {
                uint32_t r[4];
                b_frag_kn(r, cs, CLD, 16 * np, k0, lane);
                ick_mma16816(acc[2 * np], a, r[0], r[1]);
                ick_mma16816(acc[2 * np + 1], a, r[2], r[3]);
            }
        }
        // dH += w * G as a read-modify-write: ALL sixteen old values of the lane are requested before the first store (a load
        // placed after a store through the same base pointer cannot be hoisted by the compiler, which serialised 16 round trips)
        __nv_bfloat162 oldv[CWB / 8][2];
#pragma unroll
        for (int j = 0; j < CWB / 8; ++j)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int d = d0 + 8 * j + 2 * tq, t = 16 * mt + g + 8 * hh;
                oldv[j][hh] = (d < D && t < T) ? *reinterpret_cast<const __nv_bfloat162*>(dH + ((size_t)b * T + t) * ld + d)
                                               : __floats2bfloat162_rn(0.f, 0.f);
            }
#pragma unroll
        for (int j = 0; j < CWB / 8; ++j) {
            const int c = 8 * j + 2 * tq, d = d0 + c;
            if (d >= D) continue;  // D is even (d-model 300): a column pair is inside or outside together
            const float w0 = __ldg(w + d), w1 = d + 1 < D ? __ldg(w + d + 1) : 0.f;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int t = 16 * mt + g + 8 * hh;
                if (t >= T) continue;
                const float g0 = acc[j][2 * hh], g1 = acc[j][2 * hh + 1];
                const float2 hv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(hs + (size_t)t * CLD + c));
                dwp[j][0] = fmaf(hv.x, g0, dwp[j][0]);
                dwp[j][1] = fmaf(hv.y, g1, dwp[j][1]);
                const float2 old = __bfloat1622float2(oldv[j][hh]);
                *reinterpret_cast<__nv_bfloat162*>(dH + ((size_t)b * T + t) * ld + d) = __floats2bfloat162_rn(old.x + w0 * g0, old.y + w1 * g1);
            }
        }
    }
    // dw: reduce the per-thread partials over the 8 row groups of the warp, then over warps through shared memory
#pragma unroll
    for (int j = 0; j < CWB / 8; ++j)
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            float v = dwp[j][x];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g == 0) red[warp * CWB + 8 * j + 2 * tq + x] = v;
        }
    __syncthreads();
    if (threadIdx.x < CWB && d0 + threadIdx.x < D) {
        float v = 0.f;
#pragma unroll
        for (int wi = 0; wi < NT / 32; ++wi) v += red[wi * CWB + threadIdx.x];
        atomicAdd(gflat + w_off + d0 + threadIdx.x, v);
    }

    // ---- C2 = ds^T * hs  (S x 64), K = time steps ------------------------------------------------------------------------
    for (int mt = warp; mt < Sp / 16; mt += NT / 32) {
        float acc[CWB / 8][4];
#pragma unroll
        for (int j = 0; j < CWB / 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        for (int k0 = 0; k0 < Tp; k0 += 16) {
            uint32_t a[4];
            a_frag_t(a, ds, DLD, 16 * mt, k0, lane);
#pragma unroll
            for (int np = 0; np < CWB / 16; ++np) {
                uint32_t r[4];
                b_frag_kn(r, hs, CLD, 16 * np, k0, lane);
                ick_mma16816(acc[2 * np], a, r[0], r[1]);
                ick_mma16816(acc[2 * np + 1], a, r[2], r[3]);
            }
        }
        float2 oldc[CWB / 8][2];  // same batching of the read-modify-write of dCtx
#pragma unroll
        for (int j = 0; j < CWB / 8; ++j)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int d = d0 + 8 * j + 2 * tq, s = 16 * mt + g + 8 * hh;
                oldc[j][hh] = (d < D && s < S) ? *reinterpret_cast<const float2*>(dCtx + ((size_t)b * S + s) * ld + d) : make_float2(0.f, 0.f);
            }
#pragma unroll
        for (int j = 0; j < CWB / 8; ++j) {
            const int d = d0 + 8 * j + 2 * tq;
            if (d >= D) continue;
            const float w0 = __ldg(w + d), w1 = d + 1 < D ? __ldg(w + d + 1) : 0.f;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int s = 16 * mt + g + 8 * hh;
                if (s >= S) continue;
                *reinterpret_cast<float2*>(dCtx + ((size_t)b * S + s) * ld + d) =
                    make_float2(oldc[j][hh].x + w0 * acc[j][2 * hh], oldc[j][hh].y + w1 * acc[j][2 * hh + 1]);
            }
        }
    }
}

template <typename K>
int set_smem(K kernel) {
    static bool done = false;
    if (!done) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX) != cudaSuccess) {
            ick_set_error("pointer: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
            return ICK_ERR_CUDA;
        }
        done = true;
    }
    return ICK_OK;
}

// Single greedy decode step (one row per image): scores[b, col0 + s] = bias + mask * sum_d (h[b, d] w[d]) ctx[b, s, d].  Per step every
// image re-reads all of its context rows (120 MB for 625 images x 301 entities) against two FLOPs per element, so this is a
// streaming matrix-vector product: 8 lanes share a 640-byte ctx row (consecutive lanes read consecutive 16-byte chunks), a thread
// requests all ten of its chunks (two rows) before the first use, the query vector h * w sits in shared memory in fp32, and an
// 8-lane shuffle tree finishes each row's dot product.  The tensor-core kernel above stages 64 rows through shared memory behind a
// barrier for a 16 x 64 x 320 product of which one row is real (2 TB/s, 36 us per launch); this one is +2.9 % on the whole greedy
// decode (24.0k vs 23.3k captions/s, profiles/README.md r02h).  For the G = 5 beam rows of an image the staged tensor-core kernel
// is the faster one (38 vs 63 us per launch), so beam search keeps it; the template parameter stays for that measurement.
constexpr int PS_W = 320;    // row width in elements (d_model padded): 40 chunks of 8
constexpr int PS_ROWS = 64;  // ctx rows per CTA: 8 warps x 4 rows x 2 passes
template <int G>
__global__ void __launch_bounds__(NT, 3) pointer_step_kernel(const bf16* __restrict__ h, const bf16* __restrict__ ctx, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const int* __restrict__ first_t,
                                                          float* __restrict__ scores, int t0, int S, int D, int lds, int col0, int lag,
                                                          int beams) {
    ick_pdl_entry();
    __shared__ __align__(16) float hw[G][PS_W];
    const int b = blockIdx.y, s0 = blockIdx.x * PS_ROWS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, l8 = lane & 7, rq = lane >> 3;
    // this thread's ctx chunks first: they are independent of hw and by far the longest latency
    uint4 v[2][5];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int sr = s0 + pass * 32 + warp * 4 + rq;
#pragma unroll
        for (int i = 0; i < 5; ++i) v[pass][i] = sr < S ? __ldg(reinterpret_cast<const uint4*>(ctx + ((size_t)b * S + sr) * PS_W) + l8 + 8 * i) : zero4();
    }
    for (int idx = threadIdx.x; idx < G * PS_W; idx += NT) {
        const int g = idx / PS_W, c = idx % PS_W;
        hw[g][c] = c < D ? __bfloat162float(h[((size_t)b * G + g) * PS_W + c]) * __ldg(w + c) : 0.f;
    }
    __syncthreads();
    const float bv = bias[0];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int sr = s0 + pass * 32 + warp * 4 + rq;
        float acc[G];
#pragma unroll
        for (int g = 0; g < G; ++g) acc[g] = 0.f;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const __nv_bfloat162* pr = reinterpret_cast<const __nv_bfloat162*>(&v[pass][i]);
            float x[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = __bfloat1622float2(pr[k]);
                x[2 * k] = f.x;
                x[2 * k + 1] = f.y;
            }
            const int c0 = (l8 + 8 * i) * 8;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float4 a = *reinterpret_cast<const float4*>(&hw[g][c0]);
                const float4 c = *reinterpret_cast<const float4*>(&hw[g][c0 + 4]);
                acc[g] = fmaf(x[0], a.x, acc[g]); acc[g] = fmaf(x[1], a.y, acc[g]); acc[g] = fmaf(x[2], a.z, acc[g]); acc[g] = fmaf(x[3], a.w, acc[g]);
                acc[g] = fmaf(x[4], c.x, acc[g]); acc[g] = fmaf(x[5], c.y, acc[g]); acc[g] = fmaf(x[6], c.z, acc[g]); acc[g] = fmaf(x[7], c.w, acc[g]);
                // keeps the compiler from hoisting all 2 * 5 * G shared-memory loads of a pass above the arithmetic (254 registers,
                // one CTA per SM: too few bytes in flight); the ctx chunks in v[][] are registers and are not affected
                asm volatile("" ::: "memory");
            }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            acc[g] += __shfl_xor_sync(0xffffffffu, acc[g], 1);
            acc[g] += __shfl_xor_sync(0xffffffffu, acc[g], 2);
            acc[g] += __shfl_xor_sync(0xffffffffu, acc[g], 4);
        }
        if (l8 == 0 && sr < S) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const bool on = first_t == nullptr || (beams ? first_t[((size_t)b * G + g) * S + sr] < t0 + lag : first_t[(size_t)b * S + sr] < t0 + g + lag);
                scores[((size_t)b * G + g) * lds + col0 + sr] = (on ? acc[g] : 0.f) + bv;
            }
        }
    }
}

bool ptr_step() {  // ICK_PTR_STEP=0: single-step pointer heads stay on the staged tensor-core kernel (A/B aid)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ICK_PTR_STEP");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

}  // namespace

int ick_pointer_fwd_mma(const void* h, const void* ctx, const float* w, const float* bias, const int* first_t, float* scores, int B, int Tn,
                        int t0, int S, int D, int ld, int ldscores, int col0, int lag, int beams, cudaStream_t stream) {
    static const bool step_beams = getenv("ICK_PTR_STEP_BEAMS") != nullptr;  // A/B aid: the streaming kernel for beam rows too
    if (Tn <= 8 && (Tn == 1 || (beams && step_beams)) && ld == PS_W && D <= PS_W && ptr_step() && (((uintptr_t)h | (uintptr_t)ctx) & 15) == 0) {
        dim3 gs((S + PS_ROWS - 1) / PS_ROWS, B);
#define ICK_PS_CASE(GG)                                                                                                                 \
    case GG:                                                                                                                            \
        ick_launch(pointer_step_kernel<GG>, gs, NT, 0, stream)((const bf16*)h, (const bf16*)ctx, w, bias, first_t, scores, t0, S, D, ldscores, col0, \
                                                               lag, beams);                                                             \
        break;
        switch (Tn) {
            ICK_PS_CASE(1)
            ICK_PS_CASE(2)
            ICK_PS_CASE(3)
            ICK_PS_CASE(4)
            ICK_PS_CASE(5)
            ICK_PS_CASE(6)
            ICK_PS_CASE(7)
            ICK_PS_CASE(8)
        }
#undef ICK_PS_CASE
        return ick_check_launch("pointer_fwd(step)");
    }
    const int Tp = (Tn + 15) / 16 * 16, KP = (D + 15) / 16 * 16;
    const size_t smem = (size_t)(Tp + CW) * (KP + 8) * 2;
    if (KP > ld || ld % 8 != 0 || smem > SMEM_MAX || (((uintptr_t)h | (uintptr_t)ctx) & 15) != 0) return ICK_ERR_UNSUPPORTED;
    int rc = set_smem(pointer_fwd_mma_kernel);
    if (rc) return rc;
    dim3 grid((S + CW - 1) / CW, B);
    ick_launch(pointer_fwd_mma_kernel, grid, NT, smem, stream)((const bf16*)h, (const bf16*)ctx, w, bias, first_t, scores, Tn, t0, S, D, ld, ldscores,
                                                       col0, lag, Tp, KP, beams);
    return ick_check_launch("pointer_fwd_mma");
}

int ick_pointer_bwd_mma(const void* dS, const void* h, const void* ctx, const float* w, const int* first_t, float* dCtx, void* dH, float* gflat,
                        int w_off, int bias_off, int B, int T, int S, int D, int ld, int ldds, int col0, int lag, cudaStream_t stream) {
    const int Tp = (T + 15) / 16 * 16, Sp = (S + 15) / 16 * 16;
    const size_t smem = ((size_t)Tp * (Sp + 8) + (size_t)(Sp + Tp) * CLD) * 2 + 8 * CWB * 4 + (size_t)Sp * 4;
    if (ld % 8 != 0 || (D & 1) != 0 || smem > SMEM_MAX || (((uintptr_t)h | (uintptr_t)ctx | (uintptr_t)dCtx) & 15) != 0 || ((uintptr_t)dH & 3) != 0)
        return ICK_ERR_UNSUPPORTED;
    int rc = set_smem(pointer_bwd_mma_kernel);
    if (rc) return rc;
    dim3 grid((D + CWB - 1) / CWB, B);
    ick_launch(pointer_bwd_mma_kernel, grid, NT, smem, stream)((const bf16*)dS, (const bf16*)h, (const bf16*)ctx, w, first_t, dCtx, (bf16*)dH, gflat, w_off,
                                                       bias_off, T, S, D, ld, ldds, col0, lag, Tp, Sp);
    return ick_check_launch("pointer_bwd_mma");
}
