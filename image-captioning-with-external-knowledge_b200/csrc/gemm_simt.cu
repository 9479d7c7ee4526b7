// CUDA-core (FFMA) GEMMs with fp32 accumulation.  These are the fp32-parity-mode GEMMs and the shape-generic fallback
// inside the library (odd shapes); the bf16 hot path uses the tcgen05/TMEM kernels in gemm_tcgen05.cu.
//
//   gemm_tn : C[M,N] (+)= A[M,K] * W[N,K]^T (+ bias) with fused epilogues      (nn.Linear forward and dgrad)
//   wgrad   : G[rowoff[n] + colmap[k]] += sum_m dY[m,n] * X[m,k]                 (nn.Linear weight/bias gradients,
//             scattered straight into the flat fp32 master-gradient buffer through the packing maps)
#include "common.cuh"
#include "ickb200.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, LDS = 132;

template <typename T>
__device__ __forceinline__ void load_chunk8(const T* base, int row, int nrows, int k, int K, int ld, float* v) {
    if (row < nrows && k + 8 <= K) {
        ld8(base + (size_t)row * ld + k, v);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (row < nrows && k + i < K) ? to_f(base[(size_t)row * ld + k + i]) : 0.f;
    }
}

template <typename TC>
__device__ __forceinline__ void store4(TC* p, const float* v, bool vec, int nvalid) {
    if (vec && nvalid >= 4) {
        if constexpr (sizeof(TC) == 4) {
            *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            uint2 u;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
            h[0] = __floats2bfloat162_rn(v[0], v[1]);
            h[1] = __floats2bfloat162_rn(v[2], v[3]);
            *reinterpret_cast<uint2*>(p) = u;
        }
    } else {
        for (int i = 0; i < 4 && i < nvalid; ++i) p[i] = from_f<TC>(v[i]);
    }
}

// epilogue modes
constexpr int EPI_NONE = 0, EPI_RELU_DROP = 1, EPI_RELU_BWD = 2;

template <typename TA, typename TW, typename TC>
__global__ void __launch_bounds__(256) gemm_tn_kernel(const TA* __restrict__ A, const TW* __restrict__ W, TC* C,
                                                      const float* __restrict__ bias, const TC* __restrict__ aux, int M,
                                                      int N, int K, int lda, int ldw, int ldc, int ldaux, int epi,
                                                      int accumulate, DropCfg drop) {
    ick_pdl_entry();
    __shared__ __align__(16) float As[2][BK][LDS];
    __shared__ __align__(16) float Ws[2][BK][LDS];
    ick_resolve_seed(drop);
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int lrow = tid & 127, lk = (tid >> 7) * 8;
    const int tx = tid & 15, ty = tid >> 4;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[8], rw[8];
    const int nk = (K + BK - 1) / BK;
    load_chunk8(A, m0 + lrow, M, lk, K, lda, ra);
    load_chunk8(W, n0 + lrow, N, lk, K, ldw, rw);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        As[0][lk + i][lrow] = ra[i];
        Ws[0][lk + i][lrow] = rw[i];
    }
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            load_chunk8(A, m0 + lrow, M, (kt + 1) * BK + lk, K, lda, ra);
            load_chunk8(W, n0 + lrow, N, (kt + 1) * BK + lk, K, ldw, rw);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[8];
            *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
            *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Ws[cur][k][tx * 4]);
            *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Ws[cur][k][64 + tx * 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                As[cur ^ 1][lk + i][lrow] = ra[i];
                Ws[cur ^ 1][lk + i][lrow] = rw[i];
            }
        }
        __syncthreads();
    }

    const bool vec = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (row >= M) continue;
        const uint32_t rmix = ick_rowmix(drop.seed, drop.site, (uint64_t)row);
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int col = n0 + jh * 64 + tx * 4;
            if (col >= N) continue;
            const int nvalid = min(4, N - col);
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x = acc[i][jh * 4 + j];
                const int c = col + j;
                if (c < N) {
                    if (bias) x += bias[c];
                    if (accumulate) x += to_f(C[(size_t)row * ldc + c]);
                    if (epi == EPI_RELU_DROP) {
                        x = fmaxf(x, 0.f) * ick_drop_mul(drop, rmix, (uint32_t)c);
                    } else if (epi == EPI_RELU_BWD) {
                        x = (to_f(aux[(size_t)row * ldaux + c]) != 0.f) ? x * drop.inv_keep : 0.f;
                    }
                }
                v[j] = x;
            }
            store4(C + (size_t)row * ldc + col, v, vec, nvalid);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
template <typename TY, typename TX>
__global__ void __launch_bounds__(256) wgrad_kernel(const TY* __restrict__ dY, const TX* __restrict__ X, float* G,
                                                    const int* __restrict__ rowoff, const int* __restrict__ colmap,
                                                    const int* __restrict__ biasoff, int M, int N, int K, int ldy, int ldx,
                                                    int m_per_split) {
    ick_pdl_entry();
    __shared__ __align__(16) float Ys[BK][LDS];
    __shared__ __align__(16) float Xs[BK][LDS];
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * BN, k0 = blockIdx.y * BN;
    const int mbeg = blockIdx.z * m_per_split, mend = min(M, mbeg + m_per_split);
    const int lm = tid >> 4, lc = (tid & 15) * 8;
    const int tx = tid & 15, ty = tid >> 4;
    const bool do_bias = (biasoff != nullptr) && blockIdx.y == 0 && tx == 0;

    float acc[8][8], bsum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        bsum[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    }
    for (int mt = mbeg; mt < mend; mt += BK) {
        float ry[8], rx[8];
        const int m = mt + lm;
        // rows are the reduction dimension here: "row < nrows" guards m, the 8-chunk guards the column extent
        if (m < mend && n0 + lc + 8 <= N) ld8(dY + (size_t)m * ldy + n0 + lc, ry);
        else
#pragma unroll
            for (int i = 0; i < 8; ++i) ry[i] = (m < mend && n0 + lc + i < N) ? to_f(dY[(size_t)m * ldy + n0 + lc + i]) : 0.f;
        if (m < mend && k0 + lc + 8 <= K) ld8(X + (size_t)m * ldx + k0 + lc, rx);
        else
#pragma unroll
            for (int i = 0; i < 8; ++i) rx[i] = (m < mend && k0 + lc + i < K) ? to_f(X[(size_t)m * ldx + k0 + lc + i]) : 0.f;
        __syncthreads();
        *reinterpret_cast<float4*>(&Ys[lm][lc]) = make_float4(ry[0], ry[1], ry[2], ry[3]);
        *reinterpret_cast<float4*>(&Ys[lm][lc + 4]) = make_float4(ry[4], ry[5], ry[6], ry[7]);
        *reinterpret_cast<float4*>(&Xs[lm][lc]) = make_float4(rx[0], rx[1], rx[2], rx[3]);
        *reinterpret_cast<float4*>(&Xs[lm][lc + 4]) = make_float4(rx[4], rx[5], rx[6], rx[7]);
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < BK; ++mm) {
            float a[8], b[8];
            *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&Ys[mm][ty * 4]);
            *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&Ys[mm][64 + ty * 4]);
            *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Xs[mm][tx * 4]);
            *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Xs[mm][64 + tx * 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            if (do_bias)
#pragma unroll
                for (int i = 0; i < 8; ++i) bsum[i] += a[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = n0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (n >= N) continue;
        if (do_bias && biasoff[n] >= 0) atomicAdd(G + biasoff[n], bsum[i]);
        const int ro = rowoff[n];
        if (ro < 0) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (k >= K) continue;
            const int cm = colmap ? colmap[k] : k;
            if (cm >= 0) atomicAdd(G + ro + cm, acc[i][j]);
        }
    }
}

template <typename TA, typename TW, typename TC>
int launch_tn(const void* A, const void* W, void* C, const float* bias, const void* aux, int M, int N, int K, int lda,
              int ldw, int ldc, int ldaux, int epi, int accumulate, DropCfg drop, cudaStream_t st) {
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    ick_launch(gemm_tn_kernel<TA, TW, TC>, grid, 256, 0, st)((const TA*)A, (const TW*)W, (TC*)C, bias, (const TC*)aux, M, N, K, lda,
                                                     ldw, ldc, ldaux, epi, accumulate, drop);
    return ick_check_launch("gemm_tn_simt");
}

}  // namespace

extern "C" int ick_gemm_tn_simt(const void* A, int a_dt, const void* W, int w_dt, void* C, int c_dt, const float* bias,
                                const void* aux, int M, int N, int K, int lda, int ldw, int ldc, int ldaux, int epi,
                                int accumulate, float drop_p, unsigned seed, unsigned site, cudaStream_t stream) {
    ICK_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_tn: bad sizes M=%d N=%d K=%d", M, N, K);
    ICK_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, "gemm_tn: lda/ldw must be multiples of 8 (lda=%d ldw=%d)", lda, ldw);
    ICK_REQUIRE(epi >= 0 && epi <= 2, "gemm_tn: bad epilogue %d", epi);
    ICK_REQUIRE(epi != EPI_RELU_BWD || aux != nullptr, "gemm_tn: relu_bwd needs aux");
    if (M == 0) return ICK_OK;
    DropCfg d = make_drop(drop_p, seed, site);
#define ICK_GO(TA, TW, TC) return launch_tn<TA, TW, TC>(A, W, C, bias, aux, M, N, K, lda, ldw, ldc, ldaux, epi, accumulate, d, stream)
    if (a_dt == ICK_F32 && w_dt == ICK_F32 && c_dt == ICK_F32) ICK_GO(float, float, float);
    if (a_dt == ICK_BF16 && w_dt == ICK_BF16 && c_dt == ICK_BF16) ICK_GO(bf16, bf16, bf16);
    if (a_dt == ICK_BF16 && w_dt == ICK_BF16 && c_dt == ICK_F32) ICK_GO(bf16, bf16, float);
    if (a_dt == ICK_F32 && w_dt == ICK_BF16 && c_dt == ICK_BF16) ICK_GO(float, bf16, bf16);
#undef ICK_GO
    ick_set_error("gemm_tn: unsupported dtype combination a=%d w=%d c=%d", a_dt, w_dt, c_dt);
    return ICK_ERR_UNSUPPORTED;
}

extern "C" int ick_wgrad_simt(const void* dY, int y_dt, const void* X, int x_dt, float* gflat, const int* rowoff,
                              const int* colmap, const int* biasoff, int M, int N, int K, int ldy, int ldx,
                              cudaStream_t stream) {
    ICK_REQUIRE(M >= 0 && N > 0 && K > 0, "wgrad: bad sizes M=%d N=%d K=%d", M, N, K);
    ICK_REQUIRE(ldy % 8 == 0 && ldx % 8 == 0, "wgrad: ldy/ldx must be multiples of 8 (ldy=%d ldx=%d)", ldy, ldx);
    ICK_REQUIRE(rowoff != nullptr, "wgrad: rowoff is required");
    if (M == 0) return ICK_OK;
    const int tiles = ((N + BN - 1) / BN) * ((K + BN - 1) / BN);
    int splits = (148 * 4 + tiles - 1) / tiles;
    const int max_splits = (M + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int mps = (M + splits - 1) / splits;
    mps = (mps + BK - 1) / BK * BK;
    splits = (M + mps - 1) / mps;
    dim3 grid((N + BN - 1) / BN, (K + BN - 1) / BN, splits);
#define ICK_GO(TY, TX)                                                                                               \
    ick_launch(wgrad_kernel<TY, TX>, grid, 256, 0, stream)((const TY*)dY, (const TX*)X, gflat, rowoff, colmap, biasoff, M, N, K, \
                                                   ldy, ldx, mps);                                                   \
    return ick_check_launch("wgrad_simt")
    if (y_dt == ICK_F32 && x_dt == ICK_F32) { ICK_GO(float, float); }
    if (y_dt == ICK_BF16 && x_dt == ICK_BF16) { ICK_GO(bf16, bf16); }
    if (y_dt == ICK_F32 && x_dt == ICK_BF16) { ICK_GO(float, bf16); }
#undef ICK_GO
    ick_set_error("wgrad: unsupported dtype combination y=%d x=%d", y_dt, x_dt);
    return ICK_ERR_UNSUPPORTED;
}
