// Row-wise tail of a Transformer decoder layer for the KV-cached decode loops (greedy predict / beam search), one launch instead
// of up to seven.  After an attention kernel everything up to the next attention is row-local (nn.TransformerDecoderLayer,
// post-LN, G/models.py:241-242):
//
//     y  = LayerNorm(x + attn_out W_o^T + b_o)                                     (always)
//     y  = LayerNorm(y + relu(y W_1^T + b_1) W_2^T + b_2)                          (optional: the feed-forward sublayer)
//     p  = y W_n^T + b_n                                                            (optional: the projection that feeds the NEXT
//                                                                                    attention - cross-attention Q, or the next
//                                                                                    layer's self-attention Q|K|V cache row)
//
// In the decode loops these GEMMs have M = images (x beams) rows and N, K <= 960: each of them alone is a ~10 us latency chain
// (TMA -> tcgen05 -> epilogue on a handful of CTAs) and the step is a string of them.  Here a CTA owns 16 rows and walks the whole
// chain with its activations in shared memory; the weights (~1.5 MB for the longest chain, L2-resident) stream straight from
// global memory into mma.sync B fragments:
//   * a lane's 16-byte load of W[n][32 kb + 8 tq .. +7] is exactly its B fragments of TWO k-steps if the 32 k-values of a block are
//     taken in the order k = 8 tq + 4 s + {0,1 | 2,3} - a permutation of the reduction index that the A fragments (one 16-byte
//     shared-memory load per row) follow, so no ldmatrix / staging of the weights is needed and every sector read is fully used;
//   * the B fragments of the next two 32-k blocks are in flight while the current one is multiplied (register double buffer);
//   * warp w owns the n-tiles t = w, w + 8, ... (at most 8 per pass: 32 accumulator registers), LayerNorm is two rows per warp.
// bf16 operands, fp32 accumulation, fp32 LayerNorm - the arithmetic of the unfused kernels except that the sublayer output is not
// rounded to bf16 before the residual addition.
#include "common.cuh"
#include "ickb200.h"
#include "mma.cuh"

namespace {

constexpr int CH_ROWS = 16;   // rows per CTA (one m16 tile)
constexpr int CH_NT = 256;    // threads per CTA
constexpr int CH_PAD = 32;    // shared-memory row padding in elements: (width + 32) * 2 bytes = 16 words mod 32 -> conflict-free LDS.128

struct ChainParams {
    const bf16* a_in;  int lda;   // attention output rows
    const bf16* x_res; int ldx;   // residual rows
    const bf16* Wo; const float* bo; const float* g1; const float* be1;
    const bf16* W1; const float* b1; const bf16* W2; const float* b2; const float* g2; const float* be2;
    const bf16* Wn; const float* bn;
    bf16* y; int ldy;             // LayerNorm output of the last sublayer
    bf16* proj; int ldproj;       // projection output
    int rows, D, DP, FF, FFP, Nn;
    int ldwo, ldw1, ldw2, ldwn;
    float eps;
};

// C[16 x N] = A[16 x K] (shared, row stride sa) * W[N x K]^T (global, row stride ldw) + bias; K a multiple of 32, N of 8.
// mode 0: fp32 to shared `outF` (stride so);  mode 1: bf16 (+ReLU) to shared `outB` (stride so);  mode 2: bf16 to global rows.
//
// One pass: warp w owns the NTW n-tiles t0 + 8 j + w (j < NTW; tiles >= ntiles are skipped, only possible when !FULL).  The k loop
// is software-pipelined over a ring of DEPTH register buffers of B fragments with BATCHED loads: all fragments of block kb + DEPTH
// are requested right after block kb has been multiplied, so a buffer's loads sit behind one scoreboard that nothing newer shares (interleaving
// load j with mma j made every mma wait for a load issued a few instructions earlier - 2 us per 32-k block).
template <int NTW, bool FULL, int DEPTH>
__device__ __forceinline__ void chain_gemm_pass(const bf16* arow0, const bf16* arow1, int nkb, const bf16* __restrict__ W, int ldw,
                                                const float* __restrict__ bias, int t0, int ntiles, int mode, bool relu, float* outF, bf16* outB,
                                                int so, bf16* outG, int ldg, int r0, int rows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    float acc[NTW][4];
    const bf16* wp[NTW];
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
        const int t = t0 + j * 8 + warp;
        wp[j] = W + (size_t)((FULL || t < ntiles) ? (t * 8 + g) : g) * ldw + 8 * tq;  // a skipped tile multiplies tile 0 again (never stored)
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    }
    uint4 bq[DEPTH][NTW];  // DEPTH 32-k blocks of B fragments in flight
#pragma unroll
    for (int s = 0; s < DEPTH; ++s)
#pragma unroll
        for (int j = 0; j < NTW; ++j) bq[s][j] = s < nkb ? *reinterpret_cast<const uint4*>(wp[j] + 32 * s) : make_uint4(0u, 0u, 0u, 0u);
    for (int kb2 = 0; kb2 < nkb; kb2 += DEPTH) {
#pragma unroll
        for (int cur = 0; cur < DEPTH; ++cur) {  // the register ring is indexed statically
            const int kb = kb2 + cur;
            if (kb < nkb) {
                const uint4 lo = *reinterpret_cast<const uint4*>(arow0 + 32 * kb);
                const uint4 hi = *reinterpret_cast<const uint4*>(arow1 + 32 * kb);
                const uint32_t a0[4] = {lo.x, hi.x, lo.y, hi.y}, a1[4] = {lo.z, hi.z, lo.w, hi.w};
#pragma unroll
                for (int j = 0; j < NTW; ++j) ick_mma16816(acc[j], a0, bq[cur][j].x, bq[cur][j].y);
#pragma unroll
                for (int j = 0; j < NTW; ++j) ick_mma16816(acc[j], a1, bq[cur][j].z, bq[cur][j].w);
                if (kb + DEPTH < nkb) {
#pragma unroll
                    for (int j = 0; j < NTW; ++j) bq[cur][j] = *reinterpret_cast<const uint4*>(wp[j] + 32 * (kb + DEPTH));
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NTW; ++j) {
        const int t = t0 + j * 8 + warp;
        if (!FULL && t >= ntiles) continue;
        const int n = t * 8 + 2 * tq;
        const float bx = bias ? bias[n] : 0.f, by = bias ? bias[n + 1] : 0.f;
        float v00 = acc[j][0] + bx, v01 = acc[j][1] + by, v10 = acc[j][2] + bx, v11 = acc[j][3] + by;
        if (relu) {
            v00 = fmaxf(v00, 0.f); v01 = fmaxf(v01, 0.f); v10 = fmaxf(v10, 0.f); v11 = fmaxf(v11, 0.f);
        }
        if (mode == 0) {
            *reinterpret_cast<float2*>(outF + (size_t)g * so + n) = make_float2(v00, v01);
            *reinterpret_cast<float2*>(outF + (size_t)(g + 8) * so + n) = make_float2(v10, v11);
        } else if (mode == 1) {
            st2(outB + (size_t)g * so + n, v00, v01);
            st2(outB + (size_t)(g + 8) * so + n, v10, v11);
        } else {
            if (r0 + g < rows) st2(outG + (size_t)(r0 + g) * ldg + n, v00, v01);
            if (r0 + g + 8 < rows) st2(outG + (size_t)(r0 + g + 8) * ldg + n, v10, v11);
        }
    }
}

__device__ __forceinline__ void chain_gemm(const bf16* As, int sa, int K, const bf16* __restrict__ W, int ldw, const float* __restrict__ bias,
                                           int N, int mode, bool relu, float* outF, bf16* outB, int so, bf16* outG, int ldg, int r0, int rows) {
    const int lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
    const int ntiles = N >> 3, nkb = K >> 5;
    const bf16* arow0 = As + (size_t)g * sa + 8 * tq;
    const bf16* arow1 = As + (size_t)(g + 8) * sa + 8 * tq;
#define ICK_PASS(NTW, FULL) chain_gemm_pass<NTW, FULL, (NTW <= 5 ? 5 : 3)>(arow0, arow1, nkb, W, ldw, bias, t0, ntiles, mode, relu, outF, outB, so, outG, ldg, r0, rows)
    int t0 = 0;
    for (; ntiles - t0 >= 64; t0 += 64) ICK_PASS(8, true);
    const int rem = ntiles - t0;  // < 64 tiles left: the shapes of the reference leave 40 (N = 320) or 56 (N = 960)
    if (rem == 40) ICK_PASS(5, true);
    else if (rem == 56) ICK_PASS(7, true);
    else if (rem > 32) ICK_PASS(8, false);
    else if (rem > 0) ICK_PASS(4, false);
#undef ICK_PASS
}

// y = LayerNorm(res + sub) over the first D columns (pad columns written as zero): bf16 to shared (stride sy) and, optionally, to
// the global rows.  res: bf16 shared (stride sr); sub: fp32 shared (stride ss).  Two rows per warp.
__device__ __forceinline__ void chain_ln(const bf16* res, int sr, const float* sub, int ss, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, int D, int DP, float eps, bf16* ys, int sy, bf16* yg, int ldy, int r0,
                                         int rows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < CH_ROWS; r += CH_NT / 32) {
        float v[12];  // DP <= 384
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int c = lane + 32 * i;
            v[i] = (c < D) ? __bfloat162float(res[(size_t)r * sr + c]) + sub[(size_t)r * ss + c] : 0.f;
            sum += v[i];
        }
        const float mean = warp_sum(sum) / (float)D;
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int c = lane + 32 * i;
            const float d = (c < D) ? v[i] - mean : 0.f;
            var = fmaf(d, d, var);
        }
        const float rstd = rsqrtf(warp_sum(var) / (float)D + eps);
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int c = lane + 32 * i;
            if (c < DP) {
                const float o = (c < D) ? (v[i] - mean) * rstd * gamma[c] + beta[c] : 0.f;
                const bf16 ob = __float2bfloat16_rn(o);
                ys[(size_t)r * sy + c] = ob;
                if (yg != nullptr && r0 + r < rows) yg[(size_t)(r0 + r) * ldy + c] = ob;
            }
        }
    }
}

__device__ __forceinline__ void l2_prefetch(const void* ptr, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(CH_NT, 1) decode_chain_kernel(const ChainParams p) {
    ick_pdl_launch();
    // The weights do not depend on the previous kernel: pull them into L2 while it drains (programmatic dependent launch lets this
    // CTA start early).  Between two steps the cross-attention streams > 1 GB through the 126 MB L2, so without this every
    // dependent round of fragment loads below pays a DRAM latency instead of an L2 hit.
    if (blockIdx.x == 0 && threadIdx.x < 4) {
        if (threadIdx.x == 0) l2_prefetch(p.Wo, (unsigned)(p.DP * p.ldwo * 2));
        if (threadIdx.x == 1 && p.W1) l2_prefetch(p.W1, (unsigned)(p.FFP * p.ldw1 * 2));
        if (threadIdx.x == 2 && p.W2) l2_prefetch(p.W2, (unsigned)(p.DP * p.ldw2 * 2));
        if (threadIdx.x == 3 && p.Wn) l2_prefetch(p.Wn, (unsigned)(p.Nn * p.ldwn * 2));
    }
    ick_pdl_wait();
    extern __shared__ __align__(16) uint8_t smem[];
    const int sD = p.DP + CH_PAD, sF = p.FFP + CH_PAD;
    bf16* bufA = reinterpret_cast<bf16*>(smem);          // attention output, later the feed-forward hidden rows (width max(DP, FFP))
    bf16* bufX = bufA + (size_t)CH_ROWS * (sF > sD ? sF : sD);  // residual rows
    bf16* bufY = bufX + (size_t)CH_ROWS * sD;                   // LayerNorm output rows
    float* bufS = reinterpret_cast<float*>(bufY + (size_t)CH_ROWS * sD);  // sublayer output, fp32, stride DP + 8
    const int sS = p.DP + 8;
    const int r0 = blockIdx.x * CH_ROWS;
    const int chunks = p.DP / 8;
    for (int idx = threadIdx.x; idx < CH_ROWS * chunks; idx += CH_NT) {
        const int r = idx / chunks, c = (idx % chunks) * 8;
        uint4 a = make_uint4(0u, 0u, 0u, 0u), x = a;
        if (r0 + r < p.rows) {
            a = *reinterpret_cast<const uint4*>(p.a_in + (size_t)(r0 + r) * p.lda + c);
            x = *reinterpret_cast<const uint4*>(p.x_res + (size_t)(r0 + r) * p.ldx + c);
        }
        *reinterpret_cast<uint4*>(bufA + (size_t)r * sD + c) = a;
        *reinterpret_cast<uint4*>(bufX + (size_t)r * sD + c) = x;
    }
    __syncthreads();
    // attention output projection + residual + LayerNorm
    chain_gemm(bufA, sD, p.DP, p.Wo, p.ldwo, p.bo, p.DP, 0, false, bufS, nullptr, sS, nullptr, 0, r0, p.rows);
    __syncthreads();
    const bool ffn = p.W1 != nullptr;
    chain_ln(bufX, sD, bufS, sS, p.g1, p.be1, p.D, p.DP, p.eps, bufY, sD, ffn ? nullptr : p.y, p.ldy, r0, p.rows);
    __syncthreads();
    if (ffn) {  // feed-forward sublayer: hidden rows in bufA, residual = bufY, result back into bufY (through bufX)
        chain_gemm(bufY, sD, p.DP, p.W1, p.ldw1, p.b1, p.FFP, 1, true, nullptr, bufA, sF, nullptr, 0, r0, p.rows);
        __syncthreads();
        chain_gemm(bufA, sF, p.FFP, p.W2, p.ldw2, p.b2, p.DP, 0, false, bufS, nullptr, sS, nullptr, 0, r0, p.rows);
        __syncthreads();
        chain_ln(bufY, sD, bufS, sS, p.g2, p.be2, p.D, p.DP, p.eps, bufX, sD, p.y, p.ldy, r0, p.rows);
        __syncthreads();
    }
    if (p.Wn != nullptr)
        chain_gemm(ffn ? bufX : bufY, sD, p.DP, p.Wn, p.ldwn, p.bn, p.Nn, 2, false, nullptr, nullptr, 0, p.proj, p.ldproj, r0, p.rows);
}

}  // namespace

extern "C" int ick_decode_chain(const void* attn_out, int lda, const void* x_res, int ldx, const void* Wo, int ldwo, const float* bo,
                                const float* gamma1, const float* beta1, const void* W1, int ldw1, const float* b1, const void* W2,
                                int ldw2, const float* b2, const float* gamma2, const float* beta2, const void* Wn, int ldwn,
                                const float* bn, void* y, int ldy, void* proj, int ldproj, int rows, int D, int DP, int FFP, int Nn,
                                float eps, cudaStream_t stream) {
    ICK_REQUIRE(rows >= 0 && D > 0 && D <= DP && DP % 32 == 0 && DP <= 384, "decode_chain: bad width D=%d DP=%d", D, DP);
    ICK_REQUIRE(W1 == nullptr || (W2 != nullptr && FFP % 32 == 0 && FFP > 0 && FFP <= 1024), "decode_chain: bad feed-forward width %d", FFP);
    ICK_REQUIRE(Wn == nullptr || (Nn % 8 == 0 && Nn > 0 && proj != nullptr), "decode_chain: bad projection width %d", Nn);
    ICK_REQUIRE(lda % 8 == 0 && ldx % 8 == 0 && ldwo % 8 == 0 && (W1 == nullptr || (ldw1 % 8 == 0 && ldw2 % 8 == 0)) &&
                    (Wn == nullptr || (ldwn % 8 == 0 && ldproj % 2 == 0)) && ldy % 2 == 0,
                "decode_chain: leading dimensions must keep 16-byte rows");
    ICK_REQUIRE(((((uintptr_t)attn_out) | ((uintptr_t)x_res) | ((uintptr_t)Wo) | ((uintptr_t)W1) | ((uintptr_t)W2) | ((uintptr_t)Wn)) & 15) == 0 &&
                    ((((uintptr_t)y) | ((uintptr_t)proj)) & 3) == 0,
                "decode_chain: misaligned operand");
    if (rows == 0) return ICK_OK;
    ChainParams p;
    p.a_in = (const bf16*)attn_out; p.lda = lda;
    p.x_res = (const bf16*)x_res; p.ldx = ldx;
    p.Wo = (const bf16*)Wo; p.bo = bo; p.g1 = gamma1; p.be1 = beta1;
    p.W1 = (const bf16*)W1; p.b1 = b1; p.W2 = (const bf16*)W2; p.b2 = b2; p.g2 = gamma2; p.be2 = beta2;
    p.Wn = (const bf16*)Wn; p.bn = bn;
    p.y = (bf16*)y; p.ldy = ldy; p.proj = (bf16*)proj; p.ldproj = ldproj;
    p.rows = rows; p.D = D; p.DP = DP; p.FF = FFP; p.FFP = W1 ? FFP : DP; p.Nn = Nn;
    p.ldwo = ldwo; p.ldw1 = ldw1; p.ldw2 = ldw2; p.ldwn = ldwn;
    p.eps = eps;
    const int sD = DP + CH_PAD, sF = p.FFP + CH_PAD;
    const size_t smem = (size_t)CH_ROWS * ((sF > sD ? sF : sD) + 2 * sD) * sizeof(bf16) + (size_t)CH_ROWS * (DP + 8) * sizeof(float);
    static size_t attr = 0;
    if (smem > attr) {
        cudaFuncSetAttribute(decode_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = smem;
    }
    ick_launch(decode_chain_kernel, (rows + CH_ROWS - 1) / CH_ROWS, CH_NT, smem, stream)(p);
    return ick_check_launch("decode_chain");
}
