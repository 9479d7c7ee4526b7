// Per-step cross-attention of the KV-cached decode loop (DecoderTransformer.predict, G/models.py:389-407: one new query per image
// against the image's memory = [196 pixel tokens; E entity tokens; F fact tokens]) as an HBM-streaming kernel.
//
// The memory K|V of one decoder layer is a contiguous (image, position, [K: H*32 | V: H*32]) bf16 array (engine._memory_kv), i.e.
// 1280 bytes per position for the reference's 10 heads and ~700 KB per image; each step reads all of it exactly once and does
// ~2 flop per byte, so the kernel is bound by HBM and everything here is about keeping enough bytes in flight:
//   * one CTA per image, five CTAs per SM (41 KB of shared memory each): for the 625-image shard of BASELINE configs[3] every CTA is
//     resident at once - no wave quantisation, HBM is shared by all images for the whole launch;
//   * a producer warp streams the image through a 4-stage ring of 8-position (10 KB) slabs with cp.async.bulk (TMA, 1-D) completing
//     on mbarriers, so ~170 KB per SM are in flight independent of what the consumers do - the CUDA-core kernel it replaces
//     (mha_decode_rows_kernel) only issued the loads of the next key pair after the online-softmax chain of the current one and
//     ran at 3.9 TB/s;
//   * 160 consumer threads = 4 key groups x (head, quarter of the head): 40 consecutive threads read one position's 640-byte K block
//     and 640-byte V block from shared memory with conflict-free 16-byte loads; the four quarter threads of a head combine their
//     partial dot products with two warp shuffles; the four key groups keep independent online softmaxes (exp2 domain) that are
//     merged once at the end through shared memory (re-using stage 0).
#include <cstdlib>

#include "attention_internal.h"
#include "common.cuh"
#include "mma.cuh"

namespace {

constexpr int HD = 32;        // padded head width
constexpr int DT_POS = 8;     // positions per stage
constexpr int DT_STAGES = 4;  // ring depth
constexpr int DT_KG = 4;      // key groups (independent online softmaxes)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

// threads: [0, DT_KG*H*4) consumers (a multiple of 32), then one producer warp
__global__ void __launch_bounds__(DT_KG * 40 + 32, 5) mha_decode_tma_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ KV,
                                                                            bf16* __restrict__ O, int H, int dh, int ldq, int ldo,
                                                                            long long batch_stride, int klen, float scale_log2) {
    ick_pdl_entry();
    extern __shared__ __align__(128) uint8_t smem[];
    const int rowbytes = 2 * H * HD * 2;  // K block | V block of one position
    const int stage_bytes = DT_POS * rowbytes;
    const uint32_t bars = smem_u32(smem + DT_STAGES * stage_bytes);  // full[DT_STAGES], empty[DT_STAGES]
    const int slots = H * 4, ncons = DT_KG * slots;
    const int tid = threadIdx.x, b = blockIdx.x;
    const int nchunk = (klen + DT_POS - 1) / DT_POS;
    if (tid == 0) {
        for (int s = 0; s < DT_STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (DT_STAGES + s), ncons / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid >= ncons) {  // ---- producer warp: one elected lane streams the image's K|V rows through the ring
        if (tid == ncons) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(KV + (size_t)b * batch_stride);
            for (int c = 0; c < nchunk; ++c) {
                const int s = c % DT_STAGES, it = c / DT_STAGES;
                if (it > 0) mbar_wait(bars + 8 * (DT_STAGES + s), (it - 1) & 1);
                const int npos = min(DT_POS, klen - c * DT_POS);
                const uint32_t bytes = (uint32_t)(npos * rowbytes);
                mbar_expect_tx(bars + 8 * s, bytes);
                bulk_g2s(smem_u32(smem + s * stage_bytes), src + (size_t)c * stage_bytes, bytes, bars + 8 * s);
            }
        }
        return;
    }
    // ---- consumers
    const int kg = tid / slots, sl = tid % slots, h = sl >> 2, qd = sl & 3;
    float q[8];
    ld8(Q + (size_t)b * ldq + h * HD + qd * 8, q);
#pragma unroll
    for (int c = 0; c < 8; ++c) q[c] = (qd * 8 + c < dh) ? q[c] * scale_log2 : 0.f;  // pad lanes never contribute
    const unsigned qmask = 0xFu << ((tid & 31) & ~3);
    const int eoff = h * HD + qd * 8, voff = H * HD;
    float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    for (int c = 0; c < nchunk; ++c) {
        const int s = c % DT_STAGES;
        mbar_wait(bars + 8 * s, (c / DT_STAGES) & 1);
        const int npos = min(DT_POS, klen - c * DT_POS);
        const bf16* stage = reinterpret_cast<const bf16*>(smem + s * stage_bytes) + eoff;
#pragma unroll
        for (int pp = 0; pp < DT_POS / DT_KG; ++pp) {
            const int p = kg + pp * DT_KG;
            if (p < npos) {  // uniform over the four quarter threads of a (position, head)
                float kx[8], vx[8];
                ld8(stage + (size_t)p * (rowbytes / 2), kx);
                ld8(stage + (size_t)p * (rowbytes / 2) + voff, vx);
                float sc = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) sc = fmaf(q[e], kx[e], sc);
                sc += __shfl_xor_sync(qmask, sc, 1);
                sc += __shfl_xor_sync(qmask, sc, 2);
                const float mnew = fmaxf(m, sc);
                const float corr = exp2f(m - mnew);
                const float pr = exp2f(sc - mnew);
                l = l * corr + pr;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(pr, vx[e], acc[e] * corr);
                m = mnew;
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(bars + 8 * (DT_STAGES + s));  // this warp is done with the stage
    }
    // merge the key groups through shared memory (stage 0 is free: every copy has landed and been consumed)
    asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
    float* red = reinterpret_cast<float*>(smem);  // [DT_KG][slots][10]
    float* mine = red + ((size_t)kg * slots + sl) * 10;
    mine[0] = m;
    mine[1] = l;
#pragma unroll
    for (int c = 0; c < 8; ++c) mine[2 + c] = acc[c];
    asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
    if (kg != 0) return;
    float mall = m;
    for (int g = 1; g < DT_KG; ++g) mall = fmaxf(mall, red[((size_t)g * slots + sl) * 10]);
    float lsum = 0.f, out[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) out[c] = 0.f;
    for (int g = 0; g < DT_KG; ++g) {
        const float* r = red + ((size_t)g * slots + sl) * 10;
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

// ---- beam-search step: the G <= 8 beams of an image against the image's memory K|V --------------------------------------------------
// Same streaming structure (one CTA per image, whole 1280-byte K|V rows copied by TMA through a ring), but the G query rows of an image
// go through the tensor cores: warp h owns head h, S = Q_h K_h^T as mma.sync m16n8k16 with the beams as the (padded) 16 rows, online
// softmax on the accumulator fragments, O += P V_h with P re-used from the accumulator registers.  The memory is read ONCE per image
// and step for all beams.  Differences to the scalar kernel that make the fragment loads work:
//   * a stage holds 16 positions, each copied by its own cp.async.bulk (one per producer lane) into a row padded to 1296 bytes: the
//     K_h / V_h sub-rows of consecutive positions then fall into different bank groups and every ldmatrix is conflict-free (with the
//     natural 1280-byte stride all eight rows of an 8x8 matrix would hit the same banks);
//   * K_h is read as the [n][k] B operand (ldmatrix), V_h as the [k][n] B operand (ldmatrix.trans) straight from the staged rows.
// The flash-attention kernel used before (fwd_pkernel, one warp per (image, head) item over per-head TMA boxes of 64 rows x 64 bytes)
// ran this at 4 TB/s (112 us per layer for 625 images x 5 beams x 548 slots); whole-row copies are what reaches HBM speed.
constexpr int DM_POS = 16;                    // positions per stage = one k-step of P V / two n-tiles of S
constexpr int DM_STAGES = 3;               // 3 x 20.7 KB per CTA: three CTAs per SM (444 slots for the 625 images of a shard)
constexpr int DM_PAD = 16;                    // bytes of padding per staged row
__global__ void __launch_bounds__(352, 3) mha_decode_tma_mma_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ KV, bf16* __restrict__ O,
                                                                    int G, int H, int dh, int ldq, int ldo, long long img_stride, int klen,
                                                                    float scale_log2) {
    ick_pdl_entry();
    extern __shared__ __align__(128) uint8_t smem[];
    const int rowbytes = 2 * H * HD * 2;           // K block | V block of one position in global memory
    const int srow = rowbytes + DM_PAD;            // staged row
    const int stage_bytes = DM_POS * srow;
    const uint32_t bars = smem_u32(smem + DM_STAGES * stage_bytes);  // full[DM_STAGES], empty[DM_STAGES]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, b = blockIdx.x;
    const int nchunk = (klen + DM_POS - 1) / DM_POS;
    if (tid == 0) {
        for (int s = 0; s < DM_STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (DM_STAGES + s), H);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == H) {  // ---- producer warp: lane p copies position p of every stage
        const uint8_t* src = reinterpret_cast<const uint8_t*>(KV + (size_t)b * img_stride);
        for (int c = 0; c < nchunk; ++c) {
            const int s = c % DM_STAGES, it = c / DM_STAGES;
            const int npos = min(DM_POS, klen - c * DM_POS);
            if (lane == 0) {
                if (it > 0) mbar_wait(bars + 8 * (DM_STAGES + s), (it - 1) & 1);
                mbar_expect_tx(bars + 8 * s, (uint32_t)(npos * rowbytes));
            }
            __syncwarp();
            if (lane < npos)
                bulk_g2s(smem_u32(smem + s * stage_bytes + lane * srow), src + ((size_t)c * DM_POS + lane) * rowbytes, (uint32_t)rowbytes, bars + 8 * s);
        }
        return;
    }
    if (warp > H) return;
    // ---- consumers: warp h = head h
    const int h = warp, g = lane >> 2, tq = lane & 3;
    uint32_t qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        qa[ks][1] = qa[ks][3] = 0u;  // rows 8..15 of the 16-row tile do not exist (G <= 8)
        qa[ks][0] = qa[ks][2] = 0u;
        if (g < G) {
            const bf16* qp = Q + ((size_t)b * G + g) * ldq + h * HD + ks * 16 + 2 * tq;
            qa[ks][0] = *reinterpret_cast<const uint32_t*>(qp);
            qa[ks][2] = *reinterpret_cast<const uint32_t*>(qp + 8);
        }
    }
    const int sld = srow / 2;  // staged row stride in elements
    float m = -INFINITY, l = 0.f, o[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    for (int c = 0; c < nchunk; ++c) {
        const int s = c % DM_STAGES;
        mbar_wait(bars + 8 * s, (c / DM_STAGES) & 1);
        const bf16* kt = reinterpret_cast<const bf16*>(smem + s * stage_bytes) + h * HD;  // [position][32] at stride sld
        const bf16* vt = kt + H * HD;
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};  // keys 0-7 / 8-15 of the stage
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t r[4];
            // [n][k] tile: matrices (keys 0-7, k lo), (keys 0-7, k hi), (keys 8-15, k lo), (keys 8-15, k hi)
            ick_ldsm_x4(r[0], r[1], r[2], r[3], ick_smem_u32(kt + (size_t)((lane & 7) + 8 * (lane >> 4)) * sld + ks * 16 + 8 * ((lane >> 3) & 1)));
            ick_mma16816(s0, qa[ks], r[0], r[1]);
            ick_mma16816(s1, qa[ks], r[2], r[3]);
        }
        // this lane holds row g, keys 2tq, 2tq+1 (s0) and 8+2tq, 8+2tq+1 (s1); rows 8-15 (elements 2, 3) are padding
        const int j0 = c * DM_POS + 2 * tq;
        if (j0 >= klen) s0[0] = -INFINITY;
        if (j0 + 1 >= klen) s0[1] = -INFINITY;
        if (j0 + 8 >= klen) s1[0] = -INFINITY;
        if (j0 + 9 >= klen) s1[1] = -INFINITY;
        float mx = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float mn = fmaxf(m, mx);  // finite: every stage holds at least one real position
        const float e = mn * scale_log2;
        const float corr = exp2f(m * scale_log2 - e);
        const float p00 = exp2f(fmaf(s0[0], scale_log2, -e)), p01 = exp2f(fmaf(s0[1], scale_log2, -e));
        const float p10 = exp2f(fmaf(s1[0], scale_log2, -e)), p11 = exp2f(fmaf(s1[1], scale_log2, -e));
        l = l * corr + (p00 + p01) + (p10 + p11);
        m = mn;
#pragma unroll
        for (int n = 0; n < 4; ++n) { o[n][0] *= corr; o[n][1] *= corr; }
        uint32_t pa[4];
        {
            __nv_bfloat162 t0 = __floats2bfloat162_rn(p00, p01), t1 = __floats2bfloat162_rn(p10, p11);
            pa[0] = *reinterpret_cast<uint32_t*>(&t0);
            pa[2] = *reinterpret_cast<uint32_t*>(&t1);
            pa[1] = pa[3] = 0u;
        }
#pragma unroll
        for (int nn = 0; nn < 2; ++nn) {
            uint32_t r[4];
            // [k][n] tile read transposed: (keys 0-7, dims n0..), (keys 8-15, dims n0..), (keys 0-7, dims n0+8..), (keys 8-15, dims n0+8..)
            ick_ldsm_x4_trans(r[0], r[1], r[2], r[3], ick_smem_u32(vt + (size_t)((lane & 7) + 8 * ((lane >> 3) & 1)) * sld + 16 * nn + 8 * (lane >> 4)));
            ick_mma16816(o[2 * nn], pa, r[0], r[1]);
            ick_mma16816(o[2 * nn + 1], pa, r[2], r[3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * (DM_STAGES + s));
    }
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    if (g < G) {
        const float inv = 1.f / l;
        bf16* op = O + ((size_t)b * G + g) * ldo + h * HD + 2 * tq;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int d0 = 8 * n + 2 * tq;
            *reinterpret_cast<__nv_bfloat162*>(op + 8 * n) = __floats2bfloat162_rn(d0 < dh ? o[n][0] * inv : 0.f, d0 + 1 < dh ? o[n][1] * inv : 0.f);
        }
    }
}

bool beam_mma_enabled() {  // ICK_DECODE_TMA_MMA=0: beam cross-attention stays on the flash-attention kernel (A/B aid)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ICK_DECODE_TMA_MMA");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

bool enabled() {  // ICK_DECODE_TMA=0: keep the CUDA-core mha_decode_rows kernel (A/B aid)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ICK_DECODE_TMA");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

}  // namespace

// K|V rows must be contiguous (ldkv == 2*H*32) - the per-layer memory buffers of the decode loops are; anything else is
// ICK_ERR_UNSUPPORTED and the caller keeps its CUDA-core kernel.
int ick_mha_decode_tma(const void* Q, const void* KV, void* O, int B, int H, int dh, int ldq, int ldkv, int ldo, long long batch_stride,
                       int klen, cudaStream_t stream) {
    const int ncons = DT_KG * H * 4;
    if (!enabled() || ldkv != 2 * H * HD || ncons % 32 != 0 || ncons > DT_KG * 40 || klen < 1 || (batch_stride % 8) != 0 ||
        ((((uintptr_t)Q) | ((uintptr_t)KV) | ((uintptr_t)O)) & 15) != 0)
        return ICK_ERR_UNSUPPORTED;
    const int rowbytes = 2 * H * HD * 2;
    const size_t smem = (size_t)DT_STAGES * DT_POS * rowbytes + 2 * DT_STAGES * 8;
    if ((size_t)DT_KG * H * 4 * 10 * sizeof(float) > (size_t)DT_POS * rowbytes) return ICK_ERR_UNSUPPORTED;  // merge scratch aliases stage 0
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(mha_decode_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
        attr_set = true;
    }
    const float sl2 = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
    ick_launch(mha_decode_tma_kernel, B, ncons + 32, smem, stream)((const bf16*)Q, (const bf16*)KV, (bf16*)O, H, dh, ldq, ldo, batch_stride, klen,
                                                                 sl2);
    return ick_check_launch("mha_decode(tma)");
}

// The G beams of every image against the image's memory K|V (contiguous rows, as above).  More than DM_STAGES * DM_POS cached
// positions are required (rows of the last, partial stage then hold finite stale data, which the zero probabilities cancel).
int ick_mha_decode_tma_mma(const void* Q, const void* KV, void* O, int images, int G, int H, int dh, int ldq, int ldkv, int ldo,
                           long long img_stride, int klen, cudaStream_t stream) {
    if (!beam_mma_enabled() || ldkv != 2 * H * HD || H < 1 || H > 10 || G < 1 || G > 8 || klen <= DM_STAGES * DM_POS || (img_stride % 8) != 0 ||
        (ldq % 2) != 0 || (ldo % 2) != 0 || ((((uintptr_t)KV)) & 15) != 0 || ((((uintptr_t)Q) | ((uintptr_t)O)) & 3) != 0)
        return ICK_ERR_UNSUPPORTED;
    const int rowbytes = 2 * H * HD * 2;
    const size_t smem = (size_t)DM_STAGES * DM_POS * (rowbytes + DM_PAD) + 2 * DM_STAGES * 8;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(mha_decode_tma_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) {
            ick_set_error("mha_decode_tma_mma: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed");
            return ICK_ERR_CUDA;
        }
        attr_set = true;
    }
    if (smem > 100 * 1024) return ICK_ERR_UNSUPPORTED;
    const float sl2 = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
    ick_launch(mha_decode_tma_mma_kernel, images, 352, smem, stream)((const bf16*)Q, (const bf16*)KV, (bf16*)O, G, H, dh, ldq, ldo, img_stride, klen,
                                                                    sl2);
    return ick_check_launch("mha_decode_beam(tma, mma)");
}
