// Internal (non-ABI) entry points shared between attention.cu (C-ABI dispatch, CUDA-core kernels) and
// attention_mma.cu (bf16 tensor-core kernels).
#pragma once
#include "common.cuh"

int ick_mha_fwd_mma(const void* Q, const void* K, const void* V, void* O, float* lse, int B, int H, int Sq, int Sk, int dh, int ldq, int ldk,
                    int ldv, int ldo, int causal, DropCfg dc, cudaStream_t stream);
int ick_mha_bwd_mma(const void* Q, const void* K, const void* V, const void* O, const void* dO, const float* lse, float* dsum, void* dQ,
                    void* dK, void* dV, int B, int H, int Sq, int Sk, int dh, int ldq, int ldk, int ldv, int ldo, int lddo, int lddq, int lddk,
                    int lddv, int causal, DropCfg dc, void* workspace, long long workspace_bytes, int dsum_ready, cudaStream_t stream);
// attention_bwd_fused.cu: ONE persistent kernel for dQ, dK and dV (mma.sync for S^T / dP^T / dK / dV, tcgen05 + TMEM for the dQ
// accumulation across key slabs).  ICK_ERR_UNSUPPORTED (nothing launched) when an (image, head) item does not fit its
// shared-memory / tensor-memory budget (more than 640 queries or keys) - the caller then runs the chunked two-kernel path.
int ick_mha_bwd_fused(const void* Q, const void* K, const void* V, const void* O, const void* dO, const float* lse, float* dsum, void* dQ,
                      void* dK, void* dV, int B, int H, int Sq, int Sk, int dh, int ldq, int ldk, int ldv, int ldo, int lddo, int lddq, int lddk,
                      int lddv, int causal, DropCfg dc, cudaStream_t stream);
// attention_fwd_tc.cu: the forward on tcgen05 / TMEM (S = Q K^T and O += P V as UMMA tiles, softmax warps on tcgen05.ld, lazily
// rescaled accumulator).  ICK_ERR_UNSUPPORTED (nothing launched) when an item's Q, K, V do not fit two shared-memory stages.
int ick_mha_fwd_tc(const void* Q, const void* K, const void* V, void* O, float* lse, int B, int H, int Sq, int Sk, int dh, int ldq, int ldk, int ldv,
                   int ldo, int causal, DropCfg dc, cudaStream_t stream);
// attention_bwd_tc.cu: the backward entirely on tcgen05 / TMEM (S^T, dP^T, dV, dK, dQ as UMMA tiles, softmax warps on tcgen05.ld).
// ICK_ERR_UNSUPPORTED (nothing launched) for more than 512 queries or when the operand rings do not fit.
int ick_mha_bwd_tc(const void* Q, const void* K, const void* V, const void* O, const void* dO, const float* lse, float* dsum, void* dQ, void* dK,
                   void* dV, int B, int H, int Sq, int Sk, int dh, int ldq, int ldk, int ldv, int ldo, int lddo, int lddq, int lddk, int lddv,
                   int causal, DropCfg dc, int dsum_ready, cudaStream_t stream);
// attention_decode_tma.cu: TMA-staged per-step cross-attention of the decode loop (contiguous K|V rows only; else ICK_ERR_UNSUPPORTED)
int ick_mha_decode_tma(const void* Q, const void* KV, void* O, int B, int H, int dh, int ldq, int ldkv, int ldo, long long batch_stride,
                       int klen, cudaStream_t stream);
// the G <= 8 beams of an image against its memory K|V on the tensor cores, whole rows streamed by TMA (same contiguity rule)
int ick_mha_decode_tma_mma(const void* Q, const void* KV, void* O, int images, int G, int H, int dh, int ldq, int ldkv, int ldo,
                           long long img_stride, int klen, cudaStream_t stream);
