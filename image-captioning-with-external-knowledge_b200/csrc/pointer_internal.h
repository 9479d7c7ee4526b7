// Internal (non-ABI) entry points of the tensor-core pointer-head kernels (pointer_mma.cu); context.cu dispatches to them for
// bf16 and keeps the CUDA-core kernels for fp32 parity and single-step decode.  Return ICK_ERR_UNSUPPORTED when the shape
// does not fit (the caller then uses the CUDA-core path).
#pragma once
#include "common.cuh"

int ick_pointer_fwd_mma(const void* h, const void* ctx, const float* w, const float* bias, const int* first_t, float* scores, int B, int Tn,
                        int t0, int S, int D, int ld, int ldscores, int col0, int lag, int beams, cudaStream_t stream);
int ick_pointer_bwd_mma(const void* dS, const void* h, const void* ctx, const float* w, const int* first_t, float* dCtx, void* dH, float* gflat,
                        int w_off, int bias_off, int B, int T, int S, int D, int ld, int ldds, int col0, int lag, cudaStream_t stream);
