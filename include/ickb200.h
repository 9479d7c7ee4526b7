/*
 * ickb200 — C ABI of the B200-native caption-decoder kernels (libickb200.so).
 *
 * The reference (sonniki/image-captioning-with-external-knowledge) has no FFI layer: its hot path is the Python
 * nn.Module API `models.DecoderTransformer.forward / .predict` (geo-aware/models.py:315,363;
 * knowledge-aware/models.py:457,516; news-knowledge-aware/models.py:440,499), whose arithmetic is stock PyTorch.
 * This header is the boundary that replaces those PyTorch library calls: every entry point takes raw DEVICE
 * pointers, explicit sizes / leading dimensions and a cudaStream_t, writes only caller-allocated outputs, keeps no
 * global state, and returns 0 on success (non-zero: see ick_last_error()).  The Python host side
 * (image-captioning-with-external-knowledge_b200/kernels.py) parses THIS FILE to build its ctypes signatures.
 *
 * Conventions
 *   dt            activation dtype: 0 = fp32, 1 = bf16 (fp32 accumulation everywhere)
 *   rows          activations are row-major (batch, position) rows of `ld` elements; logical width D = 300, pad
 *                 columns [D, ld) hold zeros ("dense layout"); projections to/from attention use the "head layout":
 *                 head h in columns [32h, 32h+30), two zero pad lanes per head (10 x 32 = 320 columns)
 *   gflat         flat fp32 gradient buffer in the reference's parameter order (state_dict order); *_off arguments
 *                 are element offsets into it; weight/bias gradients ACCUMULATE (+=) like autograd
 *   dropout       counter-hash masks: keep iff hash(seed, site, element index) >= p * 2^32, scaled by 1/(1-p);
 *                 p = 0 disables.  Backward kernels regenerate the mask from (seed, site).
 */
#ifndef ICKB200_H
#define ICKB200_H

#include <cuda_runtime_api.h>

#define ICK_ABI_VERSION 10

#ifdef __cplusplus
extern "C" {
#endif

const char* ick_last_error(void);
int ick_abi_version(void);
/* number of kernels this library has launched so far in the process (every launch, whichever entry point issued it; replays of a
 * captured CUDA graph are not launches of the library and are not counted) */
int ick_launch_count(long long* out);

/* ---- dense projections: nn.Linear inside nn.MultiheadAttention / Transformer*Layer, fc_vocab --------------------- */
/* C[M,N] (+)= A[M,K] * W[N,K]^T + bias.  epi: 0 none, 1 ReLU+dropout (linear1 -> activation -> dropout),
 * 2 ReLU/dropout backward (C = aux != 0 ? acc / (1-p) : 0).  Replaces F.linear at G/models.py:303 (fc_vocab) and inside
 * torch's TransformerEncoderLayer/DecoderLayer (constructed G/models.py:241-244).  CUDA-core path (fp32 parity mode). */
int ick_gemm_tn_simt(const void* A, int a_dt, const void* W, int w_dt, void* C, int c_dt, const float* bias, const void* aux,
                     int M, int N, int K, int lda, int ldw, int ldc, int ldaux, int epi, int accumulate, float drop_p,
                     unsigned seed, unsigned site, cudaStream_t stream);
/* gflat[rowoff[n] + colmap[k]] += sum_m dY[m,n] X[m,k]; gflat[biasoff[n]] += sum_m dY[m,n].  rowoff/colmap/biasoff
 * translate packed (padded, head-layout) indices to the reference's parameter layout; -1 skips.  autograd of F.linear. */
int ick_wgrad_simt(const void* dY, int y_dt, const void* X, int x_dt, float* gflat, const int* rowoff, const int* colmap,
                   const int* biasoff, int M, int N, int K, int ldy, int ldx, cudaStream_t stream);

/* Same contracts on the 5th-generation tensor cores (tcgen05.mma, TMEM accumulators, TMA operand staging); bf16 operands.
 * ick_gemm_tn_tc additionally needs K % 64 == 0 or TMA zero-fill (handled), lda/ldw multiples of 8. */
int ick_gemm_tn_tc(const void* A, const void* W, void* C, int c_dt, const float* bias, const void* aux, int M, int N, int K, int lda,
                   int ldw, int ldc, int ldaux, int epi, int accumulate, float drop_p, unsigned seed, unsigned site,
                   cudaStream_t stream);
/* Two row groups with their own weights in one launch: rows [0, M_split) use (W, bias, site), rows [M_split, M) use
 * (W2, bias2, site2, dropout row index relative to M_split); M_split must be a multiple of 128 (pad the first group).  Used
 * to run the entity and the fact Transformer encoder stacks (same layer shapes, different parameters: G/models.py:243-244,
 * K/models.py:321-324) in lockstep, so the small fact-sized GEMMs ride along with the entity-sized ones.  W and W2 share
 * ldw. */
int ick_gemm_tn_tc_dual(const void* A, const void* W, const void* W2, void* C, int c_dt, const float* bias, const float* bias2,
                        const void* aux, int M, int M_split, int N, int K, int lda, int ldw, int ldc, int ldaux, int epi,
                        int accumulate, float drop_p, unsigned seed, unsigned site, unsigned site2, cudaStream_t stream);
/* The out-projection input gradient of an attention block, dO = dB W (W = out_proj.weight^T packed K-major), which ALSO writes
 * the attention backward's row term dsum[(b*H + h)*S + i] = sum_d dO[b*S+i, 32h+d] * O[b*S+i, 32h+d] from the epilogue (dO rounded
 * to bf16 first, as stored) - pass dsum_ready = 1 to ick_mha_bwd afterwards.  W2 != NULL: two row groups as in ick_gemm_tn_tc_dual
 * (rows [0, rows0) -> dsum with S tokens per sequence, rows [M_split, M) -> dsum2 with S2); W2 == NULL: rows0 = M.  N = H*32 = ldo
 * columns of head layout.  Replaces the out_proj half of F.multi_head_attention_forward's autograd (G/models.py:241-244). */
int ick_gemm_tn_tc_rowdot(const void* A, const void* W, const void* W2, void* C, const void* O, float* dsum, float* dsum2, int M,
                          int M_split, int rows0, int N, int K, int lda, int ldw, int ldc, int ldo, int S, int S2, int H,
                          cudaStream_t stream);
/* Linear + residual + dropout + LayerNorm, the post-LN sublayer tail x = norm(x + dropout(sublayer(x))) of
 * nn.TransformerEncoderLayer / nn.TransformerDecoderLayer (constructed G/models.py:241-244) with the sublayer's last Linear
 * (out_proj / linear2) folded in:  S = X + dropout(A W^T + bias);  Y = LN(S) * gamma + beta;  mean / rstd per row are kept for
 * ick_add_ln_bwd.  W: [320, K] bf16 (d_model padded to 320 rows, pad rows zero), bias: 320 floats (pad zero, 16-byte aligned),
 * d = 300 logical columns, S / Y rows of lds / ldy >= 320 elements (pad columns written as zeros).  W2 != NULL: rows
 * [M_split, M) use (W2, bias2, gamma2, beta2, site2; dropout rows counted from M_split), M_split % 128 == 0.  One tcgen05 launch
 * instead of ick_gemm_tn_tc + ick_add_ln_fwd: the sublayer output never leaves tensor memory before the normalisation. */
int ick_gemm_add_ln_tc(const void* A, const void* W, const void* W2, const float* bias, const float* bias2, const void* X, void* S,
                       void* Y, float* mean, float* rstd, const float* gamma, const float* beta, const float* gamma2,
                       const float* beta2, int M, int M_split, int K, int d, int lda, int ldw, int ldx, int lds, int ldy, float eps,
                       float drop_p, unsigned seed, unsigned site, unsigned site2, cudaStream_t stream);
/* workspace (optional, fp32, >= splits*N*ceil32(K)*4 bytes): row-slice partial tiles are stored there and reduced by a
 * second kernel; without it the partials are added with fp32 atomics. */
int ick_wgrad_tc(const void* dY, const void* X, float* gflat, const int* rowoff, const int* colmap, const int* biasoff, int M, int N,
                 int K, int ldy, int ldx, void* workspace, long long workspace_bytes, cudaStream_t stream);
/* The weight (+ bias) gradients of up to 8 linear layers in ONE persistent tensor-core launch plus one reduce launch: the
 * work items of all problems are balanced over the SMs in a single wave (the backward of one nn.TransformerEncoderLayer /
 * nn.TransformerDecoderLayer, G/models.py:241-244, holds 4 / 6 such small problems).  Every argument array is a HOST array
 * of `nprob` entries (device pointers / sizes per problem, same meaning as ick_wgrad_tc); colmap[i] / biasoff[i] may be
 * NULL.  workspace: fp32 scratch for the row-split partial tiles (sum_i splits_i * N_i * (ceil32(K_i) + 1) * 4 bytes). */
int ick_wgrad_group_tc(int nprob, const void* const* dY, const void* const* X, const int* const* rowoff, const int* const* colmap,
                       const int* const* biasoff, const int* M, const int* N, const int* K, const int* ldy, const int* ldx,
                       float* gflat, void* workspace, long long workspace_bytes, cudaStream_t stream);

/* ---- attention: F.multi_head_attention_forward inside the Transformer layers ------------------------------------------ */
/* O = dropout(softmax(Q K^T / sqrt(dh) [+ causal mask])) V per (batch, head); lse[b,h,i] (log2 domain) is saved for backward.
 * Causal mask = _generate_square_subsequent_mask, G/models.py:256-262. */
int ick_mha_fwd(const void* Q, const void* K, const void* V, void* O, float* lse, int dt, int B, int H, int Sq, int Sk, int dh,
                int ldq, int ldk, int ldv, int ldo, int causal, float drop_p, unsigned seed, unsigned site, cudaStream_t stream);
int ick_mha_bwd(const void* Q, const void* K, const void* V, const void* O, const void* dO, const float* lse, float* dsum, void* dQ,
                void* dK, void* dV, int dt, int B, int H, int Sq, int Sk, int dh, int ldq, int ldk, int ldv, int ldo, int lddo,
                int lddq, int lddk, int lddv, int causal, float drop_p, unsigned seed, unsigned site, void* workspace,
                long long workspace_bytes, int dsum_ready, cudaStream_t stream);
/* dsum_ready != 0: dsum already holds rowsum(dO * O) (written by ick_gemm_tn_tc_rowdot); implementations that would launch a
 * row-dot kernel skip it, the others recompute it. */
/* workspace (optional, bf16 path): B*H*ceil(Sk/16)*2*ceil(Sq/64) KiB.  With it the backward runs as rowsum(dO*O) -> dK/dV (which
 * also stores dS^T) -> dQ = dS K as a GEMM over the stored dS; without it dS is recomputed by a separate dQ kernel. */
/* one query per (batch, head) against klen cached keys/values (KV-cached greedy decode; predict() re-decodes instead) */
int ick_mha_decode(const void* Q, const void* K, const void* V, void* O, int dt, int B, int H, int dh, int ldq, int ldk, int ldv,
                   int ldo, long long kbatch_stride, long long vbatch_stride, int klen, cudaStream_t stream);

/* ---- residual + dropout + LayerNorm (post-LN sublayer tails of the Transformer layers) --------------------------------- */
/* sub <- s = x + dropout(sub); y[map(r)] = LN(s).  map: out_row = (r / map_s_in) * map_s_out + map_off + r % map_s_in
 * (map_s_in = 0: identity) lets the last encoder layer write into the memory buffer = torch.cat at K/models.py:497-499. */
int ick_add_ln_fwd(const void* x, void* sub, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int dt,
                   int rows, int d, int ldx, int lds, int ldy, float eps, int map_s_in, int map_s_out, int map_off, float drop_p,
                   unsigned seed, unsigned site, cudaStream_t stream);
int ick_add_ln_bwd(const void* dy, const void* s, const float* mean, const float* rstd, const float* gamma, void* dres, void* dsub,
                   float* dgamma, float* dbeta, int dt, int rows, int d, int lddy, int lds, int ldres, int ldsub, int map_s_in,
                   int map_s_out, int map_off, int acc_res, float drop_p, unsigned seed, unsigned site, cudaStream_t stream);
/* Two row groups of the same buffers - rows [0, rows0) and [row1_start, row1_start + rows1) - with their own LayerNorm
 * parameters, dropout sites and row maps in one launch (the lockstep entity / fact encoder stacks).  Rows and dropout row
 * indices are group-local; a zero row map means "same row as the input".  d = 300 in rows of 320 elements only. */
int ick_add_ln_fwd_dual(const void* x, void* sub, void* y, float* mean, float* rstd, int dt, int d, int ld, float eps, int rows0,
                        int rows1, int row1_start, const float* gamma0, const float* beta0, const float* gamma1, const float* beta1,
                        int map0_s_in, int map0_s_out, int map0_off, int map1_s_in, int map1_s_out, int map1_off, float drop_p,
                        unsigned seed, unsigned site0, unsigned site1, cudaStream_t stream);
int ick_add_ln_bwd_dual(const void* dy, const void* s, const float* mean, const float* rstd, void* dres, void* dsub, int dt, int d,
                        int ld, int rows0, int rows1, int row1_start, const float* gamma0, const float* gamma1, float* dgamma0,
                        float* dbeta0, float* dgamma1, float* dbeta1, int map0_s_in, int map0_s_out, int map0_off, int map1_s_in,
                        int map1_s_out, int map1_off, int acc_res, float drop_p, unsigned seed, unsigned site0, unsigned site1,
                        cudaStream_t stream);

/* ---- context preparation ------------------------------------------------------------------------------------------------- */
/* EntityEncoder.forward: variant 0 = geo (G/models.py:82-104), 1 = knowledge (K/models.py:82-133), 2 = news (N/models.py:79-134) */
int ick_entity_encode_fwd(const float* entities, const long long* facts, const float* type_emb, const void* word_emb, void* out,
                          int dt, int variant, int B, int E, int C, int F, int D, int ld, int ldw, int ntypes, int V,
                          cudaStream_t stream);
int ick_entity_encode_bwd(const float* dEnt, const float* entities, const long long* facts, const float* type_emb,
                          const void* word_emb, float* gflat, int type_off, int word_off, int dt, int variant, int B, int E, int C,
                          int F, int D, int ld, int ldw, int ntypes, int V, cudaStream_t stream);
/* FactEncoder.forward, K/models.py:170-188 */
int ick_fact_encode_fwd(const long long* facts, const void* ent_enc, const float* pred_emb, void* out, int dt, int B, int E, int F,
                        int D, int ld, int NP, cudaStream_t stream);
int ick_fact_encode_bwd(const float* dFact, const long long* facts, float* dEnt, float* gflat, int pred_off, int B, int E, int F,
                        int D, int ld, int NP, cudaStream_t stream);
/* CaptionEmbedder.forward (K/models.py:209-259) fused with *sqrt(d) and PositionEncoder (K/models.py:505-507); positions
 * [t0, t0+Tn) of captions with row stride Tstride.
 * `group` (here and in the three entry points below that take it): that many consecutive caption rows share ONE image context
 * (entity / fact encodings, fact list) - the beams of an image during beam-search decoding; 1 = every row has its own. */
int ick_caption_embed_fwd(const long long* captions, const long long* masks, const void* word_emb, const void* ent_enc,
                          const void* fact_enc, const float* pe, void* out, int dt, int B, int Tstride, int t0, int Tn, int V, int E,
                          int F, int D, int ld, int ldw, int pad, float scale, int group, float drop_p, unsigned seed,
                          unsigned site, cudaStream_t stream);
int ick_caption_embed_bwd(const void* dX, const long long* captions, const long long* masks, float* dEnt, float* dFact, float* gflat,
                          int word_off, int dt, int B, int T, int V, int E, int F, int D, int ld, int pad, float scale, float drop_p,
                          unsigned seed, unsigned site, cudaStream_t stream);
/* encoder_out (B, D, P) fp32 channel-major -> rows [b*M, b*M+P) of the memory buffer; encoder_out.permute(2,0,1), G/models.py:347 */
int ick_pixels_fwd(const float* encoder_out, void* memory, int dt, int B, int D, int P, int M, int ld, cudaStream_t stream);
int ick_pixels_bwd(const void* dmemory, float* d_encoder_out, int dt, int B, int D, int P, int M, int ld, cudaStream_t stream);
/* Encoder hand-off: AdaptiveAvgPool2d((Hout, Wout)) of the ResNet trunk output x (B, C, Hin, Win) fp32, written
 * as rows[(b*Hout + oy)*Wout + ox][c] - the K-major A operand of the 1x1 convolution conv1 (G/models.py:32,43-45), which then
 * is ick_gemm_tn_* with W = conv1.weight (emb_dim, C); ick_pixels_bwd(M = P) turns the GEMM output into the (B, emb_dim, P)
 * layout Encoder.forward returns (G/models.py:46). */
int ick_pool_rows_fwd(const float* x, void* rows, int dt, int B, int C, int Hin, int Win, int Hout, int Wout, int ldo,
                      cudaStream_t stream);
/* Its backward (the trunk is fine-tuned: Encoder.fine_tune(True), G/models.py:49-60): d rows -> d x (B, C, Hin, Win) fp32, every
 * input pixel summing grad / window size over the output cells whose window covers it (AdaptiveAvgPool2d backward).  The weight
 * gradient of conv1 is ick_wgrad_tc / ick_wgrad_simt over (d rows of conv1's output, pooled rows), its input gradient ick_gemm_tn_*. */
int ick_pool_rows_bwd(const void* drows, float* dx, int dt, int B, int C, int Hin, int Win, int Hout, int Wout, int ldo,
                      cudaStream_t stream);

/* Input pipeline, device half.  raw: the HDF5 storage format of the images - fp16 (N, 3, H, W), values in
 * [0, 255] (G/create_input_files.py:99-101).  out = (float(half(raw / 255)) - mean[c]) / std[c], i.e. CaptionDataset.__getitem__'s
 * `imgs[i] / 255.` (an fp16 numpy division, G/datasets.py:44) followed by train.py's transforms.Normalize (G/train.py:139-141),
 * written as dt (fp32: bit-identical to the reference's tensor) in NCHW or, channels_last != 0, NHWC order.  mean / stdev: HOST
 * arrays of C floats.  C = 3, H*W a multiple of 8. */
int ick_image_prep(const void* raw_f16, void* out, int dt, long long N, int C, int HW, const float* mean, const float* stdev,
                   int channels_last, cudaStream_t stream);

/* ---- context indicators + predicate gate: get_context_indicators K/models.py:380-418, fc_predicate K/models.py:436-437 ---- */
/* NP: number of predicates if every predicate id is known to lie in [0, NP) (the duplicate-predicate search then uses two shared
 * tables instead of comparing all fact pairs), 0 = unknown. */
int ick_fact_first_mention(const long long* captions, const long long* facts, int* first_t, int* tmin, int B, int T, int F, int V,
                           int E, int group, int NP, cudaStream_t stream);
int ick_pred_gate_fwd(const int* tmin, const long long* facts, const float* WpT, const float* bias, const void* h, void* gate,
                      void* hg, int dt, int B, int Tn, int t0, int F, int D, int ld, int ldp, int NP, int lag, int group,
                      cudaStream_t stream);
int ick_gate_mul_bwd(const void* dHG, const void* h, const void* gate, void* dG, void* dH, int dt, long long n, cudaStream_t stream);
int ick_pred_gate_bwd(const void* dG, const int* tmin, const long long* facts, float* gflat, int wp_off, int dt, int B, int T, int F,
                      int D, int ld, int NP, int lag, cudaStream_t stream);

/* ---- pointer heads: fc_entity / fc_fact over h*ctx, get_scores K/models.py:440-452 ----------------------------------------- */
int ick_pointer_fwd(const void* h, const void* ctx, const float* w, const float* bias, const int* first_t, float* scores, int dt,
                    int B, int Tn, int t0, int S, int D, int ld, int ldscores, int col0, int lag, int group, cudaStream_t stream);
int ick_pointer_bwd(const void* dS, const void* h, const void* ctx, const float* w, const int* first_t, float* dCtx, void* dH,
                    float* gflat, int w_off, int bias_off, int dt, int B, int T, int S, int D, int ld, int ldds, int col0, int lag,
                    cudaStream_t stream);

/* ---- loss / optimizer / packing --------------------------------------------------------------------------------------------- */
/* pack_padded_sequence + CrossEntropyLoss(ignore_index=pad), G/train.py:275-281.  loss_acc[0] += sum of row losses,
 * loss_acc[1] += number of kept rows; dscores = softmax - onehot (unscaled, zeros on dropped rows) or NULL. */
int ick_ce_fwd_bwd(const float* scores, const long long* captions_sorted, const int* decode_len, float* loss_acc, void* dscores,
                   int dt, int B, int T, int W, int lds, int ldd, int pad, cudaStream_t stream);
/* grad = g * grad_scale / max(count[0],1), clamped to +-clip (ut.clip_gradient, G/utils.py:75-85), torch.optim.Adam update
 * (G/train.py:85-88, 292); then every parameter element is scattered to its packed copies: dstA/dstB index the dt-typed
 * packT (forward and transposed layouts), dstC the fp32 packF.  update = 0 only repacks. */
int ick_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                  float bias_corr1, float bias_corr2, float clip, const float* count, float grad_scale, const int* dstA,
                  const int* dstB, const int* dstC, void* packT, int dt, float* packF, int update, const int* step_dev,
                  const float* lr_dev, cudaStream_t stream);
/* Dropout kernels launched after this call add *seed_dev (device memory, may be NULL) to their seed argument when they
 * run: a captured CUDA graph then draws fresh masks on every replay.  step_dev / lr_dev of ick_adam_step play the same
 * role for the Adam bias corrections and the learning rate. */
int ick_set_seed_source(const unsigned* seed_dev);
int ick_cast2d(const void* src, int src_dt, void* dst, int dst_dt, long long rows, int cols, int lds, int ldd, cudaStream_t stream);
int ick_accum_f32(const void* src, int dt, float* dst, long long n, cudaStream_t stream);
int ick_colsum(const void* x, int dt, float* out, long long rows, int cols, int ld, cudaStream_t stream);

/* ---- greedy decode step: argmax, top-2, <end> check, repetition clean-up, next input token/mask; G/models.py:409-442 -------- */
int ick_greedy_select(const float* scores, int W, int lds, long long* output, int* second, long long* captions, long long* masks,
                      int* done, float* margins, int B, int step, int Tmax, int V, int E, int has_facts, int end_tok,
                      cudaStream_t stream);


/* ---- decode loops: the row-wise tail of a decoder layer between two attention kernels in ONE launch (decode_chain.cu) ----------
 * y = LayerNorm(x_res + attn_out Wo^T + bo) [gamma1, beta1];  if W1: y = LayerNorm(y + relu(y W1^T + b1) W2^T + b2) [gamma2, beta2];
 * y (rows, ldy) is written;  if Wn: proj (rows, ldproj) = y Wn^T + bn (Nn columns) - the cross-attention query projection or the
 * next layer's self-attention q|k|v cache row (nn.TransformerDecoderLayer, post-LN, G/models.py:241-242; one position per row).
 * bf16 activations and K-major packed weights [N, ldw] (the same operand copies the GEMM entry points take), fp32 biases /
 * LayerNorm parameters, fp32 accumulation.  D real columns in rows of DP (pad columns zero), DP and FFP multiples of 32. */
int ick_decode_chain(const void* attn_out, int lda, const void* x_res, int ldx, const void* Wo, int ldwo, const float* bo,
                     const float* gamma1, const float* beta1, const void* W1, int ldw1, const float* b1, const void* W2, int ldw2,
                     const float* b2, const float* gamma2, const float* beta2, const void* Wn, int ldwn, const float* bn, void* y,
                     int ldy, void* proj, int ldproj, int rows, int D, int DP, int FFP, int Nn, float eps, cudaStream_t stream);

/* ---- beam-search decoding: EXTENSION, no reference counterpart (the reference's predict() is greedy; BASELINE.json
 * asks for beam-5).  The algorithm is the beam search of the Show-Attend-Tell tutorial the reference's READMEs name as their
 * starting point (G/README.md:37), restated in oracle/decoder_oracle.py:beam_search over the reference-pinned scoring function. -- */
/* `rows` query rows, `group` consecutive rows = the beams of one image, one query per (row, head) against klen cached positions.
 * anc == NULL: the beams share the image's keys/values (cross-attention over the memory): row j of image i is at
 *   K + i*kimg_stride + j*ldk, read once per image for all its beams.
 * anc != NULL: position-major self-attention cache; the key of query row r at position j is row
 *   (r/group)*group + anc[r*anc_ld + j] of the block at K + j*kpos_stride, i.e. anc names the beam slot that held row r's
 *   ancestor at step j - re-ordering beams rewrites only that table. */
int ick_mha_decode_beam(const void* Q, const void* K, const void* V, void* O, int dt, int rows, int group, int H, int dh, int ldq,
                        int ldk, int ldv, int ldo, long long kimg_stride, long long vimg_stride, int klen, const int* anc, int anc_ld,
                        long long kpos_stride, long long vpos_stride, cudaStream_t stream);
/* One beam-search step for `images` images of `group` beam slots each (rows img*group + slot of scores / cum / histories).
 * ksel[img] = k, the number of captions still wanted (starts at group; live rows = 1 at step 0, else k).  Candidates
 * cum[row] + log_softmax(scores[row])[c] over the live rows; the k best in descending order: token <end> completes a caption
 * (kept in result/best if it beats the best completed one; k decreases), the others become beams 0..k'-1: histories
 * tok/mask/anc (rows of Tmax) are copied parent -> new slot from *_in to *_out and extended at position step+1 with the token,
 * its mask class (0 word, 1 entity, 2 fact) and the slot itself; cum is updated in place.  result (images, Tmax): tokens
 * without <start>, <end> included, pad-filled.  At the last step an image without a completed caption takes its best live beam.
 * Two launches: per live row the log-sum-exp and its `group` best columns (HBM-bound pass over the scores), then per image the
 * merge and the bookkeeping.  workspace: images*group*group*8 bytes of scratch for the per-row candidates. */
int ick_beam_select(const float* scores, int W, int lds, float* cum, int* ksel, const long long* tok_in, const long long* mask_in,
                    long long* tok_out, long long* mask_out, const int* anc_in, int* anc_out, float* best, long long* result,
                    int images, int group, int step, int Tmax, int V, int E, int has_facts, int end_tok, int pad_tok, void* workspace,
                    long long workspace_bytes, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ICKB200_H */
