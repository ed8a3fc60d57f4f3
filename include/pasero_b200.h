/*
 * pasero_b200 — C ABI of the B200-native kernels behind Pasero's speech-translation training step.
 *
 * The reference (naver/pasero) is pure Python and has no FFI of its own; its extension surface is the
 * Python model/task registry (pasero/config.py:91-135). These entry points are what the Python host
 * layer (pasero_b200/*.py, mirroring pasero.models / pasero.tasks / pasero.training) binds with ctypes;
 * each declaration cites the reference code whose work it takes over.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; pb200_last_error_string() describes the
 *     last failure on the calling thread. Nothing throws, nothing aborts.
 *   - the caller owns every buffer, including workspaces (sizes from pb200_*_workspace_bytes); the
 *     library never allocates or frees device memory.
 *   - all work is asynchronous and ordered on `stream` (a cudaStream_t passed as void*); no host
 *     synchronisation, no global mutable state apart from read-only tables => CUDA-graph capturable.
 *   - device = the CUDA context current on the calling thread. Pointers are device pointers unless
 *     stated otherwise, 16-byte aligned, row-major, bf16 for activations/weights, fp32/fp64 statistics.
 *   - there is no CPU fallback: without a GPU every compute entry point fails with a CUDA error.
 */
#ifndef PASERO_B200_H
#define PASERO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_ABI_VERSION 1

int32_t pb200_abi_version(void);
const char* pb200_last_error_string(void);
/* number of SMs of the current device (148 on B200); <= 0 when no device is usable */
int32_t pb200_num_sms(void);

/* ------------------------------------------------------------------------------------------------
 * Audio frontend — fused log-mel fbank (+ utterance CMVN).
 * Oracle: torchaudio.compliance.kaldi.fbank(wave * 2**15, num_mel_bins=80, sample_frequency=16000)
 * (kaldi.py:514-646) followed by per-utterance mean/variance normalisation; the reference itself only
 * consumes pre-extracted features (pasero/files.py:103-192, pasero/tasks/speech_translation.py:42-86).
 *   pcm            fp32 samples of all utterances, concatenated
 *   sample_offsets int64[batch+1] start of each utterance in `pcm` (device)
 *   frame_offsets  int64[batch+1] start row of each utterance in `feats` (device);
 *                  frames(b) = 1 + (samples(b) - 400) / 160  (pb200_fbank_num_frames)
 *   max_frames     host-side upper bound on frames per utterance (grid sizing)
 *   pcm_scale      multiplier applied to samples on load (32768 for [-1,1) floats)
 *   cmvn           bit 0: normalise per utterance; bit 1 (needs feats_bf16_padded): the fp32 rows are only scratch
 *                  (they keep the raw log-mel values), the normalised features are written to the padded layout alone
 *   feats          fp32 [total_frames, 80] out (16-byte aligned); normalised in place when cmvn == 1
 *   feats_bf16_padded  optional bf16 [batch, padded_frames, 80]: features of utterance b at rows
 *                  [front_pad, front_pad + frames(b)), zeros elsewhere (conv subsampler input layout)
 *   workspace      pb200_fbank_workspace_bytes(batch) bytes (fp64 statistics + tile counters). It must be ZERO on first
 *                  use; every successful call leaves it zeroed again, so a caller that keeps the buffer (one per stream)
 *                  never issues a memset. One launch does everything (framing .. CMVN .. padded layout).
 * ---------------------------------------------------------------------------------------------- */
int pb200_fbank_init(void);
int64_t pb200_fbank_num_frames(int64_t num_samples);
size_t pb200_fbank_workspace_bytes(int32_t batch);
int pb200_fbank_cmvn(const float* pcm, const int64_t* sample_offsets, const int64_t* frame_offsets, int32_t batch,
                     int32_t max_frames, float pcm_scale, int32_t cmvn, float* feats, void* feats_bf16_padded,
                     int32_t padded_frames, int32_t front_pad, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * bf16 GEMM on tcgen05/TMEM with TMA-fed operands:  C[z] = epilogue(alpha * op(A[z]) op(B[z])^T)
 * op(A) is [m x k], op(B) is [n x k]. An operand is "K-major" when its reduction dimension is the
 * contiguous one (x_mn_major = 0: stored [rows = m|n][cols = k]) and "MN-major" otherwise
 * (x_mn_major = 1: stored [rows = k][cols = m|n]). ld* = elements between stored rows (may be smaller
 * than the row length: overlapping rows express the strided-conv im2col view without materialising it).
 * Takes over nn.Linear / nn.Conv1d / Embedding.projection of the reference
 * (pasero/models/modules.py:610-615, 737, 829-834, 935-947; transformer.py:999-1019, 1224-1244).
 * ---------------------------------------------------------------------------------------------- */
typedef struct pb200_gemm_params {
    const void* a;
    const void* b;
    void* c;
    int64_t m, n, k;
    int64_t batch;            /* number of z slices (>= 1) */
    int64_t lda, ldb, ldc;
    int64_t a_batch_stride, b_batch_stride, c_batch_stride; /* elements; 0 = operand shared by all z */
    int32_t a_mn_major, b_mn_major;
    int32_t reduce_batch;     /* 1: C = sum_z op(A[z]) op(B[z])^T (weight gradients over utterances) */
    const void* bias;         /* bf16 [n] or NULL */
    const void* residual;     /* bf16, C layout, added after the activation, or NULL */
    const void* relu_mask;    /* bf16, C layout: result zeroed where mask <= 0 (ReLU backward), or NULL */
    void* relu_bits_out;      /* with bias + act == 1: uint32 [n / 32][relu_bits_ld] sign map of the output (bit c % 32 of
                                 word [c / 32][r] = C[r][c] > 0), or NULL. n % 32 == 0, single batch */
    const void* relu_bits;    /* such a map used as the ReLU-backward mask instead of relu_mask (1 bit per element instead of
                                 a bf16 tensor read in the epilogue), or NULL */
    int64_t relu_bits_ld;     /* words between consecutive 32-column groups (>= m) */
    int32_t act;              /* 0 none, 1 ReLU, 2 GELU (erf) */
    int32_t out_fp32;         /* 0: C is bf16, 1: C is fp32 */
    int32_t accumulate;       /* 1: C += result. With out_fp32 and no other epilogue option, every (split-K) partial tile
                                 is TMA-reduce-added straight into C: no workspace, no memset, no reduction kernel */
    float alpha;
    int32_t split_k;          /* 0 = choose automatically, >= 1 = force */
    int32_t tile_n;           /* 0 = choose automatically, 128 or 256 = force */
    int32_t cluster_m;        /* 0 = choose automatically; 1, 2 or 4 CTAs per cluster sharing the B tile by TMA multicast */
    void* workspace;          /* fp32 split-K partials; may be NULL when split_k == 1 */
    size_t workspace_bytes;
    float* a_rowsum;          /* fp32 [m] or NULL: a_rowsum[i] += sum_k op(A)[i][k], computed by the tensor core next to the
                                 product. For weight gradients (op(A) = dY^T) this is the bias gradient, so dY is not read
                                 again by a column-sum kernel. Needs a_mn_major, out_fp32 + accumulate and no other epilogue
                                 option; additions are atomic (order not fixed, like the split-K reduce-adds) */
} pb200_gemm_params;

size_t pb200_gemm_workspace_bytes(int64_t m, int64_t n, int64_t batch, int32_t max_splits);
int pb200_gemm_bf16(const pb200_gemm_params* params, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-head attention, head_dim 64, length-masked (kv_len) and/or causal; replaces
 * MultiheadAttention.forward's SDPA / flash-attn call (pasero/models/modules.py:648-734).
 * q/k/v rows are tokens ([B*L, ld]), head h occupies columns [h*64, h*64+64) from the given pointer.
 * lse: fp32 [B, H, Lq] (written by fwd, read by bwd); delta: unused (kept for ABI stability; may be NULL).
 * Any Lq / Lk: blocks of 256 keys (and 256 queries in backward) are processed on chip by tcgen05 kernels, longer
 * sequences are tiled over such blocks and merged exactly (workspace, see below).
 * ---------------------------------------------------------------------------------------------- */
typedef struct pb200_attn_params {
    const void* q; const void* k; const void* v;
    void* o;
    float* lse;
    const int32_t* kv_len;    /* int32 [B] valid key count per batch element, or NULL */
    int64_t ldq, ldk, ldv, ldo;
    int32_t B, H, Lq, Lk, head_dim;
    int32_t causal;
    float scale;              /* 1/sqrt(head_dim) when `scaled` */
    /* backward only */
    const void* dout; int64_t lddo;
    float* delta;
    void* dq; void* dk; void* dv;
    int64_t lddq, lddk, lddv;
    /* attention dropout (pasero/models/modules.py:686-731; 0 = off). Keep decision of (b, h, query, key) =
     * Philox(seed, offset + (*offset_base << 12)), 16 bits each. drop_mask: uint32 [B, H, Lq, ceil(Lk / 32)], bit j of word w
     * = key 32 w + j kept; WRITTEN by fwd, READ by bwd (required when dropout_p > 0). */
    float dropout_p;
    uint64_t seed, offset;
    const int64_t* offset_base;
    uint32_t* drop_mask;
    /* partial results when Lk > 256 (fwd) or Lq > 256 / Lk > 256 (bwd): pb200_attention_workspace_bytes; may be NULL otherwise */
    void* workspace; size_t workspace_bytes;
} pb200_attn_params;

size_t pb200_attention_workspace_bytes(const pb200_attn_params* params, int32_t backward);
int pb200_attention_fwd(const pb200_attn_params* params, void* stream);
int pb200_attention_bwd(const pb200_attn_params* params, void* stream);

/* ------------------------------------------------------------------------------------------------
 * sum = residual + dropout(x);  y = LayerNorm(sum) * gamma + beta   (gamma == NULL: no LayerNorm)
 * (pasero/models/transformer.py:1043-1054, 1073-1086, 1322-1339, 1389-1409). Dropout mask =
 * Philox(seed, offset + (*offset_base << 12), element / 8); backward regenerates it. offset_base (device int64,
 * may be NULL) lets a captured CUDA graph draw fresh masks on every replay.
 * ---------------------------------------------------------------------------------------------- */
int pb200_dropout_add_ln_fwd(const void* x, const void* residual, const void* gamma, const void* beta, int64_t M,
                             int64_t D, float eps, float p, uint64_t seed, uint64_t offset, const int64_t* offset_base,
                             void* sum_out, void* y_out, float* mean_out, float* rstd_out, void* keep_out, void* stream);
/* keep_out (optional, uint8 [M * D / 8]): the dropout keep bits (bit j of byte e / 8 = element e + j kept); passing it
 * back as keep_mask spares the backward kernel the Philox rounds. */
/* The Linear that produces x fused with the tail above, for d_model = 512 (one kernel: A [M, K] bf16 @ w [512, K]^T + bias ->
 * dropout -> + residual -> `sum_out` -> LayerNorm -> `y_out`; mean / rstd / keep bits as pb200_dropout_add_ln_fwd writes them).
 * Replaces nn.Linear + dropout + residual + nn.LayerNorm of a post-norm sub-block (pasero/models/modules.py:737,
 * transformer.py:1016, 1043-1054, 1073-1086) without x ever reaching HBM. The keep decisions are a counter hash of
 * (seed, offset + (*offset_base << 12), element), NOT the Philox stream of the unfused kernel: keep_out is required when
 * p > 0 and is what backward reads. */
int pb200_linear_dropout_add_ln_fwd(const void* a, const void* w, const void* bias, const void* residual, const void* gamma,
                                    const void* beta, int64_t M, int64_t K, float eps, float p, uint64_t seed, uint64_t offset,
                                    const int64_t* offset_base, void* sum_out, void* y_out, float* mean_out, float* rstd_out,
                                    void* keep_out, void* stream);
size_t pb200_ln_bwd_workspace_bytes(int64_t M, int64_t D);
/* backward: dx / dresidual, dgamma, dbeta and (optional) dbias_lin = column sums of dx, i.e. the bias gradient of the
 * Linear whose output was `x` (saves a separate pass over dx). accumulate: bit 0 = add into dgamma/dbeta, bit 1 = add
 * into dbias_lin, bit 2 = leave the CTA partials in `workspace` and skip the final reduction: the caller runs
 * pb200_ln_bwd_reduce on them later / on another stream (the parameter gradients have no consumer inside backward, so
 * that small kernel need not sit on the critical path). */
int pb200_dropout_add_ln_bwd(const void* dy, const void* dsum_extra, const void* sum, const void* gamma,
                             const float* mean, const float* rstd, int64_t M, int64_t D, float p, uint64_t seed,
                             uint64_t offset, const int64_t* offset_base, void* dresidual, void* dx, void* dgamma, void* dbeta,
                             void* dbias_lin, int32_t accumulate, const void* keep_mask, void* workspace, size_t workspace_bytes,
                             void* stream);
int pb200_ln_bwd_reduce(const void* workspace, int64_t M, int64_t D, void* dgamma, void* dbeta, void* dbias_lin,
                        int32_t accumulate, void* stream);

/* column sums of a bf16 [M, N] matrix -> bf16 [N] (bias gradients) */
size_t pb200_colsum_workspace_bytes(int64_t M, int64_t N);
int pb200_colsum_bf16(const void* x, int64_t M, int64_t N, int64_t ld, void* out, int32_t accumulate, void* workspace,
                      size_t workspace_bytes, void* stream);
/* the same sums into an fp32 [N] vector (fp32 main gradients of biases, see pb200_grad32_flush) */
int pb200_colsum_f32(const void* x, int64_t M, int64_t N, int64_t ld, float* out, int32_t accumulate, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Embedding epilogue: out[r] = dropout(bf16(bf16(src_row * scale) + pos[r % L])), src_row = src[ids[r]]
 * when ids != NULL else src[r]   (pasero/models/transformer.py:739-744, 874-878; modules.py:415-457,916-933)
 * Backward: dx (no ids) or fp32 atomics into dtable[ids[r]] (rows == padding_idx skipped).
 * ---------------------------------------------------------------------------------------------- */
int pb200_embed_fwd(const int64_t* ids, const void* src, const void* pos, int64_t M, int64_t D, int64_t L, float scale,
                    float p, uint64_t seed, uint64_t offset, const int64_t* offset_base, int64_t num_embeddings, void* out,
                    void* stream);
int pb200_embed_bwd(const void* dy, const int64_t* ids, int64_t M, int64_t D, float scale, float p, uint64_t seed,
                    uint64_t offset, const int64_t* offset_base, int64_t padding_idx, int64_t num_embeddings, void* dx,
                    float* dtable, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Conv subsampler pieces (pasero/models/modules.py:774-834): GLU over channels writing the next conv's
 * zero-padded input layout, its backward, the col2im gather of the strided-conv input gradient, and the
 * zero-padding copy of the first conv's input.
 * ---------------------------------------------------------------------------------------------- */
int pb200_glu_fwd(const void* in, int32_t B, int32_t T, int32_t C, void* out, int32_t Tp, int32_t front_pad, void* stream);
int pb200_glu_bwd(const void* dy, const void* in, int64_t M, int32_t C, void* din, void* stream);
/* GELU (erf) for the Whisper-style variant (pasero/config.py:2539-2566: activation_fn / conv_activation = gelu):
 * forward writes gelu(in [B, T, C]) into rows [front_pad, front_pad + T) of a zero-padded [B, Tp, C] layout */
int pb200_gelu_fwd(const void* in, int32_t B, int32_t T, int32_t C, void* out, int32_t Tp, int32_t front_pad, void* stream);
int pb200_gelu_bwd(const void* dy, const void* pre, int64_t n, void* dx, void* stream);
/* backward of a stand-alone ReLU (the encoder's in_linear, pasero/models/transformer.py:647-651): dx = dy where h > 0 */
int pb200_relu_bwd(const void* dy, const void* h, int64_t n, void* dx, void* stream);
int pb200_col2im(const void* dcol, int32_t B, int32_t Tin, int32_t Tout, int32_t C, int32_t KW, int32_t stride,
                 void* dx, void* stream);
int pb200_pad_rows(const void* in, int32_t B, int32_t T, int32_t C, void* out, int32_t Tp, int32_t front_pad, void* stream);
int pb200_f32_to_bf16(const float* in, int64_t n, void* out, int32_t accumulate, void* stream);
/* Device-side collate of a feature batch (Task.collate right-padding + cast, pasero/tasks/task.py:531-571, utils.py:709-736):
 * `rows` = the utterances' feature rows back to back (fp16 when rows_fp16 != 0 — the NumpyFile default, pasero/files.py:122-159
 * — else fp32), offsets int64 [B + 1] in rows; utterance b -> rows [front_pad, front_pad + len_b) of out[b] (bf16 [B, Tp, C]),
 * zeros elsewhere. */
int pb200_collate_features(const void* rows, int32_t rows_fp16, const int64_t* offsets, int32_t B, int32_t C, void* out,
                           int32_t Tp, int32_t front_pad, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Label-smoothed cross-entropy over bf16 logits [B*T, V] (row stride ld), targets int64 read as
 * target[b * target_stride0 + t]; rows whose target == padding_idx contribute nothing. Overwrites the
 * logits with d loss / d logits (bf16) when write_grad != 0. stats (fp64): [0] += loss (nats, summed),
 * [1] += nll (nats, summed), [2] += number of non-pad targets, [7] += number of targets outside [0, V) (treated as
 * padding; the caller raises, as F.cross_entropy does).  stats has >= 8 entries.  (pasero/models/transformer.py:323-380)
 * ---------------------------------------------------------------------------------------------- */
int pb200_ls_ce_fwd_bwd(void* logits, const int64_t* target, int64_t target_stride0, int64_t B, int64_t T, int64_t V,
                        int64_t ld, int64_t padding_idx, float label_smoothing, float grad_scale, int32_t write_grad,
                        double* stats, float* row_loss, void* stream);

/* The same loss for large vocabularies, evaluated over column chunks [chunk_begin, chunk_begin + chunk_cols) of the
 * projection so that no [B*T, V] tensor exists in HBM (the reference holds three fp32 ones, transformer.py:355-372):
 *   pass 1, per chunk: pb200_ce_chunk_stats folds the chunk's logits (bf16 [B*T, chunk_cols], row stride ld) into row_state
 *           (fp32 [4][B*T]: running max, sum exp, sum of logits, target logit; chunk_begin == 0 initialises it);
 *   pb200_ce_chunk_finish: row_state -> stats ([0] loss, [1] nll, [2] tokens, [7] bad targets) and lse fp32 [B*T];
 *   pass 2, per chunk: pb200_ce_chunk_grad overwrites the recomputed chunk with grad_scale * d loss / d logits (the caller
 *           runs the chunk's dgrad / wgrad GEMMs on it; grad_scale: the loss scale of float16 training, 1 otherwise). */
int pb200_ce_chunk_stats(const void* logits, int64_t ld, const int64_t* target, int64_t target_stride0, int64_t B, int64_t T,
                         int64_t chunk_begin, int64_t chunk_cols, int64_t padding_idx, float* row_state, void* stream);
int pb200_ce_chunk_finish(const float* row_state, const int64_t* target, int64_t target_stride0, int64_t B, int64_t T, int64_t V,
                          int64_t padding_idx, float label_smoothing, double* stats, float* lse, void* stream);
int pb200_ce_chunk_grad(void* logits, int64_t ld, const int64_t* target, int64_t target_stride0, int64_t B, int64_t T,
                        int64_t chunk_begin, int64_t chunk_cols, int64_t V, float label_smoothing, float grad_scale,
                        const float* lse, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer over the flat arenas (pasero/training.py:455-487; pasero/optimization.py:56-149, 390-427).
 *   hyper fp32[16] (device): [0] lr [1] beta1 [2] beta2 [3] eps [4] weight_decay [5] 1-beta1^t
 *        [6] sqrt(1-beta2^t) [7] clip_norm [8] dp_size [9] 1-beta1 [10] 1-beta2 [11] lr/(1-beta1^t)
 *        [12] weight_decay*lr
 *   stats fp64[8] (device): [2] num_tokens (gradient divisor), [4] squared gradient norm (out; reduced in a fixed
 *        order so that data-parallel replicas stay bit-identical). The sqnorm workspace is zero-initialised once.
 *
 * fp32 main gradients: weight-gradient GEMMs may reduce-add straight into an fp32 accumulation arena (pb200_gemm_bf16
 * with out_fp32 = accumulate = 1); pb200_grad32_flush rounds the given element ranges ([n][2] begin/end, device,
 * multiples of 8) into the bf16 gradient arena (what the allreduce and the optimizer read) and zeroes the source.
 * ---------------------------------------------------------------------------------------------- */
int pb200_grad32_flush(float* acc32, void* grad_bf16, const int64_t* ranges, int32_t n_ranges, int64_t max_len, void* stream);
/* optimizer kernels */
size_t pb200_grad_sqnorm_workspace_bytes(void);
int pb200_grad_sqnorm(const void* grad, int64_t n, const float* hyper, double* stats, void* workspace,
                      size_t workspace_bytes, void* stream);
/* master: NULL (parameters updated in place with the reference's 16-bit rounding points, optimization.Adam with
 * memory_efficient_fp16) or an fp32 copy of the parameters (MixedPrecisionAdam / FlatFP16Adam): the update is applied to it and
 * `param` receives its rounded image. hyper[13] = 1 / loss scale, hyper[14] != 0 selects the fp32-gradient semantics in the norm.
 * pb200_grad_sqnorm also sets stats[5] = 1 when the norm is not finite; pb200_adam_step then changes nothing. */
int pb200_adam_step(void* param, float* master, const void* grad, float* exp_avg, float* exp_avg_sq, int64_t n, const float* hyper,
                    const double* stats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PASERO_B200_H */
