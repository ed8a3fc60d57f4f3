// pasero_b200 — C entry points of the fused multi-head attention (forward + backward), head_dim 64, bf16 in / fp32
// softmax; the kernels are in attention_tc.cu (tcgen05 / TMEM, every sequence length; round 1's mma.sync fallback for
// more than 256 keys is gone).
//
// Attention that takes key lengths instead of a materialised mask tensor:
//   * encoder self-attention and decoder cross-attention: key-padding mask from `kv_len[b]`
//     (reference builds a float -inf mask (B,1,1,S) and calls SDPA: pasero/models/modules.py:659-677,712-718)
//   * decoder self-attention: causal, padding mask dropped (modules.py:602-605, 688)
//   * attention dropout on the probabilities (modules.py:686, 702, 716, 731)
// Q/K/V are read in place from the fused projection outputs ([B*L, ld] rows, head h at column h*64), the
// output is written as [B*Lq, H*64] so the out-projection GEMM consumes it without a transpose.
// Backward rebuilds P from the saved log-sum-exp (no O(L^2) tensor is stored apart from the 1-bit dropout mask).
#include "common.cuh"
#include "../../include/pasero_b200.h"

namespace pb200 {
int attn_fwd_tc(const pb200_attn_params* q, cudaStream_t stream);
int attn_bwd_tc(const pb200_attn_params* q, cudaStream_t stream);
size_t attn_tc_workspace_bytes(const pb200_attn_params* q, bool backward);

static int check_attn(const pb200_attn_params* q) {
    PB_CHECK(q != nullptr, "null params");
    PB_CHECK(q->head_dim == 64, "attention kernels support head_dim 64 only (got %d)", q->head_dim);
    PB_CHECK(q->ldq % 8 == 0 && q->ldk % 8 == 0 && q->ldv % 8 == 0 && q->ldo % 8 == 0, "attention strides must be multiples of 8");
    PB_CHECK(q->B >= 0 && q->H > 0 && q->Lq >= 0 && q->Lk >= 0, "bad attention shape");
    return 0;
}
}  // namespace pb200

using namespace pb200;

extern "C" size_t pb200_attention_workspace_bytes(const pb200_attn_params* q, int32_t backward) {
    if (q == nullptr || q->B <= 0 || q->Lq <= 0 || q->Lk <= 0) return 0;
    return attn_tc_workspace_bytes(q, backward != 0);
}

extern "C" int pb200_attention_fwd(const pb200_attn_params* q, void* stream_) {
    if (int rc = check_attn(q)) return rc;
    if (q->B == 0 || q->Lq == 0) return 0;
    PB_CHECK(q->q && q->k && q->v && q->o, "attention_fwd: null tensor");
    return attn_fwd_tc(q, (cudaStream_t)stream_);
}

extern "C" int pb200_attention_bwd(const pb200_attn_params* q, void* stream_) {
    if (int rc = check_attn(q)) return rc;
    if (q->B == 0 || q->Lq == 0) return 0;
    PB_CHECK(q->q && q->k && q->v && q->o && q->dout && q->lse && q->dq && q->dk && q->dv, "attention_bwd: null tensor");
    PB_CHECK(q->lddo % 8 == 0 && q->lddq % 8 == 0 && q->lddk % 8 == 0 && q->lddv % 8 == 0, "attention strides must be multiples of 8");
    return attn_bwd_tc(q, (cudaStream_t)stream_);
}
