// pasero_b200 — symbol names of the float16 build of the kernels. Every .cu of this directory is compiled twice: as it
// stands (16-bit element type = bfloat16) and with -DPB200_HALF (IEEE half), where this header renames the namespace and
// every exported entry point (suffix _f16) so that both builds link into one library. The three library-wide entry points
// of runtime.cu exist once. Generated from include/pasero_b200.h by pasero_b200/build.py: do not edit.
#pragma once
#define pb200 pb200_h
#define pb200_adam_step pb200_adam_step_f16
#define pb200_attention_bwd pb200_attention_bwd_f16
#define pb200_attention_fwd pb200_attention_fwd_f16
#define pb200_attention_workspace_bytes pb200_attention_workspace_bytes_f16
#define pb200_ce_chunk_finish pb200_ce_chunk_finish_f16
#define pb200_ce_chunk_grad pb200_ce_chunk_grad_f16
#define pb200_ce_chunk_stats pb200_ce_chunk_stats_f16
#define pb200_col2im pb200_col2im_f16
#define pb200_collate_features pb200_collate_features_f16
#define pb200_colsum_bf16 pb200_colsum_bf16_f16
#define pb200_colsum_f32 pb200_colsum_f32_f16
#define pb200_colsum_workspace_bytes pb200_colsum_workspace_bytes_f16
#define pb200_dropout_add_ln_bwd pb200_dropout_add_ln_bwd_f16
#define pb200_dropout_add_ln_fwd pb200_dropout_add_ln_fwd_f16
#define pb200_embed_bwd pb200_embed_bwd_f16
#define pb200_embed_fwd pb200_embed_fwd_f16
#define pb200_f32_to_bf16 pb200_f32_to_bf16_f16
#define pb200_fbank_cmvn pb200_fbank_cmvn_f16
#define pb200_fbank_init pb200_fbank_init_f16
#define pb200_fbank_num_frames pb200_fbank_num_frames_f16
#define pb200_fbank_workspace_bytes pb200_fbank_workspace_bytes_f16
#define pb200_gelu_bwd pb200_gelu_bwd_f16
#define pb200_gelu_fwd pb200_gelu_fwd_f16
#define pb200_gemm_bf16 pb200_gemm_bf16_f16
#define pb200_gemm_workspace_bytes pb200_gemm_workspace_bytes_f16
#define pb200_glu_bwd pb200_glu_bwd_f16
#define pb200_glu_fwd pb200_glu_fwd_f16
#define pb200_grad32_flush pb200_grad32_flush_f16
#define pb200_grad_sqnorm pb200_grad_sqnorm_f16
#define pb200_grad_sqnorm_workspace_bytes pb200_grad_sqnorm_workspace_bytes_f16
#define pb200_linear_dropout_add_ln_fwd pb200_linear_dropout_add_ln_fwd_f16
#define pb200_ln_bwd_reduce pb200_ln_bwd_reduce_f16
#define pb200_ln_bwd_workspace_bytes pb200_ln_bwd_workspace_bytes_f16
#define pb200_ls_ce_fwd_bwd pb200_ls_ce_fwd_bwd_f16
#define pb200_pad_rows pb200_pad_rows_f16
#define pb200_relu_bwd pb200_relu_bwd_f16
