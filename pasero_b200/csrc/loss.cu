// pasero_b200 — fused label-smoothed cross-entropy over bf16 logits.
//
// One CTA per target position streams the row of logits twice (second pass hits L1/L2): online
// max / sum-exp / sum-of-logits in fp32, then overwrites the logits in place with
// d(loss)/d(logits) = g * (softmax - (1-eps) onehot - eps/V) in bf16. Neither an fp32 copy of the
// logits nor the softmax probabilities are ever materialised (the reference makes an fp32 copy and two
// log-softmax passes: pasero/models/transformer.py:355-372). Returns both the smoothed loss and the
// plain NLL (the reference's second no-grad pass) plus the non-pad target count, accumulated in fp64.
#include "common.cuh"
#include "../../include/pasero_b200.h"

namespace pb200 {

static constexpr int CE_THREADS = 256;

__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = (threadIdx.x < CE_THREADS / 32) ? red[threadIdx.x] : 0.f;
    if (warp == 0) { r = warp_sum(r); if (lane == 0) red[0] = r; }
    __syncthreads();
    r = red[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_reduce_max(float v, float* red) {
    v = warp_max(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = (threadIdx.x < CE_THREADS / 32) ? red[threadIdx.x] : -INFINITY;
    if (warp == 0) { r = warp_max(r); if (lane == 0) red[0] = r; }
    __syncthreads();
    r = red[0];
    __syncthreads();
    return r;
}

// stats (fp64): [0] += sum loss (nats), [1] += sum nll (nats), [2] += #non-pad targets, [7] += #targets out of range
__global__ void __launch_bounds__(CE_THREADS)
ls_ce_kernel(bf16* __restrict__ logits, const long long* __restrict__ target, long long target_stride0,
             int T, int V, long long ld, long long padding_idx, float eps, float grad_scale, int write_grad,
             double* __restrict__ stats, float* __restrict__ row_loss) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[CE_THREADS / 32];
    const long long row = blockIdx.x;
    // targets may be a strided [B, T] view (decoder_input[:, 1:]); row = b * T + t
    const long long tgt = target[(row / T) * target_stride0 + (row % T)];
    bf16* x = logits + row * ld;
    const int v8 = (V + 7) >> 3;   // the last vector may hold padding columns (c >= V): masked below, written back as 0
    // a target id outside [0, V) (F.cross_entropy raises for it): no loss, zero gradient, and stats[7] counts it so that
    // the host raises when it reads the step's scalars (Trainer.read_logs) instead of training on a wrong loss
    const bool bad = tgt != padding_idx && (tgt < 0 || tgt >= V);
    if (bad && threadIdx.x == 0) atomicAdd(&stats[7], 1.0);
    if (tgt == padding_idx || bad) {
        if (write_grad)
            for (int i = threadIdx.x; i < v8; i += CE_THREADS) reinterpret_cast<uint4*>(x)[i] = make_uint4(0, 0, 0, 0);
        if (row_loss && threadIdx.x == 0) row_loss[row] = 0.f;
        return;
    }
    float mx = -INFINITY, sx = 0.f;
    for (int i = threadIdx.x; i < v8; i += CE_THREADS) {
        const uint4 v = reinterpret_cast<const uint4*>(x)[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(w[j]);
            const int c = i * 8 + 2 * j;
            if (c < V) { mx = fmaxf(mx, f.x); sx += f.x; }
            if (c + 1 < V) { mx = fmaxf(mx, f.y); sx += f.y; }
        }
    }
    mx = block_reduce_max(mx, red);
    sx = block_reduce_sum(sx, red);
    float se = 0.f;
    for (int i = threadIdx.x; i < v8; i += CE_THREADS) {
        const uint4 v = reinterpret_cast<const uint4*>(x)[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(w[j]);
            const int c = i * 8 + 2 * j;
            if (c < V) se += __expf(f.x - mx);
            if (c + 1 < V) se += __expf(f.y - mx);
        }
    }
    se = block_reduce_sum(se, red);
    const float lse = mx + logf(se);
    const float xt = bf16_to_f(x[tgt]);
    const float nll = lse - xt;
    const float smooth = lse - sx / V;
    const float loss = (1.f - eps) * nll + eps * smooth;
    __syncthreads();  // everyone has read x[tgt] before the row is overwritten
    if (threadIdx.x == 0) {
        atomicAdd(&stats[0], (double)loss);
        atomicAdd(&stats[1], (double)nll);
        atomicAdd(&stats[2], 1.0);
        if (row_loss) row_loss[row] = loss;
    }
    if (!write_grad) return;
    const float uni = eps / V;
    for (int i = threadIdx.x; i < v8; i += CE_THREADS) {
        const uint4 v = reinterpret_cast<const uint4*>(x)[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(w[j]);
            const int c = i * 8 + 2 * j;
            float g0 = __expf(f.x - lse) - uni, g1 = __expf(f.y - lse) - uni;
            if (c == tgt) g0 -= (1.f - eps);
            if (c + 1 == tgt) g1 -= (1.f - eps);
            if (c >= V) g0 = 0.f;
            if (c + 1 >= V) g1 = 0.f;
            o[j] = pack_bf16x2(g0 * grad_scale, g1 * grad_scale);
        }
        reinterpret_cast<uint4*>(x)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}


// ---------------------------------------------------------------------------------------------
// Chunked variant for large vocabularies (BASELINE config 5: V = 70 376): the projection is evaluated in column chunks
// that stay L2 resident, so neither the logits nor their gradient ever exist as an [M, V] tensor in HBM
// (pasero/models/transformer.py:355-372 makes an fp32 copy of them: 3 * M * V * 4 bytes). Pass 1 folds each chunk into
// per-row running statistics (online log-sum-exp), pass 2 recomputes the chunk and turns it into d loss / d logits in
// place, which the caller feeds straight into the dgrad / wgrad GEMMs of that chunk.
// Row state (fp32 [4][M]): running max, sum of exp(x - max), sum of logits, target logit.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CE_THREADS)
ce_chunk_stats_kernel(const bf16* __restrict__ logits, long long ld, const long long* __restrict__ target,
                      long long target_stride0, int T, int c0, int vc, long long padding_idx, float* __restrict__ state, int M) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[CE_THREADS / 32];
    const long long row = blockIdx.x;
    const long long tgt = target[(row / T) * target_stride0 + (row % T)];
    if (tgt == padding_idx) return;
    const bf16* x = logits + row * ld;
    const int v8 = (vc + 7) >> 3;
    float mx = -INFINITY, sx = 0.f;
    for (int i = threadIdx.x; i < v8; i += CE_THREADS) {
        const uint4 v = reinterpret_cast<const uint4*>(x)[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(w[j]);
            const int c = i * 8 + 2 * j;
            if (c < vc) { mx = fmaxf(mx, f.x); sx += f.x; }
            if (c + 1 < vc) { mx = fmaxf(mx, f.y); sx += f.y; }
        }
    }
    mx = block_reduce_max(mx, red);
    sx = block_reduce_sum(sx, red);
    float se = 0.f;
    for (int i = threadIdx.x; i < v8; i += CE_THREADS) {
        const uint4 v = reinterpret_cast<const uint4*>(x)[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(w[j]);
            const int c = i * 8 + 2 * j;
            if (c < vc) se += __expf(f.x - mx);
            if (c + 1 < vc) se += __expf(f.y - mx);
        }
    }
    se = block_reduce_sum(se, red);
    if (threadIdx.x == 0) {
        float* m = state, *s = state + M, *sum = state + 2 * M, *xt = state + 3 * M;
        if (c0 == 0) { m[row] = mx; s[row] = se; sum[row] = sx; xt[row] = 0.f; }
        else {
            const float mo = m[row], mn = fmaxf(mo, mx);
            s[row] = s[row] * __expf(mo - mn) + se * __expf(mx - mn);
            m[row] = mn;
            sum[row] += sx;
        }
        if (tgt >= c0 && tgt < c0 + vc) xt[row] = bf16_to_f(x[tgt - c0]);
    }
}

// loss / nll / token count of every row from the finished row state; lse[row] = log-sum-exp (0 for padding rows)
__global__ void ce_chunk_finish_kernel(const float* __restrict__ state, const long long* __restrict__ target,
                                       long long target_stride0, int T, int M, int V, long long padding_idx, float eps,
                                       double* __restrict__ stats, float* __restrict__ lse) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    float loss = 0.f, nll = 0.f, cnt = 0.f, bad = 0.f;
    if (row < M) {
        const long long tgt = target[(row / T) * target_stride0 + (row % T)];
        float l = 0.f;
        if (tgt != padding_idx) {
            if (tgt < 0 || tgt >= V) bad = 1.f;
            else {
                l = state[row] + logf(state[M + row]);
                nll = l - state[3 * M + row];
                loss = (1.f - eps) * nll + eps * (l - state[2 * M + row] / V);
                cnt = 1.f;
            }
        }
        lse[row] = (cnt > 0.f) ? l : INFINITY;      // +inf: exp(x - lse) = 0 and the one-hot / smoothing terms are skipped below
    }
    loss = warp_sum(loss); nll = warp_sum(nll); cnt = warp_sum(cnt); bad = warp_sum(bad);
    if ((threadIdx.x & 31) == 0) {
        if (cnt > 0.f) { atomicAdd(&stats[0], (double)loss); atomicAdd(&stats[1], (double)nll); atomicAdd(&stats[2], (double)cnt); }
        if (bad > 0.f) atomicAdd(&stats[7], (double)bad);
    }
}

// chunk of logits -> chunk of d loss / d logits, in place
__global__ void __launch_bounds__(CE_THREADS)
ce_chunk_grad_kernel(bf16* __restrict__ logits, long long ld, const long long* __restrict__ target, long long target_stride0,
                     int T, int c0, int vc, int V, float eps, float grad_scale, const float* __restrict__ lse) {
    pdl_trigger();
    pdl_wait();
    const long long row = blockIdx.x;
    const long long tgt = target[(row / T) * target_stride0 + (row % T)];
    bf16* x = logits + row * ld;
    const int v8 = (vc + 7) >> 3;
    const float l = lse[row];
    const bool live = l < INFINITY;
    const float uni = live ? eps / V : 0.f, hot = live ? (1.f - eps) : 0.f;
    for (int i = threadIdx.x; i < v8; i += CE_THREADS) {
        const uint4 v = reinterpret_cast<const uint4*>(x)[i];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(w[j]);
            const int c = i * 8 + 2 * j;
            float g0 = __expf(f.x - l) - uni, g1 = __expf(f.y - l) - uni;
            if (c0 + c == tgt) g0 -= hot;
            if (c0 + c + 1 == tgt) g1 -= hot;
            if (c >= vc) g0 = 0.f;
            if (c + 1 >= vc) g1 = 0.f;
            o[j] = pack_bf16x2(g0 * grad_scale, g1 * grad_scale);
        }
        reinterpret_cast<uint4*>(x)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_ls_ce_fwd_bwd(void* logits, const int64_t* target, int64_t target_stride0, int64_t B, int64_t T,
                                   int64_t V, int64_t ld, int64_t padding_idx, float label_smoothing, float grad_scale,
                                   int32_t write_grad, double* stats, float* row_loss, void* stream_) {
    if (B * T == 0) return 0;
    PB_CHECK(ld % 8 == 0 && ld >= (V + 7) / 8 * 8, "cross-entropy: ld (%lld) must be a multiple of 8 covering V (%lld) rounded up to 8",
             (long long)ld, (long long)V);
    PB_CHECK(stats != nullptr, "cross-entropy: stats buffer required");
    PB_CUDA(launch_pdl(ls_ce_kernel, dim3((unsigned)(B * T)), dim3(CE_THREADS), 0, (cudaStream_t)stream_,
                       (bf16*)logits, (const long long*)target, target_stride0, (int)T, (int)V, ld, padding_idx,
                       label_smoothing, grad_scale, write_grad, stats, row_loss));
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pb200_ce_chunk_stats(const void* logits, int64_t ld, const int64_t* target, int64_t target_stride0, int64_t B,
                                    int64_t T, int64_t chunk_begin, int64_t chunk_cols, int64_t padding_idx, float* row_state,
                                    void* stream_) {
    if (B * T == 0 || chunk_cols == 0) return 0;
    PB_CHECK(ld % 8 == 0 && ld >= (chunk_cols + 7) / 8 * 8 && row_state, "ce_chunk_stats: ld must be a multiple of 8 covering the chunk");
    PB_CUDA(launch_pdl(ce_chunk_stats_kernel, dim3((unsigned)(B * T)), dim3(CE_THREADS), 0, (cudaStream_t)stream_,
                       (const bf16*)logits, ld, (const long long*)target, target_stride0, (int)T, (int)chunk_begin,
                       (int)chunk_cols, padding_idx, row_state, (int)(B * T)));
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pb200_ce_chunk_finish(const float* row_state, const int64_t* target, int64_t target_stride0, int64_t B, int64_t T,
                                     int64_t V, int64_t padding_idx, float label_smoothing, double* stats, float* lse, void* stream_) {
    if (B * T == 0) return 0;
    PB_CHECK(row_state && stats && lse, "ce_chunk_finish: null pointer");
    const int M = (int)(B * T);
    ce_chunk_finish_kernel<<<(M + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(row_state, (const long long*)target, target_stride0, (int)T,
                                                                             M, (int)V, padding_idx, label_smoothing, stats, lse);
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pb200_ce_chunk_grad(void* logits, int64_t ld, const int64_t* target, int64_t target_stride0, int64_t B, int64_t T,
                                   int64_t chunk_begin, int64_t chunk_cols, int64_t V, float label_smoothing, float grad_scale,
                                   const float* lse, void* stream_) {
    if (B * T == 0 || chunk_cols == 0) return 0;
    PB_CHECK(ld % 8 == 0 && ld >= (chunk_cols + 7) / 8 * 8 && lse, "ce_chunk_grad: ld must be a multiple of 8 covering the chunk");
    PB_CUDA(launch_pdl(ce_chunk_grad_kernel, dim3((unsigned)(B * T)), dim3(CE_THREADS), 0, (cudaStream_t)stream_,
                       (bf16*)logits, ld, (const long long*)target, target_stride0, (int)T, (int)chunk_begin,
                       (int)chunk_cols, (int)V, label_smoothing, grad_scale, lse));
    PB_LAUNCH_CHECK();
    return 0;
}
