// pasero_b200 — fused gradient rescale + global-norm clip + Adam over the flat parameter arena.
//
// Semantics follow Trainer.train_step / clip_grad_norm_ / Adam.step of the reference
// (pasero/training.py:455-487, pasero/optimization.py:56-149, 390-427), including its bf16 rounding
// points: grads are bf16; `grad.mul_(dp_size / num_tokens)` and `grad.mul_(clip_coef)` each round to
// bf16; moments and the update are fp32; the parameter is rounded back to bf16.
// All step scalars live in device memory so the whole optimizer is CUDA-graph replayable and needs no
// host synchronisation (the reference does ~10 launches x 353 tensors and two host syncs).
//
//   hyper (fp32[16]): [0] lr  [1] beta1  [2] beta2  [3] eps  [4] weight_decay  [5] 1 - beta1^t
//                     [6] sqrt(1 - beta2^t)  [7] clip_norm  [8] grad multiplier numerator (dp_size)
//                     [9] 1 - beta1  [10] 1 - beta2  [11] lr / (1 - beta1^t)  [12] weight_decay * lr
//                     [13] 1 / loss scale (GradScaler.unscale_, 1 when there is no loss scaling)
//                     [14] != 0: fp32-gradient semantics (MixedPrecisionAdam / FlatFP16Adam, optimization.py:167-298): the
//                          16-bit gradients are converted to fp32 BEFORE unscaling / normalising / clipping, so those
//                          multiplications do not round to 16 bits, and the update is applied to an fp32 master copy
//                     (derived scalars are computed by the host in fp64, like the reference's Python floats)
//   stats (fp64[8]) : [2] num_tokens (divisor; <=0 -> 1)   [4] out: sum of squares of the scaled grads
//                     [5] out: 1 when that sum is not finite (an inf / nan gradient: GradScaler's found_inf) — the Adam
//                         kernel then leaves parameters and moments untouched (GradScaler.step skips the update)
//                     [6] in: number of micro-batches that failed on some rank (data loading, CUDA OOM; summed over the ranks
//                         by the caller): the step must not update anything then (pasero/training.py:444-446) -> [5] is raised too
#include "common.cuh"
#include "../../include/pasero_b200.h"

namespace pb200 {

__device__ __forceinline__ float grad_coeff(const float* hyper, const double* stats) {
    const double nt = stats[2];
    return (float)((double)hyper[8] / (nt > 0.0 ? nt : 1.0));
}

// Deterministic two-phase reduction: every block writes its partial sum, the last block to finish (atomic ticket)
// adds the partials in index order. Data-parallel replicas therefore compute bit-identical norms from their
// bit-identical (allreduced) gradients and never drift apart.
__global__ void __launch_bounds__(256)
grad_sqnorm_kernel(const bf16* __restrict__ grad, long long n8, const float* __restrict__ hyper,
                   double* __restrict__ stats, double* __restrict__ partials, unsigned int* __restrict__ ticket) {
    pdl_trigger();
    pdl_wait();
    const float coeff = grad_coeff(hyper, stats);
    const float inv_scale = hyper[13] != 0.f ? hyper[13] : 1.0f;    // (0 = not set: no loss scaling)
    const bool f32g = hyper[14] != 0.f;
    float acc = 0.f;
    auto add = [&](const uint4& v) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(w[j]);
            float a, b;
            if (f32g) {
                a = f.x * inv_scale * coeff; b = f.y * inv_scale * coeff;
            } else {
                a = f.x; b = f.y;
                if (inv_scale != 1.0f) { a = round_bf16(a * inv_scale); b = round_bf16(b * inv_scale); }
                a = round_bf16(a * coeff); b = round_bf16(b * coeff);
            }
            acc += a * a + b * b;
        }
    };
    // four 16-byte loads in flight per thread: with one, the 4 x 256-thread CTAs an SM holds keep 16 KB outstanding, which at
    // HBM latency is a third of the bandwidth (the arena is read once, nothing else hides the latency)
    const long long stride = (long long)gridDim.x * blockDim.x;
    const uint4* g4 = reinterpret_cast<const uint4*>(grad);
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n8; i += 4 * stride) {
        const uint4 v0 = __ldg(g4 + i), v1 = __ldg(g4 + i + stride), v2 = __ldg(g4 + i + 2 * stride), v3 = __ldg(g4 + i + 3 * stride);
        add(v0); add(v1); add(v2); add(v3);
    }
    for (; i < n8; i += stride) add(__ldg(g4 + i));
    __shared__ double red[8];
    __shared__ bool is_last;
    double d = warp_sum_d((double)acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        partials[blockIdx.x] = t;
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    // the last CTA adds the partials: every thread its share (a fixed order, so replicas stay bit-identical), then the block;
    // one thread walking all of them is a chain of dependent L2 round trips as long as the rest of the kernel
    __threadfence();
    double t = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) t += reinterpret_cast<volatile double*>(partials)[i];
    t = warp_sum_d(t);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        stats[4] = t;
        stats[5] = (isfinite(t) && stats[6] == 0.0) ? 0.0 : 1.0;     // [6]: a failed micro-batch on some rank -> no update either
        *ticket = 0;
    }
}

__global__ void __launch_bounds__(256)
adam_kernel(bf16* __restrict__ param, float* __restrict__ master, const bf16* __restrict__ grad, float* __restrict__ exp_avg,
            float* __restrict__ exp_avg_sq, long long n8, const float* __restrict__ hyper,
            const double* __restrict__ stats) {
    pdl_trigger();
    pdl_wait();
    if (stats[5] != 0.0) return;      // an inf / nan gradient was found: the step is skipped (GradScaler.step)
    const float inv_scale = hyper[13] != 0.f ? hyper[13] : 1.0f;
    const bool f32g = master != nullptr;
    const float beta1 = hyper[1], beta2 = hyper[2], eps = hyper[3];
    const float bc2_sqrt = hyper[6], clip = hyper[7];
    const float omb1 = hyper[9], omb2 = hyper[10], step_size = hyper[11], wd_lr = hyper[12];
    const float coeff = grad_coeff(hyper, stats);
    float clip_coef = 1.0f;
    if (clip > 0.f) {
        const float total_norm = (float)sqrt(stats[4]);
        clip_coef = fminf(clip / (total_norm + 1e-6f), 1.0f);
    }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const uint4 gv = __ldg(reinterpret_cast<const uint4*>(grad) + i);
        const uint4 pv = reinterpret_cast<const uint4*>(param)[i];
        float4 m0 = reinterpret_cast<const float4*>(exp_avg)[2 * i], m1 = reinterpret_cast<const float4*>(exp_avg)[2 * i + 1];
        float4 v0 = reinterpret_cast<const float4*>(exp_avg_sq)[2 * i], v1 = reinterpret_cast<const float4*>(exp_avg_sq)[2 * i + 1];
        const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w}, pw[4] = {pv.x, pv.y, pv.z, pv.w};
        float m[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
        float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        float g[8], p[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 gf = unpack_bf16x2(gw[j]);
            const float2 pf = unpack_bf16x2(pw[j]);
            g[2 * j] = gf.x; g[2 * j + 1] = gf.y; p[2 * j] = pf.x; p[2 * j + 1] = pf.y;
        }
        if (f32g) {   // the update works on the fp32 master copy; the 16-bit parameters are its rounded image
            const float4 q0 = reinterpret_cast<const float4*>(master)[2 * i], q1 = reinterpret_cast<const float4*>(master)[2 * i + 1];
            p[0] = q0.x; p[1] = q0.y; p[2] = q0.z; p[3] = q0.w; p[4] = q1.x; p[5] = q1.y; p[6] = q1.z; p[7] = q1.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float gj;
            if (f32g) {
                gj = g[j] * inv_scale * coeff;
                if (clip > 0.f) gj *= clip_coef;
            } else {
                gj = g[j];
                if (inv_scale != 1.0f) gj = round_bf16(gj * inv_scale);
                gj = round_bf16(gj * coeff);
                if (clip > 0.f) gj = round_bf16(gj * clip_coef);
            }
            m[j] = m[j] * beta1 + gj * omb1;
            v[j] = v[j] * beta2 + gj * gj * omb2;
            const float denom = sqrtf(v[j]) / bc2_sqrt + eps;
            float pj = p[j];
            if (wd_lr != 0.f) pj += pj * (-wd_lr);
            pj += -step_size * (m[j] / denom);
            p[j] = pj;
        }
        uint4 po;
        po.x = pack_bf16x2(p[0], p[1]); po.y = pack_bf16x2(p[2], p[3]);
        po.z = pack_bf16x2(p[4], p[5]); po.w = pack_bf16x2(p[6], p[7]);
        reinterpret_cast<uint4*>(param)[i] = po;
        if (f32g) {
            reinterpret_cast<float4*>(master)[2 * i] = make_float4(p[0], p[1], p[2], p[3]);
            reinterpret_cast<float4*>(master)[2 * i + 1] = make_float4(p[4], p[5], p[6], p[7]);
        }
        reinterpret_cast<float4*>(exp_avg)[2 * i] = make_float4(m[0], m[1], m[2], m[3]);
        reinterpret_cast<float4*>(exp_avg)[2 * i + 1] = make_float4(m[4], m[5], m[6], m[7]);
        reinterpret_cast<float4*>(exp_avg_sq)[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(exp_avg_sq)[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
    }
}

// fp32 main-gradient ranges -> bf16 gradient arena (overwrite); the fp32 source is zeroed so that the next step's
// weight-gradient GEMMs can reduce-add into it again. ranges: [n][2] (begin, end) element offsets, multiples of 8.
__global__ void __launch_bounds__(256)
grad32_flush_kernel(float* __restrict__ acc, bf16* __restrict__ out, const long long* __restrict__ ranges) {
    pdl_trigger();
    pdl_wait();
    const long long begin = ranges[2 * blockIdx.y], end = ranges[2 * blockIdx.y + 1];
    const long long n8 = (end - begin) >> 3;
    float4* a4 = reinterpret_cast<float4*>(acc + begin);
    uint4* o4 = reinterpret_cast<uint4*>(out + begin);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = a4[2 * i], b = a4[2 * i + 1];
        uint4 o;
        o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w); o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
        o4[i] = o;
        a4[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
        a4[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_grad32_flush(float* acc32, void* grad_bf16, const int64_t* ranges, int32_t n_ranges, int64_t max_len,
                                  void* stream_) {
    if (n_ranges <= 0 || max_len <= 0) return 0;
    PB_CHECK(acc32 && grad_bf16 && ranges, "grad32_flush: null pointer");
    long long bx = (max_len / 8 + 255) / 256;
    const long long cap = (8LL * num_sms() + n_ranges - 1) / n_ranges;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)n_ranges);
    PB_CUDA(launch_pdl(grad32_flush_kernel, grid, dim3(256), 0, (cudaStream_t)stream_, acc32, (bf16*)grad_bf16, (const long long*)ranges));
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t pb200_grad_sqnorm_workspace_bytes(void) { return (size_t)(4 * num_sms() + 2) * sizeof(double); }

// stats[4] = sum of squares of bf16(grad * dp_size / num_tokens); workspace must be zero-initialised once (the
// kernel leaves it ready for the next call).
extern "C" int pb200_grad_sqnorm(const void* grad, int64_t n, const float* hyper, double* stats, void* workspace,
                                 size_t workspace_bytes, void* stream_) {
    if (n == 0) return 0;
    PB_CHECK(n % 8 == 0, "gradient arena length (%lld) must be a multiple of 8", (long long)n);
    PB_CHECK(workspace && workspace_bytes >= pb200_grad_sqnorm_workspace_bytes(), "grad_sqnorm workspace too small");
    long long blocks = (n / 8 + 255) / 256;
    if (blocks > 4LL * num_sms()) blocks = 4LL * num_sms();
    double* partials = reinterpret_cast<double*>(workspace) + 1;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
    PB_CUDA(launch_pdl(grad_sqnorm_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream_, (const bf16*)grad, (long long)(n / 8), hyper, stats, partials, ticket));
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pb200_adam_step(void* param, float* master, const void* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                               const float* hyper, const double* stats, void* stream_) {
    if (n == 0) return 0;
    PB_CHECK(n % 8 == 0, "parameter arena length (%lld) must be a multiple of 8", (long long)n);
    long long blocks = (n / 8 + 255) / 256;
    if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
    PB_CUDA(launch_pdl(adam_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream_, (bf16*)param, master, (const bf16*)grad, exp_avg, exp_avg_sq,
                       (long long)(n / 8), hyper, stats));
    PB_LAUNCH_CHECK();
    return 0;
}
