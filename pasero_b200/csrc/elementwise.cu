// pasero_b200 — memory-bound elementwise / gather kernels of the ST encoder-decoder:
//   * embedding epilogue: (gather | features) * sqrt(d) + positional rows, dropout
//     (pasero/models/transformer.py:739-744, 874-878; modules.py:415-457, 916-933)
//   * GLU of the conv subsampler, written straight into the zero-padded layout that the next
//     convolution's TMA descriptor reads (pasero/models/modules.py:829-834)
//   * col2im gather for the strided-conv input gradient
// All tensors bf16, 16-byte vectorised, fp32 math with the reference's bf16 rounding points.
#include "common.cuh"
#include <cuda_fp16.h>
#include "../../include/pasero_b200.h"

namespace pb200 {

// out[r, :] = dropout( bf16( bf16(src_row * scale) + pos[r % L, :] ) ),  src_row = table[ids[r]] or x[r]
__global__ void embed_fwd_kernel(const long long* __restrict__ ids, const bf16* __restrict__ src,
                                 const bf16* __restrict__ pos, long long M, int D, int L, float scale, float p,
                                 uint64_t seed, uint64_t offset, const long long* __restrict__ offset_base,
                                 long long num_embeddings, bf16* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    if (offset_base) offset += (uint64_t)(*offset_base) << 12;
    const Philox rng(seed);
    const uint32_t thresh = dropout_thresh16(p);
    const float keep_scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    const int d8 = D >> 3;
    const long long total = M * d8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / d8;
        const int c = (int)(i % d8) * 8;
        long long srow = r;
        if (ids) {
            srow = ids[r];
            if (srow < 0) srow = 0;                         // reference clips ids at 0 (modules.py:923)
            if (srow >= num_embeddings) srow = num_embeddings - 1;  // reference asserts; never trap on device
        }
        const uint4 xv = __ldg(reinterpret_cast<const uint4*>(src + srow * D + c));
        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
        uint32_t pw[4] = {0, 0, 0, 0};
        if (pos) {
            const uint4 pv = __ldg(reinterpret_cast<const uint4*>(pos + (long long)(r % L) * D + c));
            pw[0] = pv.x; pw[1] = pv.y; pw[2] = pv.z; pw[3] = pv.w;
        }
        uint32_t keep = 0xffu;
        if (p > 0.f) keep = dropout_keep8(rng, offset, (uint64_t)i, thresh);
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 a = unpack_bf16x2(xw[j]);
            const float2 q = unpack_bf16x2(pw[j]);
            float v0 = round_bf16(round_bf16(a.x * scale) + q.x);
            float v1 = round_bf16(round_bf16(a.y * scale) + q.y);
            if (p > 0.f) {
                v0 = ((keep >> (2 * j)) & 1) ? v0 * keep_scale : 0.f;
                v1 = ((keep >> (2 * j + 1)) & 1) ? v1 * keep_scale : 0.f;
            }
            ow[j] = pack_bf16x2(v0, v1);
        }
        *reinterpret_cast<uint4*>(out + r * D + c) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// dsrc = bf16( bf16(dy * mask / (1-p)) * scale ); with ids: scatter-add (fp32 atomics) into dtable
__global__ void embed_bwd_kernel(const bf16* __restrict__ dy, const long long* __restrict__ ids, long long M, int D,
                                 float scale, float p, uint64_t seed, uint64_t offset, const long long* __restrict__ offset_base,
                                 long long padding_idx, long long num_embeddings, bf16* __restrict__ dx,
                                 float* __restrict__ dtable) {
    pdl_trigger();
    pdl_wait();
    if (offset_base) offset += (uint64_t)(*offset_base) << 12;
    const Philox rng(seed);
    const uint32_t thresh = dropout_thresh16(p);
    const float keep_scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    const int d8 = D >> 3;
    const long long total = M * d8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / d8;
        const int c = (int)(i % d8) * 8;
        const uint4 dv = __ldg(reinterpret_cast<const uint4*>(dy + r * D + c));
        const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
        uint32_t keep = 0xffu;
        if (p > 0.f) keep = dropout_keep8(rng, offset, (uint64_t)i, thresh);
        float g[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 a = unpack_bf16x2(dw[j]);
            float v0 = a.x, v1 = a.y;
            if (p > 0.f) {
                v0 = ((keep >> (2 * j)) & 1) ? round_bf16(v0 * keep_scale) : 0.f;
                v1 = ((keep >> (2 * j + 1)) & 1) ? round_bf16(v1 * keep_scale) : 0.f;
            }
            g[2 * j] = round_bf16(v0 * scale);
            g[2 * j + 1] = round_bf16(v1 * scale);
        }
        if (ids) {
            long long row = ids[r];
            if (row < 0) row = 0;
            if (row >= num_embeddings) row = num_embeddings - 1;
            if (row == padding_idx) continue;
            float* dst = dtable + row * D + c;
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(dst + j, g[j]);
        } else {
            uint4 ov;
            ov.x = pack_bf16x2(g[0], g[1]); ov.y = pack_bf16x2(g[2], g[3]);
            ov.z = pack_bf16x2(g[4], g[5]); ov.w = pack_bf16x2(g[6], g[7]);
            *reinterpret_cast<uint4*>(dx + r * D + c) = ov;
        }
    }
}

// GLU over channels: in [B*T, 2C] (value | gate) -> out[b, front_pad + t, c] of a [B, Tp, C] buffer whose
// remaining rows are written as zeros (the next conv's implicit padding).
__global__ void glu_fwd_kernel(const bf16* __restrict__ in, int B, int T, int C, bf16* __restrict__ out, int Tp,
                               int front_pad) {
    pdl_trigger();
    pdl_wait();
    const int c8 = C >> 3;
    const long long total = (long long)B * Tp * c8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c8) * 8;
        const long long rp = i / c8;
        const int tp = (int)(rp % Tp);
        const int b = (int)(rp / Tp);
        const int t = tp - front_pad;
        uint4 ov = make_uint4(0, 0, 0, 0);
        if (t >= 0 && t < T) {
            const bf16* row = in + ((long long)b * T + t) * (2 * C);
            const uint4 av = __ldg(reinterpret_cast<const uint4*>(row + c));
            const uint4 gv = __ldg(reinterpret_cast<const uint4*>(row + C + c));
            const uint32_t aw[4] = {av.x, av.y, av.z, av.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
            uint32_t ow[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 a = unpack_bf16x2(aw[j]);
                const float2 g = unpack_bf16x2(gw[j]);
                ow[j] = pack_bf16x2(a.x / (1.0f + __expf(-g.x)), a.y / (1.0f + __expf(-g.y)));
            }
            ov = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        *reinterpret_cast<uint4*>(out + rp * C + c) = ov;
    }
}

// d(value) = dy * sigmoid(g); d(gate) = dy * a * sigmoid(g) * (1 - sigmoid(g)); dy is [B*T, C] contiguous
__global__ void glu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ in, long long M, int C,
                               bf16* __restrict__ din) {
    pdl_trigger();
    pdl_wait();
    const int c8 = C >> 3;
    const long long total = M * c8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c8) * 8;
        const long long r = i / c8;
        const uint4 dv = __ldg(reinterpret_cast<const uint4*>(dy + r * C + c));
        const uint4 av = __ldg(reinterpret_cast<const uint4*>(in + r * 2 * C + c));
        const uint4 gv = __ldg(reinterpret_cast<const uint4*>(in + r * 2 * C + C + c));
        const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w}, aw[4] = {av.x, av.y, av.z, av.w},
                       gw[4] = {gv.x, gv.y, gv.z, gv.w};
        uint32_t oa[4], og[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 d = unpack_bf16x2(dw[j]);
            const float2 a = unpack_bf16x2(aw[j]);
            const float2 g = unpack_bf16x2(gw[j]);
            const float s0 = 1.0f / (1.0f + __expf(-g.x)), s1 = 1.0f / (1.0f + __expf(-g.y));
            oa[j] = pack_bf16x2(d.x * s0, d.y * s1);
            og[j] = pack_bf16x2(d.x * a.x * s0 * (1.f - s0), d.y * a.y * s1 * (1.f - s1));
        }
        *reinterpret_cast<uint4*>(din + r * 2 * C + c) = make_uint4(oa[0], oa[1], oa[2], oa[3]);
        *reinterpret_cast<uint4*>(din + r * 2 * C + C + c) = make_uint4(og[0], og[1], og[2], og[3]);
    }
}

// GELU (erf) forward into a (possibly zero-padded) [B, Tp, C] layout: out[b, front_pad + t] = gelu(in[b, t]) for
// t < T, zeros elsewhere — the Whisper-style conv front end (the next conv's TMA view reads the padded layout) and,
// with B = 1, Tp = T, the FFN activation.
__global__ void gelu_fwd_kernel(const bf16* __restrict__ in, int B, int T, int C, bf16* __restrict__ out, int Tp,
                                int front_pad) {
    pdl_trigger();
    pdl_wait();
    const int c8 = C >> 3;
    const long long total = (long long)B * Tp * c8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c8) * 8;
        const long long rp = i / c8;
        const int tp = (int)(rp % Tp);
        const int b = (int)(rp / Tp);
        const int t = tp - front_pad;
        uint4 ov = make_uint4(0, 0, 0, 0);
        if (t >= 0 && t < T) {
            const uint4 xv = __ldg(reinterpret_cast<const uint4*>(in + ((long long)b * T + t) * C + c));
            const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
            uint32_t ow[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 x = unpack_bf16x2(xw[j]);
                ow[j] = pack_bf16x2(0.5f * x.x * (1.f + erff(x.x * 0.70710678f)), 0.5f * x.y * (1.f + erff(x.y * 0.70710678f)));
            }
            ov = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        *reinterpret_cast<uint4*>(out + rp * C + c) = ov;
    }
}

// dx = dy where the ReLU output h > 0, else 0 (backward of a stand-alone ReLU, e.g. the encoder's in_linear)
__global__ void relu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ h, long long n8, bf16* __restrict__ dx) {
    pdl_trigger();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const uint4 dv = __ldg(reinterpret_cast<const uint4*>(dy) + i);
        const uint4 hv = __ldg(reinterpret_cast<const uint4*>(h) + i);
        const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w}, hw[4] = {hv.x, hv.y, hv.z, hv.w};
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // bf16 > 0  <=>  sign bit clear and not +0
            const uint32_t lo = ((int)(hw[j] << 16) > 0) ? 0x0000ffffu : 0u;
            const uint32_t hi = ((int)(hw[j] & 0xffff0000u) > 0) ? 0xffff0000u : 0u;
            ow[j] = dw[j] & (lo | hi);
        }
        reinterpret_cast<uint4*>(dx)[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// GELU (erf) elementwise fwd/bwd for the Whisper-style conv front end and FFN backward
__global__ void gelu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ pre, long long n8,
                                bf16* __restrict__ dx) {
    pdl_trigger();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
         i += (long long)gridDim.x * blockDim.x) {
        const uint4 dv = __ldg(reinterpret_cast<const uint4*>(dy) + i);
        const uint4 xv = __ldg(reinterpret_cast<const uint4*>(pre) + i);
        const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w}, xw[4] = {xv.x, xv.y, xv.z, xv.w};
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 d = unpack_bf16x2(dw[j]);
            const float2 x = unpack_bf16x2(xw[j]);
            const float c0 = 0.5f * (1.f + erff(x.x * 0.70710678f)) + x.x * 0.3989422804f * __expf(-0.5f * x.x * x.x);
            const float c1 = 0.5f * (1.f + erff(x.y * 0.70710678f)) + x.y * 0.3989422804f * __expf(-0.5f * x.y * x.y);
            ow[j] = pack_bf16x2(d.x * c0, d.y * c1);
        }
        reinterpret_cast<uint4*>(dx)[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// col2im for conv1d(kernel=KW, stride, padding=KW/2): dcol [B, Tout, KW, C] -> dx [B, Tin, C]
//   dx[b, s, c] = sum_{tap : (s + pad - tap) % stride == 0, t = (s + pad - tap)/stride in [0,Tout)} dcol[b,t,tap,c]
__global__ void col2im_kernel(const bf16* __restrict__ dcol, int B, int Tin, int Tout, int C, int KW, int stride,
                              bf16* __restrict__ dx) {
    pdl_trigger();
    pdl_wait();
    const int c8 = C >> 3;
    const int pad = KW / 2;
    const long long total = (long long)B * Tin * c8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c8) * 8;
        const long long r = i / c8;
        const int s = (int)(r % Tin);
        const int b = (int)(r / Tin);
        float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int tap = 0; tap < KW; ++tap) {
            const int u = s + pad - tap;
            if (u < 0 || (u % stride) != 0) continue;
            const int t = u / stride;
            if (t >= Tout) continue;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(dcol + (((long long)b * Tout + t) * KW + tap) * C + c));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(w[j]);
                a[2 * j] += f.x; a[2 * j + 1] += f.y;
            }
        }
        uint4 ov;
        ov.x = pack_bf16x2(a[0], a[1]); ov.y = pack_bf16x2(a[2], a[3]);
        ov.z = pack_bf16x2(a[4], a[5]); ov.w = pack_bf16x2(a[6], a[7]);
        *reinterpret_cast<uint4*>(dx + r * C + c) = ov;
    }
}

// [B, T, C] bf16 -> zero-padded [B, Tp, C] bf16 (front_pad zero rows, zeros after T)
__global__ void pad_rows_kernel(const bf16* __restrict__ in, int B, int T, int C, bf16* __restrict__ out, int Tp,
                                int front_pad) {
    pdl_trigger();
    pdl_wait();
    const int c8 = C >> 3;
    const long long total = (long long)B * Tp * c8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c8) * 8;
        const long long rp = i / c8;
        const int t = (int)(rp % Tp) - front_pad;
        const int b = (int)(rp / Tp);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (t >= 0 && t < T) v = __ldg(reinterpret_cast<const uint4*>(in + ((long long)b * T + t) * C + c));
        *reinterpret_cast<uint4*>(out + rp * C + c) = v;
    }
}

// fp32 -> bf16 with optional accumulate into the bf16 destination (embedding gradient fold-in)
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, long long n, bf16* __restrict__ out, int accumulate) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = in[i];
        if (accumulate) v += bf16_to_f(out[i]);
        out[i] = f_to_bf16(v);
    }
}

static int ew_grid(long long work_items, int threads = 256) {
    long long g = (work_items + threads - 1) / threads;
    const long long cap = 8LL * num_sms();
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_embed_fwd(const int64_t* ids, const void* src, const void* pos, int64_t M, int64_t D, int64_t L,
                               float scale, float p, uint64_t seed, uint64_t offset, const int64_t* offset_base,
                               int64_t num_embeddings, void* out, void* stream_) {
    if (M == 0) return 0;
    PB_CHECK(D % 8 == 0, "embed: D must be a multiple of 8");
    PB_CHECK(L > 0, "embed: L must be positive");
    PB_CUDA(launch_pdl(embed_fwd_kernel, dim3(ew_grid(M * (D / 8))), dim3(256), 0, (cudaStream_t)stream_,
                       (const long long*)ids, (const bf16*)src, (const bf16*)pos, M, (int)D, (int)L, scale, p, seed, offset,
                       (const long long*)offset_base, num_embeddings, (bf16*)out));
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pb200_embed_bwd(const void* dy, const int64_t* ids, int64_t M, int64_t D, float scale, float p,
                               uint64_t seed, uint64_t offset, const int64_t* offset_base, int64_t padding_idx,
                               int64_t num_embeddings, void* dx, float* dtable, void* stream_) {
    if (M == 0) return 0;
    PB_CHECK(D % 8 == 0, "embed: D must be a multiple of 8");
    PB_CHECK((ids != nullptr) == (dtable != nullptr), "embed_bwd: ids and dtable go together");
    PB_CHECK(ids != nullptr || dx != nullptr, "embed_bwd: dx required without ids");
    PB_CUDA(launch_pdl(embed_bwd_kernel, dim3(ew_grid(M * (D / 8))), dim3(256), 0, (cudaStream_t)stream_,
                       (const bf16*)dy, (const long long*)ids, M, (int)D, scale, p, seed, offset,
                       (const long long*)offset_base, padding_idx, num_embeddings, (bf16*)dx, dtable));
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pb200_glu_fwd(const void* in, int32_t B, int32_t T, int32_t C, void* out, int32_t Tp, int32_t front_pad,
                             void* stream_) {
    if (B == 0) return 0;
    PB_CHECK(C % 8 == 0 && Tp >= T + front_pad, "glu: bad shape");
    PB_CUDA(launch_pdl(glu_fwd_kernel, dim3(ew_grid((long long)B * Tp * (C / 8))), dim3(256), 0, (cudaStream_t)stream_,
                       (const bf16*)in, B, T, C, (bf16*)out, Tp, front_pad));
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pb200_glu_bwd(const void* dy, const void* in, int64_t M, int32_t C, void* din, void* stream_) {
    if (M == 0) return 0;
    PB_CHECK(C % 8 == 0, "glu: C must be a multiple of 8");
    PB_CUDA(launch_pdl(glu_bwd_kernel, dim3(ew_grid(M * (C / 8))), dim3(256), 0, (cudaStream_t)stream_,
                       (const bf16*)dy, (const bf16*)in, M, C, (bf16*)din));
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pb200_relu_bwd(const void* dy, const void* h, int64_t n, void* dx, void* stream_) {
    if (n == 0) return 0;
    PB_CHECK(n % 8 == 0, "relu_bwd: n must be a multiple of 8");
    PB_CUDA(launch_pdl(relu_bwd_kernel, dim3(ew_grid(n / 8)), dim3(256), 0, (cudaStream_t)stream_,
                       (const bf16*)dy, (const bf16*)h, n / 8, (bf16*)dx));
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pb200_gelu_fwd(const void* in, int32_t B, int32_t T, int32_t C, void* out, int32_t Tp, int32_t front_pad,
                              void* stream_) {
    if (B == 0 || T == 0) return 0;
    PB_CHECK(C % 8 == 0 && Tp >= T + front_pad, "gelu_fwd: bad shape");
    PB_CUDA(launch_pdl(gelu_fwd_kernel, dim3(ew_grid((long long)B * Tp * (C / 8))), dim3(256), 0, (cudaStream_t)stream_,
                       (const bf16*)in, B, T, C, (bf16*)out, Tp, front_pad));
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pb200_gelu_bwd(const void* dy, const void* pre, int64_t n, void* dx, void* stream_) {
    if (n == 0) return 0;
    PB_CHECK(n % 8 == 0, "gelu_bwd: n must be a multiple of 8");
    PB_CUDA(launch_pdl(gelu_bwd_kernel, dim3(ew_grid(n / 8)), dim3(256), 0, (cudaStream_t)stream_,
                       (const bf16*)dy, (const bf16*)pre, n / 8, (bf16*)dx));
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pb200_col2im(const void* dcol, int32_t B, int32_t Tin, int32_t Tout, int32_t C, int32_t KW,
                            int32_t stride, void* dx, void* stream_) {
    if (B == 0) return 0;
    PB_CHECK(C % 8 == 0 && stride >= 1 && KW >= 1, "col2im: bad shape");
    PB_CUDA(launch_pdl(col2im_kernel, dim3(ew_grid((long long)B * Tin * (C / 8))), dim3(256), 0, (cudaStream_t)stream_,
                       (const bf16*)dcol, B, Tin, Tout, C, KW, stride, (bf16*)dx));
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pb200_pad_rows(const void* in, int32_t B, int32_t T, int32_t C, void* out, int32_t Tp, int32_t front_pad,
                              void* stream_) {
    if (B == 0) return 0;
    PB_CHECK(C % 8 == 0 && Tp >= T + front_pad, "pad_rows: bad shape");
    PB_CUDA(launch_pdl(pad_rows_kernel, dim3(ew_grid((long long)B * Tp * (C / 8))), dim3(256), 0, (cudaStream_t)stream_,
                       (const bf16*)in, B, T, C, (bf16*)out, Tp, front_pad));
    PB_LAUNCH_CHECK();
    return 0;
}

// Device-side collate of a feature batch (Task.collate's right-padding + `.to(dtype)`, pasero/tasks/task.py:531-571,
// pasero/utils.py:709-736): the utterances' rows arrive back to back (fp16 as the reference stores them in a NumpyFile, or
// fp32), `offsets` [B + 1] (rows) delimits them; utterance b lands at rows [front_pad, front_pad + len_b) of out[b], bf16,
// zeros elsewhere — either the plain (B, Tp, C) batch tensor (front_pad = 0) or the conv subsampler's padded layout.
template <typename T>
__global__ void collate_features_kernel(const T* __restrict__ rows, const long long* __restrict__ offsets, int B, int C,
                                        bf16* __restrict__ out, int Tp, int front_pad) {
    const int c8 = C >> 3;
    const long long total = (long long)B * Tp * c8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c8) * 8;
        const long long rp = i / c8;
        const int t = (int)(rp % Tp) - front_pad;
        const int b = (int)(rp / Tp);
        const long long begin = offsets[b];
        const int len = (int)(offsets[b + 1] - begin);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (t >= 0 && t < len) {
            const T* src = rows + (begin + t) * C + c;
            float f[8];
            if (sizeof(T) == 2) {
                const uint4 h = __ldg(reinterpret_cast<const uint4*>(src));
                const __half2* h2 = reinterpret_cast<const __half2*>(&h);
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float2 x = __half22float2(h2[e]); f[2 * e] = x.x; f[2 * e + 1] = x.y; }
            } else {
                const float4 a = __ldg(reinterpret_cast<const float4*>(src)), bq = __ldg(reinterpret_cast<const float4*>(src) + 1);
                f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = bq.x; f[5] = bq.y; f[6] = bq.z; f[7] = bq.w;
            }
            v = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
        *reinterpret_cast<uint4*>(out + rp * C + c) = v;
    }
}

extern "C" int pb200_collate_features(const void* rows, int32_t rows_fp16, const int64_t* offsets, int32_t B, int32_t C, void* out,
                                      int32_t Tp, int32_t front_pad, void* stream_) {
    if (B == 0 || Tp == 0) return 0;
    PB_CHECK(C % 8 == 0 && rows && offsets && out, "collate_features: C must be a multiple of 8, pointers non-null");
    const auto grid = ew_grid((long long)B * Tp * (C / 8));
    if (rows_fp16)
        collate_features_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream_>>>((const __half*)rows, (const long long*)offsets, B, C,
                                                                              (bf16*)out, Tp, front_pad);
    else
        collate_features_kernel<float><<<grid, 256, 0, (cudaStream_t)stream_>>>((const float*)rows, (const long long*)offsets, B, C,
                                                                             (bf16*)out, Tp, front_pad);
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pb200_f32_to_bf16(const float* in, int64_t n, void* out, int32_t accumulate, void* stream_) {
    if (n == 0) return 0;
    f32_to_bf16_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream_>>>(in, n, (bf16*)out, accumulate);
    PB_LAUNCH_CHECK();
    return 0;
}
