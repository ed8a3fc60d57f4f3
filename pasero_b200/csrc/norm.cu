// pasero_b200 — fused dropout + residual + LayerNorm (forward / backward), bf16 I/O, fp32 math.
//
//   sum = residual + dropout(x)            (bf16, saved for backward — same rounding as the reference)
//   y   = LayerNorm(sum) * gamma + beta    (optional)
//
// Replaces the three ATen kernels per sub-block of the reference's `residual + self.dropout(x)` followed
// by `nn.LayerNorm` (pasero/models/transformer.py:1043-1054, 1073-1086, 1322-1339). Dropout masks come
// from Philox(seed, offset, element/8) and are regenerated in backward.
#include "common.cuh"
#include "../../include/pasero_b200.h"

namespace pb200 {

static constexpr int LN_WARPS = 4;

template <int NCH>  // NCH = ceil(D / 256): chunks of 8 elements per lane
__global__ void __launch_bounds__(LN_WARPS * 32)
dropout_add_ln_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ residual,
                          const bf16* __restrict__ gamma, const bf16* __restrict__ beta, int M, int D, float eps,
                          float p, uint64_t seed, uint64_t offset, const long long* __restrict__ offset_base,
                          bf16* __restrict__ sum_out, bf16* __restrict__ y_out, float* __restrict__ mean_out,
                          float* __restrict__ rstd_out, uint8_t* __restrict__ keep_out) {
    pdl_trigger();
    pdl_wait();
    if (offset_base) offset += (uint64_t)(*offset_base) << 12;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const Philox rng(seed);
    const uint32_t thresh = dropout_thresh16(p);
    const float keep_scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    float v[NCH][8];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = c * 256 + lane * 8;
        if (col < D) {
            const long long e = (long long)row * D + col;
            const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x + e));
            const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
            uint32_t keep = 0xffu;
            if (p > 0.f) {
                keep = dropout_keep8(rng, offset, (uint64_t)e >> 3, thresh);
                if (keep_out) keep_out[e >> 3] = (uint8_t)keep;   // one byte per 8 elements: backward skips the Philox rounds
            }
            uint32_t rw[4] = {0, 0, 0, 0};
            if (residual) {
                const uint4 rv = __ldg(reinterpret_cast<const uint4*>(residual + e));
                rw[0] = rv.x; rw[1] = rv.y; rw[2] = rv.z; rw[3] = rv.w;
            }
            uint32_t ow[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 a = unpack_bf16x2(xw[j]);
                if (p > 0.f) {  // dropout output is a bf16 tensor in the reference
                    a.x = ((keep >> (2 * j)) & 1) ? round_bf16(a.x * keep_scale) : 0.f;
                    a.y = ((keep >> (2 * j + 1)) & 1) ? round_bf16(a.y * keep_scale) : 0.f;
                }
                const float2 r = unpack_bf16x2(rw[j]);
                ow[j] = pack_bf16x2(a.x + r.x, a.y + r.y);
                const float2 o = unpack_bf16x2(ow[j]);
                v[c][2 * j] = o.x; v[c][2 * j + 1] = o.y;
                s += o.x + o.y;
            }
            if (sum_out) *reinterpret_cast<uint4*>(sum_out + e) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[c][j] = 0.f;
        }
    }
    if (!gamma) return;
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = c * 256 + lane * 8;
        if (col < D) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = v[c][j] - mean; q += d * d; }
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
    if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = c * 256 + lane * 8;
        if (col < D) {
            const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gamma + col));
            const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
            uint32_t bw[4] = {0, 0, 0, 0};
            if (beta) {
                const uint4 bv = __ldg(reinterpret_cast<const uint4*>(beta + col));
                bw[0] = bv.x; bw[1] = bv.y; bw[2] = bv.z; bw[3] = bv.w;
            }
            uint32_t ow[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 g = unpack_bf16x2(gw[j]);
                const float2 b = unpack_bf16x2(bw[j]);
                ow[j] = pack_bf16x2((v[c][2 * j] - mean) * rstd * g.x + b.x, (v[c][2 * j + 1] - mean) * rstd * g.y + b.y);
            }
            *reinterpret_cast<uint4*>(y_out + (long long)row * D + col) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
    }
}

// backward: dsum = LN'(dy) (or dy when no LN) [+ dsum_extra]; dx = dropout'(dsum).
// Column sums needed by the parameter gradients are accumulated on the way: dgamma = sum dy * xhat, dbeta = sum dy and
// (optionally) dlin = sum dx, the bias gradient of the Linear whose output fed this block (it would otherwise cost a
// separate pass over dx). One partial row [3][D] per CTA -> partial[blockIdx], reduced by reduce_partials_multi.
// One warp per row, 8 warps per CTA, the next row's operands are requested before the current row is processed so
// that enough bytes are in flight per SM to cover HBM latency.
static constexpr int LNB_WARPS = 8;

template <int NCH>
__global__ void __launch_bounds__(LNB_WARPS * 32, NCH <= 2 ? 2 : 1)
dropout_add_ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ dsum_extra,
                          const bf16* __restrict__ sum, const bf16* __restrict__ gamma,
                          const float* __restrict__ mean, const float* __restrict__ rstd, int M, int D, float p,
                          uint64_t seed, uint64_t offset, const long long* __restrict__ offset_base,
                          bf16* __restrict__ dresidual, bf16* __restrict__ dx, float* __restrict__ partial, int want_lin,
                          const uint8_t* __restrict__ keep_in) {
    pdl_trigger();
    pdl_wait();
    if (offset_base) offset += (uint64_t)(*offset_base) << 12;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Philox rng(seed);
    const uint32_t thresh = dropout_thresh16(p);
    const float keep_scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    float dg[NCH][8], db[NCH][8], dl[NCH][8];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) { dg[c][j] = 0.f; db[c][j] = 0.f; dl[c][j] = 0.f; }
    float gm[NCH][8];   // gamma of this lane's columns
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int col = c * 256 + lane * 8;
        uint4 gv = make_uint4(ELEM_ONE_X2, ELEM_ONE_X2, ELEM_ONE_X2, ELEM_ONE_X2);
        if (gamma && col < D) gv = __ldg(reinterpret_cast<const uint4*>(gamma + col));
        const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 g2 = unpack_bf16x2(gw[j]); gm[c][2 * j] = g2.x; gm[c][2 * j + 1] = g2.y; }
    }

    const int stride = gridDim.x * LNB_WARPS;
    int row = blockIdx.x * LNB_WARPS + warp;
    uint4 nd[NCH], ns[NCH];
    uint32_t nk[NCH];   // stored dropout keep bits of the next row (when keep_in is given)
    float nmu = 0.f, nrs = 1.f;
    auto fetch = [&](int r) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int col = c * 256 + lane * 8;
            nd[c] = make_uint4(0, 0, 0, 0); ns[c] = make_uint4(0, 0, 0, 0); nk[c] = 0xffu;
            if (r < M && col < D) {
                const long long e = (long long)r * D + col;
                nd[c] = __ldg(reinterpret_cast<const uint4*>(dy + e));
                if (gamma) ns[c] = __ldg(reinterpret_cast<const uint4*>(sum + e));
                if (keep_in && p > 0.f) nk[c] = (uint32_t)__ldg(keep_in + (e >> 3));
            }
        }
        if (gamma && r < M) { nmu = mean[r]; nrs = rstd[r]; }
    };
    fetch(row);
    for (; row < M; row += stride) {
        uint4 cd[NCH], cs[NCH];
        uint32_t ck[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) { cd[c] = nd[c]; cs[c] = ns[c]; ck[c] = nk[c]; }
        const float mu = nmu, rs = nrs;
        fetch(row + stride);
        float s1 = 0.f, s2 = 0.f;
        if (gamma) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const uint32_t dw[4] = {cd[c].x, cd[c].y, cd[c].z, cd[c].w};
                const uint32_t sw[4] = {cs[c].x, cs[c].y, cs[c].z, cs[c].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 d = unpack_bf16x2(dw[j]);
                    const float2 sx = unpack_bf16x2(sw[j]);
                    const float x0 = (sx.x - mu) * rs, x1 = (sx.y - mu) * rs;
                    dg[c][2 * j] += d.x * x0; dg[c][2 * j + 1] += d.y * x1;
                    db[c][2 * j] += d.x; db[c][2 * j + 1] += d.y;
                    const float g0 = d.x * gm[c][2 * j], g1 = d.y * gm[c][2 * j + 1];
                    s1 += g0 + g1;
                    s2 += g0 * x0 + g1 * x1;
                }
            }
            s1 = warp_sum(s1) / D; s2 = warp_sum(s2) / D;
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int col = c * 256 + lane * 8;
            if (col < D) {
                const long long e = (long long)row * D + col;
                const uint32_t dw[4] = {cd[c].x, cd[c].y, cd[c].z, cd[c].w};
                const uint32_t sw[4] = {cs[c].x, cs[c].y, cs[c].z, cs[c].w};
                float ds[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 d = unpack_bf16x2(dw[j]);
                    if (gamma) {
                        const float2 sx = unpack_bf16x2(sw[j]);
                        const float x0 = (sx.x - mu) * rs, x1 = (sx.y - mu) * rs;
                        ds[2 * j] = rs * (d.x * gm[c][2 * j] - s1 - x0 * s2);
                        ds[2 * j + 1] = rs * (d.y * gm[c][2 * j + 1] - s1 - x1 * s2);
                    } else {
                        ds[2 * j] = d.x; ds[2 * j + 1] = d.y;
                    }
                }
                if (dsum_extra) {
                    const uint4 ev = __ldg(reinterpret_cast<const uint4*>(dsum_extra + e));
                    const uint32_t ew[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 x2 = unpack_bf16x2(ew[j]);
                        ds[2 * j] += x2.x; ds[2 * j + 1] += x2.y;
                    }
                }
                uint32_t ow[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) ow[j] = pack_bf16x2(ds[2 * j], ds[2 * j + 1]);
                if (dresidual) *reinterpret_cast<uint4*>(dresidual + e) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                if (p > 0.f) {
                    const uint32_t keep = keep_in ? ck[c] : dropout_keep8(rng, offset, (uint64_t)e >> 3, thresh);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 r = unpack_bf16x2(ow[j]);  // gradient reaching dropout is bf16
                        const float a = ((keep >> (2 * j)) & 1) ? r.x * keep_scale : 0.f;
                        const float b = ((keep >> (2 * j + 1)) & 1) ? r.y * keep_scale : 0.f;
                        ow[j] = pack_bf16x2(a, b);
                    }
                }
                if (dx && (dx != dresidual || p > 0.f)) *reinterpret_cast<uint4*>(dx + e) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                if (want_lin) {   // column sums of dx exactly as stored (bf16)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 r = unpack_bf16x2(ow[j]);
                        dl[c][2 * j] += r.x; dl[c][2 * j + 1] += r.y;
                    }
                }
            }
        }
    }
    if (!partial) return;
    // CTA-level sum: every warp parks its column sums in shared memory, then each thread adds the 8 warps' values of
    // its columns in warp order (deterministic) and writes the CTA's partial row
    extern __shared__ float red[];   // [LNB_WARPS][3][NCH * 256]
    constexpr int W = NCH * 256;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = c * 256 + lane * 8 + j;
            red[(warp * 3 + 0) * W + i] = dg[c][j];
            red[(warp * 3 + 1) * W + i] = db[c][j];
            red[(warp * 3 + 2) * W + i] = dl[c][j];
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) {
        const int which = i / D, col = i % D;
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < LNB_WARPS; ++w) a += red[(w * 3 + which) * W + col];
        partial[((long long)blockIdx.x * 3 + which) * D + col] = a;
    }
}

// The shape the training step launches 42 times: LayerNorm present, D == NCH * 256 (every lane's chunks are in range), the
// dropout keep bits stored by forward (or no dropout), no extra gradient operand, all three column sums wanted. Same
// arithmetic and the same partial layout as the general kernel above; what it sheds is instruction issue, which is what
// bounds the general kernel (ncu: issue slots 69 % busy at 16 warps per SM, DRAM 21 %): operand options are compile-time, the
// element offset is carried instead of recomputed per chunk, packed pairs are split by bit operations, the normalised input
// is one FFMA, 1 / D is a constant and the row statistics are folded into the two gradient FFMAs: 376 SASS instructions per
// row of 512 against 736 (cuobjdump -sass; an alternating pair of row buffers, which would also save the 20 register copies
// per row, does not fit the 128-register cap of two CTAs per SM without spilling). ncu, 16 000 x 512: 19.9 -> 13.4 us
// (profiles/r02_ncu_ln_bwd_fast.txt); staging the rows three ahead through cp.async changed nothing and was dropped.
template <int NCH>
struct LnbRow {
    uint4 d[NCH], s[NCH];
    uint32_t k[NCH];
    float mu, rs;
};

// a packed pair as two fp32 values by bit operations on the word (no 16-bit sub-register traffic)
__device__ __forceinline__ float2 unpack2(uint32_t w) { return make_float2(elem_lo_f(w), elem_hi_f(w)); }

template <int NCH, bool DROP, bool LIN>
__global__ void __launch_bounds__(LNB_WARPS * 32, NCH <= 2 ? 2 : 1)
ln_bwd_fast_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ sum, const bf16* __restrict__ gamma,
                   const float* __restrict__ mean, const float* __restrict__ rstd, int M, float keep_scale,
                   bf16* __restrict__ dresidual, bf16* __restrict__ dx, float* __restrict__ partial,
                   const uint8_t* __restrict__ keep_in) {
    constexpr int D = NCH * 256;
    constexpr float inv_D = 1.0f / (float)D;
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = gridDim.x * LNB_WARPS;
    const long long step = (long long)stride * D;
    int row = blockIdx.x * LNB_WARPS + warp;
    long long e = (long long)row * D + lane * 8;      // element offset of this lane's first chunk in the current row

    auto load = [&](LnbRow<NCH>& R, int r, long long er) {
        if (r < M) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                R.d[c] = __ldg(reinterpret_cast<const uint4*>(dy + er) + c * 32);
                R.s[c] = __ldg(reinterpret_cast<const uint4*>(sum + er) + c * 32);
                if (DROP) R.k[c] = (uint32_t)__ldg(keep_in + (er >> 3) + c * 32);
            }
            R.mu = __ldg(mean + r); R.rs = __ldg(rstd + r);
        }
    };
    // the first row is requested before gamma is staged (a warp handles ~7 rows of a 16 000-row launch: a second exposed
    // round trip to HBM at the start is a tenth of the kernel)
    LnbRow<NCH> N;
    load(N, row, e);
    // gamma as fp32 in shared memory, laid out so that a lane's four columns are one conflict-free 128-bit load: 16 more
    // registers per lane would not fit beside the 48 column accumulators and the two row buffers under the 128-register cap
    __shared__ float4 gs4[NCH * 2 * 32];
    if (threadIdx.x < NCH * 64) {
        const int q = threadIdx.x >> 5, l = threadIdx.x & 31;
        const uint2 gv = __ldg(reinterpret_cast<const uint2*>(gamma + (q >> 1) * 256 + l * 8 + (q & 1) * 4));
        const float2 g01 = unpack2(gv.x), g23 = unpack2(gv.y);
        gs4[threadIdx.x] = make_float4(g01.x, g01.y, g23.x, g23.y);
    }
    __syncthreads();
    const uint32_t gs_lane = (uint32_t)__cvta_generic_to_shared(gs4 + lane);
    auto gamma4 = [&](int q) {      // asm volatile: re-read where it is used instead of being hoisted into (spilled) registers
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(gs_lane + q * 512));
        return v;
    };
    float dg[NCH][8], db[NCH][8], dl[NCH][8];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) { dg[c][j] = 0.f; db[c][j] = 0.f; dl[c][j] = 0.f; }
    auto process = [&](const LnbRow<NCH>& R, long long er) {
        const float rs = R.rs, nmr = -R.mu * rs;
        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const uint32_t dw[4] = {R.d[c].x, R.d[c].y, R.d[c].z, R.d[c].w};
            const uint32_t sw[4] = {R.s[c].x, R.s[c].y, R.s[c].z, R.s[c].w};
            const float4 ga = gamma4(2 * c), gb = gamma4(2 * c + 1);
            const float gm[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 d = unpack2(dw[j]);
                const float2 sx = unpack2(sw[j]);
                const float x0 = fmaf(sx.x, rs, nmr), x1 = fmaf(sx.y, rs, nmr);
                dg[c][2 * j] = fmaf(d.x, x0, dg[c][2 * j]); dg[c][2 * j + 1] = fmaf(d.y, x1, dg[c][2 * j + 1]);
                db[c][2 * j] += d.x; db[c][2 * j + 1] += d.y;
                const float g0 = d.x * gm[2 * j], g1 = d.y * gm[2 * j + 1];
                s1a += g0; s1b += g1;
                s2a = fmaf(g0, x0, s2a); s2b = fmaf(g1, x1, s2b);
            }
        }
        const float a = warp_sum(s1a + s1b) * (inv_D * rs);     // rs * mean(g)
        const float b = warp_sum(s2a + s2b) * (inv_D * rs);     // rs * mean(g * xhat)
        const float nb = -b * rs, c0 = -(a + nmr * b);          // -xhat * b - a = s * nb + c0
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const uint32_t dw[4] = {R.d[c].x, R.d[c].y, R.d[c].z, R.d[c].w};
            const uint32_t sw[4] = {R.s[c].x, R.s[c].y, R.s[c].z, R.s[c].w};
            const float4 ga = gamma4(2 * c), gb = gamma4(2 * c + 1);
            const float gm[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
            uint32_t ow[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {    // rs * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma
                const float2 d = unpack2(dw[j]);
                const float2 sx = unpack2(sw[j]);
                ow[j] = pack_bf16x2(fmaf(d.x * gm[2 * j], rs, fmaf(sx.x, nb, c0)),
                                    fmaf(d.y * gm[2 * j + 1], rs, fmaf(sx.y, nb, c0)));
            }
            if (DROP) {
                *(reinterpret_cast<uint4*>(dresidual + er) + c * 32) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                const uint32_t keep = R.k[c];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 r = unpack2(ow[j]);      // the gradient reaching dropout is 16-bit
                    ow[j] = pack_bf16x2(((keep >> (2 * j)) & 1) ? r.x * keep_scale : 0.f,
                                        ((keep >> (2 * j + 1)) & 1) ? r.y * keep_scale : 0.f);
                }
            }
            *(reinterpret_cast<uint4*>(dx + er) + c * 32) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            if (LIN) {                                          // column sums of dx exactly as stored
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 r = unpack2(ow[j]);
                    dl[c][2 * j] += r.x; dl[c][2 * j + 1] += r.y;
                }
            }
        }
    };

    for (; row < M; row += stride, e += step) {
        const LnbRow<NCH> C = N;
        load(N, row + stride, e + step);
        process(C, e);
    }
    // CTA-level sum, as in the general kernel (128-bit shared stores)
    extern __shared__ float red[];   // [LNB_WARPS][3][D]
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        float4* r0 = reinterpret_cast<float4*>(red + (warp * 3 + 0) * D + c * 256 + lane * 8);
        float4* r1 = reinterpret_cast<float4*>(red + (warp * 3 + 1) * D + c * 256 + lane * 8);
        float4* r2 = reinterpret_cast<float4*>(red + (warp * 3 + 2) * D + c * 256 + lane * 8);
        r0[0] = make_float4(dg[c][0], dg[c][1], dg[c][2], dg[c][3]); r0[1] = make_float4(dg[c][4], dg[c][5], dg[c][6], dg[c][7]);
        r1[0] = make_float4(db[c][0], db[c][1], db[c][2], db[c][3]); r1[1] = make_float4(db[c][4], db[c][5], db[c][6], db[c][7]);
        r2[0] = make_float4(dl[c][0], dl[c][1], dl[c][2], dl[c][3]); r2[1] = make_float4(dl[c][4], dl[c][5], dl[c][6], dl[c][7]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < LNB_WARPS; ++w) acc += red[w * 3 * D + i];
        partial[(long long)blockIdx.x * 3 * D + i] = acc;
    }
}

// out[n] (bf16 or fp32) = (accumulate ? out[n] : 0) + sum_p partial[p * stride + n]
// block = 32 columns x 8 partial-row lanes: coalesced 128-byte reads, tree-free smem reduce over the 8 lanes.
__device__ __forceinline__ float out_load(const bf16* p) { return bf16_to_f(*p); }
__device__ __forceinline__ float out_load(const float* p) { return *p; }
__device__ __forceinline__ void out_store(bf16* p, float v) { *p = f_to_bf16(v); }
__device__ __forceinline__ void out_store(float* p, float v) { *p = v; }
template <typename OutT>
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partial, int P, long long stride, int N,
                       OutT* __restrict__ out, int accumulate) {
    pdl_trigger();
    pdl_wait();
    const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + cx;
    float a = 0.f;
    if (n < N) {
        int p = py;
        for (; p + 24 < P; p += 32) {   // 4 independent loads in flight
            const float v0 = partial[(long long)p * stride + n], v1 = partial[(long long)(p + 8) * stride + n];
            const float v2 = partial[(long long)(p + 16) * stride + n], v3 = partial[(long long)(p + 24) * stride + n];
            a += (v0 + v1) + (v2 + v3);
        }
        for (; p < P; p += 8) a += partial[(long long)p * stride + n];
    }
    __shared__ float red[8][33];
    red[py][cx] = a;
    __syncthreads();
    if (py == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += red[r][cx];
        if (accumulate) s += out_load(out + n);
        out_store(out + n, s);
    }
}

// up to 3 vectors in one launch (blockIdx.y selects): out_k[n] = (acc_k ? out_k[n] : 0) + sum_p partial[p*stride + k*N + n]
struct ReduceOut3 { bf16* out[3]; int accumulate[3]; };
__global__ void __launch_bounds__(256)
reduce_partials_multi_kernel(const float* __restrict__ partial, int P, long long stride, int N, ReduceOut3 o) {
    pdl_trigger();
    pdl_wait();
    bf16* out = o.out[blockIdx.y];
    if (!out) return;
    const float* src = partial + (long long)blockIdx.y * N;
    const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + cx;
    float a = 0.f;
    if (n < N) {
        int q = py;
        for (; q + 24 < P; q += 32) {
            const float v0 = src[(long long)q * stride + n], v1 = src[(long long)(q + 8) * stride + n];
            const float v2 = src[(long long)(q + 16) * stride + n], v3 = src[(long long)(q + 24) * stride + n];
            a += (v0 + v1) + (v2 + v3);
        }
        for (; q < P; q += 8) a += src[(long long)q * stride + n];
    }
    __shared__ float red[8][33];
    red[py][cx] = a;
    __syncthreads();
    if (py == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += red[r][cx];
        if (o.accumulate[blockIdx.y]) s += bf16_to_f(out[n]);
        out[n] = f_to_bf16(s);
    }
}

// column sums of a bf16 [M, N] matrix (bias gradients): partial[split][N]
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16* __restrict__ x, int M, int N, long long ld, float* __restrict__ partial) {
    pdl_trigger();
    pdl_wait();
    const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;   // 32 column groups of 8, 8 row lanes
    const int col = blockIdx.x * 256 + cg * 8;
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (col < N) {
        for (int r = blockIdx.y * 8 + rl; r < M; r += gridDim.y * 8) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (long long)r * ld + col));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(w[j]);
                a[2 * j] += f.x; a[2 * j + 1] += f.y;
            }
        }
    }
    __shared__ float red[8][256];
#pragma unroll
    for (int j = 0; j < 8; ++j) red[rl][cg * 8 + j] = a[j];
    __syncthreads();
    const int c = threadIdx.x;
    if (blockIdx.x * 256 + c < N) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += red[r][c];
        partial[(long long)blockIdx.y * N + blockIdx.x * 256 + c] = s;
    }
}

static int ln_bwd_grid(int M) {
    int g = (M + LNB_WARPS - 1) / LNB_WARPS;
    const int cap = 2 * num_sms();
    return g < cap ? g : cap;
}

static int colsum_splits(long long M, long long N) {
    const long long col_blocks = (N + 255) / 256;
    long long splits = (4LL * num_sms() + col_blocks - 1) / col_blocks;
    const long long max_splits = (M + 15) / 16;
    if (splits > max_splits) splits = max_splits;
    if (splits > 1024) splits = 1024;
    if (splits < 1) splits = 1;
    return (int)splits;
}

}  // namespace pb200

using namespace pb200;

extern "C" size_t pb200_ln_bwd_workspace_bytes(int64_t M, int64_t D) {
    return (size_t)ln_bwd_grid((int)M) * 3 * (size_t)D * sizeof(float);
}

extern "C" int pb200_dropout_add_ln_fwd(const void* x, const void* residual, const void* gamma, const void* beta,
                                        int64_t M, int64_t D, float eps, float p, uint64_t seed, uint64_t offset,
                                        const int64_t* offset_base, void* sum_out, void* y_out, float* mean_out,
                                        float* rstd_out, void* keep_out, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (M == 0) return 0;
    PB_CHECK(D % 8 == 0 && D <= 1024 && D > 0, "LayerNorm width %lld unsupported (multiple of 8, <= 1024)", (long long)D);
    PB_CHECK(gamma == nullptr || y_out != nullptr, "y_out required with LayerNorm");
    PB_CHECK(gamma != nullptr || sum_out != nullptr, "sum_out required without LayerNorm");
    PB_CHECK(p >= 0.f && p < 1.f, "dropout p out of range");
    const int nch = (int)((D + 255) / 256);
    const int grid = (int)((M + LN_WARPS - 1) / LN_WARPS);
#define PB_LN_FWD(N_)                                                                                         \
    PB_CUDA(launch_pdl(dropout_add_ln_fwd_kernel<N_>, dim3(grid), dim3(LN_WARPS * 32), 0, stream,                \
        (const bf16*)x, (const bf16*)residual, (const bf16*)gamma, (const bf16*)beta, (int)M, (int)D, eps, p,  \
        seed, offset, (const long long*)offset_base, (bf16*)sum_out, (bf16*)y_out, mean_out, rstd_out,         \
        (uint8_t*)keep_out))
    switch (nch) { case 1: PB_LN_FWD(1); break; case 2: PB_LN_FWD(2); break; case 3: PB_LN_FWD(3); break; default: PB_LN_FWD(4); }
#undef PB_LN_FWD
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pb200_dropout_add_ln_bwd(const void* dy, const void* dsum_extra, const void* sum, const void* gamma,
                                        const float* mean, const float* rstd, int64_t M, int64_t D, float p,
                                        uint64_t seed, uint64_t offset, const int64_t* offset_base, void* dresidual, void* dx,
                                        void* dgamma, void* dbeta, void* dbias_lin, int32_t accumulate, const void* keep_mask,
                                        void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (M == 0) return 0;
    PB_CHECK(D % 8 == 0 && D <= 1024 && D > 0, "LayerNorm width %lld unsupported", (long long)D);
    const int grid = ln_bwd_grid((int)M);
    float* partial = nullptr;
    if ((gamma && (dgamma || dbeta)) || dbias_lin) {
        PB_CHECK(workspace && workspace_bytes >= pb200_ln_bwd_workspace_bytes(M, D), "LN backward workspace too small");
        partial = reinterpret_cast<float*>(workspace);
    }
    const int nch = (int)((D + 255) / 256);
    // the training step's shape: see ln_bwd_fast_kernel
    const bool drop = p > 0.f;
    if (gamma && sum && mean && rstd && !dsum_extra && partial && dx && D == 256LL * nch && nch >= 2 &&
        (drop ? (keep_mask && dresidual && dresidual != dx) : (!dresidual || dresidual == dx))) {
        const bool lin = dbias_lin != nullptr;
        const float keep_scale = drop ? 1.0f / (1.0f - p) : 1.0f;
#define PB_LN_FAST(N_, DROP_, LIN_)                                                                                    \
        do {                                                                                                           \
            PB_CUDA(cudaFuncSetAttribute(ln_bwd_fast_kernel<N_, DROP_, LIN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         LNB_WARPS * 3 * N_ * 256 * 4));                                               \
            PB_CUDA(launch_pdl(ln_bwd_fast_kernel<N_, DROP_, LIN_>, dim3(grid), dim3(LNB_WARPS * 32),                  \
                               (size_t)(LNB_WARPS * 3 * N_ * 256 * 4), stream, (const bf16*)dy, (const bf16*)sum,      \
                               (const bf16*)gamma, mean, rstd, (int)M, keep_scale, (bf16*)dresidual, (bf16*)dx, partial, \
                               (const uint8_t*)keep_mask));                                                            \
        } while (0)
#define PB_LN_FAST2(N_)                                                                                                \
        do {                                                                                                           \
            if (drop) { if (lin) PB_LN_FAST(N_, true, true); else PB_LN_FAST(N_, true, false); }                       \
            else { if (lin) PB_LN_FAST(N_, false, true); else PB_LN_FAST(N_, false, false); }                          \
        } while (0)
        switch (nch) { case 2: PB_LN_FAST2(2); break; case 3: PB_LN_FAST2(3); break; default: PB_LN_FAST2(4); }
#undef PB_LN_FAST2
#undef PB_LN_FAST
        PB_LAUNCH_CHECK();
        if (!(accumulate & 4))
            return pb200_ln_bwd_reduce(workspace, M, D, dgamma, dbeta, dbias_lin, accumulate, stream_);
        return 0;
    }
#define PB_LN_BWD(N_)                                                                                        \
    PB_CUDA(cudaFuncSetAttribute(dropout_add_ln_bwd_kernel<N_>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                 LNB_WARPS * 3 * N_ * 256 * 4));                                                \
    PB_CUDA(launch_pdl(dropout_add_ln_bwd_kernel<N_>, dim3(grid), dim3(LNB_WARPS * 32),                         \
                       (size_t)(partial ? LNB_WARPS * 3 * N_ * 256 * 4 : 0), stream,                              \
        (const bf16*)dy, (const bf16*)dsum_extra, (const bf16*)sum, (const bf16*)gamma, mean, rstd, (int)M,   \
        (int)D, p, seed, offset, (const long long*)offset_base, (bf16*)dresidual, (bf16*)dx, partial,          \
        dbias_lin != nullptr ? 1 : 0, (const uint8_t*)keep_mask))
    switch (nch) { case 1: PB_LN_BWD(1); break; case 2: PB_LN_BWD(2); break; case 3: PB_LN_BWD(3); break; default: PB_LN_BWD(4); }
#undef PB_LN_BWD
    PB_LAUNCH_CHECK();
    if (partial && !(accumulate & 4))
        return pb200_ln_bwd_reduce(workspace, M, D, gamma ? dgamma : nullptr, gamma ? dbeta : nullptr, dbias_lin, accumulate, stream_);
    return 0;
}

extern "C" int pb200_ln_bwd_reduce(const void* workspace, int64_t M, int64_t D, void* dgamma, void* dbeta, void* dbias_lin,
                                   int32_t accumulate, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (M == 0) return 0;
    PB_CHECK(workspace != nullptr, "LN backward reduce: null workspace");
    ReduceOut3 o;
    o.out[0] = (bf16*)dgamma; o.out[1] = (bf16*)dbeta; o.out[2] = (bf16*)dbias_lin;
    o.accumulate[0] = o.accumulate[1] = accumulate & 1; o.accumulate[2] = (accumulate >> 1) & 1;
    dim3 rgrid((unsigned)((D + 31) / 32), 3);
    PB_CUDA(launch_pdl(reduce_partials_multi_kernel, rgrid, dim3(256), 0, stream, (const float*)workspace, ln_bwd_grid((int)M),
                       (long long)(3 * D), (int)D, o));
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t pb200_colsum_workspace_bytes(int64_t M, int64_t N) {
    return (size_t)colsum_splits(M, N) * (size_t)N * sizeof(float);
}

static int colsum_impl(const void* x, int64_t M, int64_t N, int64_t ld, void* out, bool out_fp32, int32_t accumulate,
                       void* workspace, size_t workspace_bytes, void* stream_);

extern "C" int pb200_colsum_bf16(const void* x, int64_t M, int64_t N, int64_t ld, void* out, int32_t accumulate,
                                 void* workspace, size_t workspace_bytes, void* stream_) {
    return colsum_impl(x, M, N, ld, out, false, accumulate, workspace, workspace_bytes, stream_);
}

extern "C" int pb200_colsum_f32(const void* x, int64_t M, int64_t N, int64_t ld, float* out, int32_t accumulate,
                                void* workspace, size_t workspace_bytes, void* stream_) {
    return colsum_impl(x, M, N, ld, out, true, accumulate, workspace, workspace_bytes, stream_);
}

static int colsum_impl(const void* x, int64_t M, int64_t N, int64_t ld, void* out, bool out_fp32, int32_t accumulate,
                       void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PB_CHECK(N % 8 == 0 && ld % 8 == 0, "colsum: N and ld must be multiples of 8");
    if (N == 0) return 0;
    const long long splits = colsum_splits(M, N);
    PB_CHECK(workspace && workspace_bytes >= pb200_colsum_workspace_bytes(M, N), "colsum workspace too small");
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)splits);
    PB_CUDA(launch_pdl(colsum_kernel, grid, dim3(256), 0, stream, (const bf16*)x, (int)M, (int)N, (long long)ld, (float*)workspace));
    PB_LAUNCH_CHECK();
    if (out_fp32)
        PB_CUDA(launch_pdl(reduce_partials_kernel<float>, dim3((unsigned)((N + 31) / 32)), dim3(256), 0, stream, (const float*)workspace,
                           (int)splits, (long long)N, (int)N, (float*)out, (int)accumulate));
    else
        PB_CUDA(launch_pdl(reduce_partials_kernel<bf16>, dim3((unsigned)((N + 31) / 32)), dim3(256), 0, stream, (const float*)workspace,
                           (int)splits, (long long)N, (int)N, (bf16*)out, (int)accumulate));
    PB_LAUNCH_CHECK();
    return 0;
}
