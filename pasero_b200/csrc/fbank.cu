// pasero_b200 — fused Kaldi-compatible log-mel filterbank + utterance CMVN for 16 kHz PCM.
//
// One pass over the PCM: framing (25 ms / 10 ms, snip_edges), per-frame DC removal, pre-emphasis
// 0.97, Povey window, zero-pad 400 -> 512, real FFT (as a 256-point complex FFT of the even/odd
// packed signal, radix-4 Stockham in shared memory, one warp per frame), power spectrum, sparse
// 80-bin mel projection, log(max(., FLT_EPSILON)). Per-utterance, per-bin sum / sum-of-squares are
// accumulated in fp64 for the CMVN pass, which normalises in place (the 20 MB feature block of a
// 64-utterance batch is still L2 resident) and can also emit the zero-padded bf16 layout that the
// conv subsampler's TMA descriptors read.
//
// The reference has no fbank code (features are pre-extracted, pasero/files.py:103-192); the
// arithmetic follows torchaudio.compliance.kaldi.fbank (kaldi.py:514-646, _get_window :154-217,
// get_mel_banks :436-511) and the fairseq-S2T utterance CMVN convention, see oracle/fbank.py.
#include "common.cuh"
#include "../../include/pasero_b200.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>

namespace pb200 {

static constexpr int WIN = 400;
static constexpr int SHIFT = 160;
static constexpr int NFFT = 512;
static constexpr int NC = NFFT / 2;   // 256-point complex FFT
static constexpr int NMEL = 80;
static constexpr int FR_PER_CTA = 32;
static constexpr int FB_WARPS = 8;
static constexpr int FB_THREADS = FB_WARPS * 32;
static constexpr int SAMPLES_PER_CTA = (FR_PER_CTA - 1) * SHIFT + WIN;  // 5360
static constexpr int MAX_MEL_TAPS = 640;

struct FbankTables {
    float window[WIN];
    float2 w256[NC];        // exp(-2 pi i k / 256)
    float2 w512[NC + 1];    // exp(-2 pi i k / 512), k = 0..256
    int mel_start[NMEL];
    int mel_len[NMEL];
    int mel_off[NMEL];
    float mel_w[MAX_MEL_TAPS];
};

__device__ FbankTables g_fbank_tables;

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// radix-4 Stockham pass over 256 complex points held in shared memory; one warp, 2 butterflies/lane
template <int NS>
__device__ __forceinline__ void fft_pass(const float2* __restrict__ src, float2* __restrict__ dst,
                                         const float2* __restrict__ tw, int lane) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int j = lane + it * 32;          // 0..63
        const int k = j & (NS - 1);
        float2 v0 = src[j], v1 = src[j + 64], v2 = src[j + 128], v3 = src[j + 192];
        if (NS > 1) {
            const int t = k * (64 / NS);       // twiddle index into W_256
            v1 = cmul(v1, tw[t]);
            v2 = cmul(v2, tw[2 * t]);
            v3 = cmul(v3, tw[3 * t]);
        }
        // radix-4 butterfly (forward transform: multiply by -i)
        const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y);
        const float2 a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
        const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
        const float2 a3 = make_float2(v1.y - v3.y, v3.x - v1.x);  // -i * (v1 - v3)
        const int d = ((j - k) << 2) + k;
        dst[d] = make_float2(a0.x + a2.x, a0.y + a2.y);
        dst[d + NS] = make_float2(a1.x + a3.x, a1.y + a3.y);
        dst[d + 2 * NS] = make_float2(a0.x - a2.x, a0.y - a2.y);
        dst[d + 3 * NS] = make_float2(a1.x - a3.x, a1.y - a3.y);
    }
}

struct FbankSmem {
    alignas(16) float samples[SAMPLES_PER_CTA + 8];
    alignas(16) float out[FR_PER_CTA][NMEL];
    alignas(16) float2 buf[FB_WARPS][2][NC];
    alignas(16) float window[WIN];
    alignas(16) float2 w256[NC];
    alignas(16) float2 w512[NC + 1];
    alignas(16) float mel_w[MAX_MEL_TAPS];
    alignas(16) float power[FB_WARPS][NC + 8];
    int mel_start[NMEL], mel_len[NMEL], mel_off[NMEL];
};

__global__ void __launch_bounds__(FB_THREADS)
fbank_kernel(const float* __restrict__ pcm, const long long* __restrict__ sample_off,
             const long long* __restrict__ frame_off, float* __restrict__ feats, double* __restrict__ stats,
             float pcm_scale) {
    extern __shared__ __align__(16) uint8_t fb_smem_raw[];
    FbankSmem& s = *reinterpret_cast<FbankSmem*>(fb_smem_raw);
    const int b = blockIdx.y;
    const long long s0 = sample_off[b];
    const long long f_begin = frame_off[b];
    const int nframes = (int)(frame_off[b + 1] - f_begin);
    const int f0 = blockIdx.x * FR_PER_CTA;
    if (f0 >= nframes) return;
    const int nf = min(FR_PER_CTA, nframes - f0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- stage constants and this tile's samples (coalesced; 16-byte vectors when aligned) ----
    const FbankTables& T = g_fbank_tables;
    for (int i = tid; i < WIN; i += FB_THREADS) s.window[i] = T.window[i];
    for (int i = tid; i < NC; i += FB_THREADS) s.w256[i] = T.w256[i];
    for (int i = tid; i <= NC; i += FB_THREADS) s.w512[i] = T.w512[i];
    for (int i = tid; i < MAX_MEL_TAPS; i += FB_THREADS) s.mel_w[i] = T.mel_w[i];
    for (int i = tid; i < NMEL; i += FB_THREADS) {
        s.mel_start[i] = T.mel_start[i]; s.mel_len[i] = T.mel_len[i]; s.mel_off[i] = T.mel_off[i];
    }
    const int nsamp = (nf - 1) * SHIFT + WIN;
    const float* src = pcm + s0 + (long long)f0 * SHIFT;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int n4 = nsamp >> 2;
        for (int i = tid; i < n4; i += FB_THREADS) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
            reinterpret_cast<float4*>(s.samples)[i] = v;
        }
        for (int i = (n4 << 2) + tid; i < nsamp; i += FB_THREADS) s.samples[i] = __ldg(src + i);
    } else {
        for (int i = tid; i < nsamp; i += FB_THREADS) s.samples[i] = __ldg(src + i);
    }
    __syncthreads();

    // ---- one warp per frame ----
    for (int fr = warp; fr < nf; fr += FB_WARPS) {
        const float* x = s.samples + fr * SHIFT;
        float2* b0 = s.buf[warp][0];
        float2* b1 = s.buf[warp][1];
        // DC offset (mean over the 400-sample frame)
        float sum = 0.f;
        for (int i = lane; i < WIN; i += 32) sum += x[i] * pcm_scale;
        const float mean = warp_sum(sum) * (1.0f / WIN);
        // pre-emphasis + window, packed as z[n] = y[2n] + i y[2n+1]; zero-padded to 256 complex points
        for (int n = lane; n < NC; n += 32) {
            float2 z = make_float2(0.f, 0.f);
            if (n < WIN / 2) {
                const int j = 2 * n;
                const float xm1 = x[j > 0 ? j - 1 : 0] * pcm_scale - mean;
                const float x0 = x[j] * pcm_scale - mean;
                const float x1 = x[j + 1] * pcm_scale - mean;
                z.x = (x0 - 0.97f * xm1) * s.window[j];
                z.y = (x1 - 0.97f * x0) * s.window[j + 1];
            }
            b0[n] = z;
        }
        __syncwarp();
        fft_pass<1>(b0, b1, s.w256, lane);  __syncwarp();
        fft_pass<4>(b1, b0, s.w256, lane);  __syncwarp();
        fft_pass<16>(b0, b1, s.w256, lane); __syncwarp();
        fft_pass<64>(b1, b0, s.w256, lane); __syncwarp();
        // untangle the packed real transform: X[k] = E[k] + W512^k O[k], power = |X[k]|^2, k = 0..256
        for (int k = lane; k <= NC; k += 32) {
            const float2 zk = b0[k & (NC - 1)];
            const float2 zc = b0[(NC - k) & (NC - 1)];
            const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
            const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
            const float2 wo = cmul(s.w512[k], o);
            const float re = e.x + wo.x, im = e.y + wo.y;
            s.power[warp][k] = re * re + im * im;
        }
        __syncwarp();
        // sparse mel projection + log
        for (int m = lane; m < NMEL; m += 32) {
            const int st = s.mel_start[m], ln = s.mel_len[m], of = s.mel_off[m];
            float acc = 0.f;
            for (int i = 0; i < ln; ++i) acc = fmaf(s.power[warp][st + i], s.mel_w[of + i], acc);
            s.out[fr][m] = logf(fmaxf(acc, 1.1920928955078125e-07f));
        }
        __syncwarp();
    }
    __syncthreads();

    // ---- coalesced store of the [nf x 80] tile, and fp64 per-bin statistics for CMVN ----
    float* dst = feats + (f_begin + f0) * NMEL;
    const float* so = &s.out[0][0];
    const int total = nf * NMEL;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        for (int i = tid; i < (total >> 2); i += FB_THREADS)
            reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(so)[i];
    } else {
        for (int i = tid; i < total; i += FB_THREADS) dst[i] = so[i];
    }
    if (stats != nullptr && tid < NMEL) {
        double sm = 0.0, sq = 0.0;
        for (int fr = 0; fr < nf; ++fr) {
            const double v = (double)s.out[fr][tid];
            sm += v; sq += v * v;
        }
        atomicAdd(&stats[((long long)b * NMEL + tid) * 2 + 0], sm);
        atomicAdd(&stats[((long long)b * NMEL + tid) * 2 + 1], sq);
    }
}

// ---------------------------------------------------------------------------------------------
// v2: register-resident FFT. 16 lanes per frame (two frames per warp), 16 complex points per lane.
//   256-point complex FFT as 16 x 16 (Cooley-Tukey): lane n1 holds z[n1 + 16 n2] and runs a 16-point DFT over n2 in
//   registers, multiplies by W256^(n1 k2), the 16 x 16 values are transposed between the lanes through a padded
//   shared-memory tile (conflict free), lane k2 runs the second 16-point DFT over n1 and holds Z[16 k1 + k2].
//   Shared memory traffic per frame drops from ~250 to ~110 accesses per lane and the strided Stockham stores (8-way
//   bank conflicts in the first two radix-4 passes of v1) disappear.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// forward 4-point DFT (W4 = -i), in place
__device__ __forceinline__ void dft4(float2& u0, float2& u1, float2& u2, float2& u3) {
    const float2 a = cadd(u0, u2), b = csub(u0, u2), c = cadd(u1, u3), d = mul_mi(csub(u1, u3));
    u0 = cadd(a, c); u1 = cadd(b, d); u2 = csub(a, c); u3 = csub(b, d);
}

// forward 16-point DFT in registers, natural order in and out: n = 4 n_a + n_b, k = k_a + 4 k_b
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508978f, R = 0.70710678118654752f;
    float2 t[4][4];   // t[n_b][k_a]
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
        float2 u0 = v[nb], u1 = v[4 + nb], u2 = v[8 + nb], u3 = v[12 + nb];
        dft4(u0, u1, u2, u3);
        t[nb][0] = u0; t[nb][1] = u1; t[nb][2] = u2; t[nb][3] = u3;
    }
    // twiddles W16^(n_b k_a) = exp(-2 pi i n_b k_a / 16)
    t[1][1] = cmul(t[1][1], make_float2(C1, -S1));
    t[1][2] = cmul(t[1][2], make_float2(R, -R));
    t[1][3] = cmul(t[1][3], make_float2(S1, -C1));
    t[2][1] = cmul(t[2][1], make_float2(R, -R));
    t[2][2] = mul_mi(t[2][2]);
    t[2][3] = cmul(t[2][3], make_float2(-R, -R));
    t[3][1] = cmul(t[3][1], make_float2(S1, -C1));
    t[3][2] = cmul(t[3][2], make_float2(-R, -R));
    t[3][3] = cmul(t[3][3], make_float2(-C1, S1));
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) {
        float2 u0 = t[0][ka], u1 = t[1][ka], u2 = t[2][ka], u3 = t[3][ka];
        dft4(u0, u1, u2, u3);
        v[ka] = u0; v[ka + 4] = u1; v[ka + 8] = u2; v[ka + 12] = u3;
    }
}

static constexpr int FB2_TILE = 16 * 17;   // float2 per half-warp: 16 x 16 transpose tile padded to 17, reused for Z[256]

struct FbankSmem2 {
    alignas(16) float samples[SAMPLES_PER_CTA + 8];
    alignas(16) float out[FR_PER_CTA][NMEL];
    alignas(16) float2 tile[FB_WARPS][2][FB2_TILE];
    alignas(16) float power[FB_WARPS][2][NC + 8];
    alignas(16) float window[WIN];
    alignas(16) float2 w256[NC];
    alignas(16) float2 w512[NC + 1];
    alignas(16) float mel_w[MAX_MEL_TAPS];
    int mel_start[NMEL], mel_len[NMEL], mel_off[NMEL];
};

__global__ void __launch_bounds__(FB_THREADS)
fbank_kernel_v2(const float* __restrict__ pcm, const long long* __restrict__ sample_off,
                const long long* __restrict__ frame_off, float* __restrict__ feats, double* __restrict__ stats,
                float pcm_scale) {
    extern __shared__ __align__(16) uint8_t fb_smem_raw[];
    FbankSmem2& s = *reinterpret_cast<FbankSmem2*>(fb_smem_raw);
    const int b = blockIdx.y;
    const long long s0 = sample_off[b];
    const long long f_begin = frame_off[b];
    const int nframes = (int)(frame_off[b + 1] - f_begin);
    const int f0 = blockIdx.x * FR_PER_CTA;
    if (f0 >= nframes) return;
    const int nf = min(FR_PER_CTA, nframes - f0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int hl = lane & 15, half = lane >> 4;

    const FbankTables& T = g_fbank_tables;
    for (int i = tid; i < WIN; i += FB_THREADS) s.window[i] = T.window[i];
    for (int i = tid; i < NC; i += FB_THREADS) s.w256[i] = T.w256[i];
    for (int i = tid; i <= NC; i += FB_THREADS) s.w512[i] = T.w512[i];
    for (int i = tid; i < MAX_MEL_TAPS; i += FB_THREADS) s.mel_w[i] = T.mel_w[i];
    for (int i = tid; i < NMEL; i += FB_THREADS) {
        s.mel_start[i] = T.mel_start[i]; s.mel_len[i] = T.mel_len[i]; s.mel_off[i] = T.mel_off[i];
    }
    const int nsamp = (nf - 1) * SHIFT + WIN;
    const float* src = pcm + s0 + (long long)f0 * SHIFT;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int n4 = nsamp >> 2;
        for (int i = tid; i < n4; i += FB_THREADS) reinterpret_cast<float4*>(s.samples)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        for (int i = (n4 << 2) + tid; i < nsamp; i += FB_THREADS) s.samples[i] = __ldg(src + i);
    } else {
        for (int i = tid; i < nsamp; i += FB_THREADS) s.samples[i] = __ldg(src + i);
    }
    __syncthreads();

    float2* tile = s.tile[warp][half];
    float* pw = s.power[warp][half];
    // two frames per warp iteration; frames beyond nf are computed on frame nf - 1 and not stored (keeps the warp converged)
    for (int fr0 = warp * 2; fr0 < nf; fr0 += FB_WARPS * 2) {
        const int fr = fr0 + half;
        const bool live = fr < nf;
        const float* x = s.samples + (live ? fr : nf - 1) * SHIFT;
        // DC offset over the 400 samples of the frame
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < WIN / 16; ++i) sum += x[hl + 16 * i] * pcm_scale;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum * (1.0f / WIN);
        // z[n] = y[2n] + i y[2n+1], n = hl + 16 n2 (pre-emphasis 0.97, Povey window, zero-padded beyond 200)
        float2 v[16];
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) {
            const int n = hl + 16 * n2;
            float2 z = make_float2(0.f, 0.f);
            if (n < WIN / 2) {
                const int j = 2 * n;
                const float xm1 = x[j > 0 ? j - 1 : 0] * pcm_scale - mean;
                const float x0 = x[j] * pcm_scale - mean;
                const float x1 = x[j + 1] * pcm_scale - mean;
                z.x = (x0 - 0.97f * xm1) * s.window[j];
                z.y = (x1 - 0.97f * x0) * s.window[j + 1];
            }
            v[n2] = z;
        }
        dft16(v);                                   // A[n1 = hl][k2]
#pragma unroll
        for (int k2 = 1; k2 < 16; ++k2) v[k2] = cmul(v[k2], s.w256[(hl * k2) & (NC - 1)]);
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) tile[k2 * 17 + hl] = v[k2];   // transpose through the padded tile
        __syncwarp();
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) v[n1] = tile[hl * 17 + n1];   // lane k2 = hl: A'[n1][k2]
        __syncwarp();
        dft16(v);                                   // Z[16 k1 + hl] in v[k1]
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) tile[16 * k1 + hl] = v[k1];
        __syncwarp();
        // untangle the packed real transform: X[k] = E[k] + W512^k O[k], power = |X[k]|^2, k = 0..256
        auto untangle = [&](int k) {
            const float2 zk = tile[k & (NC - 1)];
            const float2 zc = tile[(NC - k) & (NC - 1)];
            const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
            const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
            const float2 wo = cmul(s.w512[k], o);
            const float re = e.x + wo.x, im = e.y + wo.y;
            pw[k] = re * re + im * im;
        };
#pragma unroll
        for (int i = 0; i < 16; ++i) untangle(hl + 16 * i);    // independent iterations: loads overlap
        if (hl == 0) untangle(NC);
        __syncwarp();
        // sparse mel projection + log: 80 bins over the 16 lanes
        for (int m = hl; m < NMEL; m += 16) {
            const int st = s.mel_start[m], ln = s.mel_len[m], of = s.mel_off[m];
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // four partial sums: the tap loop is a latency chain otherwise
            int i = 0;
            for (; i + 4 <= ln; i += 4) {
                a0 = fmaf(pw[st + i], s.mel_w[of + i], a0);
                a1 = fmaf(pw[st + i + 1], s.mel_w[of + i + 1], a1);
                a2 = fmaf(pw[st + i + 2], s.mel_w[of + i + 2], a2);
                a3 = fmaf(pw[st + i + 3], s.mel_w[of + i + 3], a3);
            }
            for (; i < ln; ++i) a0 = fmaf(pw[st + i], s.mel_w[of + i], a0);
            const float acc = (a0 + a1) + (a2 + a3);
            if (live) s.out[fr][m] = logf(fmaxf(acc, 1.1920928955078125e-07f));
        }
        __syncwarp();
    }
    __syncthreads();

    float* dst = feats + (f_begin + f0) * NMEL;
    const float* so = &s.out[0][0];
    const int total = nf * NMEL;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        for (int i = tid; i < (total >> 2); i += FB_THREADS)
            reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(so)[i];
    } else {
        for (int i = tid; i < total; i += FB_THREADS) dst[i] = so[i];
    }
    if (stats != nullptr && tid < NMEL) {
        double sm = 0.0, sq = 0.0;
        for (int fr = 0; fr < nf; ++fr) {
            const double val = (double)s.out[fr][tid];
            sm += val; sq += val * val;
        }
        atomicAdd(&stats[((long long)b * NMEL + tid) * 2 + 0], sm);
        atomicAdd(&stats[((long long)b * NMEL + tid) * 2 + 1], sq);
    }
}

// v3 (experimental, PB200_FBANK=v3): v2 with a leaner shared-memory footprint — the power spectrum reuses the transpose
// tile, the log-mel rows go straight to global memory and the CMVN statistics are pre-reduced in shared memory — so that
// three CTAs (24 warps) instead of two fit an SM: v2 is issue/latency bound at ~0.3 IPC per scheduler with 4 warps each.
struct FbankSmem3 {
    alignas(16) float samples[SAMPLES_PER_CTA + 8];
    double stat[NMEL][2];
    alignas(16) float2 tile[FB_WARPS][2][FB2_TILE];
    alignas(16) float window[WIN];
    alignas(16) float2 w256[NC];
    alignas(16) float2 w512[NC + 1];
    alignas(16) float mel_w[MAX_MEL_TAPS];
    int mel_start[NMEL], mel_len[NMEL], mel_off[NMEL];
};

__global__ void __launch_bounds__(FB_THREADS)
fbank_kernel_v3(const float* __restrict__ pcm, const long long* __restrict__ sample_off,
                const long long* __restrict__ frame_off, float* __restrict__ feats, double* __restrict__ stats,
                float pcm_scale) {
    extern __shared__ __align__(16) uint8_t fb_smem_raw[];
    FbankSmem3& s = *reinterpret_cast<FbankSmem3*>(fb_smem_raw);
    const int b = blockIdx.y;
    const long long s0 = sample_off[b];
    const long long f_begin = frame_off[b];
    const int nframes = (int)(frame_off[b + 1] - f_begin);
    const int f0 = blockIdx.x * FR_PER_CTA;
    if (f0 >= nframes) return;
    const int nf = min(FR_PER_CTA, nframes - f0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int hl = lane & 15, half = lane >> 4;

    const FbankTables& T = g_fbank_tables;
    for (int i = tid; i < WIN; i += FB_THREADS) s.window[i] = T.window[i];
    for (int i = tid; i < NC; i += FB_THREADS) s.w256[i] = T.w256[i];
    for (int i = tid; i <= NC; i += FB_THREADS) s.w512[i] = T.w512[i];
    for (int i = tid; i < MAX_MEL_TAPS; i += FB_THREADS) s.mel_w[i] = T.mel_w[i];
    for (int i = tid; i < NMEL; i += FB_THREADS) {
        s.mel_start[i] = T.mel_start[i]; s.mel_len[i] = T.mel_len[i]; s.mel_off[i] = T.mel_off[i];
        s.stat[i][0] = 0.0; s.stat[i][1] = 0.0;
    }
    const int nsamp = (nf - 1) * SHIFT + WIN;
    const float* src = pcm + s0 + (long long)f0 * SHIFT;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int n4 = nsamp >> 2;
        for (int i = tid; i < n4; i += FB_THREADS) reinterpret_cast<float4*>(s.samples)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        for (int i = (n4 << 2) + tid; i < nsamp; i += FB_THREADS) s.samples[i] = __ldg(src + i);
    } else {
        for (int i = tid; i < nsamp; i += FB_THREADS) s.samples[i] = __ldg(src + i);
    }
    __syncthreads();

    float2* tile = s.tile[warp][half];
    float* pw = reinterpret_cast<float*>(tile);   // the power spectrum (257 floats) overwrites the tile once Z has been read
    float* dst = feats + (f_begin + f0) * NMEL;
    float st_sum[NMEL / 16], st_sq[NMEL / 16];     // CMVN statistics of this lane's 5 mel bins over this half-warp's frames
#pragma unroll
    for (int i = 0; i < NMEL / 16; ++i) { st_sum[i] = 0.f; st_sq[i] = 0.f; }
    // two frames per warp iteration; frames beyond nf are computed on frame nf - 1 and not stored (keeps the warp converged)
    for (int fr0 = warp * 2; fr0 < nf; fr0 += FB_WARPS * 2) {
        const int fr = fr0 + half;
        const bool live = fr < nf;
        const float* x = s.samples + (live ? fr : nf - 1) * SHIFT;
        // DC offset over the 400 samples of the frame
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < WIN / 16; ++i) sum += x[hl + 16 * i] * pcm_scale;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum * (1.0f / WIN);
        // z[n] = y[2n] + i y[2n+1], n = hl + 16 n2 (pre-emphasis 0.97, Povey window, zero-padded beyond 200)
        float2 v[16];
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) {
            const int n = hl + 16 * n2;
            float2 z = make_float2(0.f, 0.f);
            if (n < WIN / 2) {
                const int j = 2 * n;
                const float xm1 = x[j > 0 ? j - 1 : 0] * pcm_scale - mean;
                const float x0 = x[j] * pcm_scale - mean;
                const float x1 = x[j + 1] * pcm_scale - mean;
                z.x = (x0 - 0.97f * xm1) * s.window[j];
                z.y = (x1 - 0.97f * x0) * s.window[j + 1];
            }
            v[n2] = z;
        }
        dft16(v);                                   // A[n1 = hl][k2]
#pragma unroll
        for (int k2 = 1; k2 < 16; ++k2) v[k2] = cmul(v[k2], s.w256[(hl * k2) & (NC - 1)]);
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) tile[k2 * 17 + hl] = v[k2];   // transpose through the padded tile
        __syncwarp();
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) v[n1] = tile[hl * 17 + n1];   // lane k2 = hl: A'[n1][k2]
        __syncwarp();
        dft16(v);                                   // Z[16 k1 + hl] in v[k1]
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) tile[16 * k1 + hl] = v[k1];
        __syncwarp();
        // untangle the packed real transform: X[k] = E[k] + W512^k O[k], power = |X[k]|^2, k = 0..256
        auto untangle = [&](int k) -> float {
            const float2 zk = tile[k & (NC - 1)];
            const float2 zc = tile[(NC - k) & (NC - 1)];
            const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
            const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
            const float2 wo = cmul(s.w512[k], o);
            const float re = e.x + wo.x, im = e.y + wo.y;
            return re * re + im * im;
        };
        float pk[17];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = untangle(hl + 16 * i);    // independent iterations: loads overlap
        pk[16] = untangle(NC);                                           // (every lane computes it, lane 0 stores it)
        __syncwarp();                                                    // every Z value has been read: the tile may go
#pragma unroll
        for (int i = 0; i < 16; ++i) pw[hl + 16 * i] = pk[i];
        if (hl == 0) pw[NC] = pk[16];
        __syncwarp();
        // sparse mel projection + log: 80 bins over the 16 lanes
#pragma unroll
        for (int mi = 0; mi < NMEL / 16; ++mi) {
            const int m = hl + 16 * mi;
            const int st = s.mel_start[m], ln = s.mel_len[m], of = s.mel_off[m];
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // four partial sums: the tap loop is a latency chain otherwise
            int i = 0;
            for (; i + 4 <= ln; i += 4) {
                a0 = fmaf(pw[st + i], s.mel_w[of + i], a0);
                a1 = fmaf(pw[st + i + 1], s.mel_w[of + i + 1], a1);
                a2 = fmaf(pw[st + i + 2], s.mel_w[of + i + 2], a2);
                a3 = fmaf(pw[st + i + 3], s.mel_w[of + i + 3], a3);
            }
            for (; i < ln; ++i) a0 = fmaf(pw[st + i], s.mel_w[of + i], a0);
            const float acc = (a0 + a1) + (a2 + a3);
            if (live) {
                const float val = logf(fmaxf(acc, 1.1920928955078125e-07f));
                dst[(long long)fr * NMEL + m] = val;
                st_sum[mi] += val; st_sq[mi] = fmaf(val, val, st_sq[mi]);
            }
        }
        __syncwarp();
    }
    if (stats != nullptr) {
#pragma unroll
        for (int mi = 0; mi < NMEL / 16; ++mi) {
            atomicAdd(&s.stat[hl + 16 * mi][0], (double)st_sum[mi]);
            atomicAdd(&s.stat[hl + 16 * mi][1], (double)st_sq[mi]);
        }
    }
    __syncthreads();

    if (stats != nullptr && tid < NMEL) {
        atomicAdd(&stats[((long long)b * NMEL + tid) * 2 + 0], s.stat[tid][0]);
        atomicAdd(&stats[((long long)b * NMEL + tid) * 2 + 1], s.stat[tid][1]);
    }
}

// CMVN: x <- (x - mean) / sqrt(max(E[x^2] - mean^2, 1e-10)) with fp64 statistics, in place; optionally
// also writes bf16 features into a zero-padded [B, padded_frames, 80] buffer (front_pad zero frames
// before each utterance, zeros after its last frame) for the conv subsampler.
__global__ void cmvn_kernel(float* __restrict__ feats, const long long* __restrict__ frame_off,
                            const double* __restrict__ stats, int normalize, bf16* __restrict__ out_bf16,
                            int padded_frames, int front_pad) {
    const int b = blockIdx.y;
    const long long f_begin = frame_off[b];
    const int nframes = (int)(frame_off[b + 1] - f_begin);
    __shared__ double s_mean[NMEL], s_rstd[NMEL];
    if (threadIdx.x < NMEL) {
        double mean = 0.0, rstd = 1.0;
        if (normalize && nframes > 0) {
            const double sm = stats[((long long)b * NMEL + threadIdx.x) * 2 + 0];
            const double sq = stats[((long long)b * NMEL + threadIdx.x) * 2 + 1];
            mean = sm / nframes;
            double var = sq / nframes - mean * mean;
            if (var < 1e-10) var = 1e-10;
            rstd = 1.0 / sqrt(var);
        }
        s_mean[threadIdx.x] = mean; s_rstd[threadIdx.x] = rstd;
    }
    __syncthreads();
    const int rows = out_bf16 ? padded_frames : nframes;
    const long long total = (long long)rows * NMEL;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i % NMEL);
        const long long r = i / NMEL;
        const long long f = out_bf16 ? r - front_pad : r;
        float v = 0.f;
        if (f >= 0 && f < nframes) {
            float* p = feats + (f_begin + f) * NMEL + m;
            v = *p;
            if (normalize) {
                v = (float)(((double)v - s_mean[m]) * s_rstd[m]);
                *p = v;
            }
        }
        if (out_bf16) out_bf16[((long long)b * padded_frames + r) * NMEL + m] = f_to_bf16(v);
    }
}

// ---------------------------------------------------------------------------------------------
// host: tables (computed once per process in fp64, mirroring torchaudio's formulas)
// ---------------------------------------------------------------------------------------------
static double mel_scale(double f) { return 1127.0 * log(1.0 + f / 700.0); }

static int upload_fbank_tables() {
    static std::mutex mu;
    static int done_for_device[64] = {0};
    int dev = 0;
    PB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 64 && done_for_device[dev]) return 0;
    static FbankTables t;
    const double PI = 3.14159265358979323846;
    for (int i = 0; i < WIN; ++i) {
        // torch.hann_window(400, periodic=False).pow(0.85), evaluated in fp32 like torch does
        const float hann = (float)(0.5 - 0.5 * cos(2.0 * PI * i / (WIN - 1)));
        t.window[i] = powf(hann, 0.85f);
    }
    for (int k = 0; k < NC; ++k) t.w256[k] = make_float2((float)cos(-2.0 * PI * k / NC), (float)sin(-2.0 * PI * k / NC));
    for (int k = 0; k <= NC; ++k) t.w512[k] = make_float2((float)cos(-2.0 * PI * k / NFFT), (float)sin(-2.0 * PI * k / NFFT));
    // get_mel_banks(80, 512, 16000, low=20, high=0 -> nyquist): triangular filters in the mel domain
    const double sample_freq = 16000.0, low = 20.0, high = 8000.0;
    const double fft_bin_width = sample_freq / NFFT;
    const double mel_low = mel_scale(low), mel_high = mel_scale(high);
    const double delta = (mel_high - mel_low) / (NMEL + 1);
    int off = 0;
    for (int m = 0; m < NMEL; ++m) {
        const double left = mel_low + m * delta, center = left + delta, right = center + delta;
        int start = -1, len = 0;
        std::vector<float> w;
        for (int k = 0; k < NC; ++k) {  // torchaudio builds 256 columns and pads the Nyquist column with 0
            const double mel = mel_scale(fft_bin_width * k);
            const double up = (mel - left) / (center - left), down = (right - mel) / (right - center);
            const double v = fmax(0.0, fmin(up, down));
            if (v > 0.0) {
                if (start < 0) start = k;
                while ((int)w.size() < k - start) w.push_back(0.f);
                w.push_back((float)v);
            }
        }
        if (start < 0) start = 0;
        len = (int)w.size();
        PB_CHECK(off + len <= MAX_MEL_TAPS, "mel filter table overflow");
        t.mel_start[m] = start; t.mel_len[m] = len; t.mel_off[m] = off;
        for (int i = 0; i < len; ++i) t.mel_w[off + i] = w[i];
        off += len;
    }
    for (int i = off; i < MAX_MEL_TAPS; ++i) t.mel_w[i] = 0.f;
    PB_CUDA(cudaMemcpyToSymbol(g_fbank_tables, &t, sizeof(FbankTables)));
    PB_CUDA(cudaFuncSetAttribute(fbank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FbankSmem)));
    PB_CUDA(cudaFuncSetAttribute(fbank_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FbankSmem2)));
    PB_CUDA(cudaFuncSetAttribute(fbank_kernel_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FbankSmem3)));
    if (dev < 64) done_for_device[dev] = 1;
    return 0;
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_fbank_init(void) { return upload_fbank_tables(); }

extern "C" int64_t pb200_fbank_num_frames(int64_t num_samples) {
    return num_samples < WIN ? 0 : 1 + (num_samples - WIN) / SHIFT;
}

extern "C" size_t pb200_fbank_workspace_bytes(int32_t batch) { return (size_t)batch * NMEL * 2 * sizeof(double); }

extern "C" int pb200_fbank_cmvn(const float* pcm, const int64_t* sample_offsets, const int64_t* frame_offsets,
                                int32_t batch, int32_t max_frames, float pcm_scale, int32_t cmvn, float* feats,
                                void* feats_bf16_padded, int32_t padded_frames, int32_t front_pad, void* workspace,
                                size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PB_CHECK(batch >= 0 && max_frames >= 0, "negative sizes");
    if (batch == 0) return 0;
    PB_CHECK(pcm && sample_offsets && frame_offsets && feats, "null pointer");
    PB_CHECK(workspace_bytes >= pb200_fbank_workspace_bytes(batch) && workspace, "fbank workspace too small");
    PB_CHECK(!feats_bf16_padded || padded_frames >= max_frames + front_pad, "padded_frames too small");
    int rc = upload_fbank_tables();
    if (rc) return rc;
    double* stats = reinterpret_cast<double*>(workspace);
    PB_CUDA(cudaMemsetAsync(stats, 0, pb200_fbank_workspace_bytes(batch), stream));
    if (max_frames > 0) {
        dim3 grid((max_frames + FR_PER_CTA - 1) / FR_PER_CTA, batch);
        static const bool use_v1 = [] { const char* e = getenv("PB200_FBANK"); return e && !strcmp(e, "v1"); }();
        static const bool use_v3 = [] { const char* e = getenv("PB200_FBANK"); return e && !strcmp(e, "v3"); }();
        if (use_v3)
            fbank_kernel_v3<<<grid, FB_THREADS, sizeof(FbankSmem3), stream>>>(
                pcm, reinterpret_cast<const long long*>(sample_offsets), reinterpret_cast<const long long*>(frame_offsets),
                feats, stats, pcm_scale);
        else if (use_v1)
            fbank_kernel<<<grid, FB_THREADS, sizeof(FbankSmem), stream>>>(
                pcm, reinterpret_cast<const long long*>(sample_offsets), reinterpret_cast<const long long*>(frame_offsets),
                feats, stats, pcm_scale);
        else
            fbank_kernel_v2<<<grid, FB_THREADS, sizeof(FbankSmem2), stream>>>(
                pcm, reinterpret_cast<const long long*>(sample_offsets), reinterpret_cast<const long long*>(frame_offsets),
                feats, stats, pcm_scale);
        PB_LAUNCH_CHECK();
    }
    if (cmvn || feats_bf16_padded) {
        const int rows = feats_bf16_padded ? padded_frames : max_frames;
        if (rows > 0) {
            int bx = (int)(((long long)rows * NMEL + 255) / 256);
            if (bx > 64) bx = 64;
            dim3 grid(bx, batch);
            cmvn_kernel<<<grid, 256, 0, stream>>>(feats, reinterpret_cast<const long long*>(frame_offsets), stats, cmvn,
                                                 reinterpret_cast<bf16*>(feats_bf16_padded), padded_frames, front_pad);
            PB_LAUNCH_CHECK();
        }
    }
    return 0;
}
