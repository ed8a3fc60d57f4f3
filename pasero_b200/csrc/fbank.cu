// pasero_b200 — fused Kaldi-compatible log-mel filterbank + utterance CMVN for 16 kHz PCM.
//
// One pass over the PCM: framing (25 ms / 10 ms, snip_edges), per-frame DC removal, pre-emphasis 0.97, Povey window,
// zero-pad 400 -> 512, real FFT (a 256-point complex FFT of the even/odd packed frame, 16 lanes per frame, 16 x 16 in
// registers), power spectrum, 80-bin mel projection, log(max(., FLT_EPSILON)), per-utterance mean / variance
// normalisation, optionally straight into the zero-padded bf16 layout that the conv subsampler's TMA descriptors read.
// fbank_kernel_v4 (below) does the transform in ONE persistent launch; cmvn_layout_kernel, launched behind it with programmatic
// dependent launch, normalises the (L2-resident) rows and writes the output layouts.
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

struct FbankTables {
    float window[WIN];
    float2 w256[NC];        // exp(-2 pi i k / 256)
    float2 w512[NC + 1];    // exp(-2 pi i k / 512), k = 0..256
};

__device__ FbankTables g_fbank_tables;

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}


// ---------------------------------------------------------------------------------------------
// Register-resident FFT: 16 lanes per frame (two frames per warp), 16 complex points per lane.
//   256-point complex FFT as 16 x 16 (Cooley-Tukey): lane n1 holds z[n1 + 16 n2] and runs a 16-point DFT over n2 in
//   registers, multiplies by W256^(n1 k2), the 16 x 16 values are transposed between the lanes through a padded
//   shared-memory tile (conflict free), lane k2 runs the second 16-point DFT over n1 and holds Z[16 k1 + k2].
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

// ---------------------------------------------------------------------------------------------
// v4 (default): ONE persistent kernel for framing, windowing, FFT, mel, log AND the utterance CMVN.
//
//   * grid = 2 CTAs per SM, each walks a contiguous range of 32-frame tiles; the tile's 5 360 PCM samples arrive by
//     cp.async into a double buffer while the previous tile is being transformed (no register staging);
//   * 16 lanes per frame (256-point complex FFT of the even/odd packed frame as 16 x 16, one padded shared-memory
//     transpose), but every per-lane constant lives in registers for the life of the CTA: the lane's 26 window taps
//     (pre-multiplied by the PCM scale), its 15 inter-stage twiddles, its untangle phase; DC removal and pre-emphasis are
//     folded into one 3-instruction form  z[j] = w[j] * (x[j] - 0.97 x[j-1] - 0.03 mean)  and the frame is read from shared
//     memory exactly once (the DC sum comes from the same registers);
//   * the real-FFT untangle exchanges Z[k] <-> Z[256-k] between lanes l and 16-l with shuffles (no shared-memory round
//     trip) and needs no twiddle table: W512^(16 k1 + l) = W32^k1 (immediates after unrolling) * W512^l (lane constant);
//   * mel projection through the triangle-complement identity: the 81 intervals between consecutive filter centres
//     partition the FFT bins; with T_j = sum of P over interval j and U_j = sum of u_k P (u_k = position of bin k inside
//     its interval), filter m = U_m + (T_{m+1} - U_{m+1}). Each power value is read once instead of twice, and the power
//     spectrum is stored in a permuted, padded layout (built on the host) in which lane l finds interval l + 16 i at a
//     fixed offset: the 20-step accumulation is fully unrolled, conflict free and free of per-lane loop bounds;
//   * the tile's log-mel rows are staged in shared memory, copied out with one coalesced pass and summed per bin in fp64
//     (E[x^2] - mean^2 needs every bit); the sums stay in registers across the tiles of one utterance and end in the
//     utterance's global accumulators with one atomicAdd per bin and run;
//   * utterance CMVN + output layouts are a second, PDL-launched, fully parallel pass (cmvn_layout_kernel) over the rows,
//     which are still L2 resident: a "last CTA of the utterance normalises it" variant inside this kernel was measured
//     first and left 148 SMs idle behind 64 finalising CTAs (~20 us of tail, ncu: SMs active 76 % of the duration).
// ---------------------------------------------------------------------------------------------
static constexpr int F4_THREADS = 256;
static constexpr int F4_HW = F4_THREADS / 16;                        // half-warps per CTA
static constexpr int F4_FRAMES = 2 * F4_HW;                          // frames per tile
static constexpr int F4_SAMPLES = (F4_FRAMES - 1) * SHIFT + WIN;     // 5360
static constexpr int F4_NGROUP = 5;                                  // mel intervals per lane (j = lane + 16 i)
// Slot of bin k = group base + F4_STRIDE * (position inside its interval) + lane. Loads (fixed position, consecutive lanes)
// are conflict free for any stride; stride 19 and rows of 400 floats (the two half-warps of a warp 16 banks apart) keep
// the scattered stores of 16 consecutive bins at <= 2 wavefronts (exhaustive search over strides / offsets; 17 gave 3.5).
static constexpr int F4_STRIDE = 19;
// widest interval (FFT bins) of each group (checked on the host) and the group's first slot = stride * running sum of the widths
__host__ __device__ constexpr int f4_maxc(int i) { return i == 0 ? 2 : i == 1 ? 2 : i == 2 ? 3 : i == 3 ? 5 : 8; }
__host__ __device__ constexpr int f4_base(int i) { return F4_STRIDE * (i == 0 ? 0 : i == 1 ? 2 : i == 2 ? 4 : i == 3 ? 7 : 12); }
static constexpr int F4_LAST = 20 * 19;                              // interval 80 (8 bins): one bin per lane 0..7
static constexpr int F4_TRASH = F4_LAST + 8;                         // bin 0 (below the 20 Hz cut-off)
static constexpr int F4_PB = 400;

struct Fbank4Tables {
    short perm[NC];        // FFT bin k -> slot in the permuted power buffer
    float uperm[F4_PB];    // slot -> u_k (0 for unused slots)
};
__device__ Fbank4Tables g_fbank4_tables;

struct Fbank4Smem {
    alignas(16) float samples[2][F4_SAMPLES + 16];      // + 16: the window's zero taps of the last frame read past the tile
    alignas(16) float2 tile[F4_HW][FB2_TILE];
    alignas(16) float pbuf[F4_HW][F4_PB];
    alignas(16) float uperm[F4_PB];
    alignas(16) float stash0[F4_HW][NMEL];     // log-mel rows of the tile's first 16 frames (the other 16 wait in `tile`)
    alignas(16) uint32_t pperm[2][16][4];   // lane -> byte offsets into its power buffer of bins 16 k1 + lane, two per word
};

static_assert(2 * (sizeof(Fbank4Smem) + 1024) <= 228 * 1024, "two fbank CTAs (256 threads x 128 registers each) must fit one SM's shared memory");

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// W32^k = exp(-2 pi i k / 32) as (cos, sin) of the positive angle; k is a compile-time constant after unrolling
__device__ __forceinline__ float2 cs32(int k) {
    switch (k) {
        case 0: return make_float2(1.f, 0.f);
        case 1: return make_float2(0.98078528040323043f, 0.19509032201612825f);
        case 2: return make_float2(0.92387953251128674f, 0.38268343236508978f);
        case 3: return make_float2(0.83146961230254524f, 0.55557023301960218f);
        case 4: return make_float2(0.70710678118654752f, 0.70710678118654752f);
        case 5: return make_float2(0.55557023301960218f, 0.83146961230254524f);
        case 6: return make_float2(0.38268343236508978f, 0.92387953251128674f);
        case 7: return make_float2(0.19509032201612825f, 0.98078528040323043f);
        case 8: return make_float2(0.f, 1.f);
        case 9: return make_float2(-0.19509032201612825f, 0.98078528040323043f);
        case 10: return make_float2(-0.38268343236508978f, 0.92387953251128674f);
        case 11: return make_float2(-0.55557023301960218f, 0.83146961230254524f);
        case 12: return make_float2(-0.70710678118654752f, 0.70710678118654752f);
        case 13: return make_float2(-0.83146961230254524f, 0.55557023301960218f);
        case 14: return make_float2(-0.92387953251128674f, 0.38268343236508978f);
        default: return make_float2(-0.98078528040323043f, 0.19509032201612825f);
    }
}

// stats == nullptr: no statistics wanted (no CMVN pass follows)
__global__ void __launch_bounds__(F4_THREADS, 2)
fbank_kernel_v4(const float* __restrict__ pcm, const long long* __restrict__ sample_off,
                const long long* __restrict__ frame_off, float* __restrict__ feats, double* __restrict__ stats,
                int batch, int tiles_per_utt, float pcm_scale) {
    pdl_trigger();      // the CMVN / layout kernel may be scheduled as soon as SMs free up (it waits for this grid's completion)
    extern __shared__ __align__(16) uint8_t fb_smem_raw[];
    Fbank4Smem& sm = *reinterpret_cast<Fbank4Smem*>(fb_smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int hl = lane & 15, half = lane >> 4, hw = warp * 2 + half;
    const unsigned FULL = 0xffffffffu;

    // ---- lane constants ----
    const FbankTables& T = g_fbank_tables;
    float wr[26];
#pragma unroll
    for (int n2 = 0; n2 < 13; ++n2) {
        const int j = 2 * hl + 32 * n2;
        wr[2 * n2] = j < WIN ? T.window[j] * pcm_scale : 0.f;
        wr[2 * n2 + 1] = j + 1 < WIN ? T.window[j + 1] * pcm_scale : 0.f;
    }
    float2 tw[16];
#pragma unroll
    for (int k2 = 1; k2 < 16; ++k2) tw[k2] = T.w256[(hl * k2) & (NC - 1)];
    const float2 g = make_float2(T.w512[hl].y, -T.w512[hl].x);      // -i * W512^hl
    if (tid < 128) {
        const int l = tid >> 3, q = tid & 7;
        sm.pperm[q >> 2][l][q & 3] = (uint32_t)g_fbank4_tables.perm[16 * (2 * q) + l] * 4u |
                         ((uint32_t)g_fbank4_tables.perm[16 * (2 * q + 1) + l] * 4u) << 16;
    }
    for (int i = tid; i < F4_PB; i += F4_THREADS) sm.uperm[i] = g_fbank4_tables.uperm[i];
    for (int i = tid; i < F4_HW * F4_PB; i += F4_THREADS) (&sm.pbuf[0][0])[i] = 0.f;    // unused slots stay 0 for good
    for (int i = tid; i < 2 * (F4_SAMPLES + 16); i += F4_THREADS) (&sm.samples[0][0])[i] = 0.f;   // never a NaN bit pattern

    __syncthreads();

    const long long total_tiles = (long long)batch * tiles_per_utt;
    const int t_hi = (int)(((long long)blockIdx.x + 1) * total_tiles / gridDim.x);
    // first non-empty tile at or after t (CTA-uniform)
    auto next_tile = [&](int t) {
        while (t < t_hi) {
            const int b = t / tiles_per_utt;
            const int nframes = (int)(frame_off[b + 1] - frame_off[b]);
            if ((t - b * tiles_per_utt) * F4_FRAMES < nframes) break;
            ++t;
        }
        return t;
    };
    auto issue_loads = [&](int t, int buf) {
        const int b = t / tiles_per_utt, f0 = (t - b * tiles_per_utt) * F4_FRAMES;
        const int nframes = (int)(frame_off[b + 1] - frame_off[b]);
        const int nf = min(F4_FRAMES, nframes - f0);
        const int nsamp = (nf - 1) * SHIFT + WIN;                 // a multiple of 4
        const float* src = pcm + sample_off[b] + (long long)f0 * SHIFT;
        float* dst = sm.samples[buf];
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            for (int i = tid; i < (nsamp >> 2); i += F4_THREADS) cp_async16(dst + 4 * i, src + 4 * i);
        } else {
            for (int i = tid; i < nsamp; i += F4_THREADS) cp_async4(dst + i, src + i);
        }
    };

    double s1 = 0.0, s2 = 0.0;      // threads 0..79: sum / sum of squares of bin `tid` over the current run of tiles
    int run_b = -1;

    // statistics of the finished run of tiles (all of utterance b) -> the utterance's fp64 accumulators
    auto flush = [&](int b) {
        if (stats != nullptr && tid < NMEL) {
            atomicAdd(&stats[((long long)b * NMEL + tid) * 2 + 0], s1);
            atomicAdd(&stats[((long long)b * NMEL + tid) * 2 + 1], s2);
        }
        s1 = s2 = 0.0;
    };

    int cur = next_tile((int)((long long)blockIdx.x * total_tiles / gridDim.x));
    int buf = 0;
    if (cur < t_hi) issue_loads(cur, buf);
    cp_async_commit();
    float2* tile = sm.tile[hw];
    float* pb = sm.pbuf[hw];

    while (cur < t_hi) {
        const int b = cur / tiles_per_utt, f0 = (cur - b * tiles_per_utt) * F4_FRAMES;
        const long long f_begin = frame_off[b];
        const int nf = min(F4_FRAMES, (int)(frame_off[b + 1] - f_begin) - f0);
        if (b != run_b) {
            if (run_b >= 0) flush(run_b);
            run_b = b;
        }
        cp_async_wait_all();
        __syncthreads();                       // the tile's samples are visible; everybody is done with the other buffer
        const int nxt = next_tile(cur + 1);
        if (nxt < t_hi) issue_loads(nxt, buf ^ 1);
        cp_async_commit();

#pragma unroll 1
        for (int rr = 0; rr < 2; ++rr) {
            const int fr = hw + F4_HW * rr;
            const bool live = fr < nf;
            const float* x = sm.samples[buf] + (live ? fr : nf - 1) * SHIFT;   // dead lanes redo the last frame (no store)
            float2 v[16];
            float xm[13];
            float sum = 0.f;
#pragma unroll
            for (int n2 = 0; n2 < 13; ++n2) {
                // n2 == 12, lanes >= 8: samples 400.. of the frame. Their window taps are 0 and the sample buffers only ever
                // hold finite values (zero-filled at start), so they are loaded like the rest and only kept out of the DC sum
                const int j = 2 * hl + 32 * n2;
                v[n2] = *reinterpret_cast<const float2*>(x + j);
                xm[n2] = x[(n2 == 0 && hl == 0) ? 0 : j - 1];
                const float part = v[n2].x + v[n2].y;
                sum += (n2 < 12 || hl < 8) ? part : 0.f;
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
            const float c = sum * ((1.f - 0.97f) / WIN);        // pre-emphasised DC offset
#pragma unroll
            for (int n2 = 0; n2 < 13; ++n2) {
                const float y0 = fmaf(-0.97f, xm[n2], v[n2].x) - c;
                const float y1 = fmaf(-0.97f, v[n2].x, v[n2].y) - c;
                v[n2] = make_float2(wr[2 * n2] * y0, wr[2 * n2 + 1] * y1);
            }
            v[13] = v[14] = v[15] = make_float2(0.f, 0.f);
            dft16(v);                                   // A[n1 = hl][k2]
#pragma unroll
            for (int k2 = 1; k2 < 16; ++k2) v[k2] = cmul(v[k2], tw[k2]);
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) tile[k2 * 17 + hl] = v[k2];
            __syncwarp();
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) v[n1] = tile[hl * 17 + n1];
            __syncwarp();
            dft16(v);                                   // Z[16 k1 + hl] in v[k1]
            // real-FFT untangle: 2 X[k] = (Z[k] + conj Z[256-k]) + (-i W512^k) (Z[k] - conj Z[256-k]);  Z[256-k] sits in
            // lane 16 - hl, register 15 - k1 (lane 0: its own register 16 - k1)
            uint32_t pq[8];
            {
                const uint4 a = *reinterpret_cast<const uint4*>(&sm.pperm[0][hl][0]);
                const uint4 bq = *reinterpret_cast<const uint4*>(&sm.pperm[1][hl][0]);
                pq[0] = a.x; pq[1] = a.y; pq[2] = a.z; pq[3] = a.w; pq[4] = bq.x; pq[5] = bq.y; pq[6] = bq.z; pq[7] = bq.w;
            }
            char* pbytes = reinterpret_cast<char*>(pb);
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
                const float2 give = v[15 - k1];
                float2 zc;
                zc.x = __shfl_sync(FULL, give.x, (16 - hl) & 15, 16);
                zc.y = __shfl_sync(FULL, give.y, (16 - hl) & 15, 16);
                if (hl == 0) zc = v[(16 - k1) & 15];
                const float2 zk = v[k1];
                const float sx = zk.x + zc.x, sy = zk.y - zc.y;
                const float dx = zk.x - zc.x, dy = zk.y + zc.y;
                const float tx = g.x * dx - g.y * dy, ty = g.x * dy + g.y * dx;
                const float2 w = cs32(k1);              // W32^k1 = (w.x, -w.y)
                const float xr = fmaf(w.x, tx, fmaf(w.y, ty, sx));
                const float xi = fmaf(w.x, ty, fmaf(-w.y, tx, sy));
                const uint32_t off = (k1 & 1) ? pq[k1 >> 1] >> 16 : pq[k1 >> 1] & 0xffffu;
                *reinterpret_cast<float*>(pbytes + off) = fmaf(xr, xr, xi * xi);    // 4 |X[k]|^2
            }
            __syncwarp();
            float um[F4_NGROUP], vm[F4_NGROUP];
#pragma unroll
            for (int i = 0; i < F4_NGROUP; ++i) {
                float ts = 0.f, us = 0.f;
#pragma unroll
                for (int t = 0; t < f4_maxc(i); ++t) {
                    const float pv = pb[f4_base(i) + F4_STRIDE * t + hl];
                    ts += pv;
                    us = fmaf(sm.uperm[f4_base(i) + F4_STRIDE * t + hl], pv, us);
                }
                um[i] = us; vm[i] = ts - us;
            }
            // interval 80: one bin per lane 0..7; lanes 8..15 read slots whose u is 1 (contribution pv - 1 * pv = 0)
            float v80;
            { const float pv = pb[F4_LAST + hl]; v80 = fmaf(-sm.uperm[F4_LAST + hl], pv, pv); }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v80 += __shfl_xor_sync(FULL, v80, o);
            __syncwarp();                               // pbuf / tile are free for the next frame
            float nb[F4_NGROUP];
#pragma unroll
            for (int i = 0; i < F4_NGROUP; ++i) nb[i] = __shfl_sync(FULL, vm[i], (hl + 1) & 15, 16);
            // the frame's row waits for the tile's coalesced copy-out: the first frame of this half-warp in the small stash,
            // the second one in the half-warp's own transpose tile (free until the next tile's first stage)
            float* dst = (rr == 0 ? &sm.stash0[hw][0] : reinterpret_cast<float*>(tile)) + hl;
#pragma unroll
            for (int i = 0; i < F4_NGROUP; ++i) {
                const float up = hl < 15 ? nb[i] : (i < F4_NGROUP - 1 ? nb[i + 1 < F4_NGROUP ? i + 1 : i] : v80);
                const float mel4 = um[i] + up;
                // mel = mel4 / 4: log(max(mel, eps)) = log(max(mel4, 4 eps)) - log 4
                const float val = logf(fmaxf(mel4, 4.f * 1.1920928955078125e-07f)) - 1.3862943611198906f;
                dst[16 * i] = val;
            }
        }
        __syncthreads();
        {   // the tile's rows: one coalesced copy to global memory, and bin `tid`'s statistics in fp64.
            auto row_of = [&](int fr) -> const float* {
                return fr < F4_HW ? &sm.stash0[fr][0] : reinterpret_cast<const float*>(sm.tile[fr - F4_HW]);
            };
            float4* out4 = reinterpret_cast<float4*>(feats + (f_begin + f0) * NMEL);
            for (int i = tid; i < nf * (NMEL / 4); i += F4_THREADS) {
                const int fr = i / (NMEL / 4), m4 = i - fr * (NMEL / 4);
                out4[i] = *reinterpret_cast<const float4*>(row_of(fr) + 4 * m4);
            }
            if (tid < NMEL && stats != nullptr) {
                for (int fr = 0; fr < nf; ++fr) {
                    const double val = (double)row_of(fr)[tid];
                    s1 += val;
                    s2 = fma(val, val, s2);
                }
            }
        }
        cur = nxt;
        buf ^= 1;
    }
    if (run_b >= 0) flush(run_b);
}


// Utterance CMVN + output layouts (second pass of v4): x <- (x - mean) / sqrt(max(E[x^2] - mean^2, 1e-10)) from the fp64
// statistics. One CTA per (utterance, block of CM_ROWS output rows); 16-byte accesses; writes the fp32 rows in place
// (mode bit 0 without bit 1) and / or the zero-padded bf16 layout (features of utterance b at rows [front_pad, front_pad +
// frames(b)), zeros elsewhere). The last CTA to finish clears the statistics and its own counter: the workspace is left
// zeroed for the next call (no memset in the step).
static constexpr int CM_ROWS = 64;      // 64 rows x 20 float4 = 5 per thread, all five loads in flight together
__global__ void __launch_bounds__(256)
cmvn_layout_kernel(float* __restrict__ feats, const long long* __restrict__ frame_off, double* __restrict__ stats,
                   int* __restrict__ done, int batch, int mode, bf16* __restrict__ padded, int padded_frames, int front_pad,
                   int row_blocks) {
    __shared__ __align__(16) float s_mean[NMEL];
    __shared__ __align__(16) float s_rstd[NMEL];
    __shared__ int s_last;
    const int b = blockIdx.x / row_blocks, rb = blockIdx.x - b * row_blocks;
    const int tid = threadIdx.x;
    pdl_wait();                 // the fbank kernel's rows and statistics are complete and visible
    const long long f_begin = frame_off[b];
    const int nframes = (int)(frame_off[b + 1] - f_begin);
    const int rows = padded != nullptr ? padded_frames : nframes;
    const int r0 = rb * CM_ROWS, r1 = min(rows, r0 + CM_ROWS);
    if (r0 < r1) {
        const bool write_f32 = (mode & 1) && !(mode & 2);
        constexpr int PER = CM_ROWS * (NMEL / 4) / 256;
        float4 x[PER];
        float4* src[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int i = tid + u * 256;
            const int r = r0 + i / (NMEL / 4), m4 = i % (NMEL / 4);
            const int f = padded != nullptr ? r - front_pad : r;
            src[u] = (r < r1 && f >= 0 && f < nframes) ? reinterpret_cast<float4*>(feats + (f_begin + f) * NMEL) + m4 : nullptr;
            x[u] = src[u] != nullptr ? __ldcg(src[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (tid < NMEL) {
            float mean = 0.f, rstd = 1.f;
            if ((mode & 1) && nframes > 0) {
                const double sm1 = __ldcg(&stats[((long long)b * NMEL + tid) * 2 + 0]);
                const double sm2 = __ldcg(&stats[((long long)b * NMEL + tid) * 2 + 1]);
                const double mu = sm1 / nframes;
                double var = sm2 / nframes - mu * mu;
                if (var < 1e-10) var = 1e-10;
                mean = (float)mu;
                rstd = (float)(1.0 / sqrt(var));
            }
            s_mean[tid] = mean; s_rstd[tid] = rstd;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int i = tid + u * 256;
            const int r = r0 + i / (NMEL / 4), m4 = i % (NMEL / 4);
            if (r >= r1) continue;
            float4 v = x[u];
            if (src[u] != nullptr) {
                const float4 mu = *reinterpret_cast<const float4*>(&s_mean[4 * m4]);
                const float4 rs = *reinterpret_cast<const float4*>(&s_rstd[4 * m4]);
                v.x = (v.x - mu.x) * rs.x; v.y = (v.y - mu.y) * rs.y; v.z = (v.z - mu.z) * rs.z; v.w = (v.w - mu.w) * rs.w;
                if (write_f32) *src[u] = v;
            }
            if (padded != nullptr)
                *reinterpret_cast<uint2*>(padded + ((long long)b * padded_frames + r) * NMEL + 4 * m4) =
                    make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        }
    }
    // every CTA has read its statistics before the last one clears them
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(done, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        for (long long i = tid; i < (long long)batch * NMEL * 2; i += 256) stats[i] = 0.0;
        if (tid == 0) *done = 0;
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
    PB_CUDA(cudaMemcpyToSymbol(g_fbank_tables, &t, sizeof(FbankTables)));
    // v4: the 81 intervals between consecutive filter centres (interval j = [lo + j delta, lo + (j + 1) delta)) partition the
    // FFT bins 1..255; u_k = position of bin k inside its interval = its weight in filter j (1 - u_k in filter j - 1)
    static Fbank4Tables t4;
    {
        int cnt[NMEL + 1] = {0}, first[NMEL + 1];
        for (int j = 0; j <= NMEL; ++j) first[j] = -1;
        int jk[NC];
        for (int k = 0; k < NC; ++k) {
            const double pos = (mel_scale(fft_bin_width * k) - mel_low) / delta;
            const int j = pos < 0.0 ? -1 : (int)floor(pos);
            jk[k] = j;
            PB_CHECK(j <= NMEL, "mel interval out of range");
            if (j >= 0) { if (first[j] < 0) first[j] = k; ++cnt[j]; }
        }
        for (int i = 0; i < F4_PB; ++i) t4.uperm[i] = 0.f;
        for (int i = 8; i < 16; ++i) t4.uperm[F4_LAST + i] = 1.f;    // lanes 8..15 of the interval-80 read: weight 1 - u = 0
        for (int k = 0; k < NC; ++k) {
            const int j = jk[k];
            int slot = F4_TRASH;
            if (j == NMEL) {
                PB_CHECK(k - first[j] < 8, "last mel interval wider than 8 bins");
                slot = F4_LAST + (k - first[j]);
            } else if (j >= 0) {
                const int i = j / 16, l = j % 16, tt = k - first[j];
                PB_CHECK(tt < f4_maxc(i), "mel interval wider than the unrolled accumulation");
                slot = f4_base(i) + F4_STRIDE * tt + l;
            }
            t4.perm[k] = (short)slot;
            if (j >= 0) t4.uperm[slot] = (float)((mel_scale(fft_bin_width * k) - mel_low) / delta - j);
        }
    }
    PB_CUDA(cudaMemcpyToSymbol(g_fbank4_tables, &t4, sizeof(Fbank4Tables)));
    PB_CUDA(cudaFuncSetAttribute(fbank_kernel_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Fbank4Smem)));
    // two CTAs per SM (2 x 108 KB + the per-block reserve): ask for the largest shared-memory carve-out
    PB_CUDA(cudaFuncSetAttribute(fbank_kernel_v4, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (dev < 64) done_for_device[dev] = 1;
    return 0;
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_fbank_init(void) { return upload_fbank_tables(); }

extern "C" int64_t pb200_fbank_num_frames(int64_t num_samples) {
    return num_samples < WIN ? 0 : 1 + (num_samples - WIN) / SHIFT;
}

// workspace: fp64 statistics [batch][80][2], then one int (completion counter of the CMVN pass). It must be ZERO when
// first used and is left zeroed by every successful call (the last CTA of the CMVN pass clears it), so a caller that
// keeps the buffer never needs a memset again.
extern "C" size_t pb200_fbank_workspace_bytes(int32_t batch) {
    return (size_t)batch * NMEL * 2 * sizeof(double) + 16;
}

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
    PB_CHECK(!(cmvn & 2) || feats_bf16_padded, "cmvn bit 1 (fp32 rows are scratch) needs the bf16 padded output");
    PB_CHECK((reinterpret_cast<uintptr_t>(feats) & 15) == 0 && (reinterpret_cast<uintptr_t>(feats_bf16_padded) & 7) == 0,
             "fbank outputs must be 16-byte aligned");
    int rc = upload_fbank_tables();
    if (rc) return rc;
    double* stats = reinterpret_cast<double*>(workspace);
    int* done = reinterpret_cast<int*>(stats + (size_t)batch * NMEL * 2);
    const bool second_pass = (cmvn & 1) || feats_bf16_padded;
    {
        const int tiles_per_utt = (max_frames + F4_FRAMES - 1) / F4_FRAMES;
        const long long total_tiles = (long long)batch * tiles_per_utt;
        static int sms = 0;
        if (sms == 0) {
            int dev = 0;
            PB_CUDA(cudaGetDevice(&dev));
            PB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        }
        if (total_tiles > 0) {
            const long long grid = total_tiles < 2LL * sms ? total_tiles : 2LL * sms;
            fbank_kernel_v4<<<(unsigned)grid, F4_THREADS, sizeof(Fbank4Smem), stream>>>(
                pcm, reinterpret_cast<const long long*>(sample_offsets), reinterpret_cast<const long long*>(frame_offsets), feats,
                (cmvn & 1) ? stats : nullptr, batch, tiles_per_utt, pcm_scale);
            PB_LAUNCH_CHECK();
        }
        if (second_pass) {
            const int rows = feats_bf16_padded ? padded_frames : max_frames;
            const int row_blocks = (rows + CM_ROWS - 1) / CM_ROWS;
            if (row_blocks > 0) {
                PB_CUDA(launch_pdl(cmvn_layout_kernel, dim3((unsigned)(batch * row_blocks)), dim3(256), 0, stream, feats,
                                   reinterpret_cast<const long long*>(frame_offsets), stats, done, batch, cmvn,
                                   reinterpret_cast<bf16*>(feats_bf16_padded), padded_frames, front_pad, row_blocks));
                PB_LAUNCH_CHECK();
            }
        }
    }
    return 0;
}
