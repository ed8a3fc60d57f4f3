// pasero_b200 — tcgen05/TMEM attention, head_dim 64, bf16 in / fp32 softmax, any sequence length.
// One block of <= 256 keys (and, in backward, <= 256 queries) is processed entirely on chip — this is the whole problem for
// the ST training step (250 encoder positions, 48 target positions). Longer sequences are tiled over such blocks by the
// grid: forward CTAs write per-key-block partial outputs + log-sum-exps that attn_combine_kernel merges (exact: the
// partials are softmax-normalised inside their block); backward CTAs own a (query block, key block) pair — P is rebuilt
// from the GLOBAL log-sum-exp, so the pair's dQ / dK / dV contributions are exact partial sums that attn_reduce_kernel adds.
// Attention dropout (pasero/models/modules.py:686-731): the forward kernel draws the keep bits of its P tile from
// Philox(seed, offset) — 16 bits per decision — writes them as one uint32 per (query, 32 keys) and the backward kernel
// reads them back (P o M / (1 - p) feeds the P V and dV products, dP o M / (1 - p) the softmax backward).
//
// Forward, one CTA per (128-query tile, head, batch element):
//   TMA loads Q [128 x 64], K and V [Lk x 64] (SWIZZLE_128B boxes cut straight out of the fused projection output),
//   S = Q K^T is ONE 128 x Lk tcgen05.mma chain into TMEM, the four softmax warps (one query row per thread) read S
//   with tcgen05.ld, take the row max and exp2 in registers, and hand P (bf16) back either in place in TMEM
//   (tcgen05.st; P is then the TMEM A operand of the second MMA) or through swizzled shared memory,
//   O = P V accumulates in TMEM and is normalised by the row sum on its way out. No online-softmax rescaling is
//   needed because the whole key range is resident.
// Backward, one CTA per (head, batch element), everything on chip, no recomputation and no atomics:
//   for each 128-key tile j and 128-query chunk i:  S^T = K_j Q_i^T and dP^T = V_j dO_i^T (TMEM), the softmax warps
//   (one key row per thread) form P^T and dS^T = P^T o (dP^T - delta) in bf16 in shared memory, then
//   dV_j += P^T dO_i, dK_j += dS^T Q_i and dQ_i += dS K_j run as three more MMA chains (dS^T in shared memory serves
//   both as the K-major A operand of dK and as the MN-major A operand of dQ).
//
// Reference semantics: MultiheadAttention.forward's SDPA call with a key-padding mask or a causal mask
// (pasero/models/modules.py:648-734).
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "../../include/pasero_b200.h"

#include <stdlib.h>

namespace pb200 {

static constexpr float LOG2E_F = 1.4426950408889634f;
static constexpr int TC_THREADS = 288;   // 8 softmax/epilogue warps + 1 TMA/MMA warp

struct AttnTcDev {
    bf16* o; float* lse; const int* kv_len;
    long long ldo;
    int B, H, Lq, Lk, causal;
    int q_box, kv_box;          // rows per TMA box (multiples of 16)
    float scale;
    // backward
    const bf16* q; const bf16* k; const bf16* v; long long ldq, ldk, ldv;   // operand loads by cp.async (see the kernel)
    const bf16* o_in; const bf16* dout; long long lddo;
    const float* lse_in;
    bf16* dq; bf16* dk; bf16* dv;
    long long lddq, lddk, lddv;
    // tiling over blocks of 256 keys / queries
    int n_kc;                   // forward: key blocks (gridDim.x = query tiles * n_kc)
    bf16* o_part; float* lse_part;          // forward partials [n_kc][B*Lq][H*64] / [n_kc][B*H*Lq] when n_kc > 1
    int n_qb, n_kb;             // backward: query / key blocks (gridDim.z = n_qb * n_kb)
    int q_tma_rows, kv_tma_rows;            // rows of the TMA boxes (backward)
    bf16* dq_part; bf16* dk_part; bf16* dv_part;   // backward partials [n_kb][B*Lq][H*64], [n_qb][B*Lk][H*64] (x2)
    // attention dropout
    uint32_t* mask; int mask_ld;            // keep bits [B*H*Lq][mask_ld] (mask_ld = ceil(Lk / 32) words)
    uint32_t thresh16; float keep_scale;    // keep iff u16 >= thresh16; 1 / (1 - p)
    unsigned long long seed, offset; const long long* offset_base;
};

// 32 keep bits (bit j <-> key 32 * word + j) of one query row
__device__ __forceinline__ uint32_t attn_keep32(const Philox& rng, unsigned long long offset, long long row, int mask_ld, int word,
                                                uint32_t thresh16) {
    const unsigned long long idx = ((unsigned long long)row * mask_ld + word) * 4ull;
    uint32_t m = 0;
#pragma unroll
    for (int g = 0; g < 4; ++g) m |= dropout_keep8(rng, offset, idx + g, thresh16) << (8 * g);
    return m;
}

// Optional in-kernel timeline (build with -DPB_ATTN_TIMING): per CTA 16 slots of clock64() at the phase boundaries of
// the MMA warp (slots 0-7) and of softmax warp 0 (slots 8-13), slot 14 = %globaltimer at entry, slot 15 = SM id.
#ifdef PB_ATTN_TIMING
__device__ long long g_attn_t[4096 * 16];
#define TSTAMP(slot) do { if (lane == 0 && cta_lin < 4096) g_attn_t[cta_lin * 16 + (slot)] = clock64(); } while (0)
#else
#define TSTAMP(slot) do { } while (0)
#endif

__device__ __forceinline__ uint64_t kdesc(const uint8_t* base, int k16) {           // K-major operand, k-step of 16
    return ptx::make_smem_desc_sw128(ptx::smem_u32(base) + k16 * 32, 16, 1024);
}
__device__ __forceinline__ uint64_t mndesc(const uint8_t* base, int k16, uint32_t lbo) {   // MN-major operand
    return ptx::make_smem_desc_sw128(ptx::smem_u32(base) + k16 * 2048, lbo, 1024);
}

// The single MMA-issuing thread is latency-bound on its own instruction stream (measured: ~100 cycles per tcgen05.mma
// with descriptors rebuilt inside a rolled loop), so the issue loops below are unrolled and carry each descriptor as a
// loop-invariant hi word plus a lo word that takes one 32-bit add per k-step (the 14-bit start-address field counts
// 16-byte units and cannot carry: shared-memory addresses stay below 256 KB).
__device__ __forceinline__ uint32_t lo32(uint64_t d) { return (uint32_t)d; }
__device__ __forceinline__ uint32_t hi32(uint64_t d) { return (uint32_t)(d >> 32); }
// 16-byte units from the start of a K-major operand (64-column blocks of 16 KB = 128 rows x 128 B) to k-step kk
__device__ __forceinline__ constexpr uint32_t kstep16(int kk) { return (uint32_t)((kk >> 2) * 1024 + (kk & 3) * 2); }

// =====================================================================================================
// forward
// =====================================================================================================
// Eight softmax warps: warps w and w + 4 own the same 32 TMEM lanes (query rows) and split the key columns in two
// halves, so that every scheduler has two (four with two CTAs per SM) warps to issue from.
// TMEM columns (P in TMEM): S [0,256); after the softmax: P(keys 0..127) [0,64) | O_even [64,128) | P(keys 128..255)
// [128,192) | O_odd [192,256) — each warp writes its packed P over S columns it has itself already consumed.
// TMEM columns (P in shared memory): S [0,256) | O_even [256,320) | O_odd [320,384).
// O = P V is accumulated in two halves (even / odd 16-key steps): back-to-back tcgen05.mma into ONE accumulator
// serialise on its ~100-cycle latency when N is only 64 (measured), two accumulators keep the pipe busy.
template <bool P_TMEM>
struct FwdCfg {
    static constexpr int Q_BYTES = 128 * 128, KV_BYTES = 256 * 128, P_BYTES = P_TMEM ? 0 : 128 * 512;
    static constexpr int SMEM_BYTES = Q_BYTES + 2 * KV_BYTES + P_BYTES + 2 * 2 * 128 * 4 + 1024 /*align*/ + 64 /*barriers*/;
    static constexpr int TMEM_COLS = P_TMEM ? 256 : 512;
    static constexpr int O_COL = P_TMEM ? 64 : 256;
    static constexpr int O2_COL = P_TMEM ? 192 : 320;
};

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool P_TMEM, bool CAUSAL, bool DROPOUT>
__global__ void __launch_bounds__(TC_THREADS, P_TMEM ? 2 : 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                   const __grid_constant__ CUtensorMap tmv, const AttnTcDev p) {
    pdl_trigger();
    using Cfg = FwdCfg<P_TMEM>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + Cfg::Q_BYTES;
    uint8_t* sV = sK + Cfg::KV_BYTES;
    uint8_t* sP = sV + Cfg::KV_BYTES;
    float* sMax = reinterpret_cast<float*>(sP + Cfg::P_BYTES);   // [2][128] per column half
    float* sSum = sMax + 256;                                    // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sSum + 256);
    uint64_t* bar_load = bars, *bar_s = bars + 1, *bar_p = bars + 2, *bar_o = bars + 3;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_qt = gridDim.x / p.n_kc;
    const int kc = blockIdx.x / n_qt;                       // key block of this CTA (256 keys)
    const int q0 = (blockIdx.x - kc * n_qt) * 128, k0 = kc * 256, h = blockIdx.y, b = blockIdx.z;
#ifdef PB_ATTN_TIMING
    const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    if (warp == 8 && lane == 0 && cta_lin < 4096) {
        unsigned long long gt; unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_attn_t[cta_lin * 16 + 14] = (long long)gt; g_attn_t[cta_lin * 16 + 15] = smid;
    }
    if (warp == 8) TSTAMP(0);
#endif

    if (warp == 8) {
        if (lane == 0) {
            ptx::prefetch_tmap(&tmq); ptx::prefetch_tmap(&tmk); ptx::prefetch_tmap(&tmv);
            ptx::mbar_init(bar_load, 1); ptx::mbar_init(bar_s, 1); ptx::mbar_init(bar_p, 256); ptx::mbar_init(bar_o, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();   // set-up above overlaps the previous kernel; global memory is touched only below
    // keys of this block are indexed from k0: [0, kvlen) are unmasked, causal: key j visible to query i iff j <= i + coff
    const int kvlen = (p.kv_len ? min(p.kv_len[b], p.Lk) : p.Lk) - k0;
    const int coff = p.Lk - p.Lq - k0;
    int kv_hi = CAUSAL ? min(kvlen, q0 + 128 + coff) : kvlen;
    kv_hi = max(kv_hi, 1);   // a block without a visible key still runs: every P is 0, the partial is (0, -inf)
    const int n1 = min(((kv_hi + 15) >> 4) << 4, p.kv_box);   // keys covered by the MMAs (multiple of 16)
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + Cfg::O_COL, tmem_o2 = tmem_base + Cfg::O2_COL;

    if (warp == 8) {
        if (lane == 0) {
            TSTAMP(1);
            ptx::mbar_arrive_expect_tx(bar_load, (uint32_t)(128 + 2 * p.kv_box) * 128u);
            ptx::tma_load_3d(&tmq, bar_load, sQ, h * 64, q0, b);
            ptx::tma_load_3d(&tmk, bar_load, sK, h * 64, k0, b);
            ptx::tma_load_3d(&tmv, bar_load, sV, h * 64, k0, b);
            ptx::mbar_wait(bar_load, 0);
            ptx::tc_fence_after();
            TSTAMP(2);
            const uint32_t idesc1 = ptx::make_idesc_bf16(128, n1, false, false);
            const uint64_t qd = kdesc(sQ, 0), kd = kdesc(sK, 0);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                ptx::umma_bf16_lh(tmem_s, lo32(qd) + 2 * k, hi32(qd), lo32(kd) + 2 * k, hi32(kd), idesc1, k > 0 ? 1u : 0u);
            ptx::umma_commit(bar_s);
            TSTAMP(3);
            ptx::mbar_wait(bar_p, 0);
            ptx::tc_fence_after();
            TSTAMP(4);
            const uint32_t idesc2 = ptx::make_idesc_bf16(128, 64, false, true);
            const int ksteps = n1 >> 4;
            const uint64_t vd = mndesc(sV, 0, 8192), pd = kdesc(sP, 0);
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                if (kk < ksteps) {
                    // packed P of keys [0,128) sits at columns [0,64), of keys [128,256) at columns [128,192)
                    const uint32_t acc = (kk & 1) ? tmem_o2 : tmem_o;
                    if (P_TMEM) ptx::umma_bf16_ts_lh(acc, tmem_s + (kk >> 3) * 128 + (kk & 7) * 8, lo32(vd) + 128 * kk, hi32(vd), idesc2, kk > 1 ? 1u : 0u);
                    else ptx::umma_bf16_lh(acc, lo32(pd) + kstep16(kk), hi32(pd), lo32(vd) + 128 * kk, hi32(vd), idesc2, kk > 1 ? 1u : 0u);
                }
            }
            ptx::umma_commit(bar_o);
            TSTAMP(5);
        }
    } else {
        const int quarter = warp & 3, half = warp >> 2;
        const int row = quarter * 32 + lane;
        const int qrow = q0 + row;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const int lim = CAUSAL ? min(kvlen, qrow + coff + 1) : kvlen;   // keys [0, lim) are visible to this row
        const int lim_lo = CAUSAL ? __shfl_sync(0xffffffffu, lim, 0) : lim;   // smallest lim of this warp's rows
        const float sl2 = p.scale * LOG2E_F;
        const int nch = (n1 + 31) >> 5;
        const int c_begin = half * 4, c_end = min(nch, half * 4 + 4);   // this warp's 32-key chunks
        const Philox rng(p.seed);
        const unsigned long long drop_offset = DROPOUT ? p.offset + (p.offset_base ? ((unsigned long long)*p.offset_base << 12) : 0ull) : 0ull;
        if (warp == 0) TSTAMP(8);
        ptx::mbar_wait(bar_s, 0);
        ptx::tc_fence_after();
        if (warp == 0) TSTAMP(9);
        float m = -INFINITY;
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_s + lane_base + c * 32, v);
            ptx::tmem_ld_wait();
            if (c * 32 + 32 <= lim_lo) {   // warp-uniform: whole chunk visible to every row
                float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    m0 = fmaxf(m0, __uint_as_float(v[j])); m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
                    m2 = fmaxf(m2, __uint_as_float(v[j + 2])); m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
                }
                m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c * 32 + j < lim) m = fmaxf(m, __uint_as_float(v[j]));
            }
        }
        sMax[half * 128 + row] = m;
        ptx::bar_sync(1, 256);
        if (warp == 0) TSTAMP(10);
        m = fmaxf(m, sMax[(half ^ 1) * 128 + row]);
        const float m_use = (m == -INFINITY) ? 0.f : m;
        const float msc = m_use * sl2;
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_s + lane_base + c * 32, v);
            ptx::tmem_ld_wait();
            uint32_t pk[16];
            uint32_t keep = 0xffffffffu;
            if (DROPOUT) {   // keep bits of this row's 32 keys: drawn here, stored for the backward pass
                const long long mrow = ((long long)b * p.H + h) * p.Lq + qrow;
                keep = attn_keep32(rng, drop_offset, mrow, p.mask_ld, (k0 >> 5) + c, p.thresh16);
                if (qrow < p.Lq && (k0 >> 5) + c < p.mask_ld) p.mask[mrow * p.mask_ld + (k0 >> 5) + c] = keep;
            }
            if (c * 32 + 32 <= lim_lo) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float p0 = ex2(fmaf(__uint_as_float(v[2 * j]), sl2, -msc));
                    float p1 = ex2(fmaf(__uint_as_float(v[2 * j + 1]), sl2, -msc));
                    sum0 += p0; sum1 += p1;          // the softmax denominator ignores dropout
                    if (DROPOUT) { p0 = (keep >> (2 * j)) & 1u ? p0 : 0.f; p1 = (keep >> (2 * j + 1)) & 1u ? p1 : 0.f; }
                    pk[j] = pack_bf16x2(p0, p1);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = c * 32 + 2 * j;
                    float p0 = col < lim ? ex2(fmaf(__uint_as_float(v[2 * j]), sl2, -msc)) : 0.f;
                    float p1 = col + 1 < lim ? ex2(fmaf(__uint_as_float(v[2 * j + 1]), sl2, -msc)) : 0.f;
                    sum0 += p0; sum1 += p1;
                    if (DROPOUT) { p0 = (keep >> (2 * j)) & 1u ? p0 : 0.f; p1 = (keep >> (2 * j + 1)) & 1u ? p1 : 0.f; }
                    pk[j] = pack_bf16x2(p0, p1);
                }
            }
            if (P_TMEM) {
                // in place, inside the S columns this warp has already consumed (chunk c -> 16 columns)
                ptx::tmem_st_32x16(tmem_s + lane_base + half * 128 + (c - c_begin) * 16, pk);
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int c8 = c * 4 + g;   // 16-byte chunk (8 keys) of this row
                    *reinterpret_cast<uint4*>(sP + (c8 >> 3) * 16384 + row * 128 + (((c8 & 7) ^ (row & 7)) << 4)) =
                        make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                }
            }
        }
        const float psum = sum0 + sum1;
        sSum[half * 128 + row] = psum;
        if (P_TMEM) ptx::tmem_st_wait(); else ptx::fence_proxy_async();
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar_p);
        if (warp == 0) TSTAMP(11);
        ptx::bar_sync(1, 256);
        const float sum = psum + sSum[(half ^ 1) * 128 + row];

        ptx::mbar_wait(bar_o, 0);
        ptx::tc_fence_after();
        if (warp == 0) TSTAMP(12);
        const float inv = sum > 0.f ? (DROPOUT ? p.keep_scale : 1.f) / sum : 0.f;
        // one key block: the output itself; several: this block's softmax-normalised partial (merged by attn_combine_kernel)
        bf16* orow = p.n_kc == 1 ? p.o + ((long long)b * p.Lq + qrow) * p.ldo + h * 64 + half * 32   // this warp: 32 of the 64 columns
                                 : p.o_part + (((long long)kc * p.B + b) * p.Lq + qrow) * (p.H * 64) + h * 64 + half * 32;
        {
            uint32_t v[32], w[32];
            ptx::tmem_ld_32x32(tmem_o + lane_base + half * 32, v);
            if (n1 > 16) ptx::tmem_ld_32x32(tmem_o2 + lane_base + half * 32, w);   // warp-uniform: odd steps exist
            ptx::tmem_ld_wait();
            if (qrow < p.Lq) {
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = (__uint_as_float(v[j]) + (n1 > 16 ? __uint_as_float(w[j]) : 0.f)) * inv;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 ov;
                    ov.x = pack_bf16x2(f[8 * g], f[8 * g + 1]); ov.y = pack_bf16x2(f[8 * g + 2], f[8 * g + 3]);
                    ov.z = pack_bf16x2(f[8 * g + 4], f[8 * g + 5]); ov.w = pack_bf16x2(f[8 * g + 6], f[8 * g + 7]);
                    *reinterpret_cast<uint4*>(orow + g * 8) = ov;
                }
            }
        }
        if (half == 0 && qrow < p.Lq) {
            float* lse_out = p.n_kc == 1 ? p.lse : p.lse_part + (long long)kc * p.B * p.H * p.Lq;
            if (lse_out) lse_out[((long long)b * p.H + h) * p.Lq + qrow] = sum > 0.f ? m_use * p.scale + logf(sum) : -INFINITY;
        }
        ptx::tc_fence_before();
        if (warp == 0) TSTAMP(13);
    }
    __syncthreads();
    if (warp == 8) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
        TSTAMP(6);
    }
}

// =====================================================================================================
// backward
// =====================================================================================================
struct BwdCfg {
    static constexpr int T = 256 * 128;                 // one [256 x 64] bf16 operand
    static constexpr int PT = 128 * 256;                // one [128 x 128] bf16 tile (two 64-column blocks)
    static constexpr int SMEM_BYTES = 4 * T + 2 * PT + 4 * 256 * 4 /* lse, delta: two items */ + 1024 /*align*/ + 128 /*barriers*/;
    // TMEM columns: S^T [0,128) dP^T [128,256) dV [256,320) dK [320,384) dQ chunk i [384 + 64 i, ...)
};

// PERSISTENT: one CTA per SM walks the (head, batch, query block, key block) items of the launch. Round 1 ran one CTA per
// item and measured ~30 % of each CTA's life waiting for its 192 KB of operands (all CTAs of a wave load in lock-step;
// shared memory and TMEM leave no room for a second resident CTA). Here a dedicated TMA warp refills every operand buffer
// as soon as the LAST MMA that reads it has retired (tcgen05.commit on a "buffer free" barrier): K_0 / V_0 of the next
// item are requested after the current item's first key tile is finished, Q_0 / dO_0 after its second-to-last iteration,
// so the next item's first S^T / dP^T products start the moment the current item's last chains are issued.
// Operand tiles are 128 rows; load groups: (K_0, Q_0), (V_0, dO_0), (Q_1, dO_1), (K_1, V_1).
// BWD_SW softmax / drain warps (TMEM lane quarter = warp & 3, column part = warp >> 2), then 1 MMA warp and 1 TMA warp.
// 16 softmax warps were measured (64 x 8 heads x 250 x 250: 100.4 us against 96.3 us with 8): the elementwise stage between the two
// groups of MMAs is not what bounds an iteration — the serial chain MMA -> commit -> tcgen05.ld -> exp / dS -> shared-memory
// stores -> fence -> MMA is, and more warps only add barrier traffic.
static constexpr int BWD_SW = 8;
static constexpr int BWD_PARTS = BWD_SW / 4;
static constexpr int BWD_ST = BWD_SW * 32;      // softmax threads
static constexpr int BWD_THREADS = BWD_ST + 64;

template <bool CAUSAL, bool DROPOUT>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                   const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmdo,
                   const AttnTcDev p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sdO = sQ + BwdCfg::T;
    uint8_t* sK = sdO + BwdCfg::T;
    uint8_t* sV = sK + BwdCfg::T;
    uint8_t* sPT = sV + BwdCfg::T;
    uint8_t* sdST = sPT + BwdCfg::PT;
    float* sLse = reinterpret_cast<float*>(sdST + BwdCfg::PT);
    float* sDelta = sLse + 512;       // [2][256] each: the current item's values and the next item's
    uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 512);
    uint64_t* bar_s = bars, *bar_p = bars + 1, *bar_acc = bars + 2;
    uint64_t* bar_ld = bars + 3;     // [4] operand group landed: (K_0, Q_0), (V_0, dO_0), (Q_1, dO_1), (K_1, V_1)
    uint64_t* bar_free = bars + 7;   // [4] buffers reusable: (K_0, V_0), (Q_0, dO_0), (Q_1, dO_1), (K_1, V_1)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pairs = p.n_qb * p.n_kb;
    const int total_items = p.H * p.B * pairs;
    const bool part = pairs > 1;

    if (warp == BWD_SW) {
        if (lane == 0) {
            ptx::prefetch_tmap(&tmq); ptx::prefetch_tmap(&tmk); ptx::prefetch_tmap(&tmv); ptx::prefetch_tmap(&tmdo);
            ptx::mbar_init(bar_s, 1); ptx::mbar_init(bar_p, BWD_ST); ptx::mbar_init(bar_acc, 1);
            for (int i = 0; i < 4; ++i) { ptx::mbar_init(&bar_ld[i], 1); ptx::mbar_init(&bar_free[i], 1); }
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_ptr, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();   // set-up above overlaps the previous kernel; global memory is touched only below
    const uint32_t t_s = tmem_base, t_dp = tmem_base + 128, t_dv = tmem_base + 256, t_dk = tmem_base + 320,
                   t_dq = tmem_base + 384;

    // item -> (head, batch, query block, key block) and the block's extents
    struct Item { int h, b, qi, kj, q0, k0, q_box, kv_box, ni, nj; };
    auto decode = [&](int item) {
        Item w;
        const int pair = item % pairs, hb = item / pairs;
        w.h = hb % p.H; w.b = hb / p.H;
        w.qi = pair / p.n_kb; w.kj = pair - w.qi * p.n_kb;
        w.q0 = w.qi * 256; w.k0 = w.kj * 256;
        w.q_box = min(256, ((p.Lq - w.q0 + 15) >> 4) << 4);
        w.kv_box = min(256, ((p.Lk - w.k0 + 15) >> 4) << 4);
        w.ni = (w.q_box + 127) >> 7; w.nj = (w.kv_box + 127) >> 7;
        return w;
    };

    if (warp == BWD_SW + 1) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const uint32_t qb = (uint32_t)p.q_tma_rows * 128u, kb = (uint32_t)p.kv_tma_rows * 128u;   // bytes per tile box
            uint32_t n = 0, n2 = 0, n3 = 0;   // items so far / items so far that used load group 2 / 3 (barrier phases)
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++n) {
                const Item w = decode(item);
                const uint32_t fpar = (n & 1) ^ 1;     // "free" barriers: the first pass falls through (fresh barriers)
                ptx::mbar_wait(&bar_free[0], fpar);    // K_0, V_0
                ptx::mbar_wait(&bar_free[1], fpar);    // Q_0, dO_0
                ptx::mbar_arrive_expect_tx(&bar_ld[0], qb + kb);
                ptx::tma_load_3d(&tmk, &bar_ld[0], sK, w.h * 64, w.k0, w.b);
                ptx::tma_load_3d(&tmq, &bar_ld[0], sQ, w.h * 64, w.q0, w.b);
                ptx::mbar_arrive_expect_tx(&bar_ld[1], qb + kb);
                ptx::tma_load_3d(&tmv, &bar_ld[1], sV, w.h * 64, w.k0, w.b);
                ptx::tma_load_3d(&tmdo, &bar_ld[1], sdO, w.h * 64, w.q0, w.b);
                ptx::mbar_wait(&bar_free[2], fpar);    // Q_1, dO_1
                if (w.ni > 1) {
                    ptx::mbar_arrive_expect_tx(&bar_ld[2], 2 * qb);
                    ptx::tma_load_3d(&tmq, &bar_ld[2], sQ + 16384, w.h * 64, w.q0 + 128, w.b);
                    ptx::tma_load_3d(&tmdo, &bar_ld[2], sdO + 16384, w.h * 64, w.q0 + 128, w.b);
                    ++n2;
                }
                ptx::mbar_wait(&bar_free[3], fpar);    // K_1, V_1
                if (w.nj > 1) {
                    ptx::mbar_arrive_expect_tx(&bar_ld[3], 2 * kb);
                    ptx::tma_load_3d(&tmk, &bar_ld[3], sK + 16384, w.h * 64, w.k0 + 128, w.b);
                    ptx::tma_load_3d(&tmv, &bar_ld[3], sV + 16384, w.h * 64, w.k0 + 128, w.b);
                    ++n3;
                }
            }
            (void)n2; (void)n3;
        }
    } else if (warp == BWD_SW) {
        // ===================== MMA issuer (whole warp, warp-uniform; one elected lane issues) =====================
        {
            const bool leader = ptx::elect_one();
            const uint32_t idesc_cd = ptx::make_idesc_bf16(128, 64, false, true);
            const uint32_t idesc_e = ptx::make_idesc_bf16(128, 64, true, true);
            uint32_t it = 0, n = 0, n2 = 0, n3 = 0;   // iterations / items / uses of load groups 2 and 3 so far (barrier phases)
#ifdef PB_ATTN_TIMING
            long long tm[16][3]; const long long tm0 = clock64();
#endif
            // S^T = K_j Q_i^T and dP^T = V_j dO_i^T of iteration (j, i). Independent chains are interleaved step by step: consecutive
            // tcgen05.mma into the same accumulator serialise on its latency when N is small (measured).
            auto issue_scores = [&](const Item& w, int j, int i, bool first_of_item, uint32_t lpar) {
                const int nq = min(128, w.q_box - 128 * i);     // query columns of this chunk inside the box
                const uint32_t idesc_ab = ptx::make_idesc_bf16(128, nq, false, false);
                const uint64_t kd = kdesc(sK + j * 16384, 0), qd = kdesc(sQ + i * 16384, 0);
                const uint64_t vd = kdesc(sV + j * 16384, 0), dod = kdesc(sdO + i * 16384, 0);
                if (first_of_item) {
                    ptx::mbar_wait(&bar_ld[0], lpar);
                    ptx::tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (leader) ptx::umma_bf16_lh(t_s, lo32(kd) + 2 * k, hi32(kd), lo32(qd) + 2 * k, hi32(qd), idesc_ab, k > 0 ? 1u : 0u);
                    ptx::mbar_wait(&bar_ld[1], lpar);
                    ptx::tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (leader) ptx::umma_bf16_lh(t_dp, lo32(vd) + 2 * k, hi32(vd), lo32(dod) + 2 * k, hi32(dod), idesc_ab, k > 0 ? 1u : 0u);
                } else {
                    if (j == 0 && i == 1) { ptx::mbar_wait(&bar_ld[2], n2 & 1); ++n2; ptx::tc_fence_after(); }   // Q_1, dO_1
                    if (j == 1 && i == 0) { ptx::mbar_wait(&bar_ld[3], n3 & 1); ++n3; ptx::tc_fence_after(); }   // K_1, V_1
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (leader) ptx::umma_bf16_lh(t_s, lo32(kd) + 2 * k, hi32(kd), lo32(qd) + 2 * k, hi32(qd), idesc_ab, k > 0 ? 1u : 0u);
                        if (leader) ptx::umma_bf16_lh(t_dp, lo32(vd) + 2 * k, hi32(vd), lo32(dod) + 2 * k, hi32(dod), idesc_ab, k > 0 ? 1u : 0u);
                    }
                }
                if (leader) ptx::umma_commit(bar_s);
            };
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++n) {
                const Item w = decode(item);
                const uint32_t lpar = n & 1;
                issue_scores(w, 0, 0, true, lpar);
                for (int j = 0; j < w.nj; ++j) {
                    const int kc = min(128, w.kv_box - 128 * j);       // key rows of this tile inside the TMA box
                    for (int i = 0; i < w.ni; ++i, ++it) {
                        const int nq = min(128, w.q_box - 128 * i);
                        const bool last = (j == w.nj - 1 && i == w.ni - 1);
#ifdef PB_ATTN_TIMING
                        if (it < 16) tm[it][0] = clock64() - tm0;
#endif
                        ptx::mbar_wait(bar_p, it & 1);     // P^T / dS^T of this iteration are in shared memory; S^T / dP^T have been read
                        ptx::tc_fence_after();
#ifdef PB_ATTN_TIMING
                        if (it < 16) tm[it][1] = clock64() - tm0;
#endif
                        // The NEXT iteration's scores go first: the softmax warps then work on them while this iteration's three
                        // gradient chains are still being issued and executed (one thread issues ~75 cycles per MMA: the 24 of them
                        // were 1.8 k of every iteration's 5.6 k cycles, in-kernel timeline). Not across items: the next item's
                        // operand loads wait for buffers that this item's last gradient chains release.
                        if (!last) {
                            const int jn = (i + 1 < w.ni) ? j : j + 1, in = (i + 1 < w.ni) ? i + 1 : 0;
                            issue_scores(w, jn, in, false, lpar);
                        }
                        const int qsteps = nq >> 4, ksteps = kc >> 4;
                        const uint64_t ptd = kdesc(sPT, 0), dstd = kdesc(sdST, 0), dst_mn = mndesc(sdST, 0, 16384);
                        const uint64_t do_mn = mndesc(sdO + i * 16384, 0, 8192), q_mn = mndesc(sQ + i * 16384, 0, 8192);
                        const uint64_t k_mn = mndesc(sK + j * 16384, 0, 8192);
                        const uint32_t acc_i = i > 0 ? 1u : 0u, acc_j = j > 0 ? 1u : 0u;
                        const uint32_t t_dqi = t_dq + 64 * i;
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) {   // the three chains interleaved
                            if (kk < qsteps) {
                                // dV_j += P^T dO_i ; dK_j += dS^T Q_i
                                if (leader) ptx::umma_bf16_lh(t_dv, lo32(ptd) + kstep16(kk), hi32(ptd), lo32(do_mn) + 128 * kk, hi32(do_mn), idesc_cd,
                                                  kk > 0 ? 1u : acc_i);
                                if (leader) ptx::umma_bf16_lh(t_dk, lo32(dstd) + kstep16(kk), hi32(dstd), lo32(q_mn) + 128 * kk, hi32(q_mn), idesc_cd,
                                                  kk > 0 ? 1u : acc_i);
                            }
                            if (kk < ksteps)   // dQ_i += dS K_j   (A = dS^T read MN-major)
                                if (leader) ptx::umma_bf16_lh(t_dqi, lo32(dst_mn) + 128 * kk, hi32(dst_mn), lo32(k_mn) + 128 * kk, hi32(k_mn), idesc_e,
                                                  kk > 0 ? 1u : acc_j);
                        }
                        if (leader) ptx::umma_commit(bar_acc);
#ifdef PB_ATTN_TIMING
                        if (it < 16) tm[it][2] = clock64() - tm0;
#endif
                        // operand buffers whose last reader has just been issued go back to the TMA warp (the commit arrives
                        // when these MMAs have retired); buffers this item never used are released with the last iteration
                        if (leader && j == 0 && i == w.ni - 1) ptx::umma_commit(&bar_free[0]);                 // K_0, V_0
                        if (leader && j == w.nj - 1 && i == 0) ptx::umma_commit(&bar_free[1]);                 // Q_0, dO_0
                        if (leader && ((w.ni > 1 && j == w.nj - 1 && i == 1) || (w.ni == 1 && last))) ptx::umma_commit(&bar_free[2]);   // Q_1, dO_1
                        if (leader && ((w.nj > 1 && j == 1 && i == w.ni - 1) || (w.nj == 1 && last))) ptx::umma_commit(&bar_free[3]);   // K_1, V_1
                    }
                }
            }
            // the last item's "buffer free" arrivals are asynchronous (tcgen05.commit) and have no consumer: they must have
            // landed in this CTA's shared memory before the CTA may exit and the memory is handed to another block
            if (n > 0)
                for (int g = 0; g < 4; ++g) ptx::mbar_wait(&bar_free[g], (n - 1) & 1);
#ifdef PB_ATTN_TIMING
            if (blockIdx.x == 0 && leader)
                for (uint32_t i = 0; i < it && i < 16; ++i)
                    printf("mma it %u: S/dP issued+commit %lld  P ready %lld  dV/dK/dQ issued+commit %lld\n", i, tm[i][0], tm[i][1], tm[i][2]);
#endif
        }
    } else {
        // ===================== softmax / drain warps =====================
        // warps w, w + 4, ... own the same 32 key rows (TMEM lanes) and split the query columns / output columns
        const int quarter = warp & 3, part_id = warp >> 2;
        const int t = quarter * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const float sl2 = p.scale * LOG2E_F;
        uint32_t it = 0;
#ifdef PB_ATTN_TIMING
        long long ts[16][4]; const long long ts0 = clock64();
#endif
        // log-sum-exp (log2 units) and delta = rowsum(dO o O) of one query row per thread for item `w`, into lse / delta buffer
        // `which`. Called for the NEXT item right after the current item's last softmax stage, so that the global-load latency
        // sits under the last gradient MMAs instead of between two items (it was ~5 k of the ~19 k idle cycles per item).
        auto item_prologue = [&](const Item& w, int which) {
            if (threadIdx.x < 256) {
                const int r = threadIdx.x;   // 0..255: query q0 + r
                const int Lq_rows = min(256, p.Lq - w.q0);
                float l2 = INFINITY, dl = 0.f;
                if (r < Lq_rows) {
                    l2 = p.lse_in[((long long)w.b * p.H + w.h) * p.Lq + w.q0 + r] * LOG2E_F;
                    const uint4* po = reinterpret_cast<const uint4*>(p.o_in + ((long long)w.b * p.Lq + w.q0 + r) * p.ldo + w.h * 64);
                    const uint4* pd = reinterpret_cast<const uint4*>(p.dout + ((long long)w.b * p.Lq + w.q0 + r) * p.lddo + w.h * 64);
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const uint4 a = __ldg(po + g), d = __ldg(pd + g);
                        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 af = unpack_bf16x2(aw[e]), df = unpack_bf16x2(dw[e]);
                            dl = fmaf(af.x, df.x, dl); dl = fmaf(af.y, df.y, dl);
                        }
                    }
                }
                sLse[which * 256 + r] = l2; sDelta[which * 256 + r] = dl;
            }
        };
        int buf = 0;
        if ((int)blockIdx.x < total_items) item_prologue(decode(blockIdx.x), 0);
        ptx::bar_sync(1, BWD_ST);
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, buf ^= 1) {
            const Item w = decode(item);
            const int h = w.h, b = w.b, q0 = w.q0, k0 = w.k0, qi = w.qi, kj = w.kj, ni = w.ni, nj = w.nj;
            const int Lq_blk = min(256, p.Lq - q0), Lk_blk = min(256, p.Lk - k0);   // rows of this block that exist
            const int kvlen = (p.kv_len ? min(p.kv_len[b], p.Lk) : p.Lk) - k0;   // block-local: keys [0, kvlen) are unmasked
            const int coff = p.Lk - p.Lq + q0 - k0;                               // causal: local key j visible to local query i iff j <= i + coff
            // lse / delta of this item were put into buffer `buf` while the previous item was finishing (or before the loop)
            const float* cLse = sLse + buf * 256;
            const float* cDelta = sDelta + buf * 256;
            const int next_item = item + gridDim.x;
            int local = 0;

            for (int j = 0; j < nj; ++j) {
                const int key = 128 * j + t;            // block-local key of this thread
                const bool rowok = key < kvlen;
                for (int i = 0; i < ni; ++i, ++it, ++local) {
                    const int nq = min(128, w.q_box - 128 * i);
                    const int nch = (nq + 31) >> 5;
                    // P^T / dS^T of the previous iteration may still be read by its gradient MMAs (issued AFTER this iteration's
                    // score MMAs): the first shared-memory store of this iteration waits for them, the arithmetic does not
                    bool wait_prev = local > 0;
#ifdef PB_ATTN_TIMING
                    if (it < 16) ts[it][0] = clock64() - ts0;
#endif
                    ptx::mbar_wait(bar_s, it & 1);
                    ptx::tc_fence_after();
#ifdef PB_ATTN_TIMING
                    if (it < 16) ts[it][1] = clock64() - ts0;
#endif
#pragma unroll 1
                    for (int c = part_id; c < nch; c += BWD_PARTS) {
                        uint32_t s[32], dp[32];
                        ptx::tmem_ld_32x32(t_s + lane_base + c * 32, s);
                        ptx::tmem_ld_32x32(t_dp + lane_base + c * 32, dp);
                        ptx::tmem_ld_wait();
                        const int qb = 128 * i + c * 32;
                        const float4* l4 = reinterpret_cast<const float4*>(cLse + qb);
                        const float4* d4 = reinterpret_cast<const float4*>(cDelta + qb);
                        // keep bits of (query qb + e, this thread's key): lane e fetches query qb + e's word of this warp's 32
                        // keys (the warp's keys share one word: key = k0 + 128 j + 32 quarter + lane), broadcast by shuffle below
                        uint32_t mword = 0xffffffffu;
                        if (DROPOUT) {
                            const int qrow = q0 + qb + lane;
                            const int word = (k0 + 128 * j + quarter * 32) >> 5;
                            mword = (qrow < p.Lq && word < p.mask_ld) ? p.mask[(((long long)b * p.H + h) * p.Lq + qrow) * p.mask_ld + word] : 0u;
                        }
                        uint4 pw4[4], dw4[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {   // 8 queries = one 16-byte chunk of P^T and of dS^T
                            const float4 la = l4[2 * g], lb = l4[2 * g + 1], da = d4[2 * g], db = d4[2 * g + 1];
                            const float lv[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
                            const float dv[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
                            float pe[8], ds[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                pe[e] = ex2(fmaf(__uint_as_float(s[8 * g + e]), sl2, -lv[e]));
                                float dpe = __uint_as_float(dp[8 * g + e]);
                                bool kept = true;
                                if (DROPOUT) {      // d(P o M / (1 - p)) / dP: the dropped entries pass no gradient, the kept ones are scaled
                                    kept = (__shfl_sync(0xffffffffu, mword, 8 * g + e) >> lane) & 1u;
                                    dpe = kept ? dpe * p.keep_scale : 0.f;
                                }
                                ds[e] = pe[e] * (dpe - dv[e]);
                                if (DROPOUT) pe[e] = kept ? pe[e] * p.keep_scale : 0.f;   // P^T as it multiplied V in the forward pass (-> dV)
                                if (CAUSAL) {
                                    const bool ok = key <= qb + 8 * g + e + coff;
                                    pe[e] = ok ? pe[e] : 0.f;
                                    ds[e] = ok ? ds[e] : 0.f;
                                }
                            }
                            uint4 pw, dw;
                            pw.x = pack_bf16x2(pe[0], pe[1]); pw.y = pack_bf16x2(pe[2], pe[3]);
                            pw.z = pack_bf16x2(pe[4], pe[5]); pw.w = pack_bf16x2(pe[6], pe[7]);
                            dw.x = pack_bf16x2(ds[0], ds[1]); dw.y = pack_bf16x2(ds[2], ds[3]);
                            dw.z = pack_bf16x2(ds[4], ds[5]); dw.w = pack_bf16x2(ds[6], ds[7]);
                            if (!rowok) { pw = make_uint4(0, 0, 0, 0); dw = make_uint4(0, 0, 0, 0); }   // masked / unloaded key row
                            pw4[g] = pw; dw4[g] = dw;
                        }
                        if (wait_prev) {
                            ptx::mbar_wait(bar_acc, (it - 1) & 1);
                            wait_prev = false;
                        }
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int c8 = c * 4 + g;
                            const int off = (c8 >> 3) * 16384 + t * 128 + (((c8 & 7) ^ (t & 7)) << 4);
                            *reinterpret_cast<uint4*>(sPT + off) = pw4[g];
                            *reinterpret_cast<uint4*>(sdST + off) = dw4[g];
                        }
                    }
                    // every thread observes every phase of bar_acc, also the ones without a chunk in this iteration: a thread that
                    // skipped one would find the NEXT wait on it (parity of two phases later) already satisfied
                    if (wait_prev) ptx::mbar_wait(bar_acc, (it - 1) & 1);
#ifdef PB_ATTN_TIMING
                    if (it < 16) ts[it][2] = clock64() - ts0;
#endif
                    ptx::fence_proxy_async();
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(bar_p);
#ifdef PB_ATTN_TIMING
                    if (it < 16) ts[it][3] = clock64() - ts0;
#endif
                    if (j == nj - 1 && i == ni - 1 && next_item < total_items) item_prologue(decode(next_item), buf ^ 1);
                    if (i == ni - 1) {   // dV_j and dK_j are complete once this iteration's MMAs retire
                        ptx::mbar_wait(bar_acc, it & 1);
                        ptx::tc_fence_after();
                        // the column parts share the drain: dV then dK, 64 columns each, in slices of 128 / BWD_PARTS columns
                        // a single block pair writes the gradients themselves, otherwise this pair's partial sums
                        // (dk_part / dv_part [n_qb][B*Lk][H*64], added up by attn_reduce_kernel)
                        constexpr int SLICE = 128 / BWD_PARTS;            // 64 (8 warps) or 32 (16 warps) columns per warp
                        const int col_all = part_id * SLICE;              // position inside [dV | dK]
                        const bool is_dk = col_all >= 64;
                        const int col = col_all & 63;
                        const long long krow = (long long)b * p.Lk + k0 + key;
                        bf16* dst = (!part ? (!is_dk ? p.dv + krow * p.lddv + h * 64 : p.dk + krow * p.lddk + h * 64)
                                           : (!is_dk ? p.dv_part : p.dk_part) + ((long long)qi * p.B * p.Lk + krow) * (p.H * 64) + h * 64) + col;
                        const float mul = !is_dk ? 1.0f : p.scale;
                        const uint32_t src = (!is_dk ? t_dv : t_dk) + lane_base + col;
#pragma unroll
                        for (int hh = 0; hh < SLICE / 32; ++hh) {
                            uint32_t v[32];
                            ptx::tmem_ld_32x32(src + hh * 32, v);
                            ptx::tmem_ld_wait();
                            if (key < Lk_blk) {
#pragma unroll
                                for (int g = 0; g < 4; ++g) {
                                    uint4 ov;
                                    ov.x = pack_bf16x2(__uint_as_float(v[8 * g]) * mul, __uint_as_float(v[8 * g + 1]) * mul);
                                    ov.y = pack_bf16x2(__uint_as_float(v[8 * g + 2]) * mul, __uint_as_float(v[8 * g + 3]) * mul);
                                    ov.z = pack_bf16x2(__uint_as_float(v[8 * g + 4]) * mul, __uint_as_float(v[8 * g + 5]) * mul);
                                    ov.w = pack_bf16x2(__uint_as_float(v[8 * g + 6]) * mul, __uint_as_float(v[8 * g + 7]) * mul);
                                    *reinterpret_cast<uint4*>(dst + hh * 32 + g * 8) = ov;
                                }
                            }
                        }
                        ptx::tc_fence_before();
                    }
                }
            }
            // dQ: every (j, i) chain has retired (the last bar_acc wait above covers all earlier MMAs); each warp drains
            // 32 of the 64 columns of its rows. The next item's first dQ MMA is issued behind its first bar_p, i.e. after
            // this drain in every softmax thread's program order.
            // the (query chunk, 32-column slice) pairs are dealt round-robin to the column parts
            for (int u = part_id; u < 2 * ni; u += BWD_PARTS) {
                const int i = u >> 1, ch = u & 1;
                const int qrow = 128 * i + t;           // block-local query
                const long long grow = (long long)b * p.Lq + q0 + qrow;
                bf16* dst = (!part ? p.dq + grow * p.lddq : p.dq_part + ((long long)kj * p.B * p.Lq + grow) * (p.H * 64)) + h * 64 + ch * 32;
                uint32_t v[32];
                ptx::tmem_ld_32x32(t_dq + 64 * i + lane_base + ch * 32, v);
                ptx::tmem_ld_wait();
                if (qrow < Lq_blk) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 ov;
                        ov.x = pack_bf16x2(__uint_as_float(v[8 * g]) * p.scale, __uint_as_float(v[8 * g + 1]) * p.scale);
                        ov.y = pack_bf16x2(__uint_as_float(v[8 * g + 2]) * p.scale, __uint_as_float(v[8 * g + 3]) * p.scale);
                        ov.z = pack_bf16x2(__uint_as_float(v[8 * g + 4]) * p.scale, __uint_as_float(v[8 * g + 5]) * p.scale);
                        ov.w = pack_bf16x2(__uint_as_float(v[8 * g + 6]) * p.scale, __uint_as_float(v[8 * g + 7]) * p.scale);
                        *reinterpret_cast<uint4*>(dst + g * 8) = ov;
                    }
                }
            }
            ptx::tc_fence_before();
            ptx::bar_sync(1, BWD_ST);     // the next item's lse / delta (written above by every thread) are visible
        }
#ifdef PB_ATTN_TIMING
        if (blockIdx.x == 0 && threadIdx.x == 0)
            for (uint32_t i = 0; i < it && i < 16; ++i)
                printf("softmax it %u: reach wait %lld  S ready %lld  math done %lld  arrived %lld\n", i, ts[i][0], ts[i][1], ts[i][2], ts[i][3]);
#endif
    }
    __syncthreads();
    if (warp == BWD_SW) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

// =====================================================================================================
// merging the per-block partials of long sequences
// =====================================================================================================
// forward: out = sum_c w_c O_c with w_c = exp(lse_c - lse), lse = log sum_c exp(lse_c); one thread per (row, head, 8 columns)
__global__ void attn_combine_kernel(const bf16* __restrict__ o_part, const float* __restrict__ lse_part, int n_kc, int B, int H,
                                    int Lq, bf16* __restrict__ o, long long ldo, float* __restrict__ lse) {
    pdl_wait();
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = (long long)B * Lq * H * 8;
    if (idx >= total) return;
    const int g = (int)(idx & 7), h = (int)((idx >> 3) % H);
    const long long row = idx / (8LL * H);             // b * Lq + q
    const int b = (int)(row / Lq), q = (int)(row - (long long)b * Lq);
    const long long lidx = ((long long)b * H + h) * Lq + q, lstride = (long long)B * H * Lq;
    float m = -INFINITY;
    for (int c = 0; c < n_kc; ++c) m = fmaxf(m, lse_part[c * lstride + lidx]);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float den = 0.f;
    if (m > -INFINITY) {
        for (int c = 0; c < n_kc; ++c) {
            const float w = __expf(lse_part[c * lstride + lidx] - m);
            den += w;
            const uint4 v = *reinterpret_cast<const uint4*>(o_part + ((long long)c * B * Lq + row) * (H * 64) + h * 64 + g * 8);
            const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = unpack_bf16x2(vw[e]);
                acc[2 * e] = fmaf(w, f.x, acc[2 * e]);
                acc[2 * e + 1] = fmaf(w, f.y, acc[2 * e + 1]);
            }
        }
    }
    const float inv = den > 0.f ? 1.f / den : 0.f;
    uint4 ov;
    ov.x = pack_bf16x2(acc[0] * inv, acc[1] * inv); ov.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
    ov.z = pack_bf16x2(acc[4] * inv, acc[5] * inv); ov.w = pack_bf16x2(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(o + row * ldo + h * 64 + g * 8) = ov;
    if (g == 0 && lse) lse[lidx] = den > 0.f ? m + logf(den) : -INFINITY;
}

// backward: out[row, :] = sum_c part[c][row, :] (fp32 accumulation); one thread per (row, 8 columns)
__global__ void attn_reduce_kernel(const bf16* __restrict__ part, int n, long long rows, int cols, bf16* __restrict__ out,
                                   long long ldout) {
    pdl_wait();
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int c8 = cols >> 3;
    if (idx >= rows * c8) return;
    const long long row = idx / c8;
    const int g = (int)(idx - row * c8);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < n; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(part + ((long long)c * rows + row) * cols + g * 8);
        const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_bf16x2(vw[e]);
            acc[2 * e] += f.x; acc[2 * e + 1] += f.y;
        }
    }
    uint4 ov;
    ov.x = pack_bf16x2(acc[0], acc[1]); ov.y = pack_bf16x2(acc[2], acc[3]);
    ov.z = pack_bf16x2(acc[4], acc[5]); ov.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(out + row * ldout + g * 8) = ov;
}

// ------------------------------------------------------------------------------------------------ host side
static int round16(int x) { return ((x + 15) / 16) * 16; }
static int blocks256(int L) { return (L + 255) / 256; }

int attn_tc_check(const pb200_attn_params* q, bool backward) {
    PB_CHECK(q->head_dim == 64 && q->Lk >= 1 && q->Lq >= 1, "attention: head_dim 64, non-empty sequences");
    auto al = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    PB_CHECK(al(q->q) && al(q->k) && al(q->v) && al(q->o), "attention: q / k / v / o must be 16-byte aligned");
    PB_CHECK(!backward || (al(q->dout) && al(q->dq) && al(q->dk) && al(q->dv)), "attention: dout / dq / dk / dv must be 16-byte aligned");
    PB_CHECK(q->dropout_p >= 0.f && q->dropout_p < 1.f, "attention: dropout_p in [0, 1)");
    PB_CHECK(q->dropout_p == 0.f || q->drop_mask, "attention dropout needs the keep-bit buffer (drop_mask)");
    return 0;
}

// bytes of partial results for sequences longer than one 256-key (backward: 256-query) block
size_t attn_tc_workspace_bytes(const pb200_attn_params* q, bool backward) {
    const size_t D = (size_t)q->H * 64;
    const size_t nq = blocks256(q->Lq), nk = blocks256(q->Lk);
    if (!backward) return nk > 1 ? nk * ((size_t)q->B * q->Lq * D * 2 + (size_t)q->B * q->H * q->Lq * 4) : 0;
    if (nq * nk == 1) return 0;
    // every block pair writes partials (also where only one of the two sums has several terms: one code path)
    return nk * (size_t)q->B * q->Lq * D * 2 + 2 * nq * (size_t)q->B * q->Lk * D * 2;
}

static AttnTcDev tc_dev(const pb200_attn_params* q) {
    AttnTcDev p = {};
    p.o = (bf16*)q->o; p.lse = q->lse; p.kv_len = q->kv_len; p.ldo = q->ldo;
    p.B = q->B; p.H = q->H; p.Lq = q->Lq; p.Lk = q->Lk; p.causal = q->causal; p.scale = q->scale;
    p.n_kc = blocks256(q->Lk); p.n_qb = blocks256(q->Lq); p.n_kb = p.n_kc;
    p.kv_box = p.n_kc > 1 ? 256 : round16(q->Lk);
    p.mask = q->drop_mask; p.mask_ld = (q->Lk + 31) / 32;
    p.thresh16 = dropout_thresh16(q->dropout_p);
    p.keep_scale = q->dropout_p > 0.f ? 1.f / (1.f - q->dropout_p) : 1.f;
    p.seed = q->seed; p.offset = q->offset; p.offset_base = (const long long*)q->offset_base;
    return p;
}

template <bool CAUSAL, bool DROPOUT>
static int launch_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcDev& p,
                      cudaStream_t stream) {
    auto kern = attn_fwd_tc_kernel<true, CAUSAL, DROPOUT>;
    static bool attr_set = false;
    if (!attr_set) {
        PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdCfg<true>::SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid(((p.Lq + 127) / 128) * p.n_kc, p.H, p.B);
    PB_CUDA(launch_pdl(kern, grid, dim3(TC_THREADS), FwdCfg<true>::SMEM_BYTES, stream, tq, tk, tv, p));
    PB_LAUNCH_CHECK();
    return 0;
}

int attn_fwd_tc(const pb200_attn_params* q, cudaStream_t stream) {
    if (int rc = attn_tc_check(q, false)) return rc;
    AttnTcDev p = tc_dev(q);
    p.q_box = 128;
    if (p.n_kc > 1) {
        PB_CHECK(q->workspace && q->workspace_bytes >= attn_tc_workspace_bytes(q, false), "attention_fwd: workspace too small for Lk > 256");
        p.o_part = (bf16*)q->workspace;
        p.lse_part = (float*)((char*)q->workspace + (size_t)p.n_kc * q->B * q->Lq * q->H * 64 * 2);
    }
    CUtensorMap tq, tk, tv;
    if (int rc = make_tmap(&tq, q->q, q->Lq, (long long)q->H * 64, q->ldq, q->B, (long long)q->Lq * q->ldq, 64, 128)) return rc;
    if (int rc = make_tmap(&tk, q->k, q->Lk, (long long)q->H * 64, q->ldk, q->B, (long long)q->Lk * q->ldk, 64, p.kv_box)) return rc;
    if (int rc = make_tmap(&tv, q->v, q->Lk, (long long)q->H * 64, q->ldv, q->B, (long long)q->Lk * q->ldv, 64, p.kv_box)) return rc;
    const bool drop = q->dropout_p > 0.f;
    int rc;
    if (p.causal) rc = drop ? launch_fwd<true, true>(tq, tk, tv, p, stream) : launch_fwd<true, false>(tq, tk, tv, p, stream);
    else rc = drop ? launch_fwd<false, true>(tq, tk, tv, p, stream) : launch_fwd<false, false>(tq, tk, tv, p, stream);
    if (rc) return rc;
    if (p.n_kc > 1) {
        const long long total = (long long)q->B * q->Lq * q->H * 8;
        PB_CUDA(launch_pdl(attn_combine_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, (const bf16*)p.o_part,
                           (const float*)p.lse_part, p.n_kc, p.B, p.H, p.Lq, p.o, p.ldo, p.lse));
        PB_LAUNCH_CHECK();
    }
    return 0;
}

template <bool CAUSAL, bool DROPOUT>
static int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                      const AttnTcDev& p, cudaStream_t stream) {
    auto kern = attn_bwd_tc_kernel<CAUSAL, DROPOUT>;
    static bool attr_set = false;
    if (!attr_set) {
        PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdCfg::SMEM_BYTES));
        attr_set = true;
    }
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        PB_CUDA(cudaGetDevice(&dev));
        PB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long items = (long long)p.H * p.B * p.n_qb * p.n_kb;
    dim3 grid((unsigned)(items < sms ? items : sms));      // persistent: one CTA per SM walks the items
    PB_CUDA(launch_pdl(kern, grid, dim3(BWD_THREADS), BwdCfg::SMEM_BYTES, stream, tq, tk, tv, tdo, p));
    PB_LAUNCH_CHECK();
    return 0;
}

int attn_bwd_tc(const pb200_attn_params* q, cudaStream_t stream) {
    if (int rc = attn_tc_check(q, true)) return rc;
    AttnTcDev p = tc_dev(q);
    p.o_in = (const bf16*)q->o; p.dout = (const bf16*)q->dout; p.lddo = q->lddo; p.lse_in = q->lse;
    p.dq = (bf16*)q->dq; p.dk = (bf16*)q->dk; p.dv = (bf16*)q->dv;
    p.lddq = q->lddq; p.lddk = q->lddk; p.lddv = q->lddv;
    p.q = (const bf16*)q->q; p.k = (const bf16*)q->k; p.v = (const bf16*)q->v;
    p.ldq = q->ldq; p.ldk = q->ldk; p.ldv = q->ldv;
    // 128-row tile boxes (see the kernel: one barrier per pair of tiles); a single short block gets a box of its own size
    p.q_tma_rows = p.n_qb > 1 ? 128 : (round16(q->Lq) < 128 ? round16(q->Lq) : 128);
    p.kv_tma_rows = p.n_kb > 1 ? 128 : (round16(q->Lk) < 128 ? round16(q->Lk) : 128);
    const bool part = p.n_qb * p.n_kb > 1;
    const size_t D = (size_t)q->H * 64;
    if (part) {
        PB_CHECK(q->workspace && q->workspace_bytes >= attn_tc_workspace_bytes(q, true), "attention_bwd: workspace too small for > 256 positions");
        p.dq_part = (bf16*)q->workspace;
        p.dk_part = p.dq_part + (size_t)p.n_kb * q->B * q->Lq * D;
        p.dv_part = p.dk_part + (size_t)p.n_qb * q->B * q->Lk * D;
    }
    CUtensorMap tq, tk, tv, tdo;
    if (int rc = make_tmap(&tq, q->q, q->Lq, (long long)q->H * 64, q->ldq, q->B, (long long)q->Lq * q->ldq, 64, p.q_tma_rows)) return rc;
    if (int rc = make_tmap(&tdo, q->dout, q->Lq, (long long)q->H * 64, q->lddo, q->B, (long long)q->Lq * q->lddo, 64, p.q_tma_rows)) return rc;
    if (int rc = make_tmap(&tk, q->k, q->Lk, (long long)q->H * 64, q->ldk, q->B, (long long)q->Lk * q->ldk, 64, p.kv_tma_rows)) return rc;
    if (int rc = make_tmap(&tv, q->v, q->Lk, (long long)q->H * 64, q->ldv, q->B, (long long)q->Lk * q->ldv, 64, p.kv_tma_rows)) return rc;
    const bool drop = q->dropout_p > 0.f;
    int rc;
    if (p.causal) rc = drop ? launch_bwd<true, true>(tq, tk, tv, tdo, p, stream) : launch_bwd<true, false>(tq, tk, tv, tdo, p, stream);
    else rc = drop ? launch_bwd<false, true>(tq, tk, tv, tdo, p, stream) : launch_bwd<false, false>(tq, tk, tv, tdo, p, stream);
    if (rc) return rc;
    if (part) {
        auto reduce = [&](const bf16* src, int n, long long rows, bf16* dst, long long ld) -> int {
            const long long total = rows * (long long)(D / 8);
            PB_CUDA(launch_pdl(attn_reduce_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, src, n, rows, (int)D, dst, ld));
            PB_LAUNCH_CHECK();
            return 0;
        };
        // the outputs are column slices of wider matrices (ld > H * 64): reduce row by row into them
        if (int r = reduce(p.dq_part, p.n_kb, (long long)q->B * q->Lq, p.dq, p.lddq)) return r;
        if (int r = reduce(p.dk_part, p.n_qb, (long long)q->B * q->Lk, p.dk, p.lddk)) return r;
        if (int r = reduce(p.dv_part, p.n_qb, (long long)q->B * q->Lk, p.dv, p.lddv)) return r;
    }
    return 0;
}

}  // namespace pb200

#if defined(PB_ATTN_TIMING) && !defined(PB200_HALF)   // (debug export, not in the header: only the bfloat16 build defines it)
extern "C" int pb200_debug_attn_timing(long long* host_out, int n_ctas) {
    if (n_ctas > 4096) n_ctas = 4096;
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(host_out, pb200::g_attn_t, sizeof(long long) * 16 * n_ctas);
}
#endif
