// pasero_b200 — tcgen05/TMEM attention for sequences that fit one tile row (Lk <= 256; backward: Lq <= 256 too),
// head_dim 64, bf16 in / fp32 softmax. This is the path the ST training step runs on (250 encoder positions,
// 48 target positions); longer sequences use the online-softmax kernels in attention.cu.
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

namespace pb200 {

static constexpr float LOG2E_F = 1.4426950408889634f;
static constexpr int TC_THREADS = 160;   // 4 softmax/epilogue warps + 1 TMA/MMA warp

struct AttnTcDev {
    bf16* o; float* lse; const int* kv_len;
    long long ldo;
    int B, H, Lq, Lk, causal;
    int q_box, kv_box;          // rows per TMA box (multiples of 16)
    float scale;
    // backward
    const bf16* o_in; const bf16* dout; long long lddo;
    const float* lse_in;
    bf16* dq; bf16* dk; bf16* dv;
    long long lddq, lddk, lddv;
};

__device__ __forceinline__ uint64_t kdesc(const uint8_t* base, int k16) {           // K-major operand, k-step of 16
    return ptx::make_smem_desc_sw128(ptx::smem_u32(base) + k16 * 32, 16, 1024);
}
__device__ __forceinline__ uint64_t mndesc(const uint8_t* base, int k16, uint32_t lbo) {   // MN-major operand
    return ptx::make_smem_desc_sw128(ptx::smem_u32(base) + k16 * 2048, lbo, 1024);
}

// =====================================================================================================
// forward
// =====================================================================================================
template <bool P_TMEM>
struct FwdCfg {
    static constexpr int Q_BYTES = 128 * 128, KV_BYTES = 256 * 128, P_BYTES = P_TMEM ? 0 : 128 * 512;
    static constexpr int SMEM_BYTES = Q_BYTES + 2 * KV_BYTES + P_BYTES + 1024 /*align*/ + 64 /*barriers*/;
    static constexpr int TMEM_COLS = P_TMEM ? 256 : 512;
    static constexpr int O_COL = P_TMEM ? 128 : 256;
};

template <bool P_TMEM>
__global__ void __launch_bounds__(TC_THREADS, P_TMEM ? 2 : 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                   const __grid_constant__ CUtensorMap tmv, const AttnTcDev p) {
    using Cfg = FwdCfg<P_TMEM>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + Cfg::Q_BYTES;
    uint8_t* sV = sK + Cfg::KV_BYTES;
    uint8_t* sP = sV + Cfg::KV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
    uint64_t* bar_load = bars, *bar_s = bars + 1, *bar_p = bars + 2, *bar_o = bars + 3;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int kvlen = p.kv_len ? min(p.kv_len[b], p.Lk) : p.Lk;
    const int coff = p.Lk - p.Lq;   // causal: key j visible to query i iff j <= i + coff
    int kv_hi = p.causal ? min(kvlen, q0 + 128 + coff) : kvlen;
    kv_hi = max(kv_hi, 1);
    const int n1 = min(((kv_hi + 15) >> 4) << 4, p.kv_box);   // keys covered by the MMAs (multiple of 16)

    if (warp == 4) {
        if (lane == 0) {
            ptx::prefetch_tmap(&tmq); ptx::prefetch_tmap(&tmk); ptx::prefetch_tmap(&tmv);
            ptx::mbar_init(bar_load, 1); ptx::mbar_init(bar_s, 1); ptx::mbar_init(bar_p, 128); ptx::mbar_init(bar_o, 1);
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
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + Cfg::O_COL;

    if (warp == 4) {
        if (lane == 0) {
            ptx::mbar_arrive_expect_tx(bar_load, (uint32_t)(128 + 2 * p.kv_box) * 128u);
            ptx::tma_load_3d(&tmq, bar_load, sQ, h * 64, q0, b);
            ptx::tma_load_3d(&tmk, bar_load, sK, h * 64, 0, b);
            ptx::tma_load_3d(&tmv, bar_load, sV, h * 64, 0, b);
            ptx::mbar_wait(bar_load, 0);
            ptx::tc_fence_after();
            const uint32_t idesc1 = ptx::make_idesc_bf16(128, n1, false, false);
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem_s, kdesc(sQ, k), kdesc(sK, k), idesc1, k > 0 ? 1u : 0u);
            ptx::umma_commit(bar_s);
            ptx::mbar_wait(bar_p, 0);
            ptx::tc_fence_after();
            const uint32_t idesc2 = ptx::make_idesc_bf16(128, 64, false, true);
            const int ksteps = n1 >> 4;
            for (int kk = 0; kk < ksteps; ++kk) {
                const uint64_t bd = mndesc(sV, kk, 8192);
                if (P_TMEM) ptx::umma_bf16_ts(tmem_o, tmem_s + kk * 8, bd, idesc2, kk > 0 ? 1u : 0u);
                else ptx::umma_bf16(tmem_o, kdesc(sP + (kk >> 2) * 16384, kk & 3), bd, idesc2, kk > 0 ? 1u : 0u);
            }
            ptx::umma_commit(bar_o);
        }
    } else {
        const int row = warp * 32 + lane;
        const int qrow = q0 + row;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const int lim = p.causal ? min(kvlen, qrow + coff + 1) : kvlen;   // keys [0, lim) are visible to this row
        const float sl2 = p.scale * LOG2E_F;
        const int nch = (n1 + 31) >> 5;
        ptx::mbar_wait(bar_s, 0);
        ptx::tc_fence_after();
        float m = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_s + lane_base + c * 32, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c * 32 + j < lim) m = fmaxf(m, __uint_as_float(v[j]));
        }
        const float m_use = (m == -INFINITY) ? 0.f : m;
        const float msc = m_use * sl2;
        float sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_s + lane_base + c * 32, v);
            ptx::tmem_ld_wait();
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int col = c * 32 + 2 * j;
                const float p0 = col < lim ? exp2f(fmaf(__uint_as_float(v[2 * j]), sl2, -msc)) : 0.f;
                const float p1 = col + 1 < lim ? exp2f(fmaf(__uint_as_float(v[2 * j + 1]), sl2, -msc)) : 0.f;
                sum += p0 + p1;
                pk[j] = pack_bf16x2(p0, p1);
            }
            if (P_TMEM) {
                ptx::tmem_st_32x16(tmem_s + lane_base + c * 16, pk);   // in place: columns already consumed above
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int c8 = c * 4 + g;   // 16-byte chunk (8 keys) of this row
                    *reinterpret_cast<uint4*>(sP + (c8 >> 3) * 16384 + row * 128 + (((c8 & 7) ^ (row & 7)) << 4)) =
                        make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                }
            }
        }
        if (P_TMEM) ptx::tmem_st_wait(); else ptx::fence_proxy_async();
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar_p);

        ptx::mbar_wait(bar_o, 0);
        ptx::tc_fence_after();
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        bf16* orow = p.o + ((long long)b * p.Lq + qrow) * p.ldo + h * 64;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_o + lane_base + half * 32, v);
            ptx::tmem_ld_wait();
            if (qrow < p.Lq) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 ov;
                    ov.x = pack_bf16x2(__uint_as_float(v[8 * g]) * inv, __uint_as_float(v[8 * g + 1]) * inv);
                    ov.y = pack_bf16x2(__uint_as_float(v[8 * g + 2]) * inv, __uint_as_float(v[8 * g + 3]) * inv);
                    ov.z = pack_bf16x2(__uint_as_float(v[8 * g + 4]) * inv, __uint_as_float(v[8 * g + 5]) * inv);
                    ov.w = pack_bf16x2(__uint_as_float(v[8 * g + 6]) * inv, __uint_as_float(v[8 * g + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + half * 32 + g * 8) = ov;
                }
            }
        }
        if (p.lse && qrow < p.Lq)
            p.lse[((long long)b * p.H + h) * p.Lq + qrow] = sum > 0.f ? m_use * p.scale + logf(sum) : -INFINITY;
        ptx::tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// =====================================================================================================
// backward
// =====================================================================================================
struct BwdCfg {
    static constexpr int T = 256 * 128;                 // one [256 x 64] bf16 operand
    static constexpr int PT = 128 * 256;                // one [128 x 128] bf16 tile (two 64-column blocks)
    static constexpr int SMEM_BYTES = 4 * T + 2 * PT + 2 * 256 * 4 + 1024 /*align*/ + 64 /*barriers*/;
    // TMEM columns: S^T [0,128) dP^T [128,256) dV [256,320) dK [320,384) dQ chunk i [384 + 64 i, ...)
};

__global__ void __launch_bounds__(TC_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                   const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmdo,
                   const AttnTcDev p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sdO = sQ + BwdCfg::T;
    uint8_t* sK = sdO + BwdCfg::T;
    uint8_t* sV = sK + BwdCfg::T;
    uint8_t* sPT = sV + BwdCfg::T;
    uint8_t* sdST = sPT + BwdCfg::PT;
    float* sLse = reinterpret_cast<float*>(sdST + BwdCfg::PT);
    float* sDelta = sLse + 256;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 256);
    uint64_t* bar_load = bars, *bar_s = bars + 1, *bar_p = bars + 2, *bar_acc = bars + 3;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.x, b = blockIdx.y;
    const int kvlen = p.kv_len ? min(p.kv_len[b], p.Lk) : p.Lk;
    const int coff = p.Lk - p.Lq;
    const int ni = (p.q_box + 127) >> 7, nj = (p.kv_box + 127) >> 7;

    if (warp == 4) {
        if (lane == 0) {
            ptx::prefetch_tmap(&tmq); ptx::prefetch_tmap(&tmk); ptx::prefetch_tmap(&tmv); ptx::prefetch_tmap(&tmdo);
            ptx::mbar_init(bar_load, 1); ptx::mbar_init(bar_s, 1); ptx::mbar_init(bar_p, 128); ptx::mbar_init(bar_acc, 1);
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
    const uint32_t t_s = tmem_base, t_dp = tmem_base + 128, t_dv = tmem_base + 256, t_dk = tmem_base + 320,
                   t_dq = tmem_base + 384;

    if (warp == 4) {
        if (lane == 0) {
            ptx::mbar_arrive_expect_tx(bar_load, (uint32_t)(2 * p.q_box + 2 * p.kv_box) * 128u);
            ptx::tma_load_3d(&tmk, bar_load, sK, h * 64, 0, b);
            ptx::tma_load_3d(&tmq, bar_load, sQ, h * 64, 0, b);
            ptx::tma_load_3d(&tmv, bar_load, sV, h * 64, 0, b);
            ptx::tma_load_3d(&tmdo, bar_load, sdO, h * 64, 0, b);
            ptx::mbar_wait(bar_load, 0);
            ptx::tc_fence_after();
            const uint32_t idesc_cd = ptx::make_idesc_bf16(128, 64, false, true);
            const uint32_t idesc_e = ptx::make_idesc_bf16(128, 64, true, true);
            int it = 0;
            for (int j = 0; j < nj; ++j) {
                const int kc = min(128, p.kv_box - 128 * j);       // key rows of this tile inside the TMA box
                for (int i = 0; i < ni; ++i, ++it) {
                    const int nq = min(128, p.q_box - 128 * i);     // query columns of this chunk inside the box
                    const uint32_t idesc_ab = ptx::make_idesc_bf16(128, nq, false, false);
#pragma unroll
                    for (int k = 0; k < 4; ++k)     // S^T = K_j Q_i^T
                        ptx::umma_bf16(t_s, kdesc(sK + j * 16384, k), kdesc(sQ + i * 16384, k), idesc_ab, k > 0 ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < 4; ++k)     // dP^T = V_j dO_i^T
                        ptx::umma_bf16(t_dp, kdesc(sV + j * 16384, k), kdesc(sdO + i * 16384, k), idesc_ab, k > 0 ? 1u : 0u);
                    ptx::umma_commit(bar_s);
                    ptx::mbar_wait(bar_p, it & 1);
                    ptx::tc_fence_after();
                    const int qsteps = nq >> 4;
                    for (int kk = 0; kk < qsteps; ++kk)   // dV_j += P^T dO_i
                        ptx::umma_bf16(t_dv, kdesc(sPT + (kk >> 2) * 16384, kk & 3), mndesc(sdO + i * 16384, kk, 8192), idesc_cd,
                                       (i > 0 || kk > 0) ? 1u : 0u);
                    for (int kk = 0; kk < qsteps; ++kk)   // dK_j += dS^T Q_i
                        ptx::umma_bf16(t_dk, kdesc(sdST + (kk >> 2) * 16384, kk & 3), mndesc(sQ + i * 16384, kk, 8192), idesc_cd,
                                       (i > 0 || kk > 0) ? 1u : 0u);
                    const int ksteps = kc >> 4;
                    for (int kk = 0; kk < ksteps; ++kk)   // dQ_i += dS K_j   (A = dS^T read MN-major)
                        ptx::umma_bf16(t_dq + 64 * i, mndesc(sdST, kk, 16384), mndesc(sK + j * 16384, kk, 8192), idesc_e,
                                       (j > 0 || kk > 0) ? 1u : 0u);
                    ptx::umma_commit(bar_acc);
                }
            }
        }
    } else {
        const int t = warp * 32 + lane;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const float sl2 = p.scale * LOG2E_F;
        // log-sum-exp (in log2 units) and delta = rowsum(dO o O) of every query row, while the TMA loads are in flight
        for (int r = t; r < 256; r += 128) {
            float l2 = INFINITY, dl = 0.f;
            if (r < p.Lq) {
                l2 = p.lse_in[((long long)b * p.H + h) * p.Lq + r] * LOG2E_F;
                const uint4* po = reinterpret_cast<const uint4*>(p.o_in + ((long long)b * p.Lq + r) * p.ldo + h * 64);
                const uint4* pd = reinterpret_cast<const uint4*>(p.dout + ((long long)b * p.Lq + r) * p.lddo + h * 64);
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
            sLse[r] = l2; sDelta[r] = dl;
        }
        ptx::bar_sync(1, 128);

        int it = 0;
        for (int j = 0; j < nj; ++j) {
            const int key = 128 * j + t;
            for (int i = 0; i < ni; ++i, ++it) {
                const int nq = min(128, p.q_box - 128 * i);
                const int nch = (nq + 31) >> 5;
                ptx::mbar_wait(bar_s, it & 1);
                ptx::tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < nch; ++c) {
                    uint32_t s[32], dp[32];
                    ptx::tmem_ld_32x32(t_s + lane_base + c * 32, s);
                    ptx::tmem_ld_32x32(t_dp + lane_base + c * 32, dp);
                    ptx::tmem_ld_wait();
                    uint32_t pk[16], dk_[16];
                    const int qb = 128 * i + c * 32;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        float pv[2], ds[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int q = qb + 2 * e + u;
                            const bool ok = key < kvlen && q < p.Lq && (!p.causal || key <= q + coff);
                            const float pe = exp2f(fmaf(__uint_as_float(s[2 * e + u]), sl2, -sLse[q]));
                            pv[u] = ok ? pe : 0.f;
                            ds[u] = ok ? pe * (__uint_as_float(dp[2 * e + u]) - sDelta[q]) : 0.f;
                        }
                        pk[e] = pack_bf16x2(pv[0], pv[1]);
                        dk_[e] = pack_bf16x2(ds[0], ds[1]);
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int c8 = c * 4 + g;   // 16-byte chunk (8 queries) of this key row
                        const int off = (c8 >> 3) * 16384 + t * 128 + (((c8 & 7) ^ (t & 7)) << 4);
                        *reinterpret_cast<uint4*>(sPT + off) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                        *reinterpret_cast<uint4*>(sdST + off) = make_uint4(dk_[4 * g], dk_[4 * g + 1], dk_[4 * g + 2], dk_[4 * g + 3]);
                    }
                }
                ptx::fence_proxy_async();
                ptx::tc_fence_before();
                ptx::mbar_arrive(bar_p);
                if (i == ni - 1) {   // dV_j and dK_j are complete once this iteration's MMAs retire
                    ptx::mbar_wait(bar_acc, it & 1);
                    ptx::tc_fence_after();
#pragma unroll
                    for (int which = 0; which < 2; ++which) {
                        bf16* dst = which == 0 ? p.dv + ((long long)b * p.Lk + key) * p.lddv + h * 64
                                               : p.dk + ((long long)b * p.Lk + key) * p.lddk + h * 64;
                        const float mul = which == 0 ? 1.0f : p.scale;
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            uint32_t v[32];
                            ptx::tmem_ld_32x32((which == 0 ? t_dv : t_dk) + lane_base + half * 32, v);
                            ptx::tmem_ld_wait();
                            if (key < p.Lk) {
#pragma unroll
                                for (int g = 0; g < 4; ++g) {
                                    uint4 ov;
                                    ov.x = pack_bf16x2(__uint_as_float(v[8 * g]) * mul, __uint_as_float(v[8 * g + 1]) * mul);
                                    ov.y = pack_bf16x2(__uint_as_float(v[8 * g + 2]) * mul, __uint_as_float(v[8 * g + 3]) * mul);
                                    ov.z = pack_bf16x2(__uint_as_float(v[8 * g + 4]) * mul, __uint_as_float(v[8 * g + 5]) * mul);
                                    ov.w = pack_bf16x2(__uint_as_float(v[8 * g + 6]) * mul, __uint_as_float(v[8 * g + 7]) * mul);
                                    *reinterpret_cast<uint4*>(dst + half * 32 + g * 8) = ov;
                                }
                            }
                        }
                    }
                    ptx::tc_fence_before();
                }
            }
        }
        // dQ: every (j, i) chain has retired (the last bar_acc wait above covers all earlier MMAs)
        for (int i = 0; i < ni; ++i) {
            const int qrow = 128 * i + t;
            bf16* dst = p.dq + ((long long)b * p.Lq + qrow) * p.lddq + h * 64;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(t_dq + 64 * i + lane_base + half * 32, v);
                ptx::tmem_ld_wait();
                if (qrow < p.Lq) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 ov;
                        ov.x = pack_bf16x2(__uint_as_float(v[8 * g]) * p.scale, __uint_as_float(v[8 * g + 1]) * p.scale);
                        ov.y = pack_bf16x2(__uint_as_float(v[8 * g + 2]) * p.scale, __uint_as_float(v[8 * g + 3]) * p.scale);
                        ov.z = pack_bf16x2(__uint_as_float(v[8 * g + 4]) * p.scale, __uint_as_float(v[8 * g + 5]) * p.scale);
                        ov.w = pack_bf16x2(__uint_as_float(v[8 * g + 6]) * p.scale, __uint_as_float(v[8 * g + 7]) * p.scale);
                        *reinterpret_cast<uint4*>(dst + half * 32 + g * 8) = ov;
                    }
                }
            }
        }
        ptx::tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ host side
static int round16(int x) { return ((x + 15) / 16) * 16; }

bool attn_tc_supported(const pb200_attn_params* q, bool backward) {
    if (q->head_dim != 64 || q->Lk < 1 || q->Lk > 256 || q->Lq < 1) return false;
    if (backward && q->Lq > 256) return false;
    auto al = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    if (!al(q->q) || !al(q->k) || !al(q->v) || !al(q->o)) return false;
    if (backward && (!al(q->dout) || !al(q->dq) || !al(q->dk) || !al(q->dv))) return false;
    return true;
}

static AttnTcDev tc_dev(const pb200_attn_params* q) {
    AttnTcDev p = {};
    p.o = (bf16*)q->o; p.lse = q->lse; p.kv_len = q->kv_len; p.ldo = q->ldo;
    p.B = q->B; p.H = q->H; p.Lq = q->Lq; p.Lk = q->Lk; p.causal = q->causal; p.scale = q->scale;
    p.kv_box = round16(q->Lk);
    return p;
}

template <bool P_TMEM>
static int launch_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcDev& p,
                      cudaStream_t stream) {
    auto kern = attn_fwd_tc_kernel<P_TMEM>;
    static bool attr_set = false;
    if (!attr_set) {
        PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdCfg<P_TMEM>::SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid((p.Lq + 127) / 128, p.H, p.B);
    kern<<<grid, TC_THREADS, FwdCfg<P_TMEM>::SMEM_BYTES, stream>>>(tq, tk, tv, p);
    PB_LAUNCH_CHECK();
    return 0;
}

int attn_fwd_tc(const pb200_attn_params* q, cudaStream_t stream, bool p_in_tmem) {
    AttnTcDev p = tc_dev(q);
    p.q_box = 128;
    CUtensorMap tq, tk, tv;
    if (int rc = make_tmap(&tq, q->q, q->Lq, (long long)q->H * 64, q->ldq, q->B, (long long)q->Lq * q->ldq, 64, 128)) return rc;
    if (int rc = make_tmap(&tk, q->k, q->Lk, (long long)q->H * 64, q->ldk, q->B, (long long)q->Lk * q->ldk, 64, p.kv_box)) return rc;
    if (int rc = make_tmap(&tv, q->v, q->Lk, (long long)q->H * 64, q->ldv, q->B, (long long)q->Lk * q->ldv, 64, p.kv_box)) return rc;
    return p_in_tmem ? launch_fwd<true>(tq, tk, tv, p, stream) : launch_fwd<false>(tq, tk, tv, p, stream);
}

int attn_bwd_tc(const pb200_attn_params* q, cudaStream_t stream) {
    AttnTcDev p = tc_dev(q);
    p.q_box = round16(q->Lq);
    p.o_in = (const bf16*)q->o; p.dout = (const bf16*)q->dout; p.lddo = q->lddo; p.lse_in = q->lse;
    p.dq = (bf16*)q->dq; p.dk = (bf16*)q->dk; p.dv = (bf16*)q->dv;
    p.lddq = q->lddq; p.lddk = q->lddk; p.lddv = q->lddv;
    CUtensorMap tq, tk, tv, tdo;
    if (int rc = make_tmap(&tq, q->q, q->Lq, (long long)q->H * 64, q->ldq, q->B, (long long)q->Lq * q->ldq, 64, p.q_box)) return rc;
    if (int rc = make_tmap(&tdo, q->dout, q->Lq, (long long)q->H * 64, q->lddo, q->B, (long long)q->Lq * q->lddo, 64, p.q_box)) return rc;
    if (int rc = make_tmap(&tk, q->k, q->Lk, (long long)q->H * 64, q->ldk, q->B, (long long)q->Lk * q->ldk, 64, p.kv_box)) return rc;
    if (int rc = make_tmap(&tv, q->v, q->Lk, (long long)q->H * 64, q->ldv, q->B, (long long)q->Lk * q->ldv, 64, p.kv_box)) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        PB_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdCfg::SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid(p.H, p.B);
    attn_bwd_tc_kernel<<<grid, TC_THREADS, BwdCfg::SMEM_BYTES, stream>>>(tq, tk, tv, tdo, p);
    PB_LAUNCH_CHECK();
    return 0;
}

}  // namespace pb200
