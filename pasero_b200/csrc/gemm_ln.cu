// pasero_b200 — output projection / fc2 GEMM with the whole post-norm tail of the sub-block in its epilogue (d_model = 512):
//
//   x   = A W^T + bias                      (bf16, the Linear's output in the reference)
//   sum = residual + dropout(x)             (bf16, saved for backward)
//   y   = LayerNorm(sum) * gamma + beta     (bf16) ; mean / rstd (fp32) and the dropout keep bits saved for backward
//
// i.e. `residual + self.dropout(x)` followed by nn.LayerNorm of pasero/models/transformer.py:1043-1054, 1073-1086 with the
// nn.Linear that produced x (modules.py:737, transformer.py:1016). The separate dropout + residual + LayerNorm kernel
// (norm.cu) costs a 16 MB write of x, its 16 MB read back and a launch per sub-block; here a whole row of the output
// (512 columns) sits in ONE TMEM lane of a 256 x 512 CTA-pair tile (tcgen05.mma.cta_group::2, a single 512-column fp32
// accumulator = all of TMEM), so mean and variance need no cross-thread reduction beyond one exchange between the two
// warps that share a lane quarter.
//
// Roles: warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warp 2 TMEM allocation, warps 4.. epilogue (they first stage
// bias / gamma / beta in shared memory as fp32): warp w owns TMEM lanes [32 (w & 3), +32) and the column group (w - 4) >> 2.
// (16 epilogue warps, four column groups, were measured too: no faster — pass 1 was bound by the LSU, see the residual path.)
//   pass 1: accumulator -> +bias -> bf16 -> dropout -> +residual -> bf16 `sum`: written back to TMEM as fp32 (tcgen05.st),
//           staged in swizzled shared memory and stored by TMA; per-row sum / sum of squares;
//   exchange of the two column halves' partial statistics through shared memory (one named barrier);
//   pass 2: TMEM -> (sum - mean) rstd gamma + beta -> bf16 -> staged -> TMA store.
// Dropout keep decisions come from a counter hash of (seed, offset, element pair) — two 32-bit multiplies per two
// elements instead of Philox's ten rounds per eight, which would double this epilogue; the bits are stored (one byte per
// eight elements, the layout norm.cu's backward kernel reads), so backward never regenerates them.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "../../include/pasero_b200.h"

#include <stdlib.h>

namespace pb200 {

namespace gl {
static constexpr int BM = 128, BK = 64, BN = 512, MMA_N = 256;
static constexpr int STAGES = 3;
static constexpr int A_BYTES = BM * BK * 2;                 // 16 KB: this CTA's 128 rows of A
static constexpr int B_BYTES = (BN / 2) * BK * 2;           // 32 KB: this CTA's half of each 256-row slice of W
static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
static constexpr int EPI_BUF_BYTES = 32 * 128;              // 32 rows x 64 bf16, SWIZZLE_128B
static constexpr int EPI_WARPS = 8;                         // NQ column groups x 4 TMEM lane quarters
static constexpr int NQ = EPI_WARPS / 4;
static constexpr int COLS = 512 / NQ;                       // columns per epilogue warp
static constexpr int SLABS = COLS / 64;                     // 64-column slabs (one TMA box) per warp and pass
static constexpr int NBUF = 2;                              // staging buffers per warp
static constexpr int EPI_THREADS = EPI_WARPS * 32;
static constexpr int THREADS = 128 + EPI_THREADS;
static constexpr int EPI_BYTES = EPI_WARPS * NBUF * EPI_BUF_BYTES;
static constexpr int STATIC_BYTES = 3 * BN * 4 + NQ * BM * 8;   // bias / gamma / beta as fp32 + the row statistics exchange (static __shared__)
static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 256 /*barriers*/ + 1024 /*align*/;   // dynamic
static_assert(SMEM_BYTES + STATIC_BYTES <= 227 * 1024, "shared memory budget");
static_assert(SLABS % 2 == 0, "the staging-buffer / residual-barrier parities below assume an even slab count per pass");
}  // namespace gl

struct GemmLnDev {
    int M, kb_total, tiles;
    const bf16* bias;
    const bf16* residual;
    const bf16* gamma;
    const bf16* beta;
    float* mean;
    float* rstd;
    uint8_t* keep;
    float eps, keep_scale;
    uint32_t thresh16;
    uint64_t seed, offset;
    const long long* offset_base;
};

__device__ __forceinline__ void st_shared_v4(const void* ptr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ptx::smem_u32(ptr)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint4 ld_shared_v4(const void* ptr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ptx::smem_u32(ptr)) : "memory");
    return v;
}

__device__ __forceinline__ uint32_t mix32(uint32_t x) {   // 32-bit finaliser with full avalanche (two multiplies)
    x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u; x ^= x >> 15;
    return x;
}

__global__ void __launch_bounds__(gl::THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tma, const __grid_constant__ CUtensorMap tmb,
               const __grid_constant__ CUtensorMap tmr, const __grid_constant__ CUtensorMap tms,
               const __grid_constant__ CUtensorMap tmy, const GemmLnDev p) {
    using namespace gl;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;
    __shared__ __align__(16) float s_bias[BN];
    __shared__ __align__(16) float s_gamma[BN];
    __shared__ __align__(16) float s_beta[BN];
    __shared__ float2 s_stat[NQ * BM];         // [column group][row]: (sum, sum of squares)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + EPI_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 1;
    uint64_t* res_bar = tmem_empty + 1;        // [epilogue warp][buffer]: the residual slab has landed in that staging buffer
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + EPI_WARPS * NBUF);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta_rank = (int)ptx::cluster_ctarank();
    const int pair_id = (int)(blockIdx.x >> 1), num_pairs = (int)(gridDim.x >> 1);
    constexpr uint16_t MC_MASK = 3;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tma); ptx::prefetch_tmap(&tmb); ptx::prefetch_tmap(&tmr); ptx::prefetch_tmap(&tms); ptx::prefetch_tmap(&tmy);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            ptx::mbar_init(&full_bar[i], 2);    // one arrival per CTA of the pair (+ the bytes of both CTAs' loads)
            ptx::mbar_init(&empty_bar[i], 1);   // one multicast commit of the pair's MMAs
        }
        ptx::mbar_init(tmem_full, 1);
        ptx::mbar_init(tmem_empty, 2 * EPI_WARPS);   // one arrival per epilogue warp of both CTAs (leader's barrier)
        for (int i = 0; i < EPI_WARPS * NBUF; ++i) ptx::mbar_init(&res_bar[i], 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) { ptx::tmem_alloc_2sm(tmem_ptr, 512); ptx::tmem_relinquish_2sm(); }
    ptx::tc_fence_before();
    ptx::cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();   // global memory is touched only below

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair_id; tile < p.tiles; tile += num_pairs) {
                const int m0 = (tile * 2 + cta_rank) * BM;
                for (int kb = 0; kb < p.kb_total; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                    ptx::tma_load_3d_2sm(&tma, &full_bar[stage], sa, kb * BK, m0, 0);
#pragma unroll
                    for (int h = 0; h < 2; ++h)   // MMA h reads rows [128 h, 128 h + 128) of both CTAs' B buffers
                        ptx::tma_load_3d_2sm(&tmb, &full_bar[stage], sb + h * (MMA_N / 2) * 128, kb * BK,
                                             h * MMA_N + cta_rank * (MMA_N / 2), 0);
                    if (cta_rank != 0) ptx::mbar_arrive_remote(&full_bar[stage], 0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) =====================
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(2 * BM, MMA_N, false, false);
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            for (int tile = pair_id; tile < p.tiles; tile += num_pairs, ++local) {
                ptx::mbar_wait(tmem_empty, (local & 1) ^ 1);
                ptx::tc_fence_after();
                for (int kb = 0; kb < p.kb_total; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t adesc = ptx::make_smem_desc_sw128(sa + k * 32, 16, 1024);
                        const uint64_t bdesc = ptx::make_smem_desc_sw128(sb + k * 32, 16, 1024);
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            ptx::umma_bf16_2sm(tmem_base + h * MMA_N, adesc, bdesc + (uint64_t)(h * (MMA_N / 2) * 128 >> 4), idesc,
                                               (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    ptx::umma_commit_2sm_mc(&empty_bar[stage], MC_MASK);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit_2sm_mc(tmem_full, MC_MASK);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp & 3;            // TMEM lane quarter
        const int cq = (warp - 4) >> 2;     // column group: columns [cq * COLS, (cq + 1) * COLS)
        const int wslot = warp - 4;         // staging buffers of this warp
        const int rloc = ew * 32 + lane;    // row inside this CTA's 128
        const int et = threadIdx.x - 128;   // epilogue thread index
        uint64_t offset = p.offset;
        if (p.offset_base) offset += (uint64_t)(*p.offset_base) << 12;
        const uint32_t key = mix32((uint32_t)p.seed ^ mix32((uint32_t)(p.seed >> 32) + 0x9e3779b9u) ^
                                   mix32((uint32_t)offset * 0x85ebca77u + (uint32_t)(offset >> 32)));
        const bool drop = p.thresh16 > 0;
        // bias / gamma / beta -> fp32 in shared memory (hidden behind the first main loop)
        for (int i = et; i < BN; i += EPI_THREADS) {
            s_bias[i] = p.bias ? bf16_to_f(p.bias[i]) : 0.f;
            s_gamma[i] = bf16_to_f(p.gamma[i]);
            s_beta[i] = p.beta ? bf16_to_f(p.beta[i]) : 0.f;
        }
        ptx::bar_sync(1, EPI_THREADS);
        int local = 0;
        for (int tile = pair_id; tile < p.tiles; tile += num_pairs, ++local) {
            const int m0 = (tile * 2 + cta_rank) * BM;
            const int row = m0 + rloc;
            const bool row_ok = row < p.M;
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + cq * COLS;
            // The residual rows reach the registers through TMA and the staging buffers (per-lane global loads of a row each
            // touch 32 lines per instruction and kept the LSU busy for longer than the whole rest of pass 1): the slab's
            // [32 x 64] box lands in the buffer that will afterwards stage the slab's `sum`. First box: requested before the
            // accumulator is waited for (every store of the previous tile has been read out of the buffers by then).
            uint8_t* const wbuf = epi_smem + wslot * NBUF * EPI_BUF_BYTES;
            uint64_t* const rbar = res_bar + wslot * NBUF;
            if (lane == 0) {
                ptx::tma_store_wait_read<0>();
                ptx::mbar_arrive_expect_tx(&rbar[0], EPI_BUF_BYTES);
                ptx::tma_load_3d(&tmr, &rbar[0], wbuf, cq * COLS, m0 + ew * 32, 0);
            }

            ptx::mbar_wait(tmem_full, local & 1);
            ptx::tc_fence_after();
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
            for (int c64 = 0; c64 < SLABS; ++c64) {
                uint8_t* buf = wbuf + (c64 & 1) * EPI_BUF_BYTES;      // (slab is even at the start of a pass)
                ptx::mbar_wait(&rbar[c64 & 1], (c64 >> 1) & 1);        // residual slab c64 has landed (SLABS loads per tile: 2 per barrier)
                uint4 rcur[8];
#pragma unroll
                for (int g = 0; g < 8; ++g) rcur[g] = ld_shared_v4(buf + lane * 128 + ((g ^ (lane & 7)) << 4));
                uint32_t kbits[2] = {0xffffffffu, 0xffffffffu};   // keep bits of this thread's 64 columns
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int cl = c64 * 64 + half * 32;            // column inside this warp's group
                    const int col = cq * COLS + cl;
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + cl, v);
                    if (half == 1 && c64 + 1 < SLABS) {
                        // next slab's residual into the other buffer, once the store that last used it (slab c64 - 1, issued a
                        // half slab ago) has been read out
                        __syncwarp();
                        if (lane == 0) {
                            ptx::tma_store_wait_read<0>();
                            ptx::mbar_arrive_expect_tx(&rbar[(c64 + 1) & 1], EPI_BUF_BYTES);
                            ptx::tma_load_3d(&tmr, &rbar[(c64 + 1) & 1], wbuf + ((c64 + 1) & 1) * EPI_BUF_BYTES, cq * COLS + (c64 + 1) * 64,
                                             m0 + ew * 32, 0);
                        }
                    }
                    if (drop) {
                        uint32_t kb = 0;
                        const uint32_t e2 = ((uint32_t)row * BN + (uint32_t)col) >> 1;   // element-pair index (M * 512 / 2 < 2^32)
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const uint32_t h = mix32((e2 + j) * 0x9e3779b1u + key);
                            kb |= ((h & 0xffffu) >= p.thresh16 ? 1u : 0u) << (2 * j);
                            kb |= ((h >> 16) >= p.thresh16 ? 1u : 0u) << (2 * j + 1);
                        }
                        kbits[half] = kb;
                    }
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        const uint4 rv = rcur[half * 4 + g8];
                        const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
                        const float4 b0 = *reinterpret_cast<const float4*>(s_bias + col + g8 * 8);
                        const float4 b1 = *reinterpret_cast<const float4*>(s_bias + col + g8 * 8 + 4);
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        // bf16 roundings go through the packed convert (F2FP), shifts and the packed bf16 add (exact sum, one
                        // rounding): the scalar F2F runs at a quarter of that rate and there are three rounding points per element
                        uint32_t ow[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t xw = pack_bf16x2(__uint_as_float(v[g8 * 8 + 2 * j]) + bb[2 * j],
                                                      __uint_as_float(v[g8 * 8 + 2 * j + 1]) + bb[2 * j + 1]);   // the Linear's bf16 output
                            if (drop) {
                                const uint32_t k2 = (kbits[half] >> (g8 * 8 + 2 * j)) & 3u;
                                xw = pack_bf16x2(elem_lo_f(xw) * p.keep_scale, elem_hi_f(xw) * p.keep_scale);
                                xw &= ((k2 & 1u) ? 0x0000ffffu : 0u) | ((k2 & 2u) ? 0xffff0000u : 0u);
                            }
                            const bf162 sum2 = __hadd2(*reinterpret_cast<const bf162*>(&xw), *reinterpret_cast<const bf162*>(&rw[j]));
                            ow[j] = *reinterpret_cast<const uint32_t*>(&sum2);
                            const float o0 = elem_lo_f(ow[j]), o1 = elem_hi_f(ow[j]);
                            v[g8 * 8 + 2 * j] = __float_as_uint(o0); v[g8 * 8 + 2 * j + 1] = __float_as_uint(o1);
                            s1 += o0 + o1;
                            s2 = fmaf(o0, o0, s2); s2 = fmaf(o1, o1, s2);
                        }
                        const int chunk = half * 4 + g8;
                        st_shared_v4(buf + lane * 128 + ((chunk ^ (lane & 7)) << 4), make_uint4(ow[0], ow[1], ow[2], ow[3]));
                    }
                    // `sum` back into the accumulator's columns for pass 2
                    {
                        uint32_t lo[16], hi[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) { lo[j] = v[j]; hi[j] = v[16 + j]; }
                        ptx::tmem_st_32x16(taddr + cl, lo);
                        ptx::tmem_st_32x16(taddr + cl + 16, hi);
                    }
                }
                if (drop && p.keep && row_ok)   // 64 columns = 8 bytes of keep bits
                    *reinterpret_cast<uint2*>(p.keep + (((long long)row * BN + cq * COLS + c64 * 64) >> 3)) = make_uint2(kbits[0], kbits[1]);
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_3d(&tms, buf, cq * COLS + c64 * 64, m0 + ew * 32, 0);
                    ptx::tma_store_commit();
                }
            }
            ptx::tmem_st_wait();
            // row statistics: the other column groups' partial sums come through shared memory
            s_stat[cq * BM + rloc] = make_float2(s1, s2);
            ptx::bar_sync(1, EPI_THREADS);
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int q = 0; q < NQ; ++q) { const float2 o = s_stat[q * BM + rloc]; t1 += o.x; t2 += o.y; }   // same order in every group
            const float mean = t1 * (1.0f / BN);
            const float var = fmaxf(t2 * (1.0f / BN) - mean * mean, 0.f);
            const float rstd = rsqrtf(var + p.eps);
            if (cq == 0 && row_ok) { p.mean[row] = mean; p.rstd[row] = rstd; }
            const float nm = -mean * rstd;
#pragma unroll 1
            for (int c64 = 0; c64 < SLABS; ++c64) {
                uint8_t* buf = wbuf + (c64 & 1) * EPI_BUF_BYTES;
                if (lane == 0) ptx::tma_store_wait_read<NBUF - 1>();
                __syncwarp();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int cl = c64 * 64 + half * 32;
                    const int col = cq * COLS + cl;
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + cl, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        const float4 g0 = *reinterpret_cast<const float4*>(s_gamma + col + g8 * 8);
                        const float4 g1 = *reinterpret_cast<const float4*>(s_gamma + col + g8 * 8 + 4);
                        const float4 e0 = *reinterpret_cast<const float4*>(s_beta + col + g8 * 8);
                        const float4 e1 = *reinterpret_cast<const float4*>(s_beta + col + g8 * 8 + 4);
                        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                        const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                        float y[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            y[j] = fmaf(fmaf(__uint_as_float(v[g8 * 8 + j]), rstd, nm), gg[j], ee[j]);
                        const int chunk = half * 4 + g8;
                        st_shared_v4(buf + lane * 128 + ((chunk ^ (lane & 7)) << 4),
                                     make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7])));
                    }
                }
                ptx::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_3d(&tmy, buf, cq * COLS + c64 * 64, m0 + ew * 32, 0);
                    ptx::tma_store_commit();
                }
            }
            // the accumulator goes back to the MMA issuer (one arrival per warp on the leader CTA's barrier); the statistics
            // slots are free once every epilogue thread of this CTA has read them
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_remote(tmem_empty, 0);
            ptx::bar_sync(1, EPI_THREADS);
        }
        if (lane == 0) ptx::tma_store_wait<0>();
    }

    ptx::tc_fence_before();
    ptx::cluster_sync_all();   // no CTA leaves while its peer may still signal its barriers / read its shared memory
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_2sm(tmem_base, 512);
    }
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_linear_dropout_add_ln_fwd(const void* a, const void* w, const void* bias, const void* residual,
                                               const void* gamma, const void* beta, int64_t M, int64_t K, float eps, float p,
                                               uint64_t seed, uint64_t offset, const int64_t* offset_base, void* sum_out,
                                               void* y_out, float* mean_out, float* rstd_out, void* keep_out, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PB_CHECK(a && w && residual && gamma && sum_out && y_out && mean_out && rstd_out, "null pointer");
    PB_CHECK(M > 0 && K > 0 && K % 8 == 0, "linear+LN: M (%lld) must be positive and K (%lld) a positive multiple of 8", (long long)M, (long long)K);
    PB_CHECK(M * (long long)gl::BN < (1LL << 32), "linear+LN: M x 512 must stay below 2^32 elements");
    PB_CHECK(p >= 0.f && p < 1.f, "dropout probability must be in [0, 1)");
    PB_CHECK(p == 0.f || keep_out, "linear+LN: dropout needs the keep-bit buffer (backward reads it)");
    PB_CHECK((reinterpret_cast<uintptr_t>(residual) & 15) == 0 && (reinterpret_cast<uintptr_t>(keep_out) & 7) == 0,
             "linear+LN: residual must be 16-byte aligned and the keep bits 8-byte aligned");
    CUtensorMap ta, tb, tr, ts, ty;
    int rc;
    if ((rc = make_tmap(&ta, a, M, K, K, 1, 0, gl::BK, gl::BM))) return rc;
    if ((rc = make_tmap(&tb, w, gl::BN, K, K, 1, 0, gl::BK, gl::MMA_N / 2))) return rc;
    if ((rc = make_tmap(&tr, residual, M, gl::BN, gl::BN, 1, 0, 64, 32))) return rc;
    if ((rc = make_tmap(&ts, sum_out, M, gl::BN, gl::BN, 1, 0, 64, 32))) return rc;
    if ((rc = make_tmap(&ty, y_out, M, gl::BN, gl::BN, 1, 0, 64, 32))) return rc;
    GemmLnDev d;
    d.M = (int)M;
    d.kb_total = (int)((K + gl::BK - 1) / gl::BK);
    d.tiles = (int)((M + 2 * gl::BM - 1) / (2 * gl::BM));
    d.bias = reinterpret_cast<const bf16*>(bias);
    d.residual = reinterpret_cast<const bf16*>(residual);
    d.gamma = reinterpret_cast<const bf16*>(gamma);
    d.beta = reinterpret_cast<const bf16*>(beta);
    d.mean = mean_out; d.rstd = rstd_out;
    d.keep = reinterpret_cast<uint8_t*>(keep_out);
    d.eps = eps;
    d.keep_scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    d.thresh16 = dropout_thresh16(p);
    d.seed = seed; d.offset = offset;
    d.offset_base = reinterpret_cast<const long long*>(offset_base);

    auto kern = gemm_ln_kernel;
    const int smem_bytes = gl::SMEM_BYTES;
    static bool attr_set = false;
    if (!attr_set) {
        PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_set = true;
    }
    const int pairs = num_sms() / 2;
    const int grid = 2 * (d.tiles < pairs ? d.tiles : pairs);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(gl::THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    PB_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tr, ts, ty, d));
    PB_LAUNCH_CHECK();
    return 0;
}
