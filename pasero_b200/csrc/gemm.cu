// pasero_b200 — persistent, warp-specialised bf16 GEMM for sm_100a.
//
//   C[z] = epilogue( alpha * op(A[z]) * op(B[z])^T )        op(A): [M x K], op(B): [N x K]
//
// TMA (cp.async.bulk.tensor, SWIZZLE_128B) stages operand tiles in shared memory, one elected thread
// issues tcgen05.mma (128 x BN x 16, bf16 in, fp32 accumulate) into a double-buffered TMEM
// accumulator, and four epilogue warps drain TMEM with tcgen05.ld while the next tile's main loop
// is already running. Both operands may be K-major (reduction dim contiguous) or MN-major, so the
// same kernel serves forward (X W^T), dgrad (dY W) and wgrad (dY^T X) without transposes.
//
// Replaces the cuBLASLt / cuDNN calls behind nn.Linear / nn.Conv1d / Embedding.projection in the
// reference (pasero/models/modules.py:610-615,737,829-834,935-947; transformer.py:999-1019).
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "../../include/pasero_b200.h"

#include <mutex>
#include <string.h>

namespace pb200 {

#ifdef PB200_HALF
static constexpr CUtensorMapDataType ELEM_TMAP_TYPE = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
#else
static constexpr CUtensorMapDataType ELEM_TMAP_TYPE = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#endif
static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int GEMM_THREADS = 256;

struct GemmDev {
    int M, N;
    int Z;               // batches iterated as independent output problems (1 when reduce_batch)
    int kb_per;          // k-blocks per batch slice = ceil(K / BK)
    int total_kb;        // k-blocks in the whole reduction (kb_per * Zred)
    int kb_per_split;    // k-blocks handled by one split
    int splits;
    int m_blks, n_blks;
    int reduce_batch;
    int a_batched, b_batched;
    // epilogue
    void* c;
    long long ldc, c_batch_stride;
    const bf16* bias;
    const bf16* residual;
    const bf16* relu_mask;
    uint32_t* relu_bits_out;       // EPI 2: bit (row, col) of word [col / 32][row] = (output > 0)
    const uint32_t* relu_bits;     // EPI 6: the same words, used as the ReLU-backward mask
    long long relu_bits_ld;        // words between consecutive 32-column groups (>= rows)
    int act, out_fp32, accumulate;
    int tma_store;       // bf16 output staged through shared memory and written with TMA
    float alpha;
    float* partial;      // split-K partials [split][Z][M][N] fp32
    float* a_rowsum;     // EPI 7: fp32 [M], += sum_k op(A)[m][k] (bias gradient folded into the weight-gradient GEMM)
};

// PAIR: the CTA pair of a 2-CTA cluster works as one tcgen05.mma.cta_group::2 of M = 256: each CTA stages its own 128
// rows of A and only HALF of the B tile (the tensor core reads the other half from the peer's shared memory), which cuts
// the L2 -> SM operand traffic per output tile by a third. The K = 512 GEMMs of this model are bound by exactly that
// traffic (ncu: ~10 TB/s of TMA loads, the chip-wide L2 limit, with the tensor pipe ~50 % busy).
template <int BN, bool PAIR = false>
struct GemmCfg {
    // BN == 512 (CTA pair only): one 256 x 512 tile per pair and a single 512-column accumulator — for the N = 512
    // GEMMs of this model every pair then owns exactly one tile (63 tiles on 74 pairs) instead of 1.7 waves of 256-wide
    // tiles, and the A operand is fetched once instead of twice
    static constexpr int STAGES = BN == 512 ? 4 : (PAIR ? 6 : ((BN == 128) ? 6 : 4));
    static constexpr int NACC = BN == 512 ? 1 : 2;          // accumulator buffers in TMEM
    static constexpr int NH = BN == 512 ? 2 : 1;            // tcgen05.mma N is at most 256: N-halves per k-step
    static constexpr int MMA_N = BN / NH;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = NACC * BN;
    static constexpr int EPI_BUF_BYTES = 32 * 128;                 // 32 rows x 64 bf16, SWIZZLE_128B
    static constexpr int EPI_BYTES = 4 /*warps*/ * 2 /*buffers*/ * EPI_BUF_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// EPI 7 = EPI 4 + row sums of op(A) (the bias gradient of the Linear whose weight gradient this GEMM computes): one more
// N = 16 MMA per k-step against an all-ones B operand into 16 extra TMEM columns (the tiles of one A row block take turns
// on it), so dY is not read a second time by a column-sum kernel. The extra columns leave room for a single accumulator only.
static constexpr int ONES_BYTES = 1152;   // 16 all-ones rows read as 8-row groups 16 B apart (any layout of ones is ones)
template <int BN, int EPI, bool PAIR>
struct GemmEpiCfg {
    static constexpr int NACC = EPI == 7 ? 1 : GemmCfg<BN, PAIR>::NACC;
    static constexpr int TMEM_COLS = EPI == 7 ? (BN >= 256 ? 512 : 256) : GemmCfg<BN, PAIR>::TMEM_COLS;
    static constexpr int ROWSUM_COL = BN;   // EPI 7: first of the 16 row-sum columns
    static constexpr int SMEM_BYTES = GemmCfg<BN, PAIR>::SMEM_BYTES + (EPI == 7 ? ONES_BYTES : 0);
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// EPI selects the epilogue at compile time (the generic one carries ~2500 SASS instructions per 64 columns when every
// option is a run-time flag, which made the kernel epilogue-bound):
//   -1 generic direct-store path (fp32 out, accumulate, split-K partials, alpha, GELU, residual)
//    0 plain bf16   1 +bias   2 +bias+ReLU (+ optional 1-bit-per-element sign map)   3 x ReLU-mask (bf16 tensor)
//    5 +residual (bf16, C layout)   6 x ReLU-mask (bit map written by EPI 2: 16x less epilogue traffic than EPI 3)
//                                                                           -> staged in smem, written by TMA
//    4 split-K: fp32 partial tile TMA-reduce-added (+=, performed in L2) into one fp32 accumulation buffer
// CL == 2: CTA pair (see GemmCfg). CL == 4: thread-block clusters of CL CTAs along M with TMA multicast of B: The CTAs of a cluster work on CL consecutive m-blocks of the same
// n-block, so they need the same B tile: each CTA fetches 1/CL of it and TMA-multicasts that slice into every CTA's
// shared memory, cutting the L2 -> SM operand traffic of these weight-small GEMMs (B re-read once per m-block) by
// up to (1 - 1/CL) of the B share.
template <int BN, bool A_MN, bool B_MN, int EPI, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma, const __grid_constant__ CUtensorMap tmb,
            const __grid_constant__ CUtensorMap tmc, const GemmDev p) {
    constexpr bool PAIR = (CL == 2);
    using Cfg = GemmCfg<BN, PAIR>;
    using ECfg = GemmEpiCfg<BN, EPI, PAIR>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int NACC = ECfg::NACC;
    static_assert(EPI != 7 || (A_MN && BN <= 256 && CL <= 2), "row sums: weight-gradient form only");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + Cfg::EPI_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    [[maybe_unused]] uint32_t* ones = reinterpret_cast<uint32_t*>(epi_smem + Cfg::EPI_BYTES + 256);   // EPI 7

    pdl_trigger();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int cta_rank = CL > 1 ? (int)ptx::cluster_ctarank() : 0;
    const int cluster_id = CL > 1 ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
    const int num_clusters = CL > 1 ? (int)(gridDim.x / CL) : (int)gridDim.x;
    constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tma);
        ptx::prefetch_tmap(&tmb);
        if (EPI >= 0) ptx::prefetch_tmap(&tmc);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            // PAIR: the leader's full barrier collects one arrival per CTA (+ the bytes of both CTAs' loads); one
            // multicast commit of the pair's MMAs releases the stage in both CTAs
            ptx::mbar_init(&full_bar[i], PAIR ? 2 : 1);
            ptx::mbar_init(&empty_bar[i], PAIR ? 1 : CL);   // multicast clusters: every CTA releases the stage
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tmem_full[i], 1);
            ptx::mbar_init(&tmem_empty[i], PAIR ? 8 : 128);   // PAIR: one arrival per epilogue warp of both CTAs (leader's barrier)
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) { ptx::tmem_alloc_2sm(tmem_ptr, ECfg::TMEM_COLS); ptx::tmem_relinquish_2sm(); }
        else { ptx::tmem_alloc(tmem_ptr, ECfg::TMEM_COLS); ptx::tmem_relinquish(); }
    }
    if (EPI == 7 && warp == 3) {   // bf16 1.0 everywhere; read by the tensor core (async proxy)
        for (int i = lane; i < ONES_BYTES / 4; i += 32) ones[i] = ELEM_ONE_X2;
        ptx::fence_proxy_async();
    }
    ptx::tc_fence_before();
    if (CL > 1) ptx::cluster_sync_all(); else __syncthreads();   // peers' barriers are initialised before any remote arrive
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();   // everything above overlapped the previous kernel's tail; operands / outputs are touched only below

    const int m_groups = (p.m_blks + CL - 1) / CL;
    const int total_tiles = p.Z * m_groups * p.n_blks * p.splits;   // cluster-level tiles

    if (warp == 0) {
        // ===================== TMA producer (one thread) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
                int r = tile;
                const int split = r % p.splits; r /= p.splits;
                const int n_blk = r % p.n_blks; r /= p.n_blks;
                const int m_blk = (r % m_groups) * CL + cta_rank; r /= m_groups;
                const int z = r;
                const int m0 = m_blk * BM, n0 = n_blk * BN;
                const int kb_begin = split * p.kb_per_split;
                const int kb_end = min(p.total_kb, kb_begin + p.kb_per_split);
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    const int zk = p.reduce_batch ? (kb / p.kb_per) : z;
                    const int k0 = (kb % p.kb_per) * BK;
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    const int za = p.a_batched ? zk : 0;
                    const int zb = p.b_batched ? zk : 0;
                    if constexpr (PAIR) {
                        // both CTAs' bytes are credited to the leader's barrier
                        if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                        if (!A_MN) {
                            ptx::tma_load_3d_2sm(&tma, &full_bar[stage], sa, k0, m0, za);
                        } else {
#pragma unroll
                            for (int j = 0; j < BM / 64; ++j)
                                ptx::tma_load_3d_2sm(&tma, &full_bar[stage], sa + j * (BK * 128), m0 + j * 64, k0, za);
                        }
                        // this CTA's half of each MMA_N-wide slice of the B tile (MMA h reads rows / columns
                        // [h * MMA_N/2, (h + 1) * MMA_N/2) of both CTAs' B buffers)
#pragma unroll
                        for (int h = 0; h < Cfg::NH; ++h) {
                            const int nh = n0 + h * Cfg::MMA_N + cta_rank * (Cfg::MMA_N / 2);
                            uint8_t* sbh = sb + h * (Cfg::MMA_N / 2) * 128;
                            if (!B_MN) {
                                ptx::tma_load_3d_2sm(&tmb, &full_bar[stage], sbh, k0, nh, zb);
                            } else {
#pragma unroll
                                for (int j = 0; j < Cfg::MMA_N / 128; ++j)
                                    ptx::tma_load_3d_2sm(&tmb, &full_bar[stage], sbh + j * (BK * 128), nh + j * 64, k0, zb);
                            }
                        }
                        if (cta_rank != 0) ptx::mbar_arrive_remote(&full_bar[stage], 0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    if (!A_MN) {
                        ptx::tma_load_3d(&tma, &full_bar[stage], sa, k0, m0, za);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j)
                            ptx::tma_load_3d(&tma, &full_bar[stage], sa + j * (BK * 128), m0 + j * 64, k0, za);
                    }
                    if (CL == 1) {
                        if (!B_MN) {
                            ptx::tma_load_3d(&tmb, &full_bar[stage], sb, k0, n0, zb);
                        } else {
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j)
                                ptx::tma_load_3d(&tmb, &full_bar[stage], sb + j * (BK * 128), n0 + j * 64, k0, zb);
                        }
                    } else if (!B_MN) {   // this CTA's BN/CL rows of the B tile, multicast to the whole cluster
                        constexpr int ROWS = BN / CL;
                        ptx::tma_load_3d_mc(&tmb, &full_bar[stage], sb + cta_rank * ROWS * 128, k0, n0 + cta_rank * ROWS, zb, MC_MASK);
                    } else {              // this CTA's share of the 64-column boxes
                        constexpr int PER = (BN / 64) / CL;
#pragma unroll
                        for (int jj = 0; jj < PER; ++jj) {
                            const int j = cta_rank * PER + jj;
                            ptx::tma_load_3d_mc(&tmb, &full_bar[stage], sb + j * (BK * 128), n0 + j * 64, k0, zb, MC_MASK);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread; the leader CTA's in PAIR mode) =====================
        if (lane == 0 && (!PAIR || cta_rank == 0)) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(PAIR ? 2 * BM : BM, Cfg::MMA_N, A_MN, B_MN);
            [[maybe_unused]] constexpr uint32_t idesc_rs = ptx::make_idesc_bf16(PAIR ? 2 * BM : BM, 16, A_MN, false);
            [[maybe_unused]] const uint64_t ones_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(ones), 16, 16);
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++local) {
                const int split = tile % p.splits;
                const int kb_begin = split * p.kb_per_split;
                const int kb_end = min(p.total_kb, kb_begin + p.kb_per_split);
                const int acc = NACC == 2 ? (local & 1) : 0;
                const uint32_t acc_phase = (NACC == 2 ? (local >> 1) : local) & 1;
                // EPI 7: the n_blks tiles that share this tile's A operand take turns on the row-sum MMA (k-step s of the
                // tile belongs to n-block s % n_blks), so the extra tensor work is spread evenly over the CTAs
                [[maybe_unused]] int rs_wait = EPI == 7 ? (tile / p.splits) % p.n_blks : 0;
                [[maybe_unused]] uint32_t rs_acc = 0;
                ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;   // (acc == 0 when there is a single accumulator)
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // K-major: 8-row groups 1024 B apart, advance 32 B per UMMA_K inside the 128 B row.
                        // MN-major: 64-element column blocks (BK*128 B apart, LBO), 8-k-row groups 1024 B
                        // apart (SBO), advance 16 k-rows = 2048 B per UMMA_K.
                        const uint64_t adesc = A_MN ? ptx::make_smem_desc_sw128(sa + k * 2048, BK * 128, 1024)
                                                    : ptx::make_smem_desc_sw128(sa + k * 32, 16, 1024);
                        const uint64_t bdesc = B_MN ? ptx::make_smem_desc_sw128(sb + k * 2048, BK * 128, 1024)
                                                    : ptx::make_smem_desc_sw128(sb + k * 32, 16, 1024);
                        if (PAIR) {
#pragma unroll
                            for (int h = 0; h < Cfg::NH; ++h)   // + 16 KB = the next MMA_N/2 rows (K-major) or 64-column boxes (MN-major)
                                ptx::umma_bf16_2sm(tmem_d + h * Cfg::MMA_N, adesc, bdesc + (uint64_t)(h * (Cfg::MMA_N / 2) * 128 >> 4), idesc,
                                                   (kb > kb_begin || k > 0) ? 1u : 0u);
                        } else {
                            ptx::umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
                        }
                        if constexpr (EPI == 7) {
                            if (rs_wait == 0) {
                                if (PAIR) ptx::umma_bf16_2sm(tmem_base + ECfg::ROWSUM_COL, adesc, ones_desc, idesc_rs, rs_acc);
                                else ptx::umma_bf16(tmem_base + ECfg::ROWSUM_COL, adesc, ones_desc, idesc_rs, rs_acc);
                                rs_acc = 1u;
                                rs_wait = p.n_blks;
                            }
                            --rs_wait;
                        }
                    }
                    // frees the smem slot (in every CTA of the cluster: peers multicast into it / the pair's MMA read it)
                    // when these MMAs retire
                    if (PAIR) ptx::umma_commit_2sm_mc(&empty_bar[stage], MC_MASK);
                    else if (CL > 1) ptx::umma_commit_mc(&empty_bar[stage], MC_MASK);
                    else ptx::umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                // accumulator ready for the epilogue (of both CTAs in PAIR mode)
                if (PAIR) ptx::umma_commit_2sm_mc(&tmem_full[acc], MC_MASK); else ptx::umma_commit(&tmem_full[acc]);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps (TMEM -> registers -> global) =====================
        const int ew = warp & 3;  // TMEM lane quarter this warp may access
        // hands the accumulator buffer back to the MMA issuer (PAIR: one arrival per warp on the leader CTA's barrier)
        auto release_accumulator = [&](uint64_t* bar) {
            if (PAIR) {
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_remote(bar, 0);
            } else {
                ptx::mbar_arrive(bar);
            }
        };
        int local = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++local) {
            int r = tile;
            const int split = r % p.splits; r /= p.splits;
            const int n_blk = r % p.n_blks; r /= p.n_blks;
            const int m_blk = (r % m_groups) * CL + cta_rank; r /= m_groups;
            const int z = r;
            const int n0 = n_blk * BN;
            const int row = m_blk * BM + ew * 32 + lane;
            const bool row_ok = row < p.M;
            const int acc = NACC == 2 ? (local & 1) : 0;
            const uint32_t acc_phase = (NACC == 2 ? (local >> 1) : local) & 1;
            const int kb_begin = split * p.kb_per_split;
            const bool empty_range = kb_begin >= p.total_kb;

            // Auxiliary epilogue operands are requested BEFORE the accumulator is waited for, and kept one 64-column slab
            // ahead inside the loop: a dependent global load per 32-column slab (measured) made these epilogues longer
            // than the next tile's main loop and the whole GEMM ~1.7x slower than its plain variant.
            uint32_t mword[BN / 32];   // EPI 6: ReLU sign bits of this lane's row for the whole tile
            uint4 aux_q[2][4];         // EPI 3 / 5: bf16 aux values of the next two 32-column slabs of this lane's row
            [[maybe_unused]] const long long aux_row = (long long)z * p.c_batch_stride + (long long)row * p.ldc;
            [[maybe_unused]] auto load_aux = [&](int c64, int half) {
                const bf16* aux = EPI == 3 ? p.relu_mask : p.residual;
                const uint32_t fill = EPI == 3 ? ELEM_ONE_X2 : 0u;
                const int colh = n0 + c64 * 64 + half * 32;
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    aux_q[half][g8] = make_uint4(fill, fill, fill, fill);
                    if (row_ok && c64 < BN / 64 && colh + g8 * 8 < p.N)
                        aux_q[half][g8] = __ldg(reinterpret_cast<const uint4*>(aux + aux_row + colh + g8 * 8));
                }
            };
            if constexpr (EPI == 6) {
#pragma unroll
                for (int w = 0; w < BN / 32; ++w) {
                    mword[w] = 0xffffffffu;
                    if (row_ok && n0 + w * 32 < p.N) mword[w] = __ldg(p.relu_bits + (long long)((n0 >> 5) + w) * p.relu_bits_ld + row);
                }
            }
            if constexpr (EPI == 3 || EPI == 5) { load_aux(0, 0); load_aux(0, 1); }

            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BN;

            if constexpr (EPI == 4 || EPI == 7) {
                // split-K partial: 32-column fp32 slabs -> swizzled smem -> TMA reduce-add into the fp32 accumulator
#pragma unroll 1
                for (int c32 = 0; c32 < BN / 32; ++c32) {
                    const int col0 = n0 + c32 * 32;
                    if (col0 >= p.N) break;   // warp-uniform
                    uint8_t* buf = epi_smem + (ew * 2 + (c32 & 1)) * Cfg::EPI_BUF_BYTES;
                    if (lane == 0) ptx::tma_store_wait_read<1>();
                    __syncwarp();
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + c32 * 32, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch)   // 8 x 16-byte chunks (4 floats) per 128-byte row
                        *reinterpret_cast<uint4*>(buf + lane * 128 + ((ch ^ (lane & 7)) << 4)) =
                            empty_range ? make_uint4(0, 0, 0, 0) : make_uint4(v[ch * 4], v[ch * 4 + 1], v[ch * 4 + 2], v[ch * 4 + 3]);
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        ptx::tma_reduce_add_3d(&tmc, buf, col0, m_blk * BM + ew * 32, z);
                        ptx::tma_store_commit();
                    }
                }
                if constexpr (EPI == 7) {
                    if ((min(p.total_kb, kb_begin + p.kb_per_split) - kb_begin) * (BK / 16) > n_blk) {   // this tile had a turn (warp-uniform);
                        // the lane's partial row sum sits in every one of the 16 extra columns
                        uint32_t v[32];
                        __syncwarp();
                        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + ECfg::ROWSUM_COL, v);
                        ptx::tmem_ld_wait();
                        if (row_ok && !empty_range) atomicAdd(p.a_rowsum + row, __uint_as_float(v[0]));
                    }
                }
                ptx::tc_fence_before();
                release_accumulator(&tmem_empty[acc]);
                continue;
            } else if constexpr (EPI >= 0) {
                // bf16 output: 64-column slabs -> swizzled smem (this warp's 32 rows) -> TMA store. The TMA clips
                // rows >= M and columns >= N, so no store masks are needed; two staging buffers per warp.
                const long long coff_row = (long long)z * p.c_batch_stride + (long long)row * p.ldc;
                // bias for the whole tile: lane l keeps columns [n0 + 8l, n0 + 8l + 8); broadcast by shuffle later, so
                // no global-load latency sits between the TMEM reads and the stores
                uint4 bias_reg = make_uint4(0, 0, 0, 0), bias_reg2 = make_uint4(0, 0, 0, 0);   // columns [0,256) / [256,512)
                if ((EPI == 1 || EPI == 2) && lane * 8 < BN && n0 + lane * 8 < p.N)
                    bias_reg = __ldg(reinterpret_cast<const uint4*>(p.bias + n0 + lane * 8));
                if ((EPI == 1 || EPI == 2) && BN > 256 && n0 + 256 + lane * 8 < p.N)
                    bias_reg2 = __ldg(reinterpret_cast<const uint4*>(p.bias + n0 + 256 + lane * 8));
#pragma unroll 1
                for (int c64 = 0; c64 < BN / 64; ++c64) {
                    const int col0 = n0 + c64 * 64;
                    if (col0 >= p.N) break;   // warp-uniform
                    uint8_t* buf = epi_smem + (ew * 2 + (c64 & 1)) * Cfg::EPI_BUF_BYTES;
                    if (lane == 0) ptx::tma_store_wait_read<1>();   // the store that last used this buffer has read it
                    __syncwarp();
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int colh = col0 + half * 32;
                        // this slab's auxiliary operands were requested one slab pair ago; refill the slot for c64 + 1
                        uint4 aux_m[4];
                        if constexpr (EPI == 3 || EPI == 5) {
#pragma unroll
                            for (int g8 = 0; g8 < 4; ++g8) aux_m[g8] = aux_q[half][g8];
                            load_aux(c64 + 1, half);
                        }
                        uint32_t mbits = 0xffffffffu;   // EPI 6: ReLU sign bits of this lane's row, columns [colh, colh + 32)
                        if constexpr (EPI == 6) {
#pragma unroll
                            for (int w = 0; w < BN / 32; ++w)
                                if (w == c64 * 2 + half) mbits = mword[w];
                        }
                        (void)colh;
                        uint32_t obits = 0;             // EPI 2: sign bits produced here
                        uint32_t v[32];
                        ptx::tmem_ld_32x32(taddr + c64 * 64 + half * 32, v);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int g8 = 0; g8 < 4; ++g8) {
                            float x[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[g8 * 8 + j]);
                            if constexpr (EPI == 1 || EPI == 2) {
                                const int src = (c64 & 3) * 8 + half * 4 + g8;
                                const uint4 br = (BN > 256 && c64 >= 4) ? bias_reg2 : bias_reg;   // warp-uniform
                                const uint32_t bw[4] = {__shfl_sync(0xffffffffu, br.x, src), __shfl_sync(0xffffffffu, br.y, src),
                                                        __shfl_sync(0xffffffffu, br.z, src), __shfl_sync(0xffffffffu, br.w, src)};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 b2 = unpack_bf16x2(bw[j]);
                                    x[2 * j] += b2.x; x[2 * j + 1] += b2.y;
                                }
                            }
                            if constexpr (EPI == 2) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    obits |= (x[j] > 0.f ? 1u : 0u) << (g8 * 8 + j);
                                    x[j] = fmaxf(x[j], 0.f);
                                }
                            }
                            if constexpr (EPI == 6) {
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    if (!((mbits >> (g8 * 8 + j)) & 1u)) x[j] = 0.f;
                            }
                            if constexpr (EPI == 3) {
                                const uint32_t mw[4] = {aux_m[g8].x, aux_m[g8].y, aux_m[g8].z, aux_m[g8].w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {   // bf16 > 0  <=>  sign bit clear and not +0
                                    if ((int)(mw[j] << 16) <= 0) x[2 * j] = 0.f;
                                    if ((int)(mw[j] & 0xffff0000u) <= 0) x[2 * j + 1] = 0.f;
                                }
                            }
                            if constexpr (EPI == 5) {   // bf16(result) + residual, rounded again: what `a + b` on two bf16 tensors gives
                                const uint32_t rw[4] = {aux_m[g8].x, aux_m[g8].y, aux_m[g8].z, aux_m[g8].w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 r2 = unpack_bf16x2(rw[j]);
                                    x[2 * j] = round_bf16(x[2 * j]) + r2.x; x[2 * j + 1] = round_bf16(x[2 * j + 1]) + r2.y;
                                }
                            }
                            uint4 ov;
                            ov.x = pack_bf16x2(x[0], x[1]); ov.y = pack_bf16x2(x[2], x[3]);
                            ov.z = pack_bf16x2(x[4], x[5]); ov.w = pack_bf16x2(x[6], x[7]);
                            const int chunk = half * 4 + g8;   // 16-byte chunk inside the 128-byte row
                            *reinterpret_cast<uint4*>(buf + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = ov;
                        }
                        if constexpr (EPI == 2) {   // lanes = consecutive rows: one coalesced 128-byte store per warp
                            if (p.relu_bits_out && row_ok && colh < p.N)
                                p.relu_bits_out[(long long)(colh >> 5) * p.relu_bits_ld + row] = obits;
                        }
                    }
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        ptx::tma_store_3d(&tmc, buf, col0, m_blk * BM + ew * 32, z);
                        ptx::tma_store_commit();
                    }
                }
                ptx::tc_fence_before();
                release_accumulator(&tmem_empty[acc]);
                continue;
            } else {

#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the predicated stores below
                ptx::tmem_ld_32x32(taddr + c * 32, v);
                ptx::tmem_ld_wait();
                const int col0 = n0 + c * 32;
                if (!row_ok || col0 >= p.N) continue;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = empty_range ? 0.f : __uint_as_float(v[j]);

                if (p.splits > 1) {
                    float* dst = p.partial + (((long long)split * p.Z + z) * p.M + row) * (long long)p.N + col0;
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        if (col0 + g * 4 < p.N)
                            *reinterpret_cast<float4*>(dst + g * 4) = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
                    continue;
                }
                const long long coff = (long long)z * p.c_batch_stride + (long long)row * p.ldc + col0;
#pragma unroll
                for (int g = 0; g < 4; ++g) {  // groups of 8 columns
                    if (col0 + g * 8 >= p.N) break;
                    float* x = f + g * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[j] *= p.alpha;
                    if (p.bias) {
                        const uint4 bv = __ldg(reinterpret_cast<const uint4*>(p.bias + col0 + g * 8));
                        const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 b2 = unpack_bf16x2(bw[j]);
                            x[2 * j] += b2.x; x[2 * j + 1] += b2.y;
                        }
                    }
                    if (p.act == 1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) x[j] = fmaxf(x[j], 0.f);
                    } else if (p.act == 2) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) x[j] = gelu_erf(x[j]);
                    }
                    if (p.relu_mask) {
                        const uint4 mv = __ldg(reinterpret_cast<const uint4*>(p.relu_mask + coff + g * 8));
                        const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 m2 = unpack_bf16x2(mw[j]);
                            if (!(m2.x > 0.f)) x[2 * j] = 0.f;
                            if (!(m2.y > 0.f)) x[2 * j + 1] = 0.f;
                        }
                    }
                    if (p.residual) {
                        const uint4 rv = __ldg(reinterpret_cast<const uint4*>(p.residual + coff + g * 8));
                        const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 r2 = unpack_bf16x2(rw[j]);
                            x[2 * j] += r2.x; x[2 * j + 1] += r2.y;
                        }
                    }
                    if (p.out_fp32) {
                        float* dst = reinterpret_cast<float*>(p.c) + coff + g * 8;
                        if (p.accumulate) {
                            const float4 o0 = *reinterpret_cast<const float4*>(dst);
                            const float4 o1 = *reinterpret_cast<const float4*>(dst + 4);
                            x[0] += o0.x; x[1] += o0.y; x[2] += o0.z; x[3] += o0.w;
                            x[4] += o1.x; x[5] += o1.y; x[6] += o1.z; x[7] += o1.w;
                        }
                        *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
                        *reinterpret_cast<float4*>(dst + 4) = make_float4(x[4], x[5], x[6], x[7]);
                    } else {
                        bf16* dst = reinterpret_cast<bf16*>(p.c) + coff + g * 8;
                        if (p.accumulate) {
                            const uint4 ov = *reinterpret_cast<const uint4*>(dst);
                            const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 o2 = unpack_bf16x2(ow[j]);
                                x[2 * j] += o2.x; x[2 * j + 1] += o2.y;
                            }
                        }
                        uint4 ov;
                        ov.x = pack_bf16x2(x[0], x[1]); ov.y = pack_bf16x2(x[2], x[3]);
                        ov.z = pack_bf16x2(x[4], x[5]); ov.w = pack_bf16x2(x[6], x[7]);
                        *reinterpret_cast<uint4*>(dst) = ov;
                    }
                }
            }
            ptx::tc_fence_before();
            release_accumulator(&tmem_empty[acc]);
            }
        }
    }

    if (EPI >= 0 && warp >= 4 && lane == 0) ptx::tma_store_wait<0>();
    ptx::tc_fence_before();
    if (CL > 1) ptx::cluster_sync_all(); else __syncthreads();   // no CTA leaves while peers may still signal its barriers
    if (warp == 2) {
        ptx::tc_fence_after();
        if (PAIR) ptx::tmem_dealloc_2sm(tmem_base, ECfg::TMEM_COLS); else ptx::tmem_dealloc(tmem_base, ECfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// split-K reduction: out = alpha * sum_s partial[s] (+bias) (+ old out) ; bf16 or fp32 out
// ------------------------------------------------------------------------------------------------
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, long long zmn, int M, int N,
                                     void* out, long long ldc, long long c_batch_stride, int out_fp32,
                                     int accumulate, float alpha, const float* __restrict__ extra_fp32) {
    const long long n4 = zmn / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * 4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < splits; ++k) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(partial + (long long)k * zmn + e));
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        s.x *= alpha; s.y *= alpha; s.z *= alpha; s.w *= alpha;
        if (extra_fp32) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(extra_fp32 + e));
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        const int col = (int)(e % N);
        const long long rz = e / N;
        const int row = (int)(rz % M);
        const long long z = rz / M;
        const long long off = z * c_batch_stride + (long long)row * ldc + col;
        if (out_fp32) {
            float* dst = reinterpret_cast<float*>(out) + off;
            if (accumulate) {
                const float4 o = *reinterpret_cast<const float4*>(dst);
                s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
            }
            *reinterpret_cast<float4*>(dst) = s;
        } else {
            bf16* dst = reinterpret_cast<bf16*>(out) + off;
            if (accumulate) {
                const uint2 o = *reinterpret_cast<const uint2*>(dst);
                const float2 a = unpack_bf16x2(o.x), b = unpack_bf16x2(o.y);
                s.x += a.x; s.y += a.y; s.z += b.x; s.w += b.y;
            }
            uint2 o;
            o.x = pack_bf16x2(s.x, s.y); o.y = pack_bf16x2(s.z, s.w);
            *reinterpret_cast<uint2*>(dst) = o;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

// Tensor map over a bf16 matrix stored as `rows` x `cols` (cols contiguous), `ld` elements between
// rows, `batch` slices `batch_stride` elements apart; box = box_cols x box_rows x 1, SWIZZLE_128B.
int make_tmap(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld,
              long long batch, long long batch_stride, int box_cols, int box_rows, bool fp32) {
    EncodeTiledFn enc = get_encode_fn();
    PB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    // The driver-level encode needs a context current on THIS thread. PyTorch's autograd worker threads skip
    // cudaSetDevice when the target is device 0 (the runtime default), so nothing has bound the primary context
    // yet when backward's first op is a GEMM; a runtime call binds it (once per thread).
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) {
        PB_CUDA(cudaFree(nullptr));
        ctx_bound = true;
    }
    PB_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "gemm operand pointer must be 16-byte aligned");
    PB_CHECK(ld % (fp32 ? 4 : 8) == 0, "gemm operand leading dimension (%lld) must be a multiple of 16 bytes", ld);
    PB_CHECK(batch <= 1 || batch_stride % 8 == 0, "gemm operand batch stride (%lld) must be a multiple of 8", batch_stride);
    const int esz = fp32 ? 4 : 2;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(batch > 0 ? batch : 1)};
    long long bs = (batch > 1 && batch_stride > 0) ? batch_stride : rows * ld;
    cuuint64_t strides[2] = {(cuuint64_t)ld * esz, (cuuint64_t)bs * esz};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : ELEM_TMAP_TYPE, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PB_CHECK(r == CUDA_SUCCESS,
             "cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld batch=%lld bstride=%lld box=%dx%d",
             (int)r, rows, cols, ld, batch, bs, box_cols, box_rows);
    return 0;
}

template <int BN, bool A_MN, bool B_MN, int EPI, int CL>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmDev& d, int grid,
                       cudaStream_t stream) {
    auto kern = gemm_kernel<BN, A_MN, B_MN, EPI, CL>;
    static bool attr_set = false;
    if (!attr_set) {
        PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmEpiCfg<BN, EPI, CL == 2>::SMEM_BYTES));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = GemmEpiCfg<BN, EPI, CL == 2>::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (CL > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CL; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    PB_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, d));
    PB_LAUNCH_CHECK();
    return 0;
}

}  // namespace pb200

using namespace pb200;

extern "C" size_t pb200_gemm_workspace_bytes(int64_t m, int64_t n, int64_t batch, int32_t max_splits) {
    (void)max_splits;   // split-K partials are reduce-added into a single fp32 [batch, m, n] buffer
    return (size_t)(batch > 0 ? batch : 1) * (size_t)m * (size_t)n * sizeof(float);
}

extern "C" int pb200_gemm_bf16(const pb200_gemm_params* q, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    PB_CHECK(q != nullptr, "null params");
    PB_CHECK(q->m > 0 && q->n > 0 && q->k > 0, "gemm dims must be positive (m=%lld n=%lld k=%lld)",
             (long long)q->m, (long long)q->n, (long long)q->k);
    PB_CHECK(q->n % 8 == 0, "gemm N (%lld) must be a multiple of 8", (long long)q->n);
    PB_CHECK(q->ldc % 8 == 0, "gemm ldc must be a multiple of 8");
    const long long Zall = q->batch > 0 ? q->batch : 1;
    const bool a_mn = q->a_mn_major != 0, b_mn = q->b_mn_major != 0;

    // tile width: wide tiles when N is large enough to keep >= ~1 wave of CTAs busy
    const long long m_blks = (q->m + BM - 1) / BM;
    const long long Zout = q->reduce_batch ? 1 : Zall;
    // fp32 C with accumulate and no other epilogue work: partial tiles of ANY split count are TMA-reduce-added straight
    // into C (fp32 main-gradient accumulation: no workspace, no memset, no reduction kernel)
    const bool direct_acc = q->out_fp32 && q->accumulate && q->alpha == 1.0f && !q->bias && !q->residual && !q->relu_mask &&
                            !q->relu_bits && !q->relu_bits_out &&
                            q->act == 0 && q->ldc % 4 == 0 && q->c_batch_stride % 4 == 0 &&
                            (reinterpret_cast<uintptr_t>(q->c) & 15) == 0;
    // 128 x 256 tiles halve the A re-reads and the per-tile overheads; fall back to 128 x 128 only when the wide
    // tiling would leave most SMs without a tile
    int bn = 128;
    if (q->n >= 256 && q->tile_n != 128) {
        const long long tiles256 = Zout * m_blks * ((q->n + 255) / 256);
        const long long kb_all = ((q->k + BK - 1) / BK) * (q->reduce_batch ? Zall : 1);
        const bool can_split = q->split_k != 1 && kb_all >= 32 && (q->workspace != nullptr || direct_acc);   // split-K fills the SMs instead
        if (tiles256 * 2 >= num_sms() || can_split || q->tile_n == 256) bn = 256;
    }
    // 256 x 512 tiles on CTA pairs (single 512-column accumulator: the epilogue is not overlapped with the next tile, but A
    // is fetched half as often): pays off for long reductions (8192^3: 1498 -> 1627 TFLOP/s); measured NOT to pay off for
    // the K <= 2048, N = 512 GEMMs of this model even though each pair would then own exactly one tile, so it is chosen
    // automatically only when K >= 4096
    {
        const long long pairs = num_sms() / 2;
        const long long tiles512 = Zout * ((m_blks + 1) / 2) * ((q->n + 511) / 512);
        const bool simple_epi = !q->out_fp32 && !q->accumulate && q->alpha == 1.0f && q->act == 0 && !q->relu_mask &&
                                !q->relu_bits && !q->relu_bits_out && !(q->bias && q->residual) &&
                                (reinterpret_cast<uintptr_t>(q->c) & 15) == 0 && q->c_batch_stride % 8 == 0 &&
                                (!q->residual || (reinterpret_cast<uintptr_t>(q->residual) & 15) == 0);
        const bool fits = q->k >= 4096 && q->n % 512 == 0 && tiles512 >= pairs;
        if (!a_mn && q->split_k == 1 && m_blks >= 2 && (q->cluster_m == 0 || q->cluster_m == 2) && simple_epi &&
            ((q->tile_n == 0 && fits) || q->tile_n == 512))
            bn = 512;
    }
    const long long n_blks = (q->n + bn - 1) / bn;
    const long long kb_per = (q->k + BK - 1) / BK;
    const long long total_kb = kb_per * (q->reduce_batch ? Zall : 1);
    const long long out_tiles = Zout * m_blks * n_blks;

    long long splits = q->split_k;
    if (splits <= 0) {
        splits = 1;
        if (out_tiles * 2 <= num_sms()) {
            // choose the split count that fills whole waves of SMs (each split keeps >= 8 k-blocks)
            double best = 0.0;
            const long long max_s = total_kb / 8 > 0 ? (total_kb / 8 < 32 ? total_kb / 8 : 32) : 1;
            for (long long s = 1; s <= max_s; ++s) {
                const long long t = out_tiles * s;
                const long long waves = (t + num_sms() - 1) / num_sms();
                const double eff = (double)t / (double)(waves * num_sms()) - 0.01 * (double)s;  // mild bias to fewer splits
                if (eff > best) { best = eff; splits = s; }
            }
        }
    }
    if (splits > total_kb) splits = total_kb;
    if (splits > 1 && !direct_acc) {   // partial tiles are reduce-added into ONE fp32 buffer
        const size_t need = (size_t)Zout * q->m * q->n * sizeof(float);
        if (!q->workspace || q->workspace_bytes < need) splits = 1;
    }
    long long kb_per_split = (total_kb + splits - 1) / splits;
    splits = (total_kb + kb_per_split - 1) / kb_per_split;
    if (splits > 1) {
        PB_CHECK(q->relu_mask == nullptr && q->residual == nullptr && q->act == 0 && q->bias == nullptr,
                 "split-K path supports only alpha/accumulate epilogues");
    }

    CUtensorMap ta, tb, tc;
    int rc;
    int epi = -1;
    if (splits > 1 || direct_acc) epi = 4;
    if (q->a_rowsum) {
        PB_CHECK(direct_acc && a_mn && Zout == 1 && bn != 512,
                 "a_rowsum needs the weight-gradient form: MN-major A, fp32 C with accumulate, no other epilogue option, one output batch");
        epi = 7;
    }
    else if (!q->out_fp32 && !q->accumulate && splits == 1 && q->c_batch_stride % 8 == 0 && q->alpha == 1.0f &&
        (reinterpret_cast<uintptr_t>(q->c) & 15) == 0 && (q->act == 0 || q->act == 1)) {
        if (q->residual) { if (!q->relu_mask && !q->relu_bits && !q->bias && q->act == 0 && (reinterpret_cast<uintptr_t>(q->residual) & 15) == 0) epi = 5; }
        else if (q->relu_bits) { if (!q->relu_mask && !q->bias && q->act == 0 && q->n % 32 == 0 && Zout == 1) epi = 6; }
        else if (q->relu_mask) { if (!q->bias && q->act == 0) epi = 3; }
        else if (q->bias) epi = q->act == 1 ? 2 : 1;
        else if (q->act == 0) epi = 0;
    }
    const bool tma_store = epi >= 0;
    if ((epi == 4 || epi == 7) && direct_acc) {
        rc = make_tmap(&tc, q->c, q->m, q->n, q->ldc, Zout, q->c_batch_stride, 32, 32, /*fp32=*/true);
        if (rc) return rc;
    } else if (epi == 4) {
        PB_CUDA(cudaMemsetAsync(q->workspace, 0, (size_t)Zout * q->m * q->n * sizeof(float), stream));
        rc = make_tmap(&tc, q->workspace, q->m, q->n, q->n, Zout, (long long)q->m * q->n, 32, 32, /*fp32=*/true);
        if (rc) return rc;
    } else if (tma_store) {
        rc = make_tmap(&tc, q->c, q->m, q->n, q->ldc, Zout, q->c_batch_stride, 64, 32);
        if (rc) return rc;
    } else {
        memset(&tc, 0, sizeof(tc));
    }
    // An operand shared by all batches (batch stride 0: the conv weight of the batched implicit GEMM) is described as ONE matrix:
    // a map with Zall slices would claim Zall x its size, i.e. memory past the end of the allocation (never addressed by the
    // kernel, which passes slice 0 for such an operand, but the extent a descriptor claims should be memory that exists).
    const long long Za = (Zall > 1 && q->a_batch_stride != 0) ? Zall : 1;
    const long long Zb = (Zall > 1 && q->b_batch_stride != 0) ? Zall : 1;
    if (!a_mn) rc = make_tmap(&ta, q->a, q->m, q->k, q->lda, Za, q->a_batch_stride, BK, BM);
    else rc = make_tmap(&ta, q->a, q->k, q->m, q->lda, Za, q->a_batch_stride, 64, BK);
    if (rc) return rc;
    // cluster size along M (B-tile multicast); 2 keeps every SM usable (4-CTA clusters strand SMs on this part)
    int cl = q->cluster_m > 0 ? q->cluster_m : 2;
    if (m_blks < 2) cl = 1;
    // small problems (decoder side: 3072 rows): independent CTAs start sooner than pairs and there is no operand traffic
    // to save when not even every SM has a tile (measured 6.5 vs 7.1 us for 3072 x 512 x 512)
    if (q->cluster_m == 0 && bn != 512 && out_tiles * splits <= num_sms()) cl = 1;
    if (cl == 4 && b_mn && bn == 128) cl = 2;
    if (cl != 1 && cl != 2 && cl != 4) cl = 1;
    if (epi == 7 && cl == 4) cl = 2;
    if (!b_mn) rc = make_tmap(&tb, q->b, q->n, q->k, q->ldb, Zb, q->b_batch_stride, BK, bn == 512 ? 128 : bn / cl);
    else rc = make_tmap(&tb, q->b, q->k, q->n, q->ldb, Zb, q->b_batch_stride, 64, BK);
    if (rc) return rc;

    GemmDev d;
    d.M = (int)q->m; d.N = (int)q->n; d.Z = (int)Zout;
    d.kb_per = (int)kb_per; d.total_kb = (int)total_kb; d.kb_per_split = (int)kb_per_split; d.splits = (int)splits;
    d.m_blks = (int)m_blks; d.n_blks = (int)n_blks;
    d.reduce_batch = q->reduce_batch ? 1 : 0;
    d.a_batched = (Zall > 1 && q->a_batch_stride != 0) ? 1 : 0;
    d.b_batched = (Zall > 1 && q->b_batch_stride != 0) ? 1 : 0;
    d.c = q->c; d.ldc = q->ldc; d.c_batch_stride = q->c_batch_stride;
    d.bias = reinterpret_cast<const bf16*>(q->bias);
    d.residual = reinterpret_cast<const bf16*>(q->residual);
    d.relu_mask = reinterpret_cast<const bf16*>(q->relu_mask);
    d.relu_bits_out = reinterpret_cast<uint32_t*>(q->relu_bits_out);
    d.relu_bits = reinterpret_cast<const uint32_t*>(q->relu_bits);
    d.relu_bits_ld = q->relu_bits_ld;
    PB_CHECK(q->relu_bits == nullptr || epi == 6, "relu_bits needs a plain bf16 GEMM with n %% 32 == 0 and a single batch");
    PB_CHECK(q->relu_bits_out == nullptr || (epi == 2 && q->n % 32 == 0 && Zout == 1),
             "relu_bits_out needs the bias+ReLU TMA epilogue with n %% 32 == 0 and a single batch");
    PB_CHECK((q->relu_bits == nullptr && q->relu_bits_out == nullptr) || q->relu_bits_ld >= q->m, "relu_bits_ld must cover the rows");
    d.act = q->act; d.out_fp32 = q->out_fp32; d.accumulate = q->accumulate; d.alpha = q->alpha;
    d.tma_store = tma_store ? 1 : 0;
    d.partial = reinterpret_cast<float*>(q->workspace);
    d.a_rowsum = q->a_rowsum;

    const long long m_groups = (m_blks + cl - 1) / cl;
    const long long cluster_tiles = Zout * m_groups * n_blks * splits;
    const long long max_clusters = num_sms() / cl;
    const int grid = (int)((cluster_tiles < max_clusters ? cluster_tiles : max_clusters) * cl);

#define PB_DISPATCH3(BN_, EPI_, CL_)                                                                         \
    do {                                                                                                      \
        if (!a_mn && !b_mn) rc = launch_gemm<BN_, false, false, EPI_, CL_>(ta, tb, tc, d, grid, stream);      \
        else if (!a_mn && b_mn) rc = launch_gemm<BN_, false, true, EPI_, CL_>(ta, tb, tc, d, grid, stream);   \
        else if (a_mn && !b_mn) rc = launch_gemm<BN_, true, false, EPI_, CL_>(ta, tb, tc, d, grid, stream);   \
        else rc = launch_gemm<BN_, true, true, EPI_, CL_>(ta, tb, tc, d, grid, stream);                       \
    } while (0)
#define PB_DISPATCH2(BN_, EPI_)                                   \
    do {                                                          \
        if (cl == 1) PB_DISPATCH3(BN_, EPI_, 1);                  \
        else if (cl == 2) PB_DISPATCH3(BN_, EPI_, 2);             \
        else PB_DISPATCH3(BN_, EPI_, 4);                          \
    } while (0)
#define PB_DISPATCH(BN_)                                          \
    do {                                                          \
        switch (epi) {                                            \
            case 0: PB_DISPATCH2(BN_, 0); break;                  \
            case 1: PB_DISPATCH2(BN_, 1); break;                  \
            case 2: PB_DISPATCH2(BN_, 2); break;                  \
            case 3: PB_DISPATCH2(BN_, 3); break;                  \
            case 4: PB_DISPATCH2(BN_, 4); break;                  \
            case 5: PB_DISPATCH2(BN_, 5); break;                  \
            case 6: PB_DISPATCH2(BN_, 6); break;                  \
            default: PB_DISPATCH2(BN_, -1); break;                \
        }                                                         \
    } while (0)
    if (epi == 7) {    // weight-gradient form only (MN-major A), CTA pair or single CTAs
        if (bn == 128) {
            if (cl == 1) { if (b_mn) rc = launch_gemm<128, true, true, 7, 1>(ta, tb, tc, d, grid, stream); else rc = launch_gemm<128, true, false, 7, 1>(ta, tb, tc, d, grid, stream); }
            else { if (b_mn) rc = launch_gemm<128, true, true, 7, 2>(ta, tb, tc, d, grid, stream); else rc = launch_gemm<128, true, false, 7, 2>(ta, tb, tc, d, grid, stream); }
        } else {
            if (cl == 1) { if (b_mn) rc = launch_gemm<256, true, true, 7, 1>(ta, tb, tc, d, grid, stream); else rc = launch_gemm<256, true, false, 7, 1>(ta, tb, tc, d, grid, stream); }
            else { if (b_mn) rc = launch_gemm<256, true, true, 7, 2>(ta, tb, tc, d, grid, stream); else rc = launch_gemm<256, true, false, 7, 2>(ta, tb, tc, d, grid, stream); }
        }
    } else if (bn == 512) {   // CTA pair, K-major A, plain / bias / residual epilogues only (see the tile choice above)
        PB_CHECK(cl == 2 && !a_mn && (epi == 0 || epi == 1 || epi == 5), "internal: 512-wide tile chosen for an unsupported variant");
        if (!b_mn) {
            if (epi == 0) rc = launch_gemm<512, false, false, 0, 2>(ta, tb, tc, d, grid, stream);
            else if (epi == 1) rc = launch_gemm<512, false, false, 1, 2>(ta, tb, tc, d, grid, stream);
            else rc = launch_gemm<512, false, false, 5, 2>(ta, tb, tc, d, grid, stream);
        } else {
            if (epi == 0) rc = launch_gemm<512, false, true, 0, 2>(ta, tb, tc, d, grid, stream);
            else if (epi == 1) rc = launch_gemm<512, false, true, 1, 2>(ta, tb, tc, d, grid, stream);
            else rc = launch_gemm<512, false, true, 5, 2>(ta, tb, tc, d, grid, stream);
        }
    } else if (bn == 128) PB_DISPATCH(128); else PB_DISPATCH(256);
#undef PB_DISPATCH3
#undef PB_DISPATCH2
#undef PB_DISPATCH
    if (rc) return rc;

    if (splits > 1 && !direct_acc) {
        const long long zmn = Zout * q->m * q->n;
        const int threads = 256;
        long long blocks = (zmn / 4 + threads - 1) / threads;
        if (blocks > 4LL * num_sms()) blocks = 4LL * num_sms();
        splitk_reduce_kernel<<<(int)blocks, threads, 0, stream>>>(d.partial, 1, zmn, d.M, d.N, q->c, q->ldc,
                                                                 q->c_batch_stride, q->out_fp32, q->accumulate,
                                                                 q->alpha, nullptr);
        PB_LAUNCH_CHECK();
    }
    return 0;
}
