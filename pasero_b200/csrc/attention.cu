// pasero_b200 — fused multi-head attention (forward + backward), head_dim 64, bf16 in / fp32 softmax.
//
// Online-softmax attention that takes key lengths instead of a materialised mask tensor:
//   * encoder self-attention and decoder cross-attention: key-padding mask from `kv_len[b]`
//     (reference builds a float -inf mask (B,1,1,S) and calls SDPA: pasero/models/modules.py:659-677,712-718)
//   * decoder self-attention: causal, padding mask dropped (modules.py:602-605, 688)
// Q/K/V are read in place from the fused projection outputs ([B*L, ld] rows, head h at column h*64), the
// output is written as [B*Lq, H*64] so the out-projection GEMM consumes it without a transpose.
// Backward recomputes P from the saved log-sum-exp (no O(L^2) tensor is stored): one kernel owns a KV
// block and accumulates dK/dV over query blocks, one owns a query block and accumulates dQ.
//
// Round-1 implementation uses mma.sync.m16n8k16 (HMMA) tiles with cp.async staging; the tcgen05/TMEM
// version is the next step for this kernel.
#include "common.cuh"
#include "../../include/pasero_b200.h"

#include <stdlib.h>
#include <string.h>

namespace pb200 {

static constexpr int HD = 64;      // head dim
static constexpr int BLK = 64;     // query / key block
static constexpr int ATT_THREADS = 128;
static constexpr float LOG2E = 1.4426950408889634f;

struct AttnParams {
    const bf16* q; const bf16* k; const bf16* v;
    bf16* o;
    float* lse;          // [B, H, Lq]
    const int* kv_len;   // [B] or null
    long long ldq, ldk, ldv, ldo;
    int B, H, Lq, Lk;
    int causal;
    float scale;
    // backward
    const bf16* dout; long long lddo;
    const float* delta;  // [B, H, Lq]
    bf16* dq; bf16* dk; bf16* dv;
    long long lddq, lddk, lddv;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// tile [64 rows][64 bf16], 16-byte chunks XOR-swizzled by (row & 7)
__device__ __forceinline__ bf16* tile_ptr(bf16* tile, int row, int chunk) {
    return tile + row * HD + ((chunk ^ (row & 7)) << 3);
}
// rows [row0, row0+64) of a [B*L, ld] matrix, head column offset already applied to `g`
__device__ __forceinline__ void load_tile(bf16* tile, const bf16* g, long long ld, int row0, int L) {
    for (int i = threadIdx.x; i < BLK * 8; i += ATT_THREADS) {
        const int r = i >> 3, c = i & 7;
        const int row = row0 + r;
        const bool ok = row < L;
        cp_async16(tile_ptr(tile, r, c), g + (long long)(ok ? row : 0) * ld + c * 8, ok ? 16 : 0);
    }
}
// A fragments (16 rows x 64 k) of a tile for this warp's rows
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[4][4], bf16* tile, int r0, int lane) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(f[ks], tile_ptr(tile, r0 + (lane & 15), 2 * ks + (lane >> 4)));
}
// C[16 x 64] += A[16 x 64] * T^T where tile T is stored [n][k] (k contiguous)
__device__ __forceinline__ void mma_nt(float (&c)[8][4], const uint32_t (&a)[4][4], bf16* tile, int lane) {
    const int id = lane >> 3, l7 = lane & 7;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            ldsm_x4(b, tile_ptr(tile, 16 * np + (id >> 1) * 8 + l7, 2 * ks + (id & 1)));
            mma16816(c[2 * np], a[ks], b[0], b[1]);
            mma16816(c[2 * np + 1], a[ks], b[2], b[3]);
        }
}
// C[16 x 64] += P[16 x 64] * T where tile T is stored [k][n] (n contiguous); P given as fp32 C-fragments
__device__ __forceinline__ void mma_pn(float (&c)[8][4], const float (&p)[8][4], bf16* tile, int lane) {
    const int id = lane >> 3, l7 = lane & 7;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[4];
        a[0] = pack_bf16x2(p[2 * ks][0], p[2 * ks][1]);
        a[1] = pack_bf16x2(p[2 * ks][2], p[2 * ks][3]);
        a[2] = pack_bf16x2(p[2 * ks + 1][0], p[2 * ks + 1][1]);
        a[3] = pack_bf16x2(p[2 * ks + 1][2], p[2 * ks + 1][3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            ldsm_x4_t(b, tile_ptr(tile, 16 * ks + (id & 1) * 8 + l7, 2 * np + (id >> 1)));
            mma16816(c[2 * np], a, b[0], b[1]);
            mma16816(c[2 * np + 1], a, b[2], b[3]);
        }
    }
}
__device__ __forceinline__ void zero_acc(float (&c)[8][4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
}
// store a 16 x 64 C-fragment block (scaled) to rows [row0+..] of a [.., ld] bf16 matrix
__device__ __forceinline__ void store_c(bf16* g, long long ld, int row0, int L, const float (&c)[8][4], float mul,
                                        int lane) {
    const int gq = lane >> 2, t = lane & 3;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int row = row0 + gq + half * 8;
        if (row >= L) continue;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
            *reinterpret_cast<uint32_t*>(g + (long long)row * ld + nt * 8 + 2 * t) =
                pack_bf16x2(c[nt][2 * half] * mul, c[nt][2 * half + 1] * mul);
    }
}

// =====================================================================================================
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const AttnParams p) {
    __shared__ __align__(128) bf16 sQ[BLK * HD];
    __shared__ __align__(128) bf16 sK[BLK * HD];
    __shared__ __align__(128) bf16 sV[BLK * HD];
    const int q0 = blockIdx.x * BLK, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gq = lane >> 2, t = lane & 3;
    const int kvlen = p.kv_len ? min(p.kv_len[b], p.Lk) : p.Lk;
    const int coff = p.Lk - p.Lq;  // causal: key j visible to query i iff j <= i + coff
    const int kv_hi = p.causal ? min(kvlen, q0 + BLK + coff) : kvlen;
    const int n_blocks = (kv_hi + BLK - 1) / BLK;
    const bf16* gq_ = p.q + (long long)b * p.Lq * p.ldq + h * HD;
    const bf16* gk_ = p.k + (long long)b * p.Lk * p.ldk + h * HD;
    const bf16* gv_ = p.v + (long long)b * p.Lk * p.ldv + h * HD;
    const float sl2 = p.scale * LOG2E;

    load_tile(sQ, gq_, p.ldq, q0, p.Lq);
    uint32_t qf[4][4];
    float o[8][4];
    zero_acc(o);
    float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};

    for (int nb = 0; nb < n_blocks; ++nb) {
        __syncthreads();
        load_tile(sK, gk_, p.ldk, nb * BLK, p.Lk);
        load_tile(sV, gv_, p.ldv, nb * BLK, p.Lk);
        cp_async_wait_all();
        __syncthreads();
        if (nb == 0) load_a_frags(qf, sQ, warp * 16, lane);
        float s[8][4];
        zero_acc(s);
        mma_nt(s, qf, sK, lane);
        float mx[2] = {m_i[0], m_i[1]};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = nb * BLK + nt * 8 + 2 * t + (e & 1);
                const int row = q0 + warp * 16 + gq + (e >> 1) * 8;
                if (col >= kvlen || (p.causal && col > row + coff)) s[nt][e] = -INFINITY;
                mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
            }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float alpha[2], msc[2], rs[2] = {0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float m_use = (mx[r] == -INFINITY) ? 0.f : mx[r];
            alpha[r] = exp2f((m_i[r] - m_use) * sl2);
            msc[r] = m_use * sl2;
            m_i[r] = mx[r];
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float pv = exp2f(s[nt][e] * sl2 - msc[e >> 1]);
                s[nt][e] = pv;
                rs[e >> 1] += pv;
            }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_i[r] = l_i[r] * alpha[r] + rs[r];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            o[nt][0] *= alpha[0]; o[nt][1] *= alpha[0];
            o[nt][2] *= alpha[1]; o[nt][3] *= alpha[1];
        }
        mma_pn(o, s, sV, lane);
    }
    if (n_blocks == 0) { cp_async_wait_all(); }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
        l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
    }
    const float inv0 = l_i[0] > 0.f ? 1.f / l_i[0] : 0.f, inv1 = l_i[1] > 0.f ? 1.f / l_i[1] : 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        o[nt][0] *= inv0; o[nt][1] *= inv0; o[nt][2] *= inv1; o[nt][3] *= inv1;
    }
    store_c(p.o + (long long)b * p.Lq * p.ldo + h * HD, p.ldo, q0 + warp * 16, p.Lq, o, 1.0f, lane);
    if (p.lse && t == 0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = q0 + warp * 16 + gq + r * 8;
            if (row < p.Lq)
                p.lse[((long long)b * p.H + h) * p.Lq + row] =
                    (l_i[r] > 0.f) ? m_i[r] * p.scale + logf(l_i[r]) : -INFINITY;
        }
    }
}

// delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]
__global__ void attn_delta_kernel(const bf16* __restrict__ o, long long ldo, const bf16* __restrict__ dout,
                                  long long lddo, int B, int H, int Lq, float* __restrict__ delta) {
    const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long total = (long long)B * H * Lq;
    if (wid >= total) return;
    const int i = (int)(wid % Lq);
    const int h = (int)((wid / Lq) % H);
    const int b = (int)(wid / ((long long)Lq * H));
    const long long row = (long long)b * Lq + i;
    const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o + row * ldo + h * HD + 2 * lane));
    const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dout + row * lddo + h * HD + 2 * lane));
    const float s = warp_sum(a.x * d.x + a.y * d.y);
    if (lane == 0) delta[wid] = s;
}

// dK, dV for one KV block: loops over query blocks. Works on the transposed problem (rows = keys).
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dkv_kernel(const AttnParams p) {
    __shared__ __align__(128) bf16 sK[BLK * HD];
    __shared__ __align__(128) bf16 sV[BLK * HD];
    __shared__ __align__(128) bf16 sQ[BLK * HD];
    __shared__ __align__(128) bf16 sdO[BLK * HD];
    __shared__ float sLse[BLK], sDelta[BLK];
    const int n0 = blockIdx.x * BLK, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gq = lane >> 2, t = lane & 3;
    const int kvlen = p.kv_len ? min(p.kv_len[b], p.Lk) : p.Lk;
    const int coff = p.Lk - p.Lq;
    const bf16* gq_ = p.q + (long long)b * p.Lq * p.ldq + h * HD;
    const bf16* gk_ = p.k + (long long)b * p.Lk * p.ldk + h * HD;
    const bf16* gv_ = p.v + (long long)b * p.Lk * p.ldv + h * HD;
    const bf16* gdo_ = p.dout + (long long)b * p.Lq * p.lddo + h * HD;
    const float* lse_ = p.lse + ((long long)b * p.H + h) * p.Lq;
    const float* delta_ = p.delta + ((long long)b * p.H + h) * p.Lq;
    const float sl2 = p.scale * LOG2E;

    load_tile(sK, gk_, p.ldk, n0, p.Lk);
    load_tile(sV, gv_, p.ldv, n0, p.Lk);
    cp_async_wait_all();
    __syncthreads();
    uint32_t kf[4][4], vf[4][4];
    load_a_frags(kf, sK, warp * 16, lane);
    load_a_frags(vf, sV, warp * 16, lane);
    float dk[8][4], dv[8][4];
    zero_acc(dk); zero_acc(dv);

    const int m_blocks = (p.Lq + BLK - 1) / BLK;
    int mb_start = 0;
    if (p.causal) { const int first_q = n0 - coff; mb_start = first_q > 0 ? first_q / BLK : 0; }
    const bool any_valid = n0 < kvlen;
    for (int mb = mb_start; mb < m_blocks && any_valid; ++mb) {
        const int m0 = mb * BLK;
        __syncthreads();
        load_tile(sQ, gq_, p.ldq, m0, p.Lq);
        load_tile(sdO, gdo_, p.lddo, m0, p.Lq);
        if (threadIdx.x < BLK) {
            const int row = m0 + threadIdx.x;
            sLse[threadIdx.x] = row < p.Lq ? lse_[row] * LOG2E : INFINITY;
            sDelta[threadIdx.x] = row < p.Lq ? delta_[row] : 0.f;
        }
        cp_async_wait_all();
        __syncthreads();
        float st[8][4], dpt[8][4];
        zero_acc(st); zero_acc(dpt);
        mma_nt(st, kf, sQ, lane);     // S^T = K Q^T   (rows: keys, cols: queries)
        mma_nt(dpt, vf, sdO, lane);   // dP^T = V dO^T
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int qc = nt * 8 + 2 * t + (e & 1);
                const int key = n0 + warp * 16 + gq + (e >> 1) * 8;
                const int qrow = m0 + qc;
                float pv = exp2f(st[nt][e] * sl2 - sLse[qc]);
                if (key >= kvlen || (p.causal && key > qrow + coff)) pv = 0.f;
                st[nt][e] = pv;
                dpt[nt][e] = pv * (dpt[nt][e] - sDelta[qc]);
            }
        mma_pn(dv, st, sdO, lane);    // dV += P^T dO
        mma_pn(dk, dpt, sQ, lane);    // dK += dS^T Q
    }
    store_c(p.dk + (long long)b * p.Lk * p.lddk + h * HD, p.lddk, n0 + warp * 16, p.Lk, dk, p.scale, lane);
    store_c(p.dv + (long long)b * p.Lk * p.lddv + h * HD, p.lddv, n0 + warp * 16, p.Lk, dv, 1.0f, lane);
}

// dQ for one query block: loops over KV blocks.
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dq_kernel(const AttnParams p) {
    __shared__ __align__(128) bf16 sQ[BLK * HD];
    __shared__ __align__(128) bf16 sdO[BLK * HD];
    __shared__ __align__(128) bf16 sK[BLK * HD];
    __shared__ __align__(128) bf16 sV[BLK * HD];
    const int q0 = blockIdx.x * BLK, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gq = lane >> 2, t = lane & 3;
    const int kvlen = p.kv_len ? min(p.kv_len[b], p.Lk) : p.Lk;
    const int coff = p.Lk - p.Lq;
    const int kv_hi = p.causal ? min(kvlen, q0 + BLK + coff) : kvlen;
    const int n_blocks = (kv_hi + BLK - 1) / BLK;
    const bf16* gq_ = p.q + (long long)b * p.Lq * p.ldq + h * HD;
    const bf16* gk_ = p.k + (long long)b * p.Lk * p.ldk + h * HD;
    const bf16* gv_ = p.v + (long long)b * p.Lk * p.ldv + h * HD;
    const bf16* gdo_ = p.dout + (long long)b * p.Lq * p.lddo + h * HD;
    const float sl2 = p.scale * LOG2E;

    load_tile(sQ, gq_, p.ldq, q0, p.Lq);
    load_tile(sdO, gdo_, p.lddo, q0, p.Lq);
    float lse_r[2], delta_r[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = q0 + warp * 16 + gq + r * 8;
        const long long idx = ((long long)b * p.H + h) * p.Lq + row;
        lse_r[r] = row < p.Lq ? p.lse[idx] * LOG2E : INFINITY;
        delta_r[r] = row < p.Lq ? p.delta[idx] : 0.f;
    }
    uint32_t qf[4][4], dof[4][4];
    float dq[8][4];
    zero_acc(dq);
    for (int nb = 0; nb < n_blocks; ++nb) {
        __syncthreads();
        load_tile(sK, gk_, p.ldk, nb * BLK, p.Lk);
        load_tile(sV, gv_, p.ldv, nb * BLK, p.Lk);
        cp_async_wait_all();
        __syncthreads();
        if (nb == 0) { load_a_frags(qf, sQ, warp * 16, lane); load_a_frags(dof, sdO, warp * 16, lane); }
        float s[8][4], dp[8][4];
        zero_acc(s); zero_acc(dp);
        mma_nt(s, qf, sK, lane);      // S = Q K^T
        mma_nt(dp, dof, sV, lane);    // dP = dO V^T
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = nb * BLK + nt * 8 + 2 * t + (e & 1);
                const int row = q0 + warp * 16 + gq + (e >> 1) * 8;
                float pv = exp2f(s[nt][e] * sl2 - lse_r[e >> 1]);
                if (col >= kvlen || (p.causal && col > row + coff)) pv = 0.f;
                s[nt][e] = pv * (dp[nt][e] - delta_r[e >> 1]);
            }
        mma_pn(dq, s, sK, lane);      // dQ += dS K
    }
    if (n_blocks == 0) { cp_async_wait_all(); }
    store_c(p.dq + (long long)b * p.Lq * p.lddq + h * HD, p.lddq, q0 + warp * 16, p.Lq, dq, p.scale, lane);
}

static int check_attn(const pb200_attn_params* q) {
    PB_CHECK(q != nullptr, "null params");
    PB_CHECK(q->head_dim == HD, "attention kernels support head_dim 64 only (got %d)", q->head_dim);
    PB_CHECK(q->ldq % 8 == 0 && q->ldk % 8 == 0 && q->ldv % 8 == 0 && q->ldo % 8 == 0, "attention strides must be multiples of 8");
    PB_CHECK(q->B >= 0 && q->H > 0 && q->Lq >= 0 && q->Lk >= 0, "bad attention shape");
    return 0;
}
static AttnParams to_dev(const pb200_attn_params* q) {
    AttnParams p;
    p.q = (const bf16*)q->q; p.k = (const bf16*)q->k; p.v = (const bf16*)q->v; p.o = (bf16*)q->o;
    p.lse = q->lse; p.kv_len = q->kv_len;
    p.ldq = q->ldq; p.ldk = q->ldk; p.ldv = q->ldv; p.ldo = q->ldo;
    p.B = q->B; p.H = q->H; p.Lq = q->Lq; p.Lk = q->Lk; p.causal = q->causal; p.scale = q->scale;
    p.dout = (const bf16*)q->dout; p.lddo = q->lddo; p.delta = q->delta;
    p.dq = (bf16*)q->dq; p.dk = (bf16*)q->dk; p.dv = (bf16*)q->dv;
    p.lddq = q->lddq; p.lddk = q->lddk; p.lddv = q->lddv;
    return p;
}

// tcgen05 path (attention_tc.cu)
bool attn_tc_supported(const pb200_attn_params* q, bool backward);
int attn_fwd_tc(const pb200_attn_params* q, cudaStream_t stream, bool p_in_tmem);
int attn_bwd_tc(const pb200_attn_params* q, cudaStream_t stream);

// PB200_ATTN selects the kernel family for experiments: "hmma" forces the mma.sync kernels, "tc_ss" the tcgen05
// forward with P staged in shared memory; default = tcgen05 (P kept in TMEM) whenever the shape fits.
static int attn_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("PB200_ATTN");
        mode = 0;
        if (e && !strcmp(e, "hmma")) mode = 1;
        else if (e && !strcmp(e, "tc_ss")) mode = 2;
    }
    return mode;
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_attention_fwd(const pb200_attn_params* q, void* stream_) {
    if (int rc = check_attn(q)) return rc;
    if (q->B == 0 || q->Lq == 0) return 0;
    PB_CHECK(q->q && q->k && q->v && q->o, "attention_fwd: null tensor");
    if (attn_mode() != 1 && attn_tc_supported(q, false)) return attn_fwd_tc(q, (cudaStream_t)stream_, attn_mode() != 2);
    const AttnParams p = to_dev(q);
    dim3 grid((p.Lq + BLK - 1) / BLK, p.H, p.B);
    attn_fwd_kernel<<<grid, ATT_THREADS, 0, (cudaStream_t)stream_>>>(p);
    PB_LAUNCH_CHECK();
    return 0;
}

extern "C" int pb200_attention_bwd(const pb200_attn_params* q, void* stream_) {
    if (int rc = check_attn(q)) return rc;
    if (q->B == 0 || q->Lq == 0) return 0;
    PB_CHECK(q->q && q->k && q->v && q->o && q->dout && q->lse && q->delta && q->dq && q->dk && q->dv,
             "attention_bwd: null tensor");
    PB_CHECK(q->lddo % 8 == 0 && q->lddq % 8 == 0 && q->lddk % 8 == 0 && q->lddv % 8 == 0, "attention strides must be multiples of 8");
    if (attn_mode() != 1 && attn_tc_supported(q, true)) return attn_bwd_tc(q, (cudaStream_t)stream_);
    const AttnParams p = to_dev(q);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long warps = (long long)p.B * p.H * p.Lq;
    attn_delta_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>(p.o, p.ldo, p.dout, p.lddo, p.B, p.H, p.Lq,
                                                                               const_cast<float*>(p.delta));
    PB_LAUNCH_CHECK();
    dim3 gkv((p.Lk + BLK - 1) / BLK, p.H, p.B);
    attn_bwd_dkv_kernel<<<gkv, ATT_THREADS, 0, stream>>>(p);
    PB_LAUNCH_CHECK();
    dim3 gq_((p.Lq + BLK - 1) / BLK, p.H, p.B);
    attn_bwd_dq_kernel<<<gq_, ATT_THREADS, 0, stream>>>(p);
    PB_LAUNCH_CHECK();
    return 0;
}
