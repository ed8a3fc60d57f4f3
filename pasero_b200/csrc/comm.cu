// pasero_b200 — gradient allreduce over NVLink peer memory (one node, <= 8 ranks, one process per GPU).
//
// Replaces the NCCL allreduce that torch DDP issues per bucket (pasero/training.py:243-250) when every rank's bf16
// gradient arena has been mapped into every other rank (CUDA IPC). One launch reduces one bucket in place:
//
//   start barrier   every rank's bucket is final (its producer kernels precede this launch in stream order)
//   reduce + share  rank r owns slice r of the bucket: 16-byte chunks are read from all ranks' arenas (peer loads over
//                   NVLink), summed in fp32 in rank order, rounded once to bf16 and written back into ALL ranks' arenas
//                   (peer stores). Slice r of a peer's arena is read only by rank r, so overwriting it in place is safe.
//   end barrier     every rank has finished reading from and writing into this rank's arena
//
// Per rank (N - 1) / N of the bucket crosses NVLink once in each direction — the two-shot minimum — and every value is
// the same fp32 sum rounded once on exactly one rank, so replicas stay bit-identical. The barriers are per block
// (block b of every rank synchronises with block b of the others through flags in peer memory, generation counters kept
// on the device), so the kernel has no host-side state and can be captured in a CUDA graph and overlapped with backward.
#include "common.cuh"
#include "../../include/pasero_b200.h"

#include <cuda.h>
#include <string.h>

namespace pb200 {

static constexpr int P2P_MAX_RANKS = 8;
static constexpr int P2P_BLOCKS = 32;        // co-resident with the GEMM CTAs (no dynamic shared memory): enough for NVLink
static constexpr int P2P_THREADS = 512;
static constexpr long long P2P_TIMEOUT_CYCLES = 20000000000LL;   // ~10 s at 1.9 GHz

struct P2PDev {
    bf16* data[P2P_MAX_RANKS];               // every rank's gradient arena, as mapped in this process
    uint32_t* sig[P2P_MAX_RANKS];            // every rank's signal pad: [2][P2P_BLOCKS][P2P_MAX_RANKS] flags + [P2P_BLOCKS] generations
                                             // + 1 word: number of barrier waits that timed out (a hang becomes an error)
    int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_peer16(const void* p) {   // peer / own arena: never through the non-coherent L1
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer16(void* p, uint4 v) {
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// flag (phase, block, writer) inside a signal pad
__device__ __forceinline__ uint32_t* flag_of(uint32_t* pad, int phase, int block, int writer) {
    return pad + ((phase * P2P_BLOCKS + block) * P2P_MAX_RANKS + writer);
}

// Block-level barrier across ranks: post generation `gen` into every peer's pad, wait until every peer posted it here.
// Generations only grow, so a peer that is already one launch ahead still satisfies the wait.
__device__ __forceinline__ void rank_barrier(const P2PDev& c, int phase, uint32_t gen) {
    __syncthreads();
    if ((int)threadIdx.x < c.world) {
        const int peer = threadIdx.x;
        __threadfence_system();
        st_release_sys(flag_of(c.sig[peer], phase, blockIdx.x, c.rank), gen);
        const uint32_t* mine = flag_of(c.sig[c.rank], phase, blockIdx.x, peer);
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(mine) - gen) < 0) {
            if (clock64() - t0 > P2P_TIMEOUT_CYCLES) {   // a peer never arrived: give up instead of hanging the GPU
                atomicAdd(c.sig[c.rank] + 2 * P2P_BLOCKS * P2P_MAX_RANKS + P2P_BLOCKS, 1u);
                break;
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(P2P_THREADS)
allreduce_p2p_kernel(const P2PDev c, long long begin, long long count) {
    uint32_t* gens = c.sig[c.rank] + 2 * P2P_BLOCKS * P2P_MAX_RANKS;
    const uint32_t gen = gens[blockIdx.x] + 1;          // read by every thread before thread 0 bumps it (barrier in between)
    rank_barrier(c, 0, gen);

    // slice of this rank: chunks of 8 bf16 (16 bytes)
    const long long chunks = count >> 3;
    const long long per = (chunks + c.world - 1) / c.world;
    const long long c0 = per * c.rank;
    const long long c1 = c0 + per < chunks ? c0 + per : chunks;
    for (long long i = c0 + (long long)blockIdx.x * P2P_THREADS + threadIdx.x; i < c1; i += (long long)P2P_BLOCKS * P2P_THREADS) {
        const long long e = begin + (i << 3);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint4 v[P2P_MAX_RANKS];
#pragma unroll
        for (int r = 0; r < P2P_MAX_RANKS; ++r)
            if (r < c.world) v[r] = ld_peer16(c.data[r] + e);
#pragma unroll
        for (int r = 0; r < P2P_MAX_RANKS; ++r) {
            if (r < c.world) {
                const uint32_t w[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack_bf16x2(w[j]);
                    acc[2 * j] += f.x; acc[2 * j + 1] += f.y;
                }
            }
        }
        uint4 o;
        o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
        o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
#pragma unroll
        for (int r = 0; r < P2P_MAX_RANKS; ++r)
            if (r < c.world) st_peer16(c.data[r] + e, o);
    }
    __threadfence_system();
    rank_barrier(c, 1, gen);
    if (threadIdx.x == 0) gens[blockIdx.x] = gen;
}

// NVLS variant: the arenas are bound to one NVSwitch multicast object (torch symmetric memory hands out its address).
// multimem.ld_reduce makes the SWITCH add the 16-byte chunk of all ranks (fp32 accumulation) and return the sum;
// multimem.st broadcasts the rounded result into every rank's arena. Per rank only 1/N of the bucket crosses its links in
// each direction (the two-shot kernel above moves (N - 1)/N). Same slice ownership, same flag barriers (unicast pads).
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const void* mc) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_16(void* mc, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void __launch_bounds__(P2P_THREADS)
allreduce_nvls_kernel(const P2PDev c, bf16* __restrict__ mc, long long begin, long long count) {
    uint32_t* gens = c.sig[c.rank] + 2 * P2P_BLOCKS * P2P_MAX_RANKS;
    const uint32_t gen = gens[blockIdx.x] + 1;
    rank_barrier(c, 0, gen);
    const long long chunks = count >> 3;
    const long long per = (chunks + c.world - 1) / c.world;
    const long long c0 = per * c.rank;
    const long long c1 = c0 + per < chunks ? c0 + per : chunks;
    for (long long i = c0 + (long long)blockIdx.x * P2P_THREADS + threadIdx.x; i < c1; i += (long long)P2P_BLOCKS * P2P_THREADS) {
        bf16* p = mc + begin + (i << 3);
        multimem_st_16(p, multimem_ld_reduce_bf16x8(p));
    }
    __threadfence_system();
    rank_barrier(c, 1, gen);
    if (threadIdx.x == 0) gens[blockIdx.x] = gen;
}

typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

}  // namespace pb200

using namespace pb200;

extern "C" size_t pb200_p2p_signal_bytes(void) {
    return (size_t)(2 * P2P_BLOCKS * P2P_MAX_RANKS + P2P_BLOCKS + 1) * sizeof(uint32_t);
}

// IPC handle (64 bytes) of the allocation that holds `ptr`, and ptr's byte offset inside it
extern "C" int pb200_p2p_export(const void* ptr, void* handle_out, int64_t* offset_out) {
    PB_CHECK(ptr && handle_out && offset_out, "p2p_export: null argument");
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PB_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
    PB_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuMemGetAddressRange entry point not available");
    CUdeviceptr base = 0;
    size_t size = 0;
    CUresult r = reinterpret_cast<GetAddressRangeFn>(fn)(&base, &size, (CUdeviceptr)ptr);
    PB_CHECK(r == CUDA_SUCCESS, "cuMemGetAddressRange failed (%d)", (int)r);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    PB_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_out), reinterpret_cast<void*>(base)));
    *offset_out = (int64_t)((CUdeviceptr)ptr - base);
    return 0;
}

// maps a peer's allocation into this process; *ptr_out = mapped base + offset
extern "C" int pb200_p2p_import(const void* handle, int64_t offset, void** ptr_out) {
    PB_CHECK(handle && ptr_out, "p2p_import: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* base = nullptr;
    PB_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out = static_cast<char*>(base) + offset;
    return 0;
}

// In-place sum of arena elements [begin, begin + count) over all ranks; begin and count multiples of 8.
// data[r] / signal[r]: rank r's gradient arena / signal pad as mapped here (own pointers at index `rank`); the signal
// pads must be zero-initialised once, before the first call on any rank.
extern "C" int pb200_allreduce_p2p(void* const* data, void* const* signal, int32_t rank, int32_t world, int64_t begin,
                                   int64_t count, void* stream_) {
    PB_CHECK(world >= 1 && world <= P2P_MAX_RANKS && rank >= 0 && rank < world, "allreduce_p2p: bad rank/world (%d/%d)", rank, world);
    PB_CHECK(begin % 8 == 0 && count % 8 == 0, "allreduce_p2p: begin and count must be multiples of 8 elements");
    if (count == 0 || world == 1) return 0;
    P2PDev c;
    for (int r = 0; r < P2P_MAX_RANKS; ++r) {
        c.data[r] = r < world ? reinterpret_cast<bf16*>(data[r]) : nullptr;
        c.sig[r] = r < world ? reinterpret_cast<uint32_t*>(signal[r]) : nullptr;
        PB_CHECK(r >= world || (c.data[r] && c.sig[r]), "allreduce_p2p: null peer pointer for rank %d", r);
    }
    c.rank = rank; c.world = world;
    allreduce_p2p_kernel<<<P2P_BLOCKS, P2P_THREADS, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(c, (long long)begin, (long long)count);
    PB_LAUNCH_CHECK();
    return 0;
}

// The same reduction through the NVSwitch multicast mapping `mc_data` of the arenas (NVLS). signal[r]: unicast signal pads
// as for pb200_allreduce_p2p; data pointers are not needed.
extern "C" int pb200_allreduce_nvls(void* mc_data, void* const* signal, int32_t rank, int32_t world, int64_t begin,
                                    int64_t count, void* stream_) {
    PB_CHECK(world >= 1 && world <= P2P_MAX_RANKS && rank >= 0 && rank < world, "allreduce_nvls: bad rank/world (%d/%d)", rank, world);
    PB_CHECK(begin % 8 == 0 && count % 8 == 0, "allreduce_nvls: begin and count must be multiples of 8 elements");
    PB_CHECK(mc_data != nullptr, "allreduce_nvls: no multicast mapping");
    if (count == 0 || world == 1) return 0;
    P2PDev c;
    for (int r = 0; r < P2P_MAX_RANKS; ++r) {
        c.data[r] = nullptr;
        c.sig[r] = r < world ? reinterpret_cast<uint32_t*>(signal[r]) : nullptr;
        PB_CHECK(r >= world || c.sig[r], "allreduce_nvls: null signal pad for rank %d", r);
    }
    c.rank = rank; c.world = world;
    allreduce_nvls_kernel<<<P2P_BLOCKS, P2P_THREADS, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
        c, reinterpret_cast<bf16*>(mc_data), (long long)begin, (long long)count);
    PB_LAUNCH_CHECK();
    return 0;
}
