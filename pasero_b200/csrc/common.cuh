// pasero_b200 — shared device/host helpers for the sm_100a kernels.
#pragma once
#ifdef PB200_HALF
#include "f16_symbols.h"   // float16 build: namespace pb200 -> pb200_h, exported names get the suffix _f16
#endif
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

// library-wide plumbing shared by the bfloat16 and the float16 build of the kernels (runtime.cu, compiled once)
namespace pb200_rt {
void set_error(const char* fmt, ...);
int num_sms();
bool pdl_enabled();
}  // namespace pb200_rt

namespace pb200 {
using pb200_rt::set_error;
using pb200_rt::num_sms;
using pb200_rt::pdl_enabled;

// ---------------------------------------------------------------------------------------------
// error reporting (C-ABI: every entry point returns 0 or a non-zero code; the text of the last
// failure on this thread is available from pb200_last_error_string())
// ---------------------------------------------------------------------------------------------
#define PB_CHECK(cond, ...)                                  \
    do {                                                     \
        if (!(cond)) {                                       \
            ::pb200_rt::set_error(__VA_ARGS__);                 \
            return 1;                                        \
        }                                                    \
    } while (0)

#define PB_CUDA(expr)                                                                     \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::pb200_rt::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                               __FILE__, __LINE__);                                       \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)

#define PB_LAUNCH_CHECK()                                                                 \
    do {                                                                                  \
        cudaError_t _e = cudaPeekAtLastError();                                           \
        if (_e != cudaSuccess) {                                                          \
            ::pb200_rt::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),\
                               __FILE__, __LINE__);                                       \
            return 3;                                                                     \
        }                                                                                 \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL). Kernels launched through launch_pdl() may be scheduled while the previous
// kernel of the stream is still draining: everything before pdl_wait() (barrier / TMEM / descriptor set-up, index
// arithmetic) overlaps the predecessor's tail, pdl_wait() returns once the predecessor has completed and its writes
// are visible. RULE: no global-memory access (read OR write) before pdl_wait(). pdl_trigger() lets the successor's
// CTAs be scheduled as soon as every CTA of this grid has started.  PB200_PDL=0 disables the launch attribute.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
// The 16-bit element type of activations, weights and gradients. The kernels are written against the names below; the
// library holds two builds of them: bfloat16 (the default; `dtype: bfloat16`) and, compiled with -DPB200_HALF, IEEE
// float16 (`dtype: float16`, pasero/training.py:289-293) — there `bf16` / `bf162` ARE __half / __half2 and every helper
// rounds to half. Same tensor-core instruction (tcgen05.mma kind::f16), same accumulators, same kernels.
#ifdef PB200_HALF
typedef __half bf16;
typedef __half2 bf162;
static constexpr uint32_t ELEM_ONE_X2 = 0x3c003c00u;     // two 1.0 values
__device__ __forceinline__ float bf16_to_f(bf16 x) { return __half2float(x); }
__device__ __forceinline__ bf16 f_to_bf16(float x) { return __float2half_rn(x); }
__device__ __forceinline__ float round_bf16(float x) { return __half2float(__float2half_rn(x)); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    bf162 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
    bf162 h = *reinterpret_cast<bf162*>(&v);
    return __half22float2(h);
}
// the low / high element of a packed pair as fp32
__device__ __forceinline__ float elem_lo_f(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu))); }
__device__ __forceinline__ float elem_hi_f(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }
#else
typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;
static constexpr uint32_t ELEM_ONE_X2 = 0x3f803f80u;     // two 1.0 values

__device__ __forceinline__ float bf16_to_f(bf16 x) { return __bfloat162float(x); }
__device__ __forceinline__ bf16 f_to_bf16(float x) { return __float2bfloat16_rn(x); }
// round an fp32 value to bf16 precision and come back (emulates the reference's bf16 in-place ops)
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    bf162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// the low / high element of a packed pair as fp32 (a bfloat16 is the top half of an fp32)
__device__ __forceinline__ float elem_lo_f(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float elem_hi_f(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
// by bit operations on the word: __bfloat1622float2 goes through 16-bit sub-registers (a PRMT per pair on top of the shifts)
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) { return make_float2(elem_lo_f(v), elem_hi_f(v)); }
#endif

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG. Dropout masks are a pure function of (seed, offset, element
// index) so the backward kernels regenerate them instead of storing a mask tensor.
// ---------------------------------------------------------------------------------------------
struct Philox {
    uint32_t key0, key1;
    __device__ __forceinline__ Philox(uint64_t seed) : key0((uint32_t)seed), key1((uint32_t)(seed >> 32)) {}
    __device__ __forceinline__ uint4 operator()(uint64_t counter, uint64_t stream) const {
        uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32);
        uint32_t c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
        uint32_t k0 = key0, k1 = key1;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// keep-mask for 8 consecutive elements starting at element index `idx8*8`.
// Each 32-bit Philox word yields two 16-bit uniforms; keep iff u16 >= p*65536.
__device__ __forceinline__ uint32_t dropout_keep8(const Philox& rng, uint64_t offset, uint64_t idx8,
                                                  uint32_t thresh16) {
    uint4 r = rng(idx8, offset);
    uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m |= ((w[i] & 0xffffu) >= thresh16 ? 1u : 0u) << (2 * i);
        m |= ((w[i] >> 16) >= thresh16 ? 1u : 0u) << (2 * i + 1);
    }
    return m;
}
__host__ __device__ __forceinline__ uint32_t dropout_thresh16(float p) {
    float t = p * 65536.0f;
    return t <= 0.f ? 0u : (t >= 65536.f ? 65536u : (uint32_t)(t + 0.5f));
}

}  // namespace pb200
