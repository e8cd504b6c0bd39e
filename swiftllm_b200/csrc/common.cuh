// Shared helpers for the sm_100a data-plane kernels.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/swiftllm_b200.h"

namespace sllm {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError -> error code + message

#define SLLM_REQUIRE(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            ::sllm::set_error(__VA_ARGS__); \
            return 1;                      \
        }                                  \
    } while (0)

// ---------------------------------------------------------------- dtype traits
template <typename T> struct Traits;
template <> struct Traits<__half> {
    using T2 = __half2;
    static __device__ __forceinline__ float2 to_f2(T2 v) { return __half22float2(v); }
    static __device__ __forceinline__ T2 from_f2(float2 v) { return __float22half2_rn(v); }
    static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct Traits<__nv_bfloat16> {
    using T2 = __nv_bfloat162;
    static __device__ __forceinline__ float2 to_f2(T2 v) { return __bfloat1622float2(v); }
    static __device__ __forceinline__ T2 from_f2(float2 v) { return __float22bfloat162_rn(v); }
    static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// 16-byte vector of 8 half/bf16 values
template <typename T> struct alignas(16) Vec8 {
    typename Traits<T>::T2 v[4];
};

template <typename T> __device__ __forceinline__ Vec8<T> ld_vec8(const T* p) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    return *reinterpret_cast<Vec8<T>*>(&u);
}
template <typename T> __device__ __forceinline__ void st_vec8(T* p, const Vec8<T>& v) {
    *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v);
}
// streaming (read-once) 16-byte load that does not allocate in L1
template <typename T> __device__ __forceinline__ Vec8<T> ld_vec8_stream(const T* p) {
    uint4 u;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
    return *reinterpret_cast<Vec8<T>*>(&u);
}

// shared-memory address of a generic pointer (operand of cp.async / ldmatrix / mbarrier / TMA / UMMA descriptors)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- per-device one-time state (a process may drive several GPUs)
static inline int current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return d;
}
// true exactly once per (call site's mask, device): e.g. cudaFuncSetAttribute, which is a per-device setting
static inline bool first_use_on_this_device(unsigned long long& mask) {
    const unsigned long long bit = 1ull << (current_device() & 63);
    const unsigned long long old = __atomic_fetch_or(&mask, bit, __ATOMIC_RELAXED);
    return (old & bit) == 0;
}

// dispatch on the dtype tag
#define SLLM_DISPATCH_DTYPE(dtype, ...)                                     \
    do {                                                                    \
        if ((dtype) == SLLM_F16) { using T = __half; __VA_ARGS__; }         \
        else if ((dtype) == SLLM_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
        else { ::sllm::set_error("unknown dtype tag %d", (int)(dtype)); return 1; } \
    } while (0)

}  // namespace sllm
