// Shared helpers for libtmx.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/tmx.h"

namespace tmx {

void set_error(const char* fmt, ...);
int  check_cuda(cudaError_t e, const char* what);
int  require_init();                 // TMX_OK if tmx_init ran for the current device
int  sm_count();

#define TMX_REQUIRE(cond, code, ...)                         \
    do { if (!(cond)) { ::tmx::set_error(__VA_ARGS__); return (code); } } while (0)
#define TMX_CUDA(expr)                                       \
    do { int _rc = ::tmx::check_cuda((expr), #expr); if (_rc) return _rc; } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- 128-bit streaming loads / stores ---------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld_keep(const void* p) {          // plain cached load (re-read soon)
    return *reinterpret_cast<const uint4*>(p);
}
__device__ __forceinline__ void st_stream(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- 16-bit <-> fp32 packing -------------------------------------------------------------------
template <typename T> struct Pack2;
template <> struct Pack2<__half> {
    __device__ static __forceinline__ float2 unpack(uint32_t u) {
        return __half22float2(*reinterpret_cast<const __half2*>(&u));
    }
    __device__ static __forceinline__ uint32_t pack(float a, float b) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __device__ static __forceinline__ float round(float a) { return __half2float(__float2half_rn(a)); }
};
template <> struct Pack2<__nv_bfloat16> {
    __device__ static __forceinline__ float2 unpack(uint32_t u) {
        return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
    }
    __device__ static __forceinline__ uint32_t pack(float a, float b) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __device__ static __forceinline__ float round(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }
};
template <> struct Pack2<float> {
    __device__ static __forceinline__ float round(float a) { return a; }
};

// 8 elements of T from a 16-byte vector (T = half / bf16)
template <typename T>
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    float2 a = Pack2<T>::unpack(v.x), b = Pack2<T>::unpack(v.y), c = Pack2<T>::unpack(v.z), d = Pack2<T>::unpack(v.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
template <typename T>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(Pack2<T>::pack(f[0], f[1]), Pack2<T>::pack(f[2], f[3]),
                      Pack2<T>::pack(f[4], f[5]), Pack2<T>::pack(f[6], f[7]));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace tmx
