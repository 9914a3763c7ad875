// k6 — residual add  y = (a + b) * inv_scale.  HBM-bound: 2 reads + 1 write, 128-bit accesses,
// two vectors in flight per thread.  Replaces [D] ResnetBlock2D's `(input + hidden) /
// output_scale_factor` (mirrored at video_gen/utils_attn.py:428-431) and the residual adds of
// BasicTransformerBlock.
#include "tmx_common.cuh"

namespace tmx {

template <typename T>
__global__ void __launch_bounds__(256)
resadd_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, size_t nvec, float inv_scale) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        float fa[8], fb[8];
        unpack8<T>(ld_stream(a + v * 8), fa);
        unpack8<T>(ld_stream(b + v * 8), fb);
#pragma unroll
        for (int j = 0; j < 8; ++j) fa[j] = (fa[j] + fb[j]) * inv_scale;
        st_stream(y + v * 8, pack8<T>(fa));
    }
}

__global__ void __launch_bounds__(256)
resadd_kernel_f32(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, size_t nvec, float inv_scale) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const uint4 ua = ld_stream(a + v * 4), ub = ld_stream(b + v * 4);
        uint4 o;
        o.x = __float_as_uint((__uint_as_float(ua.x) + __uint_as_float(ub.x)) * inv_scale);
        o.y = __float_as_uint((__uint_as_float(ua.y) + __uint_as_float(ub.y)) * inv_scale);
        o.z = __float_as_uint((__uint_as_float(ua.z) + __uint_as_float(ub.z)) * inv_scale);
        o.w = __float_as_uint((__uint_as_float(ua.w) + __uint_as_float(ub.w)) * inv_scale);
        st_stream(y + v * 4, o);
    }
}

}  // namespace tmx

using namespace tmx;

extern "C" int tmx_resadd_fwd(const void* a, const void* b, void* y, size_t n, float inv_scale,
                              int dtype, void* stream) {
    TMX_REQUIRE(a && b && y, TMX_EINVAL, "resadd: null pointer");
    TMX_REQUIRE(n > 0 && n % 8 == 0, TMX_ESHAPE, "resadd: n=%zu must be a positive multiple of 8", n);
    TMX_REQUIRE(aligned16(a) && aligned16(b) && aligned16(y), TMX_EALIGN, "resadd: 16-byte alignment");
    if (int rc = require_init()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nvec = dtype == TMX_F32 ? n / 4 : n / 8;
    size_t blocks = (nvec + 255) / 256;
    const size_t cap = (size_t)sm_count() * 16;              // grid-stride beyond 16 CTAs per SM
    if (blocks > cap) blocks = cap;
    switch (dtype) {
        case TMX_F16:  resadd_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>((const __half*)a, (const __half*)b, (__half*)y, nvec, inv_scale); break;
        case TMX_BF16: resadd_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)y, nvec, inv_scale); break;
        case TMX_F32:  resadd_kernel_f32<<<(unsigned)blocks, 256, 0, st>>>((const float*)a, (const float*)b, (float*)y, nvec, inv_scale); break;
        default: set_error("resadd: unsupported dtype %d", dtype); return TMX_EDTYPE;
    }
    return check_cuda(cudaGetLastError(), "resadd_kernel launch");
}
