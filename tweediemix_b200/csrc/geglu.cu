// k9a — GEGLU gating  y[m, j] = x[m, j] * gelu(x[m, F + j])  (exact erf GELU, as F.gelu's default).
// HBM-bound: reads the [M, 2F] projection once, writes [M, F] once, 128-bit accesses, fp32 math,
// one rounding on store.  Replaces [D] diffusers GEGLU.forward's chunk + gelu + mul (three eager
// kernels, five passes over the 4d-wide tensor) inside BasicTransformerBlock.ff — 70 sites per
// U-Net forward, the largest elementwise tensor of the step.
#include "tmx_common.cuh"

namespace tmx {

// exact-GELU  0.5 v (1 + erf(v / sqrt 2))  with erf from Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7, far
// below the 2^-9 / 2^-12 output rounding): one MUFU.RCP + one MUFU.EX2 + 7 FMAs instead of libdevice erff's
// ~25-instruction branchy path, which made this kernel issue-bound instead of HBM-bound.
__device__ __forceinline__ float gelu_erf(float v) {
    const float z = fabsf(v) * 0.70710678118654752f;
    const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
    float p = fmaf(t, 1.061405429f, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float e = p * t * exp2f(-1.4426950408889634f * z * z);     // 1 - erf(z), z >= 0
    const float half_erfc = 0.5f * e;                                 // Phi(-|v|)
    return v >= 0.f ? v * (1.f - half_erfc) : v * half_erfc;
}

template <typename T>
__global__ void __launch_bounds__(256)
geglu_kernel(const T* __restrict__ x, T* __restrict__ y, size_t rows, int FV /* F/8 */) {
    const size_t total = rows * (size_t)FV;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t m = i / FV;
        const int j = (int)(i - m * FV);
        const T* row = x + m * (size_t)FV * 16;
        float h[8], g[8];
        unpack8<T>(ld_stream(row + (size_t)j * 8), h);
        unpack8<T>(ld_stream(row + (size_t)(FV + j) * 8), g);
#pragma unroll
        for (int k = 0; k < 8; ++k) h[k] *= gelu_erf(g[k]);
        st_stream(y + i * 8, pack8<T>(h));
    }
}

}  // namespace tmx

using namespace tmx;

extern "C" int tmx_geglu_fwd(const void* x, void* y, size_t rows, int F, int dtype, void* stream) {
    TMX_REQUIRE(x && y, TMX_EINVAL, "geglu: null pointer");
    TMX_REQUIRE(rows > 0 && F > 0 && F % 8 == 0, TMX_ESHAPE, "geglu: rows=%zu F=%d (F must be a positive multiple of 8)", rows, F);
    TMX_REQUIRE(aligned16(x) && aligned16(y), TMX_EALIGN, "geglu: 16-byte alignment");
    if (int rc = require_init()) return rc;
    const int FV = F / 8;
    const size_t total = rows * (size_t)FV;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case TMX_F16:  geglu_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>((const __half*)x, (__half*)y, rows, FV); break;
        case TMX_BF16: geglu_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, rows, FV); break;
        default: set_error("geglu: unsupported dtype %d (fp16/bf16 only)", dtype); return TMX_EDTYPE;
    }
    return check_cuda(cudaGetLastError(), "geglu_kernel launch");
}
