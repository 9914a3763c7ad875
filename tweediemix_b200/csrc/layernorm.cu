// k8 — LayerNorm over the last dimension of a [rows, D] 16-bit tensor (the three norms of every
// BasicTransformerBlock: 210 sites per U-Net forward, 3.1e8 elements per sample-forward — more
// traffic than GroupNorm and the residual adds together).
//
// HBM-bound, one pass: one warp owns one row, holds it in registers as packed 128-bit vectors
// (D <= 2048 -> at most 8 vectors per lane), computes mean and the CENTRED variance in fp32 with warp
// shuffles (two-pass in registers: no E[x^2]-E[x]^2 cancellation), normalises and stores once.
// 8 rows per 256-thread CTA, <= 64 registers so 8 CTAs (64 warps, >= 80 KB of loads in flight) fit an
// SM.  Replaces [D] F.layer_norm in BasicTransformerBlock.norm1/2/3 (fp32 statistics under autocast,
// SURVEY App. B), which ATen runs at ~1/6 of the HBM roofline for these shapes.
#include "tmx_common.cuh"

namespace tmx {

template <typename T, int NV>
__global__ void __launch_bounds__(256, 4)
layernorm_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 T* __restrict__ y, long long rows, int D, float eps) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int nvec = D >> 3;                       // 8 elements per 128-bit vector
    const T* xr = x + row * (long long)D;
    uint4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) v[i] = ld_stream(xr + (size_t)c * 8);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (lane + i * 32 < nvec) {
            float f[8];
            unpack8<T>(v[i], f);
            s += ((f[0] + f[1]) + (f[2] + f[3])) + ((f[4] + f[5]) + (f[6] + f[7]));
        }
    }
    const float mean = warp_sum(s) / (float)D;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (lane + i * 32 < nvec) {
            float f[8];
            unpack8<T>(v[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = f[j] - mean; ss = fmaf(d, d, ss); }
        }
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)D + eps);
    T* yr = y + row * (long long)D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) {
            float f[8];
            unpack8<T>(v[i], f);
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c + 1);
            f[0] = fmaf((f[0] - mean) * rstd, g0.x, b0.x); f[1] = fmaf((f[1] - mean) * rstd, g0.y, b0.y);
            f[2] = fmaf((f[2] - mean) * rstd, g0.z, b0.z); f[3] = fmaf((f[3] - mean) * rstd, g0.w, b0.w);
            f[4] = fmaf((f[4] - mean) * rstd, g1.x, b1.x); f[5] = fmaf((f[5] - mean) * rstd, g1.y, b1.y);
            f[6] = fmaf((f[6] - mean) * rstd, g1.z, b1.z); f[7] = fmaf((f[7] - mean) * rstd, g1.w, b1.w);
            st_stream(yr + (size_t)c * 8, pack8<T>(f));
        }
    }
}

template <typename T>
static int launch_ln(const void* x, const float* gamma, const float* beta, void* y, long long rows, int D, float eps, cudaStream_t st) {
    const unsigned blocks = (unsigned)((rows + 7) / 8);
    const int nv = (D / 8 + 31) / 32;
    switch (nv) {
        case 1: layernorm_kernel<T, 1><<<blocks, 256, 0, st>>>((const T*)x, gamma, beta, (T*)y, rows, D, eps); break;
        case 2: layernorm_kernel<T, 2><<<blocks, 256, 0, st>>>((const T*)x, gamma, beta, (T*)y, rows, D, eps); break;
        case 3: layernorm_kernel<T, 3><<<blocks, 256, 0, st>>>((const T*)x, gamma, beta, (T*)y, rows, D, eps); break;
        case 4: layernorm_kernel<T, 4><<<blocks, 256, 0, st>>>((const T*)x, gamma, beta, (T*)y, rows, D, eps); break;
        case 5: layernorm_kernel<T, 5><<<blocks, 256, 0, st>>>((const T*)x, gamma, beta, (T*)y, rows, D, eps); break;
        case 6: layernorm_kernel<T, 6><<<blocks, 256, 0, st>>>((const T*)x, gamma, beta, (T*)y, rows, D, eps); break;
        case 7: layernorm_kernel<T, 7><<<blocks, 256, 0, st>>>((const T*)x, gamma, beta, (T*)y, rows, D, eps); break;
        default: layernorm_kernel<T, 8><<<blocks, 256, 0, st>>>((const T*)x, gamma, beta, (T*)y, rows, D, eps); break;
    }
    return check_cuda(cudaGetLastError(), "layernorm_kernel launch");
}

}  // namespace tmx

using namespace tmx;

extern "C" int tmx_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y,
                                 size_t rows, int D, float eps, int dtype, void* stream) {
    TMX_REQUIRE(x && gamma && beta && y, TMX_EINVAL, "layernorm: null pointer");
    TMX_REQUIRE(rows > 0 && D > 0, TMX_EINVAL, "layernorm: non-positive size");
    TMX_REQUIRE(D % 8 == 0 && D <= 2048, TMX_ESHAPE, "layernorm: D=%d must be a multiple of 8 and <= 2048", D);
    TMX_REQUIRE(rows <= 0x7fffffffULL * 8ULL, TMX_ESHAPE, "layernorm: too many rows");
    TMX_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), TMX_EALIGN, "layernorm: 16-byte alignment");
    if (int rc = require_init()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case TMX_F16:  return launch_ln<__half>(x, gamma, beta, y, (long long)rows, D, eps, st);
        case TMX_BF16: return launch_ln<__nv_bfloat16>(x, gamma, beta, y, (long long)rows, D, eps, st);
    }
    set_error("layernorm: unsupported dtype %d (fp16/bf16 only)", dtype);
    return TMX_EDTYPE;
}
