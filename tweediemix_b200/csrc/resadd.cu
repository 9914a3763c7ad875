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

// y[r, c] = (a[r, c] + bias[c] (+ b[r, c])) * inv_scale over a [rows, C] (NHWC) tensor: the bias of a
// cuDNN convolution (ATen adds it with a separate, non-vectorised elementwise kernel: 51 launches and
// 1.5 ms per fused step) folded into the ResNet tail add.  b == nullptr: plain per-channel bias add.
template <typename T, bool HAS_B>
__global__ void __launch_bounds__(256)
bias_resadd_kernel(const T* __restrict__ a, const T* __restrict__ b, const float* __restrict__ bias, T* __restrict__ y,
                   size_t nvec, int cvec, float inv_scale) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        float fa[8], fb[8];
        unpack8<T>(ld_stream(a + v * 8), fa);
        if (HAS_B) unpack8<T>(ld_stream(b + v * 8), fb);
        const int c = (int)(v % (size_t)cvec);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * c), b1 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * c + 1);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) fa[j] = (fa[j] + bb[j] + (HAS_B ? fb[j] : 0.f)) * inv_scale;
        st_stream(y + v * 8, pack8<T>(fa));
    }
}

// h = a + b (rounded to T, stored), n = LayerNorm(h) * gamma + beta (stored): the residual add of a
// BasicTransformerBlock fused with the LayerNorm that consumes its result.  One warp per row, the row
// stays in registers: 2 reads + 2 writes instead of (2R + 1W) + (1R + 1W) and one launch instead of two.
// The statistics are taken from the ROUNDED h, i.e. exactly the tensor the un-fused LayerNorm would read.
#ifndef TMX_RLN_WARPS
#define TMX_RLN_WARPS 2
#endif
constexpr int kRlnWarps = TMX_RLN_WARPS;       // rows (= warps) per CTA.  Measured (profiles/r02z_kbench_resadd_ln_cta_size.txt): 2-warp CTAs (16 resident
                                               // per SM) balance the 4096-row sites over 148 SMs better than 8-warp ones: 10.1 -> 9.3 us at D = 1280
template <typename T, int NV>
__global__ void __launch_bounds__(32 * kRlnWarps, 32 / kRlnWarps)      // 32 resident warps / SM: the CTAs of a 4 x 1024-token site run as ONE wave (60 registers, no spills)
resadd_layernorm_kernel(const T* __restrict__ a, const T* __restrict__ b, const float* __restrict__ gamma,
                        const float* __restrict__ beta, T* __restrict__ h_out, T* __restrict__ n_out,
                        long long rows, int D, float eps) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kRlnWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int nvec = D >> 3;
    const T* ar = a + row * (long long)D;
    const T* br = b + row * (long long)D;
    uint4 va[NV], vb[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) { va[i] = ld_stream(ar + (size_t)c * 8); vb[i] = ld_stream(br + (size_t)c * 8); }
    }
    T* hr = h_out + row * (long long)D;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) {
            float fa[8], fb[8];
            unpack8<T>(va[i], fa);
            unpack8<T>(vb[i], fb);
#pragma unroll
            for (int j = 0; j < 8; ++j) fa[j] += fb[j];
            va[i] = pack8<T>(fa);                       // rounded h
            st_stream(hr + (size_t)c * 8, va[i]);
            unpack8<T>(va[i], fa);
            s += ((fa[0] + fa[1]) + (fa[2] + fa[3])) + ((fa[4] + fa[5]) + (fa[6] + fa[7]));
        }
    }
    const float mean = warp_sum(s) / (float)D;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (lane + i * 32 < nvec) {
            float f[8];
            unpack8<T>(va[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = f[j] - mean; ss = fmaf(d, d, ss); }
        }
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)D + eps);
    T* nr = n_out + row * (long long)D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = lane + i * 32;
        if (c < nvec) {
            float f[8];
            unpack8<T>(va[i], f);
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c + 1);
            f[0] = fmaf((f[0] - mean) * rstd, g0.x, b0.x); f[1] = fmaf((f[1] - mean) * rstd, g0.y, b0.y);
            f[2] = fmaf((f[2] - mean) * rstd, g0.z, b0.z); f[3] = fmaf((f[3] - mean) * rstd, g0.w, b0.w);
            f[4] = fmaf((f[4] - mean) * rstd, g1.x, b1.x); f[5] = fmaf((f[5] - mean) * rstd, g1.y, b1.y);
            f[6] = fmaf((f[6] - mean) * rstd, g1.z, b1.z); f[7] = fmaf((f[7] - mean) * rstd, g1.w, b1.w);
            st_stream(nr + (size_t)c * 8, pack8<T>(f));
        }
    }
}

template <typename T>
static int launch_resadd_ln(const void* a, const void* b, const float* gamma, const float* beta, void* h, void* n,
                            long long rows, int D, float eps, cudaStream_t st) {
    const unsigned blocks = (unsigned)((rows + kRlnWarps - 1) / kRlnWarps);
    const int nv = (D / 8 + 31) / 32;
#define TMX_RLN(NVV) resadd_layernorm_kernel<T, NVV><<<blocks, 32 * kRlnWarps, 0, st>>>((const T*)a, (const T*)b, gamma, beta, (T*)h, (T*)n, rows, D, eps)
    switch (nv) {
        case 1: TMX_RLN(1); break;
        case 2: TMX_RLN(2); break;
        case 3: TMX_RLN(3); break;
        case 4: TMX_RLN(4); break;
        case 5: TMX_RLN(5); break;
        case 6: TMX_RLN(6); break;
        case 7: TMX_RLN(7); break;
        default: TMX_RLN(8); break;
    }
#undef TMX_RLN
    return check_cuda(cudaGetLastError(), "resadd_layernorm_kernel launch");
}

}  // namespace tmx

using namespace tmx;

extern "C" int tmx_bias_resadd_fwd(const void* a, const void* b, const float* bias, void* y, size_t rows, int C,
                                   float inv_scale, int dtype, void* stream) {
    TMX_REQUIRE(a && bias && y, TMX_EINVAL, "bias_resadd: null pointer");
    TMX_REQUIRE(rows > 0 && C > 0 && C % 8 == 0, TMX_ESHAPE, "bias_resadd: C=%d must be a positive multiple of 8", C);
    TMX_REQUIRE(aligned16(a) && aligned16(b) && aligned16(y) && aligned16(bias), TMX_EALIGN, "bias_resadd: 16-byte alignment");
    if (int rc = require_init()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nvec = rows * (size_t)(C / 8);
    size_t blocks = (nvec + 255) / 256;
    const size_t cap = (size_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    const int cvec = C / 8;
#define TMX_BRA(TT) do { if (b) bias_resadd_kernel<TT, true><<<(unsigned)blocks, 256, 0, st>>>((const TT*)a, (const TT*)b, bias, (TT*)y, nvec, cvec, inv_scale); \
                         else   bias_resadd_kernel<TT, false><<<(unsigned)blocks, 256, 0, st>>>((const TT*)a, nullptr, bias, (TT*)y, nvec, cvec, inv_scale); } while (0)
    switch (dtype) {
        case TMX_F16:  TMX_BRA(__half); break;
        case TMX_BF16: TMX_BRA(__nv_bfloat16); break;
        default: set_error("bias_resadd: unsupported dtype %d (fp16/bf16 only)", dtype); return TMX_EDTYPE;
    }
#undef TMX_BRA
    return check_cuda(cudaGetLastError(), "bias_resadd_kernel launch");
}

extern "C" int tmx_resadd_layernorm_fwd(const void* a, const void* b, const float* gamma, const float* beta,
                                        void* h_out, void* n_out, size_t rows, int D, float eps, int dtype, void* stream) {
    TMX_REQUIRE(a && b && gamma && beta && h_out && n_out, TMX_EINVAL, "resadd_layernorm: null pointer");
    TMX_REQUIRE(rows > 0 && D > 0, TMX_EINVAL, "resadd_layernorm: non-positive size");
    TMX_REQUIRE(D % 8 == 0 && D <= 2048, TMX_ESHAPE, "resadd_layernorm: D=%d must be a multiple of 8 and <= 2048", D);
    TMX_REQUIRE(rows <= 0x7fffffffULL * (unsigned long long)kRlnWarps, TMX_ESHAPE, "resadd_layernorm: too many rows");
    TMX_REQUIRE(aligned16(a) && aligned16(b) && aligned16(h_out) && aligned16(n_out) && aligned16(gamma) && aligned16(beta),
                TMX_EALIGN, "resadd_layernorm: 16-byte alignment");
    TMX_REQUIRE(n_out != a && n_out != b && n_out != h_out, TMX_EINVAL, "resadd_layernorm: n_out must not alias the other buffers");
    if (int rc = require_init()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case TMX_F16:  return launch_resadd_ln<__half>(a, b, gamma, beta, h_out, n_out, (long long)rows, D, eps, st);
        case TMX_BF16: return launch_resadd_ln<__nv_bfloat16>(a, b, gamma, beta, h_out, n_out, (long long)rows, D, eps, st);
    }
    set_error("resadd_layernorm: unsupported dtype %d (fp16/bf16 only)", dtype);
    return TMX_EDTYPE;
}

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
