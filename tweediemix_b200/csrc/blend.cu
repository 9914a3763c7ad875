// k7 — fused CFG combine + Tweedie x0 + mask-weighted concept blend + DDIM update, and its
// concept-parallel partial/finish split.  HBM-bound elementwise work: 128-bit coalesced loads,
// every operand read once from HBM (masks are re-read per channel from L1), one pass.
//
// Reference arithmetic: fusion_generation/fusion_sampling.py:376-386 (blend), :430 (DDIM),
// :471-472 (last step), :392-403 / :407-412 / :421-428 (weights / null-mask forms).
#include "tmx_common.cuh"
#include <cmath>

namespace tmx {

constexpr int kMaxConcepts = 16;

struct BlendCoef {
    float s_t;        // sqrt(1 - a_t)
    float sqrt_at;    // sqrt(a_t)
    float inv_sqrt_at;// 1 / sqrt(a_t), fp32 (what ATen's CUDA div-by-CPU-scalar multiplies with)
    float sqrt_an;    // sqrt(a_next)
    float s_n;        // sqrt(1 - a_next)
    float g;
    int   is_last;
    int   has_w;
    float w[kMaxConcepts];
};

template <typename T> struct Vec8 {            // 8 x 16-bit in one 128-bit load
    static __device__ __forceinline__ void load(const T* p, float (&f)[8]) { unpack8<T>(ld_stream(p), f); }
};
template <> struct Vec8<float> {
    static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
        uint4 a = ld_stream(p), b = ld_stream(p + 4);
        f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
        f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
    }
};

// 4 consecutive elements: 64-bit load for 16-bit types, 128-bit for fp32.  With 4 pixels per thread every
// load instruction of a warp covers whole 32-byte sectors (a 128-bit fp32 load per thread = 512 contiguous
// bytes per warp), which the 8-pixel layout did not: its two half-sector fp32 loads fetched every sector
// of x / masks twice from L2.
template <typename T> struct Vec4 {
    static __device__ __forceinline__ void load(const T* p, float (&f)[4]) {
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
        float2 a = Pack2<T>::unpack(r.x), b = Pack2<T>::unpack(r.y);
        f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
    }
};
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float (&f)[4]) {
        uint4 a = ld_stream(p);
        f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
    }
};
__device__ __forceinline__ void load4_f32_cached(const float* p, float (&f)[4]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
}
__device__ __forceinline__ void store4_f32(float* p, const float (&f)[4]) {
    st_stream(p, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
}

__device__ __forceinline__ void load8_f32_cached(const float* p, float (&f)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8_f32(float* p, const float (&f)[8]) {
    st_stream(p, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
    st_stream(p + 4, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
}

// One thread: one image, one channel, 4 consecutive pixels.  KT > 0 is the compile-time concept count
// (all 2 + 2*K loads are issued before the first use: at one image the kernel is pure latency, at
// thousands it must keep ~35 KB per SM in flight); KT == 0 is the generic run-time loop.
// The Tweedie divide is a multiplication by 1/sqrt(a_t) rounded to fp32 once on the host — exactly what
// ATen's CUDA `tensor / cpu_scalar` does (it multiplies by the reciprocal), so this is not an approximation
// of the reference but its arithmetic.
template <typename T, bool REF, int KT>
__global__ void __launch_bounds__(256)
blend_kernel(const float* __restrict__ x, const T* __restrict__ eps, const float* __restrict__ masks,
             float* __restrict__ x_out, float* __restrict__ x0_out,
             int imgs, int K, int C, int HW, const __grid_constant__ BlendCoef cf) {
    constexpr int KMAX = KT > 0 ? KT : 1;
    const int quads = HW >> 2;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)imgs * C * quads) return;
    const int p = (int)(idx % quads) << 2;
    const int ch = (int)((idx / quads) % C);
    const int img = (int)(idx / ((long long)quads * C));
    const int Kc = KT > 0 ? KT : K;

    const size_t chw = (size_t)C * HW;
    const size_t off = (size_t)ch * HW + p;
    const float* xi = x + (size_t)img * chw + off;
    const T* ei = eps + (size_t)img * (Kc + 1) * chw + off;

    float xv[4], eu[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
    Vec4<float>::load(xi, xv);
    Vec4<T>::load(ei, eu);

    auto fold = [&](const float (&ec)[4], const float (&m)[4], int c) {
        const float w = cf.has_w ? cf.w[c] : 1.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float t;
            if (REF) {
                // eps dtype arithmetic exactly as torch evaluates :383 under autocast (App. B)
                float d = Pack2<T>::round(__fsub_rn(ec[i], eu[i]));
                d = Pack2<T>::round(__fmul_rn(cf.g, d));
                float e = Pack2<T>::round(__fadd_rn(eu[i], d));
                float se = Pack2<T>::round(__fmul_rn(cf.s_t, e));
                t = __fmul_rn(__fsub_rn(xv[i], se), cf.inv_sqrt_at);
                if (masks) t = __fmul_rn(m[i], t);
                if (cf.has_w) t = __fmul_rn(w, t);
                acc[i] = __fadd_rn(acc[i], t);
            } else {
                const float e = fmaf(cf.g, ec[i] - eu[i], eu[i]);
                t = fmaf(-cf.s_t, e, xv[i]) * cf.inv_sqrt_at;
                acc[i] = fmaf(masks ? m[i] * w : w, t, acc[i]);
            }
        }
    };

    if constexpr (KT > 0) {
        float ec[KMAX][4], m[KMAX][4];
#pragma unroll
        for (int c = 0; c < KT; ++c) {
            Vec4<T>::load(ei + (size_t)(1 + c) * chw, ec[c]);
            if (masks) load4_f32_cached(masks + (size_t)c * HW + p, m[c]);
            else { m[c][0] = m[c][1] = m[c][2] = m[c][3] = 1.f; }
        }
#pragma unroll
        for (int c = 0; c < KT; ++c) fold(ec[c], m[c], c);
    } else {
#pragma unroll 2
        for (int c = 0; c < K; ++c) {
            float ec[4], m[4] = {1.f, 1.f, 1.f, 1.f};
            Vec4<T>::load(ei + (size_t)(1 + c) * chw, ec);
            if (masks) load4_f32_cached(masks + (size_t)c * HW + p, m);
            fold(ec, m, c);
        }
    }
    if (x0_out) store4_f32(x0_out + (size_t)img * chw + off, acc);
    float out[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (cf.is_last) {
            out[i] = acc[i];
        } else if (REF) {
            float nz = Pack2<T>::round(__fmul_rn(cf.s_n, eu[i]));
            out[i] = __fadd_rn(__fmul_rn(cf.sqrt_an, acc[i]), nz);
        } else {
            out[i] = fmaf(cf.sqrt_an, acc[i], cf.s_n * eu[i]);
        }
    }
    store4_f32(x_out + (size_t)img * chw + off, out);
}

struct PartialRows {
    int n;
    int ids[kMaxConcepts + 1];
    float w[kMaxConcepts + 1];        // weight of the concept each local row carries (unused for row id 0)
};

// One thread: one image, one channel, 8 pixels.  acc[img][0] = sum m_c eps_c ; acc[img][1] = eps_u or 0.
template <typename T>
__global__ void __launch_bounds__(256)
blend_partial_kernel(const T* __restrict__ eps_rows, const float* __restrict__ masks, float* __restrict__ acc,
                     int imgs, int C, int HW, const __grid_constant__ PartialRows rows) {
    const int groups = HW >> 3;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)imgs * C * groups) return;
    const int p = (int)(idx % groups) << 3;
    const int ch = (int)((idx / groups) % C);
    const int img = (int)(idx / ((long long)groups * C));
    const size_t chw = (size_t)C * HW;
    float a[8], u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = 0.f; u[i] = 0.f; }
    for (int r = 0; r < rows.n; ++r) {
        float e[8];
        Vec8<T>::load(eps_rows + ((size_t)img * rows.n + r) * chw + (size_t)ch * HW + p, e);
        const int id = rows.ids[r];
        if (id == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = e[i];
        } else {
            const float w = rows.w[r];
            if (masks) {
                float m[8];
                load8_f32_cached(masks + (size_t)(id - 1) * HW + p, m);
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fmaf(w * m[i], e[i], a[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fmaf(w, e[i], a[i]);
            }
        }
    }
    float* o = acc + (size_t)img * 2 * chw + (size_t)ch * HW + p;
    store8_f32(o, a);
    store8_f32(o + chw, u);
}

__global__ void __launch_bounds__(256)
blend_finish_kernel(const float* __restrict__ x, const float* __restrict__ acc, const float* __restrict__ masks,
                    float* __restrict__ x_out, float* __restrict__ x0_out,
                    int imgs, int K, int C, int HW, const __grid_constant__ BlendCoef cf) {
    const int groups = HW >> 3;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)imgs * groups) return;
    const int img = (int)(idx / groups);
    const int p = (int)(idx - (long long)img * groups) << 3;
    const size_t chw = (size_t)C * HW;
    float M[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) M[i] = 0.f;
    for (int c = 0; c < K; ++c) {
        const float w = cf.has_w ? cf.w[c] : 1.f;
        if (masks) {
            float m[8];
            load8_f32_cached(masks + (size_t)c * HW + p, m);
#pragma unroll
            for (int i = 0; i < 8; ++i) M[i] = fmaf(w, m[i], M[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) M[i] += w;
        }
    }
    const float k_u = cf.s_t * (1.f - cf.g), k_a = cf.s_t * cf.g;
    for (int ch = 0; ch < C; ++ch) {
        float xv[8], a[8], u[8], x0[8], out[8];
        Vec8<float>::load(x + (size_t)img * chw + (size_t)ch * HW + p, xv);
        Vec8<float>::load(acc + (size_t)img * 2 * chw + (size_t)ch * HW + p, a);
        Vec8<float>::load(acc + (size_t)img * 2 * chw + chw + (size_t)ch * HW + p, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x0[i] = (M[i] * xv[i] - k_u * M[i] * u[i] - k_a * a[i]) * cf.inv_sqrt_at;
            out[i] = cf.is_last ? x0[i] : fmaf(cf.sqrt_an, x0[i], cf.s_n * u[i]);
        }
        if (x0_out) store8_f32(x0_out + (size_t)img * chw + (size_t)ch * HW + p, x0);
        store8_f32(x_out + (size_t)img * chw + (size_t)ch * HW + p, out);
    }
}

static BlendCoef make_coef(float a_t, float a_next, float g, int is_last, const float* w, int K) {
    BlendCoef cf{};
    cf.s_t = sqrtf(1.0f - a_t);        // fp32 like the reference's 0-dim fp32 tensors (App. B)
    cf.sqrt_at = sqrtf(a_t);
    cf.inv_sqrt_at = 1.0f / cf.sqrt_at;
    cf.sqrt_an = sqrtf(a_next);
    cf.s_n = sqrtf(1.0f - a_next);
    cf.g = g;
    cf.is_last = is_last;
    cf.has_w = w != nullptr;
    for (int i = 0; i < K && w; ++i) cf.w[i] = w[i];
    return cf;
}

template <typename T, bool REF>
static void launch_blend_k(const float* x, const T* eps, const float* masks, float* x_out, float* x0_out,
                           int imgs, int K, int C, int HW, const BlendCoef& cf, unsigned blocks, cudaStream_t st) {
    switch (K) {
        case 1: blend_kernel<T, REF, 1><<<blocks, 256, 0, st>>>(x, eps, masks, x_out, x0_out, imgs, K, C, HW, cf); break;
        case 2: blend_kernel<T, REF, 2><<<blocks, 256, 0, st>>>(x, eps, masks, x_out, x0_out, imgs, K, C, HW, cf); break;
        case 3: blend_kernel<T, REF, 3><<<blocks, 256, 0, st>>>(x, eps, masks, x_out, x0_out, imgs, K, C, HW, cf); break;
        case 4: blend_kernel<T, REF, 4><<<blocks, 256, 0, st>>>(x, eps, masks, x_out, x0_out, imgs, K, C, HW, cf); break;
        default: blend_kernel<T, REF, 0><<<blocks, 256, 0, st>>>(x, eps, masks, x_out, x0_out, imgs, K, C, HW, cf); break;
    }
}

template <typename T>
static int launch_blend(const float* x, const void* eps, const float* masks, float* x_out, float* x0_out,
                        int imgs, int K, int C, int HW, const BlendCoef& cf, int round_mode, cudaStream_t st) {
    const long long total = (long long)imgs * C * (HW >> 2);
    const long long nblk = (total + 255) / 256;
    if (nblk > 0x7fffffffLL) { set_error("tweedie_blend: problem too large for one launch"); return TMX_ESHAPE; }
    const unsigned blocks = (unsigned)nblk;
    if (round_mode == TMX_ROUND_REF)
        launch_blend_k<T, true>(x, (const T*)eps, masks, x_out, x0_out, imgs, K, C, HW, cf, blocks, st);
    else
        launch_blend_k<T, false>(x, (const T*)eps, masks, x_out, x0_out, imgs, K, C, HW, cf, blocks, st);
    return check_cuda(cudaGetLastError(), "blend_kernel launch");
}

}  // namespace tmx

using namespace tmx;

extern "C" int tmx_tweedie_blend_ddim_fwd(const float* x, const void* eps, const float* masks,
                                          const float* weights, float* x_out, float* x0_out,
                                          int imgs, int K, int C, int HW,
                                          float a_t, float a_next, float g, int is_last,
                                          int eps_dtype, int round_mode, void* stream) {
    TMX_REQUIRE(x && eps && x_out, TMX_EINVAL, "tweedie_blend: null pointer");
    TMX_REQUIRE(imgs > 0 && C > 0 && HW > 0, TMX_EINVAL, "tweedie_blend: non-positive size");
    TMX_REQUIRE(K >= 1 && K <= kMaxConcepts, TMX_ESHAPE, "tweedie_blend: K=%d outside [1,%d]", K, kMaxConcepts);
    TMX_REQUIRE(HW % 8 == 0, TMX_ESHAPE, "tweedie_blend: HW=%d must be a multiple of 8", HW);
    TMX_REQUIRE(aligned16(x) && aligned16(eps) && aligned16(x_out) && aligned16(masks) && aligned16(x0_out),
                TMX_EALIGN, "tweedie_blend: pointers must be 16-byte aligned");
    TMX_REQUIRE(a_t > 0.f && a_t <= 1.f && a_next > 0.f && a_next <= 1.f, TMX_EINVAL,
                "tweedie_blend: alphas must be in (0,1], got %g %g", a_t, a_next);
    TMX_REQUIRE(round_mode == TMX_ROUND_FP32 || round_mode == TMX_ROUND_REF, TMX_EINVAL, "tweedie_blend: bad round_mode");
    if (int rc = require_init()) return rc;
    BlendCoef cf = make_coef(a_t, a_next, g, is_last, weights, K);
    cudaStream_t st = (cudaStream_t)stream;
    switch (eps_dtype) {
        case TMX_F16:  return launch_blend<__half>(x, eps, masks, x_out, x0_out, imgs, K, C, HW, cf, round_mode, st);
        case TMX_BF16: return launch_blend<__nv_bfloat16>(x, eps, masks, x_out, x0_out, imgs, K, C, HW, cf, round_mode, st);
        case TMX_F32:  return launch_blend<float>(x, eps, masks, x_out, x0_out, imgs, K, C, HW, cf, round_mode, st);
    }
    set_error("tweedie_blend: unsupported eps dtype %d", eps_dtype);
    return TMX_EDTYPE;
}

extern "C" int tmx_blend_partial_fwd(const void* eps_rows, const float* masks, const float* weights,
                                     const int* row_ids, float* acc, int imgs, int R, int K, int C, int HW,
                                     int eps_dtype, void* stream) {
    TMX_REQUIRE((eps_rows || R == 0) && (row_ids || R == 0) && acc, TMX_EINVAL, "blend_partial: null pointer");
    TMX_REQUIRE(imgs > 0 && C > 0 && HW > 0 && R >= 0, TMX_EINVAL, "blend_partial: bad size");
    TMX_REQUIRE(K >= 1 && K <= kMaxConcepts && R <= K + 1, TMX_ESHAPE, "blend_partial: K=%d R=%d unsupported", K, R);
    TMX_REQUIRE(HW % 8 == 0, TMX_ESHAPE, "blend_partial: HW=%d must be a multiple of 8", HW);
    TMX_REQUIRE(aligned16(eps_rows) && aligned16(masks) && aligned16(acc), TMX_EALIGN, "blend_partial: 16-byte alignment");
    if (int rc = require_init()) return rc;
    PartialRows rows{};
    rows.n = R;
    for (int r = 0; r < R; ++r) {
        TMX_REQUIRE(row_ids[r] >= 0 && row_ids[r] <= K, TMX_EINVAL, "blend_partial: row id %d outside [0,%d]", row_ids[r], K);
        rows.ids[r] = row_ids[r];
        rows.w[r] = (weights && row_ids[r] > 0) ? weights[row_ids[r] - 1] : 1.f;
    }
    const long long total = (long long)imgs * C * (HW >> 3);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    switch (eps_dtype) {
        case TMX_F16:  blend_partial_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)eps_rows, masks, acc, imgs, C, HW, rows); break;
        case TMX_BF16: blend_partial_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)eps_rows, masks, acc, imgs, C, HW, rows); break;
        case TMX_F32:  blend_partial_kernel<float><<<blocks, 256, 0, st>>>((const float*)eps_rows, masks, acc, imgs, C, HW, rows); break;
        default: set_error("blend_partial: unsupported eps dtype %d", eps_dtype); return TMX_EDTYPE;
    }
    return check_cuda(cudaGetLastError(), "blend_partial_kernel launch");
}

extern "C" int tmx_blend_finish_fwd(const float* x, const float* acc, const float* masks, const float* weights,
                                    float* x_out, float* x0_out, int imgs, int K, int C, int HW,
                                    float a_t, float a_next, float g, int is_last, void* stream) {
    TMX_REQUIRE(x && acc && x_out, TMX_EINVAL, "blend_finish: null pointer");
    TMX_REQUIRE(imgs > 0 && C > 0 && HW > 0, TMX_EINVAL, "blend_finish: non-positive size");
    TMX_REQUIRE(K >= 1 && K <= kMaxConcepts, TMX_ESHAPE, "blend_finish: K=%d unsupported", K);
    TMX_REQUIRE(HW % 8 == 0, TMX_ESHAPE, "blend_finish: HW=%d must be a multiple of 8", HW);
    TMX_REQUIRE(aligned16(x) && aligned16(acc) && aligned16(masks) && aligned16(x_out) && aligned16(x0_out),
                TMX_EALIGN, "blend_finish: 16-byte alignment");
    TMX_REQUIRE(a_t > 0.f && a_t <= 1.f && a_next > 0.f && a_next <= 1.f, TMX_EINVAL, "blend_finish: bad alphas");
    if (int rc = require_init()) return rc;
    BlendCoef cf = make_coef(a_t, a_next, g, is_last, weights, K);
    const long long total = (long long)imgs * (HW >> 3);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    blend_finish_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, acc, masks, x_out, x0_out, imgs, K, C, HW, cf);
    return check_cuda(cudaGetLastError(), "blend_finish_kernel launch");
}
