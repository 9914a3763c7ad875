// k11 / k12 — the two pieces TweedieMix adds to the I2VGen-XL video loop (BASELINE configs[4]).
//
// k11  tmx_vpred_cfg_ddim_fwd : classifier-free guidance + v-prediction Tweedie estimate + deterministic DDIM update of the
//      video latents, one pass (video_gen/pipeline_i2vgen_xl.py:694-713):
//          v   = v_u + g (v_c - v_u)
//          eps = sqrt(a_t) v + sqrt(1 - a_t) x          x0 = sqrt(a_t) x - sqrt(1 - a_t) v
//          x'  = sqrt(a_next) x0 + sqrt(1 - a_next) eps
//      The reference spends 11 elementwise launches and two permute copies of the [B, C, F, H, W] latents per step; the op
//      is position-wise, so it runs on the flat memory with no permute.  HBM-bound: 3 reads + 1 (+1 for x0) write.
// k12  tmx_frame_inject_fwd   : frame-0 residual-feature injection on a ResNet output viewed as [b, t, frame]
//      (video_gen/utils_attn.py:433-456): y[b, t>=1] = interp * y[b, 0] + (1 - interp) * y[b, t]; interp = 1 is the plain
//      replacement of `injection_schedule`, 0 < interp < 1 the `injection_schedule2` blend.  In place, frame 0 untouched.
//
// TMX_ROUND_FP32: fp32 arithmetic on the loaded values, one rounding on store.  TMX_ROUND_REF: every product and sum is rounded
// to the tensor dtype where PyTorch's type promotion rounds it (all operands are 16-bit tensors, the schedule scalars are
// 0-dim fp32 tensors / Python floats and stay fp32), which reproduces the reference's fp16 pipeline bit for bit.
#include "tmx_common.cuh"

namespace tmx {
namespace k11 {

struct Coef { float g, sa, sb, sna, snb; };        // guidance, sqrt(a_t), sqrt(1-a_t), sqrt(a_next), sqrt(1-a_next)

template <typename T, bool REF>
__device__ __forceinline__ void vpred_one(float x, float vu, float vc, const Coef& c, float& xn, float& x0) {
    auto r = [](float v) { return REF ? Pack2<T>::round(v) : v; };
    const float v = r(vu + r(c.g * r(vc - vu)));
    const float eps = r(r(c.sa * v) + r(c.sb * x));
    x0 = r(r(c.sa * x) - r(c.sb * v));
    xn = r(r(c.sna * x0) + r(c.snb * eps));
}

template <typename T, bool REF>
__global__ void __launch_bounds__(256)
vpred_kernel(const T* __restrict__ x, const T* __restrict__ vu, const T* __restrict__ vc, T* __restrict__ xn, T* __restrict__ x0o,
             size_t nvec, Coef c) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float fx[8], fu[8], fc[8], o[8], z[8];
        unpack8<T>(ld_stream(x + i * 8), fx);
        unpack8<T>(ld_stream(vu + i * 8), fu);
        unpack8<T>(ld_stream(vc + i * 8), fc);
#pragma unroll
        for (int j = 0; j < 8; ++j) vpred_one<T, REF>(fx[j], fu[j], fc[j], c, o[j], z[j]);
        st_stream(xn + i * 8, pack8<T>(o));
        if (x0o) st_stream(x0o + i * 8, pack8<T>(z));
    }
}

template <bool REF>
__global__ void __launch_bounds__(256)
vpred_kernel_f32(const float* __restrict__ x, const float* __restrict__ vu, const float* __restrict__ vc, float* __restrict__ xn,
                 float* __restrict__ x0o, size_t nvec, Coef c) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const float4 a = reinterpret_cast<const float4*>(x)[i], u = reinterpret_cast<const float4*>(vu)[i], t = reinterpret_cast<const float4*>(vc)[i];
        float4 o, z;
        vpred_one<float, false>(a.x, u.x, t.x, c, o.x, z.x);
        vpred_one<float, false>(a.y, u.y, t.y, c, o.y, z.y);
        vpred_one<float, false>(a.z, u.z, t.z, c, o.z, z.z);
        vpred_one<float, false>(a.w, u.w, t.w, c, o.w, z.w);
        reinterpret_cast<float4*>(xn)[i] = o;
        if (x0o) reinterpret_cast<float4*>(x0o)[i] = z;
    }
}

// y viewed as [groups][T][fvec] vectors of 8: every thread blends one vector of frame t >= 1 with the same vector of frame 0
template <typename T, bool REF>
__global__ void __launch_bounds__(256)
inject_kernel(T* __restrict__ y, size_t fvec, int frames, size_t total, float interp) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t per_group = fvec * (size_t)(frames - 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t g = i / per_group, rem = i - g * per_group;
        const size_t t = rem / fvec + 1, v = rem - (t - 1) * fvec;
        const T* first = y + (g * frames) * fvec * 8 + v * 8;
        T* dst = y + (g * frames + t) * fvec * 8 + v * 8;
        float f0[8], ft[8];
        unpack8<T>(ld_keep(first), f0);
        unpack8<T>(ld_stream(dst), ft);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (REF) ft[j] = Pack2<T>::round(Pack2<T>::round(interp * f0[j]) + Pack2<T>::round((1.f - interp) * ft[j]));
            else ft[j] = interp * f0[j] + (1.f - interp) * ft[j];
        }
        st_stream(dst, interp == 1.f ? pack8<T>(f0) : pack8<T>(ft));
    }
}

static unsigned blocks_for(size_t n) {
    const size_t want = (n + 255) / 256, cap = (size_t)sm_count() * 16;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace k11
}  // namespace tmx

using namespace tmx;
using namespace tmx::k11;

extern "C" int tmx_vpred_cfg_ddim_fwd(const void* x, const void* v_uncond, const void* v_cond, void* x_next, void* x0_out,
                                      size_t n, float a_t, float a_next, float guidance, int dtype, int round_mode, void* stream) {
    TMX_REQUIRE(x && v_uncond && v_cond && x_next, TMX_EINVAL, "vpred: null pointer");
    TMX_REQUIRE(n > 0 && n % 8 == 0, TMX_ESHAPE, "vpred: element count %zu must be a positive multiple of 8", n);
    TMX_REQUIRE(a_t > 0.f && a_t <= 1.f && a_next > 0.f && a_next <= 1.f, TMX_EINVAL, "vpred: alphas must lie in (0, 1]");
    TMX_REQUIRE(round_mode == TMX_ROUND_FP32 || round_mode == TMX_ROUND_REF, TMX_EINVAL, "vpred: bad round_mode");
    TMX_REQUIRE(aligned16(x) && aligned16(v_uncond) && aligned16(v_cond) && aligned16(x_next) && aligned16(x0_out), TMX_EALIGN, "vpred: 16-byte alignment");
    if (int rc = require_init()) return rc;
    Coef c{guidance, sqrtf(a_t), sqrtf(1.f - a_t), sqrtf(a_next), sqrtf(1.f - a_next)};
    cudaStream_t st = (cudaStream_t)stream;
    const bool ref = round_mode == TMX_ROUND_REF;
    switch (dtype) {
        case TMX_F16:
            if (ref) vpred_kernel<__half, true><<<blocks_for(n / 8), 256, 0, st>>>((const __half*)x, (const __half*)v_uncond, (const __half*)v_cond, (__half*)x_next, (__half*)x0_out, n / 8, c);
            else     vpred_kernel<__half, false><<<blocks_for(n / 8), 256, 0, st>>>((const __half*)x, (const __half*)v_uncond, (const __half*)v_cond, (__half*)x_next, (__half*)x0_out, n / 8, c);
            break;
        case TMX_BF16:
            if (ref) vpred_kernel<__nv_bfloat16, true><<<blocks_for(n / 8), 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)v_uncond, (const __nv_bfloat16*)v_cond, (__nv_bfloat16*)x_next, (__nv_bfloat16*)x0_out, n / 8, c);
            else     vpred_kernel<__nv_bfloat16, false><<<blocks_for(n / 8), 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)v_uncond, (const __nv_bfloat16*)v_cond, (__nv_bfloat16*)x_next, (__nv_bfloat16*)x0_out, n / 8, c);
            break;
        case TMX_F32:
            vpred_kernel_f32<false><<<blocks_for(n / 4), 256, 0, st>>>((const float*)x, (const float*)v_uncond, (const float*)v_cond, (float*)x_next, (float*)x0_out, n / 4, c);
            break;
        default:
            set_error("vpred: unsupported dtype %d", dtype);
            return TMX_EDTYPE;
    }
    return check_cuda(cudaGetLastError(), "vpred_kernel launch");
}

extern "C" int tmx_frame_inject_fwd(void* y, int groups, int frames, size_t frame_elems, float interp, int dtype, int round_mode, void* stream) {
    TMX_REQUIRE(y, TMX_EINVAL, "frame_inject: null pointer");
    TMX_REQUIRE(groups > 0 && frames >= 1 && frame_elems > 0 && frame_elems % 8 == 0, TMX_ESHAPE,
                "frame_inject: groups=%d frames=%d frame_elems=%zu (multiple of 8)", groups, frames, frame_elems);
    TMX_REQUIRE(interp >= 0.f && interp <= 1.f, TMX_EINVAL, "frame_inject: interp must lie in [0, 1]");
    TMX_REQUIRE(dtype == TMX_F16 || dtype == TMX_BF16, TMX_EDTYPE, "frame_inject: dtype %d unsupported (fp16/bf16 only)", dtype);
    TMX_REQUIRE(round_mode == TMX_ROUND_FP32 || round_mode == TMX_ROUND_REF, TMX_EINVAL, "frame_inject: bad round_mode");
    TMX_REQUIRE(aligned16(y), TMX_EALIGN, "frame_inject: 16-byte alignment");
    if (int rc = require_init()) return rc;
    if (frames == 1) return TMX_OK;
    const size_t fvec = frame_elems / 8, total = (size_t)groups * (frames - 1) * fvec;
    cudaStream_t st = (cudaStream_t)stream;
    const bool ref = round_mode == TMX_ROUND_REF;
    if (dtype == TMX_F16) {
        if (ref) inject_kernel<__half, true><<<blocks_for(total), 256, 0, st>>>((__half*)y, fvec, frames, total, interp);
        else     inject_kernel<__half, false><<<blocks_for(total), 256, 0, st>>>((__half*)y, fvec, frames, total, interp);
    } else {
        if (ref) inject_kernel<__nv_bfloat16, true><<<blocks_for(total), 256, 0, st>>>((__nv_bfloat16*)y, fvec, frames, total, interp);
        else     inject_kernel<__nv_bfloat16, false><<<blocks_for(total), 256, 0, st>>>((__nv_bfloat16*)y, fvec, frames, total, interp);
    }
    return check_cuda(cudaGetLastError(), "inject_kernel launch");
}
