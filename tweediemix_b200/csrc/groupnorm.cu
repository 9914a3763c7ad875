// k4 / k5 — GroupNorm (+ per-(n,c) additive bias before the norm, + SiLU), NHWC and NCHW.
//
// HBM-bound.  Activations that fit in the shared memory of one wave (N * floor(#SM / N) CTAs, one per SM,
// <= ~190 KB each: every 32x32 and 64x64 site of the SDXL U-Net at batch 4) take the FUSED path: one
// cooperative launch loads the CTA's contiguous NHWC slab once, keeps it in shared memory while the CTAs
// of a batch row meet at a sense-reversing barrier in global memory (the last arriver folds the partial
// sums), then normalises out of shared memory: x is read once and y written once (2 * s * elems, the
// algorithmic minimum).  Larger activations take two launches: `stats` reads x once (128-bit coalesced loads, 4 rows in flight per
// thread) and writes per-CTA shifted partial sums; `apply` re-reads x — from the 126 MB L2 for
// every SDXL activation that fits — folds the statistics, gamma, beta and the additive bias into
// one fma per element, applies SiLU and stores once.  Statistics are fp32 sums of (x - pivot)
// with a per-(n,group) pivot taken from the data, so E[d^2] - E[d]^2 has no catastrophic
// cancellation; partials are combined in a fixed order (deterministic, no atomics).
//
// Replaces [D] diffusers ResnetBlock2D norm1/norm2 + SiLU (mirrored in the reference at
// video_gen/utils_attn.py:391-431), conv_norm_out + SiLU, and Transformer2DModel.norm (no act).
#include "tmx_common.cuh"
#include <cstdlib>
#include <type_traits>

namespace tmx {

constexpr int kGnMaxParts = 1024;     // upper bound on partials per (n, group)
constexpr int kGnMaxThreads = 512;
// Rows in flight per thread and resident CTAs per SM of the two-launch kernels.  Measured (profiles/r02z_kbench_groupnorm_two_launch.txt):
// at 2 CTAs per SM the 64-register cap spilled the row buffers of both kernels (72 / 36 bytes of spill stores inside the streaming
// loops); ONE 480-512-thread CTA per SM with 8 rows (128 B) in flight per thread and no spills is 12-25 % faster at every site.
#ifndef TMX_GN_U
#define TMX_GN_U 8
#endif
#ifndef TMX_GN_US
#define TMX_GN_US TMX_GN_U
#endif
constexpr int kGnU = TMX_GN_U;          // rows in flight per thread in the apply kernel of the two-launch path
constexpr int kGnUS = TMX_GN_US;        // ... in the stats kernel (no output registers: more rows fit)
#ifndef TMX_GN_MINB
#define TMX_GN_MINB 1
#endif

template <typename T> struct V8 {
    static constexpr int kBytes = 16;
    static __device__ __forceinline__ void load(const T* p, float (&f)[8], bool keep) {
        unpack8<T>(keep ? ld_keep(p) : ld_stream(p), f);
    }
    static __device__ __forceinline__ void store(T* p, const float (&f)[8]) { st_stream(p, pack8<T>(f)); }
};
template <> struct V8<float> {
    static __device__ __forceinline__ void load(const float* p, float (&f)[8], bool) {
        float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
        st_stream(p, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
        st_stream(p + 4, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
    }
};

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

__device__ __forceinline__ float silu(float y) { return __fdividef(y, 1.f + __expf(-y)); }
// y * sigmoid(y) = h + h * tanh(h), h = y / 2: ONE MUFU op per element (tanh.approx, rel. error ~2^-11) instead of
// two (ex2 + rcp).  The apply pass of a 16-bit activation is MUFU-bound otherwise; the error is below the output rounding.
__device__ __forceinline__ float silu_fast(float y) {
    const float h = 0.5f * y;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}
template <typename T> __device__ __forceinline__ float silu_for(float y) { return sizeof(T) == 2 ? silu_fast(y) : silu(y); }



// Two-source input (the channel concatenation [x | x2] of an up-block ResNet, never materialised): channel vector tx of row r lives
// in x (row stride C1) when tx*8 < C1, else in x2 (row stride C - C1).  x2 == nullptr: one source with row stride C.
template <typename T>
struct GnSrc {
    const T* col;       // this thread's channel vector at pixel 0 of sample n
    size_t cs;          // row (pixel) stride of its source, in elements
    __device__ __forceinline__ GnSrc(const T* x, const T* x2, int C, int C1, int HW, int n, int tx) {
        const int c0 = tx * 8;
        if (x2 != nullptr && c0 >= C1) { cs = (size_t)(C - C1); col = x2 + (size_t)n * HW * cs + (c0 - C1); }
        else { cs = (size_t)(x2 != nullptr ? C1 : C); col = x + (size_t)n * HW * cs + c0; }
    }
};
// value of channel c at pixel 0 of sample n (pivot of the shifted sums)
template <typename T>
__device__ __forceinline__ float gn_first(const T* x, const T* x2, int C, int C1, int HW, int n, int c) {
    if (x2 != nullptr && c >= C1) return to_f32<T>(x2[(size_t)n * HW * (C - C1) + (c - C1)]);
    return to_f32<T>(x[(size_t)n * HW * (x2 != nullptr ? C1 : C) + c]);
}

// Per-CTA reduction of the per-thread channel sums (S, SS)[8] to per-group sums, deterministic order.
// sh: [8][RY][CV] float2 (conflict-free: consecutive threads write consecutive words), then [C] float2 channel totals.
// Stage A: the first CV threads add the RY row lanes of their 8 channels; stage B: one warp per group adds its cpg
// channels (lanes stride the channels: no integer division anywhere).  Returns with the group sum valid in lane 0
// through the callback `emit(g, s, ss)`.
template <typename F>
__device__ __forceinline__ void gn_cta_group_sums(float2* sh, const float (&S)[8], const float (&SS)[8],
                                                  int C, int CV, int RY, int G, int cpg, int tx, int ty, F emit) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[((size_t)j * RY + ty) * CV + tx] = make_float2(S[j], SS[j]);
    __syncthreads();
    float2* ch = sh + (size_t)8 * RY * CV;                // [C] channel totals
    if (threadIdx.x < CV) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float s = 0.f, ss = 0.f;
            for (int t = 0; t < RY; ++t) { const float2 v = sh[((size_t)j * RY + t) * CV + threadIdx.x]; s += v.x; ss += v.y; }
            ch[threadIdx.x * 8 + j] = make_float2(s, ss);
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int g = warp; g < G; g += nwarps) {
        float s = 0.f, ss = 0.f;
        for (int i = lane; i < cpg; i += 32) { const float2 v = ch[g * cpg + i]; s += v.x; ss += v.y; }
        s = warp_sum(s); ss = warp_sum(ss);
        if (lane == 0) emit(g, s, ss);
    }
}

// ------------------------------------------------------------------------------------------ NHWC
// grid (P, N); block = CV * RY threads (CV = C/8 channel vectors, RY row lanes), 2 CTAs per SM.
// Thread (ty, tx) owns channel vector tx and rows r = row0 + ty, + RY, ...
//
// stats: per-CTA shifted partial sums -> part[(n*G+g)*P + p]; the LAST CTA of a row n to finish
// (ticket counter in the workspace) folds the P partials of every group in a fixed order and
// publishes stat[n*G+g] = (mean, rstd), so `apply` starts with two floats per group instead of
// re-reducing the partials in every CTA.
template <typename T>
__global__ void __launch_bounds__(kGnMaxThreads, TMX_GN_MINB)
gn_stats_nhwc(const T* __restrict__ x, const T* __restrict__ x2, int C1, const float* __restrict__ add, float2* __restrict__ part,
              float2* __restrict__ stat, unsigned int* __restrict__ tickets,
              int C, int HW, int G, int rows_per_cta, float eps) {
    extern __shared__ float2 sh[];                      // [8][RY][CV] partials + [C] channel totals (gn_cta_group_sums)
    __shared__ float s_piv[64];
    __shared__ unsigned int s_ticket;
    const int CV = C >> 3, RY = blockDim.x / CV, cpg = C / G;
    const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
    const int n = blockIdx.y, P = gridDim.x;
    const int row0 = blockIdx.x * rows_per_cta;
    const int row1 = min(HW, row0 + rows_per_cta);
    const GnSrc<T> src(x, x2, C, C1, HW, n, tx);
    const T* col = src.col;
    const size_t cs = src.cs;

    // first batch of row loads goes out before anything else (the pivot staging below overlaps it)
    int r = row0 + ty;
    uint4 q[kGnUS];
    const bool full0 = (sizeof(T) == 2) && (r + (kGnUS - 1) * RY < row1);
    if (full0) {
#pragma unroll
        for (int u = 0; u < kGnUS; ++u) q[u] = ld_keep(col + (size_t)(r + u * RY) * cs);
    }
    // pivot = x'[n, pixel 0, first channel of the group]: sums of (x' - pivot) do not cancel
    for (int g = threadIdx.x; g < G; g += blockDim.x)
        s_piv[g] = gn_first<T>(x, x2, C, C1, HW, n, g * cpg) + (add ? add[(size_t)n * C + g * cpg] : 0.f);
    __syncthreads();
    float kk[8], S[8], SS[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = tx * 8 + j;
        kk[j] = s_piv[c / cpg] - (add ? add[(size_t)n * C + c] : 0.f);
        S[j] = 0.f; SS[j] = 0.f;
    }
    if constexpr (sizeof(T) == 2) {
        if (full0) {
            while (true) {
                const int rn = r + kGnUS * RY;
                const bool more = rn + (kGnUS - 1) * RY < row1;
                uint4 qn[kGnUS];
                if (more) {
#pragma unroll
                    for (int u = 0; u < kGnUS; ++u) qn[u] = ld_keep(col + (size_t)(rn + u * RY) * cs);
                }
#pragma unroll
                for (int u = 0; u < kGnUS; ++u) {
                    float f[8];
                    unpack8<T>(q[u], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const float d = f[j] - kk[j]; S[j] += d; SS[j] = fmaf(d, d, SS[j]); }
                }
                r = rn;
                if (!more) break;
#pragma unroll
                for (int u = 0; u < kGnUS; ++u) q[u] = qn[u];
            }
        }
    }
    for (; r < row1; r += RY) {
        float f0[8];
        V8<T>::load(col + (size_t)r * cs, f0, true);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d0 = f0[j] - kk[j]; S[j] += d0; SS[j] = fmaf(d0, d0, SS[j]); }
    }
    gn_cta_group_sums(sh, S, SS, C, CV, RY, G, cpg, tx, ty,
                      [&](int g, float s, float ss) { part[((size_t)n * G + g) * P + blockIdx.x] = make_float2(s, ss); });
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;   // blockDim.x is a multiple of 32 (plan_nhwc)
    // ---- last CTA of this row finalises the statistics
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[n], 1u);
    __syncthreads();
    if (s_ticket != (unsigned)(P - 1)) return;
    __threadfence();
    const float inv = 1.f / ((float)HW * (float)cpg);
    for (int g = warp; g < G; g += nwarps) {
        float s = 0.f, ss = 0.f;
        const float2* pp = part + ((size_t)n * G + g) * P;
        for (int i = lane; i < P; i += 32) { const float2 v = __ldcg(pp + i); s += v.x; ss += v.y; }
        s = warp_sum(s); ss = warp_sum(ss);
        if (lane == 0) {
            const float md = s * inv;
            const float var = fmaxf(ss * inv - md * md, 0.f);
            stat[n * G + g] = make_float2(s_piv[g] + md, rsqrtf(var + eps));
        }
    }
    if (threadIdx.x == 0) tickets[n] = 0u;              // leave the workspace ready for the next launch
}

template <typename T>
__global__ void __launch_bounds__(kGnMaxThreads, TMX_GN_MINB)
gn_apply_nhwc(const T* __restrict__ x, const T* __restrict__ x2, int C1, const float* __restrict__ gamma, const float* __restrict__ beta,
              const float* __restrict__ add, const float2* __restrict__ stat, T* __restrict__ y,
              int C, int HW, int G, int rows_per_cta, int act) {
    const int CV = C >> 3, RY = blockDim.x / CV, cpg = C / G;
    const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
    const int n = blockIdx.y;
    const int row0 = blockIdx.x * rows_per_cta, row1 = min(HW, row0 + rows_per_cta);
    const GnSrc<T> src(x, x2, C, C1, HW, n, tx);
    const T* col = src.col;
    const size_t cs = src.cs;
    T* ycol = y + (size_t)n * HW * C + (size_t)tx * 8;
    int r = row0 + ty;
    uint4 q[kGnU];
    const bool full0 = (sizeof(T) == 2) && (r + (kGnU - 1) * RY < row1);
    if (full0) {
#pragma unroll
        for (int u = 0; u < kGnU; ++u) q[u] = ld_stream(col + (size_t)(r + u * RY) * cs);
    }
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = tx * 8 + j;
        const float2 st = __ldcg(stat + n * G + c / cpg);
        a[j] = st.y * gamma[c];
        const float adj = add ? add[(size_t)n * C + c] : 0.f;
        b[j] = fmaf(adj - st.x, a[j], beta[c]);
    }
    if constexpr (sizeof(T) == 2) {
        if (full0) {
            while (true) {
                const int rn = r + kGnU * RY;
                const bool more = rn + (kGnU - 1) * RY < row1;
                uint4 qn[kGnU];
                if (more) {
#pragma unroll
                    for (int u = 0; u < kGnU; ++u) qn[u] = ld_stream(col + (size_t)(rn + u * RY) * cs);
                }
#pragma unroll
                for (int u = 0; u < kGnU; ++u) {
                    float f[8];
                    unpack8<T>(q[u], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float v = fmaf(f[j], a[j], b[j]);
                        f[j] = act ? silu_for<T>(v) : v;
                    }
                    st_stream(ycol + (size_t)(r + u * RY) * C, pack8<T>(f));
                }
                r = rn;
                if (!more) break;
#pragma unroll
                for (int u = 0; u < kGnU; ++u) q[u] = qn[u];
            }
        }
    }
    for (; r < row1; r += RY) {
        float f[8];
        V8<T>::load(col + (size_t)r * cs, f, false);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = fmaf(f[j], a[j], b[j]);
            f[j] = act ? silu_for<T>(v) : v;
        }
        V8<T>::store(ycol + (size_t)r * C, f);
    }
}


// --------------------------------------------------------------------------------- NHWC, fused
// grid (P, N) <= #SM CTAs, launched cooperatively (all co-resident); block = CV * RY threads as above.
// smem: [rows_per_cta][CV] uint4 slab | [RY][C] float2 reduction scratch.
// sync[n] = arrival counter of batch row n (left at zero), sync[1024 + n] = barrier generation (zero-initialised once).
template <typename T>
__global__ void __launch_bounds__(kGnMaxThreads, 1)
gn_fused_nhwc(const T* __restrict__ x, const T* __restrict__ x2, int C1, const float* __restrict__ gamma, const float* __restrict__ beta,
              const float* __restrict__ add, float2* __restrict__ part, float2* __restrict__ stat,
              unsigned int* __restrict__ sync, T* __restrict__ y,
              int C, int HW, int G, int rows_per_cta, float eps, int act) {
    extern __shared__ uint4 gn_smem[];
    __shared__ float s_piv[64];
    __shared__ float2 s_stat[64];
    const int CV = C >> 3, RY = blockDim.x / CV, cpg = C / G;
    const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
    const int n = blockIdx.y, P = gridDim.x;
    const int row0 = blockIdx.x * rows_per_cta;
    const int row1 = min(HW, row0 + rows_per_cta);
    uint4* slab = gn_smem;
    float2* sh = reinterpret_cast<float2*>(gn_smem + (size_t)rows_per_cta * CV);
    const GnSrc<T> src(x, x2, C, C1, HW, n, tx);
    const T* col = src.col;
    const size_t cs = src.cs;
    const bool active = ty < RY;                          // blockDim.x == CV * RY exactly, kept for clarity
    unsigned int* cnt = sync + n;
    unsigned int* gen = sync + 1024 + n;
    unsigned int gen0 = 0;
    if (threadIdx.x == 0) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen0) : "l"(gen) : "memory");

    // ---- phase 1: one pass over the slab: global -> registers -> (statistics, shared memory)
    int r = row0 + ty;
    uint4 q[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) if (active && r + u * RY < row1) q[u] = ld_stream(col + (size_t)(r + u * RY) * cs);
    for (int g = threadIdx.x; g < G; g += blockDim.x)
        s_piv[g] = gn_first<T>(x, x2, C, C1, HW, n, g * cpg) + (add ? add[(size_t)n * C + g * cpg] : 0.f);
    __syncthreads();
    float kk[8], S[8], SS[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = tx * 8 + j;
        kk[j] = s_piv[c / cpg] - (add ? add[(size_t)n * C + c] : 0.f);
        S[j] = 0.f; SS[j] = 0.f;
    }
    while (r < row1) {
        const int rn = r + 8 * RY;
        uint4 qn[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) if (rn + u * RY < row1) qn[u] = ld_stream(col + (size_t)(rn + u * RY) * cs);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (r + u * RY < row1) {
                slab[(size_t)(r + u * RY - row0) * CV + tx] = q[u];
                float f[8];
                unpack8<T>(q[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float d = f[j] - kk[j]; S[j] += d; SS[j] = fmaf(d, d, SS[j]); }
            }
        }
        r = rn;
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = qn[u];
    }
    gn_cta_group_sums(sh, S, SS, C, CV, RY, G, cpg, tx, ty,
                      [&](int g, float s, float ss) { part[((size_t)n * G + g) * P + blockIdx.x] = make_float2(s, ss); });
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;   // blockDim.x is a multiple of 32 (plan_fused)
    // ---- sense-reversing barrier over the P CTAs of batch row n (arrival counter re-armed by the last arriver,
    // generation word read at kernel entry), then EVERY CTA folds the P partials itself in the same fixed order
    // (bit-identical statistics everywhere, and no finalise -> publish -> re-read round trips).
    float gm[8], bt[8], ad[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {                         // issued before the wait: off the critical path
        const int c = tx * 8 + j;
        gm[j] = gamma[c]; bt[j] = beta[c];
        ad[j] = add ? add[(size_t)n * C + c] : 0.f;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(cnt, 1u) == (unsigned)(P - 1)) {                   // last arriver: re-arm the counter, open the barrier
            *cnt = 0u;
            asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(gen), "r"(gen0 + 1u) : "memory");
        } else {
            unsigned int cur, polls = 0;
            long long t0 = 0;
            while (true) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(gen) : "memory");
                if (cur != gen0) break;
                if ((++polls & 1023u) == 0) {                             // bounded: a lost wake-up must trap, not hang
                    const long long now = clock64();
                    if (t0 == 0) t0 = now; else if (now - t0 > 4000000000LL) __trap();
                }
            }
        }
    }
    __syncthreads();
    {
        const float inv = 1.f / ((float)HW * (float)cpg);
        float s4[4] = {0.f, 0.f, 0.f, 0.f}, ss4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 4; ++k) {                     // G <= 64 and >= 16 warps... up to 4 groups per warp, loads independent
            const int g = warp + k * nwarps;
            if (g < G) {
                const float2* pp = part + ((size_t)n * G + g) * P;
                for (int i = lane; i < P; i += 32) { const float2 v = __ldcg(pp + i); s4[k] += v.x; ss4[k] += v.y; }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int g = warp + k * nwarps;
            if (g < G) {
                const float s = warp_sum(s4[k]), ss = warp_sum(ss4[k]);
                if (lane == 0) {
                    const float md = s * inv;
                    const float var = fmaxf(ss * inv - md * md, 0.f);
                    s_stat[g] = make_float2(s_piv[g] + md, rsqrtf(var + eps));
                }
            }
        }
        for (int g = warp + 4 * nwarps; g < G; g += nwarps) {              // (only for blocks of fewer than G/4 warps)
            float s = 0.f, ss = 0.f;
            const float2* pp = part + ((size_t)n * G + g) * P;
            for (int i = lane; i < P; i += 32) { const float2 v = __ldcg(pp + i); s += v.x; ss += v.y; }
            s = warp_sum(s); ss = warp_sum(ss);
            if (lane == 0) {
                const float md = s * inv;
                s_stat[g] = make_float2(s_piv[g] + md, rsqrtf(fmaxf(ss * inv - md * md, 0.f) + eps));
            }
        }
    }
    __syncthreads();

    // ---- phase 2: normalise out of shared memory
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float2 st = s_stat[(tx * 8 + j) / cpg];
        a[j] = st.y * gm[j];
        b[j] = fmaf(ad[j] - st.x, a[j], bt[j]);
    }
    T* ycol = y + (size_t)n * HW * C + (size_t)tx * 8;
    for (int rr = row0 + ty; rr < row1; rr += RY) {
        float f[8];
        unpack8<T>(slab[(size_t)(rr - row0) * CV + tx], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float v = fmaf(f[j], a[j], b[j]);
            f[j] = act ? silu_for<T>(v) : v;
        }
        st_stream(ycol + (size_t)rr * C, pack8<T>(f));
    }
}

// ------------------------------------------------------- NHWC, one CTA (or one CLUSTER) per (sample, group)
// The (n, g) slab — cpg channels x HW pixels — is owned by ONE thread-block cluster of CS = 1, 2, 4 or 8 CTAs, each holding a
// contiguous range of pixels (<= ~200 KB) in its shared memory: loaded once (runs of cpg * 2 bytes per pixel; 4-, 8- or 16-byte
// vectors), reduced inside the CTA, the CS partial sums exchanged through DISTRIBUTED shared memory behind a cluster barrier,
// normalised out of shared memory.  No grid barrier, no partials in global memory, no cooperative launch, x read once and y
// written once at every SDXL site (the 128x128 sites took two launches and a re-read of x before): the cost is one launch + one
// load + one block (+ cluster) reduction + one store.
// grid (G * CS, N), cluster (CS, 1, 1); block = a multiple of lcm(vpp, 32) threads (vpp = vectors per pixel run), so every thread
// keeps ONE channel vector: its gamma / beta / add and its source (two-source input [x | x2], never concatenated) are fixed.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int VL> struct GnVec;
template <> struct GnVec<8> { using type = uint4; };
template <> struct GnVec<4> { using type = uint2; };
template <> struct GnVec<2> { using type = uint32_t; };

template <typename T, int VL>        // VL = elements per vector: 8 (cpg % 8 == 0), 4 (cpg % 4 == 0) or 2 (cpg % 2 == 0)
__global__ void __launch_bounds__(1024, 1)
gn_group_slab(const T* __restrict__ x, const T* __restrict__ x2, int C1, const float* __restrict__ gamma, const float* __restrict__ beta,
              const float* __restrict__ add, T* __restrict__ y, int C, int HW, int cpg, int CS, int ppc, float eps, int act) {
    using VT = typename GnVec<VL>::type;
    extern __shared__ uint4 gn_smem[];
    __shared__ float2 s_red[32];
    __shared__ float2 s_part;                          // this CTA's shifted sums, read by the whole cluster
    __shared__ float2 s_stat;
    VT* slab = reinterpret_cast<VT*>(gn_smem);
    const int rank = CS > 1 ? (int)cluster_ctarank() : 0;
    const int g = blockIdx.x / CS, n = blockIdx.y;
    const int pbeg = rank * ppc, pend = min(HW, pbeg + ppc);          // this CTA's pixels
    const int vpp = cpg / VL;                          // vectors per pixel run
    const int nt = blockDim.x;                         // multiple of vpp
    const int v = threadIdx.x % vpp;                   // this thread's vector inside the run
    const int c0 = g * cpg + v * VL;                   // its first channel
    const int pstep = nt / vpp;
    const int pix0 = pbeg + threadIdx.x / vpp;
    // source of this channel vector
    const T* src;
    size_t cs;
    if (x2 != nullptr && c0 >= C1) { cs = (size_t)(C - C1); src = x2 + (size_t)n * HW * cs + (c0 - C1); }
    else { cs = (size_t)(x2 != nullptr ? C1 : C); src = x + (size_t)n * HW * cs + c0; }
    // pivot of the shifted sums: the group's first channel at pixel 0 (+ its add), the same for every thread of the cluster
    float piv;
    {
        const int cg = g * cpg;
        piv = gn_first<T>(x, x2, C, C1, HW, n, cg) + (add ? add[(size_t)n * C + cg] : 0.f);
    }
    float ad[VL], kk[VL];
#pragma unroll
    for (int j = 0; j < VL; ++j) { ad[j] = add ? add[(size_t)n * C + c0 + j] : 0.f; kk[j] = piv - ad[j]; }

    auto unpack = [&](const VT& q, float (&f)[VL]) {
        if constexpr (VL == 8) unpack8<T>(q, f);
        else if constexpr (VL == 4) { float2 a = Pack2<T>::unpack(q.x), b = Pack2<T>::unpack(q.y); f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; }
        else { float2 a = Pack2<T>::unpack(q); f[0] = a.x; f[1] = a.y; }
    };
    auto ldv = [&](int pix) -> VT {
        const T* p = src + (size_t)pix * cs;
        if constexpr (VL == 8) return ld_stream(p);
        else if constexpr (VL == 4) { uint2 r; asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p)); return r; }
        else { uint32_t r; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p)); return r; }
    };
    // ---- phase 1: global -> registers -> (shifted sums, shared slab), 64 bytes (32 for 4-byte vectors) in flight per thread.
    // (Measured: staging the whole slab by cp.async in one round trip is not faster — 10.1 vs 9.4 us at C=1280, 32x32 — the
    // kernel is bound by launch latency, the two half-duplex phases of a CTA and, with SiLU, ~1.4 us of MUFU per 80 KB slab.)
    float S = 0.f, SS = 0.f;
    constexpr int U = VL == 8 ? 4 : 8;
    for (int pix = pix0; pix < pend; pix += U * pstep) {
        VT q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (pix + u * pstep < pend) q[u] = ldv(pix + u * pstep);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pp = pix + u * pstep;
            if (pp < pend) {
                slab[(size_t)(pp - pbeg) * vpp + v] = q[u];
                float f[VL];
                unpack(q[u], f);
#pragma unroll
                for (int j = 0; j < VL; ++j) { const float d = f[j] - kk[j]; S += d; SS = fmaf(d, d, SS); }
            }
        }
    }
    // ---- block reduction (fixed order: lanes by shuffle tree, warps in index order), then the cluster's CTAs in rank order
    S = warp_sum(S); SS = warp_sum(SS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = nt >> 5;
    if (lane == 0) s_red[warp] = make_float2(S, SS);
    float gm[VL], bt[VL];
#pragma unroll
    for (int j = 0; j < VL; ++j) { gm[j] = gamma[c0 + j]; bt[j] = beta[c0 + j]; }
    __syncthreads();
    if (warp == 0) {
        float2 t = lane < nwarps ? s_red[lane] : make_float2(0.f, 0.f);
        const float s = warp_sum(t.x), ss = warp_sum(t.y);
        if (lane == 0) s_part = make_float2(s, ss);
    }
    if (CS > 1) { cluster_arrive(); cluster_wait(); } else __syncthreads();      // every CTA's s_part is written and visible
    if (warp == 0) {
        float s = 0.f, ss = 0.f;
        if (CS > 1) {
            if (lane < CS) {                                                     // one DSMEM load per lane, summed by a fixed shuffle tree
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((uint32_t)__cvta_generic_to_shared(&s_part)), "r"(lane));
                asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(s), "=f"(ss) : "r"(remote) : "memory");
            }
            s = warp_sum(s); ss = warp_sum(ss);
        } else { s = s_part.x; ss = s_part.y; }
        if (lane == 0) {
            const float inv = 1.f / ((float)HW * (float)cpg);
            const float md = s * inv;
            s_stat = make_float2(piv + md, rsqrtf(fmaxf(ss * inv - md * md, 0.f) + eps));
        }
    }
    if (CS > 1) cluster_arrive();                      // "I have read my peers' s_part": waited for at the end, before any CTA may exit
    __syncthreads();
    const float2 st = s_stat;
    float a[VL], b[VL];
#pragma unroll
    for (int j = 0; j < VL; ++j) { a[j] = st.y * gm[j]; b[j] = fmaf(ad[j] - st.x, a[j], bt[j]); }

    // ---- phase 2: normalise out of shared memory (each thread re-reads exactly the vectors it wrote: no barrier needed for them)
    T* ycol = y + (size_t)n * HW * C + c0;
    for (int pix = pix0; pix < pend; pix += pstep) {
        float f[VL];
        unpack(slab[(size_t)(pix - pbeg) * vpp + v], f);
#pragma unroll
        for (int j = 0; j < VL; ++j) {
            const float val = fmaf(f[j], a[j], b[j]);
            f[j] = act ? silu_for<T>(val) : val;
        }
        T* dst = ycol + (size_t)pix * C;
        if constexpr (VL == 8) st_stream(dst, pack8<T>(f));
        else if constexpr (VL == 4) {
            const uint32_t w0 = Pack2<T>::pack(f[0], f[1]), w1 = Pack2<T>::pack(f[2], f[3]);
            asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" :: "l"(dst), "r"(w0), "r"(w1) : "memory");
        } else {
            asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(dst), "r"(Pack2<T>::pack(f[0], f[1])) : "memory");
        }
    }
    if (CS > 1) cluster_wait();
}

// ------------------------------------------------------------------------------------------ NCHW
// (n, g) slab = cpg * HW contiguous elements.  grid (S, N*G), 256 threads, vectors of 8.
template <typename T>
__global__ void __launch_bounds__(256)
gn_stats_nchw(const T* __restrict__ x, const float* __restrict__ add, float2* __restrict__ part,
              float* __restrict__ pivots, int C, int HW, int G, int vec_per_cta) {
    __shared__ float2 red[8];
    const int cpg = C / G, ng = blockIdx.y, n = ng / G, g = ng % G, S = gridDim.x;
    const T* slab = x + ((size_t)n * C + (size_t)g * cpg) * HW;
    const float* addn = add ? add + (size_t)n * C + g * cpg : nullptr;
    const float piv = to_f32<T>(slab[0]) + (addn ? addn[0] : 0.f);
    const int nvec = (cpg * HW) >> 3;
    const int v0 = blockIdx.x * vec_per_cta, v1 = min(nvec, v0 + vec_per_cta);
    float s = 0.f, ss = 0.f;
    for (int v = v0 + threadIdx.x; v < v1; v += 256) {
        float f[8];
        V8<T>::load(slab + (size_t)v * 8, f, true);
        const float k = piv - (addn ? addn[(v * 8) / HW] : 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = f[j] - k; s += d; ss = fmaf(d, d, ss); }
    }
    s = warp_sum(s); ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = make_float2(s, ss);
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < 8; ++w) { a += red[w].x; b += red[w].y; }
        part[(size_t)ng * S + blockIdx.x] = make_float2(a, b);
        if (blockIdx.x == 0) pivots[ng] = piv;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
gn_apply_nchw(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              const float* __restrict__ add, const float2* __restrict__ part,
              const float* __restrict__ pivots, T* __restrict__ y,
              int C, int HW, int G, int vec_per_cta, int S_stats, float eps, int act) {
    __shared__ float s_mr[2];
    const int cpg = C / G, ng = blockIdx.y, n = ng / G, g = ng % G;
    const T* slab = x + ((size_t)n * C + (size_t)g * cpg) * HW;
    T* yslab = y + ((size_t)n * C + (size_t)g * cpg) * HW;
    const float* addn = add ? add + (size_t)n * C + g * cpg : nullptr;
    if (threadIdx.x == 0) {
        float s = 0.f, ss = 0.f;
        for (int i = 0; i < S_stats; ++i) { const float2 v = part[(size_t)ng * S_stats + i]; s += v.x; ss += v.y; }
        const float inv = 1.f / ((float)HW * (float)cpg);
        const float md = s * inv;
        s_mr[0] = pivots[ng] + md;
        s_mr[1] = rsqrtf(fmaxf(ss * inv - md * md, 0.f) + eps);
    }
    __syncthreads();
    const float mean = s_mr[0], rstd = s_mr[1];
    const int nvec = (cpg * HW) >> 3;
    const int v0 = blockIdx.x * vec_per_cta, v1 = min(nvec, v0 + vec_per_cta);
    for (int v = v0 + threadIdx.x; v < v1; v += 256) {
        const int cl = (v * 8) / HW, c = g * cpg + cl;
        const float a = rstd * gamma[c];
        const float b = fmaf((addn ? addn[cl] : 0.f) - mean, a, beta[c]);
        float f[8];
        V8<T>::load(slab + (size_t)v * 8, f, false);
#pragma unroll
        for (int j = 0; j < 8; ++j) { float t = fmaf(f[j], a, b); f[j] = act ? silu_for<T>(t) : t; }
        V8<T>::store(yslab + (size_t)v * 8, f);
    }
}

struct GnPlan { int threads, RY, rows_per_cta, P; };

static int gcd_int(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }
// Row lanes per CTA: as many as fit in kGnMaxThreads with CV * RY a multiple of 32 (the per-group reductions
// use full-warp shuffles); 0 if C admits no such block (then the shape is rejected).
static int gn_row_lanes(int CV) {
    const int mult = 32 / gcd_int(CV, 32);
    return (kGnMaxThreads / CV / mult) * mult;
}

static GnPlan plan_nhwc(int N, int C, int HW) {
    GnPlan p;
    const int CV = C / 8;
    p.RY = gn_row_lanes(CV);
    p.threads = CV * p.RY;
    int want = (TMX_GN_MINB * sm_count() + N - 1) / N;       // one resident wave: TMX_GN_MINB CTAs per SM over the whole grid
    if (want < 1) want = 1;
    int rows = (HW + want - 1) / want;
    const int min_rows = 2 * (kGnU > kGnUS ? kGnU : kGnUS) * p.RY;                    // >= two pipelined batches of rows per thread when possible
    if (rows < min_rows) rows = min_rows;
    rows = ((rows + p.RY - 1) / p.RY) * p.RY;
    if (rows > HW) rows = HW;
    p.rows_per_cta = rows;
    p.P = (HW + rows - 1) / rows;
    if (p.P > kGnMaxParts) { p.rows_per_cta = (HW + kGnMaxParts - 1) / kGnMaxParts; p.P = (HW + p.rows_per_cta - 1) / p.rows_per_cta; }
    return p;
}

// Fused path: P = floor(#SM / N) CTAs per batch row (one CTA per SM, all co-resident), slab + scratch in shared memory.
constexpr size_t kGnFusedSmemMax = 200 * 1024;
struct GnFusedPlan { bool ok; int threads, rows_per_cta, P; size_t smem; };
static int g_gn_force_two_pass = 0;      // test hook
static GnFusedPlan plan_fused(int N, int C, int HW) {
    GnFusedPlan p{false, 0, 0, 0, 0};
    if (N > 1024 || N > sm_count()) return p;
    const int CV = C / 8;
    const int RY = gn_row_lanes(CV);
    if (RY == 0) return p;
    int P = sm_count() / N;
    if (P > kGnMaxParts) P = kGnMaxParts;
    int rows = (HW + P - 1) / P;
    P = (HW + rows - 1) / rows;
    p.threads = CV * RY;
    p.rows_per_cta = rows;
    p.P = P;
    p.smem = (size_t)rows * C * 2 + ((size_t)RY * C + C) * sizeof(float2);
    p.ok = p.smem <= kGnFusedSmemMax;
    return p;
}

// One cluster of CS CTAs per (sample, group): each CTA's share of the slab (cpg * ceil(HW / CS) 16-bit elements) must fit in shared
// memory, the group's pixel run must be a whole number of 4-, 8- or 16-byte vectors, and there must be enough (n, g) pairs to
// occupy the device.
constexpr size_t kGnSlabSmemMax = 208 * 1024;
struct GnSlabPlan { bool ok; int vl, threads, cs, ppc; size_t smem; };
static int g_gn_slab = 1;                 // test hook (tmx_groupnorm_set_variant 3 / 4 / 5): 0 = never, 1 = yes, 2 = single CTAs only (no clusters)
static GnSlabPlan plan_slab(int N, int C, int HW, int G) {
    GnSlabPlan p{false, 0, 0, 0, 0, 0};
    const int cpg = C / G;
    if (cpg % 2 != 0 || N > 65535 || (long long)N * G < 24) return p;
    p.vl = cpg % 8 == 0 ? 8 : (cpg % 4 == 0 ? 4 : 2);
    const int vpp = cpg / p.vl;
    const int unit = vpp / gcd_int(vpp, 32) * 32;             // lcm(vpp, 32)
    if (unit > 1024) return p;
    for (int cs = 1; cs <= 8; cs *= 2) {
        const int ppc = (HW + cs - 1) / cs;
        const size_t smem = (size_t)cpg * ppc * 2;
        if (smem > kGnSlabSmemMax) continue;
        // clusters only with 16-byte vectors: measured (profiles/r02r_kbench_groupnorm_slab.txt) the 128x128 sites (cpg = 10 / 20 / 30:
        // 4- and 8-byte runs per pixel) are SLOWER per group than the two-launch path with its full-row 16-byte accesses
        if (cs > 1 && (g_gn_slab == 2 || p.vl != 8 || (long long)(cs - 1) * ppc >= HW)) break;
        p.cs = cs; p.ppc = ppc; p.smem = smem;
        p.threads = 1024 / unit * unit;
        const long long vecs = (long long)ppc * vpp;
        while (p.threads - unit >= 256 && (long long)(p.threads - unit) * 4 >= vecs) p.threads -= unit;    // small slabs: >= 4 vectors per thread
        p.ok = true;
        break;
    }
    return p;
}

template <typename T, int VL>
static int launch_slab(const GnSlabPlan& sp, const void* x, const void* x2, int C1, const float* gamma, const float* beta, const float* add,
                       void* y, int N, int C, int HW, int G, float eps, int act, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G * sp.cs, N);
    cfg.blockDim = dim3(sp.threads);
    cfg.dynamicSmemBytes = sp.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = sp.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = sp.cs > 1 ? 1 : 0;
    return check_cuda(cudaLaunchKernelEx(&cfg, gn_group_slab<T, VL>, (const T*)x, (const T*)x2, C1, gamma, beta, add, (T*)y, C, HW, C / G, sp.cs, sp.ppc, eps, act),
                      "gn_group_slab launch");
}

// workspace layout: [tickets + generations: 2048 x u32][stat: N*G x float2][pivots (NCHW path): N*G x f32][part: N*G*kGnMaxParts x float2]
constexpr size_t kGnTicketBytes = 2048 * sizeof(unsigned int);   // [0,1024) arrival counters / tickets (left at 0), [1024,2048) fused-path barrier generations
static size_t gn_align(size_t v) { return (v + 15) & ~(size_t)15; }

template <typename T>
static int run_gn(const void* x, const float* gamma, const float* beta, const float* add, void* y, void* ws,
                  int N, int C, int HW, int G, float eps, int act, int layout, cudaStream_t st,
                  const void* x2v = nullptr, int C1 = 0) {
    unsigned int* tickets = (unsigned int*)ws;
    float2* stat = (float2*)((char*)ws + kGnTicketBytes);
    float* pivots = (float*)((char*)stat + gn_align((size_t)N * G * sizeof(float2)));
    float2* part = (float2*)((char*)pivots + gn_align((size_t)N * G * sizeof(float)));
    if (layout == TMX_NHWC) {
        if constexpr (sizeof(T) == 2) {
            GnSlabPlan sp = plan_slab(N, C, HW, G);
            if (sp.ok && g_gn_slab && g_gn_force_two_pass == 0) {
                if (sp.vl == 8) return launch_slab<T, 8>(sp, x, x2v, C1, gamma, beta, add, y, N, C, HW, G, eps, act, st);
                if (sp.vl == 4) return launch_slab<T, 4>(sp, x, x2v, C1, gamma, beta, add, y, N, C, HW, G, eps, act, st);
                return launch_slab<T, 2>(sp, x, x2v, C1, gamma, beta, add, y, N, C, HW, G, eps, act, st);
            }
            GnFusedPlan fp = plan_fused(N, C, HW);
            if (fp.ok && g_gn_force_two_pass != 1) {
                const T* xx = (const T*)x; T* yy = (T*)y;
                const T* xx2 = (const T*)x2v;
                unsigned int* sync = tickets;
                void* args[] = {&xx, &xx2, &C1, &gamma, &beta, &add, &part, &stat, &sync, &yy, &C, &HW, &G, &fp.rows_per_cta, &eps, &act};
                if (g_gn_force_two_pass == 2) {
                    gn_fused_nhwc<T><<<dim3(fp.P, N), fp.threads, fp.smem, st>>>(xx, xx2, C1, gamma, beta, add, part, stat, sync, yy, C, HW, G, fp.rows_per_cta, eps, act);
                    return check_cuda(cudaGetLastError(), "gn_fused_nhwc launch");
                }
                cudaError_t e = cudaLaunchCooperativeKernel((void*)gn_fused_nhwc<T>, dim3(fp.P, N), dim3(fp.threads), args, fp.smem, st);
                if (e == cudaSuccess) return TMX_OK;
                if (e != cudaErrorCooperativeLaunchTooLarge) return check_cuda(e, "gn_fused_nhwc launch");
                (void)cudaGetLastError();                                  // the device cannot co-schedule the grid: take the two-launch path
            }
        }
        GnPlan p = plan_nhwc(N, C, HW);
        dim3 grid(p.P, N);
        size_t smem = ((size_t)p.RY * C + C) * sizeof(float2);
        gn_stats_nhwc<T><<<grid, p.threads, smem, st>>>((const T*)x, (const T*)x2v, C1, add, part, stat, tickets, C, HW, G, p.rows_per_cta, eps);
        TMX_CUDA(cudaGetLastError());
        gn_apply_nhwc<T><<<grid, p.threads, 0, st>>>((const T*)x, (const T*)x2v, C1, gamma, beta, add, stat, (T*)y, C, HW, G, p.rows_per_cta, act);
        return check_cuda(cudaGetLastError(), "gn_apply_nhwc launch");
    }
    const int cpg = C / G, nvec = cpg * HW / 8;
    int S = (4 * sm_count() + N * G - 1) / (N * G);
    int vec_per = (nvec + S - 1) / S;
    if (vec_per < 1024) vec_per = 1024;                       // >= 4 vectors per thread
    S = (nvec + vec_per - 1) / vec_per;
    dim3 grid(S, N * G);
    gn_stats_nchw<T><<<grid, 256, 0, st>>>((const T*)x, add, part, pivots, C, HW, G, vec_per);
    TMX_CUDA(cudaGetLastError());
    gn_apply_nchw<T><<<grid, 256, 0, st>>>((const T*)x, gamma, beta, add, part, pivots, (T*)y, C, HW, G, vec_per, S, eps, act);
    return check_cuda(cudaGetLastError(), "gn_apply_nchw launch");
}

int groupnorm_init() {
    // TMX_GN_TWO_PASS=1 forces the two-launch path (ncu's kernel replay cannot re-run the fused kernel's inter-CTA barrier)
    if (const char* e = getenv("TMX_GN_TWO_PASS")) g_gn_force_two_pass = (e[0] == '1') ? 1 : g_gn_force_two_pass;
    // stats kernel: (RY*C + C) float2 of dynamic smem = 32 KB + 8*C bytes: above the 48 KB default for C > 2048.
    TMX_CUDA(cudaFuncSetAttribute(gn_stats_nhwc<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    TMX_CUDA(cudaFuncSetAttribute(gn_stats_nhwc<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    TMX_CUDA(cudaFuncSetAttribute(gn_stats_nhwc<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    TMX_CUDA(cudaFuncSetAttribute(gn_group_slab<__half, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGnSlabSmemMax));
    TMX_CUDA(cudaFuncSetAttribute(gn_group_slab<__half, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGnSlabSmemMax));
    TMX_CUDA(cudaFuncSetAttribute(gn_group_slab<__nv_bfloat16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGnSlabSmemMax));
    TMX_CUDA(cudaFuncSetAttribute(gn_group_slab<__nv_bfloat16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGnSlabSmemMax));
    TMX_CUDA(cudaFuncSetAttribute(gn_group_slab<__half, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGnSlabSmemMax));
    TMX_CUDA(cudaFuncSetAttribute(gn_group_slab<__nv_bfloat16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGnSlabSmemMax));
    TMX_CUDA(cudaFuncSetAttribute(gn_fused_nhwc<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGnFusedSmemMax));
    TMX_CUDA(cudaFuncSetAttribute(gn_fused_nhwc<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGnFusedSmemMax));
    return TMX_OK;
}

}  // namespace tmx

using namespace tmx;

extern "C" int tmx_groupnorm_set_variant(int v) {
    // 0: defaults (per-group slab kernel when the (n, g) slab fits in shared memory, else the cooperative fused kernel when the
    // activation fits in one wave of shared memory, else two launches); 1: always two launches; 2: fused kernel by plain launch;
    // 3: per-group slab kernel off (fused / two-launch as before); 4: on again; 5: on, single CTAs only (no clusters)
    TMX_REQUIRE(v >= 0 && v <= 5, TMX_EINVAL, "groupnorm_set_variant: 0 .. 5");
    if (v >= 3) { g_gn_slab = v == 3 ? 0 : (v == 4 ? 1 : 2); return TMX_OK; }
    g_gn_force_two_pass = v;
    if (v == 0) g_gn_slab = 1;
    return TMX_OK;
}

extern "C" int tmx_groupnorm_launches(int N, int C, int HW, int layout, int dtype) {
    // kernels tmx_groupnorm_fwd launches for this shape: 1 = fused cooperative kernel, 2 = stats + apply
    if (layout == TMX_NHWC && (dtype == TMX_F16 || dtype == TMX_BF16) && g_gn_force_two_pass == 0 && g_gn_slab &&
        N > 0 && C > 0 && HW > 0 && C % 32 == 0 && plan_slab(N, C, HW, 32).ok) return 1;
    if (layout == TMX_NHWC && (dtype == TMX_F16 || dtype == TMX_BF16) && g_gn_force_two_pass != 1 &&
        N > 0 && C > 0 && HW > 0 && C % 8 == 0 && plan_fused(N, C, HW).ok) return 1;
    return 2;
}

extern "C" size_t tmx_groupnorm_workspace_bytes(int N, int C, int HW, int G, int layout) {
    (void)C; (void)HW; (void)layout;
    if (N <= 0 || G <= 0) return 0;
    return kGnTicketBytes + gn_align((size_t)N * G * sizeof(float2)) + gn_align((size_t)N * G * sizeof(float))
         + (size_t)N * G * kGnMaxParts * sizeof(float2);
}

extern "C" int tmx_groupnorm_fwd(const void* x, const float* gamma, const float* beta, const float* add,
                                 void* y, void* workspace, int N, int C, int HW, int G, float eps,
                                 int act, int layout, int dtype, void* stream) {
    TMX_REQUIRE(x && gamma && beta && y && workspace, TMX_EINVAL, "groupnorm: null pointer");
    TMX_REQUIRE(N > 0 && C > 0 && HW > 0 && G > 0, TMX_EINVAL, "groupnorm: non-positive size");
    TMX_REQUIRE(N <= 1024, TMX_ESHAPE, "groupnorm: N=%d exceeds 1024 rows per launch", N);
    TMX_REQUIRE(G <= 64 && C % G == 0, TMX_ESHAPE, "groupnorm: C=%d not divisible by G=%d (G<=64)", C, G);
    TMX_REQUIRE(act == TMX_ACT_NONE || act == TMX_ACT_SILU, TMX_EINVAL, "groupnorm: bad act %d", act);
    TMX_REQUIRE(aligned16(x) && aligned16(y) && aligned16(workspace), TMX_EALIGN, "groupnorm: 16-byte alignment");
    if (layout == TMX_NHWC) {
        TMX_REQUIRE(C % 8 == 0 && C / 8 <= kGnMaxThreads, TMX_ESHAPE, "groupnorm NHWC: C=%d must be a multiple of 8 and <= %d", C, 8 * kGnMaxThreads);
        TMX_REQUIRE(gn_row_lanes(C / 8) > 0, TMX_ESHAPE, "groupnorm NHWC: no warp-aligned block for C=%d", C);
    } else if (layout == TMX_NCHW) {
        TMX_REQUIRE(HW % 8 == 0, TMX_ESHAPE, "groupnorm NCHW: HW=%d must be a multiple of 8", HW);
    } else {
        set_error("groupnorm: bad layout %d", layout);
        return TMX_EINVAL;
    }
    if (int rc = require_init()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case TMX_F32:  return run_gn<float>(x, gamma, beta, add, y, workspace, N, C, HW, G, eps, act, layout, st);
        case TMX_F16:  return run_gn<__half>(x, gamma, beta, add, y, workspace, N, C, HW, G, eps, act, layout, st);
        case TMX_BF16: return run_gn<__nv_bfloat16>(x, gamma, beta, add, y, workspace, N, C, HW, G, eps, act, layout, st);
    }
    set_error("groupnorm: unsupported dtype %d", dtype);
    return TMX_EDTYPE;
}

extern "C" int tmx_groupnorm_cat_fwd(const void* x1, const void* x2, int C1, const float* gamma, const float* beta, const float* add,
                                     void* y, void* workspace, int N, int C, int HW, int G, float eps, int act, int dtype, void* stream) {
    TMX_REQUIRE(x1 && x2 && gamma && beta && y && workspace, TMX_EINVAL, "groupnorm_cat: null pointer");
    TMX_REQUIRE(N > 0 && C > 0 && HW > 0 && G > 0 && N <= 1024, TMX_EINVAL, "groupnorm_cat: bad size");
    TMX_REQUIRE(C1 > 0 && C1 < C && C1 % 8 == 0 && C % 8 == 0, TMX_ESHAPE, "groupnorm_cat: C1=%d must be a multiple of 8 inside (0, C=%d)", C1, C);
    TMX_REQUIRE(G <= 64 && C % G == 0, TMX_ESHAPE, "groupnorm_cat: C=%d not divisible by G=%d (G<=64)", C, G);
    TMX_REQUIRE(C / 8 <= kGnMaxThreads && gn_row_lanes(C / 8) > 0, TMX_ESHAPE, "groupnorm_cat: no warp-aligned block for C=%d", C);
    TMX_REQUIRE(act == TMX_ACT_NONE || act == TMX_ACT_SILU, TMX_EINVAL, "groupnorm_cat: bad act %d", act);
    TMX_REQUIRE(dtype == TMX_F16 || dtype == TMX_BF16, TMX_EDTYPE, "groupnorm_cat: dtype %d unsupported (fp16/bf16 NHWC only)", dtype);
    TMX_REQUIRE(aligned16(x1) && aligned16(x2) && aligned16(y) && aligned16(workspace), TMX_EALIGN, "groupnorm_cat: 16-byte alignment");
    if (int rc = require_init()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TMX_F16) return run_gn<__half>(x1, gamma, beta, add, y, workspace, N, C, HW, G, eps, act, TMX_NHWC, st, x2, C1);
    return run_gn<__nv_bfloat16>(x1, gamma, beta, add, y, workspace, N, C, HW, G, eps, act, TMX_NHWC, st, x2, C1);
}
