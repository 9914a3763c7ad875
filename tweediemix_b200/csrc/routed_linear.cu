// k3 — per-row routed projection: one weight matrix (and / or one set of rank-r LoRA factors) per batch row.
//
//   y[b] = x[b] @ W[b]^T                                   (grouped GEMM, tcgen05 / TMEM / TMA)
//   y[b] += segment-wise (x[b] @ down[b]^T) @ up[b]^T      (rank-r deltas, bandwidth-bound CUDA-core kernel)
//
// Replaces the per-row nn.Linear + torch.cat of fusion_generation/utils_custom.py:64-82 (row i+1 of the K+1 batch goes
// through concept i's to_k / to_v) and the rank-4 deltas of utils_lora.py:65-79,113-119 (LoRALinearLayer, model_lora.py:28-48)
// on q, k, v and the output projection.
//
// Grouped GEMM: CTA = (128 output columns, 128 rows of x[b], batch row b).  The A tile (x) and the B tile (W[b], which
// is [Nout, Kin] row-major = K-major, exactly the layout of the K operand of S = Q K^T) are staged by TMA with
// SWIZZLE_128B through a 4-stage ring; one elected thread issues tcgen05.mma (kind::f16, M = N = 128, K = 16) into a
// 128-column TMEM accumulator; four epilogue warps read it back (thread == row), round once and store.  Rows beyond M
// and columns beyond Nout are zero-filled by TMA on the way in and masked on the way out.  The text-only cross-attention
// K/V of the custom variant (M = 77) is the intended shape: one launch instead of K+1 cuBLAS calls, W[b] streamed once.
//
// LoRA delta: one warp per token row, 3 rows per warp, 8 warps per CTA; down[b] / up[b] converted to fp32 in shared
// memory once per CTA; t = x_row . down^T by lane-strided 128-bit loads + warp reduction, then a read-modify-write of
// the y row.  Output column n belongs to segment s = n / (Nout / nseg) and uses t[s*r .. s*r + r): a packed q|k|v
// projection is ONE launch with nseg = 3 and no block-diagonal zero padding.
#include "tmx_common.cuh"
#include <cuda.h>

// The two-rows-per-warp LoRA kernel (lora_delta_kernel2) is the default; -DTMX_LORA_V1 selects the row-at-a-time one.
#if !defined(TMX_LORA_V1) && !defined(TMX_LORA_V2)
#define TMX_LORA_V2
#endif

namespace tmx {
namespace k3 {

constexpr int kMaxRows = 16;          // batch rows per launch
constexpr int kBM = 128, kBN = 128, kBK = 64;
constexpr int kTile = kBM * kBK * 2;  // 16 KiB
constexpr int kStages = 4;
constexpr int kGemmThreads = 192;     // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kGemmSmem = 1024 + kStages * 2 * kTile + (2 * kStages + 1) * 8 + 16;

__device__ unsigned int g_k3_timeout_flag = 0;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {      // bounded: a protocol bug traps, never hangs
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++polls & 255u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) { atomicExch(&g_k3_timeout_flag, 1u); __trap(); }
        }
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// SWIZZLE_128B K-major operand descriptors: low word = start address >> 4 | LBO (16 B) << 16; high word = SBO 1024 B,
// version 1, layout SWIZZLE_128B (same encoding as the attention kernel's Q / K operands).
constexpr uint32_t kDescHi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc(bool bf16, int M, int N) {
    return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
#define K3_R8(r, o)  "=r"(r[o+0]), "=r"(r[o+1]), "=r"(r[o+2]), "=r"(r[o+3]), "=r"(r[o+4]), "=r"(r[o+5]), "=r"(r[o+6]), "=r"(r[o+7])
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : K3_R8(r, 0), K3_R8(r, 8), K3_R8(r, 16), K3_R8(r, 24) : "r"(taddr) : "memory");
}

struct WeightMaps { CUtensorMap w[kMaxRows]; };

template <bool BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
routed_gemm_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ WeightMaps maps,
                   void* __restrict__ y, int M, int Kin, int Nout) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint32_t sA = smem_u32(smem);
    asm volatile("mov.u32 %0, %0;" : "+r"(sA));
    const uint32_t sB = sA + kStages * kTile;
    const uint32_t full = sB + kStages * kTile, empty = full + 8 * kStages, acc_full = empty + 8 * kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 2 * kStages * kTile + (2 * kStages + 1) * 8);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int nt = blockIdx.x, mt = blockIdx.y, b = blockIdx.z;
    const int KT = Kin / kBK;

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tm_x)) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&maps.w[b])) : "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        // ---- TMA producer
        int st = 0;
        uint32_t ph = 0;
        for (int kt = 0; kt < KT; ++kt) {
            mbar_wait(empty + 8 * st, ph ^ 1u);
            if (elect_one()) {
                mbar_expect_tx(full + 8 * st, 2 * kTile);
                tma_load_3d(sA + st * kTile, &tm_x, full + 8 * st, kt * kBK, mt * kBM, b);
                tma_load_2d(sB + st * kTile, &maps.w[b], full + 8 * st, kt * kBK, nt * kBN);
            }
            if (++st == kStages) { st = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        // ---- MMA issuer: acc[128 x 128] += x_tile[128 x 64] . W_tile[128 x 64]^T, four K = 16 steps per stage
        constexpr uint32_t idesc = make_idesc(BF16, kBM, kBN);
        int st = 0;
        uint32_t ph = 0;
        for (int kt = 0; kt < KT; ++kt) {
            mbar_wait(full + 8 * st, ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_lo = desc_lo(sA + st * kTile), b_lo = desc_lo(sB + st * kTile);
                umma_ss(tmem_base, a_lo, b_lo, idesc, kt > 0 ? 1u : 0u);
                umma_ss(tmem_base, a_lo + 2, b_lo + 2, idesc, 1u);
                umma_ss(tmem_base, a_lo + 4, b_lo + 4, idesc, 1u);
                umma_ss(tmem_base, a_lo + 6, b_lo + 6, idesc, 1u);
                umma_commit(empty + 8 * st);
                if (kt == KT - 1) umma_commit(acc_full);
            }
            if (++st == kStages) { st = 0; ph ^= 1u; }
        }
    } else {
        // ---- epilogue: thread == row of the tile == TMEM lane
        const int quarter = warp & 3;
        const int row = mt * kBM + quarter * 32 + lane;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        uint8_t* dst = reinterpret_cast<uint8_t*>(y) + (((size_t)b * M + (size_t)row) * (size_t)Nout + (size_t)nt * kBN) * 2;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            uint32_t acc[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + c4 * 32, acc);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < M) {
#pragma unroll
                for (int c = 0; c < 32; c += 8) {
                    if (nt * kBN + c4 * 32 + c + 8 <= Nout) {
                        uint4 v;
                        if constexpr (BF16) {
                            v.x = Pack2<__nv_bfloat16>::pack(__uint_as_float(acc[c]), __uint_as_float(acc[c + 1]));
                            v.y = Pack2<__nv_bfloat16>::pack(__uint_as_float(acc[c + 2]), __uint_as_float(acc[c + 3]));
                            v.z = Pack2<__nv_bfloat16>::pack(__uint_as_float(acc[c + 4]), __uint_as_float(acc[c + 5]));
                            v.w = Pack2<__nv_bfloat16>::pack(__uint_as_float(acc[c + 6]), __uint_as_float(acc[c + 7]));
                        } else {
                            v.x = Pack2<__half>::pack(__uint_as_float(acc[c]), __uint_as_float(acc[c + 1]));
                            v.y = Pack2<__half>::pack(__uint_as_float(acc[c + 2]), __uint_as_float(acc[c + 3]));
                            v.z = Pack2<__half>::pack(__uint_as_float(acc[c + 4]), __uint_as_float(acc[c + 5]));
                            v.w = Pack2<__half>::pack(__uint_as_float(acc[c + 6]), __uint_as_float(acc[c + 7]));
                        }
                        *reinterpret_cast<uint4*>(dst + (c4 * 32 + c) * 2) = v;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(128) : "memory");
    }
}

// ------------------------------------------------------------------------------------------- LoRA delta
struct LoraPtrs { const void* down[kMaxRows]; const void* up[kMaxRows]; };
constexpr int kLoraWarps = 8, kLoraRowsPerWarp = 3;   // 24 rows per CTA: the 3 x 1024 routed rows of a 32x32 site fill 128 CTAs = one wave
constexpr size_t kLoraSmemMax = 200 * 1024;

// y[b, m, n] += sum_j t[seg(n) * r + j] * up[b][n, j],   t[q] = sum_k x[b, m, k] * down[b][q, k]
// SR = nseg * r (4, 8, 12 or 16).  smem: down as fp32 [SR][Kin], up as fp32 TRANSPOSED [r][Nout] (conflict-free column reads).
template <typename T, int SR>
__global__ void __launch_bounds__(kLoraWarps * 32, 1)
lora_delta_kernel(const T* __restrict__ x, T* __restrict__ y, const __grid_constant__ LoraPtrs ptrs,
                  int M, int Kin, int Nout, int r, int seg_cols) {
    extern __shared__ float lora_smem[];
    const int b = blockIdx.y;
    const T* down = reinterpret_cast<const T*>(ptrs.down[b]);
    const T* up = reinterpret_cast<const T*>(ptrs.up[b]);
    if (down == nullptr) return;                               // this batch row is not routed (row 0: the unconditional row)
    float* s_down = lora_smem;                                 // [SR][Kin]
    float* s_up = lora_smem + (size_t)SR * Kin;                // [r][Nout]  (transposed: lanes read consecutive columns)
    __shared__ float s_t[kLoraWarps][16];                      // per-warp t, so the segment lookup is a (broadcast) smem read
    for (int i = threadIdx.x; i < SR * Kin / 8; i += blockDim.x) {
        float f[8];
        unpack8<T>(ld_keep(down + (size_t)i * 8), f);
        *reinterpret_cast<float4*>(s_down + (size_t)i * 8) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(s_down + (size_t)i * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
    for (int i = threadIdx.x; i < Nout * r / 8; i += blockDim.x) {
        float f[8];
        unpack8<T>(ld_keep(up + (size_t)i * 8), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = i * 8 + e;                         // element (n, j) of up[Nout][r]
            const int n = idx / r, j = idx - n * r;
            s_up[(size_t)j * Nout + n] = f[e];
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec_in = Kin >> 3, nvec_out = Nout >> 3;
    for (int rr = 0; rr < kLoraRowsPerWarp; ++rr) {
        const int m = (blockIdx.x * kLoraWarps + warp) * kLoraRowsPerWarp + rr;
        if (m >= M) break;                                     // warp-uniform
        const T* xr = x + ((size_t)b * M + m) * Kin;
        float t[SR];
#pragma unroll
        for (int q = 0; q < SR; ++q) t[q] = 0.f;
        for (int v = lane; v < nvec_in; v += 32) {
            float f[8];
            unpack8<T>(ld_stream(xr + (size_t)v * 8), f);
#pragma unroll
            for (int q = 0; q < SR; ++q) {
                const float4 d0 = *reinterpret_cast<const float4*>(s_down + (size_t)q * Kin + v * 8);
                const float4 d1 = *reinterpret_cast<const float4*>(s_down + (size_t)q * Kin + v * 8 + 4);
                t[q] = fmaf(f[0], d0.x, fmaf(f[1], d0.y, fmaf(f[2], d0.z, fmaf(f[3], d0.w,
                       fmaf(f[4], d1.x, fmaf(f[5], d1.y, fmaf(f[6], d1.z, fmaf(f[7], d1.w, t[q]))))))));
            }
        }
#pragma unroll
        for (int q = 0; q < SR; ++q) t[q] = warp_sum(t[q]);
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < SR; ++q) s_t[warp][q] = t[q];
        }
        __syncwarp();
        T* yr = y + ((size_t)b * M + m) * Nout;
        for (int v = lane; v < nvec_out; v += 32) {
            const float* ts = &s_t[warp][((v * 8) / seg_cols) * r];  // 8 consecutive columns never straddle a segment (seg_cols % 8 == 0)
            float f[8];
            unpack8<T>(ld_keep(yr + (size_t)v * 8), f);
            for (int j = 0; j < r; ++j) {
                const float tj = ts[j];
                const float4 u0 = *reinterpret_cast<const float4*>(s_up + (size_t)j * Nout + v * 8);
                const float4 u1 = *reinterpret_cast<const float4*>(s_up + (size_t)j * Nout + v * 8 + 4);
                f[0] = fmaf(tj, u0.x, f[0]); f[1] = fmaf(tj, u0.y, f[1]); f[2] = fmaf(tj, u0.z, f[2]); f[3] = fmaf(tj, u0.w, f[3]);
                f[4] = fmaf(tj, u1.x, f[4]); f[5] = fmaf(tj, u1.y, f[5]); f[6] = fmaf(tj, u1.z, f[6]); f[7] = fmaf(tj, u1.w, f[7]);
            }
            *reinterpret_cast<uint4*>(yr + (size_t)v * 8) = pack8<T>(f);
        }
    }
}

// Same contract, two token rows per warp in flight: every shared-memory factor load feeds two rows (half the LDS
// traffic of the row-at-a-time kernel) and the two dependent FMA chains interleave.
template <typename T, int SR>
__global__ void __launch_bounds__(kLoraWarps * 32, 1)
lora_delta_kernel2(const T* __restrict__ x, T* __restrict__ y, const __grid_constant__ LoraPtrs ptrs,
                   int M, int Kin, int Nout, int r, int seg_cols) {
    extern __shared__ float lora_smem[];
    const int b = blockIdx.y;
    const T* down = reinterpret_cast<const T*>(ptrs.down[b]);
    const T* up = reinterpret_cast<const T*>(ptrs.up[b]);
    if (down == nullptr) return;
    float* s_down = lora_smem;                                 // [SR][Kin]
    float* s_up = lora_smem + (size_t)SR * Kin;                // [r][Nout]
    __shared__ float s_t[kLoraWarps][2][16];
    for (int i = threadIdx.x; i < SR * Kin / 8; i += blockDim.x) {
        float f[8];
        unpack8<T>(ld_keep(down + (size_t)i * 8), f);
        *reinterpret_cast<float4*>(s_down + (size_t)i * 8) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(s_down + (size_t)i * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
    for (int i = threadIdx.x; i < Nout * r / 8; i += blockDim.x) {
        float f[8];
        unpack8<T>(ld_keep(up + (size_t)i * 8), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = i * 8 + e;
            const int n = idx / r, j = idx - n * r;
            s_up[(size_t)j * Nout + n] = f[e];
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec_in = Kin >> 3, nvec_out = Nout >> 3;
    const int row0 = (blockIdx.x * kLoraWarps + warp) * kLoraRowsPerWarp;
    for (int rr = 0; rr < kLoraRowsPerWarp; rr += 2) {
        const int m0 = row0 + rr, m1 = m0 + 1;
        if (m0 >= M) break;                                    // warp-uniform
        const bool has1 = (rr + 1 < kLoraRowsPerWarp) && (m1 < M);
        const T* x0 = x + ((size_t)b * M + m0) * Kin;
        const T* x1 = x + ((size_t)b * M + (has1 ? m1 : m0)) * Kin;
        float t0[SR], t1[SR];
#pragma unroll
        for (int q = 0; q < SR; ++q) { t0[q] = 0.f; t1[q] = 0.f; }
        for (int v = lane; v < nvec_in; v += 32) {
            float f0[8], f1[8];
            unpack8<T>(ld_stream(x0 + (size_t)v * 8), f0);
            unpack8<T>(ld_stream(x1 + (size_t)v * 8), f1);
#pragma unroll
            for (int q = 0; q < SR; ++q) {
                const float4 d0 = *reinterpret_cast<const float4*>(s_down + (size_t)q * Kin + v * 8);
                const float4 d1 = *reinterpret_cast<const float4*>(s_down + (size_t)q * Kin + v * 8 + 4);
                t0[q] = fmaf(f0[0], d0.x, fmaf(f0[1], d0.y, fmaf(f0[2], d0.z, fmaf(f0[3], d0.w,
                        fmaf(f0[4], d1.x, fmaf(f0[5], d1.y, fmaf(f0[6], d1.z, fmaf(f0[7], d1.w, t0[q]))))))));
                t1[q] = fmaf(f1[0], d0.x, fmaf(f1[1], d0.y, fmaf(f1[2], d0.z, fmaf(f1[3], d0.w,
                        fmaf(f1[4], d1.x, fmaf(f1[5], d1.y, fmaf(f1[6], d1.z, fmaf(f1[7], d1.w, t1[q]))))))));
            }
        }
#pragma unroll
        for (int q = 0; q < SR; ++q) { t0[q] = warp_sum(t0[q]); t1[q] = warp_sum(t1[q]); }
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < SR; ++q) { s_t[warp][0][q] = t0[q]; s_t[warp][1][q] = t1[q]; }
        }
        __syncwarp();
        T* y0 = y + ((size_t)b * M + m0) * Nout;
        T* y1 = y + ((size_t)b * M + (has1 ? m1 : m0)) * Nout;
        for (int v = lane; v < nvec_out; v += 32) {
            const int so = ((v * 8) / seg_cols) * r;
            float f0[8], f1[8];
            unpack8<T>(ld_keep(y0 + (size_t)v * 8), f0);
            unpack8<T>(ld_keep(y1 + (size_t)v * 8), f1);
            for (int j = 0; j < r; ++j) {
                const float a0 = s_t[warp][0][so + j], a1 = s_t[warp][1][so + j];
                const float4 u0 = *reinterpret_cast<const float4*>(s_up + (size_t)j * Nout + v * 8);
                const float4 u1 = *reinterpret_cast<const float4*>(s_up + (size_t)j * Nout + v * 8 + 4);
                f0[0] = fmaf(a0, u0.x, f0[0]); f0[1] = fmaf(a0, u0.y, f0[1]); f0[2] = fmaf(a0, u0.z, f0[2]); f0[3] = fmaf(a0, u0.w, f0[3]);
                f0[4] = fmaf(a0, u1.x, f0[4]); f0[5] = fmaf(a0, u1.y, f0[5]); f0[6] = fmaf(a0, u1.z, f0[6]); f0[7] = fmaf(a0, u1.w, f0[7]);
                f1[0] = fmaf(a1, u0.x, f1[0]); f1[1] = fmaf(a1, u0.y, f1[1]); f1[2] = fmaf(a1, u0.z, f1[2]); f1[3] = fmaf(a1, u0.w, f1[3]);
                f1[4] = fmaf(a1, u1.x, f1[4]); f1[5] = fmaf(a1, u1.y, f1[5]); f1[6] = fmaf(a1, u1.z, f1[6]); f1[7] = fmaf(a1, u1.w, f1[7]);
            }
            *reinterpret_cast<uint4*>(y0 + (size_t)v * 8) = pack8<T>(f0);
            if (has1) *reinterpret_cast<uint4*>(y1 + (size_t)v * 8) = pack8<T>(f1);
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, bool bf16) {
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank,
                          const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("routed_linear: cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return TMX_ECUDA; }
    return TMX_OK;
}

template <typename T>
static int launch_lora(const void* x, void* y, const LoraPtrs& ptrs, int B, int M, int Kin, int Nout, int r, int nseg, cudaStream_t st) {
    const int SR = nseg * r;
    const size_t smem = ((size_t)SR * Kin + (size_t)Nout * r) * sizeof(float);
    dim3 grid((M + kLoraWarps * kLoraRowsPerWarp - 1) / (kLoraWarps * kLoraRowsPerWarp), B);
    const int seg_cols = Nout / nseg;
    switch (SR) {
#ifdef TMX_LORA_V2
#define lora_delta_kernel lora_delta_kernel2
#endif
        case 4:  lora_delta_kernel<T, 4><<<grid, kLoraWarps * 32, smem, st>>>((const T*)x, (T*)y, ptrs, M, Kin, Nout, r, seg_cols); break;
        case 8:  lora_delta_kernel<T, 8><<<grid, kLoraWarps * 32, smem, st>>>((const T*)x, (T*)y, ptrs, M, Kin, Nout, r, seg_cols); break;
        case 12: lora_delta_kernel<T, 12><<<grid, kLoraWarps * 32, smem, st>>>((const T*)x, (T*)y, ptrs, M, Kin, Nout, r, seg_cols); break;
        case 16: lora_delta_kernel<T, 16><<<grid, kLoraWarps * 32, smem, st>>>((const T*)x, (T*)y, ptrs, M, Kin, Nout, r, seg_cols); break;
        default: set_error("routed_linear: nseg * r = %d unsupported (4, 8, 12 or 16)", SR); return TMX_ESHAPE;
    }
#ifdef TMX_LORA_V2
#undef lora_delta_kernel
#endif
    return check_cuda(cudaGetLastError(), "lora_delta_kernel launch");
}

}  // namespace k3

int routed_init() {
    using namespace k3;
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        TMX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        TMX_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, TMX_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
        g_encode = (EncodeTiledFn)fn;
    }
    TMX_CUDA(cudaFuncSetAttribute(routed_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    TMX_CUDA(cudaFuncSetAttribute(routed_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
#define K3_ATTR(T, SR) TMX_CUDA(cudaFuncSetAttribute(lora_delta_kernel<T, SR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLoraSmemMax)); \
                       TMX_CUDA(cudaFuncSetAttribute(lora_delta_kernel2<T, SR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLoraSmemMax))
    K3_ATTR(__half, 4); K3_ATTR(__half, 8); K3_ATTR(__half, 12); K3_ATTR(__half, 16);
    K3_ATTR(__nv_bfloat16, 4); K3_ATTR(__nv_bfloat16, 8); K3_ATTR(__nv_bfloat16, 12); K3_ATTR(__nv_bfloat16, 16);
#undef K3_ATTR
    return TMX_OK;
}

}  // namespace tmx

using namespace tmx;
using namespace tmx::k3;

extern "C" int tmx_routed_linear_fwd(const void* x, const void* const* w, const void* const* lora_down,
                                     const void* const* lora_up, void* y, int B, int M, int Kin, int Nout,
                                     int r, int nseg, int dtype, void* stream) {
    TMX_REQUIRE(x && y, TMX_EINVAL, "routed_linear: null pointer");
    TMX_REQUIRE(w || (lora_down && lora_up), TMX_EINVAL, "routed_linear: neither weights nor LoRA factors given");
    TMX_REQUIRE(B > 0 && B <= kMaxRows && M > 0 && Kin > 0 && Nout > 0, TMX_ESHAPE, "routed_linear: B=%d (<= %d), M=%d, Kin=%d, Nout=%d", B, kMaxRows, M, Kin, Nout);
    TMX_REQUIRE(dtype == TMX_F16 || dtype == TMX_BF16, TMX_EDTYPE, "routed_linear: dtype %d unsupported (fp16/bf16 only)", dtype);
    TMX_REQUIRE(Kin % 64 == 0 && Nout % 8 == 0, TMX_ESHAPE, "routed_linear: Kin=%d must be a multiple of 64 and Nout=%d of 8", Kin, Nout);
    TMX_REQUIRE(aligned16(x) && aligned16(y), TMX_EALIGN, "routed_linear: 16-byte alignment");
    if ((lora_down == nullptr) != (lora_up == nullptr)) { set_error("routed_linear: lora_down and lora_up must be given together"); return TMX_EINVAL; }
    if (int rc = require_init()) return rc;
    const bool bf16 = dtype == TMX_BF16;
    cudaStream_t st = (cudaStream_t)stream;

    if (w) {
        WeightMaps maps;
        for (int b = 0; b < B; ++b) {
            TMX_REQUIRE(w[b] && aligned16(w[b]), TMX_EINVAL, "routed_linear: w[%d] is null or misaligned", b);
            cuuint64_t dims[2] = {(cuuint64_t)Kin, (cuuint64_t)Nout};
            cuuint64_t strides[1] = {(cuuint64_t)Kin * 2};
            cuuint32_t box[2] = {kBK, kBN};
            if (int rc = encode(&maps.w[b], w[b], 2, dims, strides, box, bf16)) return rc;
        }
        for (int b = B; b < kMaxRows; ++b) maps.w[b] = maps.w[0];
        CUtensorMap mx;
        cuuint64_t dims[3] = {(cuuint64_t)Kin, (cuuint64_t)M, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)Kin * 2, (cuuint64_t)M * Kin * 2};
        cuuint32_t box[3] = {kBK, kBM, 1};
        if (int rc = encode(&mx, x, 3, dims, strides, box, bf16)) return rc;
        dim3 grid((Nout + kBN - 1) / kBN, (M + kBM - 1) / kBM, B);
        if (bf16) routed_gemm_kernel<true><<<grid, kGemmThreads, kGemmSmem, st>>>(mx, maps, y, M, Kin, Nout);
        else      routed_gemm_kernel<false><<<grid, kGemmThreads, kGemmSmem, st>>>(mx, maps, y, M, Kin, Nout);
        if (int rc = check_cuda(cudaGetLastError(), "routed_gemm_kernel launch")) return rc;
    }
    if (lora_down) {
        TMX_REQUIRE(r >= 1 && nseg >= 1 && Nout % nseg == 0 && (Nout / nseg) % 8 == 0 && (Nout * r) % 8 == 0, TMX_ESHAPE,
                    "routed_linear: r=%d nseg=%d do not tile Nout=%d", r, nseg, Nout);
        TMX_REQUIRE(((size_t)nseg * r * Kin + (size_t)Nout * r) * sizeof(float) <= kLoraSmemMax, TMX_ESHAPE,
                    "routed_linear: LoRA factors (%d x %d, %d x %d) do not fit in shared memory", nseg * r, Kin, Nout, r);
        LoraPtrs ptrs;
        bool any = false;
        for (int b = 0; b < kMaxRows; ++b) {
            ptrs.down[b] = b < B ? lora_down[b] : nullptr;
            ptrs.up[b] = b < B ? lora_up[b] : nullptr;
            if (b < B) {
                if ((ptrs.down[b] == nullptr) != (ptrs.up[b] == nullptr)) { set_error("routed_linear: row %d has only one LoRA factor", b); return TMX_EINVAL; }
                TMX_REQUIRE(aligned16(ptrs.down[b]) && aligned16(ptrs.up[b]), TMX_EALIGN, "routed_linear: LoRA factors must be 16-byte aligned");
                any |= ptrs.down[b] != nullptr;
            }
        }
        if (any) {
            if (bf16) return launch_lora<__nv_bfloat16>(x, y, ptrs, B, M, Kin, Nout, r, nseg, st);
            return launch_lora<__half>(x, y, ptrs, B, M, Kin, Nout, r, nseg, st);
        }
    }
    return TMX_OK;
}
