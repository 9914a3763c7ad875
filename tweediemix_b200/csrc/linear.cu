// k10 — dense projection with a fused epilogue on the per-step path:  y = epilogue(x @ W^T)
//
//   x [M, K] 16-bit row-major (row stride ldx), W [N, K] 16-bit row-major (an nn.Linear weight), fp32 accumulation
//   epilogue NONE  : y[M, N]   = acc + bias (+ residual)
//   epilogue GEGLU : y[M, N/2] = (acc_v + bias_v) * gelu(acc_g + bias_g)      (W / bias rows interleaved 32 value | 32 gate)
//   LoRA tail      : acc += t[M, 16] @ up[b][N, 16]^T for the batch row b the tile belongs to (rank-r deltas of
//                    utils_lora.py:65-79,113-119 folded into the GEMM as ONE extra K = 16 MMA per tile)
//
// Replaces, in the reference's hooked attention forward, to_q / to_k / to_v (utils_custom.py:63-89, utils_lora.py:65-79) and
// to_out[0] (+ the residual add diffusers does right after; utils_custom.py:106, utils_lora.py:113-121), and in [D]
// BasicTransformerBlock the GEGLU feed-forward (proj -> x * gelu(gate) -> Linear -> + residual).
//
// Blackwell-native structure:
//   * PERSISTENT: one 320-thread CTA per SM walks the tiles (128 rows x BN columns, BN = 128 or 256) in a static round-robin
//     order with the row-tile index fastest, so the CTAs of a wave share one W panel out of L2;
//   * warp-specialised: warp 0 = TMA producer (cp.async.bulk.tensor 2-D, SWIZZLE_128B) through a 4- or 5-stage ring of
//     (x tile, W tile) 64-deep K blocks; warp 1 = MMA issuer (one elected thread, tcgen05.mma kind::f16, M = 128, N = BN,
//     K = 16) into one of TWO TMEM accumulator stages, so the tensor pipe starts tile i+1 while tile i is still being
//     drained; warps 2-9 = epilogue (thread == row == TMEM lane; two warps per lane quarter split the columns);
//   * the epilogue works on 128-column halves of the accumulator: tcgen05.ld -> + bias (+ residual, prefetched into shared
//     memory by TMA while the main loop still runs) (GEGLU) -> one rounding -> swizzled st.shared -> TMA store, so every
//     global access of the kernel is a full-line bulk copy and M / N tails are clipped by the tensor maps;
//   * mbarriers only (full/empty ring, accumulator full/empty, residual landed) plus one named barrier among the epilogue
//     warps around the shared staging tile.
#include "tmx_common.cuh"
#include <cuda.h>
#include <type_traits>

namespace tmx {
namespace k10 {

constexpr int kBM = 128, kBK = 64;
constexpr int kATile = kBM * kBK * 2;            // 16 KiB
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);   // 320
constexpr int kSlab = kBM * 128;                 // one 64-column (128 B) slab of the staging tile: 16 KiB
constexpr int kMaxBatchRows = 16;

template <int BN> struct Cfg {
    static constexpr int kBTile = BN * kBK * 2;                  // 16 / 32 KiB
    static constexpr int kStages = BN == 256 ? 4 : 5;
    static constexpr int kRing = kStages * (kATile + kBTile);    // 192 / 160 KiB
    static constexpr int kStage = 2 * kSlab;                     // staging tile: 128 rows x 128 columns, 32 KiB
    static constexpr int kBars = 2 * kStages + 4 + 1;
    static constexpr int kSmem = 1024 + kRing + kStage + kBars * 8 + 16;
    static constexpr int kTmemCols = 2 * BN;                     // two accumulator stages
    static constexpr int kHalves = BN / 128;
};

__device__ unsigned int g_k10_timeout_flag = 0;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
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
            else if (now - t0 > 4000000000LL) { atomicExch(&g_k10_timeout_flag, 1u); __trap(); }
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
        :: "l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" :: "n"(kEpiWarps * 32) : "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// SWIZZLE_128B K-major operand descriptors (same encoding as the attention / routed kernels): low word = start address >> 4
// | LBO (16 B) << 16; high word = SBO 1024 B, version 1, layout SWIZZLE_128B.
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
#define K10_R8(r, o)  "=r"(r[o+0]), "=r"(r[o+1]), "=r"(r[o+2]), "=r"(r[o+3]), "=r"(r[o+4]), "=r"(r[o+5]), "=r"(r[o+6]), "=r"(r[o+7])
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : K10_R8(r, 0), K10_R8(r, 8), K10_R8(r, 16), K10_R8(r, 24) : "r"(taddr) : "memory");
}

__device__ __forceinline__ float gelu_erf(float g) { return 0.5f * g * (1.f + erff(g * 0.70710678118654752f)); }

struct UpMaps { CUtensorMap m[kMaxBatchRows]; };

struct Params {
    int M, N, K;               // problem (N = accumulator columns = rows of W)
    int MT, NT;                // tile counts
    int epilogue;              // TMX_EPI_*
    int has_res, has_tail;
    int rows_per_batch;        // LoRA tail: rows of x per batch row (multiple of 128)
    unsigned tail_mask;        // bit b set = batch row b carries LoRA factors
    const float* bias;         // [N] fp32 or null
};

// advance (stage, phase) of the K ring
#define K10_NEXT(st, ph, NST) do { if (++(st) == (NST)) { (st) = 0; (ph) ^= 1u; } } while (0)

template <bool BF16, int BN>
__global__ void __launch_bounds__(kThreads, 1)
linear_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
              const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_r,
              const __grid_constant__ CUtensorMap tm_t, const __grid_constant__ UpMaps tm_up, const Params p) {
    using C = Cfg<BN>;
    constexpr int NST = C::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint32_t sA = smem_u32(smem);
    asm volatile("mov.u32 %0, %0;" : "+r"(sA));                        // keep the base in a register (no S2UR chain per use)
    const uint32_t sB = sA + NST * kATile;
    const uint32_t sC = sB + NST * C::kBTile;                          // staging tile (1024-aligned: all tiles are multiples of 1 KiB)
    const uint32_t full = sC + C::kStage, empty = full + 8 * NST;
    const uint32_t acc_full = empty + 8 * NST, acc_empty = acc_full + 16, res_full = acc_empty + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::kRing + C::kStage + C::kBars * 8);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int KT = p.K / kBK;
    const int total = p.MT * p.NT;

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, kEpiWarps); }
        mbar_init(res_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tm_x)) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tm_w)) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tm_y)) : "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(C::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        // =========================================== TMA producer ===========================================
        int st = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const int nt = t / p.MT, mt = t - nt * p.MT;
            for (int kt = 0; kt < KT; ++kt) {
                mbar_wait(empty + 8 * st, ph ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(full + 8 * st, kATile + C::kBTile);
                    tma_load_2d(sA + st * kATile, &tm_x, full + 8 * st, kt * kBK, mt * kBM);
                    tma_load_2d(sB + st * C::kBTile, &tm_w, full + 8 * st, kt * kBK, nt * BN);
                }
                K10_NEXT(st, ph, NST);
            }
            if (p.has_tail) {
                const int b = (mt * kBM) / p.rows_per_batch;
                if ((p.tail_mask >> b) & 1u) {
                    mbar_wait(empty + 8 * st, ph ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(full + 8 * st, kATile + C::kBTile);
                        tma_load_2d(sA + st * kATile, &tm_t, full + 8 * st, 0, mt * kBM);
                        tma_load_2d(sB + st * C::kBTile, &tm_up.m[b], full + 8 * st, 0, nt * BN);
                    }
                    K10_NEXT(st, ph, NST);
                }
            }
        }
    } else if (warp == 1) {
        // =========================================== MMA issuer =============================================
        constexpr uint32_t idesc = make_idesc(BF16, kBM, BN);
        int st = 0;
        uint32_t ph = 0, it = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
            const uint32_t a = it & 1u;
            const int mt = t % p.MT;
            mbar_wait(acc_empty + 8 * a, ((it >> 1) & 1u) ^ 1u);        // the epilogue has drained this accumulator stage
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + a * BN;
            for (int kt = 0; kt < KT; ++kt) {
                mbar_wait(full + 8 * st, ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_lo = desc_lo(sA + st * kATile), b_lo = desc_lo(sB + st * C::kBTile);
                    umma_ss(d_tmem, a_lo, b_lo, idesc, kt > 0 ? 1u : 0u);
                    umma_ss(d_tmem, a_lo + 2, b_lo + 2, idesc, 1u);
                    umma_ss(d_tmem, a_lo + 4, b_lo + 4, idesc, 1u);
                    umma_ss(d_tmem, a_lo + 6, b_lo + 6, idesc, 1u);
                    umma_commit(empty + 8 * st);
                }
                K10_NEXT(st, ph, NST);
            }
            if (p.has_tail) {
                const int b = (mt * kBM) / p.rows_per_batch;
                if ((p.tail_mask >> b) & 1u) {                          // ONE K = 16 step: t[128 x 16] . up[b][BN x 16]^T
                    mbar_wait(full + 8 * st, ph);
                    tc_fence_after();
                    if (elect_one()) {
                        umma_ss(d_tmem, desc_lo(sA + st * kATile), desc_lo(sB + st * C::kBTile), idesc, 1u);
                        umma_commit(empty + 8 * st);
                    }
                    K10_NEXT(st, ph, NST);
                }
            }
            if (elect_one()) umma_commit(acc_full + 8 * a);
        }
    } else {
        // =========================================== epilogue ===============================================
        using T = typename std::conditional<BF16, __nv_bfloat16, __half>::type;
        const int ew = warp - 2;                                         // 0..7
        const int quarter = warp & 3;                                    // TMEM lane quarter this warp may access
        const int cg = ew >> 2;                                          // 64-column group of the 128-column half
        const int row = quarter * 32 + lane;                             // row in the tile == TMEM lane
        const bool leader = (ew == 0 && lane == 0);
        const bool geglu = p.epilogue == TMX_EPI_GEGLU;
        const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // this thread's 128-byte row of its slab, 16-byte chunks XOR-swizzled with the row (SWIZZLE_128B)
        const uint32_t c_row = sC + (geglu ? 0 : cg * kSlab) + row * 128;
        const uint32_t sw = (uint32_t)(row & 7);
        uint32_t it = 0, nres = 0;
        bool first = true;
        if (p.has_res && leader && (int)blockIdx.x < total) {            // residual of the first half
            const int nt = blockIdx.x / p.MT, mt = blockIdx.x - nt * p.MT;
            mbar_expect_tx(res_full, 2 * kSlab);
            tma_load_2d(sC, &tm_r, res_full, nt * BN, mt * kBM);
            tma_load_2d(sC + kSlab, &tm_r, res_full, nt * BN + 64, mt * kBM);
        }
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
            const uint32_t a = it & 1u;
            const int nt = t / p.MT, mt = t - nt * p.MT;
            mbar_wait(acc_full + 8 * a, (it >> 1) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int h = 0; h < C::kHalves; ++h) {
                uint32_t acc[64];
                const uint32_t t_col = t_row + a * BN + h * 128 + cg * 64;
                tmem_ld32(t_col, acc);
                tmem_ld32(t_col + 32, acc + 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (h == C::kHalves - 1) {                               // accumulator stage fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + 8 * a);
                }
                const int n_acc = nt * BN + h * 128 + cg * 64;           // first accumulator column of this thread's 64
                if (!first) epi_bar();                                   // the leader has drained the previous TMA store (and issued the residual load)
                first = false;
                if (p.has_res) { mbar_wait(res_full, nres & 1u); ++nres; }
                if (!geglu) {
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8) {                     // 8 columns = one 16-byte chunk
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(acc[c8 * 8 + e]);
                        const int n = n_acc + c8 * 8;
                        if (p.bias != nullptr && n < p.N) {
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
                            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                        }
                        const uint32_t addr = c_row + (((uint32_t)c8 ^ sw) << 4);
                        if (p.has_res) {
                            uint4 rv;
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rv.x), "=r"(rv.y), "=r"(rv.z), "=r"(rv.w) : "r"(addr));
                            float r8[8];
                            unpack8<T>(rv, r8);
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] += r8[e];
                        }
                        const uint4 v = pack8<T>(f);
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                    }
                } else {
                    // 64 accumulator columns = 32 value | 32 gate  ->  32 outputs at output column (n_acc / 2)
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        float f[8];
                        const int n = n_acc + c8 * 8;
                        float bv[8] = {0, 0, 0, 0, 0, 0, 0, 0}, bg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                        if (p.bias != nullptr && n < p.N) {
                            const float4 v0 = __ldg(reinterpret_cast<const float4*>(p.bias + n)), v1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
                            const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 32)), g1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 36));
                            bv[0] = v0.x; bv[1] = v0.y; bv[2] = v0.z; bv[3] = v0.w; bv[4] = v1.x; bv[5] = v1.y; bv[6] = v1.z; bv[7] = v1.w;
                            bg[0] = g0.x; bg[1] = g0.y; bg[2] = g0.z; bg[3] = g0.w; bg[4] = g1.x; bg[5] = g1.y; bg[6] = g1.z; bg[7] = g1.w;
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float v = __uint_as_float(acc[c8 * 8 + e]) + bv[e];
                            const float g = __uint_as_float(acc[32 + c8 * 8 + e]) + bg[e];
                            // the un-fused path rounds the projection to 16 bits before the gating kernel reads it
                            f[e] = Pack2<T>::round(v) * gelu_erf(Pack2<T>::round(g));
                        }
                        const uint4 v = pack8<T>(f);
                        const uint32_t addr = c_row + (((uint32_t)(cg * 4 + c8) ^ sw) << 4);
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                    }
                }
                fence_async_smem();                                      // generic-proxy writes -> visible to the TMA store
                epi_bar();
                if (leader) {
                    const int n0 = nt * BN + h * 128;                    // (a half that starts beyond N holds only zero padding)
                    if (!geglu) {
                        if (n0 < p.N) tma_store_2d(&tm_y, sC, n0, mt * kBM);
                        if (n0 + 64 < p.N) tma_store_2d(&tm_y, sC + kSlab, n0 + 64, mt * kBM);
                    } else if (n0 < p.N) {
                        tma_store_2d(&tm_y, sC, n0 >> 1, mt * kBM);
                    }
                    tma_store_commit();
                    tma_store_wait_read();                               // the staging tile may be overwritten
                    if (p.has_res) {                                     // prefetch the residual of the NEXT half (it lands while the main loop of that tile runs)
                        int t2 = t, h2 = h + 1;
                        if (h2 == C::kHalves) { h2 = 0; t2 = t + gridDim.x; }
                        if (t2 < total) {
                            const int nt2 = t2 / p.MT, mt2 = t2 - nt2 * p.MT;
                            mbar_expect_tx(res_full, 2 * kSlab);
                            tma_load_2d(sC, &tm_r, res_full, nt2 * BN + h2 * 128, mt2 * kBM);
                            tma_load_2d(sC + kSlab, &tm_r, res_full, nt2 * BN + h2 * 128 + 64, mt2 * kBM);
                        }
                    }
                }
            }
        }
        if (leader) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(C::kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------- LoRA t = x . down^T
// t[b*M + m, q] = sum_k x[b, m, k] * down[b][q, k] for q < SR = nseg * rank (<= 16); columns [SR, 64) of the 64-wide t rows
// stay zero (the buffer is zeroed once by the caller) — t is the A operand of the GEMM's K = 16 tail step.
struct DownPtrs { const void* down[kMaxBatchRows]; };
constexpr int kTWarps = 8, kTRowsPerWarp = 4;

template <typename T, int SR>
__global__ void __launch_bounds__(kTWarps * 32)
lora_t_kernel(const T* __restrict__ x, T* __restrict__ t_out, const __grid_constant__ DownPtrs ptrs, int M, int K, long long ldx) {
    extern __shared__ float s_down[];                                    // [SR][K] fp32
    const int b = blockIdx.y;
    const T* down = reinterpret_cast<const T*>(ptrs.down[b]);
    if (down == nullptr) return;
    for (int i = threadIdx.x; i < SR * K / 8; i += blockDim.x) {
        float f[8];
        unpack8<T>(ld_keep(down + (size_t)i * 8), f);
        *reinterpret_cast<float4*>(s_down + (size_t)i * 8) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(s_down + (size_t)i * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec = K >> 3;
    const int row0 = (blockIdx.x * kTWarps + warp) * kTRowsPerWarp;
#pragma unroll 1
    for (int rr = 0; rr < kTRowsPerWarp; rr += 2) {
        const int m0 = row0 + rr, m1 = m0 + 1;
        if (m0 >= M) break;
        const bool has1 = m1 < M;
        const T* x0 = x + ((size_t)b * M + m0) * ldx;
        const T* x1 = x + ((size_t)b * M + (has1 ? m1 : m0)) * ldx;
        float t0[SR], t1[SR];
#pragma unroll
        for (int q = 0; q < SR; ++q) { t0[q] = 0.f; t1[q] = 0.f; }
        for (int v = lane; v < nvec; v += 32) {
            float f0[8], f1[8];
            unpack8<T>(ld_keep(x0 + (size_t)v * 8), f0);
            unpack8<T>(ld_keep(x1 + (size_t)v * 8), f1);
#pragma unroll
            for (int q = 0; q < SR; ++q) {
                const float4 d0 = *reinterpret_cast<const float4*>(s_down + (size_t)q * K + v * 8);
                const float4 d1 = *reinterpret_cast<const float4*>(s_down + (size_t)q * K + v * 8 + 4);
                t0[q] = fmaf(f0[0], d0.x, fmaf(f0[1], d0.y, fmaf(f0[2], d0.z, fmaf(f0[3], d0.w,
                        fmaf(f0[4], d1.x, fmaf(f0[5], d1.y, fmaf(f0[6], d1.z, fmaf(f0[7], d1.w, t0[q]))))))));
                t1[q] = fmaf(f1[0], d0.x, fmaf(f1[1], d0.y, fmaf(f1[2], d0.z, fmaf(f1[3], d0.w,
                        fmaf(f1[4], d1.x, fmaf(f1[5], d1.y, fmaf(f1[6], d1.z, fmaf(f1[7], d1.w, t1[q]))))))));
            }
        }
#pragma unroll
        for (int q = 0; q < SR; ++q) { t0[q] = warp_sum(t0[q]); t1[q] = warp_sum(t1[q]); }
        // lane q writes element q (SR <= 16 lanes per row; 16-bit stores, 32 B per row)
        float mine0 = 0.f, mine1 = 0.f;
#pragma unroll
        for (int q = 0; q < SR; ++q) { if (lane == q) { mine0 = t0[q]; mine1 = t1[q]; } }
        if (lane < SR) {
            t_out[((size_t)b * M + m0) * 64 + lane] = (T)mine0;
            if (has1) t_out[((size_t)b * M + m1) * 64 + lane] = (T)mine1;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_force_bn = 0;     // tuning hook: 0 = heuristic, 128 / 256 = forced tile width

// 2-D map over a row-major [rows, cols] 16-bit matrix with row stride ld (elements): dims (cols, rows), box (64, box_rows)
static int encode2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows, bool bf16) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("linear: cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return TMX_ECUDA; }
    return TMX_OK;
}

// Tile width: fewest (waves x tile time).  A 128-wide tile re-reads the x tile twice as often and runs the tensor pipe at the
// shared-memory operand bandwidth, so it must win by more than kNarrowPenalty to be chosen.
static int pick_bn(int M, int N) {
    if (g_force_bn == 128 || g_force_bn == 256) return g_force_bn;
    const long long MT = (M + kBM - 1) / kBM, sms = sm_count();
    auto cost = [&](int bn) { const long long tiles = MT * ((N + bn - 1) / bn); return (double)((tiles + sms - 1) / sms) * bn; };
    constexpr double kNarrowPenalty = 1.12;
    return cost(128) * kNarrowPenalty < cost(256) ? 128 : 256;
}

}  // namespace k10

int linear_init() {
    using namespace k10;
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        TMX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        TMX_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, TMX_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
        g_encode = (EncodeTiledFn)fn;
    }
    TMX_CUDA(cudaFuncSetAttribute(linear_kernel<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::kSmem));
    TMX_CUDA(cudaFuncSetAttribute(linear_kernel<false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::kSmem));
    TMX_CUDA(cudaFuncSetAttribute(linear_kernel<true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::kSmem));
    TMX_CUDA(cudaFuncSetAttribute(linear_kernel<false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::kSmem));
#define K10_ATTR(T, SR) TMX_CUDA(cudaFuncSetAttribute(lora_t_kernel<T, SR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))
    K10_ATTR(__half, 4); K10_ATTR(__half, 8); K10_ATTR(__half, 12); K10_ATTR(__half, 16);
    K10_ATTR(__nv_bfloat16, 4); K10_ATTR(__nv_bfloat16, 8); K10_ATTR(__nv_bfloat16, 12); K10_ATTR(__nv_bfloat16, 16);
#undef K10_ATTR
    return TMX_OK;
}

}  // namespace tmx

using namespace tmx;
using namespace tmx::k10;

extern "C" int tmx_linear_set_variant(int bn) {
    TMX_REQUIRE(bn == 0 || bn == 128 || bn == 256, TMX_EINVAL, "linear_set_variant: tile width must be 0 (auto), 128 or 256");
    g_force_bn = bn;
    return TMX_OK;
}

extern "C" int tmx_linear_fwd(const void* x, const void* w, const float* bias, const void* residual, void* y,
                              int M, int N, int K, int64_t ldx, int64_t ldr, int64_t ldy, int epilogue,
                              const void* lora_t, const void* const* lora_up, int lora_rows_per_batch, int lora_batch,
                              int dtype, void* stream) {
    TMX_REQUIRE(x && w && y, TMX_EINVAL, "linear: null pointer");
    TMX_REQUIRE(M > 0 && N > 0 && K > 0, TMX_EINVAL, "linear: non-positive size");
    TMX_REQUIRE(dtype == TMX_F16 || dtype == TMX_BF16, TMX_EDTYPE, "linear: dtype %d unsupported (fp16/bf16 only)", dtype);
    TMX_REQUIRE(epilogue == TMX_EPI_NONE || epilogue == TMX_EPI_GEGLU, TMX_EINVAL, "linear: unknown epilogue %d", epilogue);
    TMX_REQUIRE(K % 64 == 0, TMX_ESHAPE, "linear: K=%d must be a multiple of 64", K);
    TMX_REQUIRE(N % 8 == 0 && (epilogue != TMX_EPI_GEGLU || N % 64 == 0), TMX_ESHAPE, "linear: N=%d must be a multiple of 8 (64 with the GEGLU epilogue)", N);
    const int n_out = epilogue == TMX_EPI_GEGLU ? N / 2 : N;
    TMX_REQUIRE(ldx >= K && ldy >= n_out && ldx % 8 == 0 && ldy % 8 == 0, TMX_ESHAPE, "linear: bad row strides ldx=%lld ldy=%lld", (long long)ldx, (long long)ldy);
    TMX_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && aligned16(bias) && aligned16(residual), TMX_EALIGN, "linear: 16-byte alignment");
    TMX_REQUIRE(!(residual && epilogue == TMX_EPI_GEGLU), TMX_EINVAL, "linear: the GEGLU epilogue takes no residual");
    TMX_REQUIRE(!residual || (ldr >= N && ldr % 8 == 0), TMX_ESHAPE, "linear: bad residual row stride %lld", (long long)ldr);
    if (int rc = require_init()) return rc;
    const bool bf16 = dtype == TMX_BF16;
    const int BN = pick_bn(M, N);

    Params p;
    p.M = M; p.N = N; p.K = K;
    p.MT = (M + kBM - 1) / kBM; p.NT = (N + BN - 1) / BN;
    p.epilogue = epilogue; p.has_res = residual != nullptr; p.bias = bias;
    p.has_tail = 0; p.rows_per_batch = 1; p.tail_mask = 0;
    CUtensorMap mx, mw, my, mr, mt;
    UpMaps ups;
    if (int rc = encode2d(&mx, x, (uint64_t)K, (uint64_t)M, (uint64_t)ldx, kBM, bf16)) return rc;
    if (int rc = encode2d(&mw, w, (uint64_t)K, (uint64_t)N, (uint64_t)K, (uint32_t)BN, bf16)) return rc;
    if (int rc = encode2d(&my, y, (uint64_t)n_out, (uint64_t)M, (uint64_t)ldy, kBM, bf16)) return rc;
    mr = my; mt = mx;
    if (residual) { if (int rc = encode2d(&mr, residual, (uint64_t)N, (uint64_t)M, (uint64_t)ldr, kBM, bf16)) return rc; }
    for (int b = 0; b < kMaxBatchRows; ++b) ups.m[b] = mw;
    if (lora_t) {
        TMX_REQUIRE(lora_up && lora_batch >= 1 && lora_batch <= kMaxBatchRows, TMX_EINVAL, "linear: LoRA tail needs lora_up and 1..%d batch rows", kMaxBatchRows);
        TMX_REQUIRE(lora_rows_per_batch > 0 && lora_rows_per_batch % kBM == 0 && (long long)lora_rows_per_batch * lora_batch == M, TMX_ESHAPE,
                    "linear: LoRA tail needs rows_per_batch (%d) to be a multiple of 128 and rows_per_batch * batch == M", lora_rows_per_batch);
        TMX_REQUIRE(aligned16(lora_t), TMX_EALIGN, "linear: lora_t alignment");
        if (int rc = encode2d(&mt, lora_t, 64, (uint64_t)M, 64, kBM, bf16)) return rc;
        for (int b = 0; b < lora_batch; ++b) {
            if (!lora_up[b]) continue;
            TMX_REQUIRE(aligned16(lora_up[b]), TMX_EALIGN, "linear: lora_up[%d] alignment", b);
            if (int rc = encode2d(&ups.m[b], lora_up[b], 64, (uint64_t)N, 64, (uint32_t)BN, bf16)) return rc;
            p.tail_mask |= 1u << b;
        }
        p.has_tail = p.tail_mask != 0;
        p.rows_per_batch = lora_rows_per_batch;
    }
    const long long tiles = (long long)p.MT * p.NT;
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    cudaStream_t st = (cudaStream_t)stream;
    if (BN == 256) {
        if (bf16) linear_kernel<true, 256><<<grid, kThreads, Cfg<256>::kSmem, st>>>(mx, mw, my, mr, mt, ups, p);
        else      linear_kernel<false, 256><<<grid, kThreads, Cfg<256>::kSmem, st>>>(mx, mw, my, mr, mt, ups, p);
    } else {
        if (bf16) linear_kernel<true, 128><<<grid, kThreads, Cfg<128>::kSmem, st>>>(mx, mw, my, mr, mt, ups, p);
        else      linear_kernel<false, 128><<<grid, kThreads, Cfg<128>::kSmem, st>>>(mx, mw, my, mr, mt, ups, p);
    }
    return check_cuda(cudaGetLastError(), "linear_kernel launch");
}

extern "C" int tmx_lora_t_fwd(const void* x, const void* const* lora_down, void* t, int B, int M, int K, int64_t ldx,
                              int sr, int dtype, void* stream) {
    TMX_REQUIRE(x && lora_down && t, TMX_EINVAL, "lora_t: null pointer");
    TMX_REQUIRE(B >= 1 && B <= kMaxBatchRows && M > 0 && K > 0 && K % 8 == 0 && ldx >= K && ldx % 8 == 0, TMX_ESHAPE, "lora_t: B=%d M=%d K=%d ldx=%lld", B, M, K, (long long)ldx);
    TMX_REQUIRE(dtype == TMX_F16 || dtype == TMX_BF16, TMX_EDTYPE, "lora_t: dtype %d unsupported", dtype);
    TMX_REQUIRE((size_t)sr * K * sizeof(float) <= 200 * 1024, TMX_ESHAPE, "lora_t: down factors (%d x %d) do not fit in shared memory", sr, K);
    TMX_REQUIRE(aligned16(x) && aligned16(t), TMX_EALIGN, "lora_t: 16-byte alignment");
    if (int rc = require_init()) return rc;
    DownPtrs ptrs;
    bool any = false;
    for (int b = 0; b < kMaxBatchRows; ++b) {
        ptrs.down[b] = b < B ? lora_down[b] : nullptr;
        TMX_REQUIRE(aligned16(ptrs.down[b]), TMX_EALIGN, "lora_t: down[%d] alignment", b);
        any |= ptrs.down[b] != nullptr;
    }
    if (!any) return TMX_OK;
    const size_t smem = (size_t)sr * K * sizeof(float);
    dim3 grid((M + kTWarps * kTRowsPerWarp - 1) / (kTWarps * kTRowsPerWarp), B);
    cudaStream_t st = (cudaStream_t)stream;
#define K10_LAUNCH(T, SR) lora_t_kernel<T, SR><<<grid, kTWarps * 32, smem, st>>>((const T*)x, (T*)t, ptrs, M, K, (long long)ldx)
    const bool bf16 = dtype == TMX_BF16;
    switch (sr) {
        case 4:  if (bf16) K10_LAUNCH(__nv_bfloat16, 4);  else K10_LAUNCH(__half, 4);  break;
        case 8:  if (bf16) K10_LAUNCH(__nv_bfloat16, 8);  else K10_LAUNCH(__half, 8);  break;
        case 12: if (bf16) K10_LAUNCH(__nv_bfloat16, 12); else K10_LAUNCH(__half, 12); break;
        case 16: if (bf16) K10_LAUNCH(__nv_bfloat16, 16); else K10_LAUNCH(__half, 16); break;
        default: set_error("lora_t: nseg * rank = %d unsupported (4, 8, 12 or 16)", sr); return TMX_ESHAPE;
    }
#undef K10_LAUNCH
    return check_cuda(cudaGetLastError(), "lora_t_kernel launch");
}
