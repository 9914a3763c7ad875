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
constexpr int kMaxTailTiles = 256;              // > any SM count: tail tiles < grid <= SMs
constexpr int kMaxSplit = 8;

template <int BN> struct Cfg {
    static constexpr int kBTile = BN * kBK * 2;                  // 16 / 24 / 32 / 40 KiB
    static constexpr int kStages = BN == 128 ? 5 : (BN == 320 ? 3 : 4);
    static constexpr int kRing = kStages * (kATile + kBTile);    // 160 / 160 / 192 / 168 KiB
    static constexpr int kStage = 2 * kSlab;                     // staging tile: 128 rows x 128 columns, 32 KiB
    static constexpr int kBars = 2 * kStages + 4 + 1;
    static constexpr int kSmem = 1024 + kRing + kStage + kBars * 8 + 16;
    static constexpr int kTmemCols = BN == 128 ? 256 : 512;      // accumulator stages of BN columns (power of two)
    static constexpr int kAccStages = 2 * BN <= 512 ? 2 : 1;     // BN = 320: one 320-column accumulator (single-wave shapes: nothing to overlap)
    static constexpr int kHalves = (BN + 127) / 128;             // BN = 192 / 320: the last half holds 64 columns
    static constexpr int kNMma = BN > 256 ? 2 : 1;               // UMMA N <= 256: a 320-wide tile is two N = 160 instructions per K step
    static constexpr int kN = BN / kNMma;
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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : K10_R8(r, 0), K10_R8(r, 8) : "r"(taddr) : "memory");
}

// Exact-GELU gate 0.5 g (1 + erf(g / sqrt 2)) with erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below the 2^-9 /
// 2^-12 rounding of the 16-bit output): one MUFU.RCP + one MUFU.EX2 + 8 FMA instead of erff's ~40 instructions — the GEGLU
// epilogue of a K = 640 tile has only ~5000 clk before the next accumulator is ready.
__device__ __forceinline__ float gelu_erf(float g) {
    const float x = fabsf(g) * 0.70710678118654752f;
    const float t = __fdividef(1.f, fmaf(0.3275911f, x, 1.f));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float e = poly * t * exp2f(-1.4426950408889634f * x * x);     // 1 - erf(|x|)
    const float one_plus_erf = g >= 0.f ? 2.f - e : e;                   // 1 + erf(x) without cancellation for negative g
    return 0.5f * g * one_plus_erf;
}

struct UpMaps { CUtensorMap m[kMaxBatchRows]; };

struct Params {
    int M, N, K;               // problem (N = accumulator columns = rows of W)
    int MT, NT;                // tile counts
    int epilogue;              // TMX_EPI_*
    int has_res, has_tail;
    int rows_per_batch;        // LoRA tail: rows of x per batch row (multiple of 128)
    unsigned tail_mask;        // bit b set = batch row b carries LoRA factors
    const float* bias;         // [N] fp32 or null
    // split-K tail: the tiles beyond the last full wave are cut `split` ways along K so that every SM has work in the last
    // round; partial accumulators meet in an fp32 workspace and are summed in a FIXED order (deterministic)
    int full_tiles;            // tiles computed whole (a multiple of the grid when split >= 2, else all tiles)
    int split;                 // S >= 2, or 0
    int tail_units;            // (tiles - full_tiles) * S
    float* ws;                 // [tail tile][S][half][128 rows][128 cols] fp32
    unsigned int* sync;        // [2][tail tile] arrival / done counters, zero between launches
};

// advance (stage, phase) of the K ring
#define K10_NEXT(st, ph, NST) do { if (++(st) == (NST)) { (st) = 0; (ph) ^= 1u; } } while (0)

// Work item i of this CTA: a whole tile (s < 0) or one K-slice of a tail tile.
struct Item { int tile, k0, kn, s, tail_idx; };
__device__ __forceinline__ bool get_item(const Params& p, int i, int KT, Item& it) {
    const int t = (int)blockIdx.x + i * (int)gridDim.x;
    if (t < p.full_tiles) { it.tile = t; it.k0 = 0; it.kn = KT; it.s = -1; it.tail_idx = 0; return true; }
    if (p.split >= 2 && i == p.full_tiles / (int)gridDim.x && (int)blockIdx.x < p.tail_units) {
        const int u = blockIdx.x;
        it.tail_idx = u / p.split;
        it.tile = p.full_tiles + it.tail_idx;
        it.s = u - it.tail_idx * p.split;
        const int base = KT / p.split, rem = KT - base * p.split;
        it.k0 = it.s * base + (it.s < rem ? it.s : rem);
        it.kn = base + (it.s < rem ? 1 : 0);
        return true;
    }
    return false;
}

template <bool BF16, int BN, bool SPLIT>
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
        Item it;
        for (int i = 0; get_item(p, i, KT, it); ++i) {
            const int nt = it.tile / p.MT, mt = it.tile - nt * p.MT;
            for (int kt = it.k0; kt < it.k0 + it.kn; ++kt) {
                mbar_wait(empty + 8 * st, ph ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(full + 8 * st, kATile + C::kBTile);
                    tma_load_2d(sA + st * kATile, &tm_x, full + 8 * st, kt * kBK, mt * kBM);
#pragma unroll
                    for (int m = 0; m < C::kNMma; ++m)                  // (TMA boxes hold <= 256 rows: a 320-wide W tile is two boxes)
                        tma_load_2d(sB + st * C::kBTile + m * (C::kN * 128), &tm_w, full + 8 * st, kt * kBK, nt * BN + m * C::kN);
                }
                K10_NEXT(st, ph, NST);
            }
            if (p.has_tail && it.s <= 0) {                              // the LoRA step rides with the whole tile / the first K-slice
                const int b = (mt * kBM) / p.rows_per_batch;
                if ((p.tail_mask >> b) & 1u) {
                    mbar_wait(empty + 8 * st, ph ^ 1u);
                    if (elect_one()) {
                        mbar_expect_tx(full + 8 * st, kATile + C::kBTile);
                        tma_load_2d(sA + st * kATile, &tm_t, full + 8 * st, 0, mt * kBM);
#pragma unroll
                        for (int m = 0; m < C::kNMma; ++m)
                            tma_load_2d(sB + st * C::kBTile + m * (C::kN * 128), &tm_up.m[b], full + 8 * st, 0, nt * BN + m * C::kN);
                    }
                    K10_NEXT(st, ph, NST);
                }
            }
        }
    } else if (warp == 1) {
        // =========================================== MMA issuer =============================================
        constexpr uint32_t idesc = make_idesc(BF16, kBM, C::kN);
        constexpr uint32_t kBHalf = (uint32_t)(C::kN * 128) >> 4;       // descriptor offset of the second N = kN half of the W tile
        int st = 0;
        uint32_t ph = 0;
        Item it;
        for (int i = 0; get_item(p, i, KT, it); ++i) {
            const uint32_t a = C::kAccStages == 2 ? ((uint32_t)i & 1u) : 0u;
            const uint32_t use = C::kAccStages == 2 ? ((uint32_t)i >> 1) : (uint32_t)i;
            const int mt = it.tile % p.MT;
            mbar_wait(acc_empty + 8 * a, (use & 1u) ^ 1u);                // the epilogue has drained this accumulator stage
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + a * BN;
            for (int kt = 0; kt < it.kn; ++kt) {
                mbar_wait(full + 8 * st, ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_lo = desc_lo(sA + st * kATile), b_lo = desc_lo(sB + st * C::kBTile);
#pragma unroll
                    for (int m = 0; m < C::kNMma; ++m) {
                        umma_ss(d_tmem + m * C::kN, a_lo, b_lo + m * kBHalf, idesc, kt > 0 ? 1u : 0u);
                        umma_ss(d_tmem + m * C::kN, a_lo + 2, b_lo + m * kBHalf + 2, idesc, 1u);
                        umma_ss(d_tmem + m * C::kN, a_lo + 4, b_lo + m * kBHalf + 4, idesc, 1u);
                        umma_ss(d_tmem + m * C::kN, a_lo + 6, b_lo + m * kBHalf + 6, idesc, 1u);
                    }
                    umma_commit(empty + 8 * st);
                }
                K10_NEXT(st, ph, NST);
            }
            if (p.has_tail && it.s <= 0) {
                const int b = (mt * kBM) / p.rows_per_batch;
                if ((p.tail_mask >> b) & 1u) {                          // ONE K = 16 step: t[128 x 16] . up[b][BN x 16]^T
                    mbar_wait(full + 8 * st, ph);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int m = 0; m < C::kNMma; ++m)
                            umma_ss(d_tmem + m * C::kN, desc_lo(sA + st * kATile), desc_lo(sB + st * C::kBTile) + m * kBHalf, idesc, it.kn > 0 ? 1u : 0u);
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
        uint32_t nres = 0;
        bool first = true;

        auto load_residual = [&](int nt, int mt, int h) {               // leader only: residual half -> staging tile
            const bool two = h * 128 + 64 < BN;                          // (BN = 192: the second half is one slab)
            mbar_expect_tx(res_full, two ? 2 * kSlab : kSlab);
            tma_load_2d(sC, &tm_r, res_full, nt * BN + h * 128, mt * kBM);
            if (two) tma_load_2d(sC + kSlab, &tm_r, res_full, nt * BN + h * 128 + 64, mt * kBM);
        };
        // ---- pieces shared by the two epilogue styles
        auto pre_half = [&]() {
            if (!first) epi_bar();                                       // the leader has drained the previous TMA store (and issued the residual load)
            first = false;
            if (p.has_res) { mbar_wait(res_full, nres & 1u); ++nres; }
        };
        auto post_half = [&](int nt, int mt, int h, bool has_next, int nt2, int mt2, int h2) {
            fence_async_smem();                                          // generic-proxy writes -> visible to the TMA store
            epi_bar();
            if (leader) {
                const int n0 = nt * BN + h * 128;                        // (a half that starts beyond N holds only zero padding)
                if (!geglu) {
                    if (n0 < p.N) tma_store_2d(&tm_y, sC, n0, mt * kBM);
                    if (h * 128 + 64 < BN && n0 + 64 < p.N) tma_store_2d(&tm_y, sC + kSlab, n0 + 64, mt * kBM);
                } else if (n0 < p.N) {
                    tma_store_2d(&tm_y, sC, n0 >> 1, mt * kBM);
                }
                tma_store_commit();
                tma_store_wait_read();                                   // the staging tile may be overwritten
                if (p.has_res && has_next) load_residual(nt2, mt2, h2);  // lands while the main loop of that tile runs
            }
        };
        // 8 accumulator columns (chunk c8 of the thread's 64) -> + bias (+ residual) -> one rounding -> staging tile
        auto chunk_plain = [&](const uint32_t* v8, int c8, int n_acc) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v8[e]);
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
            const uint4 o = pack8<T>(f);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
        };
        // 8 value + 8 gate accumulator columns -> 8 outputs (output chunk `oc` of the thread's 32) -> staging tile
        auto chunk_geglu = [&](const uint32_t* v8, const uint32_t* g8, int oc, int n_acc) {
            float f[8];
            const int n = n_acc + oc * 8;
            float bv[8] = {0, 0, 0, 0, 0, 0, 0, 0}, bg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (p.bias != nullptr && n < p.N) {
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(p.bias + n)), v1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 32)), g1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 36));
                bv[0] = v0.x; bv[1] = v0.y; bv[2] = v0.z; bv[3] = v0.w; bv[4] = v1.x; bv[5] = v1.y; bv[6] = v1.z; bv[7] = v1.w;
                bg[0] = g0.x; bg[1] = g0.y; bg[2] = g0.z; bg[3] = g0.w; bg[4] = g1.x; bg[5] = g1.y; bg[6] = g1.z; bg[7] = g1.w;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float a = __uint_as_float(v8[e]) + bv[e];
                const float b = __uint_as_float(g8[e]) + bg[e];
                // the un-fused path rounds the projection to 16 bits before the gating kernel reads it
                f[e] = Pack2<T>::round(a) * gelu_erf(Pack2<T>::round(b));
            }
            const uint4 o = pack8<T>(f);
            const uint32_t addr = c_row + (((uint32_t)(cg * 4 + oc) ^ sw) << 4);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
        };
        // Whole-tile style: all 64 columns of the thread in registers at once (two wide TMEM loads in flight, straight-line code)
        auto emit_half_wide = [&](uint32_t t_col, uint32_t release_bar, int nt, int mt, int h, bool has_next, int nt2, int mt2, int h2) {
            const int n_acc = nt * BN + h * 128 + cg * 64;
            const bool mine = h * 128 + cg * 64 < BN;                    // BN = 192: column group 1 has nothing in the second half
            uint32_t acc[64];
            if (mine) {
                tmem_ld32(t_col, acc);
                tmem_ld32(t_col + 32, acc + 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            if (release_bar) {                                           // accumulator stage fully read: hand it back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(release_bar);
            }
            pre_half();
            if (mine) {
                if (!geglu) {
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8) chunk_plain(acc + c8 * 8, c8, n_acc);
                } else {
#pragma unroll
                    for (int oc = 0; oc < 4; ++oc) chunk_geglu(acc + oc * 8, acc + 32 + oc * 8, oc, n_acc);
                }
            }
            post_half(nt, mt, h, has_next, nt2, mt2, h2);
        };
        // Split-tile style: the thread's 64 columns arrive 16 at a time through `fetch16(col, v)` (TMEM, or the summed workspace
        // partials) — few live registers, so the tail path fits beside the rest under the 168-register cap of a 10-warp CTA.
        auto emit_half = [&](auto&& fetch16, int nt, int mt, int h, bool has_next, int nt2, int mt2, int h2) {
            const int n_acc = nt * BN + h * 128 + cg * 64;              // first accumulator column of this thread's 64
            pre_half();
            if (!geglu) {
#pragma unroll 1
                for (int q4 = 0; q4 < 4; ++q4) {                         // 16 columns = two 16-byte chunks
                    uint32_t v[16];
                    fetch16(q4 * 16, v);
                    chunk_plain(v, q4 * 2, n_acc);
                    chunk_plain(v + 8, q4 * 2 + 1, n_acc);
                }
            } else {
#pragma unroll 1
                for (int q2 = 0; q2 < 2; ++q2) {                         // 16 value + 16 gate columns -> 16 outputs
                    uint32_t v[16], g[16];
                    fetch16(q2 * 16, v);
                    fetch16(32 + q2 * 16, g);
                    chunk_geglu(v, g, q2 * 2, n_acc);
                    chunk_geglu(v + 8, g + 8, q2 * 2 + 1, n_acc);
                }
            }
            post_half(nt, mt, h, has_next, nt2, mt2, h2);
        };

        // workspace layout of one (tail tile, slice, half): [epilogue warp][float4 index within the thread's 64 columns][lane],
        // so that every warp-wide access is one contiguous 512-byte run
        constexpr size_t kHalfF4 = (size_t)kEpiWarps * 16 * 32;          // float4 per half = 128 x 128 fp32
        Item it, nx;
        nx.tile = 0;
        if (p.has_res && leader && get_item(p, 0, KT, it) && it.s < 0) load_residual(it.tile / p.MT, it.tile % p.MT, 0);
        for (int i = 0; get_item(p, i, KT, it); ++i) {
            const uint32_t a = C::kAccStages == 2 ? ((uint32_t)i & 1u) : 0u;
            const uint32_t use = C::kAccStages == 2 ? ((uint32_t)i >> 1) : (uint32_t)i;
            const int nt = it.tile / p.MT, mt = it.tile - nt * p.MT;
            mbar_wait(acc_full + 8 * a, use & 1u);
            tc_fence_after();
            const uint32_t t_acc = t_row + a * BN + cg * 64;
            if (it.s < 0) {
                // ---- whole tile
                const bool more = get_item(p, i + 1, KT, nx) && nx.s < 0;
#pragma unroll 1
                for (int h = 0; h < C::kHalves; ++h) {
                    const bool last_half = h == C::kHalves - 1;
                    const uint32_t t_col = t_acc + h * 128;
                    if constexpr (!SPLIT) {
                        emit_half_wide(t_col, last_half ? acc_empty + 8 * a : 0u, nt, mt, h, !last_half || more, last_half ? nx.tile / p.MT : nt, last_half ? nx.tile % p.MT : mt, last_half ? 0 : h + 1);
                    } else {
                        emit_half([&](int col, uint32_t (&v)[16]) { tmem_ld16(t_col + col, v); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); },
                                  nt, mt, h, !last_half || more, last_half ? nx.tile / p.MT : nt, last_half ? nx.tile % p.MT : mt, last_half ? 0 : h + 1);
                    }
                }
                if constexpr (SPLIT) {
                    tc_fence_before();                                   // accumulator stage fully read: hand it back to the MMA warp
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + 8 * a);
                }
            } else if constexpr (SPLIT) {
                // ---- one K-slice of a tail tile: park the partial accumulator in the workspace, then (slices 0 .. halves-1) sum
                // the S partials of half h = s in slice order and emit it
                float4* wbase = reinterpret_cast<float4*>(p.ws) + ((size_t)it.tail_idx * p.split + it.s) * (C::kHalves * kHalfF4) + (size_t)ew * 16 * 32 + lane;
#pragma unroll 1
                for (int h = 0; h < C::kHalves; ++h) {
#pragma unroll 1
                    for (int q4 = 0; q4 < 4; ++q4) {
                        uint32_t v[16];
                        tmem_ld16(t_acc + h * 128 + q4 * 16, v);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4)
                            __stcg(wbase + h * kHalfF4 + (size_t)(q4 * 4 + c4) * 32,
                                   make_float4(__uint_as_float(v[4 * c4]), __uint_as_float(v[4 * c4 + 1]), __uint_as_float(v[4 * c4 + 2]), __uint_as_float(v[4 * c4 + 3])));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + 8 * a);
                __threadfence();                                         // partials visible device-wide before the arrival below
                epi_bar();
                unsigned int* cnt = p.sync + it.tail_idx;
                unsigned int* done = p.sync + kMaxTailTiles + it.tail_idx;
                if (leader) atomicAdd(cnt, 1u);
                if (it.s < C::kHalves) {
                    const int h = it.s;
                    if (leader) {
                        if (p.has_res) load_residual(nt, mt, h);
                        unsigned int cur, polls = 0;
                        long long t0 = 0;
                        while (true) {                                   // all S slices are co-resident (cooperative launch): bounded spin
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(cnt) : "memory");
                            if (cur >= (unsigned)p.split) break;
                            if ((++polls & 1023u) == 0) {
                                const long long now = clock64();
                                if (t0 == 0) t0 = now; else if (now - t0 > 4000000000LL) { atomicExch(&g_k10_timeout_flag, 2u); __trap(); }
                            }
                        }
                    }
                    epi_bar();
                    const float4* rbase = reinterpret_cast<const float4*>(p.ws) + (size_t)it.tail_idx * p.split * (C::kHalves * kHalfF4) + h * kHalfF4 + (size_t)ew * 16 * 32 + lane;
                    const int S = p.split;
                    emit_half([&](int col, uint32_t (&v)[16]) {
                                  float4 acc4[4];
#pragma unroll
                                  for (int c4 = 0; c4 < 4; ++c4) acc4[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
                                  for (int s2 = 0; s2 < S; ++s2) {       // fixed summation order: bit-reproducible
                                      const float4* src = rbase + (size_t)s2 * (C::kHalves * kHalfF4) + (size_t)(col >> 2) * 32;
#pragma unroll
                                      for (int c4 = 0; c4 < 4; ++c4) {
                                          const float4 w = __ldcg(src + c4 * 32);
                                          acc4[c4].x += w.x; acc4[c4].y += w.y; acc4[c4].z += w.z; acc4[c4].w += w.w;
                                      }
                                  }
#pragma unroll
                                  for (int c4 = 0; c4 < 4; ++c4) {
                                      v[4 * c4] = __float_as_uint(acc4[c4].x); v[4 * c4 + 1] = __float_as_uint(acc4[c4].y);
                                      v[4 * c4 + 2] = __float_as_uint(acc4[c4].z); v[4 * c4 + 3] = __float_as_uint(acc4[c4].w);
                                  }
                              },
                              nt, mt, h, false, 0, 0, 0);
                    if (leader) {                                        // last finaliser of the tile re-arms its counters for the next launch
                        const unsigned nfinal = (unsigned)(p.split < C::kHalves ? p.split : C::kHalves);
                        if (atomicAdd(done, 1u) == nfinal - 1u) { *cnt = 0u; *done = 0u; __threadfence(); }
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
static int g_split_mode = 0;   // tuning hook: 0 = split-K tail when a workspace is given and K is long, 1 = never, 2 = at any K

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

// Split-K plan for `tiles` tiles on `sms` SMs: the r = tiles % sms tiles of the last, partly filled round are cut S ways along
// K (S <= sms / r, <= kMaxSplit, >= kMinSliceK K blocks per slice) when a workspace is available and the K loop is long
// enough for the workspace round trip to pay (measured, profiles/r02d_kbench_linear.txt: it does at K = 5120, not at 1280).
constexpr int kMinSplitKT = 40, kMinSliceK = 8;
struct Plan { int bn, full_tiles, split, tail_units, grid; double cost; };
static Plan make_plan(int M, int N, int KT, int bn, int sms, bool can_split, int epilogue = TMX_EPI_NONE, bool has_res = false) {
    Plan pl;
    pl.bn = bn;
    const long long tiles = (long long)((M + kBM - 1) / kBM) * ((N + bn - 1) / bn);
    const int q = (int)(tiles / sms), r = (int)(tiles % sms);
    int S = 0;
    const bool eager = g_split_mode == 2;                              // tuning hook: split whenever a slice keeps >= 4 K blocks
    if (can_split && r > 0 && (KT >= kMinSplitKT || eager) && bn != 192 && bn != 320) {
        S = sms / r;
        if (S > kMaxSplit) S = kMaxSplit;
        const int min_slice = eager ? 4 : kMinSliceK;
        if (S > KT / min_slice) S = KT / min_slice;
        if (S < 2) S = 0;
    }
    // Cost model fitted to profiles/r02j_kbench_linear.txt.  A tile takes the longer of its main loop — KT K-blocks of four
    // MMAs at bn/2 clk each, slowed by the shared-memory operand bandwidth for narrower tiles (0.78 / 0.93 of the 256-wide rate) —
    // and its epilogue (~4000 clk per 128 columns, more with a residual or the GEGLU gating), which bounds the K = 640 shapes.
    const double rate = bn == 256 ? 1.0 : (bn == 320 ? 0.95 : (bn == 192 ? 0.93 : 0.78));
    const double main_clk = KT * 2.0 * bn / rate;
    const double epi_clk = bn / 128.0 * (4000.0 + (has_res ? 1000.0 : 0.0) + (epilogue == TMX_EPI_GEGLU ? 2000.0 : 0.0));
    // one accumulator stage (320-wide tiles): the epilogue of a tile does not overlap the next main loop
    const double tile_clk = bn == 320 ? (q + (r > 0) > 1 ? main_clk + epi_clk : main_clk) : (main_clk > epi_clk ? main_clk : epi_clk);
    double rounds;
    if (S >= 2) {
        pl.full_tiles = q * sms; pl.split = S; pl.tail_units = r * S;
        pl.grid = q > 0 ? sms : pl.tail_units;
        rounds = q + 1.0 / S + (bn == 256 ? 0.25 : 0.5);               // + workspace round trip, finalisers' reads (one finaliser per 128-wide tile), slower epilogue style
    } else {
        pl.full_tiles = (int)tiles; pl.split = 0; pl.tail_units = 0;
        pl.grid = (int)(tiles < sms ? tiles : sms);
        rounds = q + (r > 0 ? 1.0 : 0.0);
    }
    pl.cost = rounds * tile_clk;
    return pl;
}

// Tile width and split: the plan with the fewest (rounds x tile time).
static Plan pick_plan(int M, int N, int KT, int epilogue, bool has_res, bool can_split) {
    const int sms = sm_count();
    can_split = can_split && g_split_mode != 1;
    if (g_force_bn == 320 && epilogue == TMX_EPI_GEGLU) return make_plan(M, N, KT, 256, sms, can_split, epilogue, has_res);
    if (g_force_bn == 128 || g_force_bn == 192 || g_force_bn == 256 || g_force_bn == 320) return make_plan(M, N, KT, g_force_bn, sms, can_split, epilogue, has_res);
    Plan best = make_plan(M, N, KT, N <= 128 ? 128 : 256, sms, can_split, epilogue, has_res);
    if (N > 128) {
        const Plan p128 = make_plan(M, N, KT, 128, sms, can_split, epilogue, has_res);
        if (p128.cost < best.cost) best = p128;
        if (epilogue != TMX_EPI_GEGLU) {                               // (a 192-wide GEGLU tile would end in half an output slab)
            const Plan p192 = make_plan(M, N, KT, 192, sms, can_split, epilogue, has_res);
            if (p192.cost < best.cost * 0.97) best = p192;
            if (N > 256) {                                             // 320-wide single-accumulator tiles: the N = 1280 / 640 shapes in ONE or two waves
                const Plan p320 = make_plan(M, N, KT, 320, sms, can_split, epilogue, has_res);
                if (p320.cost < best.cost * 0.97) best = p320;
            }
        }
    }
    return best;
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
#define K10_KATTR(B16, BN, SP) TMX_CUDA(cudaFuncSetAttribute(linear_kernel<B16, BN, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::kSmem))
    K10_KATTR(true, 128, false); K10_KATTR(false, 128, false); K10_KATTR(true, 192, false); K10_KATTR(false, 192, false);
    K10_KATTR(true, 256, false); K10_KATTR(false, 256, false); K10_KATTR(true, 320, false); K10_KATTR(false, 320, false);
    K10_KATTR(true, 128, true); K10_KATTR(false, 128, true); K10_KATTR(true, 256, true); K10_KATTR(false, 256, true);
#undef K10_KATTR
#define K10_ATTR(T, SR) TMX_CUDA(cudaFuncSetAttribute(lora_t_kernel<T, SR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))
    K10_ATTR(__half, 4); K10_ATTR(__half, 8); K10_ATTR(__half, 12); K10_ATTR(__half, 16);
    K10_ATTR(__nv_bfloat16, 4); K10_ATTR(__nv_bfloat16, 8); K10_ATTR(__nv_bfloat16, 12); K10_ATTR(__nv_bfloat16, 16);
#undef K10_ATTR
    return TMX_OK;
}

}  // namespace tmx

using namespace tmx;
using namespace tmx::k10;

extern "C" int tmx_linear_set_variant(int v) {
    // v = tile width (0 auto, 128, 256) + 1000 to disable the split-K tail
    const int bn = v % 1000, nosplit = v / 1000;
    TMX_REQUIRE(v >= 0 && (bn == 0 || bn == 128 || bn == 192 || bn == 256 || bn == 320) && nosplit <= 2, TMX_EINVAL,
                "linear_set_variant: 0 | 128 | 192 | 256 | 320 (+ 1000 = no split-K tail, + 2000 = split-K tail at any K)");
    g_force_bn = bn;
    g_split_mode = nosplit;
    return TMX_OK;
}

extern "C" size_t tmx_linear_workspace_bytes(void) {
    // counters (2 x kMaxTailTiles words, zeroed ONCE by the caller) + fp32 partial tiles: at most one K-slice per SM, 128 KiB each
    return 4096 + (size_t)256 * 2 * 128 * 128 * sizeof(float);
}

extern "C" int tmx_linear_fwd(const void* x, const void* w, const float* bias, const void* residual, void* y,
                              int M, int N, int K, int64_t ldx, int64_t ldr, int64_t ldy, int epilogue,
                              const void* lora_t, const void* const* lora_up, int lora_rows_per_batch, int lora_batch,
                              void* workspace, int dtype, void* stream) {
    TMX_REQUIRE(x && w && y, TMX_EINVAL, "linear: null pointer");
    TMX_REQUIRE(M > 0 && N > 0 && K > 0, TMX_EINVAL, "linear: non-positive size");
    TMX_REQUIRE(dtype == TMX_F16 || dtype == TMX_BF16, TMX_EDTYPE, "linear: dtype %d unsupported (fp16/bf16 only)", dtype);
    TMX_REQUIRE(epilogue == TMX_EPI_NONE || epilogue == TMX_EPI_GEGLU, TMX_EINVAL, "linear: unknown epilogue %d", epilogue);
    TMX_REQUIRE(K % 64 == 0, TMX_ESHAPE, "linear: K=%d must be a multiple of 64", K);
    TMX_REQUIRE(N % 8 == 0 && (epilogue != TMX_EPI_GEGLU || N % 64 == 0), TMX_ESHAPE, "linear: N=%d must be a multiple of 8 (64 with the GEGLU epilogue)", N);
    const int n_out = epilogue == TMX_EPI_GEGLU ? N / 2 : N;
    TMX_REQUIRE(ldx >= K && ldy >= n_out && ldx % 8 == 0 && ldy % 8 == 0, TMX_ESHAPE, "linear: bad row strides ldx=%lld ldy=%lld", (long long)ldx, (long long)ldy);
    TMX_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && aligned16(bias) && aligned16(residual) && aligned16(workspace), TMX_EALIGN, "linear: 16-byte alignment");
    TMX_REQUIRE(!(residual && epilogue == TMX_EPI_GEGLU), TMX_EINVAL, "linear: the GEGLU epilogue takes no residual");
    TMX_REQUIRE(!residual || (ldr >= N && ldr % 8 == 0), TMX_ESHAPE, "linear: bad residual row stride %lld", (long long)ldr);
    if (int rc = require_init()) return rc;
    const bool bf16 = dtype == TMX_BF16;
    const int KT = K / kBK;
    Plan pl = pick_plan(M, N, KT, epilogue, residual != nullptr, workspace != nullptr);
    if (pl.bn == 192 && epilogue == TMX_EPI_GEGLU) pl = make_plan(M, N, KT, 256, sm_count(), workspace != nullptr && g_split_mode != 1, epilogue, false);
    const int BN = pl.bn;

    Params p;
    p.M = M; p.N = N; p.K = K;
    p.MT = (M + kBM - 1) / kBM; p.NT = (N + BN - 1) / BN;
    p.epilogue = epilogue; p.has_res = residual != nullptr; p.bias = bias;
    p.has_tail = 0; p.rows_per_batch = 1; p.tail_mask = 0;
    p.full_tiles = pl.full_tiles; p.split = pl.split; p.tail_units = pl.tail_units;
    p.sync = reinterpret_cast<unsigned int*>(workspace);
    p.ws = workspace ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 4096) : nullptr;
    CUtensorMap mx, mw, my, mr, mt;
    UpMaps ups;
    if (int rc = encode2d(&mx, x, (uint64_t)K, (uint64_t)M, (uint64_t)ldx, kBM, bf16)) return rc;
    const uint32_t w_box = (uint32_t)(BN > 256 ? BN / 2 : BN);          // TMA boxes hold <= 256 rows
    if (int rc = encode2d(&mw, w, (uint64_t)K, (uint64_t)N, (uint64_t)K, w_box, bf16)) return rc;
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
            if (int rc = encode2d(&ups.m[b], lora_up[b], 64, (uint64_t)N, 64, w_box, bf16)) return rc;
            p.tail_mask |= 1u << b;
        }
        p.has_tail = p.tail_mask != 0;
        p.rows_per_batch = lora_rows_per_batch;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const bool sp = pl.split >= 2;
    const void* fn = nullptr;
#define K10_PICK(BNv, SPv) (bf16 ? (const void*)linear_kernel<true, BNv, SPv> : (const void*)linear_kernel<false, BNv, SPv>)
    if (BN == 256) fn = sp ? K10_PICK(256, true) : K10_PICK(256, false);
    else if (BN == 320) fn = K10_PICK(320, false);
    else if (BN == 192) fn = K10_PICK(192, false);
    else fn = sp ? K10_PICK(128, true) : K10_PICK(128, false);
#undef K10_PICK
    const size_t smem = BN == 320 ? Cfg<320>::kSmem : (BN == 256 ? Cfg<256>::kSmem : (BN == 192 ? Cfg<192>::kSmem : Cfg<128>::kSmem));
    void* args[] = {&mx, &mw, &my, &mr, &mt, &ups, &p};
    if (pl.split >= 2) {
        // the K-slices of a tail tile wait for each other: the grid (<= one CTA per SM) must be co-resident
        return check_cuda(cudaLaunchCooperativeKernel(fn, dim3(pl.grid), dim3(kThreads), args, smem, st), "linear_kernel cooperative launch");
    }
    return check_cuda(cudaLaunchKernel(fn, dim3(pl.grid), dim3(kThreads), args, smem, st), "linear_kernel launch");
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
