// k1 / k2 — attention forward for head_dim 64 on sm_100a:  O = softmax(scale * Q K^T) V.
//
// Blackwell-native structure (no mma.sync / wgmma):
//   * PERSISTENT, STREAM-K SCHEDULED: the work is the linear space of (unit, K/V tile) iterations,
//     a unit being two 128-row query tiles of one (b, h) that are processed together ("slots" 0 and
//     1) and share every K/V tile (a lone last tile of an odd count runs in slot 0 alone).  One
//     384-thread CTA per SM takes an EQUAL contiguous share of that space, so a unit may be cut
//     along its K/V axis between neighbouring CTAs: the CTA holding the head of a unit owns it, the
//     others leave their un-normalised partial (O, m, l) in a global workspace (a CTA's first
//     piece, so it is ready long before the owner — whose head piece comes LAST in its range —
//     asks for it) and the owner folds them in during its epilogue, in CTA order (deterministic).
//     Every SM gets the same number of tile iterations whatever B*H*tiles is, instead of whole
//     query tiles quantised over 148 SMs (measured before: 5 vs 4.3 steps at B=4 / N=4096, and
//     the same 2 steps at B=2 as at B=4 for N=1024).  The next step's Q (double-buffered) and K/V
//     tiles are prefetched while the current step computes: no per-step prologue bubble;
//   * Q, K, V tiles are staged global -> shared by TMA (cp.async.bulk.tensor.4d, SWIZZLE_128B)
//     straight from the [B, N, H, 64] tensors diffusers hands the hook, no permute;
//   * S = Q K^T and O += P V are tcgen05.mma (kind::f16, M=128) issued by ONE elected thread per
//     slot, accumulators live in TMEM (512 columns: per slot S[0,128) O[128,192) P[192,256)); P is
//     written back to TMEM as packed 16-bit and consumed as the A operand of the second MMA (TS
//     form), so P never touches shared memory;
//   * warp-specialised: warp 0 = TMA producer (+ TMEM alloc), warps 1-2 = MMA issuers (one per
//     slot), warps 4-11 = two 128-thread softmax groups (thread == row == TMEM lane), linked by
//     mbarriers; tcgen05.commit signals MMA completion.  setmaxnreg moves registers from the
//     utility warps (56) to the softmax warps (224) so a full 128-column S row stays in registers;
//   * online softmax in the exp2 domain with lazy rescaling of O (only when the running max
//     moves by more than 2^8), exact because the row sum is accumulated against the same reference;
//   * the S buffer of a slot is released (s_free) as soon as its 128 columns sit in registers, so
//     the tensor pipe computes S(j+1) WHILE the softmax of tile j runs (also across steps);
//   * head dim 64 makes the kernel exp-bound, not MMA-bound: 16 MUFU.EX2 / clk / SM against 8192
//     tensor FLOP / clk / SM is 1024 vs 512 clk per 128x128 tile.  The scale/subtract and the row
//     sum use packed FFMA2 / FADD2 (two fp32 lanes per issue slot) and one pair of exponentials in
//     every kPolyEvery pairs is evaluated on the FMA pipe (Cody-Waite split + degree-3 minimax
//     polynomial) to take load off the MUFU;
//   * a partial last K/V tile (cross-attention: Nk = 77) only pays for ceil(valid/32) softmax
//     chunks and ceil(valid/16) P V k-steps.
//
// Replaces the einsum -> softmax -> einsum of fusion_generation/utils_custom.py:91-105 and
// utils_lora.py:99-113 (which materialise the [B*h, N, N] score tensor).
#include "tmx_common.cuh"
#include <cuda.h>

namespace tmx {

constexpr int kD = 64;               // head dim
constexpr int kBM = 128;             // query rows per tile  (UMMA M)
constexpr int kBN = 128;             // kv rows per tile     (UMMA N of QK^T, K of PV)
constexpr int kTileBytes = kBM * kD * 2;         // 16 KiB, one 128x64 16-bit tile
constexpr float kRescaleThreshold = 8.0f;        // log2 domain
#ifndef TMX_ATTN_POLY_EVERY
#define TMX_ATTN_POLY_EVERY 3
#endif
constexpr int kPolyEvery = TMX_ATTN_POLY_EVERY;  // one PAIR of exponentials in every kPolyEvery pairs goes to the FMA pipe (0 = all MUFU)

__device__ unsigned int g_attn_timeout_flag = 0;

// Optional event trace of CTA 0 (debug builds only: -DTMX_ATTN_TRACE): role r in {0 TMA, 1 MMA slot 0, 2 softmax
// slot 0 / half 0 / quarter 0} appends (tag, clock64) pairs; read back with tmx_attn_debug_trace().
#ifdef TMX_ATTN_TRACE
constexpr int kTraceLen = 4096;
__device__ long long g_attn_trace[4][kTraceLen];
#define TMX_TRACE_DECL(role, cond) const bool trace_on_ = (cond) && blockIdx.x == 0; int trace_n_ = 0; constexpr int trace_role_ = (role);
#define TMX_TRACE(tag) do { if (trace_on_ && trace_n_ + 2 <= kTraceLen) { g_attn_trace[trace_role_][trace_n_++] = (tag); g_attn_trace[trace_role_][trace_n_++] = clock64(); } } while (0)
#else
#define TMX_TRACE_DECL(role, cond)
#define TMX_TRACE(tag) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// All mbarrier helpers take the 32-bit shared-window address (computed ONCE per kernel from the generic
// pointer): a generic->shared conversion per call costs an S2UR + ULEA chain in every hot loop.
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
// Bounded wait: a protocol bug must end in a trap (CUDA error), never in a hung GPU.  try_wait suspends
// the warp in hardware for a while, so the loop body is rarely executed; the clock is only read
// every 256 failed polls to keep the spinning warps off the issue ports of the softmax warps.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++polls & 255u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) {          // ~2 s
                atomicExch(&g_attn_timeout_flag, 1u);
                __trap();
            }
        }
    }
}
// One lane of a CONVERGED warp (cute::elect_one_sync): keeps the operands of the guarded tcgen05 / TMA
// instructions warp-uniform for the compiler, so they live in uniform registers instead of being
// funnelled through a per-instruction R2UR "uniformisation" loop (measured: ~170 clk per tcgen05.mma).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :: "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar),
           "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(slot), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// Same, with descriptors given as their LOW words (start address + LBO); the high word (SBO = 1024 B,
// version 1, SWIZZLE_128B) is the same for every operand tile of this kernel.  Keeping the per-MMA
// arithmetic to one add on a precomputed low word matters: the issuing thread works in the slow uniform
// datapath, and a shift/mask/or chain per descriptor cost ~70 clk per tcgen05.mma.
constexpr uint32_t kDescHi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | ((16u >> 4) << 16); }
template <bool ACC>
__device__ __forceinline__ void umma_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "n"(ACC ? 1 : 0), "r"(kDescHi) : "memory");
}
__device__ __forceinline__ void umma_ts_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
        :: "r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi) : "memory");
}
__device__ __forceinline__ void umma_commit_addr(uint32_t bar_addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar_addr) : "memory");
}

#define TMX_R8(r, o)  "=r"(r[o+0]), "=r"(r[o+1]), "=r"(r[o+2]), "=r"(r[o+3]), "=r"(r[o+4]), "=r"(r[o+5]), "=r"(r[o+6]), "=r"(r[o+7])
#define TMX_I8(r, o)  "r"(r[o+0]), "r"(r[o+1]), "r"(r[o+2]), "r"(r[o+3]), "r"(r[o+4]), "r"(r[o+5]), "r"(r[o+6]), "r"(r[o+7])
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets lane (base_lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : TMX_R8(r, 0), TMX_R8(r, 8), TMX_R8(r, 16), TMX_R8(r, 24) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr), TMX_I8(r, 0), TMX_I8(r, 8), TMX_I8(r, 16), TMX_I8(r, 24) : "memory");
}
__device__ __forceinline__ void tmem_st64(uint32_t taddr, const uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, "
        "%33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, "
        "%49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, %64};"
        :: "r"(taddr), TMX_I8(r, 0), TMX_I8(r, 8), TMX_I8(r, 16), TMX_I8(r, 24), TMX_I8(r, 32), TMX_I8(r, 40), TMX_I8(r, 48), TMX_I8(r, 56)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :: "r"(taddr), TMX_I8(r, 0), TMX_I8(r, 8) : "memory");
}

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 process two fp32 lanes per issue slot) ----
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// 2^x for a PAIR on the FMA/ALU pipes (no MUFU): x = n + f with n = round(x), f in [-0.5, 0.5];
// 2^f by a degree-3 minimax polynomial (max rel. error 7.5e-5 << the 2^-9 / 2^-12 rounding of P),
// 2^n by adding n to the exponent field.
__device__ __forceinline__ void ex2_poly2(uint64_t x2, float& e0, float& e1) {
    float x0, x1;
    upk2(x2, x0, x1);
    x0 = fmaxf(x0, -126.f);
    x1 = fmaxf(x1, -126.f);
    x2 = pk2(x0, x1);
    const uint64_t xr = fadd2(x2, pk2(12582912.f, 12582912.f));          // 1.5 * 2^23: low mantissa bits hold round(x)
    const uint64_t t = fadd2(xr, pk2(-12582912.f, -12582912.f));         // round(x) as a float
    const uint64_t f = ffma2(t, pk2(-1.f, -1.f), x2);                    // x - round(x)
    uint64_t p = ffma2(f, pk2(0.05517132208f, 0.05517132208f), pk2(0.24261054397f, 0.24261054397f));
    p = ffma2(p, f, pk2(0.69326096773f, 0.69326096773f));
    p = ffma2(p, f, pk2(0.99992811680f, 0.99992811680f));
    float p0, p1, r0, r1;
    upk2(p, p0, p1);
    upk2(xr, r0, r1);
    e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23));
    e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23));
}

// ------------------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14) | LBO>>4 [16,30) |
// SBO>>4 [32,46) | version=1 [46,48) | layout_type [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16, fp32 accumulate.
__host__ __device__ constexpr uint32_t make_idesc(bool a_bf16, bool b_bf16, int M, int N, bool b_mn_major) {
    return (1u << 4)                              // c_format = F32
         | ((a_bf16 ? 1u : 0u) << 7)              // a_format (0 = F16, 1 = BF16)
         | ((b_bf16 ? 1u : 0u) << 10)             // b_format
         | (0u << 15)                             // a_major = K
         | ((b_mn_major ? 1u : 0u) << 16)         // b_major
         | ((uint32_t)(N >> 3) << 17)             // n_dim
         | ((uint32_t)(M >> 4) << 24);            // m_dim
}

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    if constexpr (BF16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    } else {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
}

// ------------------------------------------------------------------------------------- kernel
constexpr int kSlots = 2;            // query tiles processed together by one CTA
constexpr int kStages = 3;           // K/V ring depth
constexpr int kQBufs = 2;            // Q double buffer (per slot)
// HV = softmax threads per query row.  Warps [0, 8*HV) softmax: slot = (warp >> 2) & 1, half = warp >> 3; then
// 4 utility warps: TMA producer, S = Q K^T issuer, two O += P V issuers (one per slot).  The utility warps sit
// at the highest warp ids and give most of their registers to the softmax warps (setmaxnreg).
template <int HV> struct AttnCfg {
    static constexpr int kUtilWarp = 8 * HV;
    static constexpr int kThreads = 32 * (kUtilWarp + 4);
    static constexpr int kRegsUtil = HV == 2 ? 32 : 56;            // 128*32 + 512*112 = 640*96 ; 128*56 + 256*224 = 384*168
    static constexpr int kRegsSoftmax = HV == 2 ? 112 : 224;
};
constexpr int kTmemCols = 512;
constexpr int kBars = 6 * kSlots + 2 * kQBufs * kSlots + 4 * kStages;
constexpr int kSmemBytes = 1024 /*align slack*/ + (kQBufs * kSlots + 2 * kStages + kSlots /*O staging*/) * kTileBytes + kBars * 8 + 16
                           + 2 * kSlots * 2 * kBM * 4 /*l_xchg*/ + 2 * kSlots * 2 * kBM * 4 /*m_xchg (HV == 2 with -DTMX_ATTN_HV2_XCHG)*/;

// One scheduling step of a CTA: K/V tiles [j0, j1) of the unit that begins at schedule index `start` (nslots query tiles).
// A step with j0 == 0 && j1 == T is a whole unit; j0 > 0 is a partial piece handed to the unit's owner through the workspace;
// j0 == 0 && j1 < T is the owner's head piece.  The (b, h, query tile) coordinates are decoded from `start` only where they
// are needed (TMA loads / the one thread that issues the store): the softmax warps have no registers to spare for them.
struct Step { int start, nslots, j0, j1; };

// Walk the CTA's range [it, end).  split == 1: linear (unit, K/V tile) iterations, UPP = units per (b, h), TPU = query tiles per
// unit.  split == 0: linear query-tile ids (whole tiles only); pairs are formed greedily inside the range, a lone tile at a range
// or head boundary runs in slot 0 alone.
__device__ __forceinline__ bool next_step(int& it, int end, int split, int T, int QT, int UPP, int TPU, Step& s) {
    if (it >= end) return false;
    s.start = it;
    if (split) {
        const int u = it / T;
        s.j0 = it - u * T;
        const int left = end - it;
        s.j1 = left < T - s.j0 ? s.j0 + left : T;
        const int qt = (u % UPP) * TPU;
        s.nslots = (TPU == 2 && qt + 1 < QT) ? 2 : 1;
        it += s.j1 - s.j0;
    } else {
        s.j0 = 0;
        s.j1 = T;
        s.nslots = (TPU == 2 && it + 1 < end && it % QT + 1 < QT) ? 2 : 1;
        it += s.nslots;
    }
    return true;
}
__device__ __forceinline__ void step_coords(int start, int split, int T, int QT, int UPP, int TPU, int H, int& b, int& h, int& qt) {
    int bh;
    if (split) {
        const int u = start / T;
        bh = u / UPP;
        qt = (u - bh * UPP) * TPU;
    } else {
        bh = start / QT;
        qt = start - bh * QT;
    }
    b = bh / H;
    h = bh - b * H;
}

// Split-unit workspace, per CTA and slot: O[kD][kBM] fp32 (column-major: a warp writes 32 consecutive rows of one
// column), then m[kBM] (reference max, log2 domain) and l[kBM] (row sum against that reference).
constexpr int kWsFloatsPerSlot = (kD + 2) * kBM;

__device__ __forceinline__ void flag_wait(const unsigned int* f) {
    unsigned int v;
    long long t0 = 0;
    unsigned int polls = 0;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
        if (v != 0) return;
        if ((++polls & 255u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) { atomicExch(&g_attn_timeout_flag, 2u); __trap(); }
        }
    }
}

// Cold paths of the split-unit protocol, kept out of line so that they do not add to the register pressure of the softmax loop.
// Hand a partial piece to the unit's owner: (O, m, l) of this thread's row, un-normalised, into the CTA's workspace slot.
__device__ __noinline__ void split_dump(float* wsp, unsigned int* flag, uint32_t t_o, int row, float m_ref, float l_sum, int bar_id, bool signal) {
    wsp[kD * kBM + row] = m_ref;
    wsp[(kD + 1) * kBM + row] = l_sum;
#pragma unroll
    for (int g = 0; g < kD / 32; ++g) {
        uint32_t o[32];
        tmem_ld32(t_o + g * 32, o);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; ++c) wsp[(g * 32 + c) * kBM + row] = __uint_as_float(o[c]);
    }
#ifndef TMX_SPLIT_NOFENCE
    __threadfence();
#endif
    asm volatile("bar.sync %0, 128;" :: "r"(bar_id) : "memory");
    if (signal) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(flag), "r"(1u) : "memory");
}
// Owner of a unit whose tail went to the following CTA(s): fold their partials into the owner's own (O in TMEM at t_o, m_ref,
// l_sum), one piece after the other in CTA order, each piece fetched as ONE batch of independent L2 loads (m, l and the 64 O
// columns of the row), then normalise and write the row into the swizzled staging tile (o_row) like the plain epilogue does.
// Returns the index behind the last piece.
template <bool BF16>
__device__ __noinline__ int split_merge(const float* ws_slot0, const unsigned int* flags_slot0, int first, int unit_end, int total_iters,
                                        uint32_t t_o, int row, int lane, float m_run, float l_run, uint32_t o_row) {
    float o[kD];
    {
        uint32_t r[32];
#pragma unroll
        for (int g = 0; g < kD / 32; ++g) {
            tmem_ld32(t_o + g * 32, r);
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) o[g * 32 + c] = __uint_as_float(r[c]);
        }
    }
    int k = first;
    for (; k < (int)gridDim.x && (int)(((long long)k * total_iters) / gridDim.x) < unit_end; ++k) {
#ifndef TMX_SPLIT_NOWAIT
        if (lane == 0) flag_wait(flags_slot0 + k * kSlots);
        __syncwarp();
#endif
        const float* wk = ws_slot0 + (size_t)k * kSlots * kWsFloatsPerSlot + row;
        float v[kD];
#pragma unroll
        for (int c = 0; c < kD; ++c) v[c] = __ldcg(wk + c * kBM);
        const float mk = __ldcg(wk + kD * kBM), lk = __ldcg(wk + (kD + 1) * kBM);
        const float m_new = fmaxf(m_run, mk);
        const float a_own = ex2(m_run - m_new), a_k = ex2(mk - m_new);
        l_run = l_run * a_own + lk * a_k;
        m_run = m_new;
#pragma unroll
        for (int c = 0; c < kD; ++c) o[c] = fmaf(v[c], a_k, o[c] * a_own);
    }
    const float inv_l = 1.f / l_run;
    const uint32_t o_sw = (uint32_t)(row & 7);
#pragma unroll
    for (int c = 0; c < kD; c += 8) {
        uint4 q;
        q.x = pack2<BF16>(o[c] * inv_l, o[c + 1] * inv_l);
        q.y = pack2<BF16>(o[c + 2] * inv_l, o[c + 3] * inv_l);
        q.z = pack2<BF16>(o[c + 4] * inv_l, o[c + 5] * inv_l);
        q.w = pack2<BF16>(o[c + 6] * inv_l, o[c + 7] * inv_l);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(o_row + ((((uint32_t)c >> 3) ^ o_sw) << 4)), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
    }
    return k;
}

template <bool BF16, int HV>
__global__ void __launch_bounds__(AttnCfg<HV>::kThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                int Nq, int Nk, int H, int QT, int UPP, int TPU, int total_iters, int split,
                float* __restrict__ ws, unsigned int* __restrict__ ws_flags, float scale_log2) {
    using Cfg = AttnCfg<HV>;
    constexpr int ST = kStages;
    constexpr int kUtilWarp = Cfg::kUtilWarp;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                                   // [kQBufs][kSlots] tiles
    uint8_t* sK = sQ + kQBufs * kSlots * kTileBytes;      // [ST]
    uint8_t* sV = sK + ST * kTileBytes;                   // [ST]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + (ST + kSlots) * kTileBytes);
    uint32_t sQ_a = smem_u32(sQ);
    asm volatile("mov.u32 %0, %0;" : "+r"(sQ_a));         // opaque: keep it in a register instead of re-deriving it (S2UR chain) at every use
    const uint32_t sK_a = sQ_a + kQBufs * kSlots * kTileBytes, sV_a = sK_a + ST * kTileBytes;
    const uint32_t sO_a = sV_a + ST * kTileBytes;         // [kSlots] output staging tiles (128 rows x 128 B, SWIZZLE_128B) for the TMA store
    const uint32_t q_full = sO_a + kSlots * kTileBytes;   // [kQBufs*kSlots]  TMA -> MMA : Q tile landed
    const uint32_t q_empty = q_full + 8 * kQBufs * kSlots; // [kQBufs*kSlots]  MMA -> TMA : last S = Q K^T of the step issued
    const uint32_t s_full = q_empty + 8 * kQBufs * kSlots; // [kSlots]  MMA -> softmax   : S(n) complete
    const uint32_t s_free = s_full + 8 * kSlots;           // [kSlots]  softmax -> MMA   : S(n) copied to registers
    const uint32_t p_full = s_free + 8 * kSlots;           // [kSlots]  softmax -> MMA   : P(n) written to TMEM
    const uint32_t pv_done = p_full + 8 * kSlots;          // [kSlots]  MMA -> softmax   : O += P(n) V(n) complete
    const uint32_t k_full = pv_done + 8 * kSlots;          // [ST]
    const uint32_t k_empty = k_full + 8 * ST;
    const uint32_t v_full = k_empty + 8 * ST;
    const uint32_t v_empty = v_full + 8 * ST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kBars);
    float* l_xchg = reinterpret_cast<float*>(tmem_slot + 4);   // [2 step parities][kSlots][2 halves][kBM] partial row sums
    float* m_xchg = l_xchg + 2 * kSlots * 2 * kBM;             // [2 tile parities][kSlots][2 halves][kBM] half-row maxima (HV2_XCHG)
    (void)m_xchg;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int T = (Nk + kBN - 1) / kBN;                   // K/V tiles per step
    const int last_valid = Nk - (T - 1) * kBN;            // valid kv rows of the last tile (1..128)
    // this CTA's equal share of the schedule space (split: (unit, K/V tile) iterations, else query tiles); total_iters is its size
    const int tile_begin = (int)(((long long)blockIdx.x * total_iters) / gridDim.x);
    const int tile_end = (int)(((long long)(blockIdx.x + 1) * total_iters) / gridDim.x);

    if (warp == kUtilWarp + 1 && lane == 0) {
        for (int i = 0; i < kQBufs * kSlots; ++i) { mbar_init(q_full + 8u * (i), 1); mbar_init(q_empty + 8u * (i), 1); }
        for (int i = 0; i < kSlots; ++i) { mbar_init(s_full + 8u * (i), 1); mbar_init(s_free + 8u * (i), 4 * HV); mbar_init(p_full + 8u * (i), 4 * HV); /* one arrival per softmax warp of the slot */ mbar_init(pv_done + 8u * (i), 1); }
        for (int i = 0; i < ST; ++i) { mbar_init(k_full + 8u * (i), 1); mbar_init(k_empty + 8u * (i), 1); mbar_init(v_full + 8u * (i), 1); mbar_init(v_empty + 8u * (i), kSlots); }
        fence_barrier_init();
    }
    if (warp == kUtilWarp) {
        if (lane == 0) { tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_o); }
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp >= kUtilWarp) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(Cfg::kRegsUtil));
        if (warp == kUtilWarp) {
            // ================================ TMA producer ================================
            // the whole warp walks the schedule (uniform control flow); one elected lane issues
            {
                int it = tile_begin;
                Step cur;
                uint32_t sidx = 0, qbits = 0;                // per (buffer, slot) use parity
                int st = 0;
                uint32_t ph = 0;
                TMX_TRACE_DECL(0, lane == 0)
                while (next_step(it, tile_end, split, T, QT, UPP, TPU, cur)) {
                    const int buf = (int)(sidx & 1u);
                    int cb, ch, cqt;
                    step_coords(cur.start, split, T, QT, UPP, TPU, H, cb, ch, cqt);
                    for (int w = 0; w < cur.nslots; ++w) {
                        const int qi = buf * kSlots + w;
                        mbar_wait(q_empty + 8u * (qi), ((qbits >> qi) & 1u) ^ 1u);
                        qbits ^= 1u << qi;
                        if (elect_one()) {
                            mbar_expect_tx(q_full + 8u * (qi), kTileBytes);
                            tma_load_4d(sQ_a + qi * kTileBytes, &tm_q, q_full + 8u * (qi), 0, ch, (cqt + w) * kBM, cb);
                        }
                    }
                    for (int j = cur.j0; j < cur.j1; ++j) {
                        mbar_wait(k_empty + 8u * (st), ph ^ 1u);
                        TMX_TRACE(0);
                        if (elect_one()) {
                            mbar_expect_tx(k_full + 8u * (st), kTileBytes);
                            tma_load_4d(sK_a + st * kTileBytes, &tm_k, k_full + 8u * (st), 0, ch, j * kBN, cb);
                        }
                        mbar_wait(v_empty + 8u * (st), ph ^ 1u);
                        TMX_TRACE(1);
                        if (elect_one()) {
                            mbar_expect_tx(v_full + 8u * (st), kTileBytes);
                            tma_load_4d(sV_a + st * kTileBytes, &tm_v, v_full + 8u * (st), 0, ch, j * kBN, cb);
                        }
                        if (++st == ST) { st = 0; ph ^= 1u; }
                    }
                    ++sidx;
                }
            }
        } else if (warp == kUtilWarp + 1) {
            // ================================ S = Q K^T issuer (both slots) ===============
            // Walks the CTA's K/V stream tile by tile and, per tile, issues S = Q K^T for every active
            // slot as soon as the slot's S buffer is free.  It is not tied to the P V issue, so it runs
            // ahead of the softmax as far as the barriers allow (also across steps).
            constexpr uint32_t idesc_qk = make_idesc(BF16, BF16, kBM, kBN, false);
            const uint32_t q_lo0 = desc_lo(sQ_a), k_lo0 = desc_lo(sK_a);
            int it = tile_begin;
            Step cur;
            uint32_t nS0 = 0, nS1 = 0, sidx = 0, qbits = 0;   // qbits: per (buffer, slot) use parity
            int st = 0;
            uint32_t ph = 0;
            TMX_TRACE_DECL(1, lane == 0)
            while (next_step(it, tile_end, split, T, QT, UPP, TPU, cur)) {
                const int buf = (int)(sidx & 1u);
                const bool two = cur.nslots == 2;
                {
                    mbar_wait(q_full + 8u * (buf * kSlots), (qbits >> (buf * kSlots)) & 1u);
                    qbits ^= 1u << (buf * kSlots);
                    if (two) { mbar_wait(q_full + 8u * (buf * kSlots + 1), (qbits >> (buf * kSlots + 1)) & 1u); qbits ^= 1u << (buf * kSlots + 1); }
                }
                const uint32_t a_lo0 = q_lo0 + (uint32_t)(buf * kSlots) * (kTileBytes >> 4);
                for (int j = cur.j0; j < cur.j1; ++j) {
                    const uint32_t b_lo = k_lo0 + (uint32_t)st * (kTileBytes >> 4);
                    mbar_wait(k_full + 8u * (st), ph);
                    TMX_TRACE(10);
                    if (nS0 > 0) mbar_wait(s_free + 8u * (0), (nS0 - 1u) & 1u);
                    TMX_TRACE(11);
                    tc_fence_after();
                    if (elect_one()) {
                        umma_ss_lo<false>(tmem_base, a_lo0, b_lo, idesc_qk);           // 4 x (K = 16): +32 B inside the 128 B swizzle row
                        umma_ss_lo<true>(tmem_base, a_lo0 + 2, b_lo + 2, idesc_qk);
                        umma_ss_lo<true>(tmem_base, a_lo0 + 4, b_lo + 4, idesc_qk);
                        umma_ss_lo<true>(tmem_base, a_lo0 + 6, b_lo + 6, idesc_qk);
                        umma_commit(s_full + 8u * (0));
                        if (!two) umma_commit(k_empty + 8u * (st));
                    }
                    TMX_TRACE(12);
                    ++nS0;
                    if (two) {
                        if (nS1 > 0) mbar_wait(s_free + 8u * (1), (nS1 - 1u) & 1u);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t a_lo1 = a_lo0 + (kTileBytes >> 4);
                            umma_ss_lo<false>(tmem_base + 256, a_lo1, b_lo, idesc_qk);
                            umma_ss_lo<true>(tmem_base + 256, a_lo1 + 2, b_lo + 2, idesc_qk);
                            umma_ss_lo<true>(tmem_base + 256, a_lo1 + 4, b_lo + 4, idesc_qk);
                            umma_ss_lo<true>(tmem_base + 256, a_lo1 + 6, b_lo + 6, idesc_qk);
                            umma_commit(s_full + 8u * (1));
                            umma_commit(k_empty + 8u * (st));
                        }
                        ++nS1;
                    }
                    if (++st == ST) { st = 0; ph ^= 1u; }
                }
                if (elect_one()) {                         // every S = Q K^T of this step has been issued: Q buffers may be refilled
                    umma_commit(q_empty + 8u * (buf * kSlots));
                    if (two) umma_commit(q_empty + 8u * (buf * kSlots + 1));
                }
                ++sidx;
            }
        } else {
            // ================================ O += P V issuer of slot w ===================
            const int w = warp - kUtilWarp - 2;
            constexpr uint32_t idesc_pv = make_idesc(BF16, BF16, kBM, kD, true);
            const uint32_t t_o = tmem_base + w * 256 + 128, t_p = t_o + 64;
            const int last_ksteps = (last_valid + 15) >> 4;
            const uint32_t v_lo0 = desc_lo(sV_a);
            int it = tile_begin;
            Step cur;
            uint32_t nPV = 0;
            int st = 0;
            uint32_t ph = 0;
            TMX_TRACE_DECL(3, w == 0 && lane == 0)
            while (next_step(it, tile_end, split, T, QT, UPP, TPU, cur)) {
                const bool act = w < cur.nslots;
                for (int j = cur.j0; j < cur.j1; ++j) {
                    const uint32_t b_lo = v_lo0 + (uint32_t)st * (kTileBytes >> 4);
                    mbar_wait(v_full + 8u * (st), ph);
                    TMX_TRACE(13);
                    if (act) {
                        mbar_wait(p_full + 8u * (w), nPV & 1u);
                        TMX_TRACE(14);
                        tc_fence_after();
                        if (elect_one()) {                                 // (K = 16 kv rows): +2048 B in V, +8 columns in P
                            if (j < T - 1 || last_ksteps == kBN / 16) {
                                umma_ts_lo(t_o, t_p, b_lo, idesc_pv, j > cur.j0 ? 1u : 0u);
#pragma unroll
                                for (int k = 1; k < kBN / 16; ++k) umma_ts_lo(t_o, t_p + k * 8, b_lo + k * 128, idesc_pv, 1u);
                            } else {
                                for (int k = 0; k < last_ksteps; ++k) umma_ts_lo(t_o, t_p + k * 8, b_lo + k * 128, idesc_pv, (j > cur.j0 || k > 0) ? 1u : 0u);
                            }
                            umma_commit(pv_done + 8u * (w));
                            umma_commit(v_empty + 8u * (st));
                        }
                        TMX_TRACE(15);
                        ++nPV;
                    } else if (elect_one()) {
                        mbar_arrive(v_empty + 8u * (st));
                    }
                    if (++st == ST) { st = 0; ph ^= 1u; }
                }
            }
        }
    } else {
        // ================================ softmax / correction / epilogue ==============
        // HV threads per query row (TMEM lane).  HV == 1: the thread owns the whole 128-column row of
        // every S tile.  HV == 2: `half` 0 exponentiates columns [0,64), half 1 columns [64,128); both
        // read the whole row for the running max (bit-identical in the two threads, so they take the
        // same lazy-rescale decisions without talking to each other); each keeps a partial row sum
        // and owns 32 of the 64 O columns; the sums meet once per step.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(Cfg::kRegsSoftmax));
        constexpr int COLS = kBN / HV;                       // S columns exponentiated by this thread
        constexpr int NCH = COLS / 32;                       // ... in chunks of 32 (16 packed P words)
        constexpr int OC = kD / HV;                          // O columns owned by this thread
        const int w = (warp >> 2) & 1;                       // slot handled by this warp
        const int half = HV == 2 ? (warp >> 3) : 0;          // column half of the S tile
        const int quarter = warp & 3;                        // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;                 // row in the tile == TMEM lane
        const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + w * 256;
        const uint32_t t_own = t_row + half * COLS;
        const uint64_t sc2 = pk2(scale_log2, scale_log2);
        int last_chunks = (last_valid - half * COLS + 31) >> 5;          // own chunks of the last K/V tile
        last_chunks = last_chunks < 0 ? 0 : (last_chunks > NCH ? NCH : last_chunks);
        const uint32_t bar_s_full = s_full + 8u * w, bar_s_free = s_free + 8u * w, bar_p_full = p_full + 8u * w, bar_pv_done = pv_done + 8u * w;
        uint32_t n = 0;                                      // tiles processed by this slot (barrier phases)
        uint32_t nstep = 0;
        TMX_TRACE_DECL(2, warp == 0 && lane == 0)
        int it = tile_begin;
        Step cur;
        while (next_step(it, tile_end, split, T, QT, UPP, TPU, cur)) {
            if (w >= cur.nslots) continue;
            float m_ref = -INFINITY;                         // running reference max, log2 domain (scaled)
            float l_sum = 0.f;                               // (partial) row sum over this thread's columns
            for (int j = cur.j0; j < cur.j1; ++j, ++n) {
                const bool last = (j == T - 1);
                const bool partial = last && last_valid < kBN;
                mbar_wait(bar_s_full, n & 1u);
                TMX_TRACE(20);
                tc_fence_after();
                uint32_t s[COLS];
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#if !defined(TMX_ATTN_EXPERIMENT_NOMAX) && !defined(TMX_ATTN_HV2_XCHG)
                if constexpr (HV == 2) {
                    // the other half of the row: only its maximum is needed
                    const uint32_t t_other = t_row + (half ^ 1) * COLS;
#pragma unroll
                    for (int q = 0; q < NCH; ++q) tmem_ld32(t_other + q * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[q * 32]));
                    tc_wait_ld();
                    if (partial) {
                        const int v = last_valid - (half ^ 1) * COLS;
#pragma unroll
                        for (int c = 0; c < COLS; ++c) if (c >= v) s[c] = 0xff800000u;   // -inf
                    }
#pragma unroll
                    for (int c = 0; c < COLS; c += 4) {
                        mx0 = fmaxf(mx0, __uint_as_float(s[c])); mx1 = fmaxf(mx1, __uint_as_float(s[c + 1]));
                        mx2 = fmaxf(mx2, __uint_as_float(s[c + 2])); mx3 = fmaxf(mx3, __uint_as_float(s[c + 3]));
                    }
                }
#endif
                // this thread's columns
#pragma unroll
                for (int q = 0; q < NCH; ++q) tmem_ld32(t_own + q * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[q * 32]));
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_s_free);                    // S(n) is consumed: the MMA warp may overwrite it
                TMX_TRACE(21);
                if (partial) {
                    const int v = last_valid - half * COLS;
#pragma unroll
                    for (int c = 0; c < COLS; ++c) if (c >= v) s[c] = 0xff800000u;       // -inf
                }
#pragma unroll
                for (int c = 0; c < COLS; c += 4) {
                    mx0 = fmaxf(mx0, __uint_as_float(s[c])); mx1 = fmaxf(mx1, __uint_as_float(s[c + 1]));
                    mx2 = fmaxf(mx2, __uint_as_float(s[c + 2])); mx3 = fmaxf(mx3, __uint_as_float(s[c + 3]));
                }
                float row_max = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
#ifdef TMX_ATTN_HV2_XCHG
                // (NOT YET RUN ON A GPU — opt-in build.)  HV == 2 without the redundant read of the other half: the two threads
                // of a row swap their half-row maxima through shared memory.  Buffers alternate with the tile parity, so ONE
                // 64-thread named barrier per tile orders both the hand-over and the reuse two tiles later.
                if constexpr (HV == 2) {
                    float* mxb = m_xchg + ((n & 1u) * kSlots + w) * 2 * kBM;
                    mxb[half * kBM + row] = row_max;
                    asm volatile("bar.sync %0, 64;" :: "r"(3 + w * 4 + quarter) : "memory");
                    row_max = fmaxf(row_max, mxb[(half ^ 1) * kBM + row]);
                }
#endif
                const float m_tile = row_max * scale_log2;
                // lazy rescale: move the reference only when the max grew by more than 2^8
                const bool bump = m_tile > m_ref + kRescaleThreshold;
                const float m_new = bump ? m_tile : m_ref;
                bool pv_waited = false;
                if (j > cur.j0 && __any_sync(0xffffffffu, bump)) {            // warp-uniform: tcgen05.ld/st are warp-collective
                    mbar_wait(bar_pv_done, (n - 1u) & 1u);               // O must hold every P V issued so far
                    tc_fence_after();
                    pv_waited = true;
                    const float alpha = bump ? ex2(m_ref - m_new) : 1.f;
                    l_sum *= alpha;
#pragma unroll
                    for (int g = 0; g < OC / 32; ++g) {
                        uint32_t o[32];
                        tmem_ld32(t_row + 128 + half * OC + g * 32, o);
                        tc_wait_ld();
#pragma unroll
                        for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
                        tmem_st32(t_row + 128 + half * OC + g * 32, o);
                    }
                }
                m_ref = m_new;
                const int chunks = last ? last_chunks : NCH;
                const uint64_t nm2 = pk2(-m_new, -m_new);
                uint64_t sum_a = pk2(0.f, 0.f), sum_b = pk2(0.f, 0.f);
                // exponentials of one pair of columns (MUFU, or the FMA-pipe polynomial for one pair in kPolyEvery)
                auto exp_pair = [&](int col, int c, uint32_t& packed) {
                    const uint64_t x2 = ffma2(pk2(__uint_as_float(s[col]), __uint_as_float(s[col + 1])), sc2, nm2);
                    float e0, e1;
                    if (kPolyEvery > 0 && (c % (kPolyEvery > 0 ? kPolyEvery : 1)) == kPolyEvery - 1) {
                        ex2_poly2(x2, e0, e1);
                    } else {
                        float x0, x1;
                        upk2(x2, x0, x1);
#ifdef TMX_ATTN_EXPERIMENT_NOEXP
                        e0 = x0; e1 = x1;
#else
                        e0 = ex2(x0);
                        e1 = ex2(x1);
#endif
                    }
                    if (c & 1) sum_b = fadd2(sum_b, pk2(e0, e1)); else sum_a = fadd2(sum_a, pk2(e0, e1));
                    packed = pack2<BF16>(e0, e1);
                };
                // P(n-1) must have been consumed by P V(n-1) before it is overwritten.  The event trace (profiles/r02z_*) shows 300-500 clk
                // of this wait per tile; moving it behind the exp block, right in front of the store it protects, was measured SLOWER
                // (227 vs 212 us at N = 4096: the poll loop in the middle of the block breaks up its MUFU / convert / store schedule).
                if (!pv_waited && j > cur.j0) {
                    mbar_wait(bar_pv_done, (n - 1u) & 1u);
                    tc_fence_after();
                }
                TMX_TRACE(23);
                if (chunks == NCH) {
                    // full tile: ONE straight-line block over all columns (no per-chunk branches or stores in between, so
                    // the tail of one chunk's MUFU / convert chain overlaps the head of the next), one wide P store
                    uint32_t p[COLS / 2];
#pragma unroll
                    for (int c = 0; c < COLS / 2; ++c) exp_pair(2 * c, c, p[c]);
                    TMX_TRACE(22);
                    if constexpr (HV == 1) tmem_st64(t_row + 192, p);
                    else tmem_st32(t_row + 192 + half * (COLS / 2), p);
                } else {
                    // partial last K/V tile (cross-attention, Nk = 77): only the chunks that hold valid columns
#pragma unroll
                    for (int q = 0; q < NCH; ++q) {
                        if (q < chunks) {
                            uint32_t p[16];
#pragma unroll
                            for (int c = 0; c < 16; ++c) exp_pair(q * 32 + 2 * c, c, p[c]);
                            tmem_st16(t_row + 192 + half * (COLS / 2) + q * 16, p);
                        }
                    }
                }
                {
                    float a0, a1, b0, b1;
                    upk2(sum_a, a0, a1);
                    upk2(sum_b, b0, b1);
                    l_sum += (a0 + a1) + (b0 + b1);
                }
                TMX_TRACE(24);
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_p_full);
                TMX_TRACE(25);
            }

            // epilogue: (the two halves exchange their partial sums, then) O / l -> global, row-contiguous per thread
            if constexpr (HV == 2) {
                float* lx = l_xchg + ((nstep & 1u) * kSlots + w) * 2 * kBM;
                lx[half * kBM + row] = l_sum;
                asm volatile("bar.sync %0, 256;" :: "r"(1 + w) : "memory");
                l_sum += lx[(half ^ 1) * kBM + row];
                ++nstep;
            }
            mbar_wait(bar_pv_done, (n - 1u) & 1u);
            tc_fence_after();
            TMX_TRACE(30);
            if constexpr (HV == 1) {
                if (cur.j0 > 0) {
                    // ---- partial piece of a unit owned by an earlier CTA: leave (O, m, l) un-normalised in this CTA's workspace
                    split_dump(ws + ((size_t)blockIdx.x * kSlots + w) * kWsFloatsPerSlot, ws_flags + blockIdx.x * kSlots + w,
                               t_row + 128, row, m_ref, l_sum, 11 + w, quarter == 0 && lane == 0);
                    tc_fence_before();
                    continue;
                }
            }
            // O / l -> 16-bit -> this slot's staging tile (row r = 128 B, 16-byte chunks XOR-swizzled with r & 7) -> ONE TMA store of
            // the 128 x 64 tile: full-line bulk writes instead of 32 scattered 16-byte stores per warp instruction, rows beyond Nq
            // clipped by the tensor map.  Barrier A: the issuing thread has seen its previous store read the tile; barrier B: all rows written.
            const uint32_t o_row = sO_a + w * kTileBytes + row * 128;
            const uint32_t o_sw = (uint32_t)(row & 7);
            // (the wait for the PREVIOUS store's shared-memory read sits here, a whole step after it was issued, not behind the issue:
            // the event trace showed the issuing thread — and with it its whole slot — parked ~1700 clk per step on that wait)
            if (half == 0 && quarter == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync %0, %1;" :: "r"(11 + w), "n"(128 * HV) : "memory");
            TMX_TRACE(31);
            int k_end = blockIdx.x + 1;                     // pieces [blockIdx.x + 1, k_end) were folded into this unit
#ifndef TMX_SPLIT_NOMERGE
            if (HV == 1 && cur.j1 < T) {
                k_end = split_merge<BF16>(ws + w * kWsFloatsPerSlot, ws_flags + w, blockIdx.x + 1, it + (T - cur.j1) /* `it` already points behind this piece */,
                                          total_iters, t_row + 128, row, lane, m_ref, l_sum, o_row);
            } else
#endif
            {
                const float inv_l = 1.f / l_sum;
#pragma unroll
                for (int g = 0; g < OC / 32; ++g) {
                    uint32_t o[32];
                    tmem_ld32(t_row + 128 + half * OC + g * 32, o);
                    tc_wait_ld();
#pragma unroll
                    for (int c = 0; c < 32; c += 8) {
                        uint4 v;
                        v.x = pack2<BF16>(__uint_as_float(o[c]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
                        v.y = pack2<BF16>(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
                        v.z = pack2<BF16>(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
                        v.w = pack2<BF16>(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
                        const uint32_t chunk = (uint32_t)((half * OC + g * 32 + c) >> 3);
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(o_row + ((chunk ^ o_sw) << 4)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync %0, %1;" :: "r"(13 + w), "n"(128 * HV) : "memory");
            TMX_TRACE(32);
            if (half == 0 && quarter == 0 && lane == 0) {
                int cb, ch, cqt;
                step_coords(cur.start, split, T, QT, UPP, TPU, H, cb, ch, cqt);
                asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                             :: "l"(reinterpret_cast<uint64_t>(&tm_o)), "r"(sO_a + w * kTileBytes), "r"(0), "r"(ch), "r"((cqt + w) * kBM), "r"(cb) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                // every row of every piece has been read (barrier B): re-arm the pieces' flags for the next launch
#ifndef TMX_SPLIT_NOWAIT
                for (int k = blockIdx.x + 1; k < k_end; ++k) ws_flags[k * kSlots + w] = 0u;
#endif
            }
            TMX_TRACE(33);
            tc_fence_before();           // order the O reads before the next step's P(0) hand-off (p_full arrive)
        }
        if (half == 0 && quarter == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kUtilWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
int short_kv_attn_init();                      // cross_attention.cu
int short_kv_attn_launch(const void* q, const void* k, const void* v, void* o, int B, int H, int Nq, int Nk,
                         int64_t q_sn, int64_t k_sn, int64_t v_sn, int64_t o_sn, float scale, bool bf16, int mt, cudaStream_t st);
static float* g_ws[64] = {nullptr};            // per device: split-unit partials, sm_count x kSlots x kWsFloatsPerSlot floats
static unsigned int* g_ws_flags[64] = {nullptr};

int attn_init() {
    int dev = 0;
    TMX_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !g_ws[dev]) {
        // One workspace per device: launches on DIFFERENT streams of one process must not run this kernel concurrently
        // (the sampler is single-stream; separate processes have separate contexts and workspaces).
        const size_t n = (size_t)sm_count() * kSlots;
        TMX_CUDA(cudaMalloc(&g_ws[dev], n * kWsFloatsPerSlot * sizeof(float)));
        TMX_CUDA(cudaMalloc(&g_ws_flags[dev], n * sizeof(unsigned int)));
        TMX_CUDA(cudaMemset(g_ws_flags[dev], 0, n * sizeof(unsigned int)));
    }
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        TMX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        TMX_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, TMX_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
        g_encode = (EncodeTiledFn)fn;
    }
    if (int rc = short_kv_attn_init()) return rc;
    TMX_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    TMX_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    TMX_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    TMX_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    return TMX_OK;
}

// [B, N, H, 64] 16-bit tensor with token stride `stride_n` elements -> 4-D map (d, h, n, b), box 64 x 1 x 128 x 1
static int make_map(CUtensorMap* m, const void* base, int B, int N, int H, int64_t stride_n, bool bf16) {
    cuuint64_t dims[4] = {(cuuint64_t)kD, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)kD * 2, (cuuint64_t)stride_n * 2, (cuuint64_t)N * (cuuint64_t)stride_n * 2};
    cuuint32_t box[4] = {kD, 1, kBM, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return TMX_ECUDA; }
    return TMX_OK;
}

static int g_variant = 0;   // test hook: 0 / 2 = pair query tiles (default), 1 = one query tile per step
#ifndef TMX_ATTN_HALVES
#define TMX_ATTN_HALVES 1
#endif
static int g_halves = TMX_ATTN_HALVES;   // softmax threads per query row (1 or 2)
static int g_short = 3;                  // K/V sequences of <= 128 rows (cross-attention) go to the streaming kernel of cross_attention.cu with
                                         // g_short 16-row tiles per warp iteration (1, 2) or, with 3, to its tcgen05 kernel k2t; 0 = keep them on k1 (test hook: 30 .. 33)
static int g_split = 1;                  // stream-K split of units along K/V between CTAs: 0 never, 1 by the cost model, 2 always (test hook: 20 / 21 / 22)

}  // namespace tmx

using namespace tmx;

#ifdef TMX_ATTN_TRACE
extern "C" __attribute__((visibility("default"))) int tmx_attn_debug_trace(long long* host, int role) {
    return check_cuda(cudaMemcpyFromSymbol(host, g_attn_trace, sizeof(long long) * kTraceLen, sizeof(long long) * kTraceLen * role), "trace copy");
}
#endif

extern "C" int tmx_attn_set_variant(int nq) {
    // 0: defaults; 1: one query tile per step; 2: pairs; 11 / 12: 1 / 2 softmax threads per query row; 20 / 21 / 22: unit split never / by cost / always; 30 .. 33: short K/V on k1 / k2s / k2s 32 rows / k2t
    if (nq == 11 || nq == 12) { g_halves = nq - 10; return TMX_OK; }
    if (nq >= 20 && nq <= 22) { g_split = nq - 20; return TMX_OK; }
    if (nq >= 30 && nq <= 33) { g_short = nq - 30; return TMX_OK; }
    TMX_REQUIRE(nq >= 0 && nq <= 2, TMX_EINVAL, "attn_set_variant: nq must be 0, 1, 2, 11, 12, 20 .. 22 or 30 .. 33");
    g_variant = nq;
    if (nq == 0) { g_halves = TMX_ATTN_HALVES; g_split = 1; g_short = 3; }
    return TMX_OK;
}

extern "C" int tmx_attn_fwd(const void* q, const void* k, const void* v, void* o,
                            int B, int H, int Nq, int Nk, int D,
                            int64_t q_stride_n, int64_t k_stride_n, int64_t v_stride_n, int64_t o_stride_n,
                            float scale, int dtype, void* stream) {
    TMX_REQUIRE(q && k && v && o, TMX_EINVAL, "attn: null pointer");
    TMX_REQUIRE(B > 0 && H > 0 && Nq > 0 && Nk > 0, TMX_EINVAL, "attn: non-positive size");
    TMX_REQUIRE(D == kD, TMX_ESHAPE, "attn: head dim %d unsupported (only 64)", D);
    TMX_REQUIRE(dtype == TMX_F16 || dtype == TMX_BF16, TMX_EDTYPE, "attn: dtype %d unsupported (fp16/bf16 only)", dtype);
    TMX_REQUIRE(scale > 0.f, TMX_EINVAL, "attn: scale must be positive");
    const int64_t hd = (int64_t)H * kD;
    TMX_REQUIRE(q_stride_n >= hd && k_stride_n >= hd && v_stride_n >= hd && o_stride_n >= hd, TMX_ESHAPE,
                "attn: token stride smaller than H*64");
    TMX_REQUIRE(q_stride_n % 8 == 0 && k_stride_n % 8 == 0 && v_stride_n % 8 == 0 && o_stride_n % 8 == 0, TMX_EALIGN,
                "attn: token strides must be multiples of 8 elements");
    TMX_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), TMX_EALIGN, "attn: 16-byte alignment");
    const int QT = (Nq + kBM - 1) / kBM;
    const long long total = (long long)B * H * QT;
    TMX_REQUIRE(total * ((Nk + kBN - 1) / kBN) < (1ll << 30), TMX_ESHAPE, "attn: too many (query tile, K/V tile) iterations");
    if (int rc = require_init()) return rc;

    const bool bf16 = dtype == TMX_BF16;
    if (Nk <= 128 && g_short != 0 && g_variant == 0 && (long long)B * H <= 65535)
        return short_kv_attn_launch(q, k, v, o, B, H, Nq, Nk, q_stride_n, k_stride_n, v_stride_n, o_stride_n, scale, bf16, g_short, (cudaStream_t)stream);
    CUtensorMap mq, mk, mv, mo;
    if (int rc = make_map(&mo, o, B, Nq, H, o_stride_n, bf16)) return rc;
    if (int rc = make_map(&mq, q, B, Nq, H, q_stride_n, bf16)) return rc;
    if (int rc = make_map(&mk, k, B, Nk, H, k_stride_n, bf16)) return rc;
    if (int rc = make_map(&mv, v, B, Nk, H, v_stride_n, bf16)) return rc;
    const float scale_log2 = scale * 1.4426950408889634f;
    cudaStream_t st = (cudaStream_t)stream;
    // Two schedules.  (A) whole query tiles: each CTA an equal contiguous range of tile ids, pairs formed inside the range.
    // (C) stream-K: units = aligned tile pairs, the (unit, K/V tile) iteration space cut into equal shares, units that straddle
    // a share boundary merged through the workspace.  C keeps every SM equally busy whatever B*H*tiles is but pays one extra
    // step per CTA plus the merge; A is quantised to whole tiles.  Pick by a cost model in units of one pair-iteration
    // (calibrated on `tools/kbench.py --only attention --batch 1|2|4`, profiles/r02p_*): a step costs kStepCost on top of its
    // iterations, a lone tile kLoneCost of a pair, the merge kMergeCost.
    const int TPU = g_variant == 1 ? 1 : 2;
    const int UPP = (QT + TPU - 1) / TPU;
    const int T = (Nk + kBN - 1) / kBN;
    const long long units = (long long)B * H * UPP;
    const int sms = sm_count();
    int split = 0;
    if (g_halves == 1 && T > 1 && g_split != 0) {
        if (g_split == 2) split = 1;
        else {
            constexpr float kStepCost = 1.5f, kLoneCost = 0.65f, kMergeCost = 3.0f;
            const long long unitsA = TPU == 2 ? (total + 1) / 2 : total;
            const int gridA = (int)(unitsA < sms ? unitsA : sms);
            const int nt = (int)((total + gridA - 1) / gridA);                  // tiles of the busiest CTA
            const float costA = TPU == 2 ? (nt / 2 + kLoneCost * (nt & 1)) * T + (nt / 2 + (nt & 1)) * kStepCost : nt * (T + kStepCost);
            const long long wantC = units * T / 2 > units ? units * T / 2 : units;
            const int gridC = (int)(wantC < sms ? wantC : sms);
            const float itersC = (float)(units * T) / gridC;
            const float costC = itersC + (itersC / T + 1.f) * kStepCost + kMergeCost;
            split = costC < costA ? 1 : 0;
        }
    }
    const long long iters = split ? units * T : total;
    long long want = split ? (iters / 2 > units ? iters / 2 : units) : (TPU == 2 ? (total + 1) / 2 : total);
    const int grid = (int)(want < sms ? want : sms);
    int dev = 0;
    TMX_CUDA(cudaGetDevice(&dev));
    float* ws = g_ws[dev];
    unsigned int* wf = g_ws_flags[dev];
#define TMX_ATTN_LAUNCH(B16, HV) attn_fwd_kernel<B16, HV><<<grid, AttnCfg<HV>::kThreads, kSmemBytes, st>>>(mq, mk, mv, mo, Nq, Nk, H, QT, UPP, TPU, (int)iters, split, ws, wf, scale_log2)
    if (g_halves == 2) { if (bf16) TMX_ATTN_LAUNCH(true, 2); else TMX_ATTN_LAUNCH(false, 2); }
    else               { if (bf16) TMX_ATTN_LAUNCH(true, 1); else TMX_ATTN_LAUNCH(false, 1); }
#undef TMX_ATTN_LAUNCH
    return check_cuda(cudaGetLastError(), "attn_fwd_kernel launch");
}
