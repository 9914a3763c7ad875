// k2s — attention against a SHORT key/value sequence (cross-attention to the 77 text tokens), head dim 64:
//   O = softmax(scale * Q K^T) V  with Nk <= 128.
//
// Why this is not the tcgen05 kernel of attention.cu: with 77 keys the op is a STREAM over Q and O.  Per query row it moves
// 256 B (Q in, O out) and computes 4 * 80 * 64 = 20 kFLOP: 40 FLOP/B against a ridge of ~250 FLOP/B, i.e. 3-6 us of HBM time per
// SDXL site and 0.5 % of the step's FLOPs.  The TMEM pipeline of attention.cu is built for long K/V streams; on a single K/V
// tile every query tile pays the whole TMA -> MMA -> tcgen05.ld -> exp -> tcgen05.st -> MMA -> epilogue latency chain with two
// tiles in flight per SM (measured 17-26 us per launch, 1.3 ms per K=3 step).  Here the latency is hidden the way a bandwidth
// kernel hides it — many independent warps per SM, each owning a block of query rows end to end:
//   * one CTA = one (batch, head) and a contiguous range of query rows; K and V (<= 128 x 64) are staged ONCE per CTA in shared
//     memory (zero-filled beyond Nk), rows padded to 144 B so that ldmatrix is conflict-free;
//   * a warp takes 16*MT query rows at a time: 16-byte cp.async loads (full 128 B rows, coalesced) into its private staging
//     tile, double-buffered so the next block is in flight while this one computes; Q, P and the accumulators never leave
//     registers: S = Q K^T and O = P V by warp-level mma (m16n8k16, fp32 accumulate), softmax in the exp2 domain on the
//     accumulator fragments, P rounded to the I/O dtype before P V (as in attention.cu and as the reference's second einsum
//     sees it, utils_custom.py:100-103), the row sum accumulated in fp32 from the unrounded exponentials;
//   * O is normalised, packed, transposed through the warp's staging tile and written as full 128 B rows.
// Replaces, for attn2, the einsum -> softmax -> einsum of fusion_generation/utils_custom.py:91-105 and utils_lora.py:99-113.
#include "tmx_common.cuh"
#include <cuda.h>

namespace tmx {

namespace xattn {

constexpr int kD = 64;
constexpr int kPad = 72;             // staged row pitch in elements (144 B): ldmatrix rows fall into distinct banks
constexpr int kWarps = 4;
#ifndef TMX_XATTN_CTAS
#define TMX_XATTN_CTAS 4
#endif
constexpr int kCtasPerSm = TMX_XATTN_CTAS;   // resident CTAs per SM asked of the compiler for MT == 1 (register budget 65536 / (128 * kCtasPerSm))

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;                                   // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <bool BF16>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if constexpr (BF16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    if constexpr (BF16) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
    else { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
}

// KP = keys padded to a multiple of 16 (80 for the 77 text tokens, 128 at most); MT = 16-row tiles per warp iteration.
template <bool BF16, int KP, int MT>
__global__ void __launch_bounds__(kWarps * 32, MT == 1 ? kCtasPerSm : 3)
short_kv_attn_kernel(const uint16_t* __restrict__ q, const uint16_t* __restrict__ k, const uint16_t* __restrict__ v, uint16_t* __restrict__ o,
                     int Nq, int Nk, int H, long long q_sn, long long k_sn, long long v_sn, long long o_sn, int rows_per_cta, float scale_log2) {
    constexpr int RB = 16 * MT;                  // query rows per warp iteration
    constexpr int NT = KP / 8;                   // key n-tiles of S
    constexpr int KS = KP / 16;                  // key k-steps of P V
    extern __shared__ __align__(16) uint16_t smem[];
    uint16_t* sK = smem;                         // [KP][kPad]
    uint16_t* sV = sK + KP * kPad;               // [KP][kPad]
    uint16_t* sQ = sV + KP * kPad;               // [kWarps][2][RB][kPad]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y / H, h = blockIdx.y - b * H;
    const int row0 = blockIdx.x * rows_per_cta;
    const int row_end = min(Nq, row0 + rows_per_cta);
    const uint16_t* qb = q + (long long)b * Nq * q_sn + (long long)h * kD;
    uint16_t* ob = o + (long long)b * Nq * o_sn + (long long)h * kD;
    const uint32_t sQw = smem_u32(sQ + warp * 2 * RB * kPad);

    // this warp's blocks: warp, warp + kWarps, ... of the CTA's row range
    const int nblk_cta = (row_end - row0 + RB - 1) / RB;
    auto prefetch = [&](int blk, int buf) {      // 16*MT rows x 8 chunks of 16 B; a quarter-warp covers one full 128 B row
#pragma unroll
        for (int i = 0; i < RB / 4; ++i) {
            const int r = (lane >> 3) + 4 * i, c = lane & 7;
            const int row = row0 + blk * RB + r;
            cp_async16(sQw + (uint32_t)(((buf * RB + r) * kPad + c * 8) * 2), qb + (long long)min(row, Nq - 1) * q_sn + c * 8, row < row_end);
        }
        cp_async_commit();
    };
    if (warp < nblk_cta) prefetch(warp, 0);

    // K / V -> shared, every 16-byte chunk in flight at once (rows beyond Nk zero-filled: they meet P == 0, but must not be NaN)
    {
        const uint16_t* kb = k + (long long)b * Nk * k_sn + (long long)h * kD;
        const uint16_t* vb = v + (long long)b * Nk * v_sn + (long long)h * kD;
        const uint32_t sK0 = smem_u32(sK), sV0 = smem_u32(sV);
#pragma unroll
        for (int i0 = 0; i0 < KP * 8; i0 += kWarps * 32) {
            const int i = i0 + threadIdx.x;
            const int r = i >> 3, c = i & 7;
            const int rs = min(r, Nk - 1);
            cp_async16(sK0 + (uint32_t)((r * kPad + c * 8) * 2), kb + (long long)rs * k_sn + c * 8, r < Nk);
            cp_async16(sV0 + (uint32_t)((r * kPad + c * 8) * 2), vb + (long long)rs * v_sn + c * 8, r < Nk);
        }
        cp_async_commit();
        cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t sK_a = smem_u32(sK), sV_a = smem_u32(sV);
    const int g = lane >> 2, t = lane & 3;
    // per-lane ldmatrix row addresses (see the fragment maps in the comments below)
    const uint32_t q_lane = (uint32_t)((((lane & 7) + 8 * ((lane >> 3) & 1)) * kPad + 8 * (lane >> 4)) * 2);       // A: rows, then k halves
    const uint32_t k_lane = (uint32_t)(((lane & 7) * kPad + 8 * (lane >> 3)) * 2);                               // B of Q K^T: key rows, 4 dim chunks
    const uint32_t v_lane = (uint32_t)((((lane & 7) + 8 * ((lane >> 3) & 1)) * kPad + 8 * (lane >> 4)) * 2);       // B of P V (.trans): key rows, 2 dim chunks

    int buf = 0;
    for (int blk = warp; blk < nblk_cta; blk += kWarps, buf ^= 1) {
        const bool more = blk + kWarps < nblk_cta;
        if (more) prefetch(blk + kWarps, buf ^ 1);
        if (more) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncwarp();
        const uint32_t sQb = sQw + (uint32_t)(buf * RB * kPad * 2);

        // ---- Q fragments (A, row-major 16x16 per k-step): a0 (row g, k 2t..), a1 (row g+8), a2 (row g, k 2t+8..), a3 (row g+8, k 2t+8..)
        uint32_t qa[MT][4][4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                ldsm_x4(sQb + (uint32_t)((m * 16 * kPad + kk * 16) * 2) + q_lane, qa[m][kk][0], qa[m][kk][1], qa[m][kk][2], qa[m][kk][3]);

        // ---- S = Q K^T: per key n-tile j two ldmatrix.x4 fetch the B fragments (b0, b1) of all four k-steps
        float s[MT][NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            uint32_t kb0[4], kb1[4];
            ldsm_x4(sK_a + (uint32_t)((j * 8 * kPad) * 2) + k_lane, kb0[0], kb1[0], kb0[1], kb1[1]);            // dims 0..31
            ldsm_x4(sK_a + (uint32_t)((j * 8 * kPad + 32) * 2) + k_lane, kb0[2], kb1[2], kb0[3], kb1[3]);       // dims 32..63
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                s[m][j][0] = s[m][j][1] = s[m][j][2] = s[m][j][3] = 0.f;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) mma16816<BF16>(s[m][j], qa[m][kk], kb0[kk], kb1[kk]);
            }
        }

        // ---- softmax on the accumulator fragments: c0, c1 = (row g, keys 8j + 2t, +1), c2, c3 = (row g + 8, same keys)
        uint32_t p[MT][KS][4];
        float inv_l[MT][2];
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int col = 8 * j + 2 * t;
                if (8 * j + 8 > Nk) {                                       // warp-uniform: only the n-tiles that reach beyond Nk
                    if (col >= Nk) { s[m][j][0] = -INFINITY; s[m][j][2] = -INFINITY; }
                    if (col + 1 >= Nk) { s[m][j][1] = -INFINITY; s[m][j][3] = -INFINITY; }
                }
                mx0 = fmaxf(mx0, fmaxf(s[m][j][0], s[m][j][1]));
                mx1 = fmaxf(mx1, fmaxf(s[m][j][2], s[m][j][3]));
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float nm0 = -mx0 * scale_log2, nm1 = -mx1 * scale_log2;
            float l0 = 0.f, l1 = 0.f;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const float e0 = ex2f(fmaf(s[m][j][0], scale_log2, nm0)), e1 = ex2f(fmaf(s[m][j][1], scale_log2, nm0));
                const float e2 = ex2f(fmaf(s[m][j][2], scale_log2, nm1)), e3 = ex2f(fmaf(s[m][j][3], scale_log2, nm1));
                l0 += e0 + e1;
                l1 += e2 + e3;
                // A fragment of P V for key k-step j / 2: n-tile 2kk -> a0 (row g), a1 (row g+8); n-tile 2kk+1 -> a2, a3
                p[m][j >> 1][(j & 1) * 2 + 0] = pack2<BF16>(e0, e1);
                p[m][j >> 1][(j & 1) * 2 + 1] = pack2<BF16>(e2, e3);
            }
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
            inv_l[m][0] = __frcp_rn(l0);                             // l >= 1 (the row maximum contributes 2^0): no range checks needed
            inv_l[m][1] = __frcp_rn(l1);
        }

        // ---- O = P V: per key k-step and PAIR of dim n-tiles one ldmatrix.x4.trans fetches (b0, b1) of both n-tiles
        float oacc[MT][8][4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int jd = 0; jd < 8; ++jd) oacc[m][jd][0] = oacc[m][jd][1] = oacc[m][jd][2] = oacc[m][jd][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                uint32_t b00, b01, b10, b11;
                ldsm_x4_t(sV_a + (uint32_t)((kk * 16 * kPad + jp * 16) * 2) + v_lane, b00, b01, b10, b11);
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    mma16816<BF16>(oacc[m][2 * jp], p[m][kk], b00, b01);
                    mma16816<BF16>(oacc[m][2 * jp + 1], p[m][kk], b10, b11);
                }
            }
        }

        // ---- O / l -> 16 bit -> staging tile (the Q buffer of this block: its fragments are in registers) -> full-row stores
        __syncwarp();
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int jd = 0; jd < 8; ++jd) {
                const uint32_t a0 = sQb + (uint32_t)(((m * 16 + g) * kPad + jd * 8 + 2 * t) * 2);
                asm volatile("st.shared.u32 [%0], %1;" :: "r"(a0), "r"(pack2<BF16>(oacc[m][jd][0] * inv_l[m][0], oacc[m][jd][1] * inv_l[m][0])) : "memory");
                asm volatile("st.shared.u32 [%0], %1;" :: "r"(a0 + 8 * kPad * 2), "r"(pack2<BF16>(oacc[m][jd][2] * inv_l[m][1], oacc[m][jd][3] * inv_l[m][1])) : "memory");
            }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < RB / 4; ++i) {
            const int r = (lane >> 3) + 4 * i, c = lane & 7;
            const int row = row0 + blk * RB + r;
            uint4 val;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w) : "r"(sQb + (uint32_t)((r * kPad + c * 8) * 2)));
            if (row < row_end) st_stream(ob + (long long)row * o_sn + c * 8, val);
        }
        __syncwarp();                    // the staging tile is free for the prefetch issued two iterations from now
    }
}

template <bool BF16, int KP, int MT>
static int launch(const void* q, const void* k, const void* v, void* o, int B, int H, int Nq, int Nk,
                  int64_t q_sn, int64_t k_sn, int64_t v_sn, int64_t o_sn, float scale_log2, cudaStream_t st) {
    constexpr int RB = 16 * MT;
    const size_t smem = (size_t)(2 * KP * kPad + kWarps * 2 * RB * kPad) * 2;
    // one resident wave: as many row chunks per (b, h) as fit the CTA slots of the device, each a multiple of kWarps * RB rows
    const int slots = sm_count() * (MT == 1 ? kCtasPerSm : 3);
    const int quantum = kWarps * RB;
    const int q_units = (Nq + quantum - 1) / quantum;
    int chunks = slots / (B * H);
    chunks = chunks < 1 ? 1 : (chunks > q_units ? q_units : chunks);
    const int rows_per_cta = ((q_units + chunks - 1) / chunks) * quantum;
    dim3 grid((Nq + rows_per_cta - 1) / rows_per_cta, B * H);
    short_kv_attn_kernel<BF16, KP, MT><<<grid, kWarps * 32, smem, st>>>(
        (const uint16_t*)q, (const uint16_t*)k, (const uint16_t*)v, (uint16_t*)o, Nq, Nk, H, q_sn, k_sn, v_sn, o_sn, rows_per_cta, scale_log2);
    return check_cuda(cudaGetLastError(), "short_kv_attn_kernel launch");
}

template <bool BF16, int KP, int MT>
static int set_smem_attr() {
    const size_t smem = (size_t)(2 * KP * kPad + kWarps * 2 * 16 * MT * kPad) * 2;
    return check_cuda(cudaFuncSetAttribute(short_kv_attn_kernel<BF16, KP, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "short_kv_attn smem attribute");
}

}  // namespace xattn


// ================================================================================================================
// k2t — the same op on the tcgen05 path, built for ONE K/V tile per query tile (Nk <= 80).
//
// attention.cu keeps two query tiles in flight per SM and runs each through the whole TMA -> MMA -> tcgen05.ld -> exp ->
// tcgen05.st -> MMA -> epilogue chain before the next tile of that slot starts: with a single K/V tile there is nothing to
// amortise that latency over.  Here the stages are decoupled and THREE query tiles are in flight per SM:
//   * persistent, one 480-thread CTA per SM over a contiguous range of (b, h, 128-row query tile) units;
//   * warp 12 = TMA producer: an 8-deep ring of Q tiles (128 x 64, one per unit: the loads run up to eight units ahead, the op is a
//     stream and HBM latency is what has to be covered) and a 2-deep ring of (K, V) pairs (80 x 64 each, rows beyond Nk zero-filled
//     by the tensor map), loaded once per run of units of the same (b, h); SWIZZLE_128B, straight from the strided tensors;
//   * warp 13 = S issuer: S(i) = Q K^T (4 x tcgen05.mma M=128 N=80 K=16) into TMEM slot i % 3 as soon as its Q tile has landed
//     and the slot is drained; warp 14 = P V issuer: O(k) = P V (5 x M=128 N=64 K=16, P read from TMEM) as soon as the softmax
//     of unit k has handed P over — two warps because the elected thread's issue cost (~70-100 clk per MMA in the uniform
//     datapath) was a ~1000 clk serial section per unit with one;
//   * warps 0-11 = three softmax / epilogue groups (group = TMEM slot; thread == row == TMEM lane): tcgen05.ld S (80 columns)
//     -> exact row max -> exp2 -> row sum -> P as packed 16-bit written over the S columns (tcgen05.st) -> hand to the MMA warp ->
//     when O is complete: tcgen05.ld -> * 1/l -> swizzled staging tile -> one TMA store; the slot is released as soon as O sits
//     in registers;
//   * TMEM: 3 x (S/P 96 + O 64 columns) of the 512; 128 registers per thread suffice (80 S columns per row), no setmaxnreg.
namespace xattn_tc {

constexpr int kD = 64, kBM = 128, kKP = 80;
constexpr int kQBytes = kBM * kD * 2;              // 16 KiB
constexpr int kKVBytes = kKP * kD * 2;             // 10 KiB (a multiple of the 1 KiB swizzle atom)
constexpr int kStages = 8, kSlots = 3, kKvSlots = 2;   // Q ring (one 16 KiB tile per unit), TMEM slots, K/V ring (one pair per (b, h) run)
constexpr int kSoftmaxWarps = 4 * kSlots;
constexpr int kThreads = 32 * (kSoftmaxWarps + 3);   // + TMA producer, S issuer, P V issuer
constexpr int kSlotCols = 160;                     // S / P at +0 (80 / 40 columns), O at +96 (64 columns)
constexpr int kBars = 2 * kStages + 2 * kKvSlots + 4 * kSlots;
constexpr int kSmemBytes = 1024 + kStages * kQBytes + kKvSlots * 2 * kKVBytes + kSlots * kQBytes + kBars * 8 + 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {       // NON-blocking probe (try_wait may suspend the warp for a while)
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {      // bounded: a protocol bug traps, never hangs
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++polls & 255u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now; else if (now - t0 > 4000000000LL) __trap();
        }
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// descriptors as in attention.cu: low word = start >> 4 | LBO (16 B) << 16; high word = SBO 1024 B, version 1, SWIZZLE_128B
constexpr uint32_t kDescHi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | ((16u >> 4) << 16); }
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kDescHi) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 db, {%2, %5};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kDescHi) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc(bool bf16, int M, int N, bool b_mn_major) {
    return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
#define XA_R8(r, o)  "=r"(r[o+0]), "=r"(r[o+1]), "=r"(r[o+2]), "=r"(r[o+3]), "=r"(r[o+4]), "=r"(r[o+5]), "=r"(r[o+6]), "=r"(r[o+7])
#define XA_I8(r, o)  "r"(r[o+0]), "r"(r[o+1]), "r"(r[o+2]), "r"(r[o+3]), "r"(r[o+4]), "r"(r[o+5]), "r"(r[o+6]), "r"(r[o+7])
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : XA_R8(r, 0), XA_R8(r, 8) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : XA_R8(r, 0), XA_R8(r, 8), XA_R8(r, 16), XA_R8(r, 24) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 :: "r"(taddr), XA_I8(r, 0), XA_I8(r, 8), XA_I8(r, 16), XA_I8(r, 24) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" :: "r"(taddr), XA_I8(r, 0) : "memory");
}
__device__ __forceinline__ float ex2f_(float x) { return xattn::ex2f(x); }
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1)
short_kv_attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                        int Nk, int H, int QT, int total_units, float scale_log2) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint32_t s0 = smem_u32(smem);
    asm volatile("mov.u32 %0, %0;" : "+r"(s0));
    const uint32_t sKV = s0 + kStages * kQBytes;                       // [kKvSlots] x (K tile, V tile)
    const uint32_t sO = sKV + kKvSlots * 2 * kKVBytes;                 // [kSlots] staging tiles (128 rows x 128 B, SWIZZLE_128B)
    const uint32_t full = sO + kSlots * kQBytes, empty = full + 8 * kStages;
    const uint32_t kv_full = empty + 8 * kStages, kv_empty = kv_full + 8 * kKvSlots;
    const uint32_t s_full = kv_empty + 8 * kKvSlots, p_full = s_full + 8 * kSlots, o_full = p_full + 8 * kSlots, slot_free = o_full + 8 * kSlots;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kStages * kQBytes + kKvSlots * 2 * kKVBytes + kSlots * kQBytes + kBars * 8);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int u_begin = (int)(((long long)blockIdx.x * total_units) / gridDim.x);
    const int n_units = (int)(((long long)(blockIdx.x + 1) * total_units) / gridDim.x) - u_begin;

    if (warp == kSoftmaxWarps + 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 1); }
        for (int i = 0; i < kKvSlots; ++i) { mbar_init(kv_full + 8 * i, 1); mbar_init(kv_empty + 8 * i, 1); }
        for (int i = 0; i < kSlots; ++i) { mbar_init(s_full + 8 * i, 1); mbar_init(p_full + 8 * i, 4); mbar_init(o_full + 8 * i, 1); mbar_init(slot_free + 8 * i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kSoftmaxWarps) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tm_q)) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tm_k)) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tm_v)) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(&tm_o)) : "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp >= kSoftmaxWarps) {
        if (warp == kSoftmaxWarps) {
            // ================================ TMA producer: Q per unit, (K, V) per run of units of one (b, h) ================================
            int prev_bh = -1, kvi = -1;
            for (int i = 0; i < n_units; ++i) {
                const int st = i % kStages;
                const int u = u_begin + i, bh = u / QT, qt = u - bh * QT, b = bh / H, h = bh - b * H;
                if (bh != prev_bh) {
                    prev_bh = bh;
                    ++kvi;
                    const int g = kvi % kKvSlots;
                    mbar_wait(kv_empty + 8 * g, ((uint32_t)(kvi / kKvSlots) & 1u) ^ 1u);
                    if (elect_one()) {
                        const uint32_t dst = sKV + g * 2 * kKVBytes, bar = kv_full + 8 * g;
                        mbar_expect_tx(bar, 2 * kKVBytes);
                        tma_load_4d(dst, &tm_k, bar, 0, h, 0, b);
                        tma_load_4d(dst + kKVBytes, &tm_v, bar, 0, h, 0, b);
                    }
                }
                mbar_wait(empty + 8 * st, ((uint32_t)(i / kStages) & 1u) ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(full + 8 * st, kQBytes);
                    tma_load_4d(s0 + st * kQBytes, &tm_q, full + 8 * st, 0, h, qt * kBM, b);
                }
            }
        } else if (warp == kSoftmaxWarps + 1) {
            // ================================ S issuer: S(i) = Q K^T as soon as Q(i) has landed and slot i % 3 is drained ================================
            // (One elected thread spends ~70-100 clk per tcgen05.mma in the uniform datapath: with S and P V issued by the same warp the
            // nine MMAs + four commits of a unit were a ~1000 clk serial section per unit; two issuer warps halve it.)
            constexpr uint32_t idesc_qk = make_idesc(BF16, kBM, kKP, false);
            int prev_bh = -1, kvi = -1;
            for (int i = 0; i < n_units; ++i) {
                const int st = i % kStages, j = i % kSlots;
                const int bh = (u_begin + i) / QT;
                if (bh != prev_bh) { prev_bh = bh; ++kvi; }
                const int g = kvi % kKvSlots;
                mbar_wait(full + 8 * st, (uint32_t)(i / kStages) & 1u);
                mbar_wait(kv_full + 8 * g, (uint32_t)(kvi / kKvSlots) & 1u);
                // the S / P columns of slot j are free as soon as P V of the slot's previous unit has completed (its o_full): the group
                // may still be busy with that unit's epilogue — only the O columns are still in use, and those are the P V issuer's concern
                if (i >= kSlots) mbar_wait(o_full + 8 * j, ((uint32_t)(i / kSlots) - 1u) & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t q_lo = desc_lo(s0 + st * kQBytes), k_lo = desc_lo(sKV + g * 2 * kKVBytes);
                    const uint32_t t_s = tmem_base + j * kSlotCols;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_ss(t_s, q_lo + 2 * ks, k_lo + 2 * ks, idesc_qk, ks > 0 ? 1u : 0u);   // + 32 B inside the swizzle row
                    umma_commit(s_full + 8 * j);
                    umma_commit(empty + 8 * st);                                   // the Q tile has been consumed
                }
            }
        } else {
            // ================================ P V issuer: O(k) = P V as soon as the softmax of unit k has handed P over ================================
            constexpr uint32_t idesc_pv = make_idesc(BF16, kBM, kD, true);
            int prev_bh = -1, kvi = -1;
            for (int k = 0; k < n_units; ++k) {
                const int j = k % kSlots;
                const int bh = (u_begin + k) / QT;
                if (bh != prev_bh) { prev_bh = bh; ++kvi; }
                const int g = kvi % kKvSlots;
                const bool last_of_run = k + 1 >= n_units || (u_begin + k + 1) / QT != bh;
                mbar_wait(p_full + 8 * j, (uint32_t)(k / kSlots) & 1u);
                if (k >= kSlots) mbar_wait(slot_free + 8 * j, ((uint32_t)(k / kSlots) - 1u) & 1u);   // O of the slot's previous unit sits in registers
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t t_p = tmem_base + j * kSlotCols, t_o = t_p + 96;
                    const uint32_t v_lo = desc_lo(sKV + g * 2 * kKVBytes + kKVBytes);
#pragma unroll
                    for (int ks = 0; ks < kKP / 16; ++ks)                         // 16 keys per step: + 2048 B in V, + 8 columns in P
                        umma_ts(t_o, t_p + ks * 8, v_lo + ks * 128, idesc_pv, ks > 0 ? 1u : 0u);
                    umma_commit(o_full + 8 * j);
                    if (last_of_run) umma_commit(kv_empty + 8 * g);                // K and V of this (b, h) run have been consumed
                }
            }
        }
    } else {
        // ================================ softmax + epilogue group of TMEM slot `slot` ================================
        const int slot = warp >> 2, quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t t_s = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * kSlotCols, t_o = t_s + 96;
        const uint32_t o_tile = sO + slot * kQBytes;
        const uint32_t o_row = o_tile + row * 128, o_sw = (uint32_t)(row & 7);
        const bool issuer = quarter == 0 && lane == 0;
        const uint64_t sc2 = pk2(scale_log2, scale_log2);
        for (int i = slot; i < n_units; i += kSlots) {
            const uint32_t par = (uint32_t)(i / kSlots) & 1u;
            mbar_wait(s_full + 8 * slot, par);
            tc_fence_after();
            uint32_t s[kKP];
            tmem_ld32(t_s, s);
            tmem_ld32(t_s + 32, s + 32);
            tmem_ld16(t_s + 64, s + 64);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = kKP - 16; c < kKP; ++c) if (c >= Nk) s[c] = 0xff800000u;          // keys beyond Nk (only the last 16 columns can be; Nk > 64 is checked on the host)
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int c = 0; c < kKP; c += 4) {
                mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s[c]), __uint_as_float(s[c + 1])));
                mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s[c + 2]), __uint_as_float(s[c + 3])));
            }
            const float nm = -fmaxf(mx0, mx1) * scale_log2;
            const uint64_t nm2 = pk2(nm, nm);
            float l0 = 0.f, l1 = 0.f;
            uint32_t p[kKP / 2];
#pragma unroll
            for (int c = 0; c < kKP / 2; ++c) {
                float x0, x1;
                upk2(ffma2(pk2(__uint_as_float(s[2 * c]), __uint_as_float(s[2 * c + 1])), sc2, nm2), x0, x1);
                const float e0 = ex2f_(x0), e1 = ex2f_(x1);
                l0 += e0; l1 += e1;
                p[c] = xattn::pack2<BF16>(e0, e1);
            }
            tmem_st32(t_s, p);                                                             // P over the (consumed) S columns
            tmem_st8(t_s + 32, p + 32);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full + 8 * slot);
            const float inv_l = __frcp_rn(l0 + l1);                                        // l >= 1: the row maximum contributes 2^0

            // ---- epilogue: O / l -> 16 bit -> this group's staging tile -> one TMA store (rows beyond Nq clipped by the tensor map)
            mbar_wait(o_full + 8 * slot, par);
            tc_fence_after();
            if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the previous store has read the staging tile
            asm volatile("bar.sync %0, 128;" :: "r"(1 + slot) : "memory");
            uint32_t o[kD];
            tmem_ld32(t_o, o);
            tmem_ld32(t_o + 32, o + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(slot_free + 8 * slot);                              // the O columns of this slot may be overwritten
#pragma unroll
            for (int c = 0; c < kD; c += 8) {
                uint4 v;
                v.x = xattn::pack2<BF16>(__uint_as_float(o[c]) * inv_l, __uint_as_float(o[c + 1]) * inv_l);
                v.y = xattn::pack2<BF16>(__uint_as_float(o[c + 2]) * inv_l, __uint_as_float(o[c + 3]) * inv_l);
                v.z = xattn::pack2<BF16>(__uint_as_float(o[c + 4]) * inv_l, __uint_as_float(o[c + 5]) * inv_l);
                v.w = xattn::pack2<BF16>(__uint_as_float(o[c + 6]) * inv_l, __uint_as_float(o[c + 7]) * inv_l);
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(o_row + ((((uint32_t)c >> 3) ^ o_sw) << 4)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync %0, 128;" :: "r"(1 + slot) : "memory");
            if (issuer) {
                const int u = u_begin + i, bh = u / QT, qt = u - bh * QT, b = bh / H, h = bh - b * H;
                asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                             :: "l"(reinterpret_cast<uint64_t>(&tm_o)), "r"(o_tile), "r"(0), "r"(h), "r"(qt * kBM), "r"(b) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kSoftmaxWarps) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(512) : "memory");
    }
}

}  // namespace xattn_tc

typedef CUresult (*XaEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static XaEncodeTiledFn g_xa_encode = nullptr;

int short_kv_attn_init() {
    using namespace xattn;
    if (!g_xa_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        TMX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        TMX_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, TMX_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
        g_xa_encode = (XaEncodeTiledFn)fn;
    }
    TMX_CUDA(cudaFuncSetAttribute(xattn_tc::short_kv_attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, xattn_tc::kSmemBytes));
    TMX_CUDA(cudaFuncSetAttribute(xattn_tc::short_kv_attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, xattn_tc::kSmemBytes));
    if (int rc = set_smem_attr<true, 80, 1>()) return rc;
    if (int rc = set_smem_attr<false, 80, 1>()) return rc;
    if (int rc = set_smem_attr<true, 80, 2>()) return rc;
    if (int rc = set_smem_attr<false, 80, 2>()) return rc;
    if (int rc = set_smem_attr<true, 128, 1>()) return rc;
    if (int rc = set_smem_attr<false, 128, 1>()) return rc;
    return TMX_OK;
}

// [B, N, H, 64] 16-bit tensor with token stride `stride_n` elements -> 4-D map (d, h, n, b), box 64 x 1 x rows x 1
static int xa_make_map(CUtensorMap* m, const void* base, int B, int N, int H, int64_t stride_n, int box_rows, bool bf16) {
    cuuint64_t dims[4] = {64, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[3] = {128, (cuuint64_t)stride_n * 2, (cuuint64_t)N * (cuuint64_t)stride_n * 2};
    cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_xa_encode(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return TMX_ECUDA; }
    return TMX_OK;
}

static int short_kv_attn_tc_launch(const void* q, const void* k, const void* v, void* o, int B, int H, int Nq, int Nk,
                                   int64_t q_sn, int64_t k_sn, int64_t v_sn, int64_t o_sn, float scale, bool bf16, cudaStream_t st) {
    using namespace xattn_tc;
    CUtensorMap mq, mk, mv, mo;
    if (int rc = xa_make_map(&mq, q, B, Nq, H, q_sn, kBM, bf16)) return rc;
    if (int rc = xa_make_map(&mo, o, B, Nq, H, o_sn, kBM, bf16)) return rc;
    if (int rc = xa_make_map(&mk, k, B, Nk, H, k_sn, kKP, bf16)) return rc;
    if (int rc = xa_make_map(&mv, v, B, Nk, H, v_sn, kKP, bf16)) return rc;
    const int QT = (Nq + kBM - 1) / kBM;
    const long long units = (long long)B * H * QT;
    const int grid = (int)(units < sm_count() ? units : sm_count());
    const float sl2 = scale * 1.4426950408889634f;
    if (bf16) short_kv_attn_tc_kernel<true><<<grid, kThreads, kSmemBytes, st>>>(mq, mk, mv, mo, Nk, H, QT, (int)units, sl2);
    else      short_kv_attn_tc_kernel<false><<<grid, kThreads, kSmemBytes, st>>>(mq, mk, mv, mo, Nk, H, QT, (int)units, sl2);
    return check_cuda(cudaGetLastError(), "short_kv_attn_tc_kernel launch");
}

// Called by tmx_attn_fwd for Nk <= 128 (arguments already validated there).  mode 1 / 2 = streaming kernel with 1 / 2 16-row tiles
// per warp iteration; mode 3 = the tcgen05 kernel k2t where it applies (64 < Nk <= 80: the 77 text tokens), else mode 1.
int short_kv_attn_launch(const void* q, const void* k, const void* v, void* o, int B, int H, int Nq, int Nk,
                         int64_t q_sn, int64_t k_sn, int64_t v_sn, int64_t o_sn, float scale, bool bf16, int mt, cudaStream_t st) {
    using namespace xattn;
    if (mt == 3) {
        if (Nk > 64 && Nk <= xattn_tc::kKP && (long long)B * H * ((Nq + 127) / 128) < (1ll << 30))
            return short_kv_attn_tc_launch(q, k, v, o, B, H, Nq, Nk, q_sn, k_sn, v_sn, o_sn, scale, bf16, st);
        mt = 1;
    }
    const float sl2 = scale * 1.4426950408889634f;
    TMX_REQUIRE((long long)B * H <= 65535, TMX_ESHAPE, "attn (short K/V): B*H = %lld exceeds the grid's y extent", (long long)B * H);
#define TMX_XA(B16, KP, MT) return launch<B16, KP, MT>(q, k, v, o, B, H, Nq, Nk, q_sn, k_sn, v_sn, o_sn, sl2, st)
    if (Nk <= 80) {
        if (mt == 2) { if (bf16) TMX_XA(true, 80, 2); else TMX_XA(false, 80, 2); }
        else         { if (bf16) TMX_XA(true, 80, 1); else TMX_XA(false, 80, 1); }
    } else {                                     // 64 S accumulators per thread already: one 16-row tile per warp iteration
        if (bf16) TMX_XA(true, 128, 1); else TMX_XA(false, 128, 1);
    }
#undef TMX_XA
}

}  // namespace tmx
