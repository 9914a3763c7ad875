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

int short_kv_attn_init() {
    using namespace xattn;
    if (int rc = set_smem_attr<true, 80, 1>()) return rc;
    if (int rc = set_smem_attr<false, 80, 1>()) return rc;
    if (int rc = set_smem_attr<true, 80, 2>()) return rc;
    if (int rc = set_smem_attr<false, 80, 2>()) return rc;
    if (int rc = set_smem_attr<true, 128, 1>()) return rc;
    if (int rc = set_smem_attr<false, 128, 1>()) return rc;
    return TMX_OK;
}

// Called by tmx_attn_fwd for Nk <= 128 (arguments already validated there).  mt = 16-row tiles per warp iteration (1 or 2).
int short_kv_attn_launch(const void* q, const void* k, const void* v, void* o, int B, int H, int Nq, int Nk,
                         int64_t q_sn, int64_t k_sn, int64_t v_sn, int64_t o_sn, float scale, bool bf16, int mt, cudaStream_t st) {
    using namespace xattn;
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
