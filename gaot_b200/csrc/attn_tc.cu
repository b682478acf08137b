// Tensor-core attention (performance path): non-causal softmax(Q K^T * scale) V on tcgen05 with TMA-staged
// bf16 tiles and TMEM accumulators, forward and backward.  Reference: GroupQueryFlashAttention.forward,
// src/model/layers/attn.py:96-116 (F.scaled_dot_product_attention) and its autograd backward.
//
// Layout: fused qkv [B*T, ld] bf16 (per token: H*hd query | Hkv*hd key | Hkv*hd value columns), one TMA
// tensor map over that matrix with a {hd, 128 rows} box whose swizzle span equals the row bytes of a head.
//
// Forward, one CTA per (128-query tile, head, sample):
//   warp 0   TMA producer: Q once, then K/V tiles of 128 keys through a 2-stage mbarrier ring
//   warp 1   MMA issuer  : S = Q K^T  (tcgen05.mma M=128,N=128, operands straight from the swizzled tiles)
//                          PV = P V   (A = P written by the softmax warps as a K-major SWIZZLE_128B tile,
//                                      B = V used in place as an MN-major operand)
//   warps 2-5 softmax    : one thread per query row reads its S row from TMEM (tcgen05.ld), online softmax in
//                          base 2, writes P (bf16) to shared memory, then folds PV into its fp32 output row
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "../../include/gaot_b200.h"

namespace {

constexpr int AT_THREADS = 320;           // forward: TMA warp, MMA warp, 8 softmax warps (2 per TMEM lane quadrant)
constexpr int AT_BQ = 128, AT_BKV = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// 3-D form: coordinate c2 selects the hi / lo plane of a split-precision tensor ([2, rows, ld] fp16)
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// shared tile -> global, element-wise fp32 add (TMA reduce; completion tracked by the issuing thread's bulk groups)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand read from tensor memory ([lanes = M rows] x [K elements, two 16-bit values per 32-bit column]), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_c), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp): rows of ROWB bytes, swizzle span == ROWB,
// 8-row atoms SBO bytes apart; layout code 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
template <int ROWB>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    constexpr uint64_t layout = ROWB == 128 ? 2 : (ROWB == 64 ? 4 : 6);
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((8 * ROWB) >> 4) << 32) | ((uint64_t)1 << 46) |
           (layout << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
        "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// issue only: several loads can be put in flight before one tmem_wait_ld() (each wait is a full TMEM round trip)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
        "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float ex2(float x) {          // single MUFU.EX2 (ex2(-inf) = 0)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// explicit shared-state-space accesses on 32-bit addresses (pointers derived from the aligned dynamic-smem base are
// generic to the compiler, which would otherwise emit LD.E / ST.E with 64-bit address arithmetic)
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }
// global load the compiler may not sink to its first use (prefetch one step ahead)
__device__ __forceinline__ float ldg_f32(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// predicated global load with nothing depending on the result until its real use (prefetch one step ahead)
__device__ __forceinline__ float ldg_f32_pred(const float* p, bool ok, float otherwise) {
    float v;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %2, 0;\n\t"
        "mov.f32 %0, %3;\n\t"
        "@p ld.global.nc.f32 %0, [%1];\n\t}"
        : "=f"(v) : "l"(p), "r"((int)ok), "f"(otherwise));
    return v;
}
// tcgen05.ld 32x32b of N (4 / 8 / 16) consecutive columns, issue only
template <int N>
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t (&r)[N]) {
    static_assert(N == 4 || N == 8 || N == 16, "unsupported column count");
    if constexpr (N == 4) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    } else if constexpr (N == 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
    } else {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
    }
}
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
// split-precision packing: x = hi + lo with hi = fp16(x), lo = fp16(x - hi); two values per 32-bit word
__device__ __forceinline__ void pack_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// address of the 16-byte piece `chunk` (0..7) of row r inside a K-major SWIZZLE_128B tile of 128-byte rows
__device__ __forceinline__ uint32_t sw128_off(int r, int chunk) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
}

#ifdef EXP_TRACE_FWD
__device__ long long g_trace_f[4 * 128];
#define FTRACE(slot) do { if (ftrace_on && (slot) < 128) g_trace_f[ftrace_base + (slot)] = clock64(); } while (0)
#else
#define FTRACE(slot) do { } while (0)
#endif
struct AttnArgs {
    void* o;            // bf16 [B*T, H*hd]  (SPLIT: fp32)
    float* lse;         // [B,H,T] natural log
    int T, H, Hkv, koff, voff;   // column offsets (elements) of the key / value blocks inside a qkv row
    float scale_log2;
    const float* qs;    // SPLIT: {scale, 1 / scale} of the qkv planes (device)
};

// Split-precision (parity) variants, SPLIT = true: qkv arrives as fp16 hi / lo planes of qkv * s (gaot_split_f16); every
// product runs as three tensor-core terms (lo*hi + hi*lo + hi*hi) into the same TMEM accumulator; the probabilities are
// split the same way (P * 2^7 = hi + lo in the forward, where the lazily rescaled P is < 2^8; P * 2^14 in the backward,
// where P <= 1); the softmax itself stays fp32.  Outputs are fp32.
constexpr float AT_PSCALE_FWD = 128.f;

// tcgen05.ld / tcgen05.st 32x32b of N (8 / 16 / 32) consecutive columns (issue only; pair with the matching wait)
template <int N>
__device__ __forceinline__ void tmem_ldN_issue(uint32_t taddr, uint32_t (&r)[N]) {
    static_assert(N == 8 || N == 16 || N == 32, "unsupported column count");
    if constexpr (N == 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    } else if constexpr (N == 16) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
    } else {
        tmem_ld32_issue(taddr, r);
    }
}
template <int N>
__device__ __forceinline__ void tmem_stN(uint32_t taddr, const uint32_t (&r)[N]) {
    static_assert(N == 8 || N == 16 || N == 32, "unsupported column count");
    if constexpr (N == 8) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    } else if constexpr (N == 16) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                     ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                       "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    } else {
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
            "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
            ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
              "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
              "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),
              "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// tcgen05.st of 8 consecutive columns, issue only (pair with tmem_wait_st)
__device__ __forceinline__ void tmem_st8_issue(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f32_(uint32_t a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }
__device__ __forceinline__ void sts_u4_(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// Forward, one CTA per (128-query tile, head, sample), two CTAs per SM:
//   warp 0    TMA producer: Q once, then K/V tiles of 128 keys through a 2-stage mbarrier ring
//   warp 1    MMA issuer  : S = Q K^T into TMEM; O += P V with the accumulator kept in TMEM for the whole row block
//   warps 2-9 softmax     : 2 warps per TMEM lane quadrant; a thread owns one query row and 64 of the 128 key columns of
//                           a tile: S -> registers (the S columns are released at once, so S(i+1) is computed during the
//                           softmax of tile i), row max exchanged with the partner warp through shared memory, base-2
//                           exponentials on MUFU, P (bf16) into one of two swizzled shared-memory tiles.
// The running maximum is only raised when the new row maximum exceeds it by more than 2^8 (lazy rescaling): the O
// accumulator is then rescaled in place in TMEM (tcgen05.ld / st); otherwise P is formed against the stale maximum, which
// is exact after the final division by the row sum.  No per-tile read-back of O.
// (hd = 64: the 144 KB tile set leaves room for one CTA per SM anyway, so the register cap of two would only cost spills)
template <int HD, bool SPLIT>
__global__ void __launch_bounds__(AT_THREADS, (SPLIT || HD > 32) ? 1 : 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm, const AttnArgs g) {
    constexpr int ROWB = HD * 2;                         // bytes per head row = swizzle span
    constexpr int NPL = SPLIT ? 2 : 1;                   // planes per operand tile (hi, lo)
    constexpr int PLANE_BYTES = 128 * ROWB;              // one plane of a 128-row Q / K / V tile
    constexpr int TILE_BYTES = NPL * PLANE_BYTES;        // a 128-row Q / K / V tile
    constexpr int PP_BYTES = AT_BQ * AT_BKV * 2;         // one plane of P: 32 KB, two 64-column SWIZZLE_128B blocks
    constexpr int P_BYTES = NPL * PP_BYTES;
    constexpr uint32_t TMEM_COLS = 256;                  // S: [0,128), O: [128,128+HD)
    constexpr int NO = HD / 2;                           // O columns per thread (the two warps of a quadrant split HD)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + TILE_BYTES;                       // [2 stages]
    uint8_t* sV = sK + 2 * TILE_BYTES;                   // [2 stages]
    uint8_t* sP = sV + 2 * TILE_BYTES;                   // [2 buffers]
    float* sX = reinterpret_cast<float*>(sP + 2 * P_BYTES);   // [2 parities][2 halves][128 rows] row-max / row-sum exchange
    uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 512);
    uint64_t* q_bar = bars;
    uint64_t* kv_full = bars + 1;      // [2]
    uint64_t* kv_empty = bars + 3;     // [2]
    uint64_t* s_full = bars + 5;       //      S(i) complete
    uint64_t* s_free = bars + 6;       //      8 warp arrivals: S(i) is in registers
    uint64_t* p_full = bars + 7;       //      8 warp arrivals: P(i) written
    uint64_t* p_free = bars + 8;       // [2]  PV(i) has finished reading P buffer i & 1
    uint64_t* o_step = bars + 10;      //      PV(i) complete (one phase per tile)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (g.H / g.Hkv);
    const int row_base = b * g.T;
    const int n_tiles = (g.T + AT_BKV - 1) / AT_BKV;

    if (threadIdx.x == 0) {
        mbar_init(q_bar, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); mbar_init(&p_free[s], 1); }
        mbar_init(s_full, 1); mbar_init(s_free, 8); mbar_init(p_full, 8); mbar_init(o_step, 1);   // 8: one arrival per softmax warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_s = tmem, tmem_o = tmem + 128;

    if (warp == 0) {
        if (lane == 0) {
            auto load = [&](uint8_t* dst, uint64_t* bar, int col, int row) {
                if constexpr (SPLIT) {
                    tma_load_3d(dst, &tm, bar, col, row, 0);
                    tma_load_3d(dst + PLANE_BYTES, &tm, bar, col, row, 1);
                } else {
                    tma_load_2d(dst, &tm, bar, col, row);
                }
            };
            mbar_expect_tx(q_bar, TILE_BYTES);
            load(sQ, q_bar, h * HD, row_base + t0);
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i & 1;
                mbar_wait(&kv_empty[s], ((uint32_t)(i >> 1) & 1u) ^ 1u);
                mbar_expect_tx(&kv_full[s], 2 * TILE_BYTES);
                load(sK + s * TILE_BYTES, &kv_full[s], g.koff + hk * HD, row_base + i * AT_BKV);
                load(sV + s * TILE_BYTES, &kv_full[s], g.voff + hk * HD, row_base + i * AT_BKV);
            }
        }
    } else if (warp == 1) {
        // S: M=128, N=128, A (Q) and B (K) K-major.  PV: M=128, N=HD, A (P) K-major, B (V) MN-major (bit 16).
        constexpr uint32_t FMT = SPLIT ? 0u : ((1u << 7) | (1u << 10));   // operand format: bf16 (1) or, for the split planes, fp16 (0)
        constexpr uint32_t IDESC_S = (1u << 4) | FMT | ((uint32_t)(AT_BKV >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t IDESC_PV = (1u << 4) | FMT | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr int NTERM = SPLIT ? 3 : 1;                 // SPLIT terms, small to large: (A lo, B hi), (A hi, B lo), (A hi, B hi)
        mbar_wait(q_bar, 0);
        auto issue_s = [&](int i) {
            const int s = i & 1;
            mbar_wait(&kv_full[s], (uint32_t)(i >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint64_t dq = make_desc<ROWB>(smem_u32(sQ)), dk = make_desc<ROWB>(smem_u32(sK + s * TILE_BYTES));
#pragma unroll
                for (int t = 0; t < NTERM; ++t) {
                    const uint64_t pa = (uint64_t)(((SPLIT && t == 0) ? PLANE_BYTES : 0) >> 4), pb = (uint64_t)(((SPLIT && t == 1) ? PLANE_BYTES : 0) >> 4);
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        umma_f16(tmem_s, dq + pa + (uint64_t)(2 * k), dk + pb + (uint64_t)(2 * k), IDESC_S, (t > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(s_full);
            }
            __syncwarp();
        };
        issue_s(0);
        for (int i = 0; i < n_tiles; ++i) {
            const int s = i & 1;
            if (i + 1 < n_tiles) {
                mbar_wait(s_free, (uint32_t)i & 1u);                 // S(i) is in the softmax warps' registers
                issue_s(i + 1);
            }
            mbar_wait(p_full, (uint32_t)i & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                // descriptors are base + compile-time offsets (start-address field counts 16-byte units)
                const uint64_t da0 = make_desc<128>(smem_u32(sP + s * P_BYTES)), db0 = make_desc<ROWB>(smem_u32(sV + s * TILE_BYTES));
#pragma unroll
                for (int t = 0; t < NTERM; ++t) {
                    const uint64_t pa = (uint64_t)(((SPLIT && t == 0) ? PP_BYTES : 0) >> 4), pb = (uint64_t)(((SPLIT && t == 1) ? PLANE_BYTES : 0) >> 4);
#pragma unroll
                    for (int k = 0; k < AT_BKV / 16; ++k) {
                        // A: 16 keys = 32 bytes inside the 128-byte swizzle row of P's column block (k / 4)
                        const uint64_t da = da0 + pa + (uint64_t)(((k >> 2) * (AT_BQ * 128) + (k & 3) * 32) >> 4);
                        // B: 16 keys = 16 rows of V, two 8-row swizzle atoms
                        const uint64_t db = db0 + pb + (uint64_t)((k * 16 * ROWB) >> 4);
                        umma_f16(tmem_o, da, db, IDESC_PV, (i > 0 || t > 0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(o_step);
                umma_commit(&p_free[s]);
                umma_commit(&kv_empty[s]);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3, hf = (warp - 2) >> 2;      // TMEM lane quadrant, column half of the tile
        const int r = q * 32 + lane;                       // query row inside the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t xs = smem_u32(sX);
        const uint32_t bar_id = 1 + q;                     // named barrier of the two warps that share this quadrant
        float mref = -INFINITY, lsum = 0.f;                // reference maximum (base-2 units), partial row sum (own columns)
        // SPLIT: the scores are those of (s q)(s k): fold 1 / s^2 into the softmax scale
        float sl2 = g.scale_log2, inv_s = 1.f;
        if constexpr (SPLIT) { inv_s = __ldg(g.qs + 1); sl2 *= inv_s * inv_s; }
#ifdef EXP_TRACE_FWD
        const bool ftrace_on = blockIdx.x == 0 && blockIdx.y == 0 && (blockIdx.z == 0 || blockIdx.z == 20) && lane == 0 && (warp == 2 || warp == 6);
        const int ftrace_base = (blockIdx.z == 0 ? 0 : 256) + (warp == 6 ? 128 : 0);
#endif
        for (int i = 0; i < n_tiles; ++i) {
            FTRACE(i * 8);
            mbar_wait(s_full, (uint32_t)i & 1u);
            FTRACE(i * 8 + 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t sv[2][32];
            tmem_ld32_issue(tmem_s + lane_addr + hf * 64, sv[0]);
            tmem_ld32_issue(tmem_s + lane_addr + hf * 64 + 32, sv[1]);
            tmem_wait_ld();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free);            // the MMA warp may overwrite S with tile i+1
            FTRACE(i * 8 + 2);
            const int kv_left = g.T - i * AT_BKV - hf * 64;   // of this thread's 64 columns, how many keys exist
            const bool full_tile = kv_left >= 64;          // warp-uniform: only the last tile can be ragged
            float mx = -INFINITY;                          // max of the RAW scores (scale > 0)
            if (full_tile) {
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
#pragma unroll
                for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], fmaxf(__uint_as_float(sv[0][j]), __uint_as_float(sv[1][j])));
                mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    mx = fmaxf(mx, (j < kv_left) ? __uint_as_float(sv[0][j]) : -INFINITY);
                    mx = fmaxf(mx, (32 + j < kv_left) ? __uint_as_float(sv[1][j]) : -INFINITY);
                }
            }
            // row maximum over all 128 columns: exchange with the partner warp (same rows, other column half)
            const uint32_t xa = xs + (uint32_t)(((i & 1) * 2 + hf) * 128 + r) * 4u, xb = xs + (uint32_t)(((i & 1) * 2 + (hf ^ 1)) * 128 + r) * 4u;
            sts_f32_(xa, mx);
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            const float mrow = fmaxf(mx, lds_f32(xb)) * sl2;
            FTRACE(i * 8 + 3);
            // lazy rescaling: identical decision in both warps of the quadrant (same rows, same maxima)
            if (__any_sync(0xffffffffu, mrow > mref + 8.f)) {
                const float mnew = fmaxf(mref, mrow);
                const float corr = ex2(mref - mnew);       // first tile: ex2(-inf) = 0
                lsum *= corr;
                if (i > 0) {
                    mbar_wait(o_step, (uint32_t)(i - 1) & 1u);        // PV(i-1) has landed in TMEM
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint32_t ov[NO];
                    tmem_ldN_issue<NO>(tmem_o + lane_addr + hf * NO, ov);
                    tmem_wait_ld();
#pragma unroll
                    for (int d = 0; d < NO; ++d) ov[d] = __float_as_uint(__uint_as_float(ov[d]) * corr);
                    tmem_stN<NO>(tmem_o + lane_addr + hf * NO, ov);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                }
                mref = mnew;
            }
            FTRACE(i * 8 + 4);
            if (i >= 2) mbar_wait(&p_free[i & 1], (uint32_t)((i - 2) >> 1) & 1u);   // PV(i-2) is done with this P buffer
            FTRACE(i * 8 + 5);
            const uint32_t pbase = smem_u32(sP) + (uint32_t)((i & 1) * P_BYTES + hf * (AT_BQ * 128));
            float ps4[4] = {0.f, 0.f, 0.f, 0.f};           // four independent summation chains
            // two copies of the loop behind ONE warp-uniform branch: the ragged-tile masks (a compare and a select per
            // score) otherwise run on every element of every tile and kept the ALU pipe as busy as the MUFU pipe
            auto emit_p = [&](auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int w = 0; w < 4; ++w) {          // 8 columns -> one 16-byte swizzled store
                        float p[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int j = w * 8 + k;
                            p[k] = ex2(fmaf(__uint_as_float(sv[c][j]), sl2, -mref));
                            if constexpr (!FULL) {
                                if (c * 32 + j >= kv_left) p[k] = 0.f;
                            }
                            ps4[k & 3] += p[k];
                        }
                        if constexpr (SPLIT) {
                            uint32_t ph[4], pl[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) pack_split2(p[2 * k] * AT_PSCALE_FWD, p[2 * k + 1] * AT_PSCALE_FWD, ph[k], pl[k]);
                            sts_u4_(pbase + sw128_off(r, c * 4 + w), ph[0], ph[1], ph[2], ph[3]);
                            sts_u4_(pbase + PP_BYTES + sw128_off(r, c * 4 + w), pl[0], pl[1], pl[2], pl[3]);
                        } else {
                            sts_u4_(pbase + sw128_off(r, c * 4 + w), pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]),
                                    pack_bf16(p[6], p[7]));
                        }
                    }
                }
            };
            if (full_tile) emit_p(std::true_type{});
            else emit_p(std::false_type{});
            lsum += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
            FTRACE(i * 8 + 6);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes of P -> async proxy (MMA)
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
            FTRACE(i * 8 + 7);
        }
        // epilogue: row sum over both halves, O / l, lse
        mbar_wait(o_step, (uint32_t)(n_tiles - 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t xa = xs + (uint32_t)((2 + hf) * 128 + r) * 4u, xb = xs + (uint32_t)((2 + (hf ^ 1)) * 128 + r) * 4u;
        // parity slot 1 of the exchange buffer may still be read by the partner for the last tile's maximum when the number
        // of tiles is even; one more named barrier orders that read before this write
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        sts_f32_(xa, lsum);
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        const float ltot = lsum + lds_f32(xb);
        uint32_t ov[NO];
        tmem_ldN_issue<NO>(tmem_o + lane_addr + hf * NO, ov);
        tmem_wait_ld();
        const int t = t0 + r;
        if (t < g.T) {
            if constexpr (SPLIT) {
                // O accumulated (2^7 P)(s V): undo both factors together with the row sum
                const float inv = inv_s / (ltot * AT_PSCALE_FWD);
                float* op = reinterpret_cast<float*>(g.o) + ((int64_t)(row_base + t) * g.H + h) * HD + hf * NO;
#pragma unroll
                for (int d = 0; d < NO; d += 4)
                    *reinterpret_cast<float4*>(op + d) = make_float4(__uint_as_float(ov[d]) * inv, __uint_as_float(ov[d + 1]) * inv,
                                                                     __uint_as_float(ov[d + 2]) * inv, __uint_as_float(ov[d + 3]) * inv);
            } else {
            const float inv = 1.f / ltot;
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(g.o) + ((int64_t)(row_base + t) * g.H + h) * HD + hf * NO;
#pragma unroll
            for (int d = 0; d < NO; d += 8)
                *reinterpret_cast<uint4*>(op + d) = make_uint4(
                    pack_bf16(__uint_as_float(ov[d]) * inv, __uint_as_float(ov[d + 1]) * inv), pack_bf16(__uint_as_float(ov[d + 2]) * inv, __uint_as_float(ov[d + 3]) * inv),
                    pack_bf16(__uint_as_float(ov[d + 4]) * inv, __uint_as_float(ov[d + 5]) * inv), pack_bf16(__uint_as_float(ov[d + 6]) * inv, __uint_as_float(ov[d + 7]) * inv));
            }
            if (hf == 0) g.lse[((int64_t)b * g.H + h) * g.T + t] = (mref + log2f(ltot)) * 0.69314718055994531f;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

// ------------------------------------------------------------------------------------------------
// Backward.  One CTA per (128-key tile, KV head, sample); it loops over the query heads of its group and
// over 128-query tiles.  Rows (TMEM lanes, threads) are KEYS for S^T, dP^T, dV, dK and QUERIES for dQ:
//   S^T  = K Q^T            dP^T = V dO^T                         (A, B K-major)                "ST(i)"
//   P^T  = exp2(S^T*c - lse2[q]),   dS^T = P^T * (dP^T - delta[q])   (per thread = per key row)   "compute(i)"
//   dV  += P^T dO           dK  += dS^T Q                         (A K-major from smem, B MN-major in place)
//   dQ   = dS K                                                    (A = dS^T tile read MN-major, B = K MN-major)  "DVKQ(i)"
// dV / dK accumulate in TMEM over all query tiles; dQ tiles are drained into an fp32 buffer with 128-bit
// red.add (different key tiles contribute to the same dQ rows).
//
// Software pipeline (the serial ST -> compute -> DVKQ -> drain chain of one step is spread over three):
//   * the compute warps pull their S^T / dP^T values into registers first and release the TMEM columns
//     (st_free), so the MMA warp issues ST(i+1) while compute(i) is still running;
//   * P^T / dS^T are double-buffered in shared memory and dQ is double-buffered in TMEM, so DVKQ(i) runs on the
//     tensor pipe during compute(i+1) and dQ(i) is drained after compute(i+1);
//   * Q / dO tiles go through a 3-stage TMA ring (stage i+1 is being multiplied while stage i-1 is still read).
// Tensor-pipe order: ST(0) ST(1) DVKQ(0) ST(2) DVKQ(1) ...
// ------------------------------------------------------------------------------------------------
#ifdef EXP_TRACE
__device__ long long g_trace[64 * 128];
#define TRACE(slot) do { if (trace_on && (slot) < 64) g_trace[trace_base + (slot)] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif
struct AttnBwdArgs {
    const float* lse;      // [B,H,T] natural log
    const float* delta;    // [B,H,T] rowsum(dO * O)
    float* dq_acc;         // fp32 [B*T, H*hd], zero-initialised  (SPLIT: unused, dQ is reduced straight into dqkv)
    void* dqkv;            // bf16 [B*T, ld]: dK and dV are written here (dQ is converted afterwards)  (SPLIT: fp32, Q columns zeroed)
    int T, H, Hkv, koff, voff, ld;
    float scale, scale_log2;
    const float* qs;       // SPLIT: {scale, 1 / scale} of the qkv planes
    const float* gs;       // SPLIT: {scale, 1 / scale} of the dO planes
};

// SPLIT (HD <= 32): every tile exists as a hi and a lo plane, so the P^T / dS^T tiles are single-buffered and Q / dO go
// through a 2-stage ring; with the per-warp statistics slab shrunk to one slab per column group and step parity the tile
// set is exactly the 227 KB a CTA may use.
//
// QW = queries per step.  128: one CTA per SM (TMEM 2 x 128 + 4 HD columns), 16 compute warps.  64 (bf16, HD <= 32): the
// step is 128 keys x 64 queries, TMEM 2 x 64 + 4 HD = 256 columns and 106 KB of tiles, so TWO CTAs of 8 compute warps share
// an SM and one CTA's exponentials run under the other's MMA / TMEM / reduction latencies; the dQ product then has M = 64.
template <int HD, bool SPLIT, int QW> struct AttnBwdCfg {
    static_assert(QW == 128 || (QW == 64 && !SPLIT && HD <= 32), "64-query steps: bf16, head dimension <= 32");
    static constexpr int CW = QW / 8;                                // compute warps: QW / 32 column groups x 4 TMEM lane quadrants
    static constexpr int THREADS = 64 + 32 * CW;                     // + TMA warp, MMA warp
    static constexpr int NPL = SPLIT ? 2 : 1;
    static constexpr int NQ = SPLIT ? 2 : (HD <= 32 ? 3 : 2);        // Q / dO stages
    static constexpr int NBUF = SPLIT ? 1 : (HD <= 32 ? 2 : 1);      // P^T / dS^T buffers (hd = 64: shared memory allows one)
    static constexpr int PLANE_BYTES = 128 * HD * 2;                 // K / V tile (128 keys)
    static constexpr int TILE_BYTES = NPL * PLANE_BYTES;
    static constexpr int QPLANE_BYTES = QW * HD * 2;                 // Q / dO tile (QW queries)
    static constexpr int QTILE_BYTES = NPL * QPLANE_BYTES;
    static constexpr int PP_BYTES = 128 * QW * 2;                    // P^T / dS^T tile: 128 keys x QW queries
    static constexpr int PT_BYTES = NPL * PP_BYTES;
    static constexpr int STAT_BYTES = SPLIT ? 2048 : CW * 256;
    // TPT (bf16, HD <= 32): P^T is handed to the dV product as a TENSOR-MEMORY A operand (tcgen05.st, QW/2 packed columns)
    // instead of a shared-memory tile; TDS (128-query steps, where the columns exist): dS^T likewise for the dK product - it
    // still goes to shared memory as well, for the transposed read of the dQ product.  The kernel is bound by the
    // shared-memory data pipe (LSU wavefronts + MMA operand fetches); per 128 x 128 scores this removes 32 KB of stores and
    // 32 (64 with TDS) KB of operand fetches.
    static constexpr bool TPT = !SPLIT && HD <= 32;
    static constexpr bool TDS = TPT && QW == 128;
    static constexpr bool DQ1 = QW == 64;                            // one dQ accumulator (TMEM columns), drained before the next emit
    static constexpr int ND = DQ1 ? 1 : 2;
    static constexpr bool TRED = TPT;                                // dQ tiles leave through one TMA reduction per step
    static constexpr int NBUF_P = TPT ? 0 : NBUF;                    // shared-memory P^T buffers
    static constexpr int DQ_BYTES = TRED ? QW * HD * 4 : 0;          // fp32 dQ tile staged for the TMA reduction
    static constexpr int SMEM = 2 * TILE_BYTES + 2 * NQ * QTILE_BYTES + (NBUF_P + NBUF) * PT_BYTES + DQ_BYTES + STAT_BYTES + 256 + (SPLIT ? 768 : 1024);
    static constexpr uint32_t TMEM_COLS = QW == 64 ? 256 : 512;
    static constexpr int CTAS = QW == 64 ? 2 : 1;                    // CTAs per SM
    static_assert(SMEM <= 227 * 1024 / CTAS - (CTAS - 1) * 1024, "backward attention tile set exceeds shared memory");
    static constexpr int ACC0 = 2 * QW + (TPT ? QW / 2 : 0) + (TDS ? QW / 2 : 0);   // first accumulator column
    static_assert(ACC0 + (2 + ND) * HD <= (int)TMEM_COLS, "backward attention accumulators exceed the TMEM allocation");
};
constexpr float AT_PSCALE_BWD = 16384.f;               // P <= 1 in the backward (true row statistics): P * 2^14 = hi + lo

// tm: qkv with a 128-row box (K / V tiles); tmq, tmdo: qkv and dO with a QW-row box (Q / dO tiles); tmacc (QW = 64): the
// fp32 dQ accumulator [B*T, H*hd] with a {hd, 64} box, swizzle span = hd * 4 bytes
template <int HD, bool SPLIT, int QW>
__global__ void __launch_bounds__((AttnBwdCfg<HD, SPLIT, QW>::THREADS), (AttnBwdCfg<HD, SPLIT, QW>::CTAS))
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmq,
                   const __grid_constant__ CUtensorMap tmdo, const __grid_constant__ CUtensorMap tmacc, const AttnBwdArgs g) {
    using Cfg = AttnBwdCfg<HD, SPLIT, QW>;
    constexpr int ROWB = HD * 2;
    constexpr int TILE_BYTES = Cfg::TILE_BYTES;
    constexpr int PLANE_BYTES = Cfg::PLANE_BYTES;
    constexpr int QTILE_BYTES = Cfg::QTILE_BYTES;
    constexpr int QPLANE_BYTES = Cfg::QPLANE_BYTES;
    constexpr int PT_BYTES = Cfg::PT_BYTES;
    constexpr int PP_BYTES = Cfg::PP_BYTES;
    constexpr int NQ = Cfg::NQ, NBUF = Cfg::NBUF, CW = Cfg::CW;
    constexpr bool TPT = Cfg::TPT, TDS = Cfg::TDS, TRED = Cfg::TRED;
    constexpr int ND = Cfg::ND;
    constexpr int NTERM = SPLIT ? 3 : 1;               // SPLIT terms, small to large: (A lo, B hi), (A hi, B lo), (A hi, B hi)
    constexpr uint32_t TMEM_COLS = Cfg::TMEM_COLS;     // S^T [0,QW) dP^T [QW,2QW) dV [2QW,..) dK [2QW+HD,..) dQ[2] [2QW+2HD,..)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sK = smem;
    uint8_t* sV = sK + TILE_BYTES;
    uint8_t* sQ = sV + TILE_BYTES;           // [NQ stages]
    uint8_t* sG = sQ + NQ * QTILE_BYTES;     // dO, [NQ stages]
    uint8_t* sPt = sG + NQ * QTILE_BYTES;    // [NBUF]
    uint8_t* sSt = sPt + Cfg::NBUF_P * PT_BYTES;   // [NBUF]
    // statistics slabs [lse2 x 32 | delta x 32]: one per compute warp, or (SPLIT) one per column group and step parity,
    // shared by the four warps of the group (they write identical values)
    uint8_t* sDq = sSt + NBUF * PT_BYTES;    // QW = 64: fp32 dQ tile [64 queries][HD], rows swizzled like the tensor map
    float* sStat = reinterpret_cast<float*>(sDq + Cfg::DQ_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStat) + Cfg::STAT_BYTES);
    if (SPLIT && (uintptr_t)(smem - smem_raw) > 768) __trap();      // the split tile set leaves 768 bytes of alignment slack
    uint64_t* kv_bar = bars;
    uint64_t* q_full = bars + 1;     // [NQ]   TMA -> MMA
    uint64_t* q_empty = bars + 4;    // [NQ]   DVKQ(i) finished reading stage i % NQ
    uint64_t* st_full = bars + 7;    //        ST(i) complete
    uint64_t* st_free = bars + 8;    //        8 warp arrivals: S^T / dP^T of step i are in registers
    uint64_t* pds_full = bars + 9;   // [2]    8 warp arrivals: P^T / dS^T buffer written
    uint64_t* done = bars + 11;      // [2]    DVKQ(i) complete: dQ[i & 1] ready, P^T / dS^T buffer free
    uint64_t* dq_empty = bars + 13;  // [2]    8 warp arrivals: dQ[i & 1] drained
    uint64_t* stage_full = bars + 15; //       CW warp arrivals: the dQ tile is staged in shared memory
    uint64_t* stage_free = bars + 16; //       the TMA reduction has read the staged tile
    uint64_t* pt_free = bars + 17;    //       TPT: the dV / dK products of a step have read P^T / dS^T from tensor memory
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * 128, hk = blockIdx.y, b = blockIdx.z;
    const int rep = g.H / g.Hkv;
    const int row_base = b * g.T;
    const int n_qt = (g.T + QW - 1) / QW;
    const int n_steps = rep * n_qt;
#ifdef EXP_TRACE
    // compute warp 2 lane 0 -> slots [0,64), MMA warp lane 0 -> slots [64,128) of CTA (0,0,z)
    const bool trace_on = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z < 64 && lane == 0 && (warp == 2 || warp == 1);
    const int trace_base = blockIdx.z * 128 + (warp == 1 ? 64 : 0);
    TRACE(0);
#endif

    if (threadIdx.x == 0) {
        mbar_init(kv_bar, 1);
        for (int s = 0; s < NQ; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
        mbar_init(st_full, 1); mbar_init(st_free, CW);               // one arrival per compute warp
        for (int s = 0; s < 2; ++s) { mbar_init(&pds_full[s], CW); mbar_init(&done[s], 1); mbar_init(&dq_empty[s], CW); }
        mbar_init(stage_full, CW); mbar_init(stage_free, 1); mbar_init(pt_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    TRACE(1);
    constexpr uint32_t ACC0 = Cfg::ACC0;   // TPT: P^T (bf16 pairs) in [2QW, 2QW + QW/2); TDS: dS^T in [2QW + QW/2, 3QW)
    const uint32_t tm_st = tmem, tm_dp = tmem + QW, tm_pt = tmem + 2 * QW, tm_ds = tmem + 2 * QW + QW / 2;
    const uint32_t tm_dv = tmem + ACC0, tm_dk = tmem + ACC0 + HD, tm_dq = tmem + ACC0 + 2 * HD;

    if (warp == 0) {
        if (lane == 0) {
            auto load = [&](uint8_t* dst, int plane_bytes, const CUtensorMap* map, uint64_t* bar, int col, int row) {
                if constexpr (SPLIT) {
                    tma_load_3d(dst, map, bar, col, row, 0);
                    tma_load_3d(dst + plane_bytes, map, bar, col, row, 1);
                } else {
                    tma_load_2d(dst, map, bar, col, row);
                }
            };
            mbar_expect_tx(kv_bar, 2 * TILE_BYTES);
            load(sK, PLANE_BYTES, &tm, kv_bar, g.koff + hk * HD, row_base + k0);
            load(sV, PLANE_BYTES, &tm, kv_bar, g.voff + hk * HD, row_base + k0);
            int s = 0, use = 0, hq = 0, qt = 0;
            for (int i = 0; i < n_steps; ++i) {
                mbar_wait(&q_empty[s], ((uint32_t)use & 1u) ^ 1u);
                mbar_expect_tx(&q_full[s], 2 * QTILE_BYTES);
                load(sQ + s * QTILE_BYTES, QPLANE_BYTES, &tmq, &q_full[s], (hk * rep + hq) * HD, row_base + qt * QW);
                load(sG + s * QTILE_BYTES, QPLANE_BYTES, &tmdo, &q_full[s], (hk * rep + hq) * HD, row_base + qt * QW);
                if (++s == NQ) { s = 0; ++use; }
                if (++qt == n_qt) { qt = 0; ++hq; }
            }
        } else if (TRED && lane == 1) {
            // second lane of the producer warp: one TMA reduction per step moves the dQ tile staged by the compute warps
            // into the fp32 accumulator (off the compute warps' critical path), then hands the staging tile back
            int hq = 0, qt = 0;
            for (int j = 0; j < n_steps; ++j) {
                mbar_wait(stage_full, (uint32_t)j & 1u);
                tma_reduce_add_2d(&tmacc, sDq, (hk * rep + hq) * HD, row_base + qt * QW);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive(stage_free);
                if (++qt == n_qt) { qt = 0; ++hq; }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp == 1) {
        constexpr uint32_t FMT = (1u << 4) | (SPLIT ? 0u : ((1u << 7) | (1u << 10)));                // fp16 planes when SPLIT
        constexpr uint32_t BASE = FMT | ((uint32_t)(128 >> 4) << 24);                                // M = 128 keys
        constexpr uint64_t PL = (uint64_t)(PLANE_BYTES >> 4), QPL = (uint64_t)(QPLANE_BYTES >> 4), PPL = (uint64_t)(PP_BYTES >> 4);   // lo-plane offsets in descriptor units
        constexpr uint32_t IDESC_ST = BASE | ((uint32_t)(QW >> 3) << 17);                            // N=QW, K-major x K-major
        constexpr uint32_t IDESC_KN = BASE | (1u << 16) | ((uint32_t)(HD >> 3) << 17);               // N=HD, B MN-major
        constexpr uint32_t IDESC_MN = FMT | ((uint32_t)(QW >> 4) << 24) | (1u << 15) | (1u << 16) | ((uint32_t)(HD >> 3) << 17);  // M=QW queries, A and B MN-major
        mbar_wait(kv_bar, 0);
        const uint64_t d_k = make_desc<ROWB>(smem_u32(sK)), d_v = make_desc<ROWB>(smem_u32(sV));
        auto issue_st = [&](int i) {
            const int s = i % NQ;
            mbar_wait(&q_full[s], (uint32_t)(i / NQ) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint64_t dq = make_desc<ROWB>(smem_u32(sQ + s * QTILE_BYTES)), dg = make_desc<ROWB>(smem_u32(sG + s * QTILE_BYTES));
#pragma unroll
                for (int t = 0; t < NTERM; ++t) {
                    const uint64_t pa = (SPLIT && t == 0) ? PL : 0, pb = (SPLIT && t == 1) ? QPL : 0;
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        umma_f16(tm_st, d_k + pa + (uint64_t)(2 * k), dq + pb + (uint64_t)(2 * k), IDESC_ST, (t > 0 || k > 0) ? 1u : 0u);
                }
#pragma unroll
                for (int t = 0; t < NTERM; ++t) {
                    const uint64_t pa = (SPLIT && t == 0) ? PL : 0, pb = (SPLIT && t == 1) ? QPL : 0;
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        umma_f16(tm_dp, d_v + pa + (uint64_t)(2 * k), dg + pb + (uint64_t)(2 * k), IDESC_ST, (t > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(st_full);
            }
            __syncwarp();
        };
        issue_st(0);
        for (int i = 0; i < n_steps; ++i) {
            TRACE(2 + 4 * i);
            if (i + 1 < n_steps) {
                mbar_wait(st_free, (uint32_t)i & 1u);                     // S^T(i) / dP^T(i) are in the compute warps' registers
                TRACE(3 + 4 * i);
                issue_st(i + 1);
            }
            TRACE(4 + 4 * i);
            const int s = i % NQ, bf = i % NBUF, d = i % ND;
            mbar_wait(&pds_full[bf], (uint32_t)(i / NBUF) & 1u);
            if constexpr (!Cfg::DQ1) {
                if (i >= ND) mbar_wait(&dq_empty[d], (uint32_t)((i - ND) / ND) & 1u);   // dQ[d] of step i-ND drained from TMEM
            }
            TRACE(5 + 4 * i);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                // all descriptors = per-step base + compile-time offset (16-byte units in the start-address field)
                const uint64_t d_q = make_desc<ROWB>(smem_u32(sQ + s * QTILE_BYTES)), d_g = make_desc<ROWB>(smem_u32(sG + s * QTILE_BYTES));
                const uint32_t pst = smem_u32(sSt + bf * PT_BYTES);
                const uint64_t d_pt = make_desc<128>(smem_u32(sPt + bf * PT_BYTES)), d_st = make_desc<128>(pst);
                const uint32_t acc_first = i > 0 ? 1u : 0u;
                if constexpr (TPT) {
#pragma unroll
                    for (int k = 0; k < QW / 16; ++k) {   // A: 16 queries = 8 packed tensor-memory columns
                        const uint64_t boff = (uint64_t)((k * 16 * ROWB) >> 4);
                        umma_f16_ts(tm_dv, tm_pt + (uint32_t)(k * 8), d_g + boff, IDESC_KN, k > 0 ? 1u : acc_first);
                        if constexpr (TDS) umma_f16_ts(tm_dk, tm_ds + (uint32_t)(k * 8), d_q + boff, IDESC_KN, k > 0 ? 1u : acc_first);
                    }
                    umma_commit(pt_free);
                    if constexpr (!TDS) {
#pragma unroll
                        for (int k = 0; k < QW / 16; ++k)
                            umma_f16(tm_dk, d_st + (uint64_t)(((k >> 2) * (128 * 128) + (k & 3) * 32) >> 4), d_q + (uint64_t)((k * 16 * ROWB) >> 4),
                                     IDESC_KN, k > 0 ? 1u : acc_first);
                    }
                } else {
#pragma unroll
                for (int t = 0; t < NTERM; ++t) {
                    const uint64_t pa = (SPLIT && t == 0) ? PPL : 0, pb = (SPLIT && t == 1) ? QPL : 0;
#pragma unroll
                    for (int k = 0; k < QW / 16; ++k) {   // contraction over the QW queries, 16 per MMA
                        const uint64_t aoff = pa + (uint64_t)(((k >> 2) * (128 * 128) + (k & 3) * 32) >> 4);
                        const uint64_t boff = pb + (uint64_t)((k * 16 * ROWB) >> 4);
                        umma_f16(tm_dv, d_pt + aoff, d_g + boff, IDESC_KN, (t > 0 || k > 0) ? 1u : acc_first);
                        umma_f16(tm_dk, d_st + aoff, d_q + boff, IDESC_KN, (t > 0 || k > 0) ? 1u : acc_first);
                    }
                }
                }
                if constexpr (Cfg::DQ1) {                   // single dQ accumulator: step i-1 must have been drained (the dV / dK
                    if (i >= 1) {                           // products above are already queued on the tensor pipe)
                        mbar_wait(&dq_empty[0], (uint32_t)(i - 1) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                }
                // dQ[q, :] = sum_keys dS^T[key, q] K[key, :]: A = the dS^T tile as an MN-major operand
                // (M = QW queries = 64-wide swizzle atoms LBO = 16 KB apart, 8-key groups SBO = 1 KB apart)
                const uint64_t a_mn0 = (uint64_t)((pst & 0x3FFFF) >> 4) | ((uint64_t)(16384 >> 4) << 16) |
                                       ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
#pragma unroll
                for (int t = 0; t < NTERM; ++t) {
                    const uint64_t pa = (SPLIT && t == 0) ? PPL : 0, pb = (SPLIT && t == 1) ? PL : 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k)        // contraction over 128 keys, 16 per MMA = 16 rows of 128 bytes
                        umma_f16(tm_dq + d * HD, a_mn0 + pa + (uint64_t)((k * 2048) >> 4), d_k + pb + (uint64_t)((k * 16 * ROWB) >> 4), IDESC_MN,
                                 (t > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&done[d]);
                umma_commit(&q_empty[s]);
            }
            __syncwarp();
        }
    } else {
        // CW compute warps: warps w, w+4, .. share TMEM lane quadrant w%4 (the hardware restricts a warp to the lanes
        // 32*(w%4)..+31) and split the columns: `qr` owns query columns [32*qr, 32*qr+32) of S^T / dP^T (= half of a
        // 64-column swizzle block of the P^T / dS^T tiles) and a 4/CW share of the dQ / dV / dK columns.
        constexpr int NC = HD / (CW / 4);                  // dQ / dV / dK columns per thread
        const int q4 = warp & 3, qr = (warp - 2) >> 2;
        const int r = q4 * 32 + lane;                      // key row for S^T.., query row for dQ
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        const bool key_ok = k0 + r < g.T;
        const bool tile_full = k0 + 128 <= g.T;            // CTA-uniform: only the last key tile can be ragged
        // per-query statistics of the warp's 32 columns: [lse2 x 32 | delta x 32] in a per-warp shared-memory slab
        // (no CTA-wide barrier), loaded from global memory one step ahead (one lse and one delta per lane)
        uint32_t stat_s = smem_u32(sStat) + (uint32_t)(SPLIT ? qr * 2 : warp - 2) * 256u;   // SPLIT: + 256 on odd steps (set per step)
        const uint32_t stat_s0 = stat_s;
        // SPLIT scale bookkeeping (all factors are powers of two except the softmax scale):
        //   scores arrive as (s q).(s k), dP^T as (s v).(sg dO);   dS'' = P (dP' kd - delta kdel) = (sigma s sg) dS,  sigma = 2^-14 / HD
        //   keeps |dS''| < 2^15 for any input (|dP'| <= HD 2^28) while typical entries stay ~2^40 above the fp16 hi / lo error floor
        float sl2 = g.scale_log2, kd = 1.f, kdel = 1.f, f_dqk = g.scale, f_dv = 1.f;
        if constexpr (SPLIT) {
            const float sq = __ldg(g.qs), inv_s = __ldg(g.qs + 1), sg = __ldg(g.gs), inv_g = __ldg(g.gs + 1);
            constexpr float SIGMA = 1.f / (16384.f * (float)HD);
            sl2 *= inv_s * inv_s;
            kd = SIGMA;
            kdel = sq * sg * SIGMA;
            f_dqk = g.scale * (inv_s * inv_s) * inv_g * (16384.f * (float)HD);
            f_dv = inv_g * (1.f / AT_PSCALE_BWD);
        }
        const float* lse_b = g.lse + (int64_t)b * g.H * g.T;
        const float* del_b = g.delta + (int64_t)b * g.H * g.T;
        float st_l, st_d;
        auto fetch_stat = [&](int h, int qt) {
            const int q = qt * QW + qr * 32 + lane;
            const int64_t off = (int64_t)h * g.T + q;
            st_l = ldg_f32_pred(lse_b + off, q < g.T, INFINITY);           // +inf switches a query column off
            st_d = ldg_f32_pred(del_b + off, q < g.T, 0.f);
        };
        // dQ of a finished step: here the TMEM lane is the QUERY row; this warp's share of the HD columns.  QW = 64: an
        // M = 64 accumulator keeps row m in lane 32 * (m / 16) + m % 16, i.e. the lower half of each warp's lanes
        auto drain_dq = [&](int j, int h, int qt) {
            const int t = QW == 64 ? (lane < 16 ? qt * QW + q4 * 16 + lane : g.T) : qt * QW + r, d = j % ND;
            mbar_wait(&done[d], (uint32_t)(j / ND) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[NC];
            tmem_ld_issue<NC>(tm_dq + d * HD + lane_addr + qr * NC, v);
            tmem_wait_ld();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&dq_empty[d]);
            if constexpr (TRED) {
                // the tile goes to global memory as ONE TMA reduction (a per-thread red.add touches a different 128-byte
                // line in every lane: 32 LSU wavefronts per instruction), issued by the producer warp once all CW warps
                // have staged their share; it hands the staging tile back when the reduction has read it.
                constexpr int RB = HD * 4;                               // bytes per staged row = swizzle span
                if (j >= 1) mbar_wait(stage_free, (uint32_t)(j - 1) & 1u);
                if (QW == 128 || lane < 16) {
                    const uint32_t rr = (uint32_t)(QW == 64 ? q4 * 16 + lane : r);
#pragma unroll
                    for (int j4 = 0; j4 < NC; j4 += 4) {
                        uint32_t a = rr * RB + (uint32_t)(qr * NC + j4) * 4u;
                        a ^= ((a >> 7) & (uint32_t)(RB / 16 - 1)) << 4;
                        sts_u4(smem_u32(sDq) + a, __float_as_uint(__uint_as_float(v[j4]) * f_dqk), __float_as_uint(__uint_as_float(v[j4 + 1]) * f_dqk),
                               __float_as_uint(__uint_as_float(v[j4 + 2]) * f_dqk), __float_as_uint(__uint_as_float(v[j4 + 3]) * f_dqk));
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(stage_full);
                return;
            }
#ifdef EXP_NO_RED
            if (t < 0) {
#else
            if (t < g.T) {
#endif
                float* dst = SPLIT ? reinterpret_cast<float*>(g.dqkv) + (int64_t)(row_base + t) * g.ld + h * HD + qr * NC
                                   : g.dq_acc + ((int64_t)(row_base + t) * g.H + h) * HD + qr * NC;
#pragma unroll
                for (int j4 = 0; j4 < NC; j4 += 4)
                    atomicAdd(reinterpret_cast<float4*>(dst + j4),
                              make_float4(__uint_as_float(v[j4]) * f_dqk, __uint_as_float(v[j4 + 1]) * f_dqk,
                                          __uint_as_float(v[j4 + 2]) * f_dqk, __uint_as_float(v[j4 + 3]) * f_dqk));
            }
        };
        // P^T / dS^T of the warp's 32 query columns: exp2, products, bf16 packing; 8 columns = one 16-byte swizzled
        // store into each tile
        auto emit = [&](auto full_tag, const uint32_t (&sv)[32], const uint32_t (&dv)[32], uint32_t bp, uint32_t bs) {
            constexpr bool FULL = decltype(full_tag)::value;
            [[maybe_unused]] uint32_t tp[8], td[8];        // TPT: 16 columns of P^T / dS^T as bf16 pairs -> one tcgen05.st each
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint32_t pp[4], ds[4];
                [[maybe_unused]] uint32_t ppl[4], dsl[4];
#pragma unroll
                for (int k = 0; k < 8; k += 4) {
#ifdef EXP_NO_LDS
                    const float4 lv = make_float4(st_l, st_l, st_l, st_l), dl = make_float4(st_d, st_d, st_d, st_d);
#else
                    const float4 lv = lds_f4(stat_s + (uint32_t)(j + k) * 4u);              // broadcast reads
                    const float4 dl = lds_f4(stat_s + 128u + (uint32_t)(j + k) * 4u);
#endif
                    float p0 = ex2(fmaf(__uint_as_float(sv[j + k]), sl2, -lv.x));
                    float p1 = ex2(fmaf(__uint_as_float(sv[j + k + 1]), sl2, -lv.y));
                    float p2 = ex2(fmaf(__uint_as_float(sv[j + k + 2]), sl2, -lv.z));
                    float p3 = ex2(fmaf(__uint_as_float(sv[j + k + 3]), sl2, -lv.w));
                    if (!FULL) {
                        if (!key_ok) { p0 = 0.f; p1 = 0.f; p2 = 0.f; p3 = 0.f; }
                    }
                    if constexpr (SPLIT) {
                        pack_split2(p0 * AT_PSCALE_BWD, p1 * AT_PSCALE_BWD, pp[k >> 1], ppl[k >> 1]);
                        pack_split2(p2 * AT_PSCALE_BWD, p3 * AT_PSCALE_BWD, pp[(k >> 1) + 1], ppl[(k >> 1) + 1]);
                        pack_split2(p0 * fmaf(__uint_as_float(dv[j + k]), kd, -dl.x * kdel), p1 * fmaf(__uint_as_float(dv[j + k + 1]), kd, -dl.y * kdel),
                                    ds[k >> 1], dsl[k >> 1]);
                        pack_split2(p2 * fmaf(__uint_as_float(dv[j + k + 2]), kd, -dl.z * kdel), p3 * fmaf(__uint_as_float(dv[j + k + 3]), kd, -dl.w * kdel),
                                    ds[(k >> 1) + 1], dsl[(k >> 1) + 1]);
                    } else {
                    pp[k >> 1] = pack_bf16(p0, p1);
                    pp[(k >> 1) + 1] = pack_bf16(p2, p3);
                    ds[k >> 1] = pack_bf16(p0 * (__uint_as_float(dv[j + k]) - dl.x), p1 * (__uint_as_float(dv[j + k + 1]) - dl.y));
                    ds[(k >> 1) + 1] = pack_bf16(p2 * (__uint_as_float(dv[j + k + 2]) - dl.z), p3 * (__uint_as_float(dv[j + k + 3]) - dl.w));
                    }
                }
                const uint32_t off = sw128_off(r, (qr & 1) * 4 + (j >> 3));
                if constexpr (TPT) {
                    sts_u4(bs + off, ds[0], ds[1], ds[2], ds[3]);          // dS^T tile: transposed read of the dQ product (and dK without TDS)
#pragma unroll
                    for (int k = 0; k < 4; ++k) { tp[(j & 8) / 2 + k] = pp[k]; td[(j & 8) / 2 + k] = ds[k]; }
                    if (j & 8) {
                        tmem_st8_issue(tm_pt + lane_addr + (uint32_t)(qr * 16 + (j >> 4) * 8), tp);
                        if constexpr (TDS) tmem_st8_issue(tm_ds + lane_addr + (uint32_t)(qr * 16 + (j >> 4) * 8), td);
                    }
                } else {
#ifdef EXP_NO_STS
                if (pp[0] == 0x12345u && ds[1] == 0x777u)
#endif
                {
                sts_u4(bp + off, pp[0], pp[1], pp[2], pp[3]);
                sts_u4(bs + off, ds[0], ds[1], ds[2], ds[3]);
                }
                }
                if constexpr (SPLIT) {
                    sts_u4(bp + PP_BYTES + off, ppl[0], ppl[1], ppl[2], ppl[3]);
                    sts_u4(bs + PP_BYTES + off, dsl[0], dsl[1], dsl[2], dsl[3]);
                }
            }
        };
        const int h0 = hk * rep;
        int f_h = h0, f_qt = 0;                             // step being fetched (one ahead)
        int c_h = h0, c_qt = 0, p_h = h0, p_qt = 0;         // current / previous step
        fetch_stat(f_h, f_qt);
        for (int i = 0; i < n_steps; ++i) {
            const int bf = i % NBUF;
            if constexpr (SPLIT) {
                // one slab per column group and step parity: the group's first warp writes it, a 128-thread named barrier
                // publishes it to the other three (the barrier of step i+1 also orders their reads of step i before the
                // rewrite of this parity's slab at step i+2)
                stat_s = stat_s0 + (uint32_t)(i & 1) * 256u;
                if (q4 == 0) {
                    sts_f32(stat_s + (uint32_t)lane * 4u, st_l * 1.4426950408889634f);
                    sts_f32(stat_s + 128u + (uint32_t)lane * 4u, st_d);
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + qr) : "memory");
            } else {
                sts_f32(stat_s + (uint32_t)lane * 4u, st_l * 1.4426950408889634f);
                sts_f32(stat_s + 128u + (uint32_t)lane * 4u, st_d);
                __syncwarp();
            }
            if (i + 1 < n_steps) {
                if (++f_qt == n_qt) { f_qt = 0; ++f_h; }
                fetch_stat(f_h, f_qt);
            }
            TRACE(2 + 6 * i);
            mbar_wait(st_full, (uint32_t)i & 1u);
            TRACE(3 + 6 * i);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t sv[32], dv[32];
            tmem_ld32_issue(tm_st + lane_addr + qr * 32, sv);
            tmem_ld32_issue(tm_dp + lane_addr + qr * 32, dv);
            tmem_wait_ld();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(st_free);           // the MMA warp may overwrite S^T / dP^T with step i+1
            TRACE(4 + 6 * i);
            if constexpr (!Cfg::DQ1) {                      // (DQ1: implied by the drain of step i-2, which waited for done(i-2))
                if (i >= NBUF) {                            // DVKQ(i - NBUF) has finished reading this P^T / dS^T buffer
                    mbar_wait(&done[(i - NBUF) % ND], (uint32_t)((i - NBUF) / ND) & 1u);
                }
            }
            if constexpr (TPT) {
                if (i >= 1) {                               // the dV / dK products of step i-1 have read P^T / dS^T from TMEM
                    mbar_wait(pt_free, (uint32_t)(i - 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
            }
            TRACE(5 + 6 * i);
            const uint32_t bp = smem_u32(sPt) + (uint32_t)(bf * PT_BYTES + (qr >> 1) * (128 * 128));   // 64-query swizzle block
            const uint32_t bs = smem_u32(sSt) + (uint32_t)(bf * PT_BYTES + (qr >> 1) * (128 * 128));
            if (tile_full) emit(std::true_type{}, sv, dv, bp, bs);
            else emit(std::false_type{}, sv, dv, bp, bs);
            if constexpr (TPT) {
                tmem_wait_st();
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&pds_full[bf]);
            TRACE(6 + 6 * i);
            if (i >= 1) drain_dq(i - 1, p_h, p_qt);
            TRACE(7 + 6 * i);
            p_h = c_h; p_qt = c_qt;
            if (++c_qt == n_qt) { c_qt = 0; ++c_h; }
        }
        drain_dq(n_steps - 1, p_h, p_qt);
        TRACE(2 + 6 * n_steps);
        // final: dV and dK rows of this key tile (done[] of the last step has completed: every MMA has retired)
        if constexpr (SPLIT) {
            float* base = reinterpret_cast<float*>(g.dqkv) + (int64_t)(row_base + k0 + r) * g.ld;
            uint32_t vv[NC], vk[NC];
            tmem_ld_issue<NC>(tm_dv + lane_addr + qr * NC, vv);
            tmem_ld_issue<NC>(tm_dk + lane_addr + qr * NC, vk);
            tmem_wait_ld();
            if (key_ok) {
                float* dvp = base + g.voff + hk * HD + qr * NC;
                float* dkp = base + g.koff + hk * HD + qr * NC;
#pragma unroll
                for (int c = 0; c < NC; c += 4) {
                    *reinterpret_cast<float4*>(dvp + c) = make_float4(__uint_as_float(vv[c]) * f_dv, __uint_as_float(vv[c + 1]) * f_dv,
                                                                      __uint_as_float(vv[c + 2]) * f_dv, __uint_as_float(vv[c + 3]) * f_dv);
                    *reinterpret_cast<float4*>(dkp + c) = make_float4(__uint_as_float(vk[c]) * f_dqk, __uint_as_float(vk[c + 1]) * f_dqk,
                                                                      __uint_as_float(vk[c + 2]) * f_dqk, __uint_as_float(vk[c + 3]) * f_dqk);
                }
            }
        } else {
            __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(g.dqkv) + (int64_t)(row_base + k0 + r) * g.ld;
            uint32_t vv[NC], vk[NC];
            tmem_ld_issue<NC>(tm_dv + lane_addr + qr * NC, vv);
            tmem_ld_issue<NC>(tm_dk + lane_addr + qr * NC, vk);
            tmem_wait_ld();
            if (key_ok) {
                uint32_t pv[NC / 2], pk[NC / 2];
#pragma unroll
                for (int c = 0; c < NC; c += 2) {
                    pv[c >> 1] = pack_bf16(__uint_as_float(vv[c]), __uint_as_float(vv[c + 1]));
                    pk[c >> 1] = pack_bf16(__uint_as_float(vk[c]) * g.scale, __uint_as_float(vk[c + 1]) * g.scale);
                }
                __nv_bfloat16* dvp = base + g.voff + hk * HD + qr * NC;
                __nv_bfloat16* dkp = base + g.koff + hk * HD + qr * NC;
                if constexpr (NC == 4) {
                    *reinterpret_cast<uint2*>(dvp) = make_uint2(pv[0], pv[1]);
                    *reinterpret_cast<uint2*>(dkp) = make_uint2(pk[0], pk[1]);
                } else {
#pragma unroll
                    for (int c = 0; c < NC / 2; c += 4) {
                        *reinterpret_cast<uint4*>(dvp + 2 * c) = make_uint4(pv[c], pv[c + 1], pv[c + 2], pv[c + 3]);
                        *reinterpret_cast<uint4*>(dkp + 2 * c) = make_uint4(pk[c], pk[c + 1], pk[c + 2], pk[c + 3]);
                    }
                }
            }
        }
    }
    TRACE(62);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
    TRACE(63);
}

// delta[b,h,t] = sum_d dO[t,h,d] * O[t,h,d]   (bf16 in, fp32 out); one thread per (token, head)
template <int HD>
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                  float* __restrict__ delta, int64_t B, int64_t T, int H) {
    const int64_t total = B * T * H;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int h = (int)(i % H);
        const int64_t bt = i / H;
        const uint4* po = reinterpret_cast<const uint4*>(o + i * HD);
        const uint4* pg = reinterpret_cast<const uint4*>(dout + i * HD);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) {
            const uint4 a = __ldg(po + c), g4 = __ldg(pg + c);
            const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a);
            const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&g4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 fx = __bfloat1622float2(x[k]), fy = __bfloat1622float2(y[k]);
                acc = fmaf(fx.x, fy.x, acc); acc = fmaf(fx.y, fy.y, acc);
            }
        }
        delta[((bt / T) * H + h) * T + (bt % T)] = acc;
    }
}

// fp32 variant for the split-precision path
template <int HD>
__global__ void attn_delta_f32_kernel(const float* __restrict__ o, const float* __restrict__ dout, float* __restrict__ delta,
                                      int64_t B, int64_t T, int H) {
    const int64_t total = B * T * H;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int h = (int)(i % H);
        const int64_t bt = i / H;
        const float4* po = reinterpret_cast<const float4*>(o + i * HD);
        const float4* pg = reinterpret_cast<const float4*>(dout + i * HD);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < HD / 4; ++c) {
            const float4 a = __ldg(po + c), b4 = __ldg(pg + c);
            acc = fmaf(a.x, b4.x, acc); acc = fmaf(a.y, b4.y, acc); acc = fmaf(a.z, b4.z, acc); acc = fmaf(a.w, b4.w, acc);
        }
        delta[((bt / T) * H + h) * T + (bt % T)] = acc;
    }
}

// dqkv[row, 0:W] = bf16(dq_acc[row, 0:W])   (W = H*hd query columns of the fused gradient)
__global__ void attn_dq_store_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, int64_t rows, int W, int ld) {
    const int64_t total = rows * (W / 4);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / (W / 4);
        const int c = (int)(i % (W / 4)) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(acc + row * W + c));
        uint2 pk;
        pk.x = pack_bf16(v.x, v.y); pk.y = pack_bf16(v.z, v.w);
        *reinterpret_cast<uint2*>(dqkv + row * ld + c) = pk;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows, ld] bf16 matrix, box {hd, box_rows}, swizzle span = hd*2 bytes
// the driver-API encode needs a current context: the autograd engine's threads have none until a runtime call binds it
inline void bind_context_once() {
    static thread_local bool bound = false;
    if (!bound) { cudaFree(nullptr); bound = true; }
}

int make_head_map(CUtensorMap* map, const void* base, int64_t rows, int64_t ld, int hd, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return GAOT_E_BADARG;
    bind_context_once();
    cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)hd, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = hd == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (hd == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT) {
        // driver API called on a thread (e.g. the autograd engine's) that has no current context yet: a runtime
        // call binds the primary context, then encode again
        cudaFree(nullptr);
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS && getenv("GAOT_DEBUG"))
        fprintf(stderr, "gaot: cuTensorMapEncodeTiled failed (%d): base %p rows %lld ld %lld hd %d box_rows %d map %p\n", (int)r, base,
                (long long)rows, (long long)ld, hd, box_rows, (void*)map);
    return r == CUDA_SUCCESS ? 0 : GAOT_E_BADARG;
}

// fp32 [rows, cols] accumulator, box {hd, box_rows}, swizzle span = hd*4 bytes (the staged dQ tile of the backward)
int make_acc_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int hd, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return GAOT_E_BADARG;
    bind_context_once();
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)hd, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = hd == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : GAOT_E_BADARG;
}

template <int HD, bool SPLIT = false>
int launch_fwd(const CUtensorMap& tm, const AttnArgs& g, int64_t B, cudaStream_t st) {
    constexpr int SMEM = (SPLIT ? 2 : 1) * (5 * 128 * HD * 2 + 2 * AT_BQ * AT_BKV * 2) + 2048 + 1024 + 256;
    static_assert(SMEM <= 227 * 1024, "forward attention tile set exceeds shared memory");
    static bool done = false;
    if (!done) {
        GAOT_CHECK(cudaFuncSetAttribute(attn_fwd_tc_kernel<HD, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        done = true;
    }
    dim3 grid((unsigned)((g.T + AT_BQ - 1) / AT_BQ), (unsigned)g.H, (unsigned)B);
    attn_fwd_tc_kernel<HD, SPLIT><<<grid, AT_THREADS, SMEM, st>>>(tm, g);
    GAOT_RETURN_LAST();
}

// split-precision planes [2, rows, ld] fp16 -> 3-D map {ld, rows, 2}, box {hd, box_rows, 1}, swizzle span = hd*2 bytes
int make_head_map_planes(CUtensorMap* map, const void* base, int64_t rows, int64_t ld, int hd, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return GAOT_E_BADARG;
    bind_context_once();
    cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)rows, 2};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * (cuuint64_t)rows * 2};
    cuuint32_t box[3] = {(cuuint32_t)hd, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle sw = hd == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (hd == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT) {
        cudaFree(nullptr);                                     // bind the primary context on this thread, then encode again
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    return r == CUDA_SUCCESS ? 0 : GAOT_E_BADARG;
}

}  // namespace

GAOT_API int gaot_attention_fwd_bf16(const void* qkv, void* o, float* lse, int64_t B, int64_t T, int64_t H, int64_t Hkv,
                                     int64_t hd, int64_t ld, float scale, void* stream) {
    if (!qkv || !o || !lse || B < 1 || T < 1 || H < 1 || Hkv < 1 || H % Hkv) return GAOT_E_BADARG;
    if (hd != 16 && hd != 32 && hd != 64) return GAOT_E_BADARG;
    if (ld < (H + 2 * Hkv) * hd || ld % 8 || (reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(o) & 15))
        return GAOT_E_ALIGN;
    if (B > 65535 || H > 65535 || B * T > 0x7fffffffLL) return GAOT_E_BADARG;
    CUtensorMap tm;
    int rc = make_head_map(&tm, qkv, B * T, ld, (int)hd, 128);
    if (rc) return rc;
    AttnArgs g{o, lse, (int)T, (int)H, (int)Hkv, (int)(H * hd), (int)((H + Hkv) * hd), scale * 1.4426950408889634f, nullptr};
    cudaStream_t st = (cudaStream_t)stream;
    if (hd == 16) return launch_fwd<16>(tm, g, B, st);
    if (hd == 32) return launch_fwd<32>(tm, g, B, st);
    return launch_fwd<64>(tm, g, B, st);
}

GAOT_API int gaot_attention_fwd_split(const void* qkv_planes, const float* qscale, float* o, float* lse, int64_t B, int64_t T,
                                      int64_t H, int64_t Hkv, int64_t hd, int64_t ld, float scale, void* stream) {
    if (!qkv_planes || !qscale || !o || !lse || B < 1 || T < 1 || H < 1 || Hkv < 1 || H % Hkv) return GAOT_E_BADARG;
    if (hd != 16 && hd != 32) return GAOT_E_BADARG;        // hd = 64 does not fit the doubled tile set; callers use the FFMA kernels
    if (ld < (H + 2 * Hkv) * hd || ld % 8 || (reinterpret_cast<uintptr_t>(qkv_planes) & 15) || (reinterpret_cast<uintptr_t>(o) & 15))
        return GAOT_E_ALIGN;
    if (B > 65535 || H > 65535 || B * T > 0x7fffffffLL) return GAOT_E_BADARG;
    CUtensorMap tm;
    int rc = make_head_map_planes(&tm, qkv_planes, B * T, ld, (int)hd, 128);
    if (rc) return rc;
    AttnArgs g{o, lse, (int)T, (int)H, (int)Hkv, (int)(H * hd), (int)((H + Hkv) * hd), scale * 1.4426950408889634f, qscale};
    cudaStream_t st = (cudaStream_t)stream;
    if (hd == 16) return launch_fwd<16, true>(tm, g, B, st);
    return launch_fwd<32, true>(tm, g, B, st);
}

namespace {
template <int HD>
int launch_bwd_split(const CUtensorMap& tm, const CUtensorMap& tmdo, const AttnBwdArgs& g, const float* o, const float* dout,
                     float* delta, int64_t B, cudaStream_t st) {
    using Cfg = AttnBwdCfg<HD, true, 128>;
    constexpr int SMEM = Cfg::SMEM;
    static bool done = false;
    if (!done) {
        GAOT_CHECK(cudaFuncSetAttribute(attn_bwd_tc_kernel<HD, true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        done = true;
    }
    const int64_t rows = B * g.T;
    attn_delta_f32_kernel<HD><<<gaot_grid(rows * g.H, 256), 256, 0, st>>>(o, dout, delta, B, g.T, g.H);
    // dQ is reduced into the query columns of dqkv by several key tiles: zero them first
    GAOT_CHECK(cudaMemset2DAsync(g.dqkv, sizeof(float) * (size_t)g.ld, 0, sizeof(float) * (size_t)(g.H * HD), (size_t)rows, st));
    dim3 grid((unsigned)((g.T + 127) / 128), (unsigned)g.Hkv, (unsigned)B);
    attn_bwd_tc_kernel<HD, true, 128><<<grid, Cfg::THREADS, SMEM, st>>>(tm, tm, tmdo, tm, g);
    GAOT_NOTE_LAUNCHES(1);
    GAOT_RETURN_LAST();
}

template <int HD, int QW>
int launch_bwd(const CUtensorMap& tm, const CUtensorMap& tmq, const CUtensorMap& tmdo, const AttnBwdArgs& g, const void* o,
               const void* dout, float* delta, int64_t B, cudaStream_t st) {
    using Cfg = AttnBwdCfg<HD, false, QW>;
    constexpr int SMEM = Cfg::SMEM;
    static bool done = false;
    if (!done) {
        GAOT_CHECK(cudaFuncSetAttribute(attn_bwd_tc_kernel<HD, false, QW>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        done = true;
    }
    const int64_t rows = B * g.T;
    attn_delta_kernel<HD><<<gaot_grid(rows * g.H, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(o),
                                                                     reinterpret_cast<const __nv_bfloat16*>(dout), delta, B, g.T, g.H);
    GAOT_CHECK(cudaMemsetAsync(g.dq_acc, 0, sizeof(float) * (size_t)(rows * g.H * HD), st));
    dim3 grid((unsigned)((g.T + 127) / 128), (unsigned)g.Hkv, (unsigned)B);
    CUtensorMap tmacc = tm;                                  // 128-query steps reduce dQ with red.add: no accumulator map
    if constexpr (Cfg::TRED) {
        int rc = make_acc_map(&tmacc, g.dq_acc, rows, (int64_t)g.H * HD, HD, QW);
        if (rc) return rc;
    }
    attn_bwd_tc_kernel<HD, false, QW><<<grid, Cfg::THREADS, SMEM, st>>>(tm, tmq, tmdo, tmacc, g);
    attn_dq_store_kernel<<<gaot_grid(rows * g.H * HD / 4, 256), 256, 0, st>>>(g.dq_acc, reinterpret_cast<__nv_bfloat16*>(g.dqkv), rows,
                                                                            g.H * HD, g.ld);
    GAOT_NOTE_LAUNCHES(2);
    GAOT_RETURN_LAST();
}
}  // namespace

GAOT_API int gaot_attention_bwd_bf16(const void* qkv, const void* o, const float* lse, const void* dout, float* delta,
                                     float* dq_acc, void* dqkv, int64_t B, int64_t T, int64_t H, int64_t Hkv, int64_t hd,
                                     int64_t ld, float scale, void* stream) {
    if (!qkv || !o || !lse || !dout || !delta || !dq_acc || !dqkv || B < 1 || T < 1 || H < 1 || Hkv < 1 || H % Hkv) return GAOT_E_BADARG;
    if (hd != 16 && hd != 32 && hd != 64) return GAOT_E_BADARG;
    if (ld < (H + 2 * Hkv) * hd || ld % 8 || (reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(dout) & 15) ||
        (reinterpret_cast<uintptr_t>(dqkv) & 15) || (reinterpret_cast<uintptr_t>(dq_acc) & 15) || (reinterpret_cast<uintptr_t>(o) & 15))
        return GAOT_E_ALIGN;
    if (B > 65535 || Hkv > 65535 || B * T > 0x7fffffffLL) return GAOT_E_BADARG;
    // head dimensions up to 32: 64-query steps, two CTAs per SM (GAOT_B200_ATTN_BWD_QW=128 selects the one-CTA kernel)
    static const bool wide = [] { const char* e = getenv("GAOT_B200_ATTN_BWD_QW"); return e && atoi(e) == 128; }();
    const int qw = (hd == 64 || wide) ? 128 : 64;
    CUtensorMap tm, tmq, tmdo;
    int rc = make_head_map(&tm, qkv, B * T, ld, (int)hd, 128);
    if (rc) return rc;
    rc = make_head_map(&tmq, qkv, B * T, ld, (int)hd, qw);
    if (rc) return rc;
    rc = make_head_map(&tmdo, dout, B * T, H * hd, (int)hd, qw);
    if (rc) return rc;
    AttnBwdArgs g{lse, delta, dq_acc, dqkv, (int)T, (int)H, (int)Hkv, (int)(H * hd), (int)((H + Hkv) * hd), (int)ld, scale,
                  scale * 1.4426950408889634f, nullptr, nullptr};
    cudaStream_t st = (cudaStream_t)stream;
    if (qw == 64) {
        if (hd == 16) return launch_bwd<16, 64>(tm, tmq, tmdo, g, o, dout, delta, B, st);
        return launch_bwd<32, 64>(tm, tmq, tmdo, g, o, dout, delta, B, st);
    }
    if (hd == 16) return launch_bwd<16, 128>(tm, tmq, tmdo, g, o, dout, delta, B, st);
    if (hd == 32) return launch_bwd<32, 128>(tm, tmq, tmdo, g, o, dout, delta, B, st);
    return launch_bwd<64, 128>(tm, tmq, tmdo, g, o, dout, delta, B, st);
}

GAOT_API int gaot_attention_bwd_split(const void* qkv_planes, const float* qscale, const float* o, const float* lse,
                                      const float* dout, const void* dout_planes, const float* gscale, float* delta, float* dqkv,
                                      int64_t B, int64_t T, int64_t H, int64_t Hkv, int64_t hd, int64_t ld, float scale,
                                      void* stream) {
    if (!qkv_planes || !qscale || !o || !lse || !dout || !dout_planes || !gscale || !delta || !dqkv || B < 1 || T < 1 || H < 1 ||
        Hkv < 1 || H % Hkv)
        return GAOT_E_BADARG;
    if (hd != 16 && hd != 32) return GAOT_E_BADARG;
    if (ld < (H + 2 * Hkv) * hd || ld % 8 || (reinterpret_cast<uintptr_t>(qkv_planes) & 15) || (reinterpret_cast<uintptr_t>(dout) & 15) ||
        (reinterpret_cast<uintptr_t>(dout_planes) & 15) || (reinterpret_cast<uintptr_t>(dqkv) & 15) || (reinterpret_cast<uintptr_t>(o) & 15))
        return GAOT_E_ALIGN;
    if (B > 65535 || Hkv > 65535 || B * T > 0x7fffffffLL) return GAOT_E_BADARG;
    CUtensorMap tm, tmdo;
    int rc = make_head_map_planes(&tm, qkv_planes, B * T, ld, (int)hd, 128);
    if (rc) return rc;
    rc = make_head_map_planes(&tmdo, dout_planes, B * T, H * hd, (int)hd, 128);
    if (rc) return rc;
    AttnBwdArgs g{lse, delta, nullptr, dqkv, (int)T, (int)H, (int)Hkv, (int)(H * hd), (int)((H + Hkv) * hd), (int)ld, scale,
                  scale * 1.4426950408889634f, qscale, gscale};
    cudaStream_t st = (cudaStream_t)stream;
    if (hd == 16) return launch_bwd_split<16>(tm, tmdo, g, o, dout, delta, B, st);
    return launch_bwd_split<32>(tm, tmdo, g, o, dout, delta, B, st);
}

#ifdef EXP_TRACE
GAOT_API int gaot_debug_trace(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(g_trace)); }
#endif

#ifdef EXP_TRACE_FWD
GAOT_API int gaot_debug_trace_fwd(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_trace_f, sizeof(g_trace_f)); }
#endif
