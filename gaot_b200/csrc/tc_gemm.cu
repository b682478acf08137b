// bf16 tensor-core GEMM for the processor's dense layers (the performance path):
//     C[M,N] (fp32 or bf16) = A[M,K] (bf16, row-major) * B[N,K]^T (bf16, row-major) + bias[n] + add[(m % add_rows), n]
// fp32 accumulation in TMEM.  Reference call sites: nn.Linear of src/model/layers/attn.py:92-94,117,151,
// src/model/gaot.py:208 and their autograd backward.
//
// Blackwell design (sm_100a only):
//   * warp 0  : TMA producer  - cp.async.bulk.tensor 2-D boxes (64 x 128 rows, SWIZZLE_128B) into a 3-stage ring,
//               completion signalled on mbarriers (complete_tx)
//   * warp 1  : TMEM allocator + MMA issuer - one lane issues tcgen05.mma.cta_group::1.kind::f16
//               (M=128, N=BN, K=16) straight from the swizzled smem tiles; tcgen05.commit frees ring slots
//               and finally publishes the accumulator
//   * warps 2-5: epilogue - tcgen05.ld (32 lanes x 32 columns per warp), bias / broadcast-add, 16-byte stores
//               (or 128-bit red.add for split-K)
//   3 stages x 32 KB leave room for two CTAs per SM, so one CTA's epilogue overlaps the other's main loop.
#include <cuda.h>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "../../include/gaot_b200.h"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;          // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int TC_STAGES = 5;          // 5 x 32 KB in flight (with 64 KB of epilogue staging: 225 KB): one stage more than a K = 256 tile, so the ring runs ahead across tiles
constexpr int TC_THREADS = 320;        // TMA warp, MMA warp, 8 epilogue warps

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
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// 3-D form: coordinate c2 selects the hi / lo plane of a split-precision operand ([2, rows, cols] fp16)
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
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
// K-major operand tile, rows of 128 bytes, SWIZZLE_128B: 8-row atoms 1024 bytes apart (SBO), LBO unused (=1),
// descriptor version 1 (sm_100), layout type 2 (cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
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
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
// TMA reduction: global[box] += smem[box] (fp32 add performed by the memory system; split-K / accumulate epilogue)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_s(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f4(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster (one TPC) work on one 256-row tile.  Each CTA keeps its own
// 128 rows of A and HALF of the B tile in shared memory; the MMA (issued by the leader, CTA rank 0) reads A from each CTA's
// own memory and the two B halves from both, and writes each CTA's 128 accumulator rows into that CTA's TMEM.  Operand
// traffic from L2 per output element is half of the single-CTA kernel's.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are counted on the LEADER CTA's mbarrier (same shared-memory offset, CTA-rank bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// commit: arrive on the mbarrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the mbarrier at this offset in the leader CTA (rank 0) of the cluster
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)) : "memory");
}

#ifdef EXP_TRACE
__device__ long long g_trace[4 * 256];
#define TRACE(slot) do { if (trace_on && (slot) < 256) g_trace[trace_base + (slot)] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif
struct TcArgs {
    void* C; int64_t M, N, K;
    const float* bias; const float* add; int64_t add_rows, ldadd;
    int out_bf16, splits, accumulate;
    int64_t gm, gn;                   // tile grid
    int act; float* pre;              // optional activation (GAOT_ACT_*) and fp32 pre-activation output
    int fast;                         // epilogue through swizzled smem boxes and TMA: 1 = stores, 2 = fp32 reduce-add (split-K / accumulate)
    // split-precision mode (SPLIT): operands are [2, rows, cols] fp16 hi / lo planes of x * scale; sa / sb point at the
    // {scale, 1 / scale} pair written by gaot_split_f16; kb_real = k-blocks of ONE term (the virtual contraction is 3 terms long)
    const float* sa; const float* sb; int64_t kb_real;
    const void* aux;                  // TC_ACT_SWIGLU_FWD: U output [M, N/2]; TC_ACT_SWIGLU_BWD: saved pre-activations [M, 2N]
    int no_alt;                       // experiment switch (GAOT_GEMM_NO_ALT): bf16 epilogue passes do not alternate staging boxes
};

// Fused SwiGLU epilogues (internal `act` codes; bf16 output, fast epilogue, two 32-column chunks per pass):
//   TC_ACT_SWIGLU_FWD: the output columns come in blocks of 64 = [a(32 features) | b(32 features)] (weight rows interleaved
//     by the caller); besides the [M, N] pre-activations the epilogue stores u = silu(a) * b into U [M, N/2] (tmAdd = U map)
//   TC_ACT_SWIGLU_BWD: the accumulator holds dU [M, N]; the epilogue loads the matching [32 x 64] box of the saved
//     pre-activations (tmAdd = their map), forms (dA | dB) = (dU b silu'(a) | dU silu(a)) in place and stores it into
//     dAB [M, 2N] (tmC): dU itself never reaches memory
constexpr int TC_ACT_SWIGLU_FWD = 3, TC_ACT_SWIGLU_BWD = 4;
__device__ __forceinline__ float sigmoid_fast(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    return __fdividef(1.f, 1.f + e);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t p) { return make_float2(__uint_as_float(p << 16), __uint_as_float(p & 0xffff0000u)); }
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

// MN-major operand tile: [64 contraction rows][64 elements = 128 bytes] blocks of 8 KB, one per 64 output
// rows/columns; SWIZZLE_128B; LBO = 8 KB between the 64-wide blocks, SBO = 1 KB between 8-row groups
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

// Persistent GEMM: C[M,N] = opA * opB, one CTA per SM walking tiles (m fastest) with a static stride.
//   AMN = false: A is [M, K] row-major (K-major operand);  AMN = true: A is [K, M] row-major (MN-major operand)
//   BMN = false: B is [N, K] row-major (K-major operand);  BMN = true: B is [K, N] row-major (MN-major operand)
// so forward / dX / dW of a Linear layer all read activations and weights in their natural layout.
// Two TMEM accumulators (2*BN columns): the epilogue of tile i overlaps the main loop of tile i+1, and the TMA
// ring keeps prefetching across tile boundaries.
//
// SPLIT = true is the split-precision (parity) variant: x = (hi + lo) / scale with hi, lo fp16, and
//     A B^T = (A_lo B_hi + A_hi B_lo + A_hi B_hi) / (scale_a scale_b)          (the lo x lo term is below 2^-22 relative)
// runs as ONE accumulation over a contraction that is three terms long: the producer maps the virtual k-block index to
// (term, real k-block) and picks the hi / lo plane of each operand with the third TMA coordinate; MMA issue, pipeline and
// epilogue are unchanged except for kind::f16 reading fp16 instead of bf16 and the 1 / (scale_a scale_b) factor.
template <int BN, bool AMN, bool BMN, int CG, bool SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAdd, const TcArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int BNL = BN / CG;                                   // B rows (output columns) held by this CTA
    constexpr int A_BYTES = TC_BM * TC_BK * 2, B_BYTES = BNL * TC_BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;       // CG = 2: rank 0 (the leader) issues the MMAs for the pair
    const int64_t worker = blockIdx.x / CG, n_workers = gridDim.x / CG;
#ifdef EXP_TRACE
    // CTA 0: MMA warp -> [0,256), epilogue warp 2 -> [256,512); CTA 1 (the peer): epilogue warp 2 -> [512,768), TMA warp -> [768,1024)
    const bool trace_on = (threadIdx.x & 31) == 0 && ((blockIdx.x == 0 && ((threadIdx.x >> 5) == 1 || (threadIdx.x >> 5) == 2)) ||
                                                       (blockIdx.x == 1 && ((threadIdx.x >> 5) == 2 || (threadIdx.x >> 5) == 0)));
    const int trace_base = blockIdx.x == 0 ? ((threadIdx.x >> 5) == 1 ? 0 : 256) : ((threadIdx.x >> 5) == 2 ? 512 : 768);
    TRACE(0);
#endif
    constexpr int PITCH = 36;                                      // fallback epilogue staging: floats per row (16-byte aligned)
    constexpr int STG_WARP = 8192;                                 // per epilogue warp: two 4 KB boxes (32 rows x 32 columns)
    float* stg_all = reinterpret_cast<float*>(smem + TC_STAGES * STAGE_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + TC_STAGES * STAGE_BYTES + 8 * STG_WARP);
    uint64_t* empty_bar = full_bar + TC_STAGES;
    uint64_t* acc_full = empty_bar + TC_STAGES;                    // [2]
    uint64_t* acc_empty = acc_full + 2;                            // [2]
    uint64_t* add_bar = acc_empty + 2;                             // [8 warps][2 boxes]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(add_bar + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t kblocks = (g.K + TC_BK - 1) / TC_BK;
    const int64_t per = (kblocks + g.splits - 1) / g.splits;
    const int64_t n_tiles = g.gm * g.gn * g.splits;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8 * CG); }
        for (int s = 0; s < 16; ++s) mbar_init(&add_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        if (g.fast) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
            if (g.add) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAdd) : "memory");
        }
    }
    if (CG == 2) cluster_sync_all();                               // the peer's mbarriers exist before anything remote touches them
    if (warp == 1) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BN)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BN)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // (32-bit index arithmetic: the launcher bounds gm * gn * splits below 2^31; a 64-bit divide costs ~150 cycles and every
    // role walks this once per tile - on the MMA warp that was 600 cycles of a 5 000-cycle tile)
    const uint32_t gm32 = (uint32_t)g.gm, gn32 = (uint32_t)g.gn;
    auto tile_coords = [&](int64_t t, int64_t& m0, int64_t& n0, int64_t& kb0, int& nkb) {
        const uint32_t t32 = (uint32_t)t;
        const uint32_t rest32 = t32 / gm32, mi32 = t32 - rest32 * gm32;
        const uint32_t sp32 = rest32 / gn32, ni32 = rest32 - sp32 * gn32;
        const int64_t mi = mi32, ni = ni32, sp = sp32;
        m0 = mi * (TC_BM * CG) + rank * TC_BM; n0 = ni * BN;
        kb0 = sp * per;
        nkb = (int)max((int64_t)0, min(kblocks, kb0 + per) - kb0);
    };

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            uint32_t it = 0;
            auto load = [&](void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int plane) {
                if (SPLIT) { if (CG == 2) tma_load_3d_pair(dst, map, bar, c0, c1, plane); else tma_load_3d(dst, map, bar, c0, c1, plane); }
                else { if (CG == 2) tma_load_2d_pair(dst, map, bar, c0, c1); else tma_load_2d(dst, map, bar, c0, c1); }
            };
            for (int64_t t = worker; t < n_tiles; t += n_workers) {
                int64_t m0, n0, kb0; int nkb;
                tile_coords(t, m0, n0, kb0, nkb);
                const int nb0 = (int)n0 + (int)rank * BNL;           // this CTA's half of the B tile
                for (int i = 0; i < nkb; ++i, ++it) {
                    const int s = (int)(it % TC_STAGES);
                    mbar_wait(&empty_bar[s], ((it / TC_STAGES) & 1u) ^ 1u);
                    TRACE(1 + it);
                    // one arrival (the leader's) per phase; both CTAs' bytes are counted on the leader's barrier
                    if (rank == 0) mbar_expect_tx(&full_bar[s], STAGE_BYTES * CG);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    // SPLIT: terms in increasing magnitude, (A lo, B hi), (A hi, B lo), (A hi, B hi)
                    const int64_t vk = kb0 + i;
                    const int term = SPLIT ? (int)(vk / g.kb_real) : 0;
                    const int kc = (int)((SPLIT ? vk - term * g.kb_real : vk) * TC_BK);
                    const int pa = term == 0 ? 1 : 0, pb = term == 1 ? 1 : 0;
                    if (!AMN) load(st, &tmA, &full_bar[s], kc, (int)m0, pa);
                    else {
#pragma unroll
                        for (int j = 0; j < TC_BM / 64; ++j) load(st + j * 8192, &tmA, &full_bar[s], (int)m0 + 64 * j, kc, pa);
                    }
                    if (!BMN) load(st + A_BYTES, &tmB, &full_bar[s], kc, nb0, pb);
                    else {
#pragma unroll
                        for (int j = 0; j < BNL / 64; ++j) load(st + A_BYTES + j * 8192, &tmB, &full_bar[s], nb0 + 64 * j, kc, pb);
                    }
                }
            }
        }
    } else if (warp == 1) {
      if (rank == 0) {
        // ---------------- MMA issuer (CG = 2: the leader CTA issues for the pair) ----------------
        // instruction descriptor: D fp32, A/B bf16, majors per template, N = BN, M = 128 (cute/arch/mma_sm100_desc.hpp)
        // (A/B format field 1 = bf16, 0 = fp16: the split-precision planes are fp16)
        constexpr uint32_t IDESC = (1u << 4) | (SPLIT ? 0u : ((1u << 7) | (1u << 10))) | (AMN ? (1u << 15) : 0u) | (BMN ? (1u << 16) : 0u) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((TC_BM * CG) >> 4) << 24);
        uint32_t it = 0, lt = 0;
        for (int64_t t = worker; t < n_tiles; t += n_workers, ++lt) {
            int64_t m0, n0, kb0; int nkb;
            tile_coords(t, m0, n0, kb0, nkb);
            const uint32_t ab = lt & 1u;
            TRACE(1 + lt * 8);
            mbar_wait(&acc_empty[ab], ((lt >> 1) & 1u) ^ 1u);       // epilogue has drained this accumulator
            TRACE(2 + lt * 8);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_acc = tmem_base + ab * BN;
            for (int i = 0; i < nkb; ++i, ++it) {
                const int s = (int)(it % TC_STAGES);
                mbar_wait(&full_bar[s], (it / TC_STAGES) & 1u);
                if (i < 4) TRACE(3 + lt * 8 + i);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
                    const uint64_t da0 = AMN ? make_smem_desc_mn(sa) : make_smem_desc_sw128(sa);
                    const uint64_t db0 = BMN ? make_smem_desc_mn(sb) : make_smem_desc_sw128(sb);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        // start-address field counts 16-byte units.  K-major: +32 bytes per K=16 step inside the
                        // 128-byte swizzle row; MN-major: 16 contraction rows = 2048 bytes
                        const uint64_t da = da0 + (uint64_t)(AMN ? (k * 2048) >> 4 : 2 * k);
                        const uint64_t db = db0 + (uint64_t)(BMN ? (k * 2048) >> 4 : 2 * k);
                        if (CG == 2) umma_f16_pair(tmem_acc, da, db, IDESC, (i > 0 || k > 0) ? 1u : 0u);
                        else umma_f16(tmem_acc, da, db, IDESC, (i > 0 || k > 0) ? 1u : 0u);
                    }
                    if (CG == 2) {
                        umma_commit_pair(&empty_bar[s]);           // frees the slot in both CTAs
                        if (i == nkb - 1) umma_commit_pair(&acc_full[ab]);
                    } else {
                        umma_commit(&empty_bar[s]);                // frees the slot once these MMAs have read it
                        if (i == nkb - 1) umma_commit(&acc_full[ab]);  // accumulator complete
                    }
                }
                __syncwarp();
            }
        }
      }
    } else {
        // ---------------- epilogue (8 warps; warp w may touch TMEM lanes 32*(w%4) .. +31) ----------------
        // TMEM hands each thread one accumulator ROW.  A per-warp 32x32 fp32 staging tile turns that into
        // row-contiguous global accesses: lane -> (row it*4 + lane/8, 4 columns (lane%8)*4), i.e. each store
        // instruction writes 4 rows x 128 bytes (fp32) / 64 bytes (bf16).  Two warps share a lane quadrant and
        // split the BN columns.  Row offsets are hoisted out of the column loop; the store loops are specialised.
        const int q = warp & 3, half = (warp - 2) >> 2;
        float alpha = 1.f;                                         // SPLIT: 1 / (scale_a * scale_b), exact powers of two
        if (SPLIT) alpha = __ldg(g.sa + 1) * __ldg(g.sb + 1);
        if (g.fast) {
            // Fast epilogue.  Per warp: 32 accumulator rows x BN/2 columns, in chunks of 32 columns.  A lane owns one
            // row: TMEM -> registers (+ bias, + the `add` tile that TMA has loaded into the same box, activation) ->
            // 16-byte stores into a swizzled 32 x 32 box (fp32: 128-byte rows, SWIZZLE_128B; bf16: 64-byte rows,
            // SWIZZLE_64B; both conflict-free) -> one TMA store per box.  Rows >= M / columns >= N are clipped by TMA.
            constexpr int NCH = BN / 64;                           // 32-column chunks per warp and tile (BN = 256: 4, 128: 2, 64: 1)
            constexpr int PCH = NCH < 2 ? NCH : 2;                 // chunks per pass = staging boxes per warp
            constexpr int NPASS = NCH / PCH;
            const uint32_t box0 = smem_u32(smem + TC_STAGES * STAGE_BYTES) + (uint32_t)(warp - 2) * STG_WARP;
            uint64_t* abar = add_bar + (warp - 2) * 2;
            uint32_t lt = 0, aload[2] = {};                        // add boxes loaded so far (mbarrier phase bookkeeping)
            for (int64_t t = worker; t < n_tiles; t += n_workers, ++lt) {
                int64_t m0, n0, kb0; int nkb;
                tile_coords(t, m0, n0, kb0, nkb);
                const uint32_t ab = lt & 1u;
                const int row0 = (int)m0 + q * 32;
                const int64_t ncol0 = n0 + half * (BN / 2);
                const int arow = g.add ? (int)((int64_t)row0 % g.add_rows) : 0;
                const uint32_t tmem_acc = tmem_base + ab * BN + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (BN / 2));
#pragma unroll
                for (int ps = 0; ps < NPASS; ++ps) {
                    // both boxes must have been read out by the previous TMA stores before they are refilled.  The plain bf16
                    // pass fills ONE 4 KB box: passes alternate between the warp's two boxes, so only the store of the pass
                    // before the previous one has to be done (the previous store drains under this pass)
                    const bool alt = PCH == 2 && g.out_bf16 && g.act < TC_ACT_SWIGLU_FWD && !g.no_alt;
                    const uint32_t boxp = box0 + ((alt && (ps & 1)) ? 4096u : 0u);
                    if (lane == 0) {
                        if (alt) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    __syncwarp();
                    if (g.add) {
#pragma unroll
                        for (int c = 0; c < PCH; ++c) {
                            if (ncol0 + (ps * PCH + c) * 32 < g.N && row0 < g.M) {
                                if (lane == 0) {
                                    mbar_expect_tx(&abar[c], 4096);
                                    tma_load_2d_s(box0 + c * 4096, &tmAdd, &abar[c], (int)ncol0 + (ps * PCH + c) * 32, arow);
                                }
                                ++aload[c];
                            }
                        }
                    }
                    if (g.act == TC_ACT_SWIGLU_BWD) {              // the [32 x 64] (a | b) boxes of this pass's two chunks
#pragma unroll
                        for (int c = 0; c < PCH; ++c) {
                            if (ncol0 + (ps * PCH + c) * 32 < g.N && row0 < g.M) {
                                if (lane == 0) {
                                    mbar_expect_tx(&abar[c], 4096);
                                    tma_load_2d_s(box0 + c * 4096, &tmAdd, &abar[c], 2 * ((int)ncol0 + (ps * PCH + c) * 32), row0);
                                }
                                ++aload[c];
                            }
                        }
                    }
                    if (ps == 0) {
                        TRACE(1 + lt * 8);
                        mbar_wait(&acc_full[ab], (lt >> 1) & 1u);
                        TRACE(2 + lt * 8);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    uint32_t r[PCH][32];
#pragma unroll
                    for (int c = 0; c < PCH; ++c) tmem_ld32_issue(tmem_acc + (ps * PCH + c) * 32, r[c]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    TRACE(3 + lt * 8 + 2 * ps);
                    if (ps == NPASS - 1) {
                        // the accumulator is in registers: hand it back to the MMA warp before the stores
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 2) mbar_arrive_leader(&acc_empty[ab]);
                            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[ab])) : "memory");
                        }
                    }
                    if (row0 >= g.M) continue;                     // warp-uniform: nothing of this warp's rows exists
#pragma unroll
                    for (int c = 0; c < PCH; ++c) {
                        const int64_t n = ncol0 + (ps * PCH + c) * 32;
                        if (n >= g.N) continue;                    // warp-uniform
                        const uint32_t box = box0 + c * 4096;
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[c][j]);
                        if (PCH == 2 && g.act == TC_ACT_SWIGLU_BWD) {
                            // own row of the (a | b) box -> (dA | dB) in place, one TMA store of the [32 x 64] box
                            mbar_wait(&abar[c], (aload[c] - 1u) & 1u);
                            const uint32_t rowa = box + (uint32_t)lane * 128u;
#pragma unroll
                            for (int w = 0; w < 4; ++w) {
                                const uint4 pa = lds_u4(rowa + (uint32_t)((w ^ (lane & 7)) << 4));
                                const uint4 pb = lds_u4(rowa + (uint32_t)(((w + 4) ^ (lane & 7)) << 4));
                                const uint32_t pav[4] = {pa.x, pa.y, pa.z, pa.w}, pbv[4] = {pb.x, pb.y, pb.z, pb.w};
                                uint32_t oa[4], ob[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const float2 a2 = unpack_bf16x2(pav[k]), b2 = unpack_bf16x2(pbv[k]);
                                    const float d0 = v[8 * w + 2 * k], d1 = v[8 * w + 2 * k + 1];
                                    const float s0 = sigmoid_fast(a2.x), s1 = sigmoid_fast(a2.y);
                                    oa[k] = pack_bf16x2(d0 * b2.x * s0 * (1.f + a2.x * (1.f - s0)), d1 * b2.y * s1 * (1.f + a2.y * (1.f - s1)));
                                    ob[k] = pack_bf16x2(d0 * a2.x * s0, d1 * a2.y * s1);
                                }
                                sts_u4(rowa + (uint32_t)((w ^ (lane & 7)) << 4), oa[0], oa[1], oa[2], oa[3]);
                                sts_u4(rowa + (uint32_t)(((w + 4) ^ (lane & 7)) << 4), ob[0], ob[1], ob[2], ob[3]);
                            }
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_2d(&tmC, box, 2 * (int)n, row0);
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            }
                            continue;
                        }
                        if (PCH == 2 && g.act == TC_ACT_SWIGLU_FWD) {
                            if (c == 1) {
                                // u = silu(a) * b (a = chunk 0 of the pass, still in registers) -> the warp's second staging box
                                // as a [32 x 32] bf16 tile (64-byte rows)
                                float u[32];
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const float a = __uint_as_float(r[0][j]);
                                    u[j] = a * sigmoid_fast(a) * v[j];
                                }
#pragma unroll
                                for (int w = 0; w < 4; ++w)
                                    sts_u4(box0 + 4096u + (uint32_t)lane * 64u + (uint32_t)((w ^ ((lane >> 1) & 3)) << 4),
                                           pack_bf16x2(u[8 * w], u[8 * w + 1]), pack_bf16x2(u[8 * w + 2], u[8 * w + 3]),
                                           pack_bf16x2(u[8 * w + 4], u[8 * w + 5]), pack_bf16x2(u[8 * w + 6], u[8 * w + 7]));
                            }
                        }
                        if (SPLIT) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= alpha;
                        }
                        if (g.bias) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (n + j < g.N) bv = __ldg(reinterpret_cast<const float4*>(g.bias + n + j));   // N % 4 == 0 on this path
                                v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
                            }
                        }
                        if (g.add) {
                            mbar_wait(&abar[c], (aload[c] - 1u) & 1u);
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 av = lds_f4(box + (uint32_t)lane * 128u + (uint32_t)((((j >> 2) ^ (lane & 7))) << 4));
                                v[j] += av.x; v[j + 1] += av.y; v[j + 2] += av.z; v[j + 3] += av.w;
                            }
                        }
                        if (g.act == GAOT_ACT_GELU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                        } else if (g.act == GAOT_ACT_RELU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                        }
                        if (g.out_bf16) {
                            if (PCH == 2) {
                                // both chunks of the pass share ONE 32-row x 64-column box (128-byte rows, SWIZZLE_128B):
                                // TMA stores are paced per box row, so 128-byte rows halve the row count
#pragma unroll
                                for (int w = 0; w < 4; ++w)
                                    sts_u4(boxp + (uint32_t)lane * 128u + (uint32_t)(((c * 4 + w) ^ (lane & 7)) << 4),
                                           pack_bf16x2(v[8 * w], v[8 * w + 1]), pack_bf16x2(v[8 * w + 2], v[8 * w + 3]),
                                           pack_bf16x2(v[8 * w + 4], v[8 * w + 5]), pack_bf16x2(v[8 * w + 6], v[8 * w + 7]));
                            } else {
#pragma unroll
                                for (int w = 0; w < 4; ++w)
                                    sts_u4(box + (uint32_t)lane * 64u + (uint32_t)((w ^ ((lane >> 1) & 3)) << 4),
                                           pack_bf16x2(v[8 * w], v[8 * w + 1]), pack_bf16x2(v[8 * w + 2], v[8 * w + 3]),
                                           pack_bf16x2(v[8 * w + 4], v[8 * w + 5]), pack_bf16x2(v[8 * w + 6], v[8 * w + 7]));
                            }
                        } else {
#pragma unroll
                            for (int w = 0; w < 8; ++w)
                                sts_f4(box + (uint32_t)lane * 128u + (uint32_t)((w ^ (lane & 7)) << 4), v[4 * w], v[4 * w + 1], v[4 * w + 2], v[4 * w + 3]);
                        }
                        if (g.out_bf16 && PCH == 2) {
                            // one store for the pass, after its last existing chunk (clipped at N by TMA)
                            const bool last = c == PCH - 1 || n + 32 >= g.N;
                            if (last) {
                                if (ps == NPASS - 1) TRACE(7 + lt * 8);
                                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                                __syncwarp();
                                if (ps == NPASS - 1) TRACE(8 + lt * 8);
                                if (lane == 0) {
                                    tma_store_2d(&tmC, boxp, (int)(ncol0 + ps * PCH * 32), row0);
                                    if (g.act == TC_ACT_SWIGLU_FWD) tma_store_2d(&tmAdd, box0 + 4096u, (int)((ncol0 + ps * PCH * 32) >> 1), row0);
                                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                                }
                            }
                        } else {
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) {
                                if (g.fast == 2) tma_reduce_add_2d(&tmC, box, (int)n, row0);
                                else tma_store_2d(&tmC, box, (int)n, row0);
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            }
                        }
                    }
                    TRACE(4 + lt * 8 + 2 * ps);
                }
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncwarp();
        } else {
        float* stg = stg_all + (warp - 2) * (32 * PITCH);
        const bool vec_ok = (g.N % 4 == 0) && (!g.add || (g.ldadd % 4 == 0));
        const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
        const bool atomic = g.splits > 1 || g.accumulate;
        uint32_t lt = 0;
        for (int64_t t = worker; t < n_tiles; t += n_workers, ++lt) {
            int64_t m0, n0, kb0; int nkb;
            tile_coords(t, m0, n0, kb0, nkb);
            const uint32_t ab = lt & 1u;
            if (nkb > 0) {
                mbar_wait(&acc_full[ab], (lt >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const bool first = kb0 == 0;
            const float* bias = first ? g.bias : nullptr;
            const float* add = first ? g.add : nullptr;
            const uint32_t tmem_acc = tmem_base + ab * BN + ((uint32_t)(q * 32) << 16);
            // per-lane row bookkeeping for the 8 store iterations (rows do not depend on the column chunk)
            int64_t coff[8], aoff[8];
#pragma unroll
            for (int it8 = 0; it8 < 8; ++it8) {
                const int64_t m = m0 + q * 32 + it8 * 4 + sub_r;
                coff[it8] = (m < g.M) ? m * g.N : -1;
                aoff[it8] = (add && m < g.M) ? (m % g.add_rows) * g.ldadd : 0;
            }
#pragma unroll 1
            for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 32) {
                uint32_t r[32];
                if (nkb > 0) tmem_ld32(tmem_acc + (uint32_t)c0, r);
                else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                const int64_t n = n0 + c0;
                if (n >= g.N) continue;                            // warp-uniform
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(stg + lane * PITCH + j) =
                        make_float4(__uint_as_float(r[j]) * alpha, __uint_as_float(r[j + 1]) * alpha, __uint_as_float(r[j + 2]) * alpha,
                                    __uint_as_float(r[j + 3]) * alpha);
                __syncwarp();
                const int64_t nc = n + sub_c;
                if ((n + 32 <= g.N) && vec_ok) {
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bias) bv = __ldg(reinterpret_cast<const float4*>(bias + nc));
                    float4 v[8];
#pragma unroll
                    for (int it8 = 0; it8 < 8; ++it8) {
                        v[it8] = *reinterpret_cast<const float4*>(stg + (it8 * 4 + sub_r) * PITCH + sub_c);
                        v[it8].x += bv.x; v[it8].y += bv.y; v[it8].z += bv.z; v[it8].w += bv.w;
                    }
                    if (add) {
                        float4 av[8];
#pragma unroll
                        for (int it8 = 0; it8 < 8; ++it8)
                            av[it8] = (coff[it8] >= 0) ? __ldg(reinterpret_cast<const float4*>(add + aoff[it8] + nc)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int it8 = 0; it8 < 8; ++it8) { v[it8].x += av[it8].x; v[it8].y += av[it8].y; v[it8].z += av[it8].z; v[it8].w += av[it8].w; }
                    }
                    if (g.pre) {
#pragma unroll
                        for (int it8 = 0; it8 < 8; ++it8)
                            if (coff[it8] >= 0) *reinterpret_cast<float4*>(g.pre + nc + coff[it8]) = v[it8];
                    }
                    if (g.act == GAOT_ACT_GELU) {
#pragma unroll
                        for (int it8 = 0; it8 < 8; ++it8) { v[it8].x = gelu_erf(v[it8].x); v[it8].y = gelu_erf(v[it8].y); v[it8].z = gelu_erf(v[it8].z); v[it8].w = gelu_erf(v[it8].w); }
                    } else if (g.act == GAOT_ACT_RELU) {
#pragma unroll
                        for (int it8 = 0; it8 < 8; ++it8) { v[it8].x = fmaxf(v[it8].x, 0.f); v[it8].y = fmaxf(v[it8].y, 0.f); v[it8].z = fmaxf(v[it8].z, 0.f); v[it8].w = fmaxf(v[it8].w, 0.f); }
                    }
                    if (g.out_bf16) {
                        __nv_bfloat16* cb = reinterpret_cast<__nv_bfloat16*>(g.C) + nc;
#pragma unroll
                        for (int it8 = 0; it8 < 8; ++it8) {
                            if (coff[it8] < 0) continue;
                            __nv_bfloat162 t0 = __floats2bfloat162_rn(v[it8].x, v[it8].y), t1 = __floats2bfloat162_rn(v[it8].z, v[it8].w);
                            uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                            *reinterpret_cast<uint2*>(cb + coff[it8]) = pk;
                        }
                    } else if (atomic) {
                        float* cf = reinterpret_cast<float*>(g.C) + nc;
#pragma unroll
                        for (int it8 = 0; it8 < 8; ++it8)
                            if (coff[it8] >= 0) atomicAdd(reinterpret_cast<float4*>(cf + coff[it8]), v[it8]);
                    } else {
                        float* cf = reinterpret_cast<float*>(g.C) + nc;
#pragma unroll
                        for (int it8 = 0; it8 < 8; ++it8)
                            if (coff[it8] >= 0) *reinterpret_cast<float4*>(cf + coff[it8]) = v[it8];
                    }
                } else {                                           // ragged N / unaligned: scalar, guarded
#pragma unroll 1
                    for (int it8 = 0; it8 < 8; ++it8) {
                        if (coff[it8] < 0) continue;
                        const float4 v4 = *reinterpret_cast<const float4*>(stg + (it8 * 4 + sub_r) * PITCH + sub_c);
                        const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
                        for (int j = 0; j < 4; ++j) {
                            if (nc + j >= g.N) break;
                            float x = vv[j];
                            if (bias) x += __ldg(bias + nc + j);
                            if (add) x += __ldg(add + aoff[it8] + nc + j);
                            if (g.pre) g.pre[coff[it8] + nc + j] = x;
                            x = (g.act == GAOT_ACT_GELU) ? gelu_erf(x) : ((g.act == GAOT_ACT_RELU) ? fmaxf(x, 0.f) : x);
                            if (g.out_bf16) reinterpret_cast<__nv_bfloat16*>(g.C)[coff[it8] + nc + j] = __float2bfloat16_rn(x);
                            else if (atomic) atomicAdd(reinterpret_cast<float*>(g.C) + coff[it8] + nc + j, x);
                            else reinterpret_cast<float*>(g.C)[coff[it8] + nc + j] = x;
                        }
                    }
                }
            }
            // this warp is done with the accumulator: let the MMA warp reuse it
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_leader(&acc_empty[ab]);
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[ab])) : "memory");
            }
        }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cluster_sync_all(); else __syncthreads();       // pair: nobody leaves while the peer may still touch its memory
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)));
    }
}

// ---- fp32 -> bf16 casts and bf16 transposes (feed the TMA descriptors) ---------------------------------
__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&a); pk.y = *reinterpret_cast<uint32_t*>(&b);
        reinterpret_cast<uint2*>(out)[i] = pk;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(in[i]);
}

// out_t[c, r] = bf16(in[r, c]); TIn = float or bf16.  64x64 tiles through padded smem, both sides coalesced.
template <typename TIn>
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const TIn* __restrict__ in, __nv_bfloat16* __restrict__ out_t,
                                                             int64_t rows, int64_t cols) {
    __shared__ __nv_bfloat16 tile[64][66];
    const int64_t r0 = (int64_t)blockIdx.y * 64, c0 = (int64_t)blockIdx.x * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;          // 64 x 4
    for (int i = ty; i < 64; i += 4) {
        const int64_t r = r0 + i, c = c0 + tx;
        float v = 0.f;
        if (r < rows && c < cols) v = (float)in[r * cols + c];
        tile[i][tx] = __float2bfloat16_rn(v);
    }
    __syncthreads();
    for (int i = ty; i < 64; i += 4) {
        const int64_t c = c0 + i, r = r0 + tx;
        if (c < cols && r < rows) out_t[c * rows + r] = tile[tx][i];
    }
}

// ---- split-precision operands: x * scale = hi + lo (fp16 each, ~22 significant bits), scale a power of two that puts the
// tensor's largest magnitude just below 2^14 (fp16 overflows at 65504; the lo plane keeps absolute errors at 2^-25 of that) ----
constexpr int SPLIT_PARTS = 1024;                                  // upper bound on absmax partials

__global__ void __launch_bounds__(256) absmax_partial_kernel(const float* __restrict__ in, int64_t n, float* __restrict__ part) {
    const int64_t n4 = n >> 2;
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(in[i]));
    __shared__ float sm[8];
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 8) {
        m = sm[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
        if (threadIdx.x == 0) part[blockIdx.x] = m;
    }
}

__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ lo,
                                                        int64_t n, const float* __restrict__ part, int nparts,
                                                        float* __restrict__ scale_out) {
    __shared__ float sm[8];
    __shared__ float s_scale;
    float m = 0.f;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) m = fmaxf(m, part[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = sm[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) a = fmaxf(a, sm[w]);
        float sc = 1.f, inv = 1.f;
        if (a > 0.f && a <= 3.0e38f) {                             // zero / inf / nan tensors are passed through unscaled
            int e;
            frexpf(a, &e);                                         // a = f * 2^e, f in [0.5, 1)  =>  a * 2^(14 - e) < 2^14
            int se = 14 - e;
            if (se > 126) se = 126;
            sc = ldexpf(1.f, se);
            inv = ldexpf(1.f, -se);
        }
        s_scale = sc;
        if (blockIdx.x == 0) { scale_out[0] = sc; scale_out[1] = inv; }
    }
    __syncthreads();
    const float sc = s_scale;
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
        const float x[4] = {v.x * sc, v.y * sc, v.z * sc, v.w * sc};
        __half h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { h[k] = __float2half_rn(x[k]); l[k] = __float2half_rn(x[k] - __half2float(h[k])); }
        reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
        reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = in[i] * sc;
        const __half h = __float2half_rn(x);
        hi[i] = h;
        lo[i] = __float2half_rn(x - __half2float(h));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// the driver-API encode needs a current context: the autograd engine's threads have none until a runtime call binds it
inline void bind_context_once() {
    static thread_local bool bound = false;
    if (!bound) { cudaFree(nullptr); bound = true; }
}

EncodeTiledFn get_encode() {
    bind_context_once();
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

// [rows, K] bf16 row-major -> 2-D map with box {64 (K), box_rows}, SWIZZLE_128B, zero fill out of bounds
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return GAOT_E_BADARG;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT) {
        // driver API called on a thread (e.g. the autograd engine's) that has no current context yet: a runtime
        // call binds the primary context, then encode again
        cudaFree(nullptr);
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    return r == CUDA_SUCCESS ? 0 : GAOT_E_BADARG;
}

// [R, cols] bf16 row-major -> 2-D map with box {64 columns, 64 rows}, SWIZZLE_128B (MN-major operand blocks)
int make_map_mn(CUtensorMap* map, const void* base, int64_t R, int64_t cols) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return GAOT_E_BADARG;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT) {
        // driver API called on a thread (e.g. the autograd engine's) that has no current context yet: a runtime
        // call binds the primary context, then encode again
        cudaFree(nullptr);
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    return r == CUDA_SUCCESS ? 0 : GAOT_E_BADARG;
}

// Split-precision operand planes [2, rows, cols] fp16 -> 3-D map {cols, rows, 2} with box {64, box_rows, 1}, SWIZZLE_128B.
// K-major use: cols = contraction, box_rows = tile rows; MN-major use: rows = contraction, box {64, 64, 1}.
int make_map_planes(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return GAOT_E_BADARG;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
    cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * (cuuint64_t)rows * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT) {
        cudaFree(nullptr);                                     // bind the primary context on this thread, then encode again
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    return r == CUDA_SUCCESS ? 0 : GAOT_E_BADARG;
}

// output / add boxes of the fast epilogue: [rows, ld] matrix of fp32 or bf16, box {32 columns, 32 rows}, swizzle span =
// the 32-column row (128 bytes fp32 / 64 bytes bf16)
int make_map_box32(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, bool bf16, int box_cols = 32) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return GAOT_E_BADARG;
    const int es = bf16 ? 2 : 4;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, 32};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const CUtensorMapSwizzle sw = (bf16 && box_cols == 32) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;   // span = row bytes
    CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT) {
        cudaFree(nullptr);                                     // bind the primary context on this thread, then encode again
        r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    return r == CUDA_SUCCESS ? 0 : GAOT_E_BADARG;
}

int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <int BN, bool AMN, bool BMN, int CG, bool SPLIT = false>
int launch_tc(const void* A, const void* B, TcArgs& g, cudaStream_t st) {
    constexpr int BNL = BN / CG;
    CUtensorMap ta, tb;
    int rc;
    if (SPLIT) {
        rc = AMN ? make_map_planes(&ta, A, g.K, g.M, 64) : make_map_planes(&ta, A, g.M, g.K, TC_BM);
        if (rc) return rc;
        rc = BMN ? make_map_planes(&tb, B, g.K, g.N, 64) : make_map_planes(&tb, B, g.N, g.K, BNL);
    } else {
        rc = AMN ? make_map_mn(&ta, A, g.K, g.M) : make_map(&ta, A, g.M, g.K, TC_BM);
        if (rc) return rc;
        rc = BMN ? make_map_mn(&tb, B, g.K, g.N) : make_map(&tb, B, g.N, g.K, BNL);
    }
    if (rc) return rc;
    static const int no_alt = getenv("GAOT_GEMM_NO_ALT") ? 1 : 0;
    g.no_alt = no_alt;
    g.gm = (g.M + TC_BM * CG - 1) / (TC_BM * CG);
    g.gn = (g.N + BN - 1) / BN;
    g.kb_real = (g.K + TC_BK - 1) / TC_BK;
    const int64_t kblocks = g.kb_real * (SPLIT ? 3 : 1), tiles = g.gm * g.gn;   // SPLIT: three terms along the contraction
    if (SPLIT) g.K = kblocks * TC_BK;                           // the kernel derives its k-block count from K
    const int workers = sm_count() / CG;                        // CTAs (CG = 1) or CTA pairs (CG = 2) resident on the chip
    int splits = 1;
    if (!g.out_bf16 && !g.act && !g.pre && tiles < workers && kblocks >= 8) {   // too few tiles for the chip: split the contraction
        splits = (int)(workers / tiles);                            // one round: every worker gets exactly one (tile, split)
        if (splits > kblocks / 4) splits = (int)(kblocks / 4);
        if (splits > 128) splits = 128;
        if (splits < 1) splits = 1;
    }
    if (splits > 1) {                                           // no empty splits: every tile must own >= 1 k-block
        const int64_t per = (kblocks + splits - 1) / splits;
        splits = (int)((kblocks + per - 1) / per);
    }
    g.splits = splits;
    if (splits > 1 && !g.accumulate) GAOT_CHECK(cudaMemsetAsync(g.C, 0, sizeof(float) * (size_t)(g.M * g.N), st));
    // fast epilogue (TMA stores): plain stores only (no split-K / accumulate / pre-activation copy), 16-byte row pitches,
    // and an `add` operand (fp32 output only) whose broadcast period is a whole number of 32-row boxes
    CUtensorMap tc = ta, tadd = ta;
    const bool reduce = splits > 1 || g.accumulate;
    g.fast = !g.pre && g.N % (g.out_bf16 ? 8 : 4) == 0 &&
             (!g.add || (!g.out_bf16 && g.add_rows % 32 == 0 && g.ldadd % 4 == 0 && !(reinterpret_cast<uintptr_t>(g.add) & 15))) &&
             (!g.bias || !(reinterpret_cast<uintptr_t>(g.bias) & 15));
    if (reduce) g.fast = (g.fast && !g.out_bf16 && !g.bias && !g.add && !g.act) ? 2 : 0;   // TMA reduce-add of fp32 partial tiles
    if (g.act == TC_ACT_SWIGLU_FWD || g.act == TC_ACT_SWIGLU_BWD) {
        if (!g.fast || BN < 128 || !g.out_bf16 || SPLIT) return GAOT_E_BADARG;   // these epilogues exist in the two-chunk bf16 pass only
        if (g.act == TC_ACT_SWIGLU_BWD) {                      // output and saved pre-activations: [M, 2N], boxes of 32 x 64
            rc = make_map_box32(&tc, g.C, g.M, 2 * g.N, 2 * g.N, true, 64);
            if (rc) return rc;
            rc = make_map_box32(&tadd, g.aux, g.M, 2 * g.N, 2 * g.N, true, 64);
        } else {                                                // pre-activations [M, N] as usual, U [M, N/2] in boxes of 32 x 32
            rc = make_map_box32(&tc, g.C, g.M, g.N, g.N, true, 64);
            if (rc) return rc;
            rc = make_map_box32(&tadd, g.aux, g.M, g.N / 2, g.N / 2, true, 32);
        }
        if (rc) return rc;
    } else if (g.fast) {
        rc = make_map_box32(&tc, g.C, g.M, g.N, g.N, g.out_bf16 != 0, (g.out_bf16 && BN >= 128) ? 64 : 32);
        if (rc) return rc;
        if (g.add) {
            rc = make_map_box32(&tadd, g.add, g.add_rows, g.N, g.ldadd, false);
            if (rc) return rc;
        }
    }
    constexpr int SMEM = TC_STAGES * (TC_BM * TC_BK * 2 + BNL * TC_BK * 2) + 8 * 8192 + 1024 + 512;
    static bool attr_done = false;
    if (!attr_done) {
        GAOT_CHECK(cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN, AMN, BMN, CG, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_done = true;
    }
    const int64_t total = tiles * splits;
    if (total > 0x7fffffffLL) return GAOT_E_BADARG;
    const unsigned grid = (unsigned)((total < workers ? total : workers) * CG);   // persistent: one CTA per SM
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(TC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GAOT_CHECK(cudaLaunchKernelEx(&cfg, gemm_bf16_tc_kernel<BN, AMN, BMN, CG, SPLIT>, ta, tb, tc, tadd, g));
    GAOT_RETURN_LAST();
}

// N >= 256: CTA pairs on 256 x 256 tiles; narrower outputs: one CTA per 128 x 128 / 128 x 64 tile
template <bool AMN, bool BMN, bool SPLIT = false>
int dispatch_tc(const void* A, const void* B, TcArgs& g, cudaStream_t st) {
    if (g.N <= 64) return launch_tc<64, AMN, BMN, 1, SPLIT>(A, B, g, st);
    if (g.N < 256 || getenv("GAOT_GEMM_NO_PAIR")) return launch_tc<128, AMN, BMN, 1, SPLIT>(A, B, g, st);
    return launch_tc<256, AMN, BMN, 2, SPLIT>(A, B, g, st);
}

}  // namespace

GAOT_API int gaot_gemm_bf16(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int out_bf16,
                            const float* bias, const float* add, int64_t add_rows, int64_t ldadd, int act, float* pre,
                            int accumulate, void* stream) {
    if (!A || !B || !C || M < 1 || N < 1 || K < 1) return GAOT_E_BADARG;
    if (K % 8 != 0 || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) ||
        (reinterpret_cast<uintptr_t>(C) & 15))
        return GAOT_E_ALIGN;                       // TMA: 16-byte aligned base and row pitch
    if (add && (add_rows < 1 || ldadd < N)) return GAOT_E_BADARG;
    if (out_bf16 && accumulate) return GAOT_E_BADARG;
    if (act < 0 || act > 2 || ((act || pre) && accumulate)) return GAOT_E_BADARG;
    if (M > 0x7fffffffLL || N > 0x7fffffffLL || K > 0x7fffffffLL) return GAOT_E_BADARG;
    TcArgs g{C, M, N, K, bias, add, add ? add_rows : 1, ldadd, out_bf16, 1, accumulate, 0, 0, act, pre, 0};
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_tc<false, false>(A, B, g, st);
}

GAOT_API int gaot_gemm_bf16_nn(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int out_bf16,
                               void* stream) {
    if (!A || !B || !C || M < 1 || N < 1 || K < 1) return GAOT_E_BADARG;
    if (K % 8 != 0 || N % 8 != 0 || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) ||
        (reinterpret_cast<uintptr_t>(C) & 15))
        return GAOT_E_ALIGN;
    if (M > 0x7fffffffLL || N > 0x7fffffffLL || K > 0x7fffffffLL) return GAOT_E_BADARG;
    TcArgs g{C, M, N, K, nullptr, nullptr, 1, 0, out_bf16, 1, 0, 0, 0, 0, nullptr, 0};
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_tc<false, true>(A, B, g, st);
}

GAOT_API int gaot_gemm_bf16_swiglu(const void* A, const void* B, void* AB, void* U, int64_t M, int64_t N, int64_t K,
                                   void* stream) {
    if (!A || !B || !AB || !U || M < 1 || N < 128 || K < 1 || N % 64 != 0) return GAOT_E_BADARG;
    if (K % 8 != 0 || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) ||
        (reinterpret_cast<uintptr_t>(AB) & 15) || (reinterpret_cast<uintptr_t>(U) & 15))
        return GAOT_E_ALIGN;
    if (M > 0x7fffffffLL || N > 0x7fffffffLL || K > 0x7fffffffLL) return GAOT_E_BADARG;
    TcArgs g{AB, M, N, K, nullptr, nullptr, 1, 0, 1, 1, 0, 0, 0, TC_ACT_SWIGLU_FWD, nullptr, 0};
    g.aux = U;
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_tc<false, false>(A, B, g, st);
}

GAOT_API int gaot_gemm_bf16_nn_swiglu_bwd(const void* dY, const void* W, const void* AB, void* dAB, int64_t M, int64_t N,
                                          int64_t K, void* stream) {
    if (!dY || !W || !AB || !dAB || M < 1 || N < 128 || K < 1 || N % 32 != 0) return GAOT_E_BADARG;
    if (K % 8 != 0 || (reinterpret_cast<uintptr_t>(dY) & 15) || (reinterpret_cast<uintptr_t>(W) & 15) ||
        (reinterpret_cast<uintptr_t>(AB) & 15) || (reinterpret_cast<uintptr_t>(dAB) & 15))
        return GAOT_E_ALIGN;
    if (M > 0x7fffffffLL || N > 0x3fffffffLL || K > 0x7fffffffLL) return GAOT_E_BADARG;
    TcArgs g{dAB, M, N, K, nullptr, nullptr, 1, 0, 1, 1, 0, 0, 0, TC_ACT_SWIGLU_BWD, nullptr, 0};
    g.aux = AB;
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_tc<false, true>(dY, W, g, st);
}

GAOT_API int gaot_gemm_bf16_tn(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t R, int accumulate,
                               void* stream) {
    if (!A || !B || !C || M < 1 || N < 1 || R < 1) return GAOT_E_BADARG;
    if (M % 8 != 0 || N % 8 != 0 || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) ||
        (reinterpret_cast<uintptr_t>(C) & 15))
        return GAOT_E_ALIGN;                       // TMA: 16-byte aligned base and row pitch
    if (M > 0x7fffffffLL || N > 0x7fffffffLL || R > 0x7fffffffLL) return GAOT_E_BADARG;
    TcArgs g{C, M, N, R, nullptr, nullptr, 1, 0, 0, 1, accumulate, 0, 0, 0, nullptr, 0};
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_tc<true, true>(A, B, g, st);
}

GAOT_API int gaot_split_ws_floats(void) { return SPLIT_PARTS; }

GAOT_API int gaot_split_f16(const float* in, void* planes, float* scale, float* ws, int64_t n, void* stream) {
    if (!in || !planes || !scale || !ws || n < 1) return GAOT_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(planes) & 15) || n % 8 != 0) return GAOT_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t parts = (n / 4 + 256 * 8 - 1) / (256 * 8);
    if (parts > SPLIT_PARTS) parts = SPLIT_PARTS;
    if (parts < 1) parts = 1;
    absmax_partial_kernel<<<(unsigned)parts, 256, 0, st>>>(in, n, ws);
    GAOT_NOTE_LAUNCHES(1);
    __half* hi = reinterpret_cast<__half*>(planes);
    split_f16_kernel<<<gaot_grid(n / 4 + 1, 256), 256, 0, st>>>(in, hi, hi + n, n, ws, (int)parts, scale);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_gemm_split(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int a_mn, int b_mn,
                             const float* sa, const float* sb, const float* bias, const float* add, int64_t add_rows,
                             int64_t ldadd, int act, float* pre, int accumulate, void* stream) {
    if (!A || !B || !C || !sa || !sb || M < 1 || N < 1 || K < 1) return GAOT_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (reinterpret_cast<uintptr_t>(C) & 15))
        return GAOT_E_ALIGN;
    // TMA: 16-byte row pitch of every plane (the stored row length is K for a K-major operand, M / N for an MN-major one)
    if ((a_mn ? M : K) % 8 != 0 || (b_mn ? N : K) % 8 != 0) return GAOT_E_ALIGN;
    if (a_mn && !b_mn) return GAOT_E_BADARG;                       // (A^T, B^T-stored) is not a form any layer needs
    if (add && (add_rows < 1 || ldadd < N)) return GAOT_E_BADARG;
    if (act < 0 || act > 2 || ((act || pre) && accumulate)) return GAOT_E_BADARG;
    if (M > 0x7fffffffLL || N > 0x7fffffffLL || 3 * K > 0x7fffffffLL) return GAOT_E_BADARG;
    TcArgs g{C, M, N, K, bias, add, add ? add_rows : 1, ldadd, 0, 1, accumulate, 0, 0, act, pre, 0, sa, sb, 0};
    cudaStream_t st = (cudaStream_t)stream;
    if (a_mn) return dispatch_tc<true, true, true>(A, B, g, st);
    if (b_mn) return dispatch_tc<false, true, true>(A, B, g, st);
    return dispatch_tc<false, false, true>(A, B, g, st);
}

GAOT_API int gaot_cast_bf16(const float* in, void* out, int64_t rows, int64_t cols, void* out_t, void* stream) {
    if (!in || (!out && !out_t) || rows < 0 || cols < 0) return GAOT_E_BADARG;
    if (rows == 0 || cols == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (out) {
        if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 7)) return GAOT_E_ALIGN;
        cast_bf16_kernel<<<gaot_grid(rows * cols / 4 + 1, 256), 256, 0, st>>>(in, reinterpret_cast<__nv_bfloat16*>(out), rows * cols);
        if (out_t) GAOT_NOTE_LAUNCHES(1);
    }
    if (out_t) {
        dim3 grid((unsigned)((cols + 63) / 64), (unsigned)((rows + 63) / 64));
        if (grid.y > 65535) return GAOT_E_BADARG;
        transpose_bf16_kernel<float><<<grid, 256, 0, st>>>(in, reinterpret_cast<__nv_bfloat16*>(out_t), rows, cols);
    }
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_transpose_bf16(const void* in, void* out_t, int64_t rows, int64_t cols, void* stream) {
    if (!in || !out_t || rows < 0 || cols < 0) return GAOT_E_BADARG;
    if (rows == 0 || cols == 0) return 0;
    dim3 grid((unsigned)((cols + 63) / 64), (unsigned)((rows + 63) / 64));
    if (grid.y > 65535) return GAOT_E_BADARG;
    transpose_bf16_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(in),
                                                                                 reinterpret_cast<__nv_bfloat16*>(out_t), rows, cols);
    GAOT_RETURN_LAST();
}

#ifdef EXP_TRACE
GAOT_API int gaot_debug_trace_gemm(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(g_trace)); }
#endif
