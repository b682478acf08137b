// Fused edge-weight kernel of AGNO (performance path, bf16 tensor cores):
//     w[e, :] = alpha[e] * MLP( cat[ y[index[e], :kc], x[row(e), :kc] ] ),   MLP = Linear(2kc -> H) GELU [Linear(H -> H) GELU]* Linear(H -> C)
// Reference: AGNO.forward, src/model/layers/agno.py:206-207 (gather / repeat of the coordinates), :229-242 (kernel MLP =
// LinearChannelMLP, mlp.py:329-337, exact-erf GELU), :250 (attention weight times kernel).  The reference materialises the
// [E, 2d] edge inputs and one [E, H] activation per layer (plus their gradients); unfused, that was ~30 passes over
// [E, 64] tensors per step and the largest part of the vx-mode (E = 3.6 M) and 3-D (E = 1.4 M) steps.
//
// Here a CTA owns 128-edge tiles (one edge = one TMEM lane = two threads, 32 columns each) and keeps the whole chain on chip:
//   layer 1 (K = 2kc <= 64, coordinates) on the FP32 pipe from registers (fp32, as everywhere else);
//   hidden / output layers as tcgen05.mma M=128, N=H|C, K=H with the bf16 activation tile written by the threads as a
//   K-major SWIZZLE_128B operand and the bf16 weights resident in shared memory for the CTA's lifetime;
//   accumulators in TMEM, read back by the owning thread for bias + GELU (+ alpha, store).
// Backward (edge_mlp_bwd_tc_kernel) recomputes the chain per tile, keeps the pre-activations as bf16 tiles in shared
// memory, back-propagates with the same MMAs (weights consumed MN-major in place) and accumulates the weight gradients
// dW_l += delta_l^T h_{l-1} ON THE TENSOR CORES into TMEM-resident M=64 accumulators for the CTA's lifetime (one
// reduction into global memory per CTA at the end); nothing of size [E, H] ever reaches HBM.
// Requires H == 64, C % 16 == 0, C <= 64, 2kc <= 64, 1 or 2 hidden (H x H) layers; other shapes use the unfused operators.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "../../include/gaot_b200.h"

namespace {

constexpr int EM_H = 64;                // hidden width = bf16 elements per 128-byte swizzle row
constexpr int EM_TILE = 128;            // edges per tile = TMEM lanes
constexpr int EM_THREADS = 256;          // two threads per edge row: each owns 32 of the 64 columns (warps w and w + 4 share a TMEM lane quadrant)
constexpr int EM_HW = 32;                // columns per thread
constexpr int EM_MAXHID = 2;            // H x H layers
constexpr int EM_MAXKIN = 64;
constexpr int EM_WBYTES = EM_H * EM_H * 2;          // one bf16 [64 rows x 64] weight tile (8 KB)
constexpr int EM_ABYTES = EM_TILE * EM_H * 2;       // one bf16 [128 x 64] activation tile (16 KB)

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
// K-major operand tile, 128-byte rows, SWIZZLE_128B, 8-row atoms 1024 bytes apart (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major view of the same [contraction rows][64 elements] tile: one 64-wide block, 8-row groups 1 KB apart
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
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
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sts_u4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t v) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}
// byte offset of the 16-byte piece `chunk` (0..7) of row r inside a SWIZZLE_128B tile of 128-byte rows
__device__ __forceinline__ uint32_t sw128_off(int r, int chunk) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
}
// GELU(x) = x Phi(x) and GELU'(x) = Phi(x) + x phi(x) from ONE exponential: erf by Abramowitz-Stegun 7.1.26
// (|error| < 1.5e-7, far below the bf16 rounding of every value these kernels hand to an MMA):
//   erf(u) = 1 - (a1 t + ... + a5 t^5) exp(-u^2),  t = 1 / (1 + p u),  u = |x| / sqrt(2),  exp(-u^2) = exp(-x^2 / 2) = sqrt(2 pi) phi(x)
// one MUFU.RCP + one MUFU.EX2 + ~12 FMA-pipe instructions, branch free (erff is ~2x that and branches on |x|).
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void gelu_both(float x, float& y, float& dy) {
    // constants folded: p u = (p / sqrt 2) |x|; the polynomial carries the factor 1/2 of (1 - erf) / 2
    float t;                                                         // 1 / (1 + p u): ONE MUFU.RCP (the argument is in [1, inf); __frcp_rn
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.23164189f, fabsf(x), 1.0f)));   // compiles to a Newton step plus a slow-path call)
    const float ex = ex2f(x * x * -0.72134752044448170f);           // exp(-x^2 / 2)
    float p = fmaf(0.5307027145f, t, -0.7265760135f);
    p = fmaf(p, t, 0.7107068705f);
    p = fmaf(p, t, -0.142248368f);
    p = fmaf(p, t, 0.127414796f);
    const float tail = p * t * ex;                                   // (1 - erf(u)) / 2
    const float cdf = x >= 0.f ? 1.0f - tail : tail;
    y = x * cdf;
    dy = fmaf(x * 0.39894228040143268f, ex, cdf);
}
__device__ __forceinline__ float gelu_fast(float x) {
    float y, dy;
    gelu_both(x, y, dy);
    return y;
}

struct EdgeMlpArgs {
    const float* y; const float* x;            // source / destination coordinates, row pitches dy / dx, first kc columns used
    int dy, dx, kc;
    const int64_t* index; const int64_t* edge_row;
    int64_t E;
    const float* alpha;                        // [E] attention weight (or 1 / degree), nullable = 1
    const float* W1; const float* b1;          // [H, 2kc], [H]
    const float* Wh[EM_MAXHID]; const float* bh[EM_MAXHID];   // [H, H], [H]
    const float* Wo; const float* bo;          // [C, H], [C]
    int n_hidden, C;
    float* w;                                  // forward: out [E, C]
    // backward
    const float* dw;                           // [E, C]
    float* dW1; float* db1; float* dWh[EM_MAXHID]; float* dbh[EM_MAXHID]; float* dWo; float* dbo;
};

// fp32 [rows x 64] global weight -> bf16 K-major SWIZZLE_128B tile of 64 rows (rows >= `rows` zero filled)
__device__ __forceinline__ void stage_weight(uint8_t* tile, const float* __restrict__ W, int rows) {
    for (int i = threadIdx.x; i < EM_H * 8; i += blockDim.x) {
        const int n = i >> 3, c = i & 7;
        uint32_t p[4] = {0u, 0u, 0u, 0u};
        if (n < rows) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(W + n * EM_H + c * 8));
            const float4 b = __ldg(reinterpret_cast<const float4*>(W + n * EM_H + c * 8 + 4));
            p[0] = pack_bf16(a.x, a.y); p[1] = pack_bf16(a.z, a.w); p[2] = pack_bf16(b.x, b.y); p[3] = pack_bf16(b.z, b.w);
        }
        sts_u4(smem_u32(tile) + sw128_off(n, c), p[0], p[1], p[2], p[3]);
    }
}

// one thread's half row (32 fp32 values, columns 32*hf ..) -> bf16 row r of a K-major SWIZZLE_128B tile
__device__ __forceinline__ void store_row_bf16(uint32_t tile, int r, int hf, const float (&v)[EM_HW]) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
        sts_u4(tile + sw128_off(r, 4 * hf + c), pack_bf16(v[8 * c], v[8 * c + 1]), pack_bf16(v[8 * c + 2], v[8 * c + 3]),
               pack_bf16(v[8 * c + 4], v[8 * c + 5]), pack_bf16(v[8 * c + 6], v[8 * c + 7]));
}

// explicit shared-state-space loads on 32-bit addresses (through the generic pointers derived from the aligned dynamic-smem
// base these reads compiled to LD.E with 64-bit address arithmetic)
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
// v[o] += bias[o] for this thread's 32 columns (warp-uniform 128-bit loads)
__device__ __forceinline__ void add_bias(uint32_t bias_s, float (&v)[EM_HW]) {
#pragma unroll
    for (int o = 0; o < EM_HW; o += 4) {
        const float4 b4 = lds_f4(bias_s + (uint32_t)o * 4u);
        v[o] += b4.x; v[o + 1] += b4.y; v[o + 2] += b4.z; v[o + 3] += b4.w;
    }
}
// layer 1 on the FP32 pipe, this thread's 32 outputs: z[o] = b1[o] + sum_i W1[o, i] * z0[i]; W1 / b1 in shared memory
template <int KIN>
__device__ __forceinline__ void layer1(const float* sW1, const float* sb1, int hf, const float (&z0)[KIN], float (&z)[EM_HW]) {
    const uint32_t w_s = smem_u32(sW1) + (uint32_t)(hf * EM_HW * KIN) * 4u, b_s = smem_u32(sb1) + (uint32_t)(hf * EM_HW) * 4u;
#pragma unroll
    for (int o = 0; o < EM_HW; ++o) {
        float acc = lds_f32(b_s + (uint32_t)o * 4u);
        if constexpr (KIN % 4 == 0) {
#pragma unroll
            for (int i = 0; i < KIN; i += 4) {
                const float4 w4 = lds_f4(w_s + (uint32_t)(o * KIN + i) * 4u);
                acc = fmaf(w4.x, z0[i], acc); acc = fmaf(w4.y, z0[i + 1], acc); acc = fmaf(w4.z, z0[i + 2], acc); acc = fmaf(w4.w, z0[i + 3], acc);
            }
        } else {
#pragma unroll
            for (int i = 0; i < KIN; ++i) acc = fmaf(lds_f32(w_s + (uint32_t)(o * KIN + i) * 4u), z0[i], acc);
        }
        z[o] = acc;
    }
}

// issue D[128 x N] = A[128 x 64] (K-major tile) * W^T (K-major weight tile [N rows x 64]) and commit to `bar`
__device__ __forceinline__ void mma_forward(uint32_t tmem_acc, uint32_t a_tile, uint32_t w_tile, int N, uint64_t* bar) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = desc_k(a_tile), db = desc_k(w_tile);
#pragma unroll
    for (int k = 0; k < EM_H / 16; ++k) umma_f16(tmem_acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k > 0 ? 1u : 0u);
    umma_commit(bar);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int KIN>
__global__ void __launch_bounds__(EM_THREADS, 3)
edge_mlp_fwd_tc_kernel(const EdgeMlpArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sWt = smem;                                             // [n_hidden + 1] weight tiles
    uint8_t* sA = sWt + (EM_MAXHID + 1) * EM_WBYTES;                 // activation tile
    float* sW1 = reinterpret_cast<float*>(sA + EM_ABYTES);           // [H, KIN]
    float* sb = sW1 + EM_H * KIN;                                    // [1 + MAXHID + 1][64] biases
    uint64_t* bar = reinterpret_cast<uint64_t*>(sb + (EM_MAXHID + 2) * EM_H);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, hf = warp >> 2;                          // TMEM lane quadrant, column half
    const int nh = g.n_hidden, C = g.C;

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    for (int l = 0; l < nh; ++l) stage_weight(sWt + l * EM_WBYTES, g.Wh[l], EM_H);
    stage_weight(sWt + nh * EM_WBYTES, g.Wo, C);
    for (int i = threadIdx.x; i < EM_H * KIN; i += blockDim.x) sW1[i] = __ldg(g.W1 + i);
    for (int i = threadIdx.x; i < EM_H; i += blockDim.x) {
        sb[i] = __ldg(g.b1 + i);
        for (int l = 0; l < nh; ++l) sb[(1 + l) * EM_H + i] = __ldg(g.bh[l] + i);
        sb[(1 + nh) * EM_H + i] = i < C ? __ldg(g.bo + i) : 0.f;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t tacc = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * EM_HW);   // this thread's lane and column half
    const uint32_t a_tile = smem_u32(sA);
    const int r = q * 32 + lane;                                     // edge row inside the tile == TMEM lane
    uint32_t phase = 0;
    constexpr int KC = KIN / 2;

    const int64_t n_tiles = (g.E + EM_TILE - 1) / EM_TILE;
    // the edge's endpoints and attention weight are fetched one tile ahead: the dependent chain index -> coordinates ->
    // layer 1 otherwise starts every tile with two exposed memory round trips
    int nj = 0, nm = 0;
    float nal = 1.f;
    {
        const int64_t e0 = (int64_t)blockIdx.x * EM_TILE + r;
        if (e0 < g.E) { nj = (int)__ldg(g.index + e0); nm = (int)__ldg(g.edge_row + e0); nal = g.alpha ? __ldg(g.alpha + e0) : 1.f; }
    }
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t e = t * EM_TILE + r;
        const bool ok = e < g.E;
        const int j = nj, m = nm;
        const float al = nal;
        {
            const int64_t en = e + (int64_t)gridDim.x * EM_TILE;
            if (en < g.E) { nj = (int)__ldg(g.index + en); nm = (int)__ldg(g.edge_row + en); nal = g.alpha ? __ldg(g.alpha + en) : 1.f; }
        }
        float v[EM_HW];
        {
            float z0[KIN];
#pragma unroll
            for (int i = 0; i < KC; ++i) { z0[i] = __ldg(g.y + (int64_t)j * g.dy + i); z0[KC + i] = __ldg(g.x + (int64_t)m * g.dx + i); }
            layer1<KIN>(sW1, sb, hf, z0, v);
        }
#pragma unroll
        for (int o = 0; o < EM_HW; ++o) v[o] = gelu_fast(v[o]);
        for (int l = 0; l <= nh; ++l) {
            store_row_bf16(a_tile, r, hf, v);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> async proxy (MMA)
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                mma_forward(tmem, a_tile, smem_u32(sWt + l * EM_WBYTES), l < nh ? EM_H : C, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t bl_s = smem_u32(sb) + (uint32_t)((1 + l) * EM_H + hf * EM_HW) * 4u;
            uint32_t acc[32];
            tmem_ld32(tacc, acc);
            tmem_wait_ld();
#pragma unroll
            for (int o = 0; o < EM_HW; ++o) v[o] = __uint_as_float(acc[o]);
            add_bias(bl_s, v);
            if (l < nh) {
#pragma unroll
                for (int o = 0; o < EM_HW; ++o) v[o] = gelu_fast(v[o]);
            }
        }
        if (ok) {
            float* dst = g.w + e * C + hf * EM_HW;
#pragma unroll
            for (int o = 0; o < EM_HW; o += 4)
                if (hf * EM_HW + o < C) *reinterpret_cast<float4*>(dst + o) = make_float4(v[o] * al, v[o + 1] * al, v[o + 2] * al, v[o + 3] * al);
        }
        // the next tile's store_row_bf16 overwrites the activation tile: its last reader (the output MMA) has completed
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u));
    }
}

template <int KIN>
int launch_fwd(const EdgeMlpArgs& g, cudaStream_t st) {
    constexpr int EM_FWD_SMEM = (EM_MAXHID + 1) * EM_WBYTES + EM_ABYTES + 4 * EM_H * KIN + 4 * (EM_MAXHID + 2) * EM_H + 64 + 1024;
    static bool done = false;
    if (!done) {
        GAOT_CHECK(cudaFuncSetAttribute(edge_mlp_fwd_tc_kernel<KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, EM_FWD_SMEM));
        done = true;
    }
    const int64_t n_tiles = (g.E + EM_TILE - 1) / EM_TILE;
    const int64_t cap = (int64_t)gaot_sm_count() * 3;
    edge_mlp_fwd_tc_kernel<KIN><<<(unsigned)(n_tiles < cap ? n_tiles : cap), EM_THREADS, EM_FWD_SMEM, st>>>(g);
    GAOT_RETURN_LAST();
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// read this thread's half of a bf16 row of a SWIZZLE_128B tile back as fp32
__device__ __forceinline__ void load_row_bf16(uint32_t tile, int r, int hf, float (&v)[EM_HW]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint4 qd = lds_u4(tile + sw128_off(r, 4 * hf + c));
        const float2 a = unpack_bf16(qd.x), b = unpack_bf16(qd.y), cc = unpack_bf16(qd.z), d = unpack_bf16(qd.w);
        v[8 * c] = a.x; v[8 * c + 1] = a.y; v[8 * c + 2] = b.x; v[8 * c + 3] = b.y;
        v[8 * c + 4] = cc.x; v[8 * c + 5] = cc.y; v[8 * c + 6] = d.x; v[8 * c + 7] = d.y;
    }
}
__device__ __forceinline__ float lds_bf16(uint32_t a) {
    unsigned short u;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(u) : "r"(a));
    return __uint_as_float((uint32_t)u << 16);
}

// g[128 x 64] = delta (K-major tile) * W (weight tile [out rows x 64] consumed MN-major: contraction over the rows), then
// dW_acc[64 x 64] (+)= delta^T * h (both tiles consumed MN-major, contraction over the 128 edges, M = 64), one commit
__device__ __forceinline__ void mma_backward(uint32_t tmem_g, uint32_t tmem_dw, uint32_t d_tile, uint32_t w_tile, uint32_t h_tile,
                                             bool first_tile, uint64_t* bar) {
    constexpr uint32_t BASE = (1u << 4) | (1u << 7) | (1u << 10);
    constexpr uint32_t IDESC_G = BASE | (1u << 16) | ((uint32_t)(EM_H >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);              // M=128, N=64
    constexpr uint32_t IDESC_W = BASE | (1u << 15) | (1u << 16) | ((uint32_t)(EM_H >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);   // M=64, N=64
    const uint64_t da = desc_k(d_tile), dbw = desc_mn(w_tile);
#pragma unroll
    for (int k = 0; k < EM_H / 16; ++k) umma_f16(tmem_g, da + (uint64_t)(2 * k), dbw + (uint64_t)((k * 2048) >> 4), IDESC_G, k > 0 ? 1u : 0u);
    const uint64_t dam = desc_mn(d_tile), dbh = desc_mn(h_tile);
#pragma unroll
    for (int k = 0; k < EM_TILE / 16; ++k)
        umma_f16(tmem_dw, dam + (uint64_t)((k * 2048) >> 4), dbh + (uint64_t)((k * 2048) >> 4), IDESC_W, (k > 0 || !first_tile) ? 1u : 0u);
    umma_commit(bar);
}

template <int KIN>
struct EmBwdSmem {
    static constexpr int Z0 = EM_TILE * (KIN + 1) * 4;
    static constexpr int BYTES = (EM_MAXHID + 1) * EM_WBYTES + (EM_MAXHID + 2) * EM_ABYTES + Z0 + 4 * EM_H * KIN + 4 * (EM_MAXHID + 2) * EM_H + 64 + 1024;
};

template <int KIN>
__global__ void __launch_bounds__(EM_THREADS, 2)
edge_mlp_bwd_tc_kernel(const EdgeMlpArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sWt = smem;                                             // [n_hidden + 1] weight tiles
    uint8_t* sZ = sWt + (EM_MAXHID + 1) * EM_WBYTES;                 // [n_hidden] pre-activation tiles z_1 .. z_nh (bf16)
    uint8_t* sS = sZ + EM_MAXHID * EM_ABYTES;                        // operand tile: h_l
    uint8_t* sD = sS + EM_ABYTES;                                    // operand tile: delta_l
    float* sZ0 = reinterpret_cast<float*>(sD + EM_ABYTES);           // [128][KIN + 1]: edge inputs, last column 1
    float* sW1 = sZ0 + EM_TILE * (KIN + 1);                          // [H, KIN]
    float* sb = sW1 + EM_H * KIN;                                    // [1 + MAXHID + 1][64] biases
    uint64_t* bar = reinterpret_cast<uint64_t*>(sb + (EM_MAXHID + 2) * EM_H);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, hf = warp >> 2;                          // TMEM lane quadrant, column half
    const int nh = g.n_hidden, C = g.C;

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    for (int l = 0; l < nh; ++l) stage_weight(sWt + l * EM_WBYTES, g.Wh[l], EM_H);
    stage_weight(sWt + nh * EM_WBYTES, g.Wo, C);
    for (int i = threadIdx.x; i < EM_H * KIN; i += blockDim.x) sW1[i] = __ldg(g.W1 + i);
    for (int i = threadIdx.x; i < EM_H; i += blockDim.x) {
        sb[i] = __ldg(g.b1 + i);
        for (int l = 0; l < nh; ++l) sb[(1 + l) * EM_H + i] = __ldg(g.bh[l] + i);
        sb[(1 + nh) * EM_H + i] = i < C ? __ldg(g.bo + i) : 0.f;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    // this thread's TMEM lane and column half; columns [0,64): activations / g, [64,128) ...: weight-gradient accumulators
    const uint32_t tacc = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * EM_HW);
    const uint32_t s_tile = smem_u32(sS), d_tile = smem_u32(sD);
    const int r = q * 32 + lane;
    uint32_t phase = 0;
    constexpr int KC = KIN / 2;
    // dW1 | db1 on the FP32 pipe: thread-owned (output o, input i) pairs, i == KIN is the bias (ones column of sZ0)
    // first-layer gradients: thread = (output column o = t % 64, quarter t / 64 of the tile's edges); it keeps dW1[o, 0..KIN) and
    // db1[o] of its edges: one delta read per edge serves all KIN + 1 products, the edge inputs are warp-uniform loads
    static_assert(EM_THREADS == 4 * EM_H, "first-layer gradient mapping: 64 columns x 4 edge quarters");
    float acc1[KIN + 1];
#pragma unroll
    for (int p = 0; p <= KIN; ++p) acc1[p] = 0.f;
    float dbacc[EM_MAXHID + 1];                                      // column sums of delta (bias gradients of bh[.] / bo): thread t
#pragma unroll                                                       // owns column t % 64 over the 32 rows of quarter t / 64
    for (int l = 0; l <= EM_MAXHID; ++l) dbacc[l] = 0.f;

    auto sync_mma = [&]() {                                          // operand tiles written by all threads -> visible to the MMA
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    };
    auto wait_mma = [&]() {
        mbar_wait(bar, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };
    auto colsum = [&](int slot) {                                    // every thread: 32 rows of one column of the delta tile
        const int c = threadIdx.x & (EM_H - 1), e0 = (threadIdx.x >> 6) * 32;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
        for (int e = 0; e < 32; e += 2) {
            s0 += lds_bf16(d_tile + sw128_off(e0 + e, c >> 3) + (uint32_t)(c & 7) * 2u);
            s1 += lds_bf16(d_tile + sw128_off(e0 + e + 1, c >> 3) + (uint32_t)(c & 7) * 2u);
        }
        dbacc[slot] += s0 + s1;
    };
    auto gelu_pair = [&](float (&v)[EM_HW], uint32_t (&gp)[EM_HW / 2]) {   // v -> GELU(v), gp = GELU'(v) packed bf16
#pragma unroll
        for (int o = 0; o < EM_HW; o += 2) {
            float h0, h1, d0, d1;
            gelu_both(v[o], h0, d0); gelu_both(v[o + 1], h1, d1);
            v[o] = h0; v[o + 1] = h1;
            gp[o >> 1] = pack_bf16(d0, d1);
        }
    };

    const int64_t n_tiles = (g.E + EM_TILE - 1) / EM_TILE;
    bool first_tile = true;
    // endpoints / attention weight one tile ahead (see the forward kernel); the upstream-gradient rows of the next tile are
    // pulled towards the SM with prefetch instructions while this tile is being processed
    int nj = 0, nm = 0;
    float nal = 0.f;
    {
        const int64_t e0 = (int64_t)blockIdx.x * EM_TILE + r;
        if (e0 < g.E) { nj = (int)__ldg(g.index + e0); nm = (int)__ldg(g.edge_row + e0); nal = g.alpha ? __ldg(g.alpha + e0) : 1.f; }
    }
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, first_tile = false) {
        const int64_t e = t * EM_TILE + r;
        const bool ok = e < g.E;
        const int j = nj, m = nm;
        const float al = ok ? nal : 0.f;
        {
            const int64_t en = e + (int64_t)gridDim.x * EM_TILE;
            if (en < g.E) {
                nj = (int)__ldg(g.index + en); nm = (int)__ldg(g.edge_row + en); nal = g.alpha ? __ldg(g.alpha + en) : 1.f;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(g.dw + en * C + hf * EM_HW));
            }
        }
        float v[EM_HW];
        // output layer first: delta_out = alpha * dw (zero beyond C and for rows past E) goes straight into the delta tile
        // (free since the previous tile's last barrier; its rows were prefetched into the L2 one tile ago)
        {
            const float* src = g.dw + (ok ? e : 0) * C + hf * EM_HW;
#pragma unroll
            for (int o = 0; o < EM_HW; o += 4) {
                float4 qd = make_float4(0.f, 0.f, 0.f, 0.f);
                if (hf * EM_HW + o < C) qd = __ldg(reinterpret_cast<const float4*>(src + o));
                v[o] = qd.x * al; v[o + 1] = qd.y * al; v[o + 2] = qd.z * al; v[o + 3] = qd.w * al;
            }
            store_row_bf16(d_tile, r, hf, v);
        }
        uint32_t gp[EM_HW / 2];                                      // GELU'(z_l) of the layer being back-propagated
        {
            float z0[KIN];
#pragma unroll
            for (int i = 0; i < KC; ++i) { z0[i] = __ldg(g.y + (int64_t)j * g.dy + i); z0[KC + i] = __ldg(g.x + (int64_t)m * g.dx + i); }
            if (hf == 0) {
#pragma unroll
                for (int i = 0; i < KIN; ++i) sZ0[r * (KIN + 1) + i] = z0[i];
                sZ0[r * (KIN + 1) + KIN] = 1.f;
            }
            layer1<KIN>(sW1, sb, hf, z0, v);
        }
        // ---- forward recompute: z_1 .. z_{nh+1}; z_1..z_nh kept as bf16 tiles, the last one only through GELU' ----
        for (int l = 1; l <= nh + 1; ++l) {
            if (l <= nh) {
                store_row_bf16(smem_u32(sZ + (l - 1) * EM_ABYTES), r, hf, v);
#pragma unroll
                for (int o = 0; o < EM_HW; ++o) v[o] = gelu_fast(v[o]);
            } else {
                gelu_pair(v, gp);
            }
            store_row_bf16(s_tile, r, hf, v);
            if (l <= nh) {
                sync_mma();
                if (threadIdx.x == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    mma_forward(tmem, s_tile, smem_u32(sWt + (l - 1) * EM_WBYTES), EM_H, bar);
                }
                wait_mma();
                const uint32_t bl_s = smem_u32(sb) + (uint32_t)(l * EM_H + hf * EM_HW) * 4u;
                uint32_t acc[32];
                tmem_ld32(tacc, acc);
                tmem_wait_ld();
#pragma unroll
                for (int o = 0; o < EM_HW; ++o) v[o] = __uint_as_float(acc[o]);
                add_bias(bl_s, v);
            }
        }
        // ---- back-propagate: weight tile L = nh (Wo) down to 0 (Wh[0]) ----
        for (int L = nh; L >= 0; --L) {
            // delta tile (sD) and h tile (sS = h_{L+1}) are written: g = delta W_L, dW_L += delta^T h
            sync_mma();
            if (threadIdx.x == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                mma_backward(tmem, tmem + 64u * (uint32_t)(1 + (nh - L)), d_tile, smem_u32(sWt + L * EM_WBYTES), s_tile, first_tile, bar);
            }
            colsum(nh - L);                                          // bias gradient of this layer (reads sD beside the MMA)
            wait_mma();
            __syncthreads();                                         // every column-sum read of sD precedes its rewrite below
            // delta of the GELU layer below: delta = g * GELU'(z)
            {
                uint32_t acc[32];
                tmem_ld32(tacc, acc);
                tmem_wait_ld();
#pragma unroll
                for (int o = 0; o < EM_HW; o += 2) {
                    const float2 d = unpack_bf16(gp[o >> 1]);
                    v[o] = __uint_as_float(acc[o]) * d.x;
                    v[o + 1] = __uint_as_float(acc[o + 1]) * d.y;
                }
            }
            store_row_bf16(d_tile, r, hf, v);                        // its readers (both MMAs, the column sums) are done
            if (L > 0) {
                // h_L and GELU'(z_L) from the stored pre-activation tile (a thread reads back the half row it wrote)
                load_row_bf16(smem_u32(sZ + (L - 1) * EM_ABYTES), r, hf, v);
                gelu_pair(v, gp);
                store_row_bf16(s_tile, r, hf, v);
            }
        }
        // ---- first layer: dW1[o, i] += sum_e delta_1[e, o] z0[e, i], db1[o] += sum_e delta_1[e, o] (FP32 pipe) ----
        __syncthreads();
        {
            const int o = threadIdx.x & (EM_H - 1), e_lo = (threadIdx.x >> 6) * (EM_TILE / 4);
            const uint32_t z_s = smem_u32(sZ0), d_col = d_tile + (uint32_t)(o & 7) * 2u;
#pragma unroll 4
            for (int ee = e_lo; ee < e_lo + EM_TILE / 4; ++ee) {
                const float d = lds_bf16(d_col + sw128_off(ee, o >> 3));
#pragma unroll
                for (int i = 0; i <= KIN; ++i) acc1[i] = fmaf(d, lds_f32(z_s + (uint32_t)(ee * (KIN + 1) + i) * 4u), acc1[i]);
            }
        }
        __syncthreads();                                             // next tile overwrites sZ0 / sS / sD
    }
    // ---- CTA-level results -> global (fp32 red.add) ----
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int L = nh; L >= 0; --L) {
        // M = 64 accumulator: row m lives in TMEM lane (m % 16) + 32 * (m / 16), i.e. in the first 16 lanes of quadrant m / 16
        float* dst = L == nh ? g.dWo : g.dWh[L];
        const int rows = L == nh ? C : EM_H;
        uint32_t acc[32];
        tmem_ld32(tacc + 64u * (uint32_t)(1 + (nh - L)), acc);
        tmem_wait_ld();
        const int row = q * 16 + lane;
        if (lane < 16 && row < rows) {
#pragma unroll
            for (int o = 0; o < EM_HW; ++o) atomicAdd(dst + row * EM_H + hf * EM_HW + o, __uint_as_float(acc[o]));
        }
        {
            float* db = L == nh ? g.dbo : g.dbh[L];
            const int c = threadIdx.x & (EM_H - 1);
            if (c < rows) atomicAdd(db + c, dbacc[nh - L]);
        }
    }
    {
        const int o = threadIdx.x & (EM_H - 1);
#pragma unroll
        for (int i = 0; i < KIN; ++i) atomicAdd(g.dW1 + o * KIN + i, acc1[i]);
        atomicAdd(g.db1 + o, acc1[KIN]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
    }
}

template <int KIN>
int launch_bwd(const EdgeMlpArgs& g, cudaStream_t st) {
    constexpr int SMEM = EmBwdSmem<KIN>::BYTES;
    static_assert(SMEM <= 227 * 1024, "edge-MLP backward tile set exceeds shared memory");
    static bool done = false;
    if (!done) {
        GAOT_CHECK(cudaFuncSetAttribute(edge_mlp_bwd_tc_kernel<KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        done = true;
    }
    const int64_t n_tiles = (g.E + EM_TILE - 1) / EM_TILE;
    const int64_t cap = (int64_t)gaot_sm_count() * (SMEM <= 110 * 1024 ? 2 : 1);
    edge_mlp_bwd_tc_kernel<KIN><<<(unsigned)(n_tiles < cap ? n_tiles : cap), EM_THREADS, SMEM, st>>>(g);
    GAOT_RETURN_LAST();
}

int check_args(const EdgeMlpArgs& g) {
    if (!g.y || !g.x || !g.index || !g.edge_row || !g.W1 || !g.b1 || !g.Wo || !g.bo || g.E < 0) return GAOT_E_BADARG;
    if (g.kc < 1 || 2 * g.kc > EM_MAXKIN || g.dy < g.kc || g.dx < g.kc) return GAOT_E_BADARG;
    if (g.n_hidden < 1 || g.n_hidden > EM_MAXHID || g.C < 16 || g.C > EM_H || g.C % 16) return GAOT_E_BADARG;
    for (int l = 0; l < g.n_hidden; ++l)
        if (!g.Wh[l] || !g.bh[l] || (reinterpret_cast<uintptr_t>(g.Wh[l]) & 15)) return GAOT_E_BADARG;
    if (reinterpret_cast<uintptr_t>(g.Wo) & 15) return GAOT_E_ALIGN;
    return 0;
}

#define GAOT_EM_DISPATCH_KIN(KIN_, FN, ...)                                   \
    switch (KIN_) {                                                          \
        case 4: return FN<4>(__VA_ARGS__);                                   \
        case 6: return FN<6>(__VA_ARGS__);                                   \
        case 32: return FN<32>(__VA_ARGS__);                                 \
        case 48: return FN<48>(__VA_ARGS__);                                 \
        default: return GAOT_E_BADARG;                                       \
    }

}  // namespace

GAOT_API int gaot_edge_mlp_supported(int64_t kin, int64_t H, int64_t n_hidden, int64_t C) {
    return (kin == 4 || kin == 6 || kin == 32 || kin == 48) && H == EM_H && n_hidden >= 1 && n_hidden <= EM_MAXHID && C >= 16 &&
           C <= EM_H && C % 16 == 0;
}

GAOT_API int gaot_edge_mlp_fwd(const float* y, int64_t dy, const float* x, int64_t dx, int64_t kc, const int64_t* index,
                               const int64_t* edge_row, int64_t E, const float* alpha, const float* W1, const float* b1,
                               const float* Wh0, const float* bh0, const float* Wh1, const float* bh1, int64_t n_hidden,
                               const float* Wo, const float* bo, int64_t C, float* w, void* stream) {
    EdgeMlpArgs g{};
    g.y = y; g.x = x; g.dy = (int)dy; g.dx = (int)dx; g.kc = (int)kc; g.index = index; g.edge_row = edge_row; g.E = E; g.alpha = alpha;
    g.W1 = W1; g.b1 = b1; g.Wh[0] = Wh0; g.bh[0] = bh0; g.Wh[1] = Wh1; g.bh[1] = bh1; g.n_hidden = (int)n_hidden;
    g.Wo = Wo; g.bo = bo; g.C = (int)C; g.w = w;
    int rc = check_args(g);
    if (rc) return rc;
    if (!w || (reinterpret_cast<uintptr_t>(w) & 15)) return GAOT_E_BADARG;
    if (E == 0) return 0;
    GAOT_EM_DISPATCH_KIN(2 * (int)kc, launch_fwd, g, (cudaStream_t)stream);
}

GAOT_API int gaot_edge_mlp_bwd(const float* y, int64_t dy, const float* x, int64_t dx, int64_t kc, const int64_t* index,
                               const int64_t* edge_row, int64_t E, const float* alpha, const float* W1, const float* b1,
                               const float* Wh0, const float* bh0, const float* Wh1, const float* bh1, int64_t n_hidden,
                               const float* Wo, const float* bo, int64_t C, const float* dw, float* dW1, float* db1, float* dWh0,
                               float* dbh0, float* dWh1, float* dbh1, float* dWo, float* dbo, void* stream) {
    EdgeMlpArgs g{};
    g.y = y; g.x = x; g.dy = (int)dy; g.dx = (int)dx; g.kc = (int)kc; g.index = index; g.edge_row = edge_row; g.E = E; g.alpha = alpha;
    g.W1 = W1; g.b1 = b1; g.Wh[0] = Wh0; g.bh[0] = bh0; g.Wh[1] = Wh1; g.bh[1] = bh1; g.n_hidden = (int)n_hidden;
    g.Wo = Wo; g.bo = bo; g.C = (int)C; g.dw = dw;
    g.dW1 = dW1; g.db1 = db1; g.dWh[0] = dWh0; g.dbh[0] = dbh0; g.dWh[1] = dWh1; g.dbh[1] = dbh1; g.dWo = dWo; g.dbo = dbo;
    int rc = check_args(g);
    if (rc) return rc;
    if (!dw || !dW1 || !db1 || !dWo || !dbo || (reinterpret_cast<uintptr_t>(dw) & 15)) return GAOT_E_BADARG;
    for (int l = 0; l < g.n_hidden; ++l)
        if (!g.dWh[l] || !g.dbh[l]) return GAOT_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    // the gradients are reduced into with red.add by every CTA: zero them first
    GAOT_CHECK(cudaMemsetAsync(dW1, 0, sizeof(float) * (size_t)(EM_H * 2 * kc), st));
    GAOT_CHECK(cudaMemsetAsync(db1, 0, sizeof(float) * EM_H, st));
    for (int l = 0; l < g.n_hidden; ++l) {
        GAOT_CHECK(cudaMemsetAsync(g.dWh[l], 0, sizeof(float) * EM_H * EM_H, st));
        GAOT_CHECK(cudaMemsetAsync(g.dbh[l], 0, sizeof(float) * EM_H, st));
    }
    GAOT_CHECK(cudaMemsetAsync(dWo, 0, sizeof(float) * (size_t)(C * EM_H), st));
    GAOT_CHECK(cudaMemsetAsync(dbo, 0, sizeof(float) * (size_t)C, st));
    if (E == 0) return 0;
    GAOT_EM_DISPATCH_KIN(2 * (int)kc, launch_bwd, g, st);
}
