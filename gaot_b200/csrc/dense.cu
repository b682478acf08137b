// fp32 dense kernels of the parity path: tiled FFMA GEMM with fused bias / broadcast-add /
// activation epilogue and split-K, column sums, activation backward, RMSNorm, SwiGLU gate,
// patchify permutations, MSE and AdamW.
// Reference: src/model/layers/mlp.py:283-337, src/model/layers/attn.py:133-172, src/model/gaot.py:182-231,
//            src/core/base_trainer.py:71, src/utils/optimizers.py:196,252-255.
#include "common.cuh"
#include "../../include/gaot_b200.h"

namespace {

constexpr int GEMM_THREADS = 256;
constexpr int BK = 16;

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == GAOT_ACT_GELU) return gelu_erf(v);
    if (act == GAOT_ACT_RELU) return fmaxf(v, 0.f);
    return v;
}

// Stage a [ROWS x BK] operand tile into smem as tile[k][row].
//   kcontig = true : src[row*ld + k]   (row-major [rows, K], e.g. activations or nn.Linear weights)
//   kcontig = false: src[k*ld + row]   ([K, rows])
template <int ROWS>
struct TileLoader {
    static constexpr int VECS = ROWS * BK / 4;                 // float4 slots per tile
    static constexpr int ITERS = (VECS + GEMM_THREADS - 1) / GEMM_THREADS;
    float4 reg[ITERS];

    __device__ __forceinline__ void fetch(const float* __restrict__ src, int64_t ld, bool kcontig, bool vec_ok,
                                          int64_t row0, int64_t nrows, int64_t k0, int64_t kend) {
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int s = threadIdx.x + it * GEMM_THREADS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s >= VECS) { reg[it] = v; continue; }
            if (kcontig) {
                const int r = s / (BK / 4), k4 = (s % (BK / 4)) * 4;
                const int64_t row = row0 + r, k = k0 + k4;
                if (row < nrows) {
                    const float* p = src + row * ld + k;
                    if (vec_ok && k + 3 < kend) v = __ldg(reinterpret_cast<const float4*>(p));
                    else {
                        if (k < kend) v.x = __ldg(p);
                        if (k + 1 < kend) v.y = __ldg(p + 1);
                        if (k + 2 < kend) v.z = __ldg(p + 2);
                        if (k + 3 < kend) v.w = __ldg(p + 3);
                    }
                }
            } else {
                const int kk = s / (ROWS / 4), r4 = (s % (ROWS / 4)) * 4;
                const int64_t row = row0 + r4, k = k0 + kk;
                if (k < kend) {
                    const float* p = src + k * ld + row;
                    if (vec_ok && row + 3 < nrows) v = __ldg(reinterpret_cast<const float4*>(p));
                    else {
                        if (row < nrows) v.x = __ldg(p);
                        if (row + 1 < nrows) v.y = __ldg(p + 1);
                        if (row + 2 < nrows) v.z = __ldg(p + 2);
                        if (row + 3 < nrows) v.w = __ldg(p + 3);
                    }
                }
            }
            reg[it] = v;
        }
    }
    __device__ __forceinline__ void commit(float (*tile)[ROWS + 4], bool kcontig) const {
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int s = threadIdx.x + it * GEMM_THREADS;
            if (s >= VECS) continue;
            if (kcontig) {
                const int r = s / (BK / 4), k4 = (s % (BK / 4)) * 4;
                tile[k4 + 0][r] = reg[it].x; tile[k4 + 1][r] = reg[it].y;
                tile[k4 + 2][r] = reg[it].z; tile[k4 + 3][r] = reg[it].w;
            } else {
                const int kk = s / (ROWS / 4), r4 = (s % (ROWS / 4)) * 4;
                *reinterpret_cast<float4*>(&tile[kk][r4]) = reg[it];
            }
        }
    }
};

struct GemmArgs {
    const float* A; const float* B; float* C;
    int64_t M, N, K, lda, ldb, ldc;
    int transA, transB;
    const float* bias; const float* add; int64_t add_rows, ldadd;
    int act; float* pre; int accumulate; int splits;
    int vecA, vecB;
};

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_f32_kernel(const GemmArgs g) {
    static_assert((BM / TM) * (BN / TN) == GEMM_THREADS, "thread tiling");
    static_assert(TM == 8, "row micro-tile is two groups of 4");
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tx = threadIdx.x % (BN / TN), ty = threadIdx.x / (BN / TN);
    const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
    // K range of this split
    const int64_t ktiles = (g.K + BK - 1) / BK;
    const int64_t per = (ktiles + g.splits - 1) / g.splits;
    const int64_t kbeg = (int64_t)blockIdx.z * per * BK;
    const int64_t kend = min(g.K, kbeg + per * BK);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    TileLoader<BM> la;
    TileLoader<BN> lb;
    const bool akc = !g.transA, bkc = g.transB != 0;
    if (kbeg < kend) {
        la.fetch(g.A, g.lda, akc, g.vecA, m0, g.M, kbeg, kend);
        lb.fetch(g.B, g.ldb, bkc, g.vecB, n0, g.N, kbeg, kend);
    }
    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
        la.commit(As, akc);
        lb.commit(Bs, bkc);
        __syncthreads();
        if (k0 + BK < kend) {
            la.fetch(g.A, g.lda, akc, g.vecA, m0, g.M, k0 + BK, kend);
            lb.fetch(g.B, g.ldb, bkc, g.vecB, n0, g.N, k0 + BK, kend);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][BM / 2 + ty * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            if constexpr (TN == 8) {
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][BN / 2 + tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
            } else if constexpr (TN == 4) {
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    const bool first = blockIdx.z == 0;
    const bool atomic = g.splits > 1;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t m = m0 + ((i < 4) ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int64_t n = n0 + ((TN == 8) ? ((j < 4) ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4)) : tx * TN + j);
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (first) {
                if (g.bias) v += __ldg(g.bias + n);
                if (g.add) v += __ldg(g.add + (m % g.add_rows) * g.ldadd + n);
            }
            float* cp = g.C + m * g.ldc + n;
            if (atomic) { atomicAdd(cp, v); continue; }
            if (g.pre) g.pre[m * g.ldc + n] = v;
            v = apply_act(v, g.act);
            if (g.accumulate) v += *cp;
            *cp = v;
        }
    }
}

template <int BM, int BN, int TM, int TN>
int launch_gemm(GemmArgs& g, cudaStream_t st) {
    const int64_t gm = (g.M + BM - 1) / BM, gn = (g.N + BN - 1) / BN;
    const int64_t tiles = gm * gn, ktiles = (g.K + BK - 1) / BK;
    int splits = 1;
    const bool can_split = (g.act == GAOT_ACT_NONE) && !g.pre && (g.ldc == g.N);
    if (can_split && tiles < gaot_sm_count() && ktiles >= 16) {
        splits = (int)((gaot_sm_count() * 4 + tiles - 1) / tiles);
        if (splits > ktiles / 8) splits = (int)(ktiles / 8);
        if (splits < 1) splits = 1;
        if (splits > 1024) splits = 1024;
    }
    g.splits = splits;
    if (splits > 1 && !g.accumulate) GAOT_CHECK(cudaMemsetAsync(g.C, 0, sizeof(float) * (size_t)(g.M * g.N), st));
    if (gn > 65535 || gm > 0x7fffffffLL) return GAOT_E_BADARG;
    dim3 grid((unsigned)gm, (unsigned)gn, (unsigned)splits);
    gemm_f32_kernel<BM, BN, TM, TN><<<grid, GEMM_THREADS, 0, st>>>(g);
    GAOT_RETURN_LAST();
}

// ------------------------------------------------------------------------------------------------
// Narrow GEMMs (both small dimensions <= 64): the MAGNO kernel MLP (reference mlp.py:329-337 on E edge
// rows, hidden 64), the geometry-embedding MLP and their backward.  A 128x64 smem-tiled GEMM is
// latency bound at K = 64; here the 64x64 weight block lives in REGISTERS (lane l owns output columns
// l and l+32 -> 2 x 64 weights), a warp streams rows four at a time through a 1 KB smem staging slab
// that is read back with broadcast 128-bit loads (1 LDS per 8 FFMA), so the kernel is FFMA bound.
//   ROWS mode:  C[r, n] = act( sum_k A[r,k] * Wt(k,n) + bias[n] + add[r % R, n] )   r over M rows
//               (forward: Wt(k,n) = B[n*ldb+k]; input gradient: Wt(k,n) = B[k*ldb+n])
//   RED  mode:  C[m, n] = sum_r A[r,m] * B[r,n]     (weight gradient; lane owns rows m = l, l+32 of C,
//               64 accumulators each; warps reduce through fp32 atomics into the zeroed C)
// ------------------------------------------------------------------------------------------------
constexpr int NW_THREADS = 256;
constexpr int NW_ROWS = 8;               // rows in flight per warp

template <bool BT>
__global__ void __launch_bounds__(NW_THREADS, 1) narrow_rows_kernel(const GemmArgs g) {
    __shared__ __align__(16) float slab_raw[(NW_THREADS / 32) * NW_ROWS * 64 + 64];   // + 64: doubles as the 64 x 65 weight staging
    float (*slab)[NW_ROWS][64] = reinterpret_cast<float (*)[NW_ROWS][64]>(slab_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = (int)g.K, N = (int)g.N;
    const int n0 = lane, n1 = lane + 32;
    // the weight block reaches the registers through shared memory: one coalesced pass of the CTA over B, then every lane
    // picks its two columns (a lane reading its own row of B straight from global memory touches 32 lines per instruction,
    // 128 instructions per warp - that prologue was most of the kernel at a few thousand rows)
    float w0[64], w1[64];
    {
        float* ws = slab_raw;                                    // 64 x 65 floats
        static_assert(sizeof(slab_raw) >= 64 * 65 * sizeof(float), "weight staging does not fit the row slab");
        for (int i = threadIdx.x; i < 64 * 64; i += NW_THREADS) {
            const int r = i >> 6, c = i & 63;                    // element (r, c) of B as stored: BT [N, K] -> (n, k); else [K, N] -> (k, n)
            const int n = BT ? r : c, k = BT ? c : r;
            ws[n * 65 + k] = (n < N && k < K) ? __ldg(g.B + (int64_t)r * g.ldb + c) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 64; ++k) { w0[k] = ws[n0 * 65 + k]; w1[k] = ws[n1 * 65 + k]; }
        __syncthreads();
    }
    const float b0 = (g.bias && n0 < N) ? __ldg(g.bias + n0) : 0.f, b1 = (g.bias && n1 < N) ? __ldg(g.bias + n1) : 0.f;
    float (*my)[64] = slab[warp];
    const int kc_n = (K + 3) >> 2;
    const int64_t gw = (int64_t)blockIdx.x * (NW_THREADS / 32) + warp, nw = (int64_t)gridDim.x * (NW_THREADS / 32);
    // software pipeline: the next group's rows are fetched into registers while this group is computed
    float pa[NW_ROWS], pb[NW_ROWS];
    auto fetch = [&](int64_t r0) {
#pragma unroll
        for (int rr = 0; rr < NW_ROWS; ++rr) {
            const int64_t r = min(r0 + rr, g.M - 1);
            const float* ap = g.A + r * g.lda;
            pa[rr] = (lane < K) ? __ldg(ap + lane) : 0.f;
            pb[rr] = (lane + 32 < K) ? __ldg(ap + lane + 32) : 0.f;
        }
    };
    if (gw * NW_ROWS < g.M) fetch(gw * NW_ROWS);
    for (int64_t r0 = gw * NW_ROWS; r0 < g.M; r0 += nw * NW_ROWS) {
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < NW_ROWS; ++rr) { my[rr][lane] = pa[rr]; my[rr][lane + 32] = pb[rr]; }
        __syncwarp();
        if (r0 + nw * NW_ROWS < g.M) fetch(r0 + nw * NW_ROWS);
        float acc0[NW_ROWS], acc1[NW_ROWS];
#pragma unroll
        for (int rr = 0; rr < NW_ROWS; ++rr) { acc0[rr] = b0; acc1[rr] = b1; }
#pragma unroll
        for (int kc = 0; kc < 16; ++kc) {
            if (kc < kc_n) {                             // warp-uniform
#pragma unroll
                for (int rr = 0; rr < NW_ROWS; ++rr) {
                    const float4 x = *reinterpret_cast<const float4*>(&my[rr][kc * 4]);
                    acc0[rr] = fmaf(x.x, w0[kc * 4 + 0], acc0[rr]); acc1[rr] = fmaf(x.x, w1[kc * 4 + 0], acc1[rr]);
                    acc0[rr] = fmaf(x.y, w0[kc * 4 + 1], acc0[rr]); acc1[rr] = fmaf(x.y, w1[kc * 4 + 1], acc1[rr]);
                    acc0[rr] = fmaf(x.z, w0[kc * 4 + 2], acc0[rr]); acc1[rr] = fmaf(x.z, w1[kc * 4 + 2], acc1[rr]);
                    acc0[rr] = fmaf(x.w, w0[kc * 4 + 3], acc0[rr]); acc1[rr] = fmaf(x.w, w1[kc * 4 + 3], acc1[rr]);
                }
            }
        }
#pragma unroll
        for (int rr = 0; rr < NW_ROWS; ++rr) {
            const int64_t r = r0 + rr;
            if (r >= g.M) break;
            float v0 = acc0[rr], v1 = acc1[rr];
            if (g.add) {
                const float* ar = g.add + (r % g.add_rows) * g.ldadd;
                if (n0 < N) v0 += __ldg(ar + n0);
                if (n1 < N) v1 += __ldg(ar + n1);
            }
            if (g.pre) {
                if (n0 < N) g.pre[r * g.ldc + n0] = v0;
                if (n1 < N) g.pre[r * g.ldc + n1] = v1;
            }
            v0 = apply_act(v0, g.act); v1 = apply_act(v1, g.act);
            float* cp = g.C + r * g.ldc;
            if (n0 < N) cp[n0] = g.accumulate ? cp[n0] + v0 : v0;
            if (n1 < N) cp[n1] = g.accumulate ? cp[n1] + v1 : v1;
        }
    }
}

__global__ void __launch_bounds__(NW_THREADS, 1) narrow_reduce_kernel(const GemmArgs g) {
    // C[m,n] (+)= sum_r A[r*lda + m] * B[r*ldb + n],  m < M <= 64, n < N <= 64, r < K (the long dimension)
    __shared__ __align__(16) float slab_raw[(NW_THREADS / 32) * NW_ROWS * 64 + 64];   // + 64: the reduction buffer is 64 x 65
    float (*slab)[NW_ROWS][64] = reinterpret_cast<float (*)[NW_ROWS][64]>(slab_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int M = (int)g.M, N = (int)g.N;
    const int m0 = lane, m1 = lane + 32;
    float c0[64], c1[64];
#pragma unroll
    for (int n = 0; n < 64; ++n) { c0[n] = 0.f; c1[n] = 0.f; }
    float (*my)[64] = slab[warp];
    const int nc_n = (N + 3) >> 2;
    const int64_t gw = (int64_t)blockIdx.x * (NW_THREADS / 32) + warp, nw = (int64_t)gridDim.x * (NW_THREADS / 32);
    float pa[NW_ROWS], pb[NW_ROWS], qa[NW_ROWS], qb[NW_ROWS];      // prefetched B row halves / own A values
    auto fetch = [&](int64_t r0) {
#pragma unroll
        for (int rr = 0; rr < NW_ROWS; ++rr) {
            const bool ok = r0 + rr < g.K;
            const int64_t r = ok ? r0 + rr : g.K - 1;
            const float* bp = g.B + r * g.ldb;
            pa[rr] = (ok && lane < N) ? __ldg(bp + lane) : 0.f;
            pb[rr] = (ok && lane + 32 < N) ? __ldg(bp + lane + 32) : 0.f;
            qa[rr] = (ok && m0 < M) ? __ldg(g.A + r * g.lda + m0) : 0.f;
            qb[rr] = (ok && m1 < M) ? __ldg(g.A + r * g.lda + m1) : 0.f;
        }
    };
    if (gw * NW_ROWS < g.K) fetch(gw * NW_ROWS);
    for (int64_t r0 = gw * NW_ROWS; r0 < g.K; r0 += nw * NW_ROWS) {
        float a0[NW_ROWS], a1[NW_ROWS];
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < NW_ROWS; ++rr) { my[rr][lane] = pa[rr]; my[rr][lane + 32] = pb[rr]; a0[rr] = qa[rr]; a1[rr] = qb[rr]; }
        __syncwarp();
        if (r0 + nw * NW_ROWS < g.K) fetch(r0 + nw * NW_ROWS);
#pragma unroll
        for (int nc = 0; nc < 16; ++nc) {
            if (nc < nc_n) {
#pragma unroll
                for (int rr = 0; rr < NW_ROWS; ++rr) {
                    const float4 x = *reinterpret_cast<const float4*>(&my[rr][nc * 4]);
                    c0[nc * 4 + 0] = fmaf(a0[rr], x.x, c0[nc * 4 + 0]); c1[nc * 4 + 0] = fmaf(a1[rr], x.x, c1[nc * 4 + 0]);
                    c0[nc * 4 + 1] = fmaf(a0[rr], x.y, c0[nc * 4 + 1]); c1[nc * 4 + 1] = fmaf(a1[rr], x.y, c1[nc * 4 + 1]);
                    c0[nc * 4 + 2] = fmaf(a0[rr], x.z, c0[nc * 4 + 2]); c1[nc * 4 + 2] = fmaf(a1[rr], x.z, c1[nc * 4 + 2]);
                    c0[nc * 4 + 3] = fmaf(a0[rr], x.w, c0[nc * 4 + 3]); c1[nc * 4 + 3] = fmaf(a1[rr], x.w, c1[nc * 4 + 3]);
                }
            }
        }
    }
    // CTA-level reduction of the 8 per-warp partials in shared memory (column-major so lanes hit distinct banks),
    // then ONE coalesced set of 64 x 64 red.adds per CTA instead of one per warp: the per-address serialisation of
    // same-address atomics in L2 was the whole cost of this kernel
    static_assert(sizeof(slab_raw) >= 64 * 65 * sizeof(float), "slab doubles as the 64 x 65 reduction buffer");
    float* cs = slab_raw;                                          // cs[n * 65 + m]: conflict-free both ways
    __syncthreads();                                               // every warp is done with its staging rows
    for (int w = 0; w < NW_THREADS / 32; ++w) {
        if (warp == w) {
#pragma unroll
            for (int n = 0; n < 64; ++n) {
                if (w == 0) { cs[n * 65 + m0] = c0[n]; cs[n * 65 + m1] = c1[n]; }
                else { cs[n * 65 + m0] += c0[n]; cs[n * 65 + m1] += c1[n]; }
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 64 * 64; i += NW_THREADS) {
        const int m = i >> 6, n = i & 63;                          // consecutive threads -> consecutive n: coalesced
        if (m < M && n < N) atomicAdd(g.C + (int64_t)m * g.ldc + n, cs[n * 65 + m]);
    }
}

int launch_narrow(GemmArgs& g, cudaStream_t st) {
    // returns 1 when the shape is not handled here
    const bool rows_mode = !g.transA && g.N <= 64 && g.K <= 64 && g.M >= 2048;
    const bool red_mode = g.transA && !g.transB && g.M <= 64 && g.N <= 64 && g.K >= 2048 && g.act == GAOT_ACT_NONE && !g.pre &&
                          !g.bias && !g.add;
    if (!rows_mode && !red_mode) return 1;
    const int warps = NW_THREADS / 32;
    if (rows_mode) {
        int64_t ctas = (g.M + warps * NW_ROWS - 1) / (warps * NW_ROWS);
        if (ctas > gaot_sm_count() * 4) ctas = gaot_sm_count() * 4;
        if (g.transB) narrow_rows_kernel<true><<<(unsigned)ctas, NW_THREADS, 0, st>>>(g);
        else narrow_rows_kernel<false><<<(unsigned)ctas, NW_THREADS, 0, st>>>(g);
    } else {
        if (!g.accumulate) {
            if (g.ldc == g.N) GAOT_CHECK(cudaMemsetAsync(g.C, 0, sizeof(float) * (size_t)(g.M * g.N), st));
            else GAOT_CHECK(cudaMemset2DAsync(g.C, sizeof(float) * (size_t)g.ldc, 0, sizeof(float) * (size_t)g.N, (size_t)g.M, st));
        }
        int64_t ctas = (g.K + warps * NW_ROWS * 2 - 1) / (warps * NW_ROWS * 2);    // >= 2 row groups per warp before the atomics
        if (ctas > gaot_sm_count()) ctas = gaot_sm_count();
        if (ctas < 1) ctas = 1;
        narrow_reduce_kernel<<<(unsigned)ctas, NW_THREADS, 0, st>>>(g);
    }
    gaot_launch_counter_add(1);
    return (int)cudaGetLastError();
}

// ---- small pointwise / reduction kernels ----------------------------------------------------------
__global__ void col_sum_kernel(const float* __restrict__ x, int64_t M, int64_t N, float* __restrict__ out, int64_t rows_per) {
    // grid (ceil(N/32), row-slices); block 32 x 8: threadIdx.x -> column (coalesced), threadIdx.y -> row lane
    __shared__ float red[8][33];
    const int64_t n = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
    float s = 0.f;
    if (n < N)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += __ldg(x + r * N + n);
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
        atomicAdd(out + n, t);
    }
}

__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ pre, float* __restrict__ dx,
                               int64_t n, int act) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float p = pre[i];
        const float gr = (act == GAOT_ACT_GELU) ? gelu_erf_grad(p) : ((act == GAOT_ACT_RELU) ? (p > 0.f ? 1.f : 0.f) : 1.f);
        dx[i] = dy[i] * gr;
    }
}

// dx (bf16) = dy * act'(pre); dy fp32 or bf16 - the gradient operand of the tensor-core dX / dW GEMMs
__global__ void act_bwd_bf16_kernel(const void* __restrict__ dy, int dy_bf16, const float* __restrict__ pre,
                                    __nv_bfloat16* __restrict__ dx, int64_t n, int act) {
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(pre) + i);
        float4 gy;
        if (dy_bf16) {
            const uint2 pk = __ldg(reinterpret_cast<const uint2*>(dy) + i);
            gy = make_float4(__uint_as_float(pk.x << 16), __uint_as_float(pk.x & 0xffff0000u), __uint_as_float(pk.y << 16),
                             __uint_as_float(pk.y & 0xffff0000u));
        } else gy = __ldg(reinterpret_cast<const float4*>(dy) + i);
        float gr[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) gr[k] = (act == GAOT_ACT_GELU) ? gelu_erf_grad(gr[k]) : ((act == GAOT_ACT_RELU) ? (gr[k] > 0.f ? 1.f : 0.f) : 1.f);
        __nv_bfloat162 t0 = __floats2bfloat162_rn(gy.x * gr[0], gy.y * gr[1]), t1 = __floats2bfloat162_rn(gy.z * gr[2], gy.w * gr[3]);
        uint2 o; o.x = *reinterpret_cast<uint32_t*>(&t0); o.y = *reinterpret_cast<uint32_t*>(&t1);
        reinterpret_cast<uint2*>(dx)[i] = o;
    }
}

constexpr int NORM_MAXE = 64;   // scalar fallback: per-lane elements, D <= 2048

// Vectorised RMSNorm (D % 4 == 0, D <= 128*KMAX): a warp owns a row, lane l holds the float4 chunks
// l, l+32, ... so every load/store is a coalesced 512-byte warp access; the row stays in registers.
template <int KMAX>
__global__ void __launch_bounds__(256) rmsnorm_fwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                              float* __restrict__ y, __nv_bfloat16* __restrict__ y16,
                                                              float* __restrict__ rstd, int64_t rows, int D, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nv = D >> 2;
    float4 gm[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) gm[k] = (lane + 32 * k < nv) ? __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = warp; r < rows; r += nwarps) {
        const float4* xp = reinterpret_cast<const float4*>(x + r * D);
        float4 xv[KMAX];
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            xv[k] = (lane + 32 * k < nv) ? __ldg(xp + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
            ss += xv[k].x * xv[k].x + xv[k].y * xv[k].y + xv[k].z * xv[k].z + xv[k].w * xv[k].w;
        }
        ss = warp_sum(ss);
        const float rs = rsqrtf(ss / (float)D + eps);
        if (lane == 0) rstd[r] = rs;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            const int c4 = lane + 32 * k;
            if (c4 < nv) {
                const float4 o = make_float4(xv[k].x * rs * gm[k].x, xv[k].y * rs * gm[k].y, xv[k].z * rs * gm[k].z, xv[k].w * rs * gm[k].w);
                if (y) reinterpret_cast<float4*>(y + r * D)[c4] = o;
                if (y16) {
                    __nv_bfloat162 t0 = __floats2bfloat162_rn(o.x, o.y), t1 = __floats2bfloat162_rn(o.z, o.w);
                    uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                    reinterpret_cast<uint2*>(y16 + r * D)[c4] = pk;
                }
            }
        }
    }
}

// (register cap: without it ptxas spends 105 registers per thread at KMAX = 2 on address arithmetic it keeps live across the row
// loop, two CTAs per SM fit, and the kernel runs at 3.3 TB/s on the latency of its one row per warp in flight)
template <int KMAX>
__global__ void __launch_bounds__(256, KMAX <= 2 ? 4 : (KMAX <= 4 ? 2 : 1)) rmsnorm_bwd_vec_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ dy16,
                                                              const float* __restrict__ x, const float* __restrict__ gamma,
                                                              const float* __restrict__ rstd, float* __restrict__ dx,
                                                              float* __restrict__ dgamma, int64_t rows, int D) {
    extern __shared__ float sdg[];                 // [D] per-CTA partial of dgamma
    for (int c = threadIdx.x; c < D; c += blockDim.x) sdg[c] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nv = D >> 2;
    float4 gm[KMAX], dg[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        gm[k] = (lane + 32 * k < nv) ? __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
        dg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int64_t r = warp; r < rows; r += nwarps) {
        float4 xv[KMAX], gv[KMAX];
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            const int c4 = lane + 32 * k;
            xv[k] = gv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c4 < nv) {
                xv[k] = __ldg(reinterpret_cast<const float4*>(x + r * D) + c4);
                if (dy) gv[k] = __ldg(reinterpret_cast<const float4*>(dy + r * D) + c4);
                if (dy16) {     // the two consumers of y (fp32 residual path, bf16 GEMM path) add up
                    const uint2 pk = __ldg(reinterpret_cast<const uint2*>(dy16 + r * D) + c4);
                    gv[k].x += __uint_as_float(pk.x << 16); gv[k].y += __uint_as_float(pk.x & 0xffff0000u);
                    gv[k].z += __uint_as_float(pk.y << 16); gv[k].w += __uint_as_float(pk.y & 0xffff0000u);
                }
            }
            dot += gv[k].x * gm[k].x * xv[k].x + gv[k].y * gm[k].y * xv[k].y + gv[k].z * gm[k].z * xv[k].z + gv[k].w * gm[k].w * xv[k].w;
        }
        dot = warp_sum(dot);
        const float rs = __ldg(rstd + r);
        const float coef = dot * rs * rs * rs / (float)D;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            const int c4 = lane + 32 * k;
            if (c4 < nv) {
                reinterpret_cast<float4*>(dx + r * D)[c4] =
                    make_float4(gv[k].x * gm[k].x * rs - xv[k].x * coef, gv[k].y * gm[k].y * rs - xv[k].y * coef,
                                gv[k].z * gm[k].z * rs - xv[k].z * coef, gv[k].w * gm[k].w * rs - xv[k].w * coef);
                dg[k].x = fmaf(gv[k].x, xv[k].x * rs, dg[k].x); dg[k].y = fmaf(gv[k].y, xv[k].y * rs, dg[k].y);
                dg[k].z = fmaf(gv[k].z, xv[k].z * rs, dg[k].z); dg[k].w = fmaf(gv[k].w, xv[k].w * rs, dg[k].w);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const int c = (lane + 32 * k) * 4;
        if (c < D) {
            atomicAdd(&sdg[c], dg[k].x); atomicAdd(&sdg[c + 1], dg[k].y); atomicAdd(&sdg[c + 2], dg[k].z); atomicAdd(&sdg[c + 3], dg[k].w);
        }
    }
    __syncthreads();
    if ((reinterpret_cast<uintptr_t>(dgamma) & 15) == 0) {
        for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4)      // D % 4 == 0 on this path
            atomicAdd(reinterpret_cast<float4*>(dgamma + c), make_float4(sdg[c], sdg[c + 1], sdg[c + 2], sdg[c + 3]));
    } else {
        for (int c = threadIdx.x; c < D; c += blockDim.x) atomicAdd(dgamma + c, sdg[c]);
    }
}


__global__ void __launch_bounds__(256) rmsnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                          float* __restrict__ y, __nv_bfloat16* __restrict__ y16,
                                                          float* __restrict__ rstd, int64_t rows, int D, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const float* xp = x + r * D;
        float ss = 0.f;
        for (int c = lane; c < D; c += 32) { const float v = xp[c]; ss = fmaf(v, v, ss); }
        ss = warp_sum(ss);
        const float rs = rsqrtf(ss / (float)D + eps);
        if (lane == 0) rstd[r] = rs;
        for (int c = lane; c < D; c += 32) {
            const float v = (xp[c] * rs) * __ldg(gamma + c);
            if (y) y[r * D + c] = v;
            if (y16) y16[r * D + c] = __float2bfloat16_rn(v);
        }
    }
}

__global__ void __launch_bounds__(256) rmsnorm_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ dy16,
                                                          const float* __restrict__ x,
                                                          const float* __restrict__ gamma, const float* __restrict__ rstd,
                                                          float* __restrict__ dx, float* __restrict__ dgamma,
                                                          int64_t rows, int D) {
    extern __shared__ float sdg[];                 // [D] per-CTA partial of dgamma
    for (int c = threadIdx.x; c < D; c += blockDim.x) sdg[c] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float dg[NORM_MAXE];
#pragma unroll
    for (int k = 0; k < NORM_MAXE; ++k) dg[k] = 0.f;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const float* xp = x + r * D;
        const float rs = rstd[r];
        auto grad_at = [&](int c) -> float {        // the two consumers of y (fp32 residual path, bf16 GEMM path) add up
            float v = dy ? dy[r * D + c] : 0.f;
            if (dy16) v += __bfloat162float(dy16[r * D + c]);
            return v;
        };
        float dot = 0.f;
        for (int c = lane; c < D; c += 32) dot = fmaf(grad_at(c) * __ldg(gamma + c), xp[c], dot);
        dot = warp_sum(dot);
        const float coef = dot * rs * rs * rs / (float)D;
        float* dxp = dx + r * D;
#pragma unroll
        for (int k = 0; k < NORM_MAXE; ++k) {
            const int c = lane + k * 32;
            if (c < D) {
                const float gy = grad_at(c), xv = xp[c];
                dxp[c] = gy * __ldg(gamma + c) * rs - xv * coef;
                dg[k] = fmaf(gy, xv * rs, dg[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NORM_MAXE; ++k) {
        const int c = lane + k * 32;
        if (c < D) atomicAdd(&sdg[c], dg[k]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) atomicAdd(dgamma + c, sdg[c]);
}

__global__ void swiglu_fwd_kernel(const float* __restrict__ ab, float* __restrict__ u, int64_t rows, int64_t F) {
    const int64_t total = rows * F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / F, c = i - r * F;
        const float a = ab[r * 2 * F + c], b = ab[r * 2 * F + F + c];
        u[i] = (a / (1.f + expf(-a))) * b;
    }
}
__global__ void swiglu_bwd_kernel(const float* __restrict__ du, const float* __restrict__ ab, float* __restrict__ dab,
                                  int64_t rows, int64_t F) {
    const int64_t total = rows * F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / F, c = i - r * F;
        const float a = ab[r * 2 * F + c], b = ab[r * 2 * F + F + c], gu = du[i];
        const float sg = 1.f / (1.f + expf(-a));
        dab[r * 2 * F + c] = gu * b * (sg + a * sg * (1.f - sg));
        dab[r * 2 * F + F + c] = gu * a * sg;
    }
}

__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t bf_pack(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
// bf16 SwiGLU: 8 elements (16 bytes) per thread; F % 8 == 0
__global__ void swiglu_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ ab, __nv_bfloat16* __restrict__ u, int64_t rows, int64_t F) {
    const int64_t f8 = F >> 3, total = rows * f8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / f8, c = (i - r * f8) << 3;
        const uint4 av = __ldg(reinterpret_cast<const uint4*>(ab + r * 2 * F + c));
        const uint4 bv = __ldg(reinterpret_cast<const uint4*>(ab + r * 2 * F + F + c));
        const uint32_t aw[4] = {av.x, av.y, av.z, av.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a0 = bf_lo(aw[k]), a1 = bf_hi(aw[k]);
            o[k] = bf_pack(a0 / (1.f + __expf(-a0)) * bf_lo(bw[k]), a1 / (1.f + __expf(-a1)) * bf_hi(bw[k]));
        }
        *reinterpret_cast<uint4*>(u + r * F + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}
__global__ void swiglu_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ du, const __nv_bfloat16* __restrict__ ab,
                                       __nv_bfloat16* __restrict__ dab, int64_t rows, int64_t F) {
    const int64_t f8 = F >> 3, total = rows * f8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / f8, c = (i - r * f8) << 3;
        const uint4 av = __ldg(reinterpret_cast<const uint4*>(ab + r * 2 * F + c));
        const uint4 bv = __ldg(reinterpret_cast<const uint4*>(ab + r * 2 * F + F + c));
        const uint4 gv = __ldg(reinterpret_cast<const uint4*>(du + r * F + c));
        const uint32_t aw[4] = {av.x, av.y, av.z, av.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
        uint32_t oa[4], ob[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float da[2], db[2];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const float a = hh ? bf_hi(aw[k]) : bf_lo(aw[k]), b = hh ? bf_hi(bw[k]) : bf_lo(bw[k]);
                const float gu = hh ? bf_hi(gw[k]) : bf_lo(gw[k]);
                const float sg = 1.f / (1.f + __expf(-a));
                da[hh] = gu * b * (sg + a * sg * (1.f - sg));
                db[hh] = gu * a * sg;
            }
            oa[k] = bf_pack(da[0], da[1]); ob[k] = bf_pack(db[0], db[1]);
        }
        *reinterpret_cast<uint4*>(dab + r * 2 * F + c) = make_uint4(oa[0], oa[1], oa[2], oa[3]);
        *reinterpret_cast<uint4*>(dab + r * 2 * F + F + c) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
    }
}

// IDX = uint32_t whenever the element count fits (it does for every BASELINE config): the index decomposition is a chain of
// divisions / remainders, and 64-bit integer division costs ~5x the 32-bit one - this copy kernel was bound by it, not by HBM
template <typename T, typename IDX>
__global__ void patch_permute_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t B_, int dims,
                                     int64_t H_, int64_t W_, int64_t D_, int64_t P_, int64_t C_, int to_tokens) {
    // iterate over token-layout elements (T = float, or float4 with C counted in vectors); channel fastest ->
    // both sides move contiguous C-runs
    const IDX B = (IDX)B_, H = (IDX)H_, W = (IDX)W_, D = (IDX)D_, P = (IDX)P_, C = (IDX)C_;
    const IDX nodes = H * W * (dims == 3 ? D : 1);
    const IDX total = B * nodes * C;
    const IDX pw = W / P, pd = (dims == 3) ? D / P : 1;
    const IDX pv = (dims == 3) ? P * P * P : P * P;
    const IDX ntok = nodes / pv;
    for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IDX)gridDim.x * blockDim.x) {
        const IDX c = i % C;
        IDX r = i / C;
        const IDX pin = r % pv; r /= pv;             // position inside the patch
        const IDX t = r % ntok;                        // token
        const IDX b = r / ntok;
        IDX node;
        if (dims == 2) {
            const IDX ti = t / pw, tj = t % pw, pi = pin / P, pj = pin % P;
            node = (ti * P + pi) * W + (tj * P + pj);
        } else {
            const IDX tk = t % pd, tj = (t / pd) % pw, ti = t / (pd * pw);
            const IDX pk = pin % P, pj = (pin / P) % P, pi = pin / (P * P);
            node = ((ti * P + pi) * W + (tj * P + pj)) * D + (tk * P + pk);
        }
        const IDX g = (b * nodes + node) * C + c;
        if (to_tokens) out[i] = in[g]; else out[g] = in[i];
    }
}

__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                  float* __restrict__ loss, float* __restrict__ dpred, int64_t n,
                                                  float inv_n, float gcoef) {
    __shared__ float red[8];
    float s = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float df = pred[i] - target[i];
        s = fmaf(df, df, s);
        if (dpred) dpred[i] = df * gcoef;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) atomicAdd(loss, t * inv_n);
    }
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                             float inv_bc1, float inv_sqrt_bc2, float gscale) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gr = g[i] * gscale;
        float pv = p[i] * (1.f - lr * wd);
        const float mv = b1 * m[i] + (1.f - b1) * gr;
        const float vv = b2 * v[i] + (1.f - b2) * gr * gr;
        m[i] = mv; v[i] = vv;
        const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
        pv -= (lr * inv_bc1) * (mv / denom);
        p[i] = pv;
    }
}

// Rollout update (gaot.py:436-476 + 431-434): u_phys = the de-normalised prediction of the step (what the reference returns),
// u_next = (u_phys - u_mean) / u_std, the normalised state carried into the next step.
__global__ void rollout_kernel(const float* __restrict__ pred, const float* __restrict__ u_in, float* __restrict__ u_phys,
                               float* __restrict__ u_next, int64_t rows, int C, float dt, int mode, const float* __restrict__ um,
                               const float* __restrict__ us, const float* __restrict__ sm, const float* __restrict__ ss) {
    const int64_t total = rows * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        float un;
        if (mode == 0) {
            un = pred[i] * us[c] + um[c];
        } else {
            const float ucur = u_in[i] * us[c] + um[c];
            const float inc = pred[i] * ss[c] + sm[c];
            un = ucur + (mode == 2 ? dt * inc : inc);
        }
        u_phys[i] = un;
        if (u_next) u_next[i] = (un - um[c]) / us[c];
    }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

GAOT_API int gaot_abi_version(void) { return 1; }

static long long g_launches = 0;
extern "C" __attribute__((visibility("default"))) long long gaot_launch_counter_add(long long n) {
    return __atomic_add_fetch(&g_launches, n, __ATOMIC_RELAXED);
}
GAOT_API long long gaot_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

GAOT_API int gaot_gemm_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                           int64_t ldb, int64_t ldc, int transA, int transB, const float* bias, const float* add,
                           int64_t add_rows, int64_t ldadd, int act, float* pre, int accumulate, void* stream) {
    if (!A || !B || !C || M < 0 || N < 0 || K < 0) return GAOT_E_BADARG;
    if (add && (add_rows < 1 || ldadd < N)) return GAOT_E_BADARG;
    if (act < 0 || act > 2 || ldc < N) return GAOT_E_BADARG;
    if (M == 0 || N == 0) return 0;
    GemmArgs g{A, B, C, M, N, K, lda, ldb, ldc, transA, transB, bias, add, add ? add_rows : 1, ldadd, act, pre,
               accumulate, 1, 0, 0};
    g.vecA = al16(A) && (lda % 4 == 0);
    g.vecB = al16(B) && (ldb % 4 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    {
        const int rc = launch_narrow(g, st);              // both small dimensions <= 64: weights-in-registers kernels
        if (rc != 1) return rc;
    }
    if (N <= 16) return launch_gemm<256, 16, 8, 2>(g, st);
    if (N <= 64) return launch_gemm<128, 64, 8, 4>(g, st);
    return launch_gemm<128, 128, 8, 8>(g, st);
}

GAOT_API int gaot_col_sum(const float* x, int64_t M, int64_t N, float* out, int accumulate, void* stream) {
    if (!x || !out || M < 0 || N < 1) return GAOT_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) GAOT_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N, st));
    if (M == 0) return 0;
    const int64_t gx = (N + 31) / 32;
    int64_t slices = (gaot_sm_count() * 8 + gx - 1) / gx;
    if (slices > (M + 63) / 64) slices = (M + 63) / 64;
    if (slices < 1) slices = 1;
    const int64_t rows_per = (M + slices - 1) / slices;
    col_sum_kernel<<<dim3((unsigned)gx, (unsigned)slices), dim3(32, 8), 0, st>>>(x, M, N, out, rows_per);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_act_bwd(const float* dy, const float* pre, float* dx, int64_t n, int act, void* stream) {
    if (!dy || !pre || !dx || n < 0 || act < 0 || act > 2) return GAOT_E_BADARG;
    if (n == 0) return 0;
    act_bwd_kernel<<<gaot_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, pre, dx, n, act);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_act_bwd_bf16(const void* dy, int dy_is_bf16, const float* pre, void* dx, int64_t n, int act, void* stream) {
    if (!dy || !pre || !dx || n < 0 || n % 4 || act < 0 || act > 2) return GAOT_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(dy) & 15) || (reinterpret_cast<uintptr_t>(pre) & 15) || (reinterpret_cast<uintptr_t>(dx) & 7)) return GAOT_E_ALIGN;
    if (n == 0) return 0;
    act_bwd_bf16_kernel<<<gaot_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(dy, dy_is_bf16, pre, reinterpret_cast<__nv_bfloat16*>(dx), n, act);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_rmsnorm_fwd(const float* x, const float* gamma, float* y, void* y_bf16, float* rstd, int64_t rows,
                              int64_t D, float eps, void* stream) {
    if (!x || !gamma || (!y && !y_bf16) || !rstd || rows < 0 || D < 1) return GAOT_E_BADARG;
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16* y16 = reinterpret_cast<__nv_bfloat16*>(y_bf16);
    const bool vec = D % 4 == 0 && D <= 1024 && !(reinterpret_cast<uintptr_t>(x) & 15) && !(reinterpret_cast<uintptr_t>(gamma) & 15) &&
                     !(reinterpret_cast<uintptr_t>(y) & 15) && !(reinterpret_cast<uintptr_t>(y_bf16) & 7);
    const int grid = gaot_grid(rows * 32, 256);
    if (vec && D <= 128) rmsnorm_fwd_vec_kernel<1><<<grid, 256, 0, st>>>(x, gamma, y, y16, rstd, rows, (int)D, eps);
    else if (vec && D <= 256) rmsnorm_fwd_vec_kernel<2><<<grid, 256, 0, st>>>(x, gamma, y, y16, rstd, rows, (int)D, eps);
    else if (vec && D <= 512) rmsnorm_fwd_vec_kernel<4><<<grid, 256, 0, st>>>(x, gamma, y, y16, rstd, rows, (int)D, eps);
    else if (vec) rmsnorm_fwd_vec_kernel<8><<<grid, 256, 0, st>>>(x, gamma, y, y16, rstd, rows, (int)D, eps);
    else rmsnorm_fwd_kernel<<<grid, 256, 0, st>>>(x, gamma, y, y16, rstd, rows, (int)D, eps);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_rmsnorm_bwd(const float* dy, const void* dy_bf16, const float* x, const float* gamma, const float* rstd,
                              float* dx, float* dgamma, int64_t rows, int64_t D, void* stream) {
    if ((!dy && !dy_bf16) || !x || !gamma || !rstd || !dx || !dgamma || rows < 0 || D < 1 || D > 32 * NORM_MAXE) return GAOT_E_BADARG;
    if (rows == 0) return 0;
    const int grid = gaot_grid(rows * 32, 256, 4);
    cudaStream_t st = (cudaStream_t)stream;
    const __nv_bfloat16* g16 = reinterpret_cast<const __nv_bfloat16*>(dy_bf16);
    const size_t sm = sizeof(float) * (size_t)D;
    const bool vec = D % 4 == 0 && D <= 1024 && !(reinterpret_cast<uintptr_t>(x) & 15) && !(reinterpret_cast<uintptr_t>(gamma) & 15) &&
                     !(reinterpret_cast<uintptr_t>(dy) & 15) && !(reinterpret_cast<uintptr_t>(dy_bf16) & 7) &&
                     !(reinterpret_cast<uintptr_t>(dx) & 15);
    if (vec && D <= 128) rmsnorm_bwd_vec_kernel<1><<<grid, 256, sm, st>>>(dy, g16, x, gamma, rstd, dx, dgamma, rows, (int)D);
    else if (vec && D <= 256) rmsnorm_bwd_vec_kernel<2><<<grid, 256, sm, st>>>(dy, g16, x, gamma, rstd, dx, dgamma, rows, (int)D);
    else if (vec && D <= 512) rmsnorm_bwd_vec_kernel<4><<<grid, 256, sm, st>>>(dy, g16, x, gamma, rstd, dx, dgamma, rows, (int)D);
    else if (vec) rmsnorm_bwd_vec_kernel<8><<<grid, 256, sm, st>>>(dy, g16, x, gamma, rstd, dx, dgamma, rows, (int)D);
    else rmsnorm_bwd_kernel<<<grid, 256, sm, st>>>(dy, g16, x, gamma, rstd, dx, dgamma, rows, (int)D);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_swiglu_fwd(const float* ab, float* u, int64_t rows, int64_t F, void* stream) {
    if (!ab || !u || rows < 0 || F < 1) return GAOT_E_BADARG;
    if (rows == 0) return 0;
    swiglu_fwd_kernel<<<gaot_grid(rows * F, 256), 256, 0, (cudaStream_t)stream>>>(ab, u, rows, F);
    GAOT_RETURN_LAST();
}
GAOT_API int gaot_swiglu_bwd(const float* du, const float* ab, float* dab, int64_t rows, int64_t F, void* stream) {
    if (!du || !ab || !dab || rows < 0 || F < 1) return GAOT_E_BADARG;
    if (rows == 0) return 0;
    swiglu_bwd_kernel<<<gaot_grid(rows * F, 256), 256, 0, (cudaStream_t)stream>>>(du, ab, dab, rows, F);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_swiglu_fwd_bf16(const void* ab, void* u, int64_t rows, int64_t F, void* stream) {
    if (!ab || !u || rows < 0 || F < 8 || F % 8) return GAOT_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(ab) & 15) || (reinterpret_cast<uintptr_t>(u) & 15)) return GAOT_E_ALIGN;
    if (rows == 0) return 0;
    swiglu_fwd_bf16_kernel<<<gaot_grid(rows * F / 8, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(ab), reinterpret_cast<__nv_bfloat16*>(u), rows, F);
    GAOT_RETURN_LAST();
}
GAOT_API int gaot_swiglu_bwd_bf16(const void* du, const void* ab, void* dab, int64_t rows, int64_t F, void* stream) {
    if (!du || !ab || !dab || rows < 0 || F < 8 || F % 8) return GAOT_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(ab) & 15) || (reinterpret_cast<uintptr_t>(du) & 15) || (reinterpret_cast<uintptr_t>(dab) & 15))
        return GAOT_E_ALIGN;
    if (rows == 0) return 0;
    swiglu_bwd_bf16_kernel<<<gaot_grid(rows * F / 8, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(du), reinterpret_cast<const __nv_bfloat16*>(ab), reinterpret_cast<__nv_bfloat16*>(dab),
        rows, F);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_patch_permute(const float* in, float* out, int64_t B, int dims, int64_t H, int64_t W, int64_t D,
                                int64_t P, int64_t C, int to_tokens, void* stream) {
    if (!in || !out || B < 1 || (dims != 2 && dims != 3) || P < 1 || C < 1) return GAOT_E_BADARG;
    if (H % P || W % P || (dims == 3 && D % P)) return GAOT_E_BADARG;
    const int64_t total = B * H * W * (dims == 3 ? D : 1) * C;
    const bool small = total < 0x7fffffffLL;           // 32-bit index arithmetic (grid-stride loop counter included)
    if (C % 4 == 0 && !(reinterpret_cast<uintptr_t>(in) & 15) && !(reinterpret_cast<uintptr_t>(out) & 15)) {
        if (small)
            patch_permute_kernel<float4, uint32_t><<<gaot_grid(total / 4, 256), 256, 0, (cudaStream_t)stream>>>(
                reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), B, dims, H, W, D, P, C / 4, to_tokens);
        else
            patch_permute_kernel<float4, int64_t><<<gaot_grid(total / 4, 256), 256, 0, (cudaStream_t)stream>>>(
                reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), B, dims, H, W, D, P, C / 4, to_tokens);
    } else if (small) {
        patch_permute_kernel<float, uint32_t><<<gaot_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, dims, H, W, D, P, C, to_tokens);
    } else {
        patch_permute_kernel<float, int64_t><<<gaot_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, dims, H, W, D, P, C, to_tokens);
    }
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_mse_fwd_bwd(const float* pred, const float* target, float* loss, float* dpred, int64_t n, float gscale,
                              void* stream) {
    if (!pred || !target || !loss || n < 1) return GAOT_E_BADARG;
    const float inv_n = (float)(1.0 / (double)n);
    mse_kernel<<<gaot_grid(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(pred, target, loss, dpred, n, inv_n,
                                                                       (float)(2.0 / (double)n) * gscale);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                             void* stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq || n < 0 || step < 1) return GAOT_E_BADARG;
    if (n == 0) return 0;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    adamw_kernel<<<gaot_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
                                                                      eps, weight_decay, (float)(1.0 / bc1),
                                                                      (float)(1.0 / sqrt(bc2)), grad_scale);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_rollout_update(const float* pred, const float* u_in, float* u_phys, float* u_next, int64_t rows, int64_t C,
                                 float dt, int mode, const float* u_mean, const float* u_std, const float* s_mean,
                                 const float* s_std, void* stream) {
    if (!pred || !u_phys || rows < 0 || C < 1 || mode < 0 || mode > 2 || !u_mean || !u_std) return GAOT_E_BADARG;
    if (mode != 0 && (!u_in || !s_mean || !s_std)) return GAOT_E_BADARG;
    if (rows == 0) return 0;
    rollout_kernel<<<gaot_grid(rows * C, 256), 256, 0, (cudaStream_t)stream>>>(pred, u_in, u_phys, u_next, rows, (int)C, dt, mode,
                                                                               u_mean, u_std, s_mean, s_std);
    GAOT_RETURN_LAST();
}

namespace {

// ---- time-conditioned affine correction (ConditionedNorm.forward, mlp.py:121-124): y[b,t,:] = x[b,t,:] * s[b,:] + o[b,:] ----
__global__ void cond_affine_fwd_kernel(const float* __restrict__ x, const float* __restrict__ s, const float* __restrict__ o,
                                       float* __restrict__ y, int64_t B, int64_t T, int D) {
    const int64_t total = B * T * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / (T * D);
        const int d = (int)(i % D);
        y[i] = fmaf(x[i], __ldg(s + b * D + d), __ldg(o + b * D + d));
    }
}
// backward: dx = dy * s; ds[b,d] = sum_t dy x; do[b,d] = sum_t dy.  Block (32 columns x 8 row lanes) owns (sample, 32 columns)
__global__ void __launch_bounds__(256) cond_affine_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                              const float* __restrict__ s, float* __restrict__ dx, float* __restrict__ ds,
                                                              float* __restrict__ dof, int64_t T, int D) {
    __shared__ float rs[8][33], ro[8][33];
    const int d = blockIdx.x * 32 + threadIdx.x;
    const int64_t b = blockIdx.y;
    float as = 0.f, ao = 0.f;
    if (d < D) {
        const float sv = __ldg(s + b * D + d);
        for (int64_t t = threadIdx.y; t < T; t += 8) {
            const int64_t i = (b * T + t) * D + d;
            const float g = dy[i];
            dx[i] = g * sv;
            as = fmaf(g, x[i], as);
            ao += g;
        }
    }
    rs[threadIdx.y][threadIdx.x] = as; ro[threadIdx.y][threadIdx.x] = ao;
    __syncthreads();
    if (threadIdx.y == 0 && d < D) {
        float a = 0.f, c = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { a += rs[k][threadIdx.x]; c += ro[k][threadIdx.x]; }
        ds[b * D + d] = a; dof[b * D + d] = c;
    }
}
}  // namespace

GAOT_API int gaot_cond_affine_fwd(const float* x, const float* scale, const float* bias, float* y, int64_t B, int64_t T, int64_t D,
                                  void* stream) {
    if (!x || !scale || !bias || !y || B < 0 || T < 0 || D < 1) return GAOT_E_BADARG;
    if (B * T == 0) return 0;
    cond_affine_fwd_kernel<<<gaot_grid(B * T * D, 256), 256, 0, (cudaStream_t)stream>>>(x, scale, bias, y, B, T, (int)D);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_cond_affine_bwd(const float* dy, const float* x, const float* scale, float* dx, float* dscale, float* dbias,
                                  int64_t B, int64_t T, int64_t D, void* stream) {
    if (!dy || !x || !scale || !dx || !dscale || !dbias || B < 0 || T < 0 || D < 1 || B > 65535) return GAOT_E_BADARG;
    if (B == 0) return 0;
    dim3 grid((unsigned)((D + 31) / 32), (unsigned)B), block(32, 8);
    cond_affine_bwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(dy, x, scale, dx, dscale, dbias, T, (int)D);
    GAOT_RETURN_LAST();
}
