// fp32 flash-style attention (parity path): non-causal softmax(Q K^T * scale) V with grouped KV
// heads, forward and backward, never materialising the T x T score matrix.
// Reference: GroupQueryFlashAttention.forward, src/model/layers/attn.py:96-116 (F.scaled_dot_product_attention).
//
// Mapping: one thread owns one query row (forward, dQ) or one key row (dK/dV); the other operand is
// staged in shared memory 64 rows at a time and read with 128-bit broadcast loads, so every LDS feeds
// 4 FFMAs of all 32 lanes.  Softmax statistics are per thread: no shuffles, no cross-thread traffic.
#include "common.cuh"
#include "../../include/gaot_b200.h"

namespace {

constexpr int ATT_THREADS = 128;
constexpr int ATT_TILE = 64;       // rows of the streamed operand per smem stage
constexpr int ATT_CHUNK = 8;       // keys per online-softmax update

template <int HD>
__device__ __forceinline__ void stage_rows(float (*dst)[HD], const float* __restrict__ src, int64_t ld, int64_t row0,
                                           int64_t nrows_total) {
    // ATT_TILE rows x HD floats, coalesced float4 loads; rows past the end are zero-filled
    constexpr int V = HD / 4;
    for (int s = threadIdx.x; s < ATT_TILE * V; s += ATT_THREADS) {
        const int r = s / V, c4 = (s % V) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < nrows_total) v = __ldg(reinterpret_cast<const float4*>(src + (row0 + r) * ld + c4));
        *reinterpret_cast<float4*>(&dst[r][c4]) = v;
    }
}

template <int HD>
__device__ __forceinline__ float dot_smem(const float (&a)[HD], const float* __restrict__ srow) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(srow + d);
        s = fmaf(a[d], kv.x, s); s = fmaf(a[d + 1], kv.y, s); s = fmaf(a[d + 2], kv.z, s); s = fmaf(a[d + 3], kv.w, s);
    }
    return s;
}
template <int HD>
__device__ __forceinline__ void axpy_smem(float (&acc)[HD], float a, const float* __restrict__ srow) {
#pragma unroll
    for (int d = 0; d < HD; d += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(srow + d);
        acc[d] = fmaf(a, vv.x, acc[d]); acc[d + 1] = fmaf(a, vv.y, acc[d + 1]);
        acc[d + 2] = fmaf(a, vv.z, acc[d + 2]); acc[d + 3] = fmaf(a, vv.w, acc[d + 3]);
    }
}

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                float* __restrict__ o, float* __restrict__ lse, int64_t T, int H, int Hkv, int64_t ldq, int64_t ldk,
                float scale) {
    __shared__ __align__(16) float Ks[ATT_TILE][HD];
    __shared__ __align__(16) float Vs[ATT_TILE][HD];
    const int h = blockIdx.y, hk = h / (H / Hkv);
    const int64_t b = blockIdx.z;
    const int64_t t = (int64_t)blockIdx.x * ATT_THREADS + threadIdx.x;
    const bool live = t < T;
    const float* kb = k + b * T * ldk + (int64_t)hk * HD;
    const float* vb = v + b * T * ldk + (int64_t)hk * HD;
    float qr[HD], acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) { qr[d] = 0.f; acc[d] = 0.f; }
    if (live) {
        const float* qp = q + (b * T + t) * ldq + (int64_t)h * HD;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(qp + d));
            qr[d] = x.x * scale; qr[d + 1] = x.y * scale; qr[d + 2] = x.z * scale; qr[d + 3] = x.w * scale;
        }
    }
    float mrun = -INFINITY, lrun = 0.f;
    for (int64_t k0 = 0; k0 < T; k0 += ATT_TILE) {
        __syncthreads();
        stage_rows<HD>(Ks, kb, ldk, k0, T);
        stage_rows<HD>(Vs, vb, ldk, k0, T);
        __syncthreads();
        const int nk = (int)min((int64_t)ATT_TILE, T - k0);
        for (int c0 = 0; c0 < nk; c0 += ATT_CHUNK) {
            float s[ATT_CHUNK];
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < ATT_CHUNK; ++c) {
                s[c] = (c0 + c < nk) ? dot_smem<HD>(qr, &Ks[c0 + c][0]) : -INFINITY;
                mx = fmaxf(mx, s[c]);
            }
            const float mnew = fmaxf(mrun, mx);
            const float corr = expf(mrun - mnew);
            lrun *= corr;
#pragma unroll
            for (int d = 0; d < HD; ++d) acc[d] *= corr;
#pragma unroll
            for (int c = 0; c < ATT_CHUNK; ++c) {
                const float p = expf(s[c] - mnew);
                lrun += p;
                axpy_smem<HD>(acc, p, &Vs[(c0 + c < nk) ? c0 + c : 0][0]);
            }
            mrun = mnew;
        }
    }
    if (live) {
        const float inv = 1.f / lrun;
        float* op = o + (b * T + t) * ((int64_t)H * HD) + (int64_t)h * HD;
#pragma unroll
        for (int d = 0; d < HD; d += 4)
            *reinterpret_cast<float4*>(op + d) = make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv);
        lse[(b * H + h) * T + t] = mrun + logf(lrun);
    }
}

// dQ and delta = rowsum(dO * O); one thread per query row
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dq_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                   const float* __restrict__ o, const float* __restrict__ lse, const float* __restrict__ dout,
                   float* __restrict__ delta, float* __restrict__ dq, int64_t T, int H, int Hkv, int64_t ldq,
                   int64_t ldk, float scale) {
    __shared__ __align__(16) float Ks[ATT_TILE][HD];
    __shared__ __align__(16) float Vs[ATT_TILE][HD];
    const int h = blockIdx.y, hk = h / (H / Hkv);
    const int64_t b = blockIdx.z;
    const int64_t t = (int64_t)blockIdx.x * ATT_THREADS + threadIdx.x;
    const bool live = t < T;
    const float* kb = k + b * T * ldk + (int64_t)hk * HD;
    const float* vb = v + b * T * ldk + (int64_t)hk * HD;
    float qr[HD], go[HD], acc[HD];
    float dl = 0.f, ls = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) { qr[d] = 0.f; go[d] = 0.f; acc[d] = 0.f; }
    if (live) {
        const float* qp = q + (b * T + t) * ldq + (int64_t)h * HD;
        const float* gp = dout + (b * T + t) * ((int64_t)H * HD) + (int64_t)h * HD;
        const float* op = o + (b * T + t) * ((int64_t)H * HD) + (int64_t)h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) { qr[d] = __ldg(qp + d) * scale; go[d] = __ldg(gp + d); dl = fmaf(go[d], __ldg(op + d), dl); }
        ls = lse[(b * H + h) * T + t];
        delta[(b * H + h) * T + t] = dl;
    }
    for (int64_t k0 = 0; k0 < T; k0 += ATT_TILE) {
        __syncthreads();
        stage_rows<HD>(Ks, kb, ldk, k0, T);
        stage_rows<HD>(Vs, vb, ldk, k0, T);
        __syncthreads();
        const int nk = (int)min((int64_t)ATT_TILE, T - k0);
        for (int c = 0; c < nk; ++c) {
            const float s = dot_smem<HD>(qr, &Ks[c][0]);
            const float p = expf(s - ls);
            const float dp = dot_smem<HD>(go, &Vs[c][0]);
            axpy_smem<HD>(acc, p * (dp - dl), &Ks[c][0]);
        }
    }
    if (live) {
        float* dp_ = dq + (b * T + t) * ldq + (int64_t)h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) dp_[d] = acc[d] * scale;
    }
}

// dK, dV; one thread per key row, loops over the query heads of its KV group
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dkv_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                    const float* __restrict__ lse, const float* __restrict__ dout, const float* __restrict__ delta,
                    float* __restrict__ dk, float* __restrict__ dv, int64_t T, int H, int Hkv, int64_t ldq, int64_t ldk,
                    float scale) {
    __shared__ __align__(16) float Qs[ATT_TILE][HD];
    __shared__ __align__(16) float Gs[ATT_TILE][HD];
    __shared__ float Ls[ATT_TILE], Ds[ATT_TILE];
    const int hk = blockIdx.y, rep = H / Hkv;
    const int64_t b = blockIdx.z;
    const int64_t t = (int64_t)blockIdx.x * ATT_THREADS + threadIdx.x;
    const bool live = t < T;
    float kr[HD], vr[HD], ak[HD], av[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) { kr[d] = 0.f; vr[d] = 0.f; ak[d] = 0.f; av[d] = 0.f; }
    if (live) {
        const float* kp = k + (b * T + t) * ldk + (int64_t)hk * HD;
        const float* vp = v + (b * T + t) * ldk + (int64_t)hk * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) { kr[d] = __ldg(kp + d) * scale; vr[d] = __ldg(vp + d); }
    }
    for (int r = 0; r < rep; ++r) {
        const int h = hk * rep + r;
        const float* qb = q + b * T * ldq + (int64_t)h * HD;
        const float* gb = dout + b * T * ((int64_t)H * HD) + (int64_t)h * HD;
        for (int64_t q0 = 0; q0 < T; q0 += ATT_TILE) {
            __syncthreads();
            stage_rows<HD>(Qs, qb, ldq, q0, T);
            stage_rows<HD>(Gs, gb, (int64_t)H * HD, q0, T);
            if (threadIdx.x < ATT_TILE) {
                const bool ok = q0 + threadIdx.x < T;
                Ls[threadIdx.x] = ok ? lse[(b * H + h) * T + q0 + threadIdx.x] : INFINITY;   // exp(-inf) = 0
                Ds[threadIdx.x] = ok ? delta[(b * H + h) * T + q0 + threadIdx.x] : 0.f;
            }
            __syncthreads();
            const int nq = (int)min((int64_t)ATT_TILE, T - q0);
            for (int c = 0; c < nq; ++c) {
                const float s = dot_smem<HD>(kr, &Qs[c][0]);          // (k*scale) . q
                const float p = expf(s - Ls[c]);
                axpy_smem<HD>(av, p, &Gs[c][0]);
                const float dp = dot_smem<HD>(vr, &Gs[c][0]);
                axpy_smem<HD>(ak, p * (dp - Ds[c]), &Qs[c][0]);
            }
        }
    }
    if (live) {
        float* dkp = dk + (b * T + t) * ldk + (int64_t)hk * HD;
        float* dvp = dv + (b * T + t) * ldk + (int64_t)hk * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) { dkp[d] = ak[d] * scale; dvp[d] = av[d]; }
    }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

GAOT_API int gaot_attention_fwd(const float* q, const float* k, const float* v, float* o, float* lse, int64_t B, int64_t T,
                                int64_t H, int64_t Hkv, int64_t hd, int64_t ldq, int64_t ldk, float scale, void* stream) {
    if (!q || !k || !v || !o || !lse || B < 1 || T < 1 || H < 1 || Hkv < 1 || H % Hkv) return GAOT_E_BADARG;
    if (!al16(q) || !al16(k) || !al16(v) || !al16(o) || ldq % 4 || ldk % 4) return GAOT_E_ALIGN;
    if (B > 65535 || H > 65535) return GAOT_E_BADARG;
    dim3 grid((unsigned)((T + ATT_THREADS - 1) / ATT_THREADS), (unsigned)H, (unsigned)B);
    cudaStream_t st = (cudaStream_t)stream;
    switch (hd) {
        case 16: attn_fwd_kernel<16><<<grid, ATT_THREADS, 0, st>>>(q, k, v, o, lse, T, (int)H, (int)Hkv, ldq, ldk, scale); break;
        case 32: attn_fwd_kernel<32><<<grid, ATT_THREADS, 0, st>>>(q, k, v, o, lse, T, (int)H, (int)Hkv, ldq, ldk, scale); break;
        case 64: attn_fwd_kernel<64><<<grid, ATT_THREADS, 0, st>>>(q, k, v, o, lse, T, (int)H, (int)Hkv, ldq, ldk, scale); break;
        default: return GAOT_E_BADARG;
    }
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_attention_bwd(const float* q, const float* k, const float* v, const float* o, const float* lse,
                                const float* dout, float* delta, float* dq, float* dk, float* dv, int64_t B, int64_t T,
                                int64_t H, int64_t Hkv, int64_t hd, int64_t ldq, int64_t ldk, float scale, void* stream) {
    if (!q || !k || !v || !o || !lse || !dout || !delta || !dq || !dk || !dv) return GAOT_E_BADARG;
    if (B < 1 || T < 1 || H < 1 || Hkv < 1 || H % Hkv || B > 65535 || H > 65535) return GAOT_E_BADARG;
    if (!al16(q) || !al16(k) || !al16(v) || !al16(dout) || ldq % 4 || ldk % 4) return GAOT_E_ALIGN;
    dim3 gq((unsigned)((T + ATT_THREADS - 1) / ATT_THREADS), (unsigned)H, (unsigned)B);
    dim3 gk((unsigned)((T + ATT_THREADS - 1) / ATT_THREADS), (unsigned)Hkv, (unsigned)B);
    cudaStream_t st = (cudaStream_t)stream;
#define GAOT_ATT_BWD(HD_)                                                                                             \
    attn_bwd_dq_kernel<HD_><<<gq, ATT_THREADS, 0, st>>>(q, k, v, o, lse, dout, delta, dq, T, (int)H, (int)Hkv, ldq, ldk, scale); \
    attn_bwd_dkv_kernel<HD_><<<gk, ATT_THREADS, 0, st>>>(q, k, v, lse, dout, delta, dk, dv, T, (int)H, (int)Hkv, ldq, ldk, scale)
    switch (hd) {
        case 16: GAOT_ATT_BWD(16); break;
        case 32: GAOT_ATT_BWD(32); break;
        case 64: GAOT_ATT_BWD(64); break;
        default: return GAOT_E_BADARG;
    }
#undef GAOT_ATT_BWD
    GAOT_NOTE_LAUNCHES(1);
    GAOT_RETURN_LAST();
}
