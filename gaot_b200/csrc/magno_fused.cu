// Algebraically fused MAGNO operators.  Both exploit that the layers around AGNO are affine
// (ChannelMLP with n_layers=1, reference src/model/layers/magno.py:127-132, 457-462 and mlp.py:274-275):
//
// (1) lift + aggregate (encoder, magno.py:273-274 + agno.py:198,245-271):
//       out[b,m,c] = sum_{e in m} w[e,c] * ( sum_i Wl[c,i] * pn[b, idx_e, i] + bl[c] )
//     The lifted features f[B,N,C] (75 MB at cfg-2) are never written or gathered: the kernel gathers the
//     Cin raw input channels (4..16 bytes per point, L2 resident) and applies the lifting on the fly.
//     Backward: dw[e,c] = sum_b g[b,m,c] * f[b,idx_e,c] with f recomputed the same way;
//               dWl[c,i] = sum_{b,m,e} g[b,m,c] w[e,c] pn[b,idx_e,i];  dbl[c] = sum_{b,m,e} g[b,m,c] w[e,c]
//     (inputs carry no gradient, so no transposed graph is needed on the encoder side).
//
// (2) aggregate + recovery + projection (decoder, agno.py:198-271 + magno.py:684-688, 640-641):
//       out[b,n,o] = sum_{e in n} < u[o,e,:], r[b, idx_e, :] >,    u[o,e,c] = (Wp Wa)[o,c] * w[e,c]
//     (the batch-independent geometry/bias term is added by the caller).  The [B,N,C] decoder activations
//     (75 MB) and the two 295k-row GEMMs disappear; the kernel gathers latent rows (1 MB per sample, L2
//     resident) and writes Cout scalars per point.
//     Backward: dr[b,j,c] = sum_{e: idx_e=j} sum_o u[o,e,c] g[b,row_e,o]   (transposed graph)
//               du[o,e,c] = sum_b g[b,row_e,o] r[b,idx_e,c]
#include <cstdlib>

#include "common.cuh"
#include "../../include/gaot_b200.h"

namespace {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void fma4(float4& a, const float4& x, const float4& y) {
    a.x = fmaf(x.x, y.x, a.x); a.y = fmaf(x.y, y.y, a.y); a.z = fmaf(x.z, y.z, a.z); a.w = fmaf(x.w, y.w, a.w);
}
__device__ __forceinline__ void fma4s(float4& a, const float4& x, float s) {
    a.x = fmaf(x.x, s, a.x); a.y = fmaf(x.y, s, a.y); a.z = fmaf(x.z, s, a.z); a.w = fmaf(x.w, s, a.w);
}
__device__ __forceinline__ float dot4(const float4& x, const float4& y) {
    return fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, x.w * y.w)));
}

// FB = samples per work item: template parameter (8 / 4 / 2 / 1 by batch size, so a small batch does no redundant work)

// ---------------------------------------------------------------------------------------------
// (1) forward.  G lanes x float4 = one channel row; item = (row, chunk of FB samples)
// ---------------------------------------------------------------------------------------------
// Work items (row, chunk of FB samples) are walked with a persistent grid-stride loop; the CSR bounds of the
// item after next and the first neighbour index of the next item are fetched while the current item is being
// processed, so an item costs one dependent memory round trip (its feature rows) instead of three.
// Edges are taken U at a time: the U neighbour indices and U weight rows of a batch are independent loads, and the U*FB*CIN
// feature loads that depend on the indices are issued together - one dependent round trip per U edges instead of per edge.
// The register budget is bounded per instantiation (3 CTAs per SM when FB*CIN <= 8) so more warps hide that round trip.
// TP = true: the raw features come from the SAMPLE-MINOR copy pn_t [Ns, B, CIN]: the FB * CIN values an edge needs for the
// item's chunk of samples are contiguous, i.e. FB * CIN / 4 128-bit loads and ONE address computation per edge instead of
// FB * CIN scalar loads with an address each (integer address arithmetic was half of this kernel's issue slots).
template <int G, int CIN, int FB, bool TP>
__global__ void __launch_bounds__(256, (FB * CIN <= 16) ? 2 : 1)
lift_aggregate_fwd_kernel(const float* __restrict__ pn, const float* __restrict__ Wl, const float* __restrict__ bl,
                          const float* __restrict__ w, const int64_t* __restrict__ index, const int64_t* __restrict__ rs,
                          float* __restrict__ out, int B, int64_t Ns, int64_t Nd, int C, int64_t n_items) {
    constexpr int U = (FB * CIN <= 8) ? 4 : 2;
    const int64_t ngroups = (int64_t)gridDim.x * blockDim.x / G;
    const int gl = threadIdx.x % G;
    const int nvec = C >> 2;
    const int c0 = min(gl, nvec - 1) * 4;
    float4 wl[CIN];                                   // Wl[c0..c0+3, i]
#pragma unroll
    for (int i = 0; i < CIN; ++i)
        wl[i] = make_float4(__ldg(Wl + (c0 + 0) * CIN + i), __ldg(Wl + (c0 + 1) * CIN + i), __ldg(Wl + (c0 + 2) * CIN + i),
                            __ldg(Wl + (c0 + 3) * CIN + i));
    const float4 blv = ldg4(bl + c0);
    int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    int64_t a = 0, e = 0, na = 0, ne = 0;
    if (item < n_items) { a = __ldg(rs + item % Nd); e = __ldg(rs + item % Nd + 1); }
    if (item + ngroups < n_items) { na = __ldg(rs + (item + ngroups) % Nd); ne = __ldg(rs + (item + ngroups) % Nd + 1); }
    for (; item < n_items; item += ngroups) {
        int64_t nna = 0, nne = 0;
        if (item + 2 * ngroups < n_items) { nna = __ldg(rs + (item + 2 * ngroups) % Nd); nne = __ldg(rs + (item + 2 * ngroups) % Nd + 1); }
        const int64_t row = item % Nd;
        const int b0 = (int)(item / Nd) * FB;
        const float* pb0 = TP ? pn + (int64_t)b0 * CIN : pn + (int64_t)min(b0, B - 1) * Ns * CIN;
        [[maybe_unused]] int pb[FB];                  // sample offsets relative to the chunk's first sample (FB * Ns * CIN < 2^31)
        if constexpr (!TP) {
#pragma unroll
            for (int bb = 0; bb < FB; ++bb) pb[bb] = (int)((int64_t)(min(b0 + bb, B - 1) - min(b0, B - 1)) * Ns * CIN);
        }
        // s[bb][i] = sum_e w[e,c] pn[b,j,i] (per lane: 4 channels), sw = sum_e w[e,c]
        float4 s[FB][CIN];
        float4 sw = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int bb = 0; bb < FB; ++bb)
#pragma unroll
            for (int i = 0; i < CIN; ++i) s[bb][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t it = a; it < e; it += U) {
            int jj[U];
            float4 wv[U];
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const bool ok = it + q < e;
                const int64_t ee = ok ? it + q : it;
                jj[q] = (int)__ldg(index + ee) * (TP ? B * CIN : CIN);
                wv[q] = ok ? ldg4(w + ee * C + c0) : make_float4(0.f, 0.f, 0.f, 0.f);      // a zero weight masks the tail
            }
            if constexpr (TP) {
                constexpr int V4 = FB * CIN / 4;
                float4 pv[U][V4];
#pragma unroll
                for (int q = 0; q < U; ++q)
#pragma unroll
                    for (int v = 0; v < V4; ++v) pv[q][v] = ldg4(pb0 + jj[q] + 4 * v);
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    sw.x += wv[q].x; sw.y += wv[q].y; sw.z += wv[q].z; sw.w += wv[q].w;
                    const float* pf = reinterpret_cast<const float*>(pv[q]);
#pragma unroll
                    for (int bb = 0; bb < FB; ++bb)
#pragma unroll
                        for (int i = 0; i < CIN; ++i) fma4s(s[bb][i], wv[q], pf[bb * CIN + i]);
                }
            } else {
#pragma unroll
            for (int q = 0; q < U; ++q) {
                sw.x += wv[q].x; sw.y += wv[q].y; sw.z += wv[q].z; sw.w += wv[q].w;
#pragma unroll
                for (int bb = 0; bb < FB; ++bb)
#pragma unroll
                    for (int i = 0; i < CIN; ++i) fma4s(s[bb][i], wv[q], __ldg(pb0 + (pb[bb] + jj[q] + i)));
            }
            }
        }
        if (gl < nvec) {
#pragma unroll
            for (int bb = 0; bb < FB; ++bb) {
                if (b0 + bb < B) {
                    float4 o = make_float4(blv.x * sw.x, blv.y * sw.y, blv.z * sw.z, blv.w * sw.w);
#pragma unroll
                    for (int i = 0; i < CIN; ++i) fma4(o, wl[i], s[bb][i]);
                    *reinterpret_cast<float4*>(out + ((int64_t)(b0 + bb) * Nd + row) * C + c0) = o;
                }
            }
        }
        a = na; e = ne; na = nna; ne = nne;
    }
}

// (1) backward: dw (atomics over the sample chunks), dWl / dbl (CTA partial sums -> atomics)
template <int G, int CIN, int FB, bool TP>
__global__ void __launch_bounds__(256)
lift_aggregate_bwd_kernel(const float* __restrict__ g, const float* __restrict__ pn, const float* __restrict__ Wl,
                          const float* __restrict__ bl, const float* __restrict__ w, const int64_t* __restrict__ index,
                          const int64_t* __restrict__ rs, float* __restrict__ dw, float* __restrict__ dWl,
                          float* __restrict__ dbl, int B, int64_t Ns, int64_t Nd, int C, int64_t n_items, int chunks) {
    extern __shared__ float sred[];                   // [C * (CIN + 1)] CTA partial of dWl | dbl
    for (int t = threadIdx.x; t < C * (CIN + 1); t += blockDim.x) sred[t] = 0.f;
    __syncthreads();
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int gl = threadIdx.x % G;
    const int nvec = C >> 2;
    const int c0 = min(gl, nvec - 1) * 4;
    float4 dl[CIN];
    float4 db = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CIN; ++i) dl[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gid < n_items) {
        const int64_t row = gid % Nd;
        const int b0 = (int)(gid / Nd) * FB;
        float4 wl[CIN];
#pragma unroll
        for (int i = 0; i < CIN; ++i)
            wl[i] = make_float4(__ldg(Wl + (c0 + 0) * CIN + i), __ldg(Wl + (c0 + 1) * CIN + i), __ldg(Wl + (c0 + 2) * CIN + i),
                                __ldg(Wl + (c0 + 3) * CIN + i));
        const float4 blv = ldg4(bl + c0);
        float4 gv[FB];
        int64_t boff[FB];
#pragma unroll
        for (int bb = 0; bb < FB; ++bb) {
            const bool ok = b0 + bb < B;
            boff[bb] = TP ? 0 : (int64_t)min(b0 + bb, B - 1) * Ns * CIN;
            gv[bb] = ok ? ldg4(g + ((int64_t)(b0 + bb) * Nd + row) * C + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const int64_t a = __ldg(rs + row), e = __ldg(rs + row + 1);
        if (a < e) {
            int64_t j = __ldg(index + a);
            for (int64_t it = a; it < e; ++it) {
                const int64_t jn = __ldg(index + ((it + 1 < e) ? it + 1 : it));
                const float4 wv = ldg4(w + it * C + c0);
                float p[FB][CIN];
                if constexpr (TP) {
                    constexpr int V4 = FB * CIN / 4;
                    const float* src = pn + (j * B + b0) * CIN;          // sample-minor copy [Ns, B, CIN]
                    float4 pv[V4];
#pragma unroll
                    for (int v = 0; v < V4; ++v) pv[v] = ldg4(src + 4 * v);
                    const float* pf = reinterpret_cast<const float*>(pv);
#pragma unroll
                    for (int bb = 0; bb < FB; ++bb)
#pragma unroll
                        for (int i = 0; i < CIN; ++i) p[bb][i] = pf[bb * CIN + i];
                } else {
#pragma unroll
                for (int bb = 0; bb < FB; ++bb)
#pragma unroll
                    for (int i = 0; i < CIN; ++i) p[bb][i] = __ldg(pn + boff[bb] + j * CIN + i);
                }
                float4 dwe = make_float4(0.f, 0.f, 0.f, 0.f);        // sum_b g * f
                float4 gsum = make_float4(0.f, 0.f, 0.f, 0.f);       // sum_b g
                float4 gp[CIN];                                       // sum_b g * pn_i
#pragma unroll
                for (int i = 0; i < CIN; ++i) gp[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int bb = 0; bb < FB; ++bb) {
                    gsum.x += gv[bb].x; gsum.y += gv[bb].y; gsum.z += gv[bb].z; gsum.w += gv[bb].w;
#pragma unroll
                    for (int i = 0; i < CIN; ++i) fma4s(gp[i], gv[bb], p[bb][i]);
                }
                // f = Wl pn + bl  ->  sum_b g f = sum_i Wl_i * gp_i + bl * gsum
                dwe = make_float4(blv.x * gsum.x, blv.y * gsum.y, blv.z * gsum.z, blv.w * gsum.w);
#pragma unroll
                for (int i = 0; i < CIN; ++i) {
                    fma4(dwe, wl[i], gp[i]);
                    fma4(dl[i], wv, gp[i]);                            // dWl += w * (sum_b g pn_i)
                }
                fma4(db, wv, gsum);                                    // dbl += w * sum_b g
                if (gl < nvec) {
                    float4* dst = reinterpret_cast<float4*>(dw + it * C + c0);
                    if (chunks > 1) atomicAdd(dst, dwe); else *dst = dwe;
                }
                j = jn;
            }
        }
    }
    if (gl < nvec) {
#pragma unroll
        for (int i = 0; i < CIN; ++i) {
            atomicAdd(&sred[(c0 + 0) * (CIN + 1) + i], dl[i].x); atomicAdd(&sred[(c0 + 1) * (CIN + 1) + i], dl[i].y);
            atomicAdd(&sred[(c0 + 2) * (CIN + 1) + i], dl[i].z); atomicAdd(&sred[(c0 + 3) * (CIN + 1) + i], dl[i].w);
        }
        atomicAdd(&sred[(c0 + 0) * (CIN + 1) + CIN], db.x); atomicAdd(&sred[(c0 + 1) * (CIN + 1) + CIN], db.y);
        atomicAdd(&sred[(c0 + 2) * (CIN + 1) + CIN], db.z); atomicAdd(&sred[(c0 + 3) * (CIN + 1) + CIN], db.w);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < C * (CIN + 1); t += blockDim.x) {
        const int c = t / (CIN + 1), i = t % (CIN + 1);
        const float v = sred[t];
        if (v != 0.f) {
            if (i < CIN) atomicAdd(dWl + c * CIN + i, v); else atomicAdd(dbl + c, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// (2) forward: out[b,n,o] = sum_e <u[o,e,:], r[b,idx_e,:]>
// ---------------------------------------------------------------------------------------------
template <int G, int COUT, int FB>
__global__ void __launch_bounds__(256)
agno_dot_fwd_kernel(const float* __restrict__ r, const float* __restrict__ u, const int64_t* __restrict__ index,
                    const int64_t* __restrict__ rs, float* __restrict__ out, int B, int64_t Ns, int64_t Nd, int64_t E,
                    int C, int64_t n_items) {
    // persistent grid-stride walk with cross-item prefetch (see lift_aggregate_fwd_kernel); every lane of a warp
    // iterates the same number of times because the group shuffles below need full warps
    const int64_t ngroups = (int64_t)gridDim.x * blockDim.x / G;
    const int gl = threadIdx.x % G;
    const int nvec = C >> 2;
    const bool act = gl < nvec;
    const int c0 = min(gl, nvec - 1) * 4;
    const int64_t first = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / G;      // first group of this warp
    int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    int64_t a = 0, e = 0, na = 0, ne = 0, j0 = 0;
    if (item < n_items) { a = __ldg(rs + item % Nd); e = __ldg(rs + item % Nd + 1); }
    if (item + ngroups < n_items) { na = __ldg(rs + (item + ngroups) % Nd); ne = __ldg(rs + (item + ngroups) % Nd + 1); }
    if (a < e) j0 = __ldg(index + a);
    for (int64_t base = first; base < n_items; base += ngroups, item += ngroups) {
        const bool valid = item < n_items;
        int64_t nna = 0, nne = 0, nj0 = 0;
        if (item + 2 * ngroups < n_items) { nna = __ldg(rs + (item + 2 * ngroups) % Nd); nne = __ldg(rs + (item + 2 * ngroups) % Nd + 1); }
        if (na < ne) nj0 = __ldg(index + na);
        const int64_t row = valid ? item % Nd : 0;
        const int b0 = valid ? (int)(item / Nd) * FB : 0;
        int64_t boff[FB];
#pragma unroll
        for (int bb = 0; bb < FB; ++bb) boff[bb] = (int64_t)min(b0 + bb, B - 1) * Ns * C;
        float acc[FB][COUT];
#pragma unroll
        for (int bb = 0; bb < FB; ++bb)
#pragma unroll
            for (int o = 0; o < COUT; ++o) acc[bb][o] = 0.f;
        int64_t j = j0;
        for (int64_t it = a; it < e; ++it) {          // a == e for invalid items
            const int64_t jn = __ldg(index + ((it + 1 < e) ? it + 1 : it));
            float4 uv[COUT], rv[FB];
#pragma unroll
            for (int o = 0; o < COUT; ++o) uv[o] = ldg4(u + ((int64_t)o * E + it) * C + c0);
#pragma unroll
            for (int bb = 0; bb < FB; ++bb) rv[bb] = ldg4(r + boff[bb] + j * C + c0);
            if (act) {
#pragma unroll
                for (int bb = 0; bb < FB; ++bb)
#pragma unroll
                    for (int o = 0; o < COUT; ++o) acc[bb][o] += dot4(uv[o], rv[bb]);
            }
            j = jn;
        }
        // reduce the per-lane partial dot products over the G lanes of the group
#pragma unroll
        for (int bb = 0; bb < FB; ++bb)
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
                float v = acc[bb][o];
#pragma unroll
                for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off, G);
                acc[bb][o] = v;
            }
        if (gl == 0 && valid) {
#pragma unroll
            for (int bb = 0; bb < FB; ++bb)
                if (b0 + bb < B)
#pragma unroll
                    for (int o = 0; o < COUT; ++o) out[((int64_t)(b0 + bb) * Nd + row) * COUT + o] = acc[bb][o];
        }
        a = na; e = ne; j0 = nj0; na = nna; ne = nne;
    }
}

// (2) dr on the transposed graph: dr[b,j,c] = sum_{k in T(j)} sum_o u[o, perm_k, c] * g[b, trow_k, o]
template <int G, int COUT, int FB>
__global__ void __launch_bounds__(256)
agno_dot_bwd_r_kernel(const float* __restrict__ g, const float* __restrict__ u, const int64_t* __restrict__ t_rows,
                      const int64_t* __restrict__ t_perm, const int64_t* __restrict__ ts, float* __restrict__ dr, int B,
                      int64_t Ns, int64_t Nd, int64_t E, int C, int64_t n_items) {
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int gl = threadIdx.x % G;
    if (gid >= n_items) return;
    const int64_t src = gid % Ns;
    const int b0 = (int)(gid / Ns) * FB;
    const int nvec = C >> 2;
    const int c0 = min(gl, nvec - 1) * 4;
    float4 acc[FB];
#pragma unroll
    for (int bb = 0; bb < FB; ++bb) acc[bb] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t a = __ldg(ts + src), e = __ldg(ts + src + 1);
    for (int64_t k = a; k < e; ++k) {
        const int64_t row = __ldg(t_rows + k), pe = __ldg(t_perm + k);
        float4 uv[COUT];
#pragma unroll
        for (int o = 0; o < COUT; ++o) uv[o] = ldg4(u + ((int64_t)o * E + pe) * C + c0);
#pragma unroll
        for (int bb = 0; bb < FB; ++bb) {
            const int64_t gb = ((int64_t)min(b0 + bb, B - 1) * Nd + row) * COUT;
#pragma unroll
            for (int o = 0; o < COUT; ++o) fma4s(acc[bb], uv[o], __ldg(g + gb + o));
        }
    }
    if (gl < nvec) {
#pragma unroll
        for (int bb = 0; bb < FB; ++bb)
            if (b0 + bb < B) *reinterpret_cast<float4*>(dr + ((int64_t)(b0 + bb) * Ns + src) * C + c0) = acc[bb];
    }
}

// (2) du[o,e,c] = sum_b g[b,row_e,o] r[b,idx_e,c]; a group owns a destination row and a slice of 8 samples: the slice's
// upstream gradients g[b,row,:] are loaded once per row, every edge then costs 8 independent row loads and 8*COUT FMAs per
// lane; slices are combined with 128-bit red.add when the batch has more than one.
template <int G, int COUT>
__global__ void __launch_bounds__(256)
agno_dot_bwd_u_kernel(const float* __restrict__ g, const float* __restrict__ r, const int64_t* __restrict__ index,
                      const int64_t* __restrict__ rs, float* __restrict__ du, int B, int64_t Ns, int64_t Nd, int64_t E, int C,
                      int bsplit, int64_t n_items) {
    constexpr int PB = 8;
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int gl = threadIdx.x % G;
    if (gid >= n_items) return;
    const int64_t row = gid % Nd;
    const int b_lo = (int)(gid / Nd) * PB;
    const int nvec = C >> 2;
    const int c0 = min(gl, nvec - 1) * 4;
    const int64_t a = __ldg(rs + row), e = __ldg(rs + row + 1);
    if (a == e) return;
    float gs[PB][COUT];
    const float* rb[PB];
#pragma unroll
    for (int q = 0; q < PB; ++q) {
        const int bq = min(b_lo + q, B - 1);
        rb[q] = r + (int64_t)bq * Ns * C + c0;
#pragma unroll
        for (int o = 0; o < COUT; ++o) gs[q][o] = (b_lo + q < B) ? __ldg(g + ((int64_t)bq * Nd + row) * COUT + o) : 0.f;
    }
    int64_t j = __ldg(index + a);
    for (int64_t it = a; it < e; ++it) {
        const int64_t jn = __ldg(index + ((it + 1 < e) ? it + 1 : it));
        const int64_t joff = j * C;
        float4 rv[PB];
#pragma unroll
        for (int q = 0; q < PB; ++q) rv[q] = ldg4(rb[q] + joff);
        float4 acc[COUT];
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < PB; ++q)
#pragma unroll
            for (int o = 0; o < COUT; ++o) fma4s(acc[o], rv[q], gs[q][o]);
        if (gl < nvec) {
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
                float4* dst = reinterpret_cast<float4*>(du + ((int64_t)o * E + it) * C + c0);
                if (bsplit > 1) atomicAdd(dst, acc[o]); else *dst = acc[o];
            }
        }
        j = jn;
    }
}

// ---------------------------------------------------------------------------------------------
// (2') the decoder operator walked by SOURCE node (latent token) on the transposed graph.
// Row-wise, every edge gathers the 256-byte latent row of every sample again: E * B * C * 4 bytes (257 MB at cfg-2) cross
// the L2 -> SM fabric for 33.5 MB of distinct data.  Source-wise, a group owns one latent row j: it reads r[b, j, :] ONCE
// per sample and serves all edges leaving j (7.7 on average) from registers:
//   forward : out[b, row_k, o] += < u[o, e_k, :], r[b, j, :] >            16-lane shuffle reduction, one fp32 red.add per (b, edge)
//   backward: dr[b, j, :]       = sum_k sum_o u[o, e_k, :] g[b, row_k, o]   owned by the group: plain stores
//             du[o, e_k, :]     = sum_b g[b, row_k, o] r[b, j, :]           owned by the group: plain stores, ONE pass for dr and du
// KE edges of the source are held in registers per pass (sources with more edges take several passes).
// item = (source j, slice of the batch); with bslices > 1 the du of a source is reduced over its slices with red.add.
// ---------------------------------------------------------------------------------------------
template <int G, int COUT, int KE>
__global__ void __launch_bounds__(256)
agno_dot_fwd_t_kernel(const float* __restrict__ r, const float* __restrict__ u, const int64_t* __restrict__ t_rows,
                      const int64_t* __restrict__ t_perm, const int64_t* __restrict__ ts, float* __restrict__ out, int B,
                      int Ns, int Nd, int64_t E, int C, int64_t n_items, int bslices) {
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int gl = threadIdx.x % G;
    const bool valid = gid < n_items;                  // whole groups fall out together, but shuffles want full warps
    const int j = valid ? (int)(gid % Ns) : 0;
    const int slice = valid ? (int)(gid / Ns) : 0;
    const int per = (B + bslices - 1) / bslices;
    const int b_lo = slice * per, b_hi = valid ? min(B, b_lo + per) : b_lo;
    const int nvec = C >> 2;
    const bool act = gl < nvec;
    const int c0 = min(gl, nvec - 1) * 4;
    const int64_t a = valid ? __ldg(ts + j) : 0, e = valid ? __ldg(ts + j + 1) : 0;
    // the groups of a warp run different trip counts (degree, batch slice): shuffles synchronise the group's own lanes only
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (((threadIdx.x & 31) / G) * G));
    for (int64_t k0 = a; k0 < e; k0 += KE) {
        int rows[KE];
        float4 uv[COUT][KE];
#pragma unroll
        for (int k = 0; k < KE; ++k) {
            const bool ok = k0 + k < e;
            const int64_t kk = ok ? k0 + k : k0;
            rows[k] = ok ? (int)__ldg(t_rows + kk) : -1;
            const int64_t pe = __ldg(t_perm + kk);
#pragma unroll
            for (int o = 0; o < COUT; ++o)
                uv[o][k] = (ok && act) ? ldg4(u + ((int64_t)o * E + pe) * C + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll 2
        for (int b = b_lo; b < b_hi; ++b) {
            const float4 rv = ldg4(r + ((int64_t)b * Ns + j) * C + c0);
            float v[COUT][KE];
#pragma unroll
            for (int o = 0; o < COUT; ++o)
#pragma unroll
                for (int k = 0; k < KE; ++k) {
                    float x = dot4(uv[o][k], rv);
#pragma unroll
                    for (int off = G / 2; off > 0; off >>= 1) x += __shfl_xor_sync(gmask, x, off, G);
                    v[o][k] = x;
                }
            // lane k of the group issues the reduction of edge k
#pragma unroll
            for (int k = 0; k < KE; ++k) {
                if (gl == (k % G) && rows[k] >= 0) {
#pragma unroll
                    for (int o = 0; o < COUT; ++o) atomicAdd(out + ((int64_t)b * Nd + rows[k]) * COUT + o, v[o][k]);
                }
            }
        }
    }
}

template <int G, int COUT, int KE>
__global__ void __launch_bounds__(256, 2)
agno_dot_bwd_t_kernel(const float* __restrict__ g, const float* __restrict__ r, const float* __restrict__ u,
                      const int64_t* __restrict__ t_rows, const int64_t* __restrict__ t_perm, const int64_t* __restrict__ ts,
                      float* __restrict__ dr, float* __restrict__ du, int B, int Ns, int Nd, int64_t E, int C, int64_t n_items,
                      int bslices) {
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int gl = threadIdx.x % G;
    if (gid >= n_items) return;
    const int j = (int)(gid % Ns);
    const int slice = (int)(gid / Ns);
    const int per = (B + bslices - 1) / bslices;
    const int b_lo = slice * per, b_hi = min(B, b_lo + per);
    const int nvec = C >> 2;
    const bool act = gl < nvec;
    const int c0 = min(gl, nvec - 1) * 4;
    const int64_t a = __ldg(ts + j), e = __ldg(ts + j + 1);
    if (a == e) {                                      // a source without edges: dr = 0
        if (dr && act)
            for (int b = b_lo; b < b_hi; ++b) *reinterpret_cast<float4*>(dr + ((int64_t)b * Ns + j) * C + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    for (int64_t k0 = a; k0 < e; k0 += KE) {
        int rows[KE];
        int pes[KE];                                   // edge ids fit 32 bits (E * C * 4 bytes of weights would not exist otherwise)
        float4 uv[COUT][KE], acc[COUT][KE];
#pragma unroll
        for (int k = 0; k < KE; ++k) {
            const bool ok = k0 + k < e;
            const int64_t kk = ok ? k0 + k : k0;
            rows[k] = (int)__ldg(t_rows + kk);
            pes[k] = ok ? (int)__ldg(t_perm + kk) : -1;
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
                uv[o][k] = (ok && dr) ? ldg4(u + ((int64_t)o * E + pes[k]) * C + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
                acc[o][k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll 2
        for (int b = b_lo; b < b_hi; ++b) {
            const float4 rv = ldg4(r + ((int64_t)b * Ns + j) * C + c0);
            float gk[COUT][KE];
#pragma unroll
            for (int k = 0; k < KE; ++k)
#pragma unroll
                for (int o = 0; o < COUT; ++o) gk[o][k] = pes[k] >= 0 ? __ldg(g + ((int64_t)b * Nd + rows[k]) * COUT + o) : 0.f;
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < KE; ++k)
#pragma unroll
                for (int o = 0; o < COUT; ++o) { fma4s(d, uv[o][k], gk[o][k]); fma4s(acc[o][k], rv, gk[o][k]); }
            if (dr && act) {
                float4* dst = reinterpret_cast<float4*>(dr + ((int64_t)b * Ns + j) * C + c0);
                if (k0 > a) { const float4 old = *dst; d.x += old.x; d.y += old.y; d.z += old.z; d.w += old.w; }   // later pass of the same owner
                *dst = d;
            }
        }
        if (du && act) {
#pragma unroll
            for (int k = 0; k < KE; ++k) {
                if (pes[k] < 0) continue;
#pragma unroll
                for (int o = 0; o < COUT; ++o) {
                    float4* dst = reinterpret_cast<float4*>(du + ((int64_t)o * E + pes[k]) * C + c0);
                    if (bslices > 1) atomicAdd(dst, acc[o][k]); else *dst = acc[o][k];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// (2'') decoder operator with one LANE per SAMPLE (batches of >= 16 samples).  A warp owns a source node j and 32 samples:
// lane b keeps the whole latent row r[b, j, :] (C floats) in registers, the weight rows u[e, :] of the edges leaving j are
// warp-uniform (broadcast) loads, so a dot product needs no cross-lane reduction at all:
//   forward : out[b, row_k, o] += < u[o, e_k, :], r[b, j, :] >     C FMAs per lane and edge, one red.add per lane
//   backward: dr[b, j, :] = sum_k sum_o g[b, row_k, o] u[o, e_k, :]  C FMAs per lane and edge, the row is stored once
// ~85 instructions per (edge, 32 samples) against ~650 for the 16-lanes-per-row forms above (shuffle reductions, per-lane
// address arithmetic); every latent row is read exactly once.  du is the one product that reduces over samples and
// stays on the row-walk kernel.
// ---------------------------------------------------------------------------------------------
// The weight rows of ALL edges of a pass (up to 16 rows of C floats) are staged in shared memory by the whole warp with
// coalesced 256-byte requests issued together - one memory round trip per pass instead of one per edge - and then read back
// as warp-uniform (broadcast) LDS.128.
template <int C4, int COUT, bool BWD>
__global__ void __launch_bounds__(128)
agno_dot_s_kernel(const float* __restrict__ r, const float* __restrict__ u, const float* __restrict__ g,
                  const int64_t* __restrict__ t_rows, const int64_t* __restrict__ t_perm, const int64_t* __restrict__ ts,
                  float* __restrict__ out, int B, int Ns, int Nd, int64_t E, int64_t n_items) {
    constexpr int C = C4 * 4;
    constexpr int EK = 16 / COUT;                      // edges per pass: EK * COUT <= 16 staged rows per warp
    __shared__ float4 su_all[4][16][C4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float4 (*su)[C4] = su_all[wib];
    const int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (item >= n_items) return;                       // whole warps leave together
    const int j = (int)(item % Ns);
    const int b = (int)(item / Ns) * 32 + lane;
    const bool valid = b < B;
    const int bc = valid ? b : B - 1;
    const int64_t a = __ldg(ts + j), e = __ldg(ts + j + 1);
    float4 acc[C4];                                    // FWD: r[b, j, :]   BWD: dr[b, j, :]
    const float4* rrow = reinterpret_cast<const float4*>(r + ((int64_t)bc * Ns + j) * C);
#pragma unroll
    for (int q = 0; q < C4; ++q) acc[q] = BWD ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(rrow + q);
    for (int64_t k0 = a; k0 < e; k0 += EK) {
        const int ne = (int)min((int64_t)EK, e - k0);
        int myrow = 0, mype = 0;
        if (lane < ne) { myrow = (int)__ldg(t_rows + k0 + lane); mype = (int)__ldg(t_perm + k0 + lane); }
        const int tot = ne * COUT * C4;                // float4 pieces to stage
        for (int i0 = 0; i0 < tot; i0 += 32) {
            const int i = i0 + lane;
            const bool ok = i < tot;
            const int kk = ok ? i / (COUT * C4) : 0, rem = ok ? i % (COUT * C4) : 0;
            const int pe = __shfl_sync(0xffffffffu, mype, kk);
            if (ok) su[kk * COUT + rem / C4][rem % C4] = ldg4(u + (((int64_t)(rem / C4) * E + pe) * C4 + rem % C4) * 4);
        }
        float gk[BWD ? EK * COUT : 1];
        if (BWD) {
#pragma unroll
            for (int kk = 0; kk < EK; ++kk) {
                const int row = __shfl_sync(0xffffffffu, myrow, kk);
#pragma unroll
                for (int o = 0; o < COUT; ++o) gk[kk * COUT + o] = kk < ne ? __ldg(g + ((int64_t)bc * Nd + row) * COUT + o) : 0.f;
            }
        }
        __syncwarp();
#pragma unroll
        for (int kk = 0; kk < EK; ++kk) {
            if (kk >= ne) break;                        // warp-uniform
            const int row = __shfl_sync(0xffffffffu, myrow, kk);
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
                const float4* urow = su[kk * COUT + o];
                if (BWD) {
#pragma unroll
                    for (int q = 0; q < C4; ++q) fma4s(acc[q], urow[q], gk[kk * COUT + o]);
                } else {
                    float d0 = 0.f, d1 = 0.f;            // two chains
#pragma unroll
                    for (int q = 0; q < C4; q += 2) {
                        d0 += dot4(urow[q], acc[q]);
                        if (q + 1 < C4) d1 += dot4(urow[q + 1], acc[q + 1]);
                    }
                    if (valid) atomicAdd(out + ((int64_t)b * Nd + row) * COUT + o, d0 + d1);
                }
            }
        }
        __syncwarp();                                  // the next pass overwrites the staged rows
    }
    if (BWD && valid) {
        float4* drow = reinterpret_cast<float4*>(out + ((int64_t)b * Ns + j) * C);
#pragma unroll
        for (int q = 0; q < C4; ++q) drow[q] = acc[q];
    }
}

// ---------------------------------------------------------------------------------------------
// (2-du) du of the decoder operator on the transposed graph with one LANE per CHANNEL: du[o, e_k, :] = sum_b g[b, row_k, o] r[b, j, :]
// for the edges e_k leaving source j.  A warp owns a source node: for every sample it reads the latent row r[b, j, :] ONCE
// (coalesced, lane l holds channels l, l + 32, ..) and the upstream gradients g[b, row_k, o] as warp-uniform (broadcast) loads,
// and accumulates the KE edge gradients in registers - no reduction across lanes, no atomics, no zero fill, every latent row
// read once per pass of KE edges (33.5 MB per pass at cfg-2) where the row walk gathers it once per edge (257 MB through the L2).
// ---------------------------------------------------------------------------------------------
template <int VPL, int COUT>
__global__ void __launch_bounds__(256)
agno_dot_du_c_kernel(const float* __restrict__ r, const float* __restrict__ g, const int64_t* __restrict__ t_rows,
                     const int64_t* __restrict__ t_perm, const int64_t* __restrict__ ts, float* __restrict__ du, int B, int Ns,
                     int Nd, int64_t E) {
    constexpr int C = 32 * VPL;
    constexpr int KE = COUT == 1 ? 8 : 4;              // edges per pass (accumulators: KE * COUT * VPL registers)
    // upstream gradients of a pass, staged per warp as [edge][o][sample]: lane b fetches g[b, row_k, o] once (one load per
    // lane and edge), the accumulation loop reads them back four samples at a time as warp-uniform 128-bit loads
    __shared__ __align__(16) float gs_all[8][KE * COUT][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float (*gs)[32] = gs_all[wib];
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= Ns) return;                               // whole warps leave together
    const int64_t a = __ldg(ts + j), e = __ldg(ts + j + 1);
    const float* rj = r + j * C + lane;
    const int64_t rstride = (int64_t)Ns * C, gstride = (int64_t)Nd * COUT;
    for (int64_t k0 = a; k0 < e; k0 += KE) {
        const int ne = (int)min((int64_t)KE, e - k0);
        int myrow = 0;
        int64_t mype = 0;
        if (lane < ne) { myrow = (int)__ldg(t_rows + k0 + lane); mype = __ldg(t_perm + k0 + lane); }
        int rows[KE];
#pragma unroll
        for (int k = 0; k < KE; ++k) rows[k] = __shfl_sync(0xffffffffu, myrow, k < ne ? k : 0);
        float tot[KE][COUT][VPL];
#pragma unroll
        for (int k = 0; k < KE; ++k)
#pragma unroll
            for (int o = 0; o < COUT; ++o)
#pragma unroll
                for (int v = 0; v < VPL; ++v) tot[k][o][v] = 0.f;
        for (int b0 = 0; b0 < B; b0 += 32) {           // slices of 32 samples
            const int bl = b0 + lane;
            __syncwarp();                              // the previous slice / pass has been read
#pragma unroll
            for (int k = 0; k < KE; ++k)
#pragma unroll
                for (int o = 0; o < COUT; ++o)
                    gs[k * COUT + o][lane] = (bl < B && k < ne) ? __ldg(g + bl * gstride + (int64_t)rows[k] * COUT + o) : 0.f;
            __syncwarp();
            const int nb = min(32, B - b0);
            // (fully unrolled, with the row loads of all eight steps hoisted, the kernel needs 80 registers - three CTAs per SM
            // instead of four - and ran 36 -> 43 us; capped at 64 registers it spills)
#pragma unroll 1
            for (int b4 = 0; b4 < nb; b4 += 4) {
                float rv[4][VPL];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int v = 0; v < VPL; ++v) rv[q][v] = __ldg(rj + (int64_t)min(b0 + b4 + q, B - 1) * rstride + 32 * v);   // a zero g masks the tail
#pragma unroll
                for (int k = 0; k < KE; ++k)
#pragma unroll
                    for (int o = 0; o < COUT; ++o) {
                        const float4 g4 = *reinterpret_cast<const float4*>(&gs[k * COUT + o][b4]);
#pragma unroll
                        for (int v = 0; v < VPL; ++v) {
                            float t = tot[k][o][v];
                            t = fmaf(g4.x, rv[0][v], t); t = fmaf(g4.y, rv[1][v], t); t = fmaf(g4.z, rv[2][v], t); t = fmaf(g4.w, rv[3][v], t);
                            tot[k][o][v] = t;
                        }
                    }
            }
        }
#pragma unroll
        for (int k = 0; k < KE; ++k) {
            const int64_t pe = __shfl_sync(0xffffffffu, mype, k < ne ? k : 0);
            if (k < ne) {
#pragma unroll
                for (int o = 0; o < COUT; ++o)
#pragma unroll
                    for (int v = 0; v < VPL; ++v) du[((int64_t)o * E + pe) * C + 32 * v + lane] = tot[k][o][v];
            }
        }
    }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int fb_of(int64_t B) {
    static int cap = 0;                                // GAOT_B200_FB: experiment knob, samples per work item of the encoder kernels
    if (!cap) { const char* e = getenv("GAOT_B200_FB"); cap = e ? atoi(e) : 8; if (cap != 1 && cap != 2 && cap != 4 && cap != 8) cap = 8; }
    const int fb = B >= 8 ? 8 : (B >= 4 ? 4 : (B >= 2 ? 2 : 1));
    return fb < cap ? fb : cap;
}
inline int group_of(int64_t C) { const int nv = (int)(C / 4); return nv <= 4 ? 4 : (nv <= 8 ? 8 : (nv <= 16 ? 16 : 32)); }

#define GAOT_DISPATCH_G(G_, ...)  \
    switch (G_) { case 4: { constexpr int G = 4; __VA_ARGS__; break; } case 8: { constexpr int G = 8; __VA_ARGS__; break; } \
                  case 16: { constexpr int G = 16; __VA_ARGS__; break; } default: { constexpr int G = 32; __VA_ARGS__; break; } }
#define GAOT_DISPATCH_FB(FB_, ...)  \
    switch (FB_) { case 8: { constexpr int FB = 8; __VA_ARGS__; break; } case 4: { constexpr int FB = 4; __VA_ARGS__; break; } \
                   case 2: { constexpr int FB = 2; __VA_ARGS__; break; } default: { constexpr int FB = 1; __VA_ARGS__; break; } }
#define GAOT_DISPATCH_N(N_, NAME, ...)  \
    switch (N_) { case 1: { constexpr int NAME = 1; __VA_ARGS__; break; } case 2: { constexpr int NAME = 2; __VA_ARGS__; break; } \
                  case 3: { constexpr int NAME = 3; __VA_ARGS__; break; } case 4: { constexpr int NAME = 4; __VA_ARGS__; break; } \
                  default: return GAOT_E_BADARG; }

// the sample-minor copy serves a launch when every work item reads whole 16-byte pieces of it
inline bool use_pn_t(const float* pn_t, int fb, int64_t B, int64_t Ns, int64_t Cin) {
    return pn_t && al16(pn_t) && (fb * Cin) % 4 == 0 && B % fb == 0 && Ns * B * Cin < 0x7fffffffLL;
}
}  // namespace

GAOT_API int gaot_lift_aggregate_fwd(const float* pn, const float* Wl, const float* bl, const float* w, const int64_t* index,
                                     const int64_t* row_splits, float* out, int64_t B, int64_t Ns, int64_t Nd, int64_t C,
                                     int64_t Cin, const float* pn_t, void* stream) {
    if (!pn || !Wl || !bl || !w || !index || !row_splits || !out || B < 1 || Ns < 0 || Nd < 0) return GAOT_E_BADARG;
    if (C % 4 || C > 128 || Cin < 1 || Cin > 4) return GAOT_E_BADARG;
    if (!al16(w) || !al16(out) || !al16(bl)) return GAOT_E_ALIGN;
    if (Nd == 0) return 0;
    const int g_ = group_of(C);
    const int fb = fb_of(B);
    const int64_t n_items = Nd * ((B + fb - 1) / fb);
    int64_t blocks = (n_items * g_ + 255) / 256;
    const int resident = fb * (int)Cin <= 16 ? 2 : 1;                                 // CTAs per SM the kernel's register bound allows
    if (blocks > (int64_t)gaot_sm_count() * resident) blocks = (int64_t)gaot_sm_count() * resident;   // persistent grid-stride loop
    cudaStream_t st = (cudaStream_t)stream;
    if (use_pn_t(pn_t, fb, B, Ns, Cin)) {
        GAOT_DISPATCH_FB(fb, GAOT_DISPATCH_G(g_, GAOT_DISPATCH_N((int)Cin, CIN, (lift_aggregate_fwd_kernel<G, CIN, FB, (FB * CIN) % 4 == 0><<<(unsigned)blocks, 256, 0, st>>>(
                                                              pn_t, Wl, bl, w, index, row_splits, out, (int)B, Ns, Nd, (int)C, n_items)))));
    } else {
        GAOT_DISPATCH_FB(fb, GAOT_DISPATCH_G(g_, GAOT_DISPATCH_N((int)Cin, CIN, (lift_aggregate_fwd_kernel<G, CIN, FB, false><<<(unsigned)blocks, 256, 0, st>>>(
                                                              pn, Wl, bl, w, index, row_splits, out, (int)B, Ns, Nd, (int)C, n_items)))));
    }
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_lift_aggregate_bwd(const float* g, const float* pn, const float* Wl, const float* bl, const float* w,
                                     const int64_t* index, const int64_t* row_splits, float* dw, float* dWl, float* dbl,
                                     int64_t B, int64_t Ns, int64_t Nd, int64_t E, int64_t C, int64_t Cin, const float* pn_t,
                                     void* stream) {
    if (!g || !pn || !Wl || !bl || !w || !index || !row_splits || !dw || !dWl || !dbl || B < 1 || E < 0) return GAOT_E_BADARG;
    if (C % 4 || C > 128 || Cin < 1 || Cin > 4) return GAOT_E_BADARG;
    if (!al16(w) || !al16(g) || !al16(dw) || !al16(bl)) return GAOT_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_CHECK(cudaMemsetAsync(dWl, 0, sizeof(float) * (size_t)(C * Cin), st));
    GAOT_CHECK(cudaMemsetAsync(dbl, 0, sizeof(float) * (size_t)C, st));
    if (Nd == 0 || E == 0) return 0;
    const int fb = fb_of(B);
    const int chunks = (int)((B + fb - 1) / fb);
    if (chunks > 1) GAOT_CHECK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)(E * C), st));
    const int g_ = group_of(C);
    const int64_t n_items = Nd * chunks;
    const int64_t blocks = (n_items * g_ + 255) / 256;
    if (blocks > 0x7fffffffLL) return GAOT_E_BADARG;
    const size_t smem = sizeof(float) * (size_t)(C * (Cin + 1));
    if (use_pn_t(pn_t, fb, B, Ns, Cin)) {
        GAOT_DISPATCH_FB(fb, GAOT_DISPATCH_G(g_, GAOT_DISPATCH_N((int)Cin, CIN, (lift_aggregate_bwd_kernel<G, CIN, FB, (FB * CIN) % 4 == 0><<<(unsigned)blocks, 256, smem, st>>>(
                                                              g, pn_t, Wl, bl, w, index, row_splits, dw, dWl, dbl, (int)B, Ns, Nd, (int)C,
                                                              n_items, chunks)))));
    } else {
        GAOT_DISPATCH_FB(fb, GAOT_DISPATCH_G(g_, GAOT_DISPATCH_N((int)Cin, CIN, (lift_aggregate_bwd_kernel<G, CIN, FB, false><<<(unsigned)blocks, 256, smem, st>>>(
                                                              g, pn, Wl, bl, w, index, row_splits, dw, dWl, dbl, (int)B, Ns, Nd, (int)C,
                                                              n_items, chunks)))));
    }
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_agno_dot_fwd(const float* r, const float* u, const int64_t* index, const int64_t* row_splits, float* out,
                               int64_t B, int64_t Ns, int64_t Nd, int64_t E, int64_t C, int64_t Cout, void* stream) {
    if (!r || !u || !index || !row_splits || !out || B < 1 || Ns < 0 || Nd < 0 || E < 0) return GAOT_E_BADARG;
    if (C % 4 || C > 128 || Cout < 1 || Cout > 4) return GAOT_E_BADARG;
    if (!al16(r) || !al16(u)) return GAOT_E_ALIGN;
    if (Nd == 0) return 0;
    const int g_ = group_of(C);
    const int fb = fb_of(B);
    const int64_t n_items = Nd * ((B + fb - 1) / fb);
    int64_t blocks = (n_items * g_ + 255) / 256;
    if (blocks > gaot_sm_count() * 2) blocks = gaot_sm_count() * 2;                 // persistent (2 resident CTAs/SM at 128 registers): grid-stride loop
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_DISPATCH_FB(fb, GAOT_DISPATCH_G(g_, GAOT_DISPATCH_N((int)Cout, COUT, (agno_dot_fwd_kernel<G, COUT, FB><<<(unsigned)blocks, 256, 0, st>>>(
                                                            r, u, index, row_splits, out, (int)B, Ns, Nd, E, (int)C, n_items)))));
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_agno_dot_bwd(const float* g, const float* r, const float* u, const int64_t* index, const int64_t* row_splits,
                               const int64_t* t_splits, const int64_t* t_rows, const int64_t* t_perm, float* dr, float* du,
                               int64_t B, int64_t Ns, int64_t Nd, int64_t E, int64_t C, int64_t Cout, void* stream) {
    if (!g || !r || !u || !index || !row_splits || B < 1 || Ns < 0 || Nd < 0 || E < 0) return GAOT_E_BADARG;
    if (dr && (!t_splits || !t_rows || !t_perm)) return GAOT_E_BADARG;
    if (C % 4 || C > 128 || Cout < 1 || Cout > 4) return GAOT_E_BADARG;
    if (!al16(r) || !al16(u) || (dr && !al16(dr)) || (du && !al16(du))) return GAOT_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int g_ = group_of(C);
    int launches = 0;
    if (dr && Ns > 0) {
        const int fb = fb_of(B);
        const int64_t n_items = Ns * ((B + fb - 1) / fb);
        const int64_t blocks = (n_items * g_ + 255) / 256;
        if (blocks > 0x7fffffffLL) return GAOT_E_BADARG;
        GAOT_DISPATCH_FB(fb, GAOT_DISPATCH_G(g_, GAOT_DISPATCH_N((int)Cout, COUT, (agno_dot_bwd_r_kernel<G, COUT, FB><<<(unsigned)blocks, 256, 0, st>>>(
                                                                g, u, t_rows, t_perm, t_splits, dr, (int)B, Ns, Nd, E, (int)C, n_items)))));
        ++launches;
    }
    if (du && Nd > 0 && E > 0) {
        const int bsplit = (int)((B + 7) / 8);                  // slices of 8 samples (agno_dot_bwd_u_kernel)
        if (bsplit > 1) GAOT_CHECK(cudaMemsetAsync(du, 0, sizeof(float) * (size_t)(Cout * E * C), st));
        const int64_t n_items = Nd * bsplit;
        const int64_t blocks = (n_items * g_ + 255) / 256;
        if (blocks > 0x7fffffffLL) return GAOT_E_BADARG;
        GAOT_DISPATCH_G(g_, GAOT_DISPATCH_N((int)Cout, COUT, (agno_dot_bwd_u_kernel<G, COUT><<<(unsigned)blocks, 256, 0, st>>>(
                                                                g, r, index, row_splits, du, (int)B, Ns, Nd, E, (int)C, bsplit, n_items))));
        ++launches;
    }
    if (launches == 0) return 0;
    if (launches == 2) GAOT_NOTE_LAUNCHES(1);
    GAOT_RETURN_LAST();
}

namespace {
// lane-per-sample kernels: C4 in {4, 8, 12, 16} (C = 16 ... 64), batches of at least 16 samples
template <bool BWD>
bool launch_dot_s(const float* r, const float* u, const float* g, const int64_t* t_rows, const int64_t* t_perm, const int64_t* ts,
                  float* out, int B, int Ns, int Nd, int64_t E, int C, int Cout, cudaStream_t st) {
    if (B < 16 || C % 16 || C > 64) return false;
    const int64_t n_items = (int64_t)Ns * ((B + 31) / 32);
    const int64_t blocks = (n_items * 32 + 127) / 128;
    if (blocks > 0x7fffffffLL) return false;
#define GAOT_S_CASE(C4_, CO_) if (C == 4 * C4_ && Cout == CO_) { \
        agno_dot_s_kernel<C4_, CO_, BWD><<<(unsigned)blocks, 128, 0, st>>>(r, u, g, t_rows, t_perm, ts, out, B, Ns, Nd, E, n_items); return true; }
    GAOT_S_CASE(4, 1) GAOT_S_CASE(8, 1) GAOT_S_CASE(12, 1) GAOT_S_CASE(16, 1)
    GAOT_S_CASE(4, 2) GAOT_S_CASE(8, 2) GAOT_S_CASE(12, 2) GAOT_S_CASE(16, 2)
    GAOT_S_CASE(4, 3) GAOT_S_CASE(8, 3) GAOT_S_CASE(12, 3) GAOT_S_CASE(16, 3)
    GAOT_S_CASE(4, 4) GAOT_S_CASE(8, 4) GAOT_S_CASE(12, 4) GAOT_S_CASE(16, 4)
#undef GAOT_S_CASE
    return false;
}
template <int G, int COUT, int KE>
void launch_dot_fwd_t(const float* r, const float* u, const int64_t* t_rows, const int64_t* t_perm, const int64_t* ts, float* out,
                      int B, int Ns, int Nd, int64_t E, int C, int64_t n_items, int bslices, cudaStream_t st) {
    const int64_t blocks = (n_items * G + 255) / 256;
    agno_dot_fwd_t_kernel<G, COUT, KE><<<(unsigned)blocks, 256, 0, st>>>(r, u, t_rows, t_perm, ts, out, B, Ns, Nd, E, C, n_items, bslices);
}
template <int G, int COUT, int KE>
void launch_dot_bwd_t(const float* g, const float* r, const float* u, const int64_t* t_rows, const int64_t* t_perm, const int64_t* ts,
                      float* dr, float* du, int B, int Ns, int Nd, int64_t E, int C, int64_t n_items, int bslices, cudaStream_t st) {
    const int64_t blocks = (n_items * G + 255) / 256;
    agno_dot_bwd_t_kernel<G, COUT, KE><<<(unsigned)blocks, 256, 0, st>>>(g, r, u, t_rows, t_perm, ts, dr, du, B, Ns, Nd, E, C, n_items,
                                                                        bslices);
}
// batch slices per source: enough groups to fill the chip (sources alone do when the batch is small and the mesh is large)
inline int slices_of(int64_t Ns, int g_, int64_t B) {
    const int64_t want = (int64_t)gaot_sm_count() * 1536;               // threads for ~75 % occupancy
    int s = 1;
    while (s < B && Ns * g_ * s < want && s < 8) s <<= 1;
    return (int)(s > B ? B : s);
}
}  // namespace

// Source-walk form of gaot_agno_dot_fwd (see the kernel comment): needs the transposed graph; `out` is zeroed here and reduced
// into with fp32 red.add (summation order inside a row is not fixed; the row-walk gaot_agno_dot_fwd is the deterministic form).
GAOT_API int gaot_agno_dot_fwd_t(const float* r, const float* u, const int64_t* t_splits, const int64_t* t_rows,
                                 const int64_t* t_perm, float* out, int64_t B, int64_t Ns, int64_t Nd, int64_t E, int64_t C,
                                 int64_t Cout, void* stream) {
    if (!r || !u || !t_splits || !t_rows || !t_perm || !out || B < 1 || Ns < 0 || Nd < 0 || E < 0) return GAOT_E_BADARG;
    if (C % 4 || C > 128 || Cout < 1 || Cout > 4) return GAOT_E_BADARG;
    if (Ns > 0x7fffffffLL || Nd > 0x7fffffffLL) return GAOT_E_BADARG;
    if (!al16(r) || !al16(u)) return GAOT_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    if (Nd == 0) return 0;
    GAOT_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)(B * Nd * Cout), st));
    if (Ns == 0 || E == 0) return 0;
    if (launch_dot_s<false>(r, u, nullptr, t_rows, t_perm, t_splits, out, (int)B, (int)Ns, (int)Nd, E, (int)C, (int)Cout, st)) GAOT_RETURN_LAST();
    const int g_ = group_of(C);
    const int sl = slices_of(Ns, g_, B);
    const int64_t n_items = Ns * sl;
    if ((n_items * g_ + 255) / 256 > 0x7fffffffLL) return GAOT_E_BADARG;
    GAOT_DISPATCH_G(g_, {
        if (Cout == 1) launch_dot_fwd_t<G, 1, 8>(r, u, t_rows, t_perm, t_splits, out, (int)B, (int)Ns, (int)Nd, E, (int)C, n_items, sl, st);
        else if (Cout == 2) launch_dot_fwd_t<G, 2, 4>(r, u, t_rows, t_perm, t_splits, out, (int)B, (int)Ns, (int)Nd, E, (int)C, n_items, sl, st);
        else if (Cout == 3) launch_dot_fwd_t<G, 3, 2>(r, u, t_rows, t_perm, t_splits, out, (int)B, (int)Ns, (int)Nd, E, (int)C, n_items, sl, st);
        else launch_dot_fwd_t<G, 4, 2>(r, u, t_rows, t_perm, t_splits, out, (int)B, (int)Ns, (int)Nd, E, (int)C, n_items, sl, st);
    });
    GAOT_RETURN_LAST();
}

// Source-walk backward: dr and du in ONE pass over the transposed graph, no gather of latent rows per edge.
GAOT_API int gaot_agno_dot_bwd_t(const float* g, const float* r, const float* u, const int64_t* index, const int64_t* row_splits,
                                 const int64_t* t_splits, const int64_t* t_rows, const int64_t* t_perm, float* dr, float* du,
                                 int64_t B, int64_t Ns, int64_t Nd, int64_t E, int64_t C, int64_t Cout, void* stream) {
    if (!g || !r || !u || !t_splits || !t_rows || !t_perm || (!dr && !du) || B < 1 || Ns < 0 || Nd < 0 || E < 0) return GAOT_E_BADARG;
    if (C % 4 || C > 128 || Cout < 1 || Cout > 4) return GAOT_E_BADARG;
    if (Ns > 0x7fffffffLL || Nd > 0x7fffffffLL) return GAOT_E_BADARG;
    if (!al16(r) || !al16(u) || (dr && !al16(dr)) || (du && !al16(du))) return GAOT_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    if (Ns == 0) return 0;
    const int g_ = group_of(C);
    if (index && row_splits && B >= 16 && C % 16 == 0 && C <= 64) {
        // batches of >= 16 samples: dr with one lane per sample on the transposed graph (each latent-row gradient is
        // formed in registers and stored once); du, the product that reduces over samples, on the row-sorted graph
        int launches = 0;
        if (dr && launch_dot_s<true>(r, u, g, t_rows, t_perm, t_splits, dr, (int)B, (int)Ns, (int)Nd, E, (int)C, (int)Cout, st)) ++launches;
        static const bool du_rows = getenv("GAOT_B200_DU_ROW_WALK") != nullptr;    // experiment switch: the row-walk du kernel
        if (du && Nd > 0 && E > 0 && (C == 32 || C == 64) && !du_rows) {
            // du with one lane per channel on the transposed graph: every edge has exactly one owner (its source), so the
            // kernel overwrites du without atomics or a zero fill
            const int64_t blocks = (Ns * 32 + 255) / 256;
            if (blocks > 0x7fffffffLL) return GAOT_E_BADARG;
            if (C == 32) { GAOT_DISPATCH_N((int)Cout, COUT, (agno_dot_du_c_kernel<1, COUT><<<(unsigned)blocks, 256, 0, st>>>(r, g, t_rows, t_perm, t_splits, du, (int)B, (int)Ns, (int)Nd, E))); }
            else { GAOT_DISPATCH_N((int)Cout, COUT, (agno_dot_du_c_kernel<2, COUT><<<(unsigned)blocks, 256, 0, st>>>(r, g, t_rows, t_perm, t_splits, du, (int)B, (int)Ns, (int)Nd, E))); }
            ++launches;
        } else if (du && Nd > 0 && E > 0) {
            const int bsplit = (int)((B + 7) / 8);
            if (bsplit > 1) GAOT_CHECK(cudaMemsetAsync(du, 0, sizeof(float) * (size_t)(Cout * E * C), st));
            const int64_t n_items = Nd * bsplit;
            const int64_t blocks = (n_items * g_ + 255) / 256;
            if (blocks > 0x7fffffffLL) return GAOT_E_BADARG;
            GAOT_DISPATCH_G(g_, GAOT_DISPATCH_N((int)Cout, COUT, (agno_dot_bwd_u_kernel<G, COUT><<<(unsigned)blocks, 256, 0, st>>>(
                                                                    g, r, index, row_splits, du, (int)B, Ns, Nd, E, (int)C, bsplit, n_items))));
            ++launches;
        }
        if (launches == 0) return 0;
        if (launches == 2) GAOT_NOTE_LAUNCHES(1);
        GAOT_RETURN_LAST();
    }
    const int sl = slices_of(Ns, g_, B);
    if (du && sl > 1 && E > 0) GAOT_CHECK(cudaMemsetAsync(du, 0, sizeof(float) * (size_t)(Cout * E * C), st));
    const int64_t n_items = Ns * sl;
    if ((n_items * g_ + 255) / 256 > 0x7fffffffLL) return GAOT_E_BADARG;
    GAOT_DISPATCH_G(g_, {
        if (Cout == 1) launch_dot_bwd_t<G, 1, 8>(g, r, u, t_rows, t_perm, t_splits, dr, du, (int)B, (int)Ns, (int)Nd, E, (int)C, n_items, sl, st);
        else if (Cout == 2) launch_dot_bwd_t<G, 2, 4>(g, r, u, t_rows, t_perm, t_splits, dr, du, (int)B, (int)Ns, (int)Nd, E, (int)C, n_items, sl, st);
        else if (Cout == 3) launch_dot_bwd_t<G, 3, 2>(g, r, u, t_rows, t_perm, t_splits, dr, du, (int)B, (int)Ns, (int)Nd, E, (int)C, n_items, sl, st);
        else launch_dot_bwd_t<G, 4, 2>(g, r, u, t_rows, t_perm, t_splits, dr, du, (int)B, (int)Ns, (int)Nd, E, (int)C, n_items, sl, st);
    });
    GAOT_RETURN_LAST();
}
