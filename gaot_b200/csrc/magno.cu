// MAGNO encoder/decoder kernels: edge inputs, cosine segment-softmax, the attentional
// gather * weight -> segment-sum aggregate (forward, and both backward products), the generic
// segment_csr drop-in and the statistical geometry features.
// Reference: src/model/layers/agno.py:148-273, src/model/layers/gemb.py:83-171,
//            src/model/layers/utils/segment_csr.py:14-55.
//
// Aggregate layout: features [B, N, C] fp32, channel-contiguous.  A group of G lanes (G*4 >= C, so
// one 128-bit load per lane covers a feature row) owns one destination row and BC batch samples:
// the edge weight row w[e,:] is loaded once per edge into registers and reused for the BC samples,
// every feature row is fetched with coalesced 16-byte loads, sums stay in registers in edge order
// (deterministic, no atomics) and one coalesced store writes the result.  Repeated reads of the
// same source row by neighbouring destination rows are served by L2 (the whole per-sample feature
// matrix is a few MB).
#include "common.cuh"
#include "../../include/gaot_b200.h"

namespace {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void fma4(float4& a, const float4& x, const float4& y) {
    a.x = fmaf(x.x, y.x, a.x); a.y = fmaf(x.y, y.y, a.y); a.z = fmaf(x.z, y.z, a.z); a.w = fmaf(x.w, y.w, a.w);
}

// ------------------------------------------------------------------------------------------------
// edge inputs / attention weights
// ------------------------------------------------------------------------------------------------
__global__ void edge_inputs_kernel(const float* __restrict__ y, int dy, const float* __restrict__ x, int dx,
                                   const int64_t* __restrict__ index, const int64_t* __restrict__ edge_row,
                                   int64_t E, float* __restrict__ z) {
    const int w = dy + dx;
    const int64_t total = E * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i / w;
        const int k = (int)(i - e * w);
        z[i] = (k < dy) ? __ldg(y + index[e] * dy + k) : __ldg(x + edge_row[e] * dx + (k - dy));
    }
}

// one thread per destination row: rows are short (mean degree 3..50) and the whole op touches E scalars
__global__ void cosine_softmax_kernel(const float* __restrict__ y, int dy, const float* __restrict__ x, int dx, int cd,
                                      const int64_t* __restrict__ index, const int64_t* __restrict__ rs, int64_t m,
                                      float* __restrict__ alpha) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= m) return;
    const int64_t a = rs[row], e = rs[row + 1];
    if (a == e) return;
    float qn = 0.f;
    for (int k = 0; k < cd; ++k) { const float v = x[row * dx + k]; qn = fmaf(v, v, qn); }
    const float qinv = 1.f / fmaxf(sqrtf(qn), 1e-12f);
    float smax = -INFINITY;
    for (int64_t i = a; i < e; ++i) {
        const float* yp = y + index[i] * dy;
        float kn = 0.f, dot = 0.f;
        for (int k = 0; k < cd; ++k) {
            const float kv = __ldg(yp + k);
            kn = fmaf(kv, kv, kn);
            dot = fmaf(kv, x[row * dx + k] * qinv, dot);
        }
        const float s = dot / fmaxf(sqrtf(kn), 1e-12f);
        alpha[i] = s;
        smax = fmaxf(smax, s);
    }
    float den = 0.f;
    for (int64_t i = a; i < e; ++i) { const float ex = expf(alpha[i] - smax); alpha[i] = ex; den += ex; }
    const float inv = 1.f / den;
    for (int64_t i = a; i < e; ++i) alpha[i] *= inv;
}

// ---- dot-product attention weights (agno.py:214-217): one warp per destination row, A-wide projections ----
// alpha[e] = softmax_row( scale * <q[row], k[index[e]]> ); lanes own the columns l, l+32, ... of the projections
__global__ void __launch_bounds__(128) dot_softmax_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, int A,
                                                              const int64_t* __restrict__ index, const int64_t* __restrict__ rs,
                                                              int64_t m, float scale, float* __restrict__ alpha) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= m) return;
    const int64_t a = rs[row], e = rs[row + 1];
    if (a == e) return;
    float smax = -INFINITY;
    for (int64_t i = a; i < e; ++i) {
        const float* kp = k + index[i] * A;
        float dot = 0.f;
        for (int c = lane; c < A; c += 32) dot = fmaf(__ldg(q + row * A + c), __ldg(kp + c), dot);
        const float s = warp_sum(dot) * scale;
        if (lane == 0) alpha[i] = s;
        smax = fmaxf(smax, s);
    }
    __syncwarp();
    float den = 0.f;
    for (int64_t i = a + lane; i < e; i += 32) { const float ex = expf(alpha[i] - smax); alpha[i] = ex; den += ex; }
    den = warp_sum(den);
    const float inv = 1.f / den;
    for (int64_t i = a + lane; i < e; i += 32) alpha[i] *= inv;
}

// backward of the above: ds_e = alpha_e (dalpha_e - sum_row alpha dalpha);  dq[row] = scale sum_e ds_e k[idx_e] (row owned),
// dk[idx_e] += scale ds_e q[row] (red.add: a source point belongs to several rows)
__global__ void __launch_bounds__(128) dot_softmax_bwd_kernel(const float* __restrict__ dalpha, const float* __restrict__ alpha,
                                                              const float* __restrict__ q, const float* __restrict__ k, int A,
                                                              const int64_t* __restrict__ index, const int64_t* __restrict__ rs,
                                                              int64_t m, float scale, float* __restrict__ dq, float* __restrict__ dk) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= m) return;
    const int64_t a = rs[row], e = rs[row + 1];
    float t = 0.f;
    for (int64_t i = a + lane; i < e; i += 32) t = fmaf(alpha[i], dalpha[i], t);
    t = warp_sum(t);
    for (int c = lane; c < A; c += 32) {
        const float qv = __ldg(q + row * A + c);
        float acc = 0.f;
        for (int64_t i = a; i < e; ++i) {
            const float ds = alpha[i] * (dalpha[i] - t) * scale;
            const int64_t j = index[i];
            acc = fmaf(ds, __ldg(k + j * A + c), acc);
            atomicAdd(dk + j * A + c, ds * qv);
        }
        dq[row * A + c] = acc;
    }
}

// out[r] = sum_c a[r,c] * b[r,c]   (gradient of the row scale in scale_rows); one warp per row
__global__ void __launch_bounds__(256) row_dot_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t rows, int C,
                                                      float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nw) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(__ldg(a + r * C + c), __ldg(b + r * C + c), s);
        s = warp_sum(s);
        if (lane == 0) out[r] = s;
    }
}

// z[e,k] = y[index[e], k] - x[edge_row[e], k], k < d   (neighbour coordinates centred on their query, gemb.py:212-213)
__global__ void edge_diff_kernel(const float* __restrict__ y, int dy, const float* __restrict__ x, int dx, int d,
                                 const int64_t* __restrict__ index, const int64_t* __restrict__ edge_row, int64_t E,
                                 float* __restrict__ z) {
    const int64_t total = E * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i / d;
        const int k = (int)(i % d);
        z[i] = __ldg(y + index[e] * dy + k) - __ldg(x + edge_row[e] * dx + k);
    }
}

// backward of the segment reduction out[r,:] = reduce_{e in row r} src[e,:]: sum -> dout[row], mean -> dout[row] / deg,
// max -> dout[row] at the FIRST edge of the row that attains the maximum (torch_scatter's argmax), 0 elsewhere
__global__ void segment_csr_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ src, const float* __restrict__ out,
                                       const int64_t* __restrict__ rs, const int64_t* __restrict__ edge_row, float* __restrict__ dsrc,
                                       int64_t E, int C, int reduce) {
    const int64_t total = E * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i / C;
        const int c = (int)(i % C);
        const int64_t row = edge_row[e], a = rs[row];
        float g = dout[row * C + c];
        if (reduce == 1) g /= (float)(rs[row + 1] - a);
        else if (reduce == 2) {
            const float mx = out[row * C + c];
            bool mine = src[i] == mx;
            for (int64_t j = a; mine && j < e; ++j) mine = src[j * C + c] != mx;
            if (!mine) g = 0.f;
        }
        dsrc[i] = g;
    }
}

__global__ void scale_rows_kernel(const float* __restrict__ in, const float* __restrict__ scale, int64_t rows, int C,
                                  float* __restrict__ out) {
    const int64_t total = rows * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i] * __ldg(scale + i / C);
}

// ------------------------------------------------------------------------------------------------
// aggregate: out[b,r,:] = sum_{e in row r} w[p(e),:] * f[b,index[e],:]
// ------------------------------------------------------------------------------------------------
template <int G, int VPL, int BC>
__global__ void __launch_bounds__(256)
agno_aggregate_kernel(const float* __restrict__ f, const float* __restrict__ w, const int64_t* __restrict__ w_perm,
                      const int64_t* __restrict__ index, const int64_t* __restrict__ rs,
                      float* __restrict__ out, int B, int64_t Ns, int64_t Nd, int C, int64_t n_items) {
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int gl = threadIdx.x % G;
    if (gid >= n_items) return;
    const int64_t row = gid % Nd;
    const int b0 = (int)(gid / Nd) * BC;
    const int nvec = C >> 2;
    // Branch-free inner loop: out-of-range batch slots / channel lanes read a valid (clamped) address and
    // are dropped at the store, so the BC+1 loads of an edge issue back to back and stay in flight together.
    int64_t boff[BC];
#pragma unroll
    for (int bb = 0; bb < BC; ++bb) boff[bb] = (int64_t)min(b0 + bb, B - 1) * Ns * C;
    int coff[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) coff[v] = min(gl + v * G, nvec - 1) * 4;
    float4 acc[BC][VPL];
#pragma unroll
    for (int bb = 0; bb < BC; ++bb)
#pragma unroll
        for (int v = 0; v < VPL; ++v) acc[bb][v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t a = __ldg(rs + row), e = __ldg(rs + row + 1);
    if (a < e) {
        // software pipeline: the index (and permuted weight row id) of edge i+1 is fetched while edge i's
        // feature rows are in flight
        int64_t j = __ldg(index + a);
        int64_t pe = w_perm ? __ldg(w_perm + a) : a;
        for (int64_t i = a; i < e; ++i) {
            const int64_t inext = (i + 1 < e) ? i + 1 : i;
            const int64_t jn = __ldg(index + inext);
            const int64_t pn = w_perm ? __ldg(w_perm + inext) : inext;
            float4 wv[VPL], fv[BC][VPL];
#pragma unroll
            for (int v = 0; v < VPL; ++v) wv[v] = ldg4(w + pe * C + coff[v]);
#pragma unroll
            for (int bb = 0; bb < BC; ++bb)
#pragma unroll
                for (int v = 0; v < VPL; ++v) fv[bb][v] = ldg4(f + boff[bb] + j * C + coff[v]);
#pragma unroll
            for (int bb = 0; bb < BC; ++bb)
#pragma unroll
                for (int v = 0; v < VPL; ++v) fma4(acc[bb][v], wv[v], fv[bb][v]);
            j = jn; pe = pn;
        }
    }
#pragma unroll
    for (int bb = 0; bb < BC; ++bb) {
        if (b0 + bb < B) {
            float* op = out + ((int64_t)(b0 + bb) * Nd + row) * C;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c4 = gl + v * G;
                if (c4 < nvec) *reinterpret_cast<float4*>(op + c4 * 4) = acc[bb][v];
            }
        }
    }
}

// any C (not a multiple of 4, or very wide): one thread per (b, row, c)
__global__ void agno_aggregate_scalar_kernel(const float* __restrict__ f, const float* __restrict__ w,
                                             const int64_t* __restrict__ w_perm, const int64_t* __restrict__ index,
                                             const int64_t* __restrict__ rs, float* __restrict__ out,
                                             int64_t B, int64_t Ns, int64_t Nd, int64_t C) {
    const int64_t total = B * Nd * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t % C, row = (t / C) % Nd, b = t / (C * Nd);
        float acc = 0.f;
        for (int64_t i = rs[row]; i < rs[row + 1]; ++i) {
            const int64_t pe = w_perm ? w_perm[i] : i;
            acc = fmaf(__ldg(w + pe * C + c), __ldg(f + (b * Ns + index[i]) * C + c), acc);
        }
        out[t] = acc;
    }
}

template <int G, int VPL>
int launch_aggregate_bc(const float* f, const float* w, const int64_t* w_perm, const int64_t* index, const int64_t* rs,
                        float* out, int64_t B, int64_t Ns, int64_t Nd, int64_t C, cudaStream_t st) {
    const int bc = B >= 8 ? 8 : (B >= 4 ? 4 : (B >= 2 ? 2 : 1));
    const int64_t n_items = Nd * ((B + bc - 1) / bc);
    const int threads = 256;
    const int64_t blocks = (n_items * G + threads - 1) / threads;
    if (blocks > 0x7fffffffLL) return GAOT_E_BADARG;
#define GAOT_AGG(BC_) agno_aggregate_kernel<G, VPL, BC_><<<(unsigned)blocks, threads, 0, st>>>( \
        f, w, w_perm, index, rs, out, (int)B, Ns, Nd, (int)C, n_items)
    switch (bc) {
        case 8: GAOT_AGG(8); break;
        case 4: GAOT_AGG(4); break;
        case 2: GAOT_AGG(2); break;
        default: GAOT_AGG(1); break;
    }
#undef GAOT_AGG
    GAOT_RETURN_LAST();
}

// ------------------------------------------------------------------------------------------------
// weight gradient: dw[e,:] = sum_b g[b,row(e),:] * f[b,index[e],:]
// A group of G lanes owns one destination row and a slice of the batch; it keeps EC edge
// accumulators in registers so g[b,row,:] is loaded once per EC edges.  With more than one batch
// slice the partial sums are combined with 128-bit red.add.
// ------------------------------------------------------------------------------------------------
template <int G, int VPL, int EC>
__global__ void __launch_bounds__(256)
agno_weight_grad_kernel(const float* __restrict__ g, const float* __restrict__ f, const int64_t* __restrict__ index,
                        const int64_t* __restrict__ rs, float* __restrict__ dw,
                        int B, int64_t Ns, int64_t Nd, int C, int bsplit, int64_t n_items) {
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int gl = threadIdx.x % G;
    if (gid >= n_items) return;
    const int64_t row = gid % Nd;
    const int slice = (int)(gid / Nd);
    const int per = (B + bsplit - 1) / bsplit;
    const int b_lo = slice * per, b_hi = min(B, b_lo + per);
    const int nvec = C >> 2;
    int coff[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) coff[v] = min(gl + v * G, nvec - 1) * 4;
    const int64_t a = __ldg(rs + row), e = __ldg(rs + row + 1);
    for (int64_t i0 = a; i0 < e; i0 += EC) {
        int64_t j[EC];                      // edges past the row end alias the last valid edge (dropped at the store)
#pragma unroll
        for (int k = 0; k < EC; ++k) j[k] = __ldg(index + min(i0 + k, e - 1)) * C;
        float4 acc[EC][VPL];
#pragma unroll
        for (int k = 0; k < EC; ++k)
#pragma unroll
            for (int v = 0; v < VPL; ++v) acc[k][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int b = b_lo; b < b_hi; ++b) {
            float4 gv[VPL], fv[EC][VPL];
            const float* gp = g + ((int64_t)b * Nd + row) * C;
            const float* fb = f + (int64_t)b * Ns * C;
#pragma unroll
            for (int v = 0; v < VPL; ++v) gv[v] = ldg4(gp + coff[v]);
#pragma unroll
            for (int k = 0; k < EC; ++k)
#pragma unroll
                for (int v = 0; v < VPL; ++v) fv[k][v] = ldg4(fb + j[k] + coff[v]);
#pragma unroll
            for (int k = 0; k < EC; ++k)
#pragma unroll
                for (int v = 0; v < VPL; ++v) fma4(acc[k][v], gv[v], fv[k][v]);
        }
#pragma unroll
        for (int k = 0; k < EC; ++k) {
            if (i0 + k < e) {
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    const int c4 = gl + v * G;
                    if (c4 < nvec) {
                        float4* dst = reinterpret_cast<float4*>(dw + (i0 + k) * C + c4 * 4);
                        if (bsplit > 1) atomicAdd(dst, acc[k][v]);
                        else *dst = acc[k][v];
                    }
                }
            }
        }
    }
}

__global__ void agno_weight_grad_scalar_kernel(const float* __restrict__ g, const float* __restrict__ f,
                                               const int64_t* __restrict__ index, const int64_t* __restrict__ edge_rows_unused,
                                               const int64_t* __restrict__ rs, float* __restrict__ dw,
                                               int64_t B, int64_t Ns, int64_t Nd, int64_t C) {
    // one thread per (row, c): walks the row's edges; tiny-C fallback only
    const int64_t total = Nd * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t % C, row = t / C;
        for (int64_t i = rs[row]; i < rs[row + 1]; ++i) {
            const int64_t jj = index[i];
            float acc = 0.f;
            for (int64_t b = 0; b < B; ++b)
                acc = fmaf(__ldg(g + (b * Nd + row) * C + c), __ldg(f + (b * Ns + jj) * C + c), acc);
            dw[i * C + c] = acc;
        }
    }
}

template <int G, int VPL>
int launch_weight_grad(const float* g, const float* f, const int64_t* index, const int64_t* rs, float* dw,
                       int64_t B, int64_t Ns, int64_t Nd, int64_t C, int64_t E, cudaStream_t st) {
    // enough groups to fill the chip: target >= 148 SMs * 2048 threads
    int bsplit = 1;
    const int64_t want = (int64_t)gaot_sm_count() * 2048;
    while (bsplit < B && Nd * G * bsplit < want && bsplit < 8) bsplit <<= 1;
    if (bsplit > B) bsplit = (int)B;
    if (bsplit > 1) GAOT_CHECK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)(E * C), st));
    const int64_t n_items = Nd * bsplit;
    const int threads = 256;
    const int64_t blocks = (n_items * G + threads - 1) / threads;
    if (blocks > 0x7fffffffLL) return GAOT_E_BADARG;
    agno_weight_grad_kernel<G, VPL, 4><<<(unsigned)blocks, threads, 0, st>>>(g, f, index, rs, dw, (int)B, Ns, Nd,
                                                                             (int)C, bsplit, n_items);
    GAOT_RETURN_LAST();
}

// ------------------------------------------------------------------------------------------------
// generic segment_csr (drop-in API; not on the fused hot path)
// ------------------------------------------------------------------------------------------------
__global__ void segment_csr_kernel(const float* __restrict__ src, const int64_t* __restrict__ rs, float* __restrict__ out,
                                   int64_t B, int64_t E, int64_t m, int64_t C, int reduce) {
    const int64_t total = B * m * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t % C, row = (t / C) % m, b = t / (C * m);
        const int64_t a = rs[row], e = rs[row + 1];
        float acc = (reduce == 2) ? -INFINITY : 0.f;
        for (int64_t i = a; i < e; ++i) {
            const float v = __ldg(src + (b * E + i) * C + c);
            acc = (reduce == 2) ? fmaxf(acc, v) : acc + v;
        }
        if (a == e) acc = 0.f;
        else if (reduce == 1) acc /= (float)(e - a);
        out[t] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// statistical geometry features (fp64 inside: the variance is a difference of moments)
// ------------------------------------------------------------------------------------------------
__device__ void eig_sym2(double a, double b, double c, double* ev) {   // [[a,b],[b,c]] -> descending
    const double mid = 0.5 * (a + c), h = 0.5 * (a - c);
    const double r = sqrt(h * h + b * b);
    ev[0] = mid + r; ev[1] = mid - r;
}
__device__ void eig_sym3(double a00, double a01, double a02, double a11, double a12, double a22, double* ev) {
    const double p1 = a01 * a01 + a02 * a02 + a12 * a12;
    const double q = (a00 + a11 + a22) / 3.0;
    const double p2 = (a00 - q) * (a00 - q) + (a11 - q) * (a11 - q) + (a22 - q) * (a22 - q) + 2.0 * p1;
    if (p2 <= 0.0) { ev[0] = ev[1] = ev[2] = q; return; }
    const double p = sqrt(p2 / 6.0);
    const double ip = 1.0 / p;
    const double b00 = (a00 - q) * ip, b11 = (a11 - q) * ip, b22 = (a22 - q) * ip;
    const double b01 = a01 * ip, b02 = a02 * ip, b12 = a12 * ip;
    double r = 0.5 * (b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02));
    r = fmin(1.0, fmax(-1.0, r));
    const double phi = acos(r) / 3.0;
    const double e0 = q + 2.0 * p * cos(phi);
    const double e2 = q + 2.0 * p * cos(phi + 2.0943951023931954923);
    ev[0] = e0; ev[1] = 3.0 * q - e0 - e2; ev[2] = e2;
}

template <int D>
__global__ void geo_stats_kernel(const float* __restrict__ src, const float* __restrict__ qry,
                                 const int64_t* __restrict__ index, const int64_t* __restrict__ rs, int64_t m,
                                 float* __restrict__ feat) {
    constexpr int F = 3 + 2 * D;
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= m) return;
    float* o = feat + row * F;
    const int64_t a = rs[row], e = rs[row + 1];
    if (a == e) {
#pragma unroll
        for (int k = 0; k < F; ++k) o[k] = 0.f;
        return;
    }
    double q[D], sp[D], sd = 0.0, sd2 = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) { q[k] = (double)qry[row * D + k]; sp[k] = 0.0; }
    for (int64_t i = a; i < e; ++i) {
        const float* pp = src + index[i] * D;
        double d2 = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) { const double pv = (double)__ldg(pp + k); sp[k] += pv; const double df = pv - q[k]; d2 += df * df; }
        sd += sqrt(d2); sd2 += d2;
    }
    const double n = (double)(e - a), inv = 1.0 / n;
    const double davg = sd * inv;
    double dvar = sd2 * inv - davg * davg;
    if (dvar < 0.0) dvar = 0.0;
    double cen[D];
#pragma unroll
    for (int k = 0; k < D; ++k) cen[k] = sp[k] * inv;
    double cov[D][D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int c = 0; c < D; ++c) cov[r][c] = 0.0;
    for (int64_t i = a; i < e; ++i) {
        const float* pp = src + index[i] * D;
        double ct[D];
#pragma unroll
        for (int k = 0; k < D; ++k) ct[k] = (double)__ldg(pp + k) - cen[k];
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
            for (int c = r; c < D; ++c) cov[r][c] += ct[r] * ct[c];
    }
    double ev[D];
    if (D == 2) eig_sym2(cov[0][0] * inv, cov[0][1] * inv, cov[1][1] * inv, ev);
    else eig_sym3(cov[0][0] * inv, cov[0][1] * inv, cov[0][D - 1] * inv, cov[1][1] * inv, cov[1][D - 1] * inv,
                  cov[D - 1][D - 1] * inv, ev);
    o[0] = (float)n; o[1] = (float)davg; o[2] = (float)dvar;
#pragma unroll
    for (int k = 0; k < D; ++k) { o[3 + k] = (float)(cen[k] - q[k]); o[3 + D + k] = (float)ev[k]; }
}

// one CTA per column: mean and unbiased std in fp64, then in-place standardisation
__global__ void __launch_bounds__(1024) col_standardize_kernel(float* __restrict__ feat, int64_t m, int F) {
    __shared__ double red[32];
    __shared__ double bc;
    const int col = blockIdx.x;
    feat += (int64_t)blockIdx.y * m * F;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto block_sum = [&](double v) -> double {
        v = warp_sum_d(v);
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        if (warp == 0) {
            double t = red[lane];
            t = warp_sum_d(t);
            if (lane == 0) bc = t;
        }
        __syncthreads();
        return bc;
    };
    double s = 0.0;
    for (int64_t r = threadIdx.x; r < m; r += blockDim.x) s += (double)feat[r * F + col];
    const double mean = block_sum(s) / (double)m;
    double ss = 0.0;
    for (int64_t r = threadIdx.x; r < m; r += blockDim.x) { const double dv = (double)feat[r * F + col] - mean; ss += dv * dv; }
    double sd = sqrt(block_sum(ss) / (double)(m - 1));
    if (sd < 1e-6) sd = 1.0;
    const double inv = 1.0 / sd;
    for (int64_t r = threadIdx.x; r < m; r += blockDim.x) feat[r * F + col] = (float)(((double)feat[r * F + col] - mean) * inv);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

GAOT_API int gaot_edge_inputs(const float* y, int64_t dy, const float* x, int64_t dx, const int64_t* index,
                              const int64_t* edge_row, int64_t E, float* z, void* stream) {
    if (!y || !x || !index || !edge_row || !z || dy < 1 || dx < 1 || E < 0) return GAOT_E_BADARG;
    if (E == 0) return 0;
    edge_inputs_kernel<<<gaot_grid(E * (dy + dx), 256), 256, 0, (cudaStream_t)stream>>>(y, (int)dy, x, (int)dx, index,
                                                                                        edge_row, E, z);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_edge_cosine_softmax(const float* y, int64_t dy, const float* x, int64_t dx, int64_t cd,
                                      const int64_t* index, const int64_t* row_splits, int64_t m, float* alpha,
                                      void* stream) {
    if (!y || !x || !index || !row_splits || !alpha || cd < 1 || cd > dy || cd > dx || m < 0) return GAOT_E_BADARG;
    if (m == 0) return 0;
    cosine_softmax_kernel<<<gaot_blocks(m, 128), 128, 0, (cudaStream_t)stream>>>(y, (int)dy, x, (int)dx, (int)cd, index,
                                                                                 row_splits, m, alpha);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_edge_dot_softmax_fwd(const float* q, const float* k, int64_t A, const int64_t* index, const int64_t* row_splits,
                                       int64_t m, float scale, float* alpha, void* stream) {
    // `index` / `alpha` may be null when the graph has no edge at all (every row empty): nothing is dereferenced then
    if (!q || !k || !row_splits || A < 1 || m < 0) return GAOT_E_BADARG;
    if (m == 0) return 0;
    dot_softmax_fwd_kernel<<<gaot_blocks(m * 32, 128), 128, 0, (cudaStream_t)stream>>>(q, k, (int)A, index, row_splits, m, scale, alpha);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_edge_dot_softmax_bwd(const float* dalpha, const float* alpha, const float* q, const float* k, int64_t A,
                                       const int64_t* index, const int64_t* row_splits, int64_t m, int64_t n, float scale,
                                       float* dq, float* dk, void* stream) {
    if (!q || !k || !row_splits || !dq || !dk || A < 1 || m < 0 || n < 0) return GAOT_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (n > 0) GAOT_CHECK(cudaMemsetAsync(dk, 0, sizeof(float) * (size_t)(n * A), st));
    if (m == 0) return 0;
    dot_softmax_bwd_kernel<<<gaot_blocks(m * 32, 128), 128, 0, st>>>(dalpha, alpha, q, k, (int)A, index, row_splits, m, scale, dq, dk);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_row_dot(const float* a, const float* b, int64_t rows, int64_t C, float* out, void* stream) {
    if (!a || !b || !out || rows < 0 || C < 1) return GAOT_E_BADARG;
    if (rows == 0) return 0;
    row_dot_kernel<<<gaot_grid(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(a, b, rows, (int)C, out);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_edge_diff(const float* y, int64_t dy, const float* x, int64_t dx, int64_t d, const int64_t* index,
                            const int64_t* edge_row, int64_t E, float* z, void* stream) {
    if (!y || !x || d < 1 || d > dy || d > dx || E < 0) return GAOT_E_BADARG;
    if (E == 0) return 0;
    if (!index || !edge_row || !z) return GAOT_E_BADARG;
    edge_diff_kernel<<<gaot_grid(E * d, 256), 256, 0, (cudaStream_t)stream>>>(y, (int)dy, x, (int)dx, (int)d, index, edge_row, E, z);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_segment_csr_bwd(const float* dout, const float* src, const float* out, const int64_t* row_splits,
                                  const int64_t* edge_row, float* dsrc, int64_t E, int64_t m, int64_t C, int reduce, void* stream) {
    if (E < 0 || m < 0 || C < 1 || reduce < 0 || reduce > 2) return GAOT_E_BADARG;
    if (E == 0) return 0;
    if (!dout || !row_splits || !edge_row || !dsrc || (reduce == 2 && (!src || !out))) return GAOT_E_BADARG;
    segment_csr_bwd_kernel<<<gaot_grid(E * C, 256), 256, 0, (cudaStream_t)stream>>>(dout, src, out, row_splits, edge_row, dsrc, E, (int)C,
                                                                                    reduce);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_scale_rows(const float* in, const float* scale, int64_t rows, int64_t C, float* out, void* stream) {
    if (!in || !scale || !out || rows < 0 || C < 1) return GAOT_E_BADARG;
    if (rows == 0) return 0;
    scale_rows_kernel<<<gaot_grid(rows * C, 256), 256, 0, (cudaStream_t)stream>>>(in, scale, rows, (int)C, out);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_agno_aggregate(const float* f, const float* w, const int64_t* w_perm, const int64_t* index,
                                 const int64_t* row_splits, float* out, int64_t B, int64_t Ns, int64_t Nd, int64_t C,
                                 void* stream) {
    if (!f || !w || !index || !row_splits || !out || B < 1 || Ns < 0 || Nd < 0 || C < 1) return GAOT_E_BADARG;
    if (Nd == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (C % 4 == 0) && C <= 512 && aligned16(f) && aligned16(w) && aligned16(out);
    if (!vec) {
        agno_aggregate_scalar_kernel<<<gaot_grid(B * Nd * C, 256), 256, 0, st>>>(f, w, w_perm, index, row_splits, out, B,
                                                                                 Ns, Nd, C);
        GAOT_RETURN_LAST();
    }
    const int nvec = (int)(C / 4);
    if (nvec <= 4) return launch_aggregate_bc<4, 1>(f, w, w_perm, index, row_splits, out, B, Ns, Nd, C, st);
    if (nvec <= 8) return launch_aggregate_bc<8, 1>(f, w, w_perm, index, row_splits, out, B, Ns, Nd, C, st);
    if (nvec <= 16) return launch_aggregate_bc<16, 1>(f, w, w_perm, index, row_splits, out, B, Ns, Nd, C, st);
    if (nvec <= 32) return launch_aggregate_bc<32, 1>(f, w, w_perm, index, row_splits, out, B, Ns, Nd, C, st);
    if (nvec <= 64) return launch_aggregate_bc<32, 2>(f, w, w_perm, index, row_splits, out, B, Ns, Nd, C, st);
    return launch_aggregate_bc<32, 4>(f, w, w_perm, index, row_splits, out, B, Ns, Nd, C, st);
}

GAOT_API int gaot_agno_weight_grad(const float* g, const float* f, const int64_t* index, const int64_t* row_splits,
                                   float* dw, int64_t B, int64_t Ns, int64_t Nd, int64_t E, int64_t C, void* stream) {
    if (!g || !f || !index || !row_splits || !dw || B < 1 || Ns < 0 || Nd < 0 || E < 0 || C < 1) return GAOT_E_BADARG;
    if (Nd == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (C % 4 == 0) && C <= 512 && aligned16(f) && aligned16(g) && aligned16(dw);
    if (!vec) {
        agno_weight_grad_scalar_kernel<<<gaot_grid(Nd * C, 128), 128, 0, st>>>(g, f, index, nullptr, row_splits, dw, B, Ns,
                                                                               Nd, C);
        GAOT_RETURN_LAST();
    }
    const int nvec = (int)(C / 4);
    if (nvec <= 4) return launch_weight_grad<4, 1>(g, f, index, row_splits, dw, B, Ns, Nd, C, E, st);
    if (nvec <= 8) return launch_weight_grad<8, 1>(g, f, index, row_splits, dw, B, Ns, Nd, C, E, st);
    if (nvec <= 16) return launch_weight_grad<16, 1>(g, f, index, row_splits, dw, B, Ns, Nd, C, E, st);
    if (nvec <= 32) return launch_weight_grad<32, 1>(g, f, index, row_splits, dw, B, Ns, Nd, C, E, st);
    if (nvec <= 64) return launch_weight_grad<32, 2>(g, f, index, row_splits, dw, B, Ns, Nd, C, E, st);
    return launch_weight_grad<32, 4>(g, f, index, row_splits, dw, B, Ns, Nd, C, E, st);
}

GAOT_API int gaot_segment_csr(const float* src, const int64_t* row_splits, float* out, int64_t B, int64_t E, int64_t m,
                              int64_t C, int reduce, void* stream) {
    if (!src || !row_splits || !out || B < 1 || E < 0 || m < 0 || C < 1 || reduce < 0 || reduce > 2) return GAOT_E_BADARG;
    if (m == 0) return 0;
    segment_csr_kernel<<<gaot_grid(B * m * C, 256), 256, 0, (cudaStream_t)stream>>>(src, row_splits, out, B, E, m, C, reduce);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_geo_stats(const float* src, const float* queries, int64_t d, const int64_t* index,
                            const int64_t* row_splits, int64_t m, float* feat, void* stream) {
    if (!src || !queries || !index || !row_splits || !feat || m < 0) return GAOT_E_BADARG;
    if (d != 2 && d != 3) return GAOT_E_BADARG;
    if (m == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (d == 2) geo_stats_kernel<2><<<gaot_blocks(m, 128), 128, 0, st>>>(src, queries, index, row_splits, m, feat);
    else geo_stats_kernel<3><<<gaot_blocks(m, 128), 128, 0, st>>>(src, queries, index, row_splits, m, feat);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_col_standardize(float* feat, int64_t m, int64_t F, int64_t groups, void* stream) {
    if (!feat || m < 1 || F < 1 || groups < 1 || groups > 65535) return GAOT_E_BADARG;
    col_standardize_kernel<<<dim3((unsigned)F, (unsigned)groups), 1024, 0, (cudaStream_t)stream>>>(feat, m, (int)F);
    GAOT_RETURN_LAST();
}
