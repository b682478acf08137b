// Radius-graph construction and CSR utilities (integer outputs, bit-exact with the oracle).
// Replaces src/model/layers/utils/neighbor_search.py and the loops of src/datasets/graph_builder.py.
//
// Design: one warp owns QPW query rows; its 32 lanes sweep the data points 32 at a time in
// ascending index, so a __ballot gives the hits of the sweep already ordered - the position of a hit
// inside its row is (hits so far) + popc(ballot below my lane).  No sort, no atomics, deterministic.
// The same sweep runs twice (count, fill) because E is data dependent.
#include "common.cuh"
#include "../../include/gaot_b200.h"

namespace {

constexpr int QPW = 4;            // query rows per warp (amortises every data load over 4 rows)
constexpr int RADIUS_THREADS = 256;

template <int D>
__device__ __forceinline__ float dist2_exact(const float (&p)[D], const float (&q)[D]) {
    // products rounded, added in dimension order, no fma contraction (oracle/gaot_oracle.py radius_csr)
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const float df = __fsub_rn(p[k], q[k]);
        acc = __fadd_rn(acc, __fmul_rn(df, df));
    }
    return acc;
}

template <int D, bool FILL>
__global__ void __launch_bounds__(RADIUS_THREADS)
radius_sweep_kernel(const float* __restrict__ data, int64_t n, const float* __restrict__ queries, int64_t m,
                    float r2, int inclusive, int64_t cap, int data_batched, int queries_batched,
                    int64_t* __restrict__ row_splits, const int64_t* __restrict__ index_offsets,
                    int64_t* __restrict__ index) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t b = blockIdx.y;
    const int64_t q0 = ((int64_t)blockIdx.x * (RADIUS_THREADS / 32) + warp) * QPW;
    if (q0 >= m) return;
    const float* dat = data + (data_batched ? b * n * D : 0);
    const float* qry = queries + (queries_batched ? b * m * D : 0);
    int64_t* rs = row_splits + b * (m + 1);

    float q[QPW][D];
    int64_t run[QPW];
    int64_t base[QPW];
    bool live[QPW];
#pragma unroll
    for (int i = 0; i < QPW; ++i) {
        live[i] = (q0 + i) < m;
        const int64_t qi = live[i] ? q0 + i : q0;
#pragma unroll
        for (int k = 0; k < D; ++k) q[i][k] = __ldg(qry + qi * D + k);
        run[i] = 0;
        base[i] = 0;
        if (FILL) base[i] = index_offsets[b] + rs[qi];
    }
    const unsigned below = (1u << lane) - 1u;
    for (int64_t j0 = 0; j0 < n; j0 += 32) {
        const int64_t j = j0 + lane;
        float p[D];
        const bool inb = j < n;
#pragma unroll
        for (int k = 0; k < D; ++k) p[k] = inb ? __ldg(dat + j * D + k) : 0.f;
        bool any_open = false;
#pragma unroll
        for (int i = 0; i < QPW; ++i) {
            const float d2 = dist2_exact<D>(p, q[i]);
            const bool hit = inb && live[i] && (inclusive ? (d2 <= r2) : (d2 < r2));
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (FILL) {
                const int64_t pos = run[i] + __popc(bal & below);
                if (hit && (cap <= 0 || pos < cap)) index[base[i] + pos] = j;
            }
            run[i] += __popc(bal);
            any_open |= live[i] && (cap <= 0 || run[i] < cap);
        }
        if (!any_open) break;     // every row of this warp reached the first-k cap
    }
    if (!FILL && lane == 0) {
#pragma unroll
        for (int i = 0; i < QPW; ++i)
            if (live[i]) rs[1 + q0 + i] = (cap > 0 && run[i] > cap) ? cap : run[i];
    }
}

// In-place inclusive scan of row_splits[1..m] per sample, row_splits[0] = 0.  One CTA per sample.
__global__ void __launch_bounds__(1024) splits_scan_kernel(int64_t* __restrict__ row_splits, int64_t m) {
    __shared__ int64_t warp_tot[32];
    __shared__ int64_t carry_s;
    int64_t* rs = row_splits + (int64_t)blockIdx.x * (m + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry_s = 0; rs[0] = 0; }
    __syncthreads();
    for (int64_t s = 0; s < m; s += 1024) {
        const int64_t i = s + threadIdx.x;
        int64_t v = (i < m) ? rs[1 + i] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) warp_tot[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int64_t w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const int64_t carry = carry_s;
        const int64_t incl = v + (warp > 0 ? warp_tot[warp - 1] : 0) + carry;
        if (i < m) rs[1 + i] = incl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
}

__global__ void edge_rows_kernel(const int64_t* __restrict__ rs, int64_t m, int64_t* __restrict__ edge_row) {
    // one warp per row: rows are short, lanes write consecutive edges
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= m) return;
    const int64_t a = rs[row], e = rs[row + 1];
    for (int64_t i = a + lane; i < e; i += 32) edge_row[i] = row;
}

__global__ void count_sources_kernel(const int64_t* __restrict__ index, int64_t E, unsigned long long* __restrict__ cnt) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(cnt + 1 + index[e], 1ull);
}

// claim slots with atomics (order inside a source list is arbitrary here, fixed by the sort below)
__global__ void transpose_fill_kernel(const int64_t* __restrict__ index, const int64_t* __restrict__ rs, int64_t m,
                                      const int64_t* __restrict__ t_splits, unsigned long long* __restrict__ cursor,
                                      int64_t* __restrict__ t_perm) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= m) return;
    const int64_t a = rs[row], e = rs[row + 1];
    for (int64_t i = a + lane; i < e; i += 32) {
        const int64_t j = index[i];
        const unsigned long long slot = atomicAdd(cursor + j, 1ull);
        t_perm[t_splits[j] + (int64_t)slot] = i;
    }
}

// per source: insertion-sort its (short) list by edge id, then derive the destination rows
__global__ void transpose_sort_kernel(const int64_t* __restrict__ t_splits, int64_t n_src,
                                      const int64_t* __restrict__ rs, int64_t m,
                                      int64_t* __restrict__ t_perm, int64_t* __restrict__ t_rows) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_src) return;
    const int64_t a = t_splits[j], e = t_splits[j + 1];
    for (int64_t i = a + 1; i < e; ++i) {
        const int64_t key = t_perm[i];
        int64_t k = i - 1;
        while (k >= a && t_perm[k] > key) { t_perm[k + 1] = t_perm[k]; --k; }
        t_perm[k + 1] = key;
    }
    for (int64_t i = a; i < e; ++i) {
        // row of edge id: binary search in row_splits (largest r with rs[r] <= id)
        const int64_t id = t_perm[i];
        int64_t lo = 0, hi = m;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (rs[mid] <= id) lo = mid; else hi = mid;
        }
        t_rows[i] = lo;
    }
}

template <bool FILL>
int radius_launch(const float* data, int64_t n, const float* queries, int64_t m, int d, float radius,
                  int inclusive, int64_t max_nbrs, int64_t batch, int db, int qb, int64_t* row_splits,
                  const int64_t* index_offsets, int64_t* index, cudaStream_t st) {
    if (!data || !queries || !row_splits || n < 0 || m < 0 || batch < 1) return GAOT_E_BADARG;
    if (d != 2 && d != 3) return GAOT_E_BADARG;
    if (FILL && (!index_offsets)) return GAOT_E_BADARG;
    const float r2 = radius * radius;   // fp32 product, same rounding as the oracle
    if (m > 0) {
        const int rows_per_cta = (RADIUS_THREADS / 32) * QPW;
        dim3 grid((unsigned)((m + rows_per_cta - 1) / rows_per_cta), (unsigned)batch);
        if (d == 2)
            radius_sweep_kernel<2, FILL><<<grid, RADIUS_THREADS, 0, st>>>(data, n, queries, m, r2, inclusive, max_nbrs,
                                                                         db, qb, row_splits, index_offsets, index);
        else
            radius_sweep_kernel<3, FILL><<<grid, RADIUS_THREADS, 0, st>>>(data, n, queries, m, r2, inclusive, max_nbrs,
                                                                         db, qb, row_splits, index_offsets, index);
    }
    if (!FILL) { splits_scan_kernel<<<(unsigned)batch, 1024, 0, st>>>(row_splits, m); GAOT_NOTE_LAUNCHES(1); }
    GAOT_RETURN_LAST();
}

}  // namespace

GAOT_API int gaot_radius_count(const float* data, int64_t n, const float* queries, int64_t m, int d,
                               float radius, int inclusive, int64_t max_nbrs, int64_t batch,
                               int data_batched, int queries_batched, int64_t* row_splits, void* stream) {
    return radius_launch<false>(data, n, queries, m, d, radius, inclusive, max_nbrs, batch, data_batched,
                                queries_batched, row_splits, nullptr, nullptr, (cudaStream_t)stream);
}

GAOT_API int gaot_radius_fill(const float* data, int64_t n, const float* queries, int64_t m, int d,
                              float radius, int inclusive, int64_t max_nbrs, int64_t batch,
                              int data_batched, int queries_batched, const int64_t* row_splits,
                              const int64_t* index_offsets, int64_t* index, void* stream) {
    if (!index) return GAOT_E_BADARG;
    return radius_launch<true>(data, n, queries, m, d, radius, inclusive, max_nbrs, batch, data_batched,
                               queries_batched, const_cast<int64_t*>(row_splits), index_offsets, index,
                               (cudaStream_t)stream);
}

GAOT_API int gaot_csr_edge_rows(const int64_t* row_splits, int64_t m, int64_t* edge_row, void* stream) {
    if (!row_splits || !edge_row || m < 0) return GAOT_E_BADARG;
    if (m == 0) return 0;
    edge_rows_kernel<<<gaot_blocks(m * 32, 256), 256, 0, (cudaStream_t)stream>>>(row_splits, m, edge_row);
    GAOT_RETURN_LAST();
}

GAOT_API int gaot_csr_transpose(const int64_t* index, const int64_t* row_splits, int64_t m, int64_t n_src,
                                int64_t E, int64_t* t_splits, int64_t* t_rows, int64_t* t_perm,
                                int64_t* cursor_ws, void* stream) {
    if (!index || !row_splits || !t_splits || !t_rows || !t_perm || !cursor_ws || m < 0 || n_src < 0 || E < 0)
        return GAOT_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    // histogram into t_splits[1..], scan in place, then claim slots through per-source cursors
    GAOT_CHECK(cudaMemsetAsync(t_splits, 0, sizeof(int64_t) * (size_t)(n_src + 1), st));
    if (E > 0) count_sources_kernel<<<gaot_grid(E, 256), 256, 0, st>>>(index, E, (unsigned long long*)t_splits);
    splits_scan_kernel<<<1, 1024, 0, st>>>(t_splits, n_src);
    if (E == 0 || n_src == 0) GAOT_RETURN_LAST();
    GAOT_CHECK(cudaMemsetAsync(cursor_ws, 0, sizeof(int64_t) * (size_t)n_src, st));
    transpose_fill_kernel<<<gaot_blocks(m * 32, 256), 256, 0, st>>>(index, row_splits, m, t_splits,
                                                                     (unsigned long long*)cursor_ws, t_perm);
    transpose_sort_kernel<<<gaot_blocks(n_src, 128), 128, 0, st>>>(t_splits, n_src, row_splits, m, t_perm, t_rows);
    GAOT_NOTE_LAUNCHES(3);
    GAOT_RETURN_LAST();
}
