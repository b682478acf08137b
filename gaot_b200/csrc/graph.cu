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

// ---------------------------------------------------------------------------------------------
// Cell-binned radius search: same graph, same canon (set of j with fp32 d2 <= r2, ascending j) as the sweep above, in
// O(candidates of the 3^d cells around a query) instead of O(n) per query.  The data points are binned into a uniform
// grid of pitch h >= 1.001 r (the 0.1 % margin dwarfs the fp32 rounding of the cell coordinate, so two points within r
// always sit in the same or in adjacent cells), counting-sorted by cell and sorted by index inside a cell; a thread per
// query then tests the points of the 3^d neighbouring cells with the SAME fp32 distance arithmetic and finally
// insertion-sorts its (short) row into ascending index.  Reference counterpart: the pure-Python `grid` back-end
// (neighbor_search.py:183-272), which returns the same sets in cell order.
// ---------------------------------------------------------------------------------------------
struct GridGeom {
    float lo[3];
    float inv_h;
    int g[3];          // cells per dimension (g[2] = 1 for d = 2)
    int ncell;
};

template <int D>
__device__ __forceinline__ void cell_coords(const float (&p)[D], const GridGeom& G, int (&c)[D]) {
#pragma unroll
    for (int k = 0; k < D; ++k) c[k] = (int)floorf((p[k] - G.lo[k]) * G.inv_h);
}

template <int D>
__global__ void grid_hist_kernel(const float* __restrict__ data, int64_t n, int data_batched, GridGeom G, int* __restrict__ cnt,
                                 int* __restrict__ cell_of) {
    const int64_t b = blockIdx.y;
    const float* dat = data + (data_batched ? b * n * D : 0);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        float p[D];
        int c[D];
#pragma unroll
        for (int k = 0; k < D; ++k) p[k] = __ldg(dat + j * D + k);
        cell_coords<D>(p, G, c);
        int cell = 0;
#pragma unroll
        for (int k = 0; k < D; ++k) cell = cell * G.g[k] + min(max(c[k], 0), G.g[k] - 1);
        cell_of[b * n + j] = cell;
        atomicAdd(cnt + b * (G.ncell + 1) + 1 + cell, 1);
    }
}

// in-place inclusive scan of cnt[1..ncell] per sample (cnt[0] = 0): cell_start.  One CTA per sample.
__global__ void __launch_bounds__(1024) scan_i32_kernel(int* __restrict__ cnt, int ncell) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    int* c = cnt + (int64_t)blockIdx.x * (ncell + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry_s = 0; c[0] = 0; }
    __syncthreads();
    for (int s = 0; s < ncell; s += 1024) {
        const int i = s + threadIdx.x;
        int v = (i < ncell) ? c[1 + i] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) warp_tot[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const int incl = v + (warp > 0 ? warp_tot[warp - 1] : 0) + carry_s;
        if (i < ncell) c[1 + i] = incl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
}

__global__ void grid_scatter_kernel(const int* __restrict__ cell_of, int64_t n, int ncell, const int* __restrict__ cell_start,
                                    int* __restrict__ cursor, int* __restrict__ sorted) {
    const int64_t b = blockIdx.y;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const int cell = cell_of[b * n + j];
        const int slot = atomicAdd(cursor + b * ncell + cell, 1);
        sorted[b * n + cell_start[b * (ncell + 1) + cell] + slot] = (int)j;
    }
}

// ascending point index inside every cell (slots were claimed in arbitrary order)
__global__ void grid_sort_cells_kernel(const int* __restrict__ cell_start, int ncell, int64_t n, int* __restrict__ sorted) {
    const int64_t b = blockIdx.y;
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= ncell) return;
    int* lst = sorted + b * n;
    const int a = cell_start[b * (ncell + 1) + cell], e = cell_start[b * (ncell + 1) + cell + 1];
    for (int i = a + 1; i < e; ++i) {
        const int key = lst[i];
        int k = i - 1;
        while (k >= a && lst[k] > key) { lst[k + 1] = lst[k]; --k; }
        lst[k + 1] = key;
    }
}

// One WARP per query.  Cells that differ only in the last coordinate are neighbours in the cell order, so the points of the
// three cells (c_last - 1 .. c_last + 1) of a (dx[, dy]) column are ONE contiguous run of the cell-sorted point list: the
// warp sweeps each run 32 points at a time, a ballot orders the hits of a sweep, and after the 3^(d-1) runs the row (a few
// ascending runs) is rank-sorted in registers into ascending point index.
template <int D, bool FILL>
__global__ void __launch_bounds__(128)
radius_grid_kernel(const float* __restrict__ data, int64_t n, const float* __restrict__ queries, int64_t m, float r2, int inclusive,
                   int data_batched, int queries_batched, GridGeom G, const int* __restrict__ cell_start, const int* __restrict__ sorted,
                   int64_t* __restrict__ row_splits, const int64_t* __restrict__ index_offsets, int64_t* __restrict__ index) {
    const int64_t b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int64_t qi = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= m) return;                               // whole warps leave together
    const float* dat = data + (data_batched ? b * n * D : 0);
    const float* qry = queries + (queries_batched ? b * m * D : 0);
    const int* cs = cell_start + (data_batched ? b * (G.ncell + 1) : 0);
    const int* srt = sorted + (data_batched ? b * n : 0);
    int64_t* rs = row_splits + b * (m + 1);
    float q[D];
    int c[D];
#pragma unroll
    for (int k = 0; k < D; ++k) q[k] = __ldg(qry + qi * D + k);
    cell_coords<D>(q, G, c);
    int64_t* dst = nullptr;
    if (FILL) dst = index + index_offsets[b] + rs[qi];
    int cnt = 0;
    const unsigned below = (1u << lane) - 1u;
    const int gl = G.g[D - 1];
    const int l_lo = max(c[D - 1] - 1, 0), l_hi = min(c[D - 1] + 1, gl - 1);      // last-coordinate cell range
    if (l_lo <= l_hi) {
        const int y_lo = D == 3 ? -1 : 0, y_hi = D == 3 ? 1 : 0;
        for (int dx = -1; dx <= 1; ++dx) {
            const int cx = c[0] + dx;
            if (cx < 0 || cx >= G.g[0]) continue;
            for (int dy = y_lo; dy <= y_hi; ++dy) {
                int col = cx;                              // index of the column of cells along the last coordinate
                if (D == 3) {
                    const int cy = c[1] + dy;
                    if (cy < 0 || cy >= G.g[1]) continue;
                    col = cx * G.g[1] + cy;
                }
                const int a = __ldg(cs + col * gl + l_lo), e = __ldg(cs + col * gl + l_hi + 1);
                for (int t0 = a; t0 < e; t0 += 32) {
                    const int t = t0 + lane;
                    const bool inb = t < e;
                    const int j = inb ? __ldg(srt + t) : 0;
                    float p[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) p[k] = __ldg(dat + (int64_t)j * D + k);
                    const float d2 = dist2_exact<D>(p, q);
                    const bool hit = inb && (inclusive ? (d2 <= r2) : (d2 < r2));
                    const unsigned bal = __ballot_sync(0xffffffffu, hit);
                    if (FILL && hit) dst[cnt + __popc(bal & below)] = j;
                    cnt += __popc(bal);
                }
            }
        }
    }
    if (!FILL) {
        if (lane == 0) rs[1 + qi] = cnt;
        return;
    }
    if (cnt < 2) return;
    __syncwarp();
    if (cnt <= 128) {                                   // rank sort: indices inside a row are distinct
        int64_t v[4];
        int rk[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int i = lane + 32 * s;
            v[s] = i < cnt ? dst[i] : 0;
            rk[s] = 0;
        }
        for (int k = 0; k < cnt; ++k) {
            const int64_t x = dst[k];                   // warp-uniform address
#pragma unroll
            for (int s = 0; s < 4; ++s) rk[s] += x < v[s];
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (lane + 32 * s < cnt) dst[rk[s]] = v[s];
    } else if (lane == 0) {
        for (int i = 1; i < cnt; ++i) {
            const int64_t key = dst[i];
            int k = i - 1;
            while (k >= 0 && dst[k] > key) { dst[k + 1] = dst[k]; --k; }
            dst[k + 1] = key;
        }
    }
}

inline int make_geom(GridGeom& G, int d, float lox, float loy, float loz, float h, int gx, int gy, int gz) {
    if (!(h > 0.f) || gx < 1 || gy < 1 || gz < 1 || (d == 2 && gz != 1)) return GAOT_E_BADARG;
    if ((int64_t)gx * gy * gz > (1LL << 26)) return GAOT_E_BADARG;
    G.lo[0] = lox; G.lo[1] = loy; G.lo[2] = loz;
    G.inv_h = 1.0f / h;
    G.g[0] = gx; G.g[1] = gy; G.g[2] = gz;
    G.ncell = gx * gy * gz;
    return 0;
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

// ---- cell-binned search (see the kernel comment above) ----
GAOT_API int gaot_grid_bin(const float* data, int64_t n, int d, int64_t batch, float lox, float loy, float loz, float h, int gx,
                           int gy, int gz, int* cell_start, int* sorted, int* cell_ws, int* cursor_ws, void* stream) {
    if (!data || !cell_start || !sorted || !cell_ws || !cursor_ws || n < 0 || batch < 1 || (d != 2 && d != 3)) return GAOT_E_BADARG;
    if (n > 0x7fffffffLL) return GAOT_E_BADARG;
    GridGeom G;
    int rc = make_geom(G, d, lox, loy, loz, h, gx, gy, gz);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_CHECK(cudaMemsetAsync(cell_start, 0, sizeof(int) * (size_t)(batch * (G.ncell + 1)), st));
    GAOT_CHECK(cudaMemsetAsync(cursor_ws, 0, sizeof(int) * (size_t)(batch * G.ncell), st));
    if (n == 0) return 0;
    dim3 gp((unsigned)gaot_grid(n, 256, 4), (unsigned)batch);
    if (d == 2) grid_hist_kernel<2><<<gp, 256, 0, st>>>(data, n, 1, G, cell_start, cell_ws);
    else grid_hist_kernel<3><<<gp, 256, 0, st>>>(data, n, 1, G, cell_start, cell_ws);
    scan_i32_kernel<<<(unsigned)batch, 1024, 0, st>>>(cell_start, G.ncell);
    grid_scatter_kernel<<<gp, 256, 0, st>>>(cell_ws, n, G.ncell, cell_start, cursor_ws, sorted);
    dim3 gc((unsigned)((G.ncell + 127) / 128), (unsigned)batch);
    grid_sort_cells_kernel<<<gc, 128, 0, st>>>(cell_start, G.ncell, n, sorted);
    GAOT_NOTE_LAUNCHES(3);
    GAOT_RETURN_LAST();
}

namespace {
template <bool FILL>
int radius_grid_launch(const float* data, int64_t n, const float* queries, int64_t m, int d, float radius, int inclusive,
                       int64_t batch, int db, int qb, float lox, float loy, float loz, float h, int gx, int gy, int gz,
                       const int* cell_start, const int* sorted, int64_t* row_splits, const int64_t* index_offsets, int64_t* index,
                       cudaStream_t st) {
    if (!data || !queries || !cell_start || !sorted || !row_splits || n < 0 || m < 0 || batch < 1 || (d != 2 && d != 3)) return GAOT_E_BADARG;
    if (FILL && (!index_offsets || !index)) return GAOT_E_BADARG;
    if (h < radius * 1.0005f) return GAOT_E_BADARG;             // the cell pitch must cover the radius with margin
    GridGeom G;
    int rc = make_geom(G, d, lox, loy, loz, h, gx, gy, gz);
    if (rc) return rc;
    const float r2 = radius * radius;                           // fp32 product, same rounding as the oracle
    if (m > 0) {
        dim3 grid((unsigned)((m + 3) / 4), (unsigned)batch);       // one warp per query, four queries per CTA
        if (d == 2) radius_grid_kernel<2, FILL><<<grid, 128, 0, st>>>(data, n, queries, m, r2, inclusive, db, qb, G, cell_start, sorted,
                                                                     row_splits, index_offsets, index);
        else radius_grid_kernel<3, FILL><<<grid, 128, 0, st>>>(data, n, queries, m, r2, inclusive, db, qb, G, cell_start, sorted,
                                                               row_splits, index_offsets, index);
    }
    if (!FILL) { splits_scan_kernel<<<(unsigned)batch, 1024, 0, st>>>(row_splits, m); GAOT_NOTE_LAUNCHES(1); }
    GAOT_RETURN_LAST();
}
}  // namespace

GAOT_API int gaot_radius_count_grid(const float* data, int64_t n, const float* queries, int64_t m, int d, float radius, int inclusive,
                                    int64_t batch, int data_batched, int queries_batched, float lox, float loy, float loz, float h,
                                    int gx, int gy, int gz, const int* cell_start, const int* sorted, int64_t* row_splits,
                                    void* stream) {
    return radius_grid_launch<false>(data, n, queries, m, d, radius, inclusive, batch, data_batched, queries_batched, lox, loy, loz, h,
                                     gx, gy, gz, cell_start, sorted, row_splits, nullptr, nullptr, (cudaStream_t)stream);
}

GAOT_API int gaot_radius_fill_grid(const float* data, int64_t n, const float* queries, int64_t m, int d, float radius, int inclusive,
                                   int64_t batch, int data_batched, int queries_batched, float lox, float loy, float loz, float h,
                                   int gx, int gy, int gz, const int* cell_start, const int* sorted, const int64_t* row_splits,
                                   const int64_t* index_offsets, int64_t* index, void* stream) {
    return radius_grid_launch<true>(data, n, queries, m, d, radius, inclusive, batch, data_batched, queries_batched, lox, loy, loz, h,
                                    gx, gy, gz, cell_start, sorted, const_cast<int64_t*>(row_splits), index_offsets, index,
                                    (cudaStream_t)stream);
}
