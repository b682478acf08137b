/* gaot_b200 - C ABI of the B200-native GAOT forward/backward hot path.
 *
 * Every entry point
 *   - takes plain device pointers + sizes (no torch types), row-major contiguous unless a leading
 *     dimension is passed, fp32 activations/weights, int64 CSR arrays exactly as the reference
 *     stores them ({"neighbors_index", "neighbors_row_splits"}, src/model/layers/utils/neighbor_search.py:80-90);
 *   - is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises, keeps no
 *     global state, and is re-entrant per stream;
 *   - returns 0 on success or a cudaError_t value (>0) / GAOT_E_* (<0) on failure.  The caller
 *     allocates every output and workspace.
 *
 * The comment on each function names the reference interface it replaces (paths relative to the
 * reference checkout /root/reference).  The reference has no FFI of its own (pure Python over
 * torch / torch_scatter / torch_cluster wheels); INTEGRATION.md shows the ctypes binding a maintainer
 * would add in those modules.
 */
#ifndef GAOT_B200_H
#define GAOT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAOT_E_BADARG (-1)   /* unsupported size / null pointer / bad flag            */
#define GAOT_E_ALIGN  (-2)   /* pointer or leading dimension violates an alignment rule */

/* activation codes for the dense epilogues */
#define GAOT_ACT_NONE 0
#define GAOT_ACT_GELU 1      /* exact erf GELU (torch F.gelu default)  */
#define GAOT_ACT_RELU 2

int gaot_abi_version(void);
/* kernels launched by this library in this process so far (diagnostics; bench.py `gpu_launches`) */
long long gaot_launch_count(void);
long long gaot_launch_counter_add(long long n);   /* host code adds launches replayed through a CUDA graph */

/* ------------------------------------------------------------------------------------------------
 * Radius graph.  Replaces NeighborSearch.forward and its back-ends
 * (src/model/layers/utils/neighbor_search.py:65-106, :108-144 `native`, :147-180 `torch_cluster`,
 * :275-335 `chunked`) and the per-sample loops of src/datasets/graph_builder.py:28-87.
 * Row q lists, in ascending j, every j with sum_k (data[j,k]-queries[q,k])^2 (<=|<) r*r, all in
 * fp32 with no fma contraction (bit-exact with the oracle).  inclusive=1,max_nbrs=0 is `native`;
 * inclusive=0,max_nbrs=32 is torch_cluster.radius as the reference calls it.
 * Two phases because E is data dependent: count writes the finished row_splits (exclusive scan,
 * row_splits[m] = E); the caller reads E, allocates index[E], and calls fill.
 * `batch` graphs are built in one launch (vx mode): data [batch, n, d] if data_batched else [n, d],
 * queries likewise; row_splits [batch, m+1] (each starting at 0), index_offsets [batch] gives where
 * each sample's edges start inside `index`.
 * ---------------------------------------------------------------------------------------------- */
int gaot_radius_count(const float* data, int64_t n, const float* queries, int64_t m, int d,
                      float radius, int inclusive, int64_t max_nbrs,
                      int64_t batch, int data_batched, int queries_batched,
                      int64_t* row_splits, void* stream);
int gaot_radius_fill(const float* data, int64_t n, const float* queries, int64_t m, int d,
                     float radius, int inclusive, int64_t max_nbrs,
                     int64_t batch, int data_batched, int queries_batched,
                     const int64_t* row_splits, const int64_t* index_offsets,
                     int64_t* index, void* stream);

/* Cell-binned form of the two entry points above: the SAME graph (same fp32 distance arithmetic, rows in ascending j) in
 * O(points of the 3^d cells around a query) instead of O(n) per query; for max_nbrs = 0 (all neighbours) only.
 * Replaces the reference's pure-Python `grid` back-end (neighbor_search.py:183-272).
 * gaot_grid_bin: counting sort of the data points [batch, n, d] by cell of a uniform grid with origin lo, pitch
 * h >= 1.0005 * radius and g = (gx, gy, gz) cells (gz = 1 for d = 2): cell_start [batch, ncell + 1], sorted [batch, n]
 * (point ids, ascending inside a cell); cell_ws [batch, n] and cursor_ws [batch, ncell] are int32 workspaces.
 * count / fill: as gaot_radius_count / gaot_radius_fill; data_batched = 0 shares one binned point set between the samples. */
int gaot_grid_bin(const float* data, int64_t n, int d, int64_t batch, float lox, float loy, float loz, float h, int gx,
                  int gy, int gz, int* cell_start, int* sorted, int* cell_ws, int* cursor_ws, void* stream);
int gaot_radius_count_grid(const float* data, int64_t n, const float* queries, int64_t m, int d, float radius, int inclusive,
                           int64_t batch, int data_batched, int queries_batched, float lox, float loy, float loz, float h,
                           int gx, int gy, int gz, const int* cell_start, const int* sorted, int64_t* row_splits,
                           void* stream);
int gaot_radius_fill_grid(const float* data, int64_t n, const float* queries, int64_t m, int d, float radius, int inclusive,
                          int64_t batch, int data_batched, int queries_batched, float lox, float loy, float loz, float h,
                          int gx, int gy, int gz, const int* cell_start, const int* sorted, const int64_t* row_splits,
                          const int64_t* index_offsets, int64_t* index, void* stream);

/* edge_row[e] = the CSR row of edge e (the reference materialises the same thing with
 * torch.repeat_interleave, src/model/layers/agno.py:206-207). */
int gaot_csr_edge_rows(const int64_t* row_splits, int64_t m, int64_t* edge_row, void* stream);

/* Source-sorted (transposed) view of a CSR graph, needed by the backward of the gather
 * (autograd's index_add for agno.py:198).  t_splits [n_src+1]; t_rows[k] = destination row of the
 * k-th transposed edge; t_perm[k] = its edge id in the original CSR.  Lists ascending in edge id. */
int gaot_csr_transpose(const int64_t* index, const int64_t* row_splits, int64_t m, int64_t n_src,
                       int64_t E, int64_t* t_splits, int64_t* t_rows, int64_t* t_perm,
                       int64_t* cursor_ws /* [n_src] workspace */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * AGNO (src/model/layers/agno.py:148-273).
 * ---------------------------------------------------------------------------------------------- */
/* z[e,:] = [ y[index[e], :dy], x[row(e), :dx] ]   (agno.py:188, 206-207, 229) */
int gaot_edge_inputs(const float* y, int64_t dy, const float* x, int64_t dx,
                     const int64_t* index, const int64_t* edge_row, int64_t E, float* z, void* stream);

/* alpha[e] = softmax over the row of cos(x_row, y_j) using the first `cd` coordinates
 * (F.normalize eps 1e-12; agno.py:218-224 and _segment_softmax :112-146). */
int gaot_edge_cosine_softmax(const float* y, int64_t dy, const float* x, int64_t dx, int64_t cd,
                             const int64_t* index, const int64_t* row_splits, int64_t m,
                             float* alpha, void* stream);

/* Dot-product attention weights (agno.py:105-110, 214-217): q [m, A] = query_proj(x[:, :cd]) per destination row,
 * k [n, A] = key_proj(y[:, :cd]) per source point (the reference projects per edge; same values),
 * alpha[e] = softmax over the row of scale * <q[row(e)], k[index[e]]>   (_segment_softmax, agno.py:112-146). */
int gaot_edge_dot_softmax_fwd(const float* q, const float* k, int64_t A, const int64_t* index, const int64_t* row_splits,
                              int64_t m, float scale, float* alpha, void* stream);

/* Backward of gaot_edge_dot_softmax_fwd: dq [m, A] is written, dk [n, A] is zeroed here and accumulated with red.add. */
int gaot_edge_dot_softmax_bwd(const float* dalpha, const float* alpha, const float* q, const float* k, int64_t A,
                              const int64_t* index, const int64_t* row_splits, int64_t m, int64_t n, float scale,
                              float* dq, float* dk, void* stream);

/* out[r] = sum_c a[r,c] * b[r,c]   (gradient of the row scale of gaot_scale_rows: dalpha[e] = <dW[e,:], kappa[e,:]>) */
int gaot_row_dot(const float* a, const float* b, int64_t rows, int64_t C, float* out, void* stream);

/* out[e,c] = scale[e] * in[e,c]    (agno.py:250 forward; alpha * dW in the backward) */
int gaot_scale_rows(const float* in, const float* scale, int64_t rows, int64_t C, float* out, void* stream);

/* out[b,r,c] = sum_{e in row r} w[p(e),c] * f[b,index[e],c],  p(e) = w_perm ? w_perm[e] : e.
 * Replaces the gather + multiply + torch_scatter.segment_csr chain
 * (agno.py:198, 245-250, 264-271; src/model/layers/utils/segment_csr.py:48-55).
 * With the transposed graph and w_perm = t_perm it is also the backward w.r.t. f. */
int gaot_agno_aggregate(const float* f, const float* w, const int64_t* w_perm,
                        const int64_t* index, const int64_t* row_splits,
                        float* out, int64_t B, int64_t Ns, int64_t Nd, int64_t C, void* stream);

/* dw[e,c] = sum_b g[b,row(e),c] * f[b,index[e],c]   (backward of agno.py:245-271 w.r.t. the edge weights) */
int gaot_agno_weight_grad(const float* g, const float* f, const int64_t* index,
                          const int64_t* row_splits, float* dw,
                          int64_t B, int64_t Ns, int64_t Nd, int64_t E, int64_t C, void* stream);

/* Fused lifting + aggregate (encoder): out[b,r,c] = sum_{e in r} w[e,c] * (sum_i Wl[c,i] pn[b,index[e],i] + bl[c]).
 * Replaces ChannelMLP lifting (magno.py:273-274, mlp.py:274-275: a single affine map) followed by the
 * gather/multiply/segment_csr chain of agno.py:198,245-271 without materialising the lifted features.
 * pn [B,Ns,Cin], Wl [C,Cin], bl [C], w [E,C]; Cin <= 4, C % 4 == 0, C <= 128.
 * pn_t (optional, may be NULL): the same features SAMPLE-MINOR, [Ns,B,Cin]; when B is a multiple of the kernel's sample
 * chunk (8 for B >= 8) and chunk * Cin % 4 == 0 the gather reads it instead - one 128-bit load per 4 samples. */
int gaot_lift_aggregate_fwd(const float* pn, const float* Wl, const float* bl, const float* w, const int64_t* index,
                            const int64_t* row_splits, float* out, int64_t B, int64_t Ns, int64_t Nd, int64_t C,
                            int64_t Cin, const float* pn_t, void* stream);
/* its backward: dw [E,C], dWl [C,Cin], dbl [C] (all overwritten); pn carries no gradient. */
int gaot_lift_aggregate_bwd(const float* g, const float* pn, const float* Wl, const float* bl, const float* w,
                            const int64_t* index, const int64_t* row_splits, float* dw, float* dWl, float* dbl,
                            int64_t B, int64_t Ns, int64_t Nd, int64_t E, int64_t C, int64_t Cin, const float* pn_t,
                            void* stream);

/* Fused aggregate + affine read-out (decoder): out[b,r,o] = sum_{e in r} <u[o,e,:], f[b,index[e],:]>, with
 * u[o,e,c] = (W_proj W_rec_a)[o,c] * w[e,c] folded by the caller.  Replaces agno.py:198-271 followed by the
 * recovery and projection 1x1 convolutions (magno.py:684-688, 640-641), both affine.  f [B,Ns,C],
 * u [Cout,E,C], out [B,Nd,Cout]; Cout <= 4, C % 4 == 0, C <= 128. */
int gaot_agno_dot_fwd(const float* f, const float* u, const int64_t* index, const int64_t* row_splits, float* out,
                      int64_t B, int64_t Ns, int64_t Nd, int64_t E, int64_t C, int64_t Cout, void* stream);
/* its backward: df [B,Ns,C] through the transposed graph (nullable), du [Cout,E,C] (nullable). */
int gaot_agno_dot_bwd(const float* g, const float* f, const float* u, const int64_t* index, const int64_t* row_splits,
                      const int64_t* t_splits, const int64_t* t_rows, const int64_t* t_perm, float* df, float* du,
                      int64_t B, int64_t Ns, int64_t Nd, int64_t E, int64_t C, int64_t Cout, void* stream);

/* Source-walk forms of the two entry points above (the default decoder path).  A group of lanes owns one SOURCE node j
 * (latent token) of the transposed graph {t_splits [Ns+1], t_rows [E] destination row of each edge, t_perm [E] edge id in the
 * row-sorted order}: it reads r[b, j, :] once per sample and serves every edge leaving j from registers, instead of
 * gathering the 256-byte latent row again for every edge and sample (agno.py:198 materialises exactly that gather).
 * Forward: `out` is zeroed, then reduced into with fp32 red.add (order inside a row not fixed; gaot_agno_dot_fwd is the
 * deterministic form).  Backward: dr and du in one pass, plain stores (du by red.add only when the batch is split).
 * Batches of >= 16 samples with C in {16, 32, 48, 64} take the lane-per-sample kernels: one warp owns a source node and 32
 * samples, a lane keeps its sample's whole latent row in registers and the edge weights are warp-uniform loads, so no
 * cross-lane reduction is needed; du (the product that reduces over samples) then runs on the row-sorted graph, which the
 * optional index / row_splits arguments of the backward provide. */
int gaot_agno_dot_fwd_t(const float* r, const float* u, const int64_t* t_splits, const int64_t* t_rows,
                        const int64_t* t_perm, float* out, int64_t B, int64_t Ns, int64_t Nd, int64_t E, int64_t C,
                        int64_t Cout, void* stream);
int gaot_agno_dot_bwd_t(const float* g, const float* r, const float* u, const int64_t* index, const int64_t* row_splits,
                        const int64_t* t_splits, const int64_t* t_rows, const int64_t* t_perm, float* dr, float* du,
                        int64_t B, int64_t Ns, int64_t Nd, int64_t E, int64_t C, int64_t Cout, void* stream);

/* z[e,k] = y[index[e],k] - x[edge_row[e],k], k < d: neighbour coordinates centred on their query point
 * (pointnet geometry embedding, gemb.py:208-213) */
int gaot_edge_diff(const float* y, int64_t dy, const float* x, int64_t dx, int64_t d, const int64_t* index,
                   const int64_t* edge_row, int64_t E, float* z, void* stream);

/* backward of gaot_segment_csr for one batch (B = 1): dsrc[e,:] from dout[row(e),:]; reduce 2 (max) sends the gradient to the
 * first edge of the row that attains the maximum (scatter_max's argmax, gemb.py:217) and needs src and out. */
int gaot_segment_csr_bwd(const float* dout, const float* src, const float* out, const int64_t* row_splits,
                         const int64_t* edge_row, float* dsrc, int64_t E, int64_t m, int64_t C, int reduce, void* stream);

/* generic segment reduction of src[B,E,C] over CSR rows, reduce 0=sum 1=mean 2=max (empty -> 0);
 * the drop-in for segment_csr(src, indptr, reduce) (segment_csr.py:14-19). */
int gaot_segment_csr(const float* src, const int64_t* row_splits, float* out,
                     int64_t B, int64_t E, int64_t m, int64_t C, int reduce, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Statistical geometry embedding, parameter-free part
 * (GeometricEmbedding._compute_statistical_features, src/model/layers/gemb.py:83-171).
 * feat [m, 3+2d] = [deg, mean dist, var dist, centroid-q (d), eig(cov) descending (d)], empty rows 0,
 * then every column standardised over the m rows (unbiased std, std<1e-6 -> 1).  fp64 inside.
 * ---------------------------------------------------------------------------------------------- */
int gaot_geo_stats(const float* src, const float* queries, int64_t d,
                   const int64_t* index, const int64_t* row_splits, int64_t m,
                   float* feat, void* stream);
/* standardises each of `groups` consecutive blocks of m rows separately (vx mode: one block per sample) */
int gaot_col_standardize(float* feat, int64_t m, int64_t F, int64_t groups, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense layers: nn.Linear / 1x1 Conv1d of src/model/layers/mlp.py:283-337, attn.py:92-94,117,151,
 * gaot.py:208.  C[M,N] = act( opA(A)[M,K] * opB(B)[K,N] + bias[n] + add[(m % add_rows), n] ),
 * opA(A)[m,k] = transA ? A[k*lda+m] : A[m*lda+k];  opB(B)[k,n] = transB ? B[n*ldb+k] : B[k*ldb+n].
 * accumulate=1 adds the result to what C already holds.  pre (nullable) receives the value before
 * the activation.  fp32 FFMA arithmetic (the parity path).
 * ---------------------------------------------------------------------------------------------- */
int gaot_gemm_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                  int64_t lda, int64_t ldb, int64_t ldc, int transA, int transB,
                  const float* bias, const float* add, int64_t add_rows, int64_t ldadd,
                  int act, float* pre, int accumulate, void* stream);

/* Tensor-core (tcgen05, TMA, TMEM) variant for the processor: A [M,K] and B [N,K] are bf16 row-major
 * (B in nn.Linear layout), fp32 accumulation; C [M,N] is fp32 (out_bf16=0) or bf16 (out_bf16=1).
 * act / pre as in gaot_gemm_f32 (pre is fp32).  accumulate=1 (fp32 C only, no act) adds into C.
 * K % 8 == 0 and 16-byte aligned bases (TMA). */
int gaot_gemm_bf16(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int out_bf16,
                   const float* bias, const float* add, int64_t add_rows, int64_t ldadd, int act, float* pre,
                   int accumulate, void* stream);
/* Gated FFN, forward (reference FFN.forward, src/model/layers/attn.py:137-146: w2(silu(w1 x) * w3 x)): the w1 / w3
 * product with the SwiGLU fused into the epilogue.  B [N,K] holds the w1 and w3 rows INTERLEAVED in blocks of 32
 * (rows [64j, 64j+32) = w1[32j..], rows [64j+32, 64j+64) = w3[32j..]), so AB [M,N] (bf16, kept for the backward) carries
 * (a | b) in column blocks of 64 and U [M,N/2] (bf16) = silu(a) * b in the feature order of w1.  N % 64 == 0, N >= 128,
 * K % 8 == 0, 16-byte aligned bases. */
int gaot_gemm_bf16_swiglu(const void* A, const void* B, void* AB, void* U, int64_t M, int64_t N, int64_t K, void* stream);
/* Gated FFN, backward of the same lines: dU = dY[M,K] W[K,N] (W = w2 in nn.Linear layout, consumed MN-major) never
 * reaches memory - the epilogue reads the saved (a | b) blocks of AB [M,2N] and writes dAB [M,2N] (bf16) =
 * (dU b silu'(a) | dU silu(a)) in the same interleaved layout.  N % 32 == 0, N >= 128, K % 8 == 0. */
int gaot_gemm_bf16_nn_swiglu_bwd(const void* dY, const void* W, const void* AB, void* dAB, int64_t M, int64_t N, int64_t K,
                                 void* stream);
/* Input-gradient form: C[M,N] (fp32 or bf16) = A[M,K] B[K,N], B row-major [K,N] consumed MN-major in place
 * (dX = dY W with W in nn.Linear layout).  K % 8 == 0, N % 8 == 0. */
int gaot_gemm_bf16_nn(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int out_bf16,
                      void* stream);
/* Weight-gradient form: C[M,N] (fp32) = A[R,M]^T B[R,N], A and B bf16 row-major with the contraction over
 * their R rows (dW = dY^T X): both operands are consumed MN-major straight from their natural layout.
 * M % 8 == 0, N % 8 == 0. */
int gaot_gemm_bf16_tn(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t R, int accumulate,
                      void* stream);
/* Split-precision (parity) tensor-core path.  The reference computes every nn.Linear of the processor in fp32
 * (src/core/base_trainer.py:63-68 admits only fp32/fp64); north_star asks for <= 1e-5 of it.  An fp32 tensor x is
 * written as two fp16 planes, x * scale = hi + lo (~22 significant bits), scale a power of two chosen from max|x| so that
 * the largest magnitude sits just below 2^14; a product then needs three tensor-core terms (lo*hi + hi*lo + hi*hi, the
 * lo*lo term is < 2^-22 relative) accumulated in fp32 in TMEM - the same tcgen05 / TMA kernel as gaot_gemm_bf16 with a
 * contraction that is three terms long.
 * gaot_split_f16: planes = fp16 [2, n] (hi plane then lo plane), scale = {scale, 1 / scale} (device), ws = workspace of
 * gaot_split_ws_floats() floats; n % 8 == 0. */
int gaot_split_ws_floats(void);
int gaot_split_f16(const float* in, void* planes, float* scale, float* ws, int64_t n, void* stream);
/* C[M,N] (fp32) = act(opA opB / (scale_a scale_b) + bias + add[(m % add_rows), n]) on split operands:
 *   a_mn = 0: A planes [2, M, K] (contraction along the stored row);   a_mn = 1: A planes [2, K, M] (contraction over rows)
 *   b_mn = 0: B planes [2, N, K] (nn.Linear weight layout);            b_mn = 1: B planes [2, K, N]
 * i.e. forward (0,0): x W^T, input gradient (0,1): dY W, weight gradient (1,1): dY^T X, each operand in its natural
 * layout.  sa / sb: the {scale, 1 / scale} pairs of gaot_split_f16.  bias / add / act / pre / accumulate as in
 * gaot_gemm_bf16. */
int gaot_gemm_split(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int a_mn, int b_mn,
                    const float* sa, const float* sb, const float* bias, const float* add, int64_t add_rows,
                    int64_t ldadd, int act, float* pre, int accumulate, void* stream);
/* out[r,c] = bf16(in[r,c]) (out nullable) and/or out_t[c,r] = bf16(in[r,c]) (out_t nullable) */
int gaot_cast_bf16(const float* in, void* out, int64_t rows, int64_t cols, void* out_t, void* stream);
/* out_t[c,r] = in[r,c], bf16 */
int gaot_transpose_bf16(const void* in, void* out_t, int64_t rows, int64_t cols, void* stream);

/* out[n] (+)= sum_m x[m,n]  (bias gradients) */
int gaot_col_sum(const float* x, int64_t M, int64_t N, float* out, int accumulate, void* stream);
/* dx (bf16) = dy * act'(pre), dy fp32 (dy_is_bf16=0) or bf16; n % 4 == 0 */
int gaot_act_bwd_bf16(const void* dy, int dy_is_bf16, const float* pre, void* dx, int64_t n, int act, void* stream);
/* dx = dy * act'(pre) */
int gaot_act_bwd(const float* dy, const float* pre, float* dx, int64_t n, int act, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Processor pieces (src/model/layers/attn.py).
 * ---------------------------------------------------------------------------------------------- */
/* RMSNorm.forward (attn.py:167-172): y = x * rsqrt(mean(x^2)+eps) * gamma; rstd[rows] saved. */
/* y (fp32) and/or y_bf16 may be requested (either nullable, not both); the backward adds the gradients of both. */
int gaot_rmsnorm_fwd(const float* x, const float* gamma, float* y, void* y_bf16, float* rstd,
                     int64_t rows, int64_t D, float eps, void* stream);
int gaot_rmsnorm_bwd(const float* dy, const void* dy_bf16, const float* x, const float* gamma, const float* rstd,
                     float* dx, float* dgamma /* accumulated, caller zeroes */,
                     int64_t rows, int64_t D, void* stream);
/* FFN gate (attn.py:150-156): u = silu(ab[:, :F]) * ab[:, F:]  with ab = x [W1;W3]^T  [rows, 2F] */
int gaot_swiglu_fwd(const float* ab, float* u, int64_t rows, int64_t F, void* stream);
int gaot_swiglu_bwd(const float* du, const float* ab, float* dab, int64_t rows, int64_t F, void* stream);
/* bf16 in / bf16 out variants (F % 8 == 0), used between tensor-core GEMMs */
int gaot_swiglu_fwd_bf16(const void* ab, void* u, int64_t rows, int64_t F, void* stream);
int gaot_swiglu_bwd_bf16(const void* du, const void* ab, void* dab, int64_t rows, int64_t F, void* stream);

/* Non-causal softmax(Q K^T * scale) V, GroupQueryFlashAttention.forward (attn.py:96-116).
 * q[(b*T+t)*ldq + h*hd + i], k/v[(b*T+t)*ldk + hk*hd + i] with hk = h / (H/Hkv);
 * o [B,T,H*hd] contiguous; lse [B,H,T].  hd in {16,32,64}. */
int gaot_attention_fwd(const float* q, const float* k, const float* v, float* o, float* lse,
                       int64_t B, int64_t T, int64_t H, int64_t Hkv, int64_t hd,
                       int64_t ldq, int64_t ldk, float scale, void* stream);
/* delta [B,H,T] is workspace.  dq/dk/dv use the same strides as q/k/v. */
int gaot_attention_bwd(const float* q, const float* k, const float* v, const float* o,
                       const float* lse, const float* dout, float* delta,
                       float* dq, float* dk, float* dv,
                       int64_t B, int64_t T, int64_t H, int64_t Hkv, int64_t hd,
                       int64_t ldq, int64_t ldk, float scale, void* stream);

/* Tensor-core variant (tcgen05 / TMA / TMEM): fused qkv [B*T, ld] bf16 with per-token columns
 * [H*hd query | Hkv*hd key | Hkv*hd value]; o [B*T, H*hd] bf16; lse [B,H,T] fp32 (natural log). */
int gaot_attention_fwd_bf16(const void* qkv, void* o, float* lse, int64_t B, int64_t T, int64_t H, int64_t Hkv,
                            int64_t hd, int64_t ld, float scale, void* stream);

/* Backward of the above.  dout [B*T, H*hd] bf16; workspaces delta [B,H,T] fp32 and dq_acc [B*T, H*hd] fp32;
 * dqkv [B*T, ld] bf16 receives dQ | dK | dV in the layout of qkv. */
int gaot_attention_bwd_bf16(const void* qkv, const void* o, const float* lse, const void* dout, float* delta,
                            float* dq_acc, void* dqkv, int64_t B, int64_t T, int64_t H, int64_t Hkv, int64_t hd,
                            int64_t ld, float scale, void* stream);

/* Split-precision (parity) tensor-core attention: the same tcgen05 kernels with fp16 hi / lo operand planes
 * (gaot_split_f16 of the fp32 fused qkv matrix [B*T, ld] -> qkv_planes [2, B*T, ld], qscale = its {scale, 1/scale}),
 * three tensor-core terms per product, fp32 softmax, probabilities split the same way, fp32 outputs.
 * o [B*T, H*hd] fp32, lse [B,H,T].  hd in {16, 32}. */
int gaot_attention_fwd_split(const void* qkv_planes, const float* qscale, float* o, float* lse, int64_t B, int64_t T,
                             int64_t H, int64_t Hkv, int64_t hd, int64_t ld, float scale, void* stream);
/* Backward: dout [B*T, H*hd] fp32 and its planes / scale; delta [B,H,T] workspace; dqkv [B*T, ld] fp32 receives
 * dQ | dK | dV in the layout of qkv. */
int gaot_attention_bwd_split(const void* qkv_planes, const float* qscale, const float* o, const float* lse,
                             const float* dout, const void* dout_planes, const float* gscale, float* delta, float* dqkv,
                             int64_t B, int64_t T, int64_t H, int64_t Hkv, int64_t hd, int64_t ld, float scale,
                             void* stream);

/* Patchify / un-patchify permutations of GAOT.process (src/model/gaot.py:182-185, 202-205, 225-231).
 * grid [B, H*W*(D), C] <-> tokens [B, T, P^dims * C]; dims 2 or 3; to_tokens 1 = patchify. */
int gaot_patch_permute(const float* in, float* out, int64_t B, int dims, int64_t H, int64_t W,
                       int64_t D, int64_t P, int64_t C, int to_tokens, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Step (src/trainer/static_trainer.py:160-178, src/core/base_trainer.py:71, src/utils/optimizers.py:252-255)
 * ---------------------------------------------------------------------------------------------- */
/* loss[0] += mean((pred-target)^2) (caller zeroes); dpred = 2 (pred-target) / n * gscale (nullable). */
int gaot_mse_fwd_bwd(const float* pred, const float* target, float* loss, float* dpred,
                     int64_t n, float gscale, void* stream);
/* torch.optim.AdamW step on a flat parameter buffer; grad is multiplied by grad_scale first
 * (1/world_size after a summing all-reduce). */
int gaot_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                    float lr, float beta1, float beta2, float eps, float weight_decay,
                    int64_t step, float grad_scale, void* stream);

/* Fused edge-weight operator of AGNO on the tensor cores (bf16 mode):
 *     w[e, :] = alpha[e] * MLP(cat[y[index[e], :kc], x[edge_row[e], :kc]]),
 *     MLP = Linear(2kc -> 64) GELU [Linear(64 -> 64) GELU] x n_hidden, Linear(64 -> C)     (exact-erf GELU)
 * replacing the gather / repeat of the edge coordinates (agno.py:206-207), the kernel MLP (agno.py:229-242 with
 * LinearChannelMLP, mlp.py:329-337) and the attention-weight product (agno.py:250), forward and backward, without ever
 * writing an [E, 64] activation or gradient to memory: one CTA keeps a 128-edge tile's whole chain in shared memory / TMEM,
 * the hidden and output layers run as tcgen05.mma on bf16 operands (first layer on the FP32 pipe from fp32 coordinates),
 * and the backward accumulates dW_l = delta_l^T h_{l-1} on the tensor cores into TMEM for the CTA's lifetime.
 * y [*, dy], x [*, dx] fp32 coordinate (or position-encoding) rows; index / edge_row [E]; alpha [E] (nullable);
 * W1 [64, 2kc], Wh0 / Wh1 [64, 64] (Wh1 unused when n_hidden == 1), Wo [C, 64] and their biases, all fp32;
 * gaot_edge_mlp_supported(2kc, H, n_hidden, C) tells whether the shape is instantiated (H = 64, C % 16 == 0, C <= 64,
 * 2kc in {4, 6, 32, 48}, n_hidden in {1, 2}).  The backward ZEROES and then reduces into dW1 ... dbo; alpha and the
 * coordinates receive no gradient (cosine / mean weights are parameter free). */
int gaot_edge_mlp_supported(int64_t kin, int64_t H, int64_t n_hidden, int64_t C);
int gaot_edge_mlp_fwd(const float* y, int64_t dy, const float* x, int64_t dx, int64_t kc, const int64_t* index,
                      const int64_t* edge_row, int64_t E, const float* alpha, const float* W1, const float* b1,
                      const float* Wh0, const float* bh0, const float* Wh1, const float* bh1, int64_t n_hidden,
                      const float* Wo, const float* bo, int64_t C, float* w, void* stream);
int gaot_edge_mlp_bwd(const float* y, int64_t dy, const float* x, int64_t dx, int64_t kc, const int64_t* index,
                      const int64_t* edge_row, int64_t E, const float* alpha, const float* W1, const float* b1,
                      const float* Wh0, const float* bh0, const float* Wh1, const float* bh1, int64_t n_hidden,
                      const float* Wo, const float* bo, int64_t C, const float* dw, float* dW1, float* db1, float* dWh0,
                      float* dbh0, float* dWh1, float* dbh1, float* dWo, float* dbo, void* stream);

/* Rollout update rule of GAOT._process_autoregressive_prediction (src/model/gaot.py:436-476) and the state carry of
 * src/model/gaot.py:431-434.  mode 0 output, 1 residual, 2 time_der.  All tensors [rows, C]; stats are per channel [C].
 * u_phys = the de-normalised prediction (the value the reference appends and returns); u_next (optional) =
 * (u_phys - u_mean) / u_std, the normalised state fed to the next step.  u_in / s_mean / s_std are unused in mode 0. */
int gaot_rollout_update(const float* pred, const float* u_in, float* u_phys, float* u_next, int64_t rows, int64_t C,
                        float dt, int mode, const float* u_mean, const float* u_std,
                        const float* s_mean, const float* s_std, void* stream);

/* Time-conditioned affine correction (ConditionedNorm.forward, src/model/layers/mlp.py:121-124; attn.py:89-90, 153-154):
 * y[b,t,:] = x[b,t,:] * scale[b,:] + bias[b,:] with per-sample scale / bias [B, D]; the backward writes dx and the per-sample
 * sums dscale[b,:] = sum_t dy * x, dbias[b,:] = sum_t dy. */
int gaot_cond_affine_fwd(const float* x, const float* scale, const float* bias, float* y, int64_t B, int64_t T, int64_t D,
                         void* stream);
int gaot_cond_affine_bwd(const float* dy, const float* x, const float* scale, float* dx, float* dscale, float* dbias,
                         int64_t B, int64_t T, int64_t D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GAOT_B200_H */
