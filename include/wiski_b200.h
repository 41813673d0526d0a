/*
 * wiski_b200.h — C ABI of the B200-native WISKI online-update hot path (libwiski_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The reference (wjmaddox/online_gp) is
 * pure Python and issues every operation below as a chain of PyTorch/GPyTorch library calls; each entry point
 * cites the reference interface it replaces (paths relative to the reference repo; "App. A" = GPyTorch semantics
 * summarised in SURVEY.md Appendix A, source not vendored by the reference).
 *
 * Conventions
 *   - every function returns 0 on success; non-zero: 1 invalid argument, 2 CUDA error, 3 unsupported shape.
 *     wiski_last_error() gives the message for the calling thread.  Functions never allocate or free caller memory.
 *   - pointers are DEVICE pointers unless the name starts with h_ (host).  Matrices are row-major, contiguous.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises unless stated.
 *   - _f32 / _f64 variants compute in float / double ("dtype" in bench.py).  Indices are int64.
 *   - panels: L, B, KL are m x r (m = prod g_i inducing points, r = root rank); grid axis 0 is the slowest.
 */
#ifndef WISKI_B200_H
#define WISKI_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WISKI_MAX_DIMS 8

const char* wiski_last_error(void);
int wiski_abi_version(void);
/* number of CUDA kernels this library has enqueued so far in this process (bench.py's gpu_launches) */
long long wiski_launch_count(void);

/* ---- k1: cubic-convolution interpolation stencils.
 * Replaces GPyTorch Interpolation.interpolate (App. A.1) reached from
 * online_gp/models/batched_fixed_noise_online_gp.py:143,205,261 and online_gp/mlls/streaming_partial_mll.py:20.
 * x [q,d] -> idx [q,4^d] (flat C-order index, dim 0 slowest), val [q,4^d].
 * h_lo[i]=grid_i[0], h_delta[i]=clamp_min(grid_i[1]-grid_i[0],1e-10) (computed by the caller in the grid's dtype),
 * h_first4/h_last4 [d*4] = first/last four grid points per dim (boundary branch), h_gmin/h_gmax = grid extrema.
 * oob_flag (device int, may be NULL) is OR-ed with 1 if any x lies outside [gmin-1e-7, gmax+1e-7]: the host wrapper
 * turns that into GPyTorch's "out of bounds for the specified grid" RuntimeError. */
int wiski_interp_fwd_f32(const float* x, int64_t q, int d, const int64_t* h_g, const float* h_lo,
                         const float* h_delta, const float* h_first4, const float* h_last4, const float* h_gmin,
                         const float* h_gmax, int64_t* idx, float* val, int* oob_flag, void* stream);
int wiski_interp_fwd_f64(const double* x, int64_t q, int d, const int64_t* h_g, const double* h_lo,
                         const double* h_delta, const double* h_first4, const double* h_last4, const double* h_gmin,
                         const double* h_gmax, int64_t* idx, double* val, int* oob_flag, void* stream);
/* d val / d x contracted with grad_val [q,4^d] -> grad_x [q,d] (autograd of A.1 through `lower_pt_rel_dists`;
 * needed by the stem update, online_gp/models/online_ski_regression.py:148-162). */
int wiski_interp_bwd_f32(const float* x, int64_t q, int d, const int64_t* h_g, const float* h_lo,
                         const float* h_delta, const float* h_first4, const float* h_last4, const float* grad_val,
                         float* grad_x, void* stream);
int wiski_interp_bwd_f64(const double* x, int64_t q, int d, const int64_t* h_g, const double* h_lo,
                         const double* h_delta, const double* h_first4, const double* h_last4, const double* grad_val,
                         double* grad_x, void* stream);

/* ---- k13 / k5: W @ src  (GPyTorch left_interp, App. A.1; batched_fixed_noise_online_gp.py:206-210,236-240).
 * out[n,:] = sum_k val[n,k] * src[idx[n,k],:], src [m,c], out [q,c].  With src = inverse-root panel B and
 * val = w / sqrt(D) this is the projection p^T = v^T B of updated_root_lazy_tensor.py:79. */
int wiski_gather_f32(const int64_t* idx, const float* val, int64_t q, int64_t s, const float* src, int64_t m,
                     int64_t c, float* out, void* stream);
int wiski_gather_f64(const int64_t* idx, const double* val, int64_t q, int64_t s, const double* src, int64_t m,
                     int64_t c, double* out, void* stream);
/* ---- k2 / k3: dst += W^T @ src without densifying W^T (replaces _sparse_left_interp_t(...).to_dense() matmuls,
 * batched_fixed_noise_online_gp.py:22-28,42-47,158-160).  src [q,c], dst [m,c] accumulated in place (atomics). */
int wiski_scatter_add_f32(const int64_t* idx, const float* val, int64_t q, int64_t s, const float* src, int64_t m,
                          int64_t c, float* dst, void* stream);
int wiski_scatter_add_f64(const int64_t* idx, const double* val, int64_t q, int64_t s, const double* src, int64_t m,
                          int64_t c, double* dst, void* stream);

/* ---- k9: Kronecker-Toeplitz MVM  Y = (T(col_0) x ... x T(col_{d-1})) X,  X,Y [m,c]
 * (replaces KroneckerProductLazyTensor._matmul / ToeplitzLazyTensor._matmul, App. A.4, reached from
 * batched_fixed_noise_online_gp.py:348,366 and streaming_partial_mll.py:29).  cols [d,gmax] device (row i holds
 * col_i in its first g_i entries; any 1/sigma^2 scaling is folded into a column by the caller),
 * work = scratch [m,c]; X, Y, work must not alias. */
int wiski_kron_toeplitz_mm_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, const float* X, int64_t c,
                               float* Y, float* work, void* stream);
int wiski_kron_toeplitz_mm_f64(const double* cols, int d, const int64_t* h_g, int64_t gmax, const double* X,
                               int64_t c, double* Y, double* work, void* stream);
/* ---- k15: gradient of sum(Z * (K X)) w.r.t. the Toeplitz columns (replaces the autograd of k9 /
 * _quad_form_derivative; online_gp/models/online_ski_regression.py:141).
 * grad_cols [d,gmax] is overwritten.  work = scratch of wiski_kron_toeplitz_bwd_work_elems(...) elements
 * ((d+1) panels [m,c] + a double accumulator), elem_size = 4 or 8. */
int64_t wiski_kron_toeplitz_bwd_work_elems(int d, int64_t m, int64_t c, int64_t gmax, int elem_size);
int wiski_kron_toeplitz_bwd_cols_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, const float* Z,
                                     const float* X, int64_t c, float* grad_cols, float* work, void* stream);
int wiski_kron_toeplitz_bwd_cols_f64(const double* cols, int d, const int64_t* h_g, int64_t gmax, const double* Z,
                                     const double* X, int64_t c, double* grad_cols, double* work, void* stream);

/* Single-axis building blocks of the two entry points above, exported for the row-sharded multi-GPU path
 * (each rank applies the local grid axes to its slab and the sharded axis in a column-sharded layout).
 * View X as [outer, g, inner]: Y[o,a,w] = sum_b col[|a-b|] X[o,b,w];  X != Y.
 * contract: acc64[k] += sum_{o,w} sum_{|a-b|=k} Z[o,a,w] P[o,b,w]   (acc64: g doubles, accumulated, caller zeroes). */
int wiski_kron_axis_apply_f32(const float* X, float* Y, const float* col, int64_t g, int64_t outer, int64_t inner,
                              void* stream);
int wiski_kron_axis_apply_f64(const double* X, double* Y, const double* col, int64_t g, int64_t outer, int64_t inner,
                              void* stream);
int wiski_kron_axis_contract_f32(const float* Z, const float* P, int64_t g, int64_t outer, int64_t inner, double* acc64,
                                 void* stream);
int wiski_kron_axis_contract_f64(const double* Z, const double* P, int64_t g, int64_t outer, int64_t inner,
                                 double* acc64, void* stream);

/* Tensor-core form of the same two building blocks for axes with g >= 64 points (fp32; g % 4 == 0, inner % 4 == 0,
 * inner >= 64): an axis of that size is a real contraction (2 g flop per element), so it runs as a batched tcgen05
 * GEMM with 3xTF32 split operands — apply: Y[o] = T X[o] with the dense symmetric Toeplitz factor built in `work`;
 * contract: S = sum_o Z[o] P[o]^T (g x g, K-sliced partials in `work`), then acc64[k] += sum_{|a-b|=k} S[a][b].
 * work = scratch of wiski_kron_axis_tc_work_elems(g, outer, inner, contract) floats (0 = shape not supported; use the
 * SIMT entry points above).  This is what replaces ToeplitzLazyTensor's FFT MVM (App. A.4) on 128^3 / 256^2 / 1024^2 grids. */
int64_t wiski_kron_axis_tc_work_elems(int64_t g, int64_t outer, int64_t inner, int contract);
int wiski_kron_axis_apply_tc_f32(const float* X, float* Y, const float* col, int64_t g, int64_t outer, int64_t inner,
                                 float* work, void* stream);
int wiski_kron_axis_contract_tc_f32(const float* Z, const float* P, int64_t g, int64_t outer, int64_t inner,
                                    double* acc64, float* work, void* stream);

/* Fused fast path of k9 / k15 for grids whose axes all have 32 points (fp32, d even, c % 16 == 0; BASELINE config 2):
 * two axes per pass, grid tiles staged in shared memory.  wiski_kron_toeplitz_mm_f32 dispatches to it by itself;
 * the pair-level entry points let the autograd layer keep the intermediate panel of the forward pass.
 *   pair p covers grid axes (2p, 2p+1);   pair_apply: Y = (T_2p x T_2p+1) X;
 *   pair_grad: acc_u += contract_{2p}(Z, T_2p+1 P), acc_v += contract_{2p+1}(T_2p Z, P), Zout = T_2p+1 T_2p Z (or NULL).
 * The SIMT form is not re-entrant across streams (coefficients live in __constant__ memory); the tensor-core form is. */
int wiski_kron_fused_supported(int d, const int64_t* h_g, int64_t c);
int wiski_kron_fused_pair_apply_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, int pair, const float* X,
                                    float* Y, int64_t c, void* stream);
int wiski_kron_fused_pair_grad_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, int pair, const float* Z,
                                   const float* P, float* Zout, int64_t c, double* acc_u64, double* acc_v64,
                                   void* stream);

/* The same two passes with explicit operand layouts, for the row <-> column exchange of the row-sharded multi-GPU
 * path: h_lay holds (ld, cw, cstride) per operand — X, Y for pair_apply; Z, P, Zout for pair_grad — where element
 * (row, col) of the operand lives at ptr + (col / cw) * cstride + row * ld + col % cw.  Plain row-major is
 * (c, c, 0); "column-chunked" (cw = c / world, ld = cw, cstride = rows * cw) is the send / receive buffer of an
 * all-to-all, so no transposing copy is needed on either side of the collective.  cw % 16 == 0, ld, cstride % 4 == 0. */
int wiski_kron_fused_pair_apply_lay_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, int pair,
                                        const float* X, float* Y, int64_t c, const int64_t* h_lay, void* stream);
int wiski_kron_fused_pair_grad_lay_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, int pair,
                                       const float* Z, const float* P, float* Zout, int64_t c, double* acc_u64,
                                       double* acc_v64, const int64_t* h_lay, void* stream);

/* Directional form of pair_grad: with dirs [d,gmax] = d col_i / d lengthscale_i it returns, accumulated into out3
 * (3 doubles):  <grad_{2p}, dirs_{2p}>, <grad_{2p+1}, dirs_{2p+1}> and <Z', K' P'> (= <grad_i, col_i> for every i),
 * which is all a stationary product kernel with one lengthscale per dimension and scalar scales needs; it replaces
 * the two 1024-FMA contractions per grid line by one 512-FMA application of the direction matrix and a dot. */
int wiski_kron_fused_pair_grad_dir_f32(const float* cols, const float* dirs, int d, const int64_t* h_g, int64_t gmax,
                                       int pair, const float* Z, const float* P, float* Zout, int64_t c, double* out3,
                                       void* stream);

/* Directional pass with explicit operand layouts (h_lay: (ld, cw, cstride) for Z, P, Zout as for pair_grad_lay). */
int wiski_kron_fused_pair_grad_dir_lay_f32(const float* cols, const float* dirs, int d, const int64_t* h_g, int64_t gmax,
                                           int pair, const float* Z, const float* P, float* Zout, int64_t c, double* out3,
                                           const int64_t* h_lay, void* stream);
/* The pair_apply / pair_grad_dir entry points above run on the tensor pipe by default (csrc/kron_tc.cu: tcgen05.mma
 * kind::tf32 with 3xTF32 split operands, the panel tile as the TMEM-resident A operand, TMA loads and stores through
 * 5-D tensor maps) and fall back to the SIMT kernels only for shapes / layouts the tensor maps cannot describe.
 * wiski_kron_tc_enable(0 / 1) switches the tensor-core form off / on (returns the previous setting; the environment
 * variable WISKI_KRON_TC=0 sets the initial state) — used by the A/B timings of bench.py and the parity tests. */
int wiski_kron_tc_enable(int on);      /* on < 0: query only */

/* "Background" launches: while the flag is set, the HBM-bound panel passes (rank-q update, skinny Gram / panel product)
 * are launched with a small resident footprint (1 - 2 CTAs per SM) and the largest shared-memory carve-out, so that the
 * host layer can put them on a side stream UNDER a tensor-bound kernel of the main stream (settings.overlap_root_update):
 * the persistent tcgen05 kernels still find their block slots, and no SM has to drain to change its L1 / shared split.
 * Returns the previous state; on < 0 queries only.  Process-global (one host thread per process). */
int wiski_set_background(int on);

/* The two pair passes for ANY two 32-point grid axes axis_u < axis_v (the remaining axes are batch indices; at most two
 * of the three index groups before / between / after the pair may be non-trivial — always true for d <= 4).
 * For a 32^4 grid the host layer pairs the axes as (1,2) + (0,3) instead of (0,1) + (2,3): a tile of the pair (0,1) is
 * 1024 rows 1024 rows apart (one 64-byte piece per 2 MB page), which the memory system serves at half the rate of the
 * 32-consecutive-row / 32-rows-apart tiles of (0,3) / (1,2) (profiles/r02_pair_kernels.md).  Adjacent even pairs fall back
 * to the SIMT kernels when the tensor-core form is off; other axes need it (status 3 otherwise).
 * h_lay may be NULL (plain row-major operands); out3 as in wiski_kron_fused_pair_grad_dir_f32 (axis_u, axis_v, scale). */
int wiski_kron_pair_apply_axes_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, int axis_u, int axis_v,
                                   const float* X, float* Y, int64_t c, const int64_t* h_lay, void* stream);
int wiski_kron_pair_grad_dir_axes_f32(const float* cols, const float* dirs, int d, const int64_t* h_g, int64_t gmax,
                                      int axis_u, int axis_v, const float* Z, const float* P, float* Zout, int64_t c,
                                      double* out3, const int64_t* h_lay, void* stream);

/* ---- row-sharded multi-GPU path: compute + layout exchange in ONE kernel over NVLink peer memory.
 * The reference has no multi-device path (SURVEY.md §8 row e); the sharded model keeps the panels row-sharded along
 * grid axis 0 and needs all rows of a column block for the axis pair that contains axis 0.  Instead of an all-to-all
 * pass between the kernels, the producing kernel's own stores go to the consumers' buffers: dst[j] is a device pointer
 * into rank j's peer-mapped receive buffer (torch.distributed._symmetric_memory / cudaIpc), at the place where THIS
 * rank's part starts.  Callers order producer and consumer with a device-side barrier over all ranks.
 *   mode 1 (row slab -> column blocks): X [m_loc, c]; column block j (c / n_dst columns) -> dst[j], a [m_loc, c / n_dst] panel
 *   mode 2 (column block -> row slabs): X [m, c], axis_u = 0; rows of axis-0 range j -> dst[j], a [m / n_dst, c] panel
 *   mode 3 = mode 2 with 32-column store boxes (c % 32 == 0; apply pass only): 128-byte row pieces instead of 64-byte ones,
 *            which NVLink carries 1.6x faster
 * Tensor-core kernels only (status 3 when the shape is not theirs: 32-point axes, c % 16 == 0, n_dst <= 8).
 * h_lay_x / h_lay_zp: (ld, cw, cstride) of the inputs (X; Z then P) or NULL for plain row-major panels. */
int wiski_kron_pair_apply_push_f32(const float* cols, int d, const int64_t* h_g, int64_t gmax, int axis_u, int axis_v,
                                   const float* X, int64_t c, const int64_t* h_lay_x, float* const* dst, int n_dst, int mode,
                                   void* stream);
int wiski_kron_pair_grad_dir_push_f32(const float* cols, const float* dirs, int d, const int64_t* h_g, int64_t gmax,
                                      int axis_u, int axis_v, const float* Z, const float* P, int64_t c, double* out3,
                                      const int64_t* h_lay_zp, float* const* dst, int n_dst, void* stream);
/* Out = P @ M with column block j written at dst[j] (mode 1 of the above for the panel GEMM); work: the scratch of
 * wiski_panel_rmul_ex_f32 (terms = 2 only, else NULL). */
int wiski_panel_rmul_push_f32(const float* P, int64_t m, int64_t r, const float* M, int64_t r2, int terms, float* const* dst,
                              int n_dst, float* work, void* stream);

/* ---- k7: panel right-multiply  Out = P @ M,  P,Out [m,r], M [r,r2], Out [m,r2]  (Out must not alias P)
 * (replaces current_root.matmul(inner_root) / current_inv_root^T.matmul(inner_inv_root),
 * updated_root_lazy_tensor.py:97-100,115-117, and Kuu_Lmat @ qmat_solve, batched_fixed_noise_online_gp.py:376). */
int wiski_panel_rmul_f32(const float* P, int64_t m, int64_t r, const float* M, int64_t r2, float* Out, void* stream);
int wiski_panel_rmul_f64(const double* P, int64_t m, int64_t r, const double* M, int64_t r2, double* Out,
                         void* stream);
/* In-place, row-local form of the same root update for q << r:  P <- P + (P @ U) @ Vt,  U [r,q], Vt [q,r]
 * ((I + p p^T)^(+-1/2) = I + P f(S) P^T; identical L L^T / B B^T as :97-117, without the r x r GEMM). q <= 32. */
int wiski_panel_lowrank_update_f32(float* P, int64_t m, int64_t r, const float* U, const float* Vt, int64_t q,
                                   void* stream);
int wiski_panel_lowrank_update_f64(double* P, int64_t m, int64_t r, const double* U, const double* Vt, int64_t q,
                                   void* stream);

/* Both panels of the rank-q root update in ONE launch: P0 <- P0 + (P0 @ U) @ Vt0, P1 <- P1 + (P1 @ U) @ Vt1 (root L
 * and inverse root B share U = p = B^T v and differ in Vt = C p^T / C' p^T).  U and both Vt are staged in shared
 * memory and a warp updates several rows at a time, so the pass stays HBM-bound for q up to 32 (the per-row form above
 * re-reads the coefficients from L1 for every row).  P1 / Vt1 may be NULL. */
int wiski_panel_lowrank_update2_f32(float* P0, float* P1, int64_t m, int64_t r, const float* U, const float* Vt0,
                                    const float* Vt1, int64_t q, void* stream);
int wiski_panel_lowrank_update2_f64(double* P0, double* P1, int64_t m, int64_t r, const double* U, const double* Vt0,
                                    const double* Vt1, int64_t q, void* stream);
/* The same launch with a by-product: Tout [m,q] = P0 U of the rows BEFORE the update (the vector the dual-layout sharded
 * model all-gathers to update its column-sharded copy) — saves a separate pass over P0. */
int wiski_panel_lowrank_update2_t_f32(float* P0, float* P1, int64_t m, int64_t r, const float* U, const float* Vt0,
                                      const float* Vt1, int64_t q, float* Tout, void* stream);
int wiski_panel_lowrank_update2_t_f64(double* P0, double* P1, int64_t m, int64_t r, const double* U, const double* Vt0,
                                      const double* Vt1, int64_t q, double* Tout, void* stream);
/* P[m,c] += T[m,q] W[q,c] in place (q <= 32), one streaming pass.  Used by the row-sharded model with the dual layout: the
 * column-sharded copy of the root panel follows the rank-q update `collect_vector` (updated_root_lazy_tensor.py:97-100)
 * through T = L p, gathered from all ranks, and W = the local columns of C p^T. */
int wiski_panel_outer_add_f32(float* P, int64_t m, int64_t c, const float* T, int64_t q, const float* W, void* stream);
int wiski_panel_outer_add_f64(double* P, int64_t m, int64_t c, const double* T, int64_t q, const double* W, void* stream);

/* ---- k10: Gram  G = A^T @ Bm,  A [m,r], Bm [m,r2], G [r,r2]  (Q - I = L^T (K L), and L^T (K b);
 * batched_fixed_noise_online_gp.py:352-355,360-361).  work = scratch of wiski_gram_work_elems(m,r,r2) elements. */
int64_t wiski_gram_work_elems(int64_t m, int64_t r, int64_t r2);
int wiski_gram_f32(const float* A, const float* Bm, int64_t m, int64_t r, int64_t r2, float* G, float* work,
                   void* stream);
int wiski_gram_f64(const double* A, const double* Bm, int64_t m, int64_t r, int64_t r2, double* G, double* work,
                   void* stream);

/* Gram whose result the caller knows to be symmetric (Q - I = L^T (K L) with K symmetric; r2 = r): tiles of the result
 * that lie entirely below the diagonal are not computed but mirrored from the upper triangle (one quarter of the tensor
 * work at r = 432).  Same scratch as wiski_gram_f32. */
int wiski_gram_sym_f32(const float* A, const float* Bm, int64_t m, int64_t r, float* G, float* work, void* stream);
/* Panel right-multiply with explicit options: nblk > 1 writes Out column-chunked (as wiski_panel_rmul_chunked_f32);
 * terms = 3: 3xTF32 split products of both operands (fp32-grade, what wiski_panel_rmul_f32 does);
 * terms = 2: the small operand M exact (M and its fp32 remainder after tf32 truncation stacked along K in `work`), the
 *            panel truncated to tf32 by the tensor core — its error is independent from row to row and averages out of
 *            sums over grid rows, while an error in M would be shared by every row;
 * terms = 1: ONE kind::tf32 product of the raw operands (relative error ~1e-3).
 * The host layer uses terms < 3 only for gradient quantities (L grad_Q feeding the hyper-parameter gradient,
 * online_gp/models/online_ski_regression.py:141), never for values (settings.backward_gemm_tf32_passes).
 * work: wiski_panel_rmul_ex_work_elems(r, r2) floats, needed for terms = 2 (may be NULL otherwise). */
int64_t wiski_panel_rmul_ex_work_elems(int64_t r, int64_t r2);
int wiski_panel_rmul_ex_f32(const float* P, int64_t m, int64_t r, const float* M, int64_t r2, int64_t nblk, int terms,
                            float* Out, float* work, void* stream);

/* Column-chunked forms of k10 / k7 for the row-sharded multi-GPU path, where K L and its gradient are kept as nblk
 * column blocks [m, r2 / nblk] (block j = columns [j r2/nblk, (j+1) r2/nblk)) — the receive / send buffers of the
 * row <-> column all-to-all:  G = A^T [B_0 | B_1 | ...]  and  [Out_0 | Out_1 | ...] = P M  in one tensor-core launch.
 * fp32, (r2 / nblk) % 32 == 0; returns 3 for shapes the tensor-core path does not take (callers then go block by
 * block through wiski_gram_f32 / wiski_panel_rmul_f32).  work as for wiski_gram_f32. */
int wiski_gram_chunked_f32(const float* A, const float* Bb, int64_t m, int64_t r, int64_t r2, int64_t nblk, float* G,
                           float* work, void* stream);
int wiski_panel_rmul_chunked_f32(const float* P, int64_t m, int64_t r, const float* M, int64_t r2, int64_t nblk,
                                 float* Outb, void* stream);
/* wiski_gram_chunked_f32 for a product the caller knows to be symmetric (see wiski_gram_sym_f32). */
int wiski_gram_chunked_sym_f32(const float* A, const float* Bb, int64_t m, int64_t r, int64_t nblk, float* G, float* work,
                               void* stream);

/* ---- k11 (CG path): fused Q-MVM  w = v + L^T (KL v),  v,w [r,c]  — one pass over both panels
 * (the matmul closure GPyTorch's linear_cg calls when r > max_cholesky_size, App. A.5).
 * work = scratch of wiski_qmv_work_elems(m,r,c) elements. */
int64_t wiski_qmv_work_elems(int64_t m, int64_t r, int64_t c);
int wiski_q_matvec_f32(const float* L, const float* KL, int64_t m, int64_t r, const float* v, int64_t c, float* w,
                       float* work, void* stream);
int wiski_q_matvec_f64(const double* L, const double* KL, int64_t m, int64_t r, const double* v, int64_t c,
                       double* w, double* work, void* stream);
/* Conjugate gradients on Q x = rhs (rhs [r,c], x [r,c]) with the fused MVM above; GPyTorch linear_cg stopping
 * rule (App. A.5): columns normalised, stop when mean residual norm < tol and it >= min(10, max_iter-1).
 * Synchronises the stream every `check_every` iterations to read the residual.  h_iters / h_resid are host outputs.
 * work = scratch of wiski_cg_work_elems(m,r,c) elements. */
int64_t wiski_cg_work_elems(int64_t m, int64_t r, int64_t c);
int wiski_cg_solve_f32(const float* L, const float* KL, int64_t m, int64_t r, const float* rhs, int64_t c,
                       float tol, int max_iter, int check_every, float* x, int* h_iters, float* h_resid,
                       float* work, void* stream);
int wiski_cg_solve_f64(const double* L, const double* KL, int64_t m, int64_t r, const double* rhs, int64_t c,
                       double tol, int max_iter, int check_every, double* x, int* h_iters, double* h_resid,
                       double* work, void* stream);

#ifdef __cplusplus
}
#endif
#endif
