/*
 * socks_b200.h — C ABI of libsocksb200.so: fp64 CUDA kernels (sm_100a) for SOCKS's
 * kernel-embedding hot path.
 *
 * The reference (ajthor/socks) is pure Python and has no FFI; what each entry point replaces
 * is therefore a span of reference Python, cited as gym_socks/<file>:<lines> below.  The
 * binding a maintainer adds on the reference side is a ctypes stub.
 *
 * Conventions
 *  - Every function returns int: 0 = ok, SOCKS_EINVAL(-k) = argument k invalid (1-based),
 *    SOCKS_ECUDA = a CUDA call failed; socks_last_error() returns a thread-local message.
 *  - All data pointers are DEVICE pointers to IEEE float64, row-major, 16-byte aligned.
 *    The caller owns every buffer, including workspaces (sizes via the *_bytes helpers).
 *  - `stream` is a cudaStream_t passed as void*; calls only enqueue work (asynchronous).
 *  - Square factorisation operands are PADDED: a logical n x n matrix lives in an
 *    np x np buffer, np = socks_padded_dim(n) (next multiple of SOCKS_TILE), leading
 *    dimension ld >= np, ld even.  socks_gram_sym fills the padding with the identity, so
 *    chol/inverse of the padded matrix restricted to [0,n) is the chol/inverse of the matrix.
 *  - Kernel: k(x,y) = exp(|x-y|^2 * scale) if divide == 0, exp(|x-y|^2 / scale) otherwise.
 *    gym_socks-style callers pass (scale = -2 sigma^2, divide = 1): the reference DIVIDES by -2 sigma^2
 *    (gym_socks/kernel/metrics.py:135-136) and the Gram builders reproduce that rounding exactly;
 *    sklearn-style callers pass (scale = -gamma, divide = 0).
 */
#ifndef SOCKS_B200_H
#define SOCKS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOCKS_TILE 128
#define SOCKS_OK 0
#define SOCKS_ECUDA 1000
#define SOCKS_EINVAL(k) (-(k))
#define SOCKS_MAX_DIM 32     /* largest state/action dimension accepted by the Gram kernels */
#define SOCKS_SR_ROWS 16     /* coefficient rows (1 normaliser + 15 time steps) per predict launch */

#define SOCKS_KERNEL_RBF 0   /* exp(|x-y|^2 (*|/) scale), gym_socks/kernel/metrics.py:69-138  */
#define SOCKS_KERNEL_ABEL 1  /* exp(|x-y|   (*|/) scale), gym_socks/kernel/metrics.py:174-206 */
/* sklearn.metrics.pairwise kernels the reference accepts through its `kernel_fn` protocol (metrics.py:239-241; pinned
 * by gym_socks/kernel/tests/test_kernel.py:18,132-147) */
#define SOCKS_KERNEL_LINEAR 2     /* x.y                                                    */
#define SOCKS_KERNEL_POLY 3       /* (scale x.y + coef0)^degree          (scale = gamma)    */
#define SOCKS_KERNEL_LAPLACIAN 4  /* exp(scale |x-y|_1)                  (scale = -gamma)   */

#define SOCKS_PROBLEM_THT 0  /* gym_socks/algorithms/reach/common.py:42-69 */
#define SOCKS_PROBLEM_FHT 1  /* gym_socks/algorithms/reach/common.py:8-39  */

typedef void* socks_stream_t;

int socks_version(void);
const char* socks_last_error(void);
int64_t socks_padded_dim(int64_t n);
/* rows x width_bytes strided copy in one enqueue (kind: 1 host->device, 2 device->host, 3 device->device): the read
 * of a column block of the (T, M) result into the caller's (pinned) ndarray, kernel_sr.py:344 `return out`. */
/* out2[0] = min_i |A[i][i]|, out2[1] = max_i |A[i][i]| (i < n): on the diagonal of a Cholesky factor,
 * (max / min)^2 is a lower bound of cond_2(G). */
int socks_diag_range(const double* A, int64_t n, int64_t ld, double* out2, socks_stream_t stream);
int socks_copy2d(void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width_bytes, int64_t rows,
                 int kind, socks_stream_t stream);

/* ---- Gram builders: gym_socks/kernel/metrics.py:119-138 (rbf_kernel) ------------------- */

/* K[i*ldk + j] = rowscale[i] * k(X_i, Y_j), i<n, j<m.  rowscale may be NULL (=1).
 * With rowscale = diag(W) this is the un-normalised beta of
 * gym_socks/algorithms/reach/kernel_sr.py:45. */
int socks_gram_rbf(const double* X, int64_t n, const double* Y, int64_t m, int d, double scale, int divide,
                   const double* rowscale, double* K, int64_t ldk, socks_stream_t stream);

/* Same with the kernel family chosen by `kind` (SOCKS_KERNEL_RBF | SOCKS_KERNEL_ABEL); the Abel kernel
 * exp(-|x-y| / sigma) (metrics.py:174-206) is what SeparatingKernelClassifier uses (scale = -sigma, divide = 1). */
int socks_gram_kernel(int kind, const double* X, int64_t n, const double* Y, int64_t m, int d, double scale,
                      int divide, const double* rowscale, double* K, int64_t ldk, socks_stream_t stream);

/* D[i*ldd + j] = |X_i - Y_j|^2 (direct differences, metrics.py:127-128 and euclidean_distance :26-66 with
 * squared=True); used for the sigma=None median heuristic (metrics.py:130-131).  socks_dist: the square root
 * (euclidean_distance with squared=False; the Abel kernel's median heuristic, metrics.py:199-200). */
int socks_sqdist(const double* X, int64_t n, const double* Y, int64_t m, int d, double* D,
                 int64_t ldd, socks_stream_t stream);
int socks_dist(const double* X, int64_t n, const double* Y, int64_t m, int d, double* D,
               int64_t ldd, socks_stream_t stream);

/* G = K(Z,Z) + diag_add * I written into the LOWER 128-tiles of the padded np x np buffer
 * (identity in the padding).  Z = X, or [X | U] concatenated for the product Gram
 * K(X,X) * K(U,U) of metrics.py:268-276.  diag_add = lambda * N (metrics.py:278-280). */
int socks_gram_sym(const double* Z, int64_t n, int d, double scale, int divide, double diag_add, double* G,
                   int64_t ldg, socks_stream_t stream);
int socks_gram_sym_kernel(int kind, const double* Z, int64_t n, int d, double scale, int divide, double diag_add,
                          double* G, int64_t ldg, socks_stream_t stream);

/* Any kernel family above: K = rowscale * k(X, Y), and the padded G = k(X,X) [* k(U,U)] + diag_add I of
 * metrics.py:264-280 with the SAME kernel for states and actions (U may be NULL).  coef0 / degree are read by
 * SOCKS_KERNEL_POLY only.  RBF/Abel without U run the tiled kernels above. */
int socks_gram_family(int kind, const double* X, int64_t n, const double* Y, int64_t m, int d, double scale,
                      int divide, double coef0, double degree, const double* rowscale, double* K, int64_t ldk,
                      socks_stream_t stream);
int socks_gram_sym_family(int kind, const double* X, int64_t n, int dx, const double* U, int du, double scale,
                          int divide, double coef0, double degree, double diag_add, double* G, int64_t ldg,
                          socks_stream_t stream);

/* ---- Regularised solve: gym_socks/kernel/metrics.py:278-280 (inv(K + lambda N I)) ------ */

int64_t socks_potrf_work_bytes(int64_t n);   /* dinv: np * 128 doubles                       */
int64_t socks_trtri_work_bytes(int64_t n);   /* scratch of socks_trtri / socks_potrf_inv (for n > 2048: recursion scratch,
                                                 one T per large merge < np^2/2 doubles, digit operands: 4.4 GB at n = 20 000) */
int64_t socks_reduce_work_bytes(int64_t cols); /* partial sums for column reductions         */

/* In-place lower Cholesky G = L L^T of the padded matrix (blocked, recursive, DMMA trailing
 * updates).  dinv receives inv(L_kk) of every 128 x 128 diagonal block.  *info_dev (device
 * int, caller-zeroed) receives 1 + index of the first non-positive pivot, else stays 0. */
int socks_potrf(double* G, int64_t n, int64_t ldg, double* dinv, int* info_dev, socks_stream_t stream);

/* M = inv(L), lower triangular, out of place (M's diagonal 128-blocks get explicit zeros above
 * the diagonal; tiles strictly above the block diagonal are not touched and never read). */
int socks_trtri(const double* L, int64_t n, int64_t ldl, const double* dinv, double* M, int64_t ldm,
                double* work, socks_stream_t stream);

/* Gout (padded np x np, lower 128-tiles) = Gin (n x n) + diag_add * I with identity padding: entry for callers that
 * bring their own Gram matrix (ConditionalEmbedding.fit, gym_socks/algorithms/kernel.py:43-66). */
int socks_pad_spd(const double* Gin, int64_t n, int64_t ldin, double diag_add, double* Gout, int64_t ldout,
                  socks_stream_t stream);

/* Factor and invert in ONE recursion (what the algorithms use): on return M = inv(L) (as socks_trtri) and, if
 * keep_l != 0, G holds L (as socks_potrf); with keep_l == 0 only the diagonal 128-blocks of G hold L and the rest
 * of G is scratch.  The panel solves run as single triangular GEMMs against the already inverted diagonal
 * blocks (L21 = A21 inv(L11)^T), so there are no leaf TRSM kernels.  work: socks_trtri_work_bytes(n). */
int socks_potrf_inv(double* G, int64_t n, int64_t ldg, double* M, int64_t ldm, double* work, int* info_dev,
                    int keep_l, socks_stream_t stream);
/* socks_potrf_inv factors right-looking over 1024-wide block columns with a look-ahead of one on three internal
 * streams that the library owns, forks from `stream` and joins back into it before returning: the next diagonal block
 * and panel (greatest priority), the trailing updates (middle), and the merges of inv(L) (least priority; each merge
 * M21 = -M22 (L21 M11) is two products that start as soon as their inputs exist, so they fill what the factorisation
 * leaves idle).  The large GEMMs -- trailing updates and merges -- run as fp64-accurate digit GEMMs on the integer
 * tensor cores (tcgen05.mma kind::i8; 7 digits in the updates, 8 in the merges). */

/* w[i] = sum_k M[k][i]^2 = diag(inv(G))[i]  (the only part of W the SR algorithms use:
 * np.einsum("ii,ij->ij", W, CXY), kernel_sr.py:45).  work: socks_reduce_work_bytes(np). */
int socks_diag_inv(const double* M, int64_t n, int64_t ldm, double* w, double* work, socks_stream_t stream);

/* y[j] = sum_{i<n} Z[i][j]^2 (general n x m matrix).  With Z = inv(L) K(X,T) this is the diagonal of
 * K^T W K that SeparatingKernelClassifier thresholds (gym_socks/algorithms/reach/separating_kernel.py:88,109).
 * work: socks_reduce_work_bytes(m). */
int socks_colsumsq(const double* Z, int64_t n, int64_t m, int64_t ldz, double* y, double* work, socks_stream_t stream);

/* X = inv(G) B = M^T (M B), in place on B (np x nrhs, nrhs a multiple of 128, zero rows in the padding):
 * `np.linalg.solve(G + lambda M I, K)` of ConditionalEmbedding.predict (gym_socks/algorithms/kernel.py:81) and the
 * `W @ c` products of the control algorithms, without forming W.  work: np * nrhs doubles. */
int socks_potrs(const double* M, int64_t n, int64_t ldm, double* B, int64_t nrhs, int64_t ldb, double* work,
                socks_stream_t stream);

/* W = M^T M = inv(G), full symmetric np x np (what regularized_inverse returns). */
int socks_lauum(const double* M, int64_t n, int64_t ldm, double* W, int64_t ldw, socks_stream_t stream);
/* The same product as an fp64-accurate digit GEMM on the integer tensor cores (8 digits); work: socks_lauum_i8_work_bytes. */
int64_t socks_lauum_i8_work_bytes(int64_t n);
int socks_lauum_i8(const double* M, int64_t n, int64_t ldm, double* W, int64_t ldw, void* work, socks_stream_t stream);

/* ---- Stochastic reachability: gym_socks/algorithms/reach/kernel_sr.py:31-69 ------------ */

/* tubes: device array [T][4][d] = (constraint.low, constraint.high, target.low, target.high),
 * the Boxes' float32 bounds promoted to float64 (gym_socks/utils/__init__.py:42-47). */

/* Fit-phase recursion over a RESIDENT un-normalised beta B (n x m, B = diag(w) K(X,pts)):
 *   out[T-1] = 1_target(pts);  out[t] = step(pts, (vf[t+1] @ B) / colsum(B))   t = T-2..0
 * with vf == out (in-place, kernel_sr.py:297-309; then n == m) or vf = fitted values (n rows).
 * T-1 HBM-bound GEMV passes over B.  work: socks_reduce_work_bytes(m) + m doubles. */
int socks_sr_recursion(const double* B, int64_t n, int64_t m, int64_t ldb, const double* pts, int d,
                       const double* tubes, int problem, int T, const double* vf, int64_t ldvf,
                       double* out, int64_t ldout, double* work, socks_stream_t stream);

/* Fused predict (kernel_sr.py:329-342 without materialising K(X,pts)):
 *   out[t][j] = step(pts_j, sum_i vf[t+1][i] w_i k(x_i,pts_j) / sum_i w_i k(x_i,pts_j)).
 * Two steps: socks_sr_pack turns the fitted state (X, w = diag(inv G), vf) into the packed operand of the predict
 * kernels once per fit (centred samples, coefficient rows, admissibility of the factored evaluation);
 * socks_sr_predict_packed evaluates any number of test-point batches against it.  socks_sr_predict does both in one
 * call (pack at the head of `work`).
 * variant: 0 = automatic (factored exponent on the tensor cores, direct form for the columns / launches where a
 * factor could over- or underflow), 1 = FMA-pipe contraction with CUDA's exp, 2 = direct form on the tensor cores,
 * 3 / 4 = automatic with the CTA shape forced to 8 / 4 warps (measurement), 6 = the same contraction on the integer
 * tensor cores (tcgen05.mma kind::i8 over fixed-point digits of the kernel values; agrees with 0 to ~3e-14, slower; kept
 * as an independent cross-check).  The sample axis is cut into fixed
 * 2048-sample splits (4096 for variant 6) summed in order, and the form chosen for a column depends only on that column and the fitted
 * state, so each output column is bit-identical however the test points are chunked or sharded. */
int64_t socks_sr_pack_bytes(int64_t n, int d, int T);
int socks_sr_pack(const double* X, const double* w, const double* vf, int64_t ldvf, int64_t n, int d, double scale,
                  int divide, int T, double* pack, socks_stream_t stream);
int64_t socks_sr_predict_packed_work_bytes(int64_t n, int64_t m);
int socks_sr_predict_packed(const double* pack, const double* X, int64_t n, int d, double scale, int divide,
                            const double* pts, int64_t m, const double* tubes, int problem, int T, double* out,
                            int64_t ldout, double* work, int variant, socks_stream_t stream);
int64_t socks_sr_predict_work_bytes(int64_t n, int64_t m);
int socks_sr_predict(const double* X, const double* w, const double* vf, int64_t ldvf, int64_t n, int d,
                     double scale, int divide, const double* pts, int64_t m, const double* tubes, int problem, int T,
                     double* out, int64_t ldout, double* work, int variant, socks_stream_t stream);

/* ---- Maximal SR: gym_socks/algorithms/reach/kernel_sr_max.py:30-79 --------------------- */

/* Bm[i][c] = coef_r[i] * CUA[i][a] for c = (r - first_row)*P + a, first_row <= r < first_row + R; zero for
 * i >= n or c >= R*P.  coef row 0 = w (the normaliser, independent of t), rows r >= 1 = vf[t0 + r] * w.
 * Bm: np x ldbm (the GEMM's right operand). */
int socks_srmax_pack(const double* CUA, int64_t n, int P, const double* w, const double* vf,
                     int64_t ldvf, int t0, int first_row, int R, double* Bm, int64_t np, int64_t ldbm,
                     socks_stream_t stream);

/* C (m x nn) = A^T B with A = Kmat (K x m, row-major: the Gram chunk K(X, pts)), B = Bm (K x nn);
 * m, nn multiples of 128, K multiple of 16.  DMMA GEMM (kernel_sr_max.py:65-70 as a contraction). */
int socks_gemm_tn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                  int64_t m, int64_t nn, int64_t K, socks_stream_t stream);

/* General form of the same kernel: C = alpha * op(A) * op(B) + beta * C, all row-major.
 *   transa == 0: A is m x K;  transa != 0: A is K x m.   transb != 0: B is nn x K ("NT");  transb == 0: B is K x nn.
 *   (transa && transb is not provided.)  flags: bit0 A(i,k)=0 for k>i, bit1 A(i,k)=0 for k<i, bit2 B(k,j)=0 for
 *   k<j, bit3 B(k,j)=0 for k>j, all at 128-tile granularity; lower_only: skip 128-tiles above the block diagonal;
 *   tile_cfg: 0 = 64x64 CTA tiles x 3 CTAs/SM (default), 1 = 128x128.  These are the SYRK/GEMM/TRMM updates of socks_potrf,
 *   socks_trtri and socks_lauum (numpy.linalg.inv's LAPACK calls, gym_socks/kernel/metrics.py:280). */
int socks_gemm_f64(int transa, int transb, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                   int64_t ldc, int64_t m, int64_t nn, int64_t K, double alpha, double beta, int flags,
                   int lower_only, int tile_cfg, socks_stream_t stream);

/* out[t0 + r - 1][j] = step(pts_j, max_a num_r[j][a] / den[j][a]) for r = 1..R-1 (kernel_sr_max.py:72-73).
 * Den == NULL: Cm = [den | num_1 .. num_{R-1}] (column blocks of P);  Den != NULL: Cm = [num_1 .. num_{R-1}] and
 * Den (m x ldden) is the normaliser computed once.  If write_last != 0 also out[T-1][j] = 1_target(pts_j). */
int socks_srmax_step(const double* Cm, int64_t ldc, const double* Den, int64_t ldden, int64_t m, int P, int R,
                     int t0, const double* pts,
                     int d, const double* tubes, int problem, int T, double* out, int64_t ldout,
                     int write_last, socks_stream_t stream);

/* ---- Backward (dynamic-programming) control: gym_socks/algorithms/control/kernel_control_bwd.py:58-122 */

/* Bm[i][r*P + a] = coefs[r][i] * CUA[i][a] (r < R), zero padding to np x ldbm: right operand of
 * C = K(X,Y)^T diag(z) CUA for z = vf[t+1] @ W and z = constraint_t @ W (kernel_control_bwd.py:90-97). */
int socks_scale_pack(const double* CUA, int64_t n, int P, const double* coefs, int64_t ldcoef, int R,
                     double* Bm, int64_t np, int64_t ldbm, socks_stream_t stream);

/* out[j] = add[j] + min_a {Cm[j][a] : Cm[j][P + a] <= 0} (all a when unconstrained or none feasible)
 * (kernel_control_bwd.py:99-115). */
int socks_rowmin_masked(const double* Cm, int64_t ldc, int64_t m, int P, int constrained, const double* add,
                        double* out, socks_stream_t stream);

/* ---- Forward control: gym_socks/algorithms/control/kernel_control_fwd.py:211-238 ------- */

/* y[r][j] = sum_i V[r][i] * Bmat[i][j] (R <= 2 row vectors, V rows ldv apart; V NULL = ones).
 * Used for z = W c (W symmetric) and for scores = q^T CUA.  work: R*socks_reduce_work_bytes(m). */
int socks_gemv_t(const double* Bmat, int64_t n, int64_t m, int64_t ldb, const double* V, int64_t ldv,
                 int R, double* y, int64_t ldy, double* work, socks_stream_t stream);

/* q[r][i] = z[r][i] * k(X_i, s), r < R (kernel_control_fwd.py:220-222). */
int socks_fwd_weights(const double* X, int64_t n, int d, double scale, int divide, const double* state,
                      const double* z, int64_t ldz, int R, double* q, int64_t ldq, socks_stream_t stream);

/* idx = argmin_k C[k] over {k : D[k] <= 0} (all k if D is NULL or the set is empty); first
 * occurrence on ties (gym_socks/algorithms/control/common.py:96-100, 166-174). idx_dev: device int. */
int socks_argmin_masked(const double* C, const double* D, int P, int* idx_dev, socks_stream_t stream);

/* ---- Linear system identification: gym_socks/algorithms/identification/kernel_linear_id.py:133-181 ------ */

/* C (p x q, row stride ldc) = A^T B for tall-skinny A (n x p) and B (n x q), p, q <= 64: Z^T Z and Z^T Y of the primal
 * normal equations (:144-145) in one pass over the samples.  Deterministic (fixed 2048-row chunks summed in order). */
int64_t socks_atb_small_work_bytes(int64_t n, int p, int q);
int socks_atb_small(const double* A, int64_t lda, int p, const double* B, int64_t ldb, int q, int64_t n, double* C,
                    int64_t ldc, double* work, socks_stream_t stream);

/* out[i][j] = sum_k A[i][k] T[j][k] (+ sum_k B[i][k] U[j][k] if U != NULL), i < r, j < m:
 * state_matrix @ T.T + input_matrix @ U.T (:173-181). */
int socks_linear_map(const double* A, int lda, int dx, const double* B, int ldb, int du, int r, const double* T,
                     const double* U, int64_t m, double* out, int64_t ldout, socks_stream_t stream);

/* ---- Composite entry points: one enqueue per algorithm step of the reference's classes ------------------ */

/* KernelSR.fit after the inverse (kernel_sr.py:286-309): B = diag(w) K(X,Y) resident in `work`, in-place recursion into
 * vf (T x n, row stride ldvf); if pack != NULL also socks_sr_pack (the predict operand). */
int64_t socks_sr_fit_work_bytes(int64_t n);
int socks_sr_fit(const double* X, const double* Y, const double* w, int64_t n, int d, double scale, int divide,
                 const double* tubes, int problem, int T, double* vf, int64_t ldvf, double* pack, double* work,
                 socks_stream_t stream);

/* KernelMaximalSR.fit after the inverse (kernel_sr_max.py:323-366): CUA = K(U,A) (n x P, out), value functions vf
 * (T x n, out).  w = diag(inv(K(X,X) * K(U,U) + lambda N I)). */
int64_t socks_srmax_fit_work_bytes(int64_t n, int P);
int socks_srmax_fit(const double* X, const double* Y, const double* U, const double* A, const double* w, int64_t n,
                    int d, int du, int P, double scale, int divide, const double* tubes, int problem, int T,
                    double* CUA, double* vf, int64_t ldvf, double* work, socks_stream_t stream);
/* The same fit with the backward recursion's contractions on the integer tensor cores (csrc/ozaki.cu: kernel values and
 * coefficients as `slices` unsigned 7-bit digits, exact int8 MMAs).  *flagged (device int) receives the number of
 * samples whose normaliser is below `tau` of the scale, i.e. outside the fixed-point range: if it is not 0 the caller
 * must redo the fit with socks_srmax_fit.  work: 1024-byte aligned, socks_srmax_fit_i8_work_bytes. */
int64_t socks_srmax_fit_i8_work_bytes(int64_t n, int P, int slices);
int socks_srmax_fit_i8(const double* X, const double* Y, const double* U, const double* A, const double* w, int64_t n, int d,
                       int du, int P, double scale, int divide, const double* tubes, int problem, int T, double* CUA,
                       double* vf, int64_t ldvf, void* work, int slices, double tau, int* flagged, socks_stream_t stream);

/* KernelMaximalSR.predict (kernel_sr_max.py:368-384): out (T x m); test points are processed `chunk` at a time (one
 * Gram tile + one DMMA GEMM for all T rows + fused divide/max/step per chunk). */
int64_t socks_srmax_predict_work_bytes(int64_t n, int P, int T, int64_t chunk);
int socks_srmax_predict(const double* X, const double* w, const double* vf, int64_t ldvf, const double* CUA, int64_t n,
                        int d, int P, double scale, int divide, const double* pts, int64_t m, const double* tubes,
                        int problem, int T, double* out, int64_t ldout, double* work, int64_t chunk,
                        socks_stream_t stream);

/* The same predict with the contraction on the INTEGER tensor cores (tcgen05.mma kind::i8, TMEM accumulators): both
 * operands are split into `slices` (6..8) unsigned 7-bit digits and every digit product is an exact int8 GEMM
 * (Ozaki splitting); sums of non-negative terms, recombined in fp64: relative error of num and den ~ 2^(-7 slices) K.
 * Points whose smallest normaliser is below tau * 2^(e_0) (far from every sample: the fixed-point digits of K(x,p) carry
 * too few significant bits there) are NOT trustworthy: their indices are appended to flagged[1..], flagged[0] = count
 * (device ints, 1 + m entries), and the caller recomputes them with socks_srmax_predict.  work: 1024-byte aligned. */
int64_t socks_srmax_predict_i8_work_bytes(int64_t n, int P, int T, int64_t chunk, int slices);
int socks_srmax_predict_i8(const double* X, const double* w, const double* vf, int64_t ldvf, const double* CUA, int64_t n,
                           int d, int P, double scale, int divide, const double* pts, int64_t m, const double* tubes,
                           int problem, int T, double* out, int64_t ldout, void* work, int64_t chunk, int slices,
                           double tau, int* flagged, socks_stream_t stream);

/* KernelControlFwd.__call__ (kernel_control_fwd.py:217-232): scores[r][a] = c_r^T W (K(X,s) * CUA)[:, a] for the R <= 2
 * row vectors c (float32-rounded cost / constraint values as float64, row stride ldc); W = inv(G) (n x n, ld ldw);
 * scores is R x P contiguous.  socks_fwd_policy adds the heuristic selection of control/common.py:96-100,166-174
 * (idx_dev: device int). */
int64_t socks_fwd_scores_work_bytes(int64_t n, int P);
int socks_fwd_scores(const double* W, int64_t ldw, const double* X, const double* CUA, int64_t n, int d, int P,
                     double scale, int divide, const double* state, const double* c, int64_t ldc, int R,
                     double* scores, double* work, socks_stream_t stream);
int socks_fwd_policy(const double* W, int64_t ldw, const double* X, const double* CUA, int64_t n, int d, int P,
                     double scale, int divide, const double* state, const double* c, int64_t ldc, int R,
                     double* scores, double* work, int* idx_dev, socks_stream_t stream);

/* *result_dev = np.median of the n x m matrix of squared (squared != 0) or plain Euclidean distances: the sigma=None
 * bandwidth of rbf_kernel (metrics.py:130-131) / abel_kernel (:199-200).  Exact order statistics by radix select. */
int64_t socks_median_work_bytes(int64_t n, int64_t m);
int socks_median_sq_dist(const double* X, int64_t n, const double* Y, int64_t m, int d, int squared, double* work,
                         double* result_dev, socks_stream_t stream);

/* ---- Context and collectives (optional: every entry point above is stateless and takes the caller's stream) ----
 * A context owns a non-blocking stream, a grow-only device workspace and, after socks_comm_init, an NCCL
 * communicator (libnccl.so.2 is bound at run time with dlopen; there is no link-time dependency).  One context per
 * GPU / rank, used from one thread.  socks_bcast / socks_allgather enqueue on the context stream: the single
 * broadcast of the fitted state (X, w, value functions -- or W for the control algorithms) and the gather of the
 * result shards are the only collectives the path has. */
typedef struct socks_ctx socks_ctx;
int socks_ctx_create(int device, socks_ctx** out);
int socks_ctx_destroy(socks_ctx* ctx);
socks_stream_t socks_ctx_stream(socks_ctx* ctx);
int socks_ctx_workspace(socks_ctx* ctx, int64_t bytes, void** ptr);
int socks_ctx_synchronize(socks_ctx* ctx);
int socks_comm_unique_id(void* id128);                                  /* 128 bytes, made on one rank, sent to all  */
int socks_comm_init(socks_ctx* ctx, int rank, int nranks, const void* id128);
int socks_bcast(socks_ctx* ctx, void* ptr, int64_t bytes, int root);    /* in place                                   */
int socks_allgather(socks_ctx* ctx, const void* send, void* recv, int64_t bytes_per_rank);

#ifdef __cplusplus
}
#endif
#endif
