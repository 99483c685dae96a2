/*
 * pdegram.h -- C ABI of libpdegram.so: the B200 (sm_100a) implementation of the
 * FD -> library -> (block average) -> Gram -> STRidge hot path of
 * anpeata/pde-discovery-laser-matter.
 *
 * The reference is pure Python and has no FFI; the boundary it offers is the set of
 * Python functions named below (paths relative to the reference root: "ks2d" =
 * scripts/ks2d_stridge_benchmark.py, "basic" = examples/basic_usage.py, "patch" =
 * scripts/patch_based_pde_discovery.py).  Each entry point cites the reference function
 * it serves.  The Python package pde_b200 binds these with ctypes and re-exposes the
 * reference signatures.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns
 *     all buffers; the library keeps only a private per-device scratch (pg_shutdown frees it)
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default)
 *   - return 0 on success, a negative PG_E* code otherwise; pg_last_error() gives the
 *     thread-local message.  No C++ exception crosses this boundary.
 *   - fields are C-contiguous doubles U[t][a0][a1]; d0/d1 are the grid spacings along
 *     a0/a1.  ks2d calls a0 "x" (ks2d:70-73); basic calls a1 "x" (basic:58).
 *
 * Statistics vector of one fold ("stats"), length PG_STATS_LEN(p) doubles:
 *   [0] n  [1] sum y  [2] sum y^2  [3..3+p) sum theta_j  [3+p..3+2p) sum theta_j*y
 *   [3+2p..) upper triangle of Theta^T Theta, row-major (i <= j)
 */
#ifndef PDEGRAM_H
#define PDEGRAM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_VERSION 103

#if defined(__GNUC__)
#define PG_API __attribute__((visibility("default")))
#else
#define PG_API
#endif

#define PG_MAX_P 16
#define PG_MAX_FOLDS 8
#define PG_STATS_LEN(p) (3 + 2 * (p) + ((p) * ((p) + 1)) / 2)

/* error codes */
#define PG_OK 0
#define PG_EINVAL (-1)   /* bad argument (shape, enum, null pointer) */
#define PG_ECUDA (-2)    /* CUDA runtime / driver error */
#define PG_ENOMEM (-3)   /* scratch allocation failed */
#define PG_EUNSUPPORTED (-4) /* valid request this build has no kernel for */

/* finite-difference dialect */
#define PG_FD_KS_PERIODIC 0 /* ks2d:63-73   periodic np.roll stencils, rows = all points of U[:-1] */
#define PG_FD_BASIC_TRIM 1  /* basic:32-72  interior stencils, rows = U[:-1, 2:-2, 2:-2]           */
#define PG_FD_SLICE_CENTRAL 2 /* analyze_results:257-274: every difference is a pair / triple of SLICES cropped to a
                               common origin, so row (t,i,j) holds u = U[t,i,j], u_x = (U[t,i,j+2]-U[t,i,j])/(2 d1),
                               u_y = (U[t,i+2,j]-U[t,i,j])/(2 d0), u_xx, u_yy from (j, j+1, j+2) / (i, i+1, i+2) and the
                               CENTRAL-time u_t = (U[t+2,i,j]-U[t,i,j])/(2 dt); rows = U[:-2, :-2, :-2] (x = a1)      */

/* candidate library (column order is the reference's) */
#define PG_LIB_KS_TRUE 0       /* p=3  lap, bih, |grad|^2                         ks2d:1095-1099 */
#define PG_LIB_KS_TRUE_ADV 1   /* p=5  + u_x, u_y                                 ks2d:1100-1102 */
#define PG_LIB_KS_RICH 2       /* p=9  1,u,u^2,u_x,u_y,lap,bih,|grad|^2,u*lap     ks2d:1048-1059 */
#define PG_LIB_KS_RICH_NOADV 3 /* p=7  rich without u_x,u_y                       ks2d:1536-1539 */
#define PG_LIB_BASIC 4         /* p=6  1,u,u_x,u_y,lap,u^2                        basic:89-99    */
#define PG_LIB_KS_GRAD 5       /* p=2  gx, gy        (pg_fd_terms only)           ks2d:70-73     */
#define PG_LIB_KS_LAP 6        /* p=1  lap           (pg_fd_terms only)           ks2d:63-67     */
#define PG_LIB_PATCH_MODEL4 7  /* p=6  1,u,u_x,u_y,lap,u^2      (pg_poly_rows)    patch:160-162  */
#define PG_LIB_PATCH_FULL 8    /* p=8  + u*u_x, u*u_y           (pg_poly_rows)    patch:163-172  */
#define PG_LIB_PATCH_DERIVS 9  /* p=6  u,u_t,u_x,u_y,u_xx,u_yy  (pg_poly_rows)    patch:240-246  */
#define PG_LIB_AR_FULL 10      /* p=13 1,u,u_x,u_y,u_xx,u_yy,lap,u^2,u*u_x,u*u_y,u^3,u_x^2,u_y^2  "Model 6" of
                                  analyze_results:618-623; Models 1-5 (:598-617) are column subsets of it, so one pass
                                  serves all six (the statistics of a subset are entries of this one)             */

/* STRidge dialect */
#define PG_STRIDGE_KS 0      /* ks2d:404-428   centre+scale X, LU solve, 1+max_iter fits      */
#define PG_STRIDGE_SKLEARN 1 /* patch:78-98    StandardScaler + Ridge(intercept), Cholesky     */
#define PG_STRIDGE_BASIC 2   /* basic:104-143  raw Gram, restart from full solve each iteration */
/* STRidge flags */
#define PG_STRIDGE_RMS_PRESCALE 1 /* ks2d:1647-1655: divide columns by sqrt(G_jj/n)+1e-12 first
                                     ('1' columns keep scale 1) and unscale the coefficients  */
#define PG_STRIDGE_NO_EPS 2       /* analyze_results:578 unscales with c / scale_ (no + 1e-12; patch:98 and ks2d:428 add it) */

/* kernel variant for pg_fd_lib_gram */
#define PG_VARIANT_AUTO 0    /* tiled TMA kernel where it applies, generic kernel for the rest */
#define PG_VARIANT_GENERIC 1 /* force the generic (reference-arithmetic) kernel everywhere    */
#define PG_VARIANT_TILED 2   /* require the tiled kernel; PG_EUNSUPPORTED if it cannot run     */

PG_API int pg_version(void);
PG_API const char *pg_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
PG_API int64_t pg_launch_count(void);
/* frees the per-device scratch of the current device */
PG_API int pg_shutdown(void);
/* number of columns of a library, or PG_EINVAL */
PG_API int pg_library_width(int library_id);

/*
 * K1 -- fused finite differences + library row + (block mean) + Gram.  Replaces, without
 * materialising Theta: gradients/laplacian (ks2d:63-73), build_dictionary(_true)
 * (ks2d:1017-1104), the forward u_t (ks2d:1511,1555), build_blockwise_dataset
 * (ks2d:358-401) and X.T@X / X.T@y (ks2d:55-58); or compute_derivatives + build_library +
 * Theta.T@Theta (basic:32-101,123-124).
 *
 *   U             [T][A0][A1]; rows come from frames 0..T-2 (frame T-1 only feeds u_t); PG_FD_SLICE_CENTRAL: frames
 *                 0..T-3 (two trailing frames), fold_of_frame then has T-2 entries
 *   bt,b0,b1      block sizes along t,a0,a1 (1,1,1 = pointwise); ragged trailing blocks
 *                 are kept with their own divisor (ks2d:384-389)
 *   fold_of_row   nullable, one fold id per block row in reference row order (t-block major,
 *                 then a0-block, then a1-block); takes precedence over fold_of_frame
 *   fold_of_frame nullable, one fold id per frame 0..T-2; a block takes the fold of its first
 *                 frame (time-holdout folds); both NULL = a single fold 0.  A NEGATIVE id (255 in
 *                 fold_of_row) excludes the row on purpose; an id >= n_folds is a caller error (counter [1])
 *   stats_out     [n_folds][PG_STATS_LEN(p)]
 *   nonfinite_out nullable, 8 int64 counters of this call:
 *                 [0] block rows skipped because a mean was not finite (ks2d:394-395).  If it is
 *                     non-zero and fold_of_row was given, the caller's row numbering no longer matches
 *                     the reference's (which renumbers after dropping): use pg_block_means +
 *                     pg_rows_gram instead.
 *                 [1] rows whose fold id was outside [0, n_folds): a caller error.  Such rows are NOT
 *                     silently dropped: when this counter is non-zero every entry of stats_out is NaN.
 *                 [2] internal (the pointwise fast path fell back to the exact non-finite handling)
 *                 [3] halo waits that timed out (pg_fd_lib_gram_halo); non-zero => stats_out is NaN
 *                 [4..7] diagnostics of the tiled blockwise kernel: SM cycle counter and %globaltimer (ns) read by
 *                     CTA 0 at its start ([4], [5]) and end ([6], [7]): ([6]-[4]) / ([7]-[5]) is the EFFECTIVE SM clock
 *                     in GHz during the launch (what NVML reports is the requested clock); 0 for other kernels
 */
PG_API int pg_fd_lib_gram(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt,
                   int fd_dialect, int library_id, int bt, int b0, int b1, const uint8_t *fold_of_row,
                   const int32_t *fold_of_frame, int n_folds, double *stats_out, int64_t *nonfinite_out,
                   int variant, void *stream);

/*
 * pg_fd_lib_gram for a time slab whose trailing frame lives on another GPU (the forward u_t of
 * ks2d:1511 is the only reader of frame T-1).  With (bt, 8m, 8n) blocks the time derivative of a block telescopes
 * to a difference of block sums of u, so the (8, 8) block MEANS of the neighbour's first frame
 * (`trailing_block_means`, [A0/8][A1/8], e.g. from pg_block_means on that frame: 1/64 of the frame's bytes) stand in
 * for the frame itself: frame T-1 of U is then a placeholder whose values are ignored.  Needs a layout the tiled
 * kernel covers completely (KS dialect, A0 % 8 == 0, A1 % 8 == 0, A1 >= 128, 16-byte aligned U), else
 * PG_EUNSUPPORTED; trailing_block_means == NULL is pg_fd_lib_gram.
 */
PG_API int pg_fd_lib_gram_tail(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt,
                   int fd_dialect, int library_id, int bt, int b0, int b1, const uint8_t *fold_of_row,
                   const int32_t *fold_of_frame, int n_folds, const double *trailing_block_means, double *stats_out,
                   int64_t *nonfinite_out, int variant, void *stream);

/*
 * pg_fd_lib_gram for a time slab whose trailing frame U[T-1] is STILL BEING WRITTEN when the call is made: a copy
 * engine is pulling it from the next rank's slab over NVLink (pg_halo_exchange).  The persistent kernel starts at
 * once and loads frame T-1 -- which only the last t-block of the slab reads -- after the 32-bit word *halo_flag
 * (device memory, written behind the transfer by a stream memory operation) has reached halo_epoch (wrap-safe
 * comparison).  One launch per slab instead of bulk + tail.  Same layout rule as pg_fd_lib_gram_tail with (bt, 8, 8)
 * blocks, else PG_EUNSUPPORTED (then make `stream` wait for the transfer and call pg_fd_lib_gram).  A wait that
 * exceeds ~4 s gives up: counter [3] is raised and the statistics are NaN.
 */
PG_API int pg_fd_lib_gram_halo(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt,
                   int fd_dialect, int library_id, int bt, int b0, int b1, const uint8_t *fold_of_row,
                   const int32_t *fold_of_frame, int n_folds, const uint32_t *halo_flag, uint32_t halo_epoch,
                   double *stats_out, int64_t *nonfinite_out, int variant, void *stream);

/*
 * The same with TWO stacks: the library columns come from U, the time derivative from Uy (same shape).  This is the
 * ks2d script's optional denoising with --denoise-space-on features (ks2d:1448-1468, 1510-1511): u_t is taken of the
 * time-smoothed stack, the features of the additionally space-smoothed one.  (bt, 8, 8) blocks of the KS dialect run
 * through the tiled kernel over U: the block mean of the forward difference telescopes to a difference of (8, 8) block
 * sums of Uy's frames k bt, which a small kernel forms first (it reads 1 / bt of Uy); everything else (pointwise rows,
 * ragged remainders, other dialects, variant = PG_VARIANT_GENERIC) uses the generic kernel (reference arithmetic).
 */
PG_API int pg_fd_lib_gram_two(const double *U, const double *Uy, int64_t T, int64_t A0, int64_t A1, double d0, double d1,
                       double dt, int fd_dialect, int library_id, int bt, int b0, int b1, const uint8_t *fold_of_row,
                       const int32_t *fold_of_frame, int n_folds, double *stats_out, int64_t *nonfinite_out, int variant,
                       void *stream);
PG_API int pg_fd_gather_rows_two(const double *U, const double *Uy, int64_t T, int64_t A0, int64_t A1, double d0, double d1,
                          double dt, int fd_dialect, int library_id, const int64_t *flat_idx, int64_t n, double *X_out,
                          double *y_out, void *stream);

/*
 * Materialised term stacks, bit-identical to the reference's NumPy arithmetic (no FMA
 * contraction, true division).  PG_FD_KS_PERIODIC: terms_out [p][T][A0][A1] over ALL T
 * frames given (ks2d:63-73, 1017-1104).  PG_FD_BASIC_TRIM with PG_LIB_BASIC: terms_out is
 * the 5 arrays (u_t, u, u_x, u_y, lap) of compute_derivatives, each [T-1][A0-4][A1-4]
 * (basic:32-72; `dt` used only here).
 */
PG_API int pg_fd_terms(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt,
                int fd_dialect, int library_id, double *terms_out, void *stream);

/*
 * K1c -- sampled pointwise rows (ks2d:1625-1636): flat_idx indexes the row space
 * ((T-1)*A0*A1 for KS, (T-1)*(A0-4)*(A1-4) for BASIC); X_out [n][p], y_out [n].
 */
PG_API int pg_fd_gather_rows(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt,
                      int fd_dialect, int library_id, const int64_t *flat_idx, int64_t n, double *X_out,
                      double *y_out, void *stream);

/*
 * Block means of k stacked arrays (the literal build_blockwise_dataset(Ut, terms, names)
 * signature, ks2d:358-401): stack [k][T][A0][A1] -> out [nrows][k], nrows =
 * ceil(T/bt)*ceil(A0/b0)*ceil(A1/b1), reference row order.  Dropping non-finite rows is
 * left to the caller.
 */
PG_API int pg_block_means(const double *stack, int k, int64_t T, int64_t A0, int64_t A1, int bt, int b0, int b1,
                   double *out, void *stream);

/*
 * Rows -> statistics for B independent problems of n rows each: X [B][n][ldx] (first p columns
 * used), y [B][n], optional fold per row [B][n].  Serves the literal stridge(X, y) signatures
 * (ks2d:404, patch:78, basic:104) and the per-patch fits (patch:420-423).
 *   shift         nullable [B][p]: statistics are taken of (X - shift), which removes the
 *                 cancellation in G - n*mu*mu^T when mean^2 >> variance (use e.g. the first row)
 *   stats_out     [B][n_folds][PG_STATS_LEN(p)]; a fold id >= n_folds is a caller error and turns the
 *                 statistics into NaN (general path; the small-problem fast path takes no folds)
 *   colminmax_out nullable [B][n_folds][2][p] = per-column min / max of the unshifted values; lets
 *                 the solver detect exactly-constant columns the way np.std()==0 does (ks2d:46-47)
 */
PG_API int pg_rows_gram(const double *X, const double *y, int64_t B, int64_t n, int p, int64_t ldx,
                 const uint8_t *fold_of_row, int n_folds, const double *shift, double *stats_out,
                 double *colminmax_out, void *stream);

/*
 * Statistics of B bootstrap resamples of ONE row set (ensemble_stridge, ks2d:603-642): weights [B][n] uint16 holds
 * how many times resample b drew row i (np.bincount of the reference's rng.choice(n, n_sub, replace=True)); a
 * resample's Gram is the multiplicity-weighted Gram, so the resampled rows are never materialised.  X [n][ldx],
 * y [n], shift [p] (nullable) are shared; stats_out [B][PG_STATS_LEN(p)]; colminmax_out [B][2][p] (nullable)
 * covers only the rows a resample drew.
 */
PG_API int pg_rows_gram_weighted(const double *X, const double *y, int64_t n, int p, int64_t ldx, const uint16_t *weights,
                          int64_t B, const double *shift, double *stats_out, double *colminmax_out, void *stream);

/*
 * K2 -- local-polynomial derivative rows (patch:193-280).  The lstsq fit of patch:231 has a
 * constant design matrix, so it is the fixed stencil W6 [6][(2rt+1)(2rs+1)^2] (rows u, u_t,
 * u_x, u_y, u_xx, u_yy; neighbour order t, y, x with x fastest).  U [T][H][W] is float32
 * (dtype 0, as the script stores it, patch:116) or float64 (dtype 1); pts [n][3] = (t,y,x).
 * X_out [n][p], y_out [n] (= u_t).
 */
PG_API int pg_poly_rows(const void *U, int dtype, int64_t T, int64_t H, int64_t W, const int32_t *pts, int64_t n,
                 const double *W6, int rt, int rs, int library_id, double *X_out, double *y_out, void *stream);

/*
 * K3 -- batched STRidge on statistics, one warp per (problem, alpha, threshold).
 *   stats        [B][PG_STATS_LEN(p)] training statistics
 *   alphas/thrs  [na] / [nt] DEVICE arrays: the sweep grid (ks2d:1720-1722)
 *   const_mask   nullable [p] uint8: columns known to be constant ('1'), forced to an exact
 *                zero coefficient as centring does in the reference
 *   signs        nullable [p] int8 (-1 / 0 / +1), PG_STRIDGE_KS only: stridge_sign_constrained
 *                (ks2d:552-600) -- a coefficient whose sign contradicts signs[j] is zeroed before the
 *                threshold mask of every iteration and after every refit
 *   colminmax    nullable [B][2][p] from pg_rows_gram: adds exact constant detection
 *   shift        nullable [B][p]: the statistics (train and held-out) are of (X - shift)
 *   eval_stats   nullable [B][PG_STATS_LEN(p)]: held-out statistics -> metrics_out
 *                [B][na][nt][2] = (r2, rmse) (ks2d:29-40) and best_out [B] = flat index
 *                a*nt+t maximising (r2, -n_active, -rmse), first maximum wins (ks2d:1731-1741)
 *   coef_out     [B][na][nt][p] coefficients in the units of the given columns
 *   relres_out   nullable [B][na][nt] (needs eval_stats): held-out ss_res / sum y^2 as the statistics give it, BEFORE
 *                the clamp at 0.  ss_res = yy - 2 c.b + c.G.c cancels when the fit is exact: below ~1e-9 the metrics
 *                are rounding noise (r2 is still 1 to ~1e-9, rmse and the sweep's tie-break are not reliable) and
 *                the caller should evaluate residuals from the rows: pg_rows_residual_ss / pg_fd_residual_ss.
 */
PG_API int pg_stridge_batched(const double *stats, int64_t B, int p, int dialect, int flags, const double *alphas,
                       int na, const double *thrs, int nt, int max_iter, const uint8_t *const_mask,
                       const int8_t *signs, const double *colminmax, const double *shift, const double *eval_stats,
                       double *coef_out, double *metrics_out, int32_t *best_out, double *relres_out, void *stream);

/*
 * scripts/patch_based_sindy.py: the regression rows of discover_pde_for_patch (sindy:300-340) for B
 * patches at once.  U [T][H][W] float64 images, origins [B][2] = (y, x) of each patch_size^2 patch; per patch the rows
 * run over frames 1..T-2 and the pixels with skip_boundary <= r, c < patch_size - skip_boundary that are multiples of
 * `subsample`, in np.where order; periodic np.roll differences INSIDE the patch (sindy:226-234), central time
 * difference (sindy:236-245).  X_out [B][n][11], y_out [B][n], n = (T-2) * n_side^2 (returned through
 * rows_per_patch_out_host, a HOST pointer, also when the output buffers are not yet known: pass B = 0).
 * scramble = 1 reproduces the reference's feature layout: it column_stacks the eleven (h, w) term arrays into (h, 11 w)
 * and views that as (h, w, 11) (sindy:269, 327-329), so feature k of pixel (r, c) is term (11 c + k) / w at column
 * (11 c + k) % w.  scramble = 0 gives the eleven terms at the pixel (what the code presumably meant).
 */
PG_API int pg_sindy_rows(const double *U, int64_t T, int64_t H, int64_t W, const int32_t *origins, int64_t B, int patch_size,
                  int skip_boundary, int subsample, double d0, double d1, double dt, int scramble, double *X_out,
                  double *y_out, int64_t *rows_per_patch_out_host, void *stream);

/*
 * build_library of basic_usage (basic:75-101) for the literal signature: four flat arrays of n values (u, u_x, u_y,
 * lap_u as compute_derivatives returned them) -> Theta_out [n][6] = [1, u, u_x, u_y, lap, u*u].
 */
PG_API int pg_basic_library_rows(const double *u, const double *u_x, const double *u_y, const double *lap_u, int64_t n,
                          double *Theta_out, void *stream);

/* dst[i] += src[i], i < n: statistics are additive over sub-slabs, passes and folds (fit_streamed, K-fold train sets) */
PG_API int pg_stats_accumulate(double *dst, const double *src, int64_t n, void *stream);

/*
 * The block-mean rows of pg_fd_lib_gram materialised: rows_out [nrows][p + 1] = (mean u_t, mean theta_0 ..) per block
 * in reference row order (t-block major), NOTHING dropped (a non-finite mean stays in its row).  Serves the rare
 * case in which rows must be renumbered after the reference drops non-finite ones (ks2d:394-395) before its random
 * row split (ks2d:1638-1641).  Generic kernel, reference arithmetic.
 */
PG_API int pg_fd_block_rows(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt,
                     int fd_dialect, int library_id, int bt, int b0, int b1, double *rows_out, void *stream);

/*
 * Exact held-out residual sums of n_coef <= 32 fitted models (r2_score / rmse, ks2d:29-40, on the test rows; the
 * sweep's arg-max key of ks2d:1731-1741 compares them): ss_out [n_coef + 1] = sum over the rows of fold `eval_fold`
 * (-1: all rows) of (y - theta . coef_j)^2, and the row count in the last entry.
 *   pg_rows_residual_ss   rows X [n][ldx], y [n] on the device (sampled pointwise rows, literal stridge(X, y) calls)
 *   pg_fd_residual_ss     straight from the field: the block rows are re-formed on the fly by the generic kernel
 *                         (same arguments as pg_fd_lib_gram), a second pass that is only needed when relres_out of
 *                         pg_stridge_batched signals cancellation (fits exact to ~1e-8: clean synthetic data)
 * coef [n_coef][p] in the units of the library columns.
 */
PG_API int pg_rows_residual_ss(const double *X, const double *y, int64_t n, int p, int64_t ldx, const uint8_t *fold_of_row,
                        int eval_fold, const double *coef, int n_coef, double *ss_out, void *stream);
PG_API int pg_fd_residual_ss(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt,
                      int fd_dialect, int library_id, int bt, int b0, int b1, const uint8_t *fold_of_row,
                      const int32_t *fold_of_frame, int n_folds, int eval_fold, const double *coef, int n_coef,
                      double *ss_out, void *stream);

/*
 * Rollout check of a discovered KS-dialect PDE (ks2d:1804-1838): explicit Euler from frame 0,
 * u_hat <- u_hat + dt * sum_k coef[k] * theta_k(u_hat) with the periodic stencils of ks2d:63-73 (terms added in
 * library order, |coef| < 1e-12 skipped), and rmse_out[k] = RMSE(U[k+1], u_hat after step k+1) (ks2d:29-32).
 *   coef      [p] DEVICE coefficients in library order        work  [2][A0][A1] DEVICE scratch frames
 *   n_steps   <= T-1                                          rmse_out [n_steps] DEVICE
 */
PG_API int pg_ks_rollout(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt,
                  int library_id, const double *coef, int n_steps, double *work, double *rmse_out, void *stream);

/*
 * Validation of the analyze_results dialect, the step after its solve.
 * pg_ar_rollout: rollout_k_rmse (analyze_results:348-395).  From every start frame t in [t0, t1 - k_steps) the model
 *   u_t = sum_k coef[k] * term(term_ids[k]) is advanced k_steps explicit-Euler steps with derivs_2d's stencils (same-grid
 *   central differences through np.pad(mode="reflect"), analyze_results:300-313; x = a1) and compared with frame
 *   t + k_steps.  term_ids index PG_LIB_AR_FULL's 13 columns; terms are added in the given order, |coef| < 1e-12
 *   skipped.  mask nullable [H][W] bytes (spatial hold-out region).  work: DEVICE scratch [2][t1-k_steps-t0][H][W].
 *   sums_out DEVICE [4] = sum e^2, sum y, sum y^2, count over the targets: rmse = sqrt([0]/[3]), nrmse = rmse / (std + 1e-12).
 * pg_one_step_ss: one_step_prediction_rmse (analyze_results:150-187): sums_out [2] = sum over t < t_max and the
 *   (masked) frame of (u[t+1] - (u[t] + dt * ut_pred[t]))^2, and the count.  u_field [t_max + 1][frame], ut_pred [t_max][frame].
 */
PG_API int pg_ar_rollout(const double *U, int64_t T, int64_t H, int64_t W, double d0, double d1, double dt,
                  const int32_t *term_ids, const double *coef, int n_terms, int k_steps, int64_t t0, int64_t t1,
                  const uint8_t *mask, double *work, double *sums_out, void *stream);
PG_API int pg_one_step_ss(const double *u_field, const double *ut_pred, int64_t t_max, int64_t frame, double dt,
                   const uint8_t *mask, double *sums_out, void *stream);

/*
 * Sums behind the reference's fit metrics (rmse / r2_score ks2d:29-40; regression_metrics patch:47-65) of
 * (y_true, y_pred), both DEVICE [n]; r = y_true - y_pred.  sums_out DEVICE [10]:
 *   [0] sum r  [1] sum r^2  [2] sum |r|  [3] sum y  [4] sum yhat
 *   [5] sum (y - mean y)^2  [6] sum (yhat - mean yhat)^2  [7] sum (y - mean y)(yhat - mean yhat)
 *   [8] sum (r - mean r)^2  [9] unused (0)
 */
PG_API int pg_fit_metrics(const double *y_true, const double *y_pred, int64_t n, double *sums_out, void *stream);
/*
 * The same sums for B problems at once, y_pred = X @ coef formed on the fly (patch:425-429: regression_metrics of every
 * patch's train / test rows): X [B][n][ldx], y [B][n], coef [B][p]; sums_out [B][10]; resid_out nullable [B][n] = y - X @ c
 * (the caller takes the median of |resid|).
 */
PG_API int pg_rows_metrics_batched(const double *X, const double *y, const double *coef, int64_t B, int64_t n, int p,
                            int64_t ldx, double *sums_out, double *resid_out, void *stream);

/*
 * Optional denoising prologue of the ks2d script (ks2d:1448-1468), the step before the hot path.
 * pg_time_moving_average: time_smooth_moving_average (ks2d:145-161), reflect-padded moving average along t
 *   through a sequential cumulative sum; bit-identical to the NumPy formulation.  window odd, out != U.
 * pg_periodic_conv: one axis of a separable circular convolution, out = sum_k weights[k] * in[.. - offsets[k] ..]
 *   (periodic); two calls with the taps of the periodic Gaussian ifft(exp(-sigma^2 k^2 / 2)) reproduce
 *   gaussian_smooth_periodic_2d (ks2d:125-142) without an FFT.  offsets / weights are DEVICE arrays.
 */
/*
 * pg_reflect_conv: one axis of scipy.ndimage.gaussian_filter (patch:335,343; analyze_results:222,250), the smoothing
 *   step before the patch / analyze_results paths: correlation with the symmetric taps weights [2 radius + 1] (DEVICE,
 *   centre at [radius]) under mode="reflect" (half-sample symmetric), accumulated in double in scipy's order (centre,
 *   then the pairs (in[l-j] + in[l+j]) * w from the outermost inwards) and rounded to the stack's dtype (0 = float32,
 *   1 = float64) once per axis: bit-identical to scipy for both dtypes.  Two calls (axis 0, then 1) filter every frame.
 */
PG_API int pg_reflect_conv(const void *in, int dtype, int64_t T, int64_t A0, int64_t A1, int axis, const double *weights,
                    int radius, void *out, void *stream);
/*
 * pg_reflect_gauss2d: both axes of the same filter (axis 0, then axis 1, the intermediate rounded to the stack's dtype
 *   as scipy's is) in one pass over the stack through shared memory: bit-identical to two pg_reflect_conv calls, one
 *   read and one write of the stack.  radius <= 32 (sigma <= 8 at scipy's truncate = 4).
 */
PG_API int pg_reflect_gauss2d(const void *in, int dtype, int64_t T, int64_t A0, int64_t A1, const double *weights, int radius,
                       void *out, void *stream);
PG_API int pg_time_moving_average(const double *U, int64_t T, int64_t A0, int64_t A1, int window, double *out, void *stream);
PG_API int pg_periodic_conv(const double *in, int64_t T, int64_t A0, int64_t A1, int axis, const int32_t *offsets,
                     const double *weights, int n_taps, double *out, void *stream);
/*
 * pg_periodic_gaussian_fft: gaussian_smooth_periodic_2d (ks2d:125-142) of every frame the reference's own way, as an
 *   FFT product: batched real-to-complex 2-D transforms (cuFFT, resolved with dlopen at first use: PG_EUNSUPPORTED
 *   where libcufft.so.11 is missing), the product with exp(-sigma^2 |k|^2 / 2) / (A0 A1) and the inverse transforms.
 *   For sigma below ~2.8 px the periodic Gaussian rings and pg_periodic_conv needs all n taps per axis; this entry
 *   point costs O(log n) per point instead.  Results agree with NumPy's FFT to ~1e-15 of the frame's magnitude.
 *   Synchronises the stream before it returns.
 */
PG_API int pg_periodic_gaussian_fft(const double *in, int64_t T, int64_t A0, int64_t A1, double sigma_px, double *out,
                             void *stream);

/*
 * Synthetic field generator for the large benchmark stacks: frames
 * t_offset..t_offset+T-1 of a smooth travelling-wave field plus counter-based noise,
 * written straight into HBM.  kind 0 = periodic (KS-shaped), 1 = laser-image-shaped [0,1].
 * Deterministic in (seed, global frame index, a0, a1), so time slabs generated on different
 * GPUs tile the same global stack.
 */
PG_API int pg_synth_field(double *U, int64_t T, int64_t A0, int64_t A1, int64_t t_offset, int64_t T_total,
                   uint64_t seed, int kind, double noise, void *stream);

/*
 * ---- Multi-GPU over NVLink peer memory (pg_comm_init, pg_halo_exchange, pg_allreduce_stats).
 * The reference has no distributed code at all; these three calls are the whole exchange of the time-slab sharding:
 * the trailing halo frame of the forward u_t (ks2d:1511) and the sum of the per-slab statistics.
 *
 * One process per GPU.  Each rank owns a zero-initialised WORKSPACE of pg_comm_workspace_bytes() bytes of device
 * memory that every peer has mapped (CUDA VMM; pde_b200.slabs uses torch symmetric memory to allocate and exchange
 * it).  peer_workspaces_host[q] is rank q's workspace AS MAPPED IN THIS PROCESS (q == rank: the local one).  Every
 * rank must issue the same sequence of pg_comm_barrier / pg_allreduce_stats calls (epochs are counted per call).
 * Device-side waits are bounded (~4 s); a wait that times out raises pg_comm_errors() and yields NaN statistics.
 */
#define PG_COMM_MAX_RANKS 16
#define PG_COMM_MAX_LEN 1536 /* doubles per all-reduce: PG_MAX_FOLDS x PG_STATS_LEN(PG_MAX_P) = 1368 */
PG_API size_t pg_comm_workspace_bytes(void);
PG_API int pg_comm_init(int rank, int world, void *const *peer_workspaces_host, void **comm_out);
PG_API int pg_comm_destroy(void *comm);
/* number of device-side waits that timed out so far (synchronises the device), or -1 */
PG_API int64_t pg_comm_errors(void *comm);
/* all ranks have reached this point of `stream` and what they wrote before it is visible to peer reads (one warp) */
PG_API int pg_comm_barrier(void *comm, void *stream);
/*
 * stats[len] <- sum over ranks, in place, in ONE launch: every rank stores its vector into every peer's workspace,
 * flags it, waits for all flags and adds the world vectors in rank order, so all ranks hold the same bits and the
 * result does not change from run to run (a ring / tree all-reduce promises neither).  len <= PG_COMM_MAX_LEN.
 */
PG_API int pg_allreduce_stats(void *comm, double *stats, int len, void *stream);
/*
 * Halo exchange without an SM: on `copy_stream`, a copy engine pulls `bytes` from peer_first_frame (the next rank's
 * frame 0, mapped peer memory) into halo_dst (this rank's U[T-1]) and a stream memory operation then publishes a new
 * epoch in the communicator's local flag word.  (*halo_flag_out, *halo_epoch_out) are what pg_fd_lib_gram_halo takes:
 * K1 can be launched on another stream immediately.  The caller orders the transfer after the peers' data is ready
 * (pg_comm_barrier on the compute stream + an event the copy stream waits for).  PG_EUNSUPPORTED when the driver has
 * no stream memory operations: then copy with cudaMemcpyAsync and make the compute stream wait for it.
 */
PG_API int pg_halo_exchange(void *comm, double *halo_dst, const double *peer_first_frame, size_t bytes, void *copy_stream,
                     const uint32_t **halo_flag_out, uint32_t *halo_epoch_out);

#ifdef __cplusplus
}
#endif
#endif /* PDEGRAM_H */
