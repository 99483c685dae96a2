// Parameter blocks and host launchers shared between api.cu and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace pg {

constexpr int GW_WARPS = 4;                 // warps per CTA in the generic kernels
constexpr int GW_THREADS = GW_WARPS * 32;

struct K1Params {
    const double *U;
    // nullable: a second stack of the same shape that the time derivative is taken of (ks2d:1448-1468 with
    // --denoise-space-on features: u_t comes from the time-smoothed stack, the library from the space-smoothed one);
    // generic kernels only
    const double *Uy;
    int64_t T, A0, A1;
    FdConsts c;
    int dialect;
    int bt, b0, b1;
    int64_t R0, R1;        // row-space extents along a0/a1 (A0,A1 for KS; A0-4,A1-4 for BASIC)
    int64_t off;           // index offset of the row space inside a frame (0 for KS, 2 for BASIC)
    int t_halo;            // trailing frames that only feed u_t: 1 (forward difference), 2 (central, PG_FD_SLICE_CENTRAL)
    int64_t nB0, nB1;      // number of blocks along a0/a1 (ceil)
    int64_t tb_lo, tb_hi, i0_lo, i0_hi, i1_lo, i1_hi;  // block-index ranges this launch covers
    const uint8_t *fold_of_row;
    const int32_t *fold_of_frame;
    int n_folds;
    double *partials;               // [parts][n_folds][S]
    unsigned long long *counters;   // [0] non-finite rows, [1] rows with an out-of-range fold id / index,
                                    // [2] set by the tiled pointwise kernel when a flushed accumulator was not finite
    const unsigned long long *run_if;  // nullable: the generic kernel only runs if *run_if != 0 (fallback launch)
    // nullable: block-mean rows [nbt][sub0][sub1][p+1] (y first) of (bt, 8, 8) sub-blocks written by the tiled kernel; the
    // generic kernel then averages the (b0/8) x (b1/8) sub-block rows of each block instead of evaluating grid points
    const double *rows8;
    int64_t sub0, sub1;
    // nullable: (8, 8) block means [A0/8][A1/8] of the frame after the last row frame; frame T-1 of U is then a
    // placeholder (tiled blockwise kernel only: pg_fd_lib_gram_tail)
    const double *tail_means;
    // nullable: frame T-1 of U is being filled by a copy engine; readable once *halo_flag >= halo_epoch
    // (tiled blockwise kernel only: pg_fd_lib_gram_halo).  counters[3] counts waits that timed out.
    const unsigned int *halo_flag;
    unsigned int halo_epoch;
    // two-stack path through the tiled blockwise kernel: block sums of Uy's frames min(k bt, T - 1) (launch_frame_block_sums8)
    const double *y_sums;
};

struct RowsParams {
    const double *X, *y;
    int64_t B, n, ldx;
    int p;
    const uint8_t *fold_of_row;  // [B][n] or null
    int n_folds;
    const double *shift;         // [B][p] or null: statistics are of (X - shift)
    const uint16_t *weights;     // [B][n] or null: row multiplicities (bootstrap resamples); X, y, shift are then SHARED by the B problems
    int chunks;                  // CTAs per problem
    double *partials;            // [B][chunks*GW_WARPS][n_folds][S]
    double *mm_partials;         // [B][chunks*GW_WARPS][n_folds][2][p] or null
    unsigned long long *counters;
};

struct StridgeParams {
    const double *stats;     // [B][S]
    int64_t B;
    int p, dialect, flags;
    const double *alphas; int na;
    const double *thrs; int nt;
    int max_iter;
    const uint8_t *const_mask;   // [p] or null
    const int8_t *signs;         // [p] or null: sign constraints (-1 / 0 / +1), ks2d:552-600
    const double *colminmax;     // [B][2][p] or null
    const double *shift;         // [B][p] or null
    const double *eval_stats;    // [B][S] or null
    double *coef_out;            // [B][na][nt][p]
    double *metrics_out;         // [B][na][nt][2] or null
    double *relres_out;          // [B][na][nt] or null: held-out ss_res / sum y^2 BEFORE clamping (cancellation indicator)
};

// What the tiled kernel covers: block indices [0,nbt) x [0,nb0) x [0,nb1) of the row space.
struct TiledPlan {
    int64_t nbt, nb0, nb1;
    int64_t n_parts;        // partial-statistics slots it writes
    size_t extra_scratch;   // bytes of scratch beyond the partials (tensor maps, work counters)
    int kernel_id;          // which specialisation
    int grid;
    int tile0, tile1, chunk_t;
    int64_t n_tiles0, n_tiles1, n_chunks;
};

int launch_k1_generic(int lib, const K1Params &P, int ctas, cudaStream_t st);
// Sum of partials[k][.] over k.  With `flag`: parts [0, n_a) are skipped if *flag != 0 and parts [n_a, n_a + n_b)
// are skipped if *flag == 0 (fast path and its conditional fallback, see tiled_pw.cu).
// `poison` (the launch's counters): the result is NaN when counters[1] (out-of-range fold ids) or counters[3] (halo
// frame never arrived) is non-zero.
int launch_reduce_partials(const double *partials, int64_t n_parts, int64_t len, double *out, int accumulate, cudaStream_t st,
                           const unsigned long long *flag = nullptr, int64_t n_a = 0, int64_t n_b = 0,
                           const unsigned long long *poison = nullptr);
// held-out residual sums (second pass): partials [ctas][J + 1] (sum r_j^2 ..., row count)
int launch_k1_generic_resid(int lib, const K1Params &P, const double *coef, int J, int eval_fold, double *partials, int ctas,
                            cudaStream_t st);
int launch_sindy_rows(const double *U, int64_t T, int64_t H, int64_t W, const int32_t *origins, int64_t B, int ps, int skip, int sub,
                      const FdConsts &k, int scramble, int n_side, int first, double *X, double *y, cudaStream_t st);
int launch_basic_library_rows(const double *u, const double *ux, const double *uy, const double *lap, int64_t n, double *Theta, cudaStream_t st);
int launch_stats_accumulate(double *dst, const double *src, int64_t n, cudaStream_t st);
int launch_k1_generic_rows(int lib, const K1Params &P, double *rows_out, int ctas, cudaStream_t st);
int launch_rows_resid(const double *X, const double *y, int64_t n, int p, int64_t ldx, const uint8_t *fold_of_row, int eval_fold,
                      const double *coef, int J, double *partials, int ctas, cudaStream_t st);
int launch_fd_terms(int dialect, int lib, const double *U, int64_t T, int64_t A0, int64_t A1, const FdConsts &c, double *out, cudaStream_t st);
int launch_fd_gather(int lib, const K1Params &P, const int64_t *flat_idx, int64_t n, double *X, double *y, cudaStream_t st);
int launch_block_means(const double *stack, int k, int64_t T, int64_t A0, int64_t A1, int bt, int b0, int b1, double *out, cudaStream_t st);
int launch_rows_gram(const RowsParams &P, double *stats, double *colminmax, cudaStream_t st);
int launch_stridge(const StridgeParams &P, int32_t *best_out, cudaStream_t st);
int launch_poly_rows(const void *U, int dtype, int64_t T, int64_t H, int64_t W, const int32_t *pts, int64_t n, const double *W6, int rt, int rs, int mode, double *X, double *y, unsigned long long *counters, cudaStream_t st);
int launch_synth(double *U, int64_t T, int64_t A0, int64_t A1, int64_t t_offset, int64_t T_total, uint64_t seed, int kind, double noise, cudaStream_t st);

// smooth.cu
int launch_time_moving_average(const double *U, int64_t T, int64_t A0, int64_t A1, int window, double *out, cudaStream_t st);
int launch_periodic_conv(const double *in, int64_t T, int64_t A0, int64_t A1, int axis, const int32_t *off, const double *w,
                         int n_taps, double *out, cudaStream_t st);

// periodic Gaussian as an FFT product (cuFFT through dlopen): hx [A0] (includes 1 / (A0 A1)), hy [A1 / 2 + 1] on the HOST
void fft_plans_release(int dev);   // pg_shutdown: destroys the cached cuFFT plans of a device
size_t periodic_gaussian_fft_scratch(int64_t T, int64_t A0, int64_t A1, int64_t *batch_out);
int launch_periodic_gaussian_fft(const double *in, int64_t T, int64_t A0, int64_t A1, const double *hx_host, const double *hy_host,
                                 double *out, void *scratch, int64_t batch, cudaStream_t st);

int launch_reflect_gauss2d(const void *in, int dtype, int64_t T, int64_t A0, int64_t A1, const double *w, int radius, void *out,
                           cudaStream_t st);
int launch_reflect_conv(const void *in, int dtype, int64_t T, int64_t A0, int64_t A1, int axis, const double *w, int radius,
                        void *out, cudaStream_t st);

// rollout.cu
int launch_ar_rollout(const double *U, int64_t H, int64_t W, const FdConsts &c, const int32_t *term_ids, const double *coef,
                      int n_terms, int k_steps, int64_t t0, int64_t n_start, const uint8_t *mask, double *work, double *partials,
                      int blocks, double *out4, cudaStream_t st);
int launch_one_step(const double *u, const double *ut, int64_t t_max, int64_t frame, double dt, const uint8_t *mask,
                    double *partials, int blocks, double *out2, cudaStream_t st);
int launch_rows_metrics_batched(const double *X, const double *y, const double *coef, int64_t B, int64_t n, int p, int64_t ldx,
                                double *sums_out, double *resid_out, cudaStream_t st);
int launch_fit_metrics(const double *y, const double *yh, int64_t n, double *partials, int blocks, double *out10, cudaStream_t st);
int rollout_blocks(int64_t A0, int64_t A1, int n_sm);
int launch_rollout(int lib, const double *U, int64_t A0, int64_t A1, const FdConsts &c, const double *coef, int n_steps,
                   double *work, double *partials, int blocks, double *rmse, cudaStream_t st);

// tiled_pw.cu (pointwise rows)
bool tiled_pw_plan(const K1Params &P, int lib, int n_sm, TiledPlan &plan);
int tiled_pw_launch(const K1Params &P, int lib, const TiledPlan &plan, double *partials, cudaStream_t st);
// boundary terms the tiled pointwise kernel leaves out of its accumulators (basic_usage library): scratch bytes, and the
// launches that add them to the reduced statistics (skipped on the device when *skip_if != 0)
size_t tiled_pw_external_scratch(const K1Params &P, int lib);
int tiled_pw_external(const K1Params &P, int lib, double *per_frame, const unsigned long long *skip_if, double *stats_out,
                      cudaStream_t st);

// tiled.cu ((bt,8,8) block means)
bool tiled_plan(const K1Params &P, int lib, int64_t nBt, int n_sm, TiledPlan &plan);
// rows8 != null: write the block-mean rows [nbt][A0/8][A1/8][p+1] instead of accumulating statistics
int tiled_launch(const K1Params &P, int lib, const TiledPlan &plan, double *partials, char *extra, cudaStream_t st,
                 double *rows8 = nullptr);
// out [nbt + 1][ceil(A0 / 8)][A1 / 8]: (8, 8) block sums of frames min(k bt, T - 1) of Uy
int launch_frame_block_sums8(const double *Uy, int64_t T, int64_t A0, int64_t A1, int bt, int64_t nbt, double *out, cudaStream_t st);

}  // namespace pg
