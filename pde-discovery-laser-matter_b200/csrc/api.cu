// extern "C" entry points of libpdegram.so (include/pdegram.h): argument validation, scratch
// management and kernel dispatch.  No exception crosses this boundary.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <map>
#include <vector>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "launch.h"

namespace pg {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// One scratch buffer per (device, stream): calls on one stream are ordered, so reuse is safe.
struct Scratch {
    void *ptr = nullptr;
    size_t bytes = 0;
};
static std::mutex g_mu;
static std::map<std::pair<int, void *>, Scratch> g_scratch;

static int scratch_for(cudaStream_t st, size_t bytes, void **ptr) {
    int dev = 0;
    PG_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    Scratch &s = g_scratch[{dev, (void *)st}];
    if (s.bytes < bytes) {
        if (s.ptr) {
            // earlier work on this stream may still read the old buffer
            PG_CUDA(cudaStreamSynchronize(st));
            PG_CUDA(cudaFree(s.ptr));
            s.ptr = nullptr;
            s.bytes = 0;
        }
        size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
        if (cudaMalloc(&s.ptr, want) != cudaSuccess) {
            cudaGetLastError();
            s.ptr = nullptr;
            PG_FAIL(PG_ENOMEM, "cannot allocate %zu bytes of device scratch", want);
        }
        s.bytes = want;
    }
    *ptr = s.ptr;
    return PG_OK;
}

static bool env_flag(const char *name) {
    const char *v = getenv(name);
    return v && *v && *v != '0';
}

static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

static bool ks_lib(int lib) {
    return lib == PG_LIB_KS_TRUE || lib == PG_LIB_KS_TRUE_ADV || lib == PG_LIB_KS_RICH || lib == PG_LIB_KS_RICH_NOADV;
}

// Validate (dialect, library, shape) and fill the row-space description.
static int describe(K1Params &P, const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt,
                    int dialect, int lib, bool fused) {
    if (!U) PG_FAIL(PG_EINVAL, "U is null");
    if (T < 1 || A0 < 1 || A1 < 1) PG_FAIL(PG_EINVAL, "bad shape T=%lld A0=%lld A1=%lld", (long long)T, (long long)A0, (long long)A1);
    if (!(d0 > 0) || !(d1 > 0) || !(dt > 0)) PG_FAIL(PG_EINVAL, "grid spacings must be positive");
    P.U = U; P.T = T; P.A0 = A0; P.A1 = A1;
    P.c = make_consts(d0, d1, dt);
    P.dialect = dialect;
    P.t_halo = 1;
    if (dialect == PG_FD_SLICE_CENTRAL) {
        if (lib != PG_LIB_AR_FULL) PG_FAIL(PG_EINVAL, "the analyze_results dialect has library PG_LIB_AR_FULL (its Models 1-5 are column subsets)");
        if (A0 < 3 || A1 < 3) PG_FAIL(PG_EINVAL, "the analyze_results dialect needs A0, A1 >= 3 (rows are U[:-2, :-2, :-2])");
        P.R0 = A0 - 2; P.R1 = A1 - 2; P.off = 0; P.t_halo = 2;
    } else if (dialect == PG_FD_KS_PERIODIC) {
        if (fused && !ks_lib(lib)) PG_FAIL(PG_EINVAL, "library %d does not belong to the KS dialect", lib);
        P.R0 = A0; P.R1 = A1; P.off = 0;
    } else if (dialect == PG_FD_BASIC_TRIM) {
        if (lib != PG_LIB_BASIC) PG_FAIL(PG_EINVAL, "the basic_usage dialect only has library PG_LIB_BASIC");
        if (A0 < 5 || A1 < 5) PG_FAIL(PG_EINVAL, "basic_usage dialect needs A0, A1 >= 5 (interior [2:-2])");
        P.R0 = A0 - 4; P.R1 = A1 - 4; P.off = 2;
    } else {
        PG_FAIL(PG_EINVAL, "unknown finite-difference dialect %d", dialect);
    }
    return PG_OK;
}

// KS dialect, (bt, 8m, 8n) blocks on a grid of whole (8, 8) sub-blocks: the tiled kernel writes the block-mean rows of
// the (bt, 8, 8) sub-blocks (EMIT), the generic kernel averages the sub-block rows of each block (equal sizes, so the
// mean of means is the block mean; ragged edge blocks hold fewer sub-blocks) and accumulates the statistics with the
// caller's folds.  The rows are (p+1)/(64 bt) of the field's bytes.  Returns 1 when the layout does not qualify.
static int two_stage_blocks(const K1Params &P, int lib, int64_t nBt, int64_t len, double *stats_out, int64_t *nonfinite_out,
                            cudaStream_t st) {
    if (P.dialect != PG_FD_KS_PERIODIC || P.b0 % 8 || P.b1 % 8 || P.A0 % 8 || P.A1 % 8) return 1;
    // (bt, 8, 8) itself is fused in one kernel, except with more than two per-row folds (its masked accumulators hold two)
    if (P.b0 == 8 && P.b1 == 8 && !(P.fold_of_row && (P.n_folds > 2 || env_flag("PG_ROWFOLD_TWO_STAGE")))) return 1;
    K1Params Q = P;
    Q.b0 = Q.b1 = 8; Q.nB0 = P.A0 / 8; Q.nB1 = P.A1 / 8;
    Q.fold_of_row = nullptr; Q.fold_of_frame = nullptr; Q.n_folds = 1;
    TiledPlan plan{};
    if (!tiled_plan(Q, lib, nBt, sm_count(), plan)) return 1;
    if (plan.nbt != nBt || plan.nb0 != Q.nB0 || plan.nb1 != Q.nB1) return 1;   // the tiles must cover the grid
    const int p = library_width(lib), S = PG_STATS_LEN(p);
    const int64_t items = nBt * P.nB0 * P.nB1;
    int64_t g = (items + GW_THREADS - 1) / GW_THREADS;
    const int ctas = (int)(g < 1 ? 1 : (g > sm_count() * 8 ? sm_count() * 8 : g));
    const int64_t gen_parts = (int64_t)ctas * GW_WARPS;
    const size_t b_parts = sizeof(double) * (size_t)(gen_parts * len + plan.n_parts * S);
    const size_t b_extra = (plan.extra_scratch + 15) / 16 * 16;
    const size_t b_rows = sizeof(double) * (size_t)(nBt * Q.nB0 * Q.nB1 * (p + 1));
    void *scr = nullptr;
    int rc = scratch_for(st, 64 + b_parts + b_extra + b_rows, &scr);
    if (rc) return rc;
    unsigned long long *counters = (unsigned long long *)scr;
    double *partials = (double *)((char *)scr + 64);
    char *extra = (char *)scr + 64 + b_parts;
    double *rows8 = (double *)(extra + b_extra);
    PG_CUDA(cudaMemsetAsync(counters, 0, 64, st));
    Q.counters = counters;
    rc = tiled_launch(Q, lib, plan, partials + gen_parts * len, extra, st, rows8);
    if (rc) return rc;
    K1Params G = P;
    G.counters = counters;
    G.tb_lo = 0; G.tb_hi = nBt; G.i0_lo = 0; G.i0_hi = P.nB0; G.i1_lo = 0; G.i1_hi = P.nB1;
    G.partials = partials; G.run_if = nullptr;
    G.rows8 = rows8; G.sub0 = Q.nB0; G.sub1 = Q.nB1;
    rc = launch_k1_generic(lib, G, ctas, st);
    if (rc) return rc;
    rc = launch_reduce_partials(partials, gen_parts, len, stats_out, 0, st, nullptr, 0, 0, counters);
    if (rc) return rc;
    if (nonfinite_out) PG_CUDA(cudaMemcpyAsync(nonfinite_out, counters, 8 * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    return PG_OK;
}

}  // namespace pg

using namespace pg;

extern "C" {

int pg_version(void) { return PG_VERSION; }
int64_t pg_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }
const char *pg_last_error(void) { return g_err; }
int pg_library_width(int library_id) {
    const int w = library_width(library_id);
    if (w < 0) PG_FAIL(PG_EINVAL, "unknown library %d", library_id);
    return w;
}

int pg_shutdown(void) {
    int dev = 0;
    PG_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto it = g_scratch.begin(); it != g_scratch.end();) {
        if (it->first.first == dev) {
            if (it->second.ptr) cudaFree(it->second.ptr);
            it = g_scratch.erase(it);
        } else {
            ++it;
        }
    }
    fft_plans_release(dev);
    return PG_OK;
}

static int fd_lib_gram_impl(const double *U, const double *Uy, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt, int fd_dialect,
                            int library_id, int bt, int b0, int b1, const uint8_t *fold_of_row, const int32_t *fold_of_frame,
                            int n_folds, const double *trailing_block_means, const uint32_t *halo_flag, uint32_t halo_epoch,
                            double *stats_out, int64_t *nonfinite_out, int variant, void *stream);

int pg_fd_lib_gram(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt, int fd_dialect,
                   int library_id, int bt, int b0, int b1, const uint8_t *fold_of_row, const int32_t *fold_of_frame,
                   int n_folds, double *stats_out, int64_t *nonfinite_out, int variant, void *stream) {
    return fd_lib_gram_impl(U, nullptr, T, A0, A1, d0, d1, dt, fd_dialect, library_id, bt, b0, b1, fold_of_row, fold_of_frame, n_folds,
                            nullptr, nullptr, 0, stats_out, nonfinite_out, variant, stream);
}

int pg_fd_lib_gram_tail(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt, int fd_dialect,
                        int library_id, int bt, int b0, int b1, const uint8_t *fold_of_row, const int32_t *fold_of_frame,
                        int n_folds, const double *trailing_block_means, double *stats_out, int64_t *nonfinite_out,
                        int variant, void *stream) {
    return fd_lib_gram_impl(U, nullptr, T, A0, A1, d0, d1, dt, fd_dialect, library_id, bt, b0, b1, fold_of_row, fold_of_frame, n_folds,
                            trailing_block_means, nullptr, 0, stats_out, nonfinite_out, variant, stream);
}

int pg_fd_lib_gram_halo(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt, int fd_dialect,
                        int library_id, int bt, int b0, int b1, const uint8_t *fold_of_row, const int32_t *fold_of_frame,
                        int n_folds, const uint32_t *halo_flag, uint32_t halo_epoch, double *stats_out,
                        int64_t *nonfinite_out, int variant, void *stream) {
    if (!halo_flag) PG_FAIL(PG_EINVAL, "halo_flag is null (use pg_fd_lib_gram)");
    return fd_lib_gram_impl(U, nullptr, T, A0, A1, d0, d1, dt, fd_dialect, library_id, bt, b0, b1, fold_of_row, fold_of_frame, n_folds,
                            nullptr, halo_flag, halo_epoch, stats_out, nonfinite_out, variant, stream);
}

int pg_fd_lib_gram_two(const double *U, const double *Uy, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt,
                       int fd_dialect, int library_id, int bt, int b0, int b1, const uint8_t *fold_of_row,
                       const int32_t *fold_of_frame, int n_folds, double *stats_out, int64_t *nonfinite_out, int variant,
                       void *stream) {
    if (!Uy) PG_FAIL(PG_EINVAL, "Uy is null (use pg_fd_lib_gram)");
    if (reinterpret_cast<uintptr_t>(Uy) & 15) variant = PG_VARIANT_GENERIC;     // the block-sum kernel reads 16-byte cells
    return fd_lib_gram_impl(U, Uy, T, A0, A1, d0, d1, dt, fd_dialect, library_id, bt, b0, b1, fold_of_row, fold_of_frame, n_folds,
                            nullptr, nullptr, 0, stats_out, nonfinite_out, variant, stream);
}

static int fd_lib_gram_impl(const double *U, const double *Uy, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt, int fd_dialect,
                            int library_id, int bt, int b0, int b1, const uint8_t *fold_of_row, const int32_t *fold_of_frame,
                            int n_folds, const double *trailing_block_means, const uint32_t *halo_flag, uint32_t halo_epoch,
                            double *stats_out, int64_t *nonfinite_out, int variant, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    K1Params P{};
    int rc = describe(P, U, T, A0, A1, d0, d1, dt, fd_dialect, library_id, true);
    if (rc) return rc;
    P.Uy = Uy;
    if (bt <= 0 || b0 <= 0 || b1 <= 0) PG_FAIL(PG_EINVAL, "block sizes must be > 0");  // ks2d:378-379
    if (n_folds < 1 || n_folds > PG_MAX_FOLDS) PG_FAIL(PG_EINVAL, "n_folds must be in 1..%d", PG_MAX_FOLDS);
    if (!stats_out) PG_FAIL(PG_EINVAL, "stats_out is null");
    if (variant < PG_VARIANT_AUTO || variant > PG_VARIANT_TILED) PG_FAIL(PG_EINVAL, "unknown variant %d", variant);
    const int p = library_width(library_id), S = PG_STATS_LEN(p);
    const int64_t Trows = T - P.t_halo;
    P.bt = bt; P.b0 = b0; P.b1 = b1;
    P.fold_of_row = fold_of_row; P.fold_of_frame = fold_of_frame; P.n_folds = n_folds;
    P.tail_means = trailing_block_means;
    P.halo_flag = halo_flag; P.halo_epoch = halo_epoch;
    if ((trailing_block_means || halo_flag) && variant == PG_VARIANT_GENERIC)
        PG_FAIL(PG_EUNSUPPORTED, "trailing_block_means / halo_flag are served by the tiled blockwise kernel only");
    const int64_t len = (int64_t)n_folds * S;
    if (Trows <= 0) {  // a single frame has no u_t: zero rows
        PG_CUDA(cudaMemsetAsync(stats_out, 0, sizeof(double) * len, st));
        if (nonfinite_out) PG_CUDA(cudaMemsetAsync(nonfinite_out, 0, 8 * sizeof(int64_t), st));
        return PG_OK;
    }
    const int64_t nBt = (Trows + bt - 1) / bt;
    P.nB0 = (P.R0 + b0 - 1) / b0; P.nB1 = (P.R1 + b1 - 1) / b1;

    if (variant != PG_VARIANT_GENERIC && !halo_flag && !Uy) {
        rc = two_stage_blocks(P, library_id, nBt, len, stats_out, nonfinite_out, st);
        if (rc <= 0) return rc;
    }
    // Tiled TMA kernels over the region they support; the generic kernel covers what is left.
    TiledPlan plan{};
    bool tiled = false, pointwise = false;
    if (variant != PG_VARIANT_GENERIC) {
        pointwise = bt == 1 && b0 == 1 && b1 == 1;
        // two stacks: the blockwise kernel takes the time derivative from block sums of Uy; pointwise rows would need
        // both stacks in shared memory and stay with the generic kernel
        tiled = pointwise ? (!Uy && tiled_pw_plan(P, library_id, sm_count(), plan)) : tiled_plan(P, library_id, nBt, sm_count(), plan);
        if (!tiled && variant == PG_VARIANT_TILED)
            PG_FAIL(PG_EUNSUPPORTED, "no tiled kernel for dialect %d library %d block (%d,%d,%d) shape (%lld,%lld,%lld)",
                    fd_dialect, library_id, bt, b0, b1, (long long)T, (long long)A0, (long long)A1);
    }
    if ((trailing_block_means || halo_flag) && !(tiled && !pointwise && plan.nbt == nBt && plan.nb0 == P.nB0 && plan.nb1 == P.nB1))
        PG_FAIL(PG_EUNSUPPORTED, "trailing_block_means / halo_flag need a layout the tiled blockwise kernel covers completely "
                "(KS dialect, (bt, 8, 8) blocks, A0 %% 8 == 0, A1 %% 8 == 0, A1 >= 128); wait for the frame on the stream instead");
    // generic launches: up to 4 boxes of block indices (the whole space when not tiled).  Box 0 of the
    // pointwise path is its conditional fallback: the tiled region again, run only if that kernel saw a
    // non-finite accumulator (the reference drops such rows; the generic kernel does it exactly).
    struct Box { int64_t t0, t1, a0, a1, c0, c1; };
    Box boxes[4];
    int nbox = 0;
    const bool fallback = tiled && pointwise;
    if (!tiled) {
        boxes[nbox++] = {0, nBt, 0, P.nB0, 0, P.nB1};
    } else {
        // plan covers block indices [0,plan.nbt) x [0,plan.nb0) x [0,plan.nb1)
        if (fallback) boxes[nbox++] = {0, plan.nbt, 0, plan.nb0, 0, plan.nb1};
        if (plan.nbt < nBt) boxes[nbox++] = {plan.nbt, nBt, 0, P.nB0, 0, P.nB1};
        if (plan.nb0 < P.nB0) boxes[nbox++] = {0, plan.nbt, plan.nb0, P.nB0, 0, P.nB1};
        if (plan.nb1 < P.nB1) boxes[nbox++] = {0, plan.nbt, 0, plan.nb0, plan.nb1, P.nB1};
    }
    const int max_ctas = sm_count() * 8;
    int box_ctas[4] = {0, 0, 0, 0};
    int64_t gen_parts = 0;
    for (int k = 0; k < nbox; ++k) {
        const int64_t items = (boxes[k].t1 - boxes[k].t0) * (boxes[k].a1 - boxes[k].a0) * (boxes[k].c1 - boxes[k].c0);
        int64_t g = (items + GW_THREADS - 1) / GW_THREADS;
        box_ctas[k] = (int)(g < 1 ? 1 : (g > max_ctas ? max_ctas : g));
        gen_parts += (int64_t)box_ctas[k] * GW_WARPS;
    }
    const int64_t tiled_parts = tiled ? plan.n_parts : 0;
    const size_t b_extra = tiled ? (plan.extra_scratch + 15) / 16 * 16 : 0;
    const size_t b_ysums = (tiled && Uy) ? sizeof(double) * (size_t)((nBt + 1) * ((A0 + 7) / 8) * (A1 / 8)) : 0;
    const size_t b_pwext = (tiled && pointwise) ? (tiled_pw_external_scratch(P, library_id) + 15) / 16 * 16 : 0;
    const size_t bytes = 64 + sizeof(double) * (size_t)((gen_parts + tiled_parts) * len) + b_extra + b_ysums + b_pwext;
    void *scr = nullptr;
    rc = scratch_for(st, bytes, &scr);
    if (rc) return rc;
    unsigned long long *counters = (unsigned long long *)scr;
    double *partials = (double *)((char *)scr + 64);
    PG_CUDA(cudaMemsetAsync(counters, 0, 64, st));
    P.counters = counters;
    int64_t part_off = 0;
    if (tiled && Uy) {
        double *ys = (double *)((char *)(partials + (gen_parts + tiled_parts) * len) + b_extra);
        rc = launch_frame_block_sums8(Uy, T, A0, A1, bt, nBt, ys, st);
        if (rc) return rc;
        P.y_sums = ys;
    }
    if (tiled) {
        rc = pointwise ? tiled_pw_launch(P, library_id, plan, partials, st)
                       : tiled_launch(P, library_id, plan, partials, (char *)(partials + (gen_parts + tiled_parts) * len), st);
        if (rc) return rc;
        part_off = tiled_parts;
    }
    for (int k = 0; k < nbox; ++k) {
        K1Params Q = P;
        Q.tb_lo = boxes[k].t0; Q.tb_hi = boxes[k].t1; Q.i0_lo = boxes[k].a0; Q.i0_hi = boxes[k].a1;
        Q.i1_lo = boxes[k].c0; Q.i1_hi = boxes[k].c1;
        Q.partials = partials + part_off * len;
        Q.run_if = (fallback && k == 0) ? counters + 2 : nullptr;
        rc = launch_k1_generic(library_id, Q, box_ctas[k], st);
        if (rc) return rc;
        part_off += (int64_t)box_ctas[k] * GW_WARPS;
    }
    if (fallback) {
        rc = launch_reduce_partials(partials, part_off, len, stats_out, 0, st, counters + 2, tiled_parts,
                                    (int64_t)box_ctas[0] * GW_WARPS, counters);
        if (rc) return rc;
        if (b_pwext)
            rc = tiled_pw_external(P, library_id, (double *)((char *)(partials + (gen_parts + tiled_parts) * len) + b_extra + b_ysums),
                                   counters + 2, stats_out, st);
    } else
        rc = launch_reduce_partials(partials, part_off, len, stats_out, 0, st, nullptr, 0, 0, counters);
    if (rc) return rc;
    if (nonfinite_out)
        PG_CUDA(cudaMemcpyAsync(nonfinite_out, counters, 8 * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    return PG_OK;
}

int pg_fd_residual_ss(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt, int fd_dialect,
                      int library_id, int bt, int b0, int b1, const uint8_t *fold_of_row, const int32_t *fold_of_frame,
                      int n_folds, int eval_fold, const double *coef, int n_coef, double *ss_out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    K1Params P{};
    int rc = describe(P, U, T, A0, A1, d0, d1, dt, fd_dialect, library_id, true);
    if (rc) return rc;
    if (bt <= 0 || b0 <= 0 || b1 <= 0) PG_FAIL(PG_EINVAL, "block sizes must be > 0");
    if (n_folds < 1 || n_folds > PG_MAX_FOLDS) PG_FAIL(PG_EINVAL, "n_folds must be in 1..%d", PG_MAX_FOLDS);
    if (eval_fold < -1 || eval_fold >= n_folds) PG_FAIL(PG_EINVAL, "eval_fold must be -1 (all rows) or a fold id");
    if (n_coef < 1 || n_coef > 32) PG_FAIL(PG_EINVAL, "n_coef must be in 1..32 (call again for more)");
    if (!coef || !ss_out) PG_FAIL(PG_EINVAL, "null buffer");
    const int64_t Trows = T - P.t_halo;
    if (Trows <= 0) {
        PG_CUDA(cudaMemsetAsync(ss_out, 0, sizeof(double) * (n_coef + 1), st));
        return PG_OK;
    }
    P.bt = bt; P.b0 = b0; P.b1 = b1;
    P.fold_of_row = fold_of_row; P.fold_of_frame = fold_of_frame; P.n_folds = n_folds;
    const int64_t nBt = (Trows + bt - 1) / bt;
    P.nB0 = (P.R0 + b0 - 1) / b0; P.nB1 = (P.R1 + b1 - 1) / b1;
    P.tb_lo = 0; P.tb_hi = nBt; P.i0_lo = 0; P.i0_hi = P.nB0; P.i1_lo = 0; P.i1_hi = P.nB1;
    const int64_t items = nBt * P.nB0 * P.nB1;
    int64_t g = (items + GW_THREADS - 1) / GW_THREADS;
    const int ctas = (int)(g < 1 ? 1 : (g > sm_count() * 8 ? sm_count() * 8 : g));
    void *scr = nullptr;
    rc = scratch_for(st, 64 + sizeof(double) * (size_t)ctas * (n_coef + 1), &scr);
    if (rc) return rc;
    PG_CUDA(cudaMemsetAsync(scr, 0, 64, st));
    P.counters = (unsigned long long *)scr;
    double *partials = (double *)((char *)scr + 64);
    rc = launch_k1_generic_resid(library_id, P, coef, n_coef, eval_fold, partials, ctas, st);
    if (rc) return rc;
    return launch_reduce_partials(partials, ctas, n_coef + 1, ss_out, 0, st, nullptr, 0, 0, P.counters);
}

int pg_sindy_rows(const double *U, int64_t T, int64_t H, int64_t W, const int32_t *origins, int64_t B, int patch_size,
                  int skip_boundary, int subsample, double d0, double d1, double dt, int scramble, double *X_out, double *y_out,
                  int64_t *rows_per_patch_out_host, void *stream) {
    if (!U || !origins || !X_out || !y_out) PG_FAIL(PG_EINVAL, "null buffer");
    if (T < 3 || patch_size < 3 || patch_size > H || patch_size > W) PG_FAIL(PG_EINVAL, "need T >= 3 and 3 <= patch_size <= H, W");
    if (skip_boundary < 0 || subsample < 1 || B < 0) PG_FAIL(PG_EINVAL, "bad skip_boundary / subsample / B");
    if (!(d0 > 0) || !(d1 > 0) || !(dt > 0)) PG_FAIL(PG_EINVAL, "grid spacings must be positive");
    // mask (sindy:308-320): rows / columns skip .. ps-skip-1 that are multiples of `subsample`
    const int first = ((skip_boundary + subsample - 1) / subsample) * subsample;
    const int last = patch_size - skip_boundary - 1;
    const int n_side = last >= first ? (last - first) / subsample + 1 : 0;
    if (rows_per_patch_out_host) *rows_per_patch_out_host = (int64_t)(T - 2) * n_side * n_side;
    if (B == 0 || n_side == 0) return PG_OK;
    return launch_sindy_rows(U, T, H, W, origins, B, patch_size, skip_boundary, subsample, make_consts(d0, d1, dt), scramble ? 1 : 0,
                             n_side, first, X_out, y_out, (cudaStream_t)stream);
}

int pg_basic_library_rows(const double *u, const double *u_x, const double *u_y, const double *lap_u, int64_t n,
                          double *Theta_out, void *stream) {
    if (n < 0) PG_FAIL(PG_EINVAL, "n < 0");
    if (n == 0) return PG_OK;
    if (!u || !u_x || !u_y || !lap_u || !Theta_out) PG_FAIL(PG_EINVAL, "null buffer");
    return launch_basic_library_rows(u, u_x, u_y, lap_u, n, Theta_out, (cudaStream_t)stream);
}

int pg_stats_accumulate(double *dst, const double *src, int64_t n, void *stream) {
    if (n < 0) PG_FAIL(PG_EINVAL, "n < 0");
    if (n == 0) return PG_OK;
    if (!dst || !src) PG_FAIL(PG_EINVAL, "null buffer");
    return launch_stats_accumulate(dst, src, n, (cudaStream_t)stream);
}

int pg_fd_block_rows(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt, int fd_dialect,
                     int library_id, int bt, int b0, int b1, double *rows_out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    K1Params P{};
    int rc = describe(P, U, T, A0, A1, d0, d1, dt, fd_dialect, library_id, true);
    if (rc) return rc;
    if (bt <= 0 || b0 <= 0 || b1 <= 0) PG_FAIL(PG_EINVAL, "block sizes must be > 0");
    if (!rows_out) PG_FAIL(PG_EINVAL, "rows_out is null");
    const int64_t Trows = T - P.t_halo;
    if (Trows <= 0) return PG_OK;
    P.bt = bt; P.b0 = b0; P.b1 = b1; P.n_folds = 1;
    const int64_t nBt = (Trows + bt - 1) / bt;
    P.nB0 = (P.R0 + b0 - 1) / b0; P.nB1 = (P.R1 + b1 - 1) / b1;
    P.tb_lo = 0; P.tb_hi = nBt; P.i0_lo = 0; P.i0_hi = P.nB0; P.i1_lo = 0; P.i1_hi = P.nB1;
    const int64_t items = nBt * P.nB0 * P.nB1;
    int64_t g = (items + GW_THREADS - 1) / GW_THREADS;
    const int ctas = (int)(g < 1 ? 1 : (g > sm_count() * 8 ? sm_count() * 8 : g));
    return launch_k1_generic_rows(library_id, P, rows_out, ctas, st);
}

int pg_rows_residual_ss(const double *X, const double *y, int64_t n, int p, int64_t ldx, const uint8_t *fold_of_row,
                        int eval_fold, const double *coef, int n_coef, double *ss_out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (p < 1 || p > PG_MAX_P) PG_FAIL(PG_EINVAL, "p must be in 1..%d", PG_MAX_P);
    if (n < 0 || ldx < p) PG_FAIL(PG_EINVAL, "bad shape n=%lld ldx=%lld", (long long)n, (long long)ldx);
    if (n_coef < 1 || n_coef > 32) PG_FAIL(PG_EINVAL, "n_coef must be in 1..32 (call again for more)");
    if (!coef || !ss_out || (n > 0 && (!X || !y))) PG_FAIL(PG_EINVAL, "null buffer");
    if (n == 0) {
        PG_CUDA(cudaMemsetAsync(ss_out, 0, sizeof(double) * (n_coef + 1), st));
        return PG_OK;
    }
    int64_t g = (n + GW_THREADS * 8 - 1) / (GW_THREADS * 8);
    const int ctas = (int)(g < 1 ? 1 : (g > sm_count() * 8 ? sm_count() * 8 : g));
    void *scr = nullptr;
    int rc = scratch_for(st, sizeof(double) * (size_t)ctas * (n_coef + 1), &scr);
    if (rc) return rc;
    rc = launch_rows_resid(X, y, n, p, ldx, fold_of_row, eval_fold, coef, n_coef, (double *)scr, ctas, st);
    if (rc) return rc;
    return launch_reduce_partials((double *)scr, ctas, n_coef + 1, ss_out, 0, st);
}

int pg_fd_terms(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt, int fd_dialect,
                int library_id, double *terms_out, void *stream) {
    K1Params P{};
    int rc = describe(P, U, T, A0, A1, d0, d1, dt, fd_dialect, library_id, false);
    if (rc) return rc;
    if (!terms_out) PG_FAIL(PG_EINVAL, "terms_out is null");
    if (fd_dialect == PG_FD_KS_PERIODIC && !ks_lib(library_id) && library_id != PG_LIB_KS_GRAD && library_id != PG_LIB_KS_LAP)
        PG_FAIL(PG_EINVAL, "library %d does not belong to the KS dialect", library_id);
    return launch_fd_terms(fd_dialect, library_id, U, T, A0, A1, P.c, terms_out, (cudaStream_t)stream);
}

static int fd_gather_impl(const double *U, const double *Uy, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt,
                          int fd_dialect, int library_id, const int64_t *flat_idx, int64_t n, double *X_out, double *y_out,
                          void *stream);

int pg_fd_gather_rows_two(const double *U, const double *Uy, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt,
                          int fd_dialect, int library_id, const int64_t *flat_idx, int64_t n, double *X_out, double *y_out,
                          void *stream) {
    if (!Uy) PG_FAIL(PG_EINVAL, "Uy is null (use pg_fd_gather_rows)");
    return fd_gather_impl(U, Uy, T, A0, A1, d0, d1, dt, fd_dialect, library_id, flat_idx, n, X_out, y_out, stream);
}

int pg_fd_gather_rows(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt,
                      int fd_dialect, int library_id, const int64_t *flat_idx, int64_t n, double *X_out, double *y_out,
                      void *stream) {
    return fd_gather_impl(U, nullptr, T, A0, A1, d0, d1, dt, fd_dialect, library_id, flat_idx, n, X_out, y_out, stream);
}

static int fd_gather_impl(const double *U, const double *Uy, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt,
                          int fd_dialect, int library_id, const int64_t *flat_idx, int64_t n, double *X_out, double *y_out,
                          void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    K1Params P{};
    int rc = describe(P, U, T, A0, A1, d0, d1, dt, fd_dialect, library_id, true);
    if (rc) return rc;
    P.Uy = Uy;
    if (n < 0) PG_FAIL(PG_EINVAL, "n < 0");
    if (n == 0) return PG_OK;
    if (!flat_idx || !X_out || !y_out) PG_FAIL(PG_EINVAL, "null buffer");
    if (T < 1 + P.t_halo) PG_FAIL(PG_EINVAL, "need at least %d frames for u_t", 1 + P.t_halo);
    void *scr = nullptr;
    rc = scratch_for(st, 64, &scr);
    if (rc) return rc;
    PG_CUDA(cudaMemsetAsync(scr, 0, 64, st));
    P.counters = (unsigned long long *)scr;
    return launch_fd_gather(library_id, P, flat_idx, n, X_out, y_out, st);
}

int pg_block_means(const double *stack, int k, int64_t T, int64_t A0, int64_t A1, int bt, int b0, int b1, double *out,
                   void *stream) {
    if (!stack || !out) PG_FAIL(PG_EINVAL, "null buffer");
    if (k < 1 || T < 1 || A0 < 1 || A1 < 1) PG_FAIL(PG_EINVAL, "bad shape");
    if (bt <= 0 || b0 <= 0 || b1 <= 0) PG_FAIL(PG_EINVAL, "block sizes must be > 0");
    return launch_block_means(stack, k, T, A0, A1, bt, b0, b1, out, (cudaStream_t)stream);
}

int pg_rows_gram(const double *X, const double *y, int64_t B, int64_t n, int p, int64_t ldx, const uint8_t *fold_of_row,
                 int n_folds, const double *shift, double *stats_out, double *colminmax_out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (p < 1 || p > PG_MAX_P) PG_FAIL(PG_EINVAL, "p must be in 1..%d", PG_MAX_P);
    if (n_folds < 1 || n_folds > PG_MAX_FOLDS) PG_FAIL(PG_EINVAL, "n_folds must be in 1..%d", PG_MAX_FOLDS);
    if (B < 0 || n < 0 || ldx < p) PG_FAIL(PG_EINVAL, "bad shape B=%lld n=%lld ldx=%lld", (long long)B, (long long)n, (long long)ldx);
    if (!stats_out) PG_FAIL(PG_EINVAL, "stats_out is null");
    if (B == 0) return PG_OK;
    if (n > 0 && (!X || !y)) PG_FAIL(PG_EINVAL, "null rows");
    const int S = PG_STATS_LEN(p);
    RowsParams R{};
    R.X = X; R.y = y; R.B = B; R.n = n; R.ldx = ldx; R.p = p; R.fold_of_row = fold_of_row; R.n_folds = n_folds;
    R.shift = shift;
    int64_t chunks = (n + GW_THREADS * 64 - 1) / (GW_THREADS * 64);  // ~64 rows per thread
    const int64_t cap = (int64_t)sm_count() * 8;
    if (chunks * B > cap) chunks = cap / B;
    if (chunks < 1) chunks = 1;
    R.chunks = (int)chunks;
    const int64_t parts = B * chunks * GW_WARPS;
    const size_t b_stats = sizeof(double) * (size_t)(parts * n_folds * S);
    const size_t b_mm = colminmax_out ? sizeof(double) * (size_t)(parts * n_folds * 2 * p) : 0;
    void *scr = nullptr;
    int rc = scratch_for(st, 64 + b_stats + b_mm, &scr);
    if (rc) return rc;
    PG_CUDA(cudaMemsetAsync(scr, 0, 64, st));
    R.counters = (unsigned long long *)scr;
    R.partials = (double *)((char *)scr + 64);
    R.mm_partials = colminmax_out ? (double *)((char *)scr + 64 + b_stats) : nullptr;
    return launch_rows_gram(R, stats_out, colminmax_out, st);
}

int pg_rows_gram_weighted(const double *X, const double *y, int64_t n, int p, int64_t ldx, const uint16_t *weights,
                          int64_t B, const double *shift, double *stats_out, double *colminmax_out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (p < 1 || p > PG_MAX_P) PG_FAIL(PG_EINVAL, "p must be in 1..%d", PG_MAX_P);
    if (B < 0 || n < 0 || ldx < p) PG_FAIL(PG_EINVAL, "bad shape B=%lld n=%lld ldx=%lld", (long long)B, (long long)n, (long long)ldx);
    if (!stats_out || !weights) PG_FAIL(PG_EINVAL, "null buffer");
    if (B == 0) return PG_OK;
    if (n > 0 && (!X || !y)) PG_FAIL(PG_EINVAL, "null rows");
    const int S = PG_STATS_LEN(p);
    RowsParams R{};
    R.X = X; R.y = y; R.B = B; R.n = n; R.ldx = ldx; R.p = p; R.fold_of_row = nullptr; R.n_folds = 1;
    R.shift = shift; R.weights = weights;
    int64_t chunks = (n + GW_THREADS * 64 - 1) / (GW_THREADS * 64);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (chunks * B > cap) chunks = cap / B;
    if (chunks < 1) chunks = 1;
    R.chunks = (int)chunks;
    const int64_t parts = B * chunks * GW_WARPS;
    const size_t b_stats = sizeof(double) * (size_t)(parts * S);
    const size_t b_mm = colminmax_out ? sizeof(double) * (size_t)(parts * 2 * p) : 0;
    void *scr = nullptr;
    int rc = scratch_for(st, 64 + b_stats + b_mm, &scr);
    if (rc) return rc;
    PG_CUDA(cudaMemsetAsync(scr, 0, 64, st));
    R.counters = (unsigned long long *)scr;
    R.partials = (double *)((char *)scr + 64);
    R.mm_partials = colminmax_out ? (double *)((char *)scr + 64 + b_stats) : nullptr;
    return launch_rows_gram(R, stats_out, colminmax_out, st);
}

int pg_poly_rows(const void *U, int dtype, int64_t T, int64_t H, int64_t W, const int32_t *pts, int64_t n,
                 const double *W6, int rt, int rs, int library_id, double *X_out, double *y_out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype != 0 && dtype != 1) PG_FAIL(PG_EINVAL, "dtype must be 0 (float32) or 1 (float64)");
    if (library_id != PG_LIB_PATCH_MODEL4 && library_id != PG_LIB_PATCH_FULL && library_id != PG_LIB_PATCH_DERIVS) PG_FAIL(PG_EINVAL, "library %d is not a patch library", library_id);
    if (rt < 0 || rs < 0 || rt > 8 || rs > 8) PG_FAIL(PG_EINVAL, "rt, rs must be in 0..8");
    if (n < 0) PG_FAIL(PG_EINVAL, "n < 0");
    if (n == 0) return PG_OK;
    if (!U || !pts || !W6 || !X_out || !y_out) PG_FAIL(PG_EINVAL, "null buffer");
    void *scr = nullptr;
    int rc = scratch_for(st, 64, &scr);
    if (rc) return rc;
    PG_CUDA(cudaMemsetAsync(scr, 0, 64, st));
    return launch_poly_rows(U, dtype, T, H, W, pts, n, W6, rt, rs, library_id == PG_LIB_PATCH_FULL ? 1 : (library_id == PG_LIB_PATCH_DERIVS ? 2 : 0), X_out, y_out,
                            (unsigned long long *)scr, st);
}

int pg_stridge_batched(const double *stats, int64_t B, int p, int dialect, int flags, const double *alphas, int na,
                       const double *thrs, int nt, int max_iter, const uint8_t *const_mask, const int8_t *signs,
                       const double *colminmax,
                       const double *shift, const double *eval_stats, double *coef_out, double *metrics_out,
                       int32_t *best_out, double *relres_out, void *stream) {
    if (p < 1 || p > PG_MAX_P) PG_FAIL(PG_EINVAL, "p must be in 1..%d", PG_MAX_P);
    if (dialect < PG_STRIDGE_KS || dialect > PG_STRIDGE_BASIC) PG_FAIL(PG_EINVAL, "unknown STRidge dialect %d", dialect);
    if (B < 0 || na < 1 || nt < 1 || max_iter < 0) PG_FAIL(PG_EINVAL, "bad sizes");
    if (!stats || !alphas || !thrs || !coef_out) PG_FAIL(PG_EINVAL, "null buffer");
    if (signs && dialect != PG_STRIDGE_KS) PG_FAIL(PG_EINVAL, "sign constraints belong to the ks2d dialect (ks2d:552-600)");
    if (dialect == PG_STRIDGE_BASIC && shift) PG_FAIL(PG_EINVAL, "the basic_usage dialect works on the raw Gram; shift must be null");
    if ((eval_stats != nullptr) != (metrics_out != nullptr)) PG_FAIL(PG_EINVAL, "eval_stats and metrics_out go together");
    if (best_out && !metrics_out) PG_FAIL(PG_EINVAL, "best_out needs eval_stats/metrics_out");
    if (relres_out && !metrics_out) PG_FAIL(PG_EINVAL, "relres_out needs eval_stats/metrics_out");
    StridgeParams P{};
    P.stats = stats; P.B = B; P.p = p; P.dialect = dialect; P.flags = flags; P.alphas = alphas; P.na = na;
    P.thrs = thrs; P.nt = nt; P.max_iter = max_iter; P.const_mask = const_mask; P.signs = signs; P.colminmax = colminmax;
    P.shift = shift; P.eval_stats = eval_stats; P.coef_out = coef_out; P.metrics_out = metrics_out;
    P.relres_out = relres_out;
    return launch_stridge(P, best_out, (cudaStream_t)stream);
}

int pg_ks_rollout(const double *U, int64_t T, int64_t A0, int64_t A1, double d0, double d1, double dt, int library_id,
                  const double *coef, int n_steps, double *work, double *rmse_out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    K1Params P{};
    int rc = describe(P, U, T, A0, A1, d0, d1, dt, PG_FD_KS_PERIODIC, library_id, true);
    if (rc) return rc;
    if (n_steps < 0 || n_steps > T - 1) PG_FAIL(PG_EINVAL, "n_steps must be in 0..T-1 (every step is compared with an observed frame)");
    if (n_steps == 0) return PG_OK;
    if (!coef || !work || !rmse_out) PG_FAIL(PG_EINVAL, "null buffer");
    const int blocks = rollout_blocks(A0, A1, sm_count());
    void *scr = nullptr;
    rc = scratch_for(st, sizeof(double) * (size_t)blocks * (size_t)n_steps, &scr);
    if (rc) return rc;
    return launch_rollout(library_id, U, A0, A1, P.c, coef, n_steps, work, (double *)scr, blocks, rmse_out, st);
}

int pg_ar_rollout(const double *U, int64_t T, int64_t H, int64_t W, double d0, double d1, double dt, const int32_t *term_ids,
                  const double *coef, int n_terms, int k_steps, int64_t t0, int64_t t1, const uint8_t *mask, double *work,
                  double *sums_out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!U || !term_ids || !coef || !work || !sums_out) PG_FAIL(PG_EINVAL, "null buffer");
    if (T < 1 || H < 2 || W < 2) PG_FAIL(PG_EINVAL, "bad shape (reflect padding needs H, W >= 2)");
    if (!(d0 > 0) || !(d1 > 0) || !(dt > 0)) PG_FAIL(PG_EINVAL, "grid spacings must be positive");
    if (n_terms < 1 || n_terms > PG_MAX_P) PG_FAIL(PG_EINVAL, "n_terms must be in 1..%d", PG_MAX_P);
    if (k_steps < 1) PG_FAIL(PG_EINVAL, "k_steps must be >= 1");
    if (t0 < 0 || t1 > T || t1 - t0 <= k_steps) PG_FAIL(PG_EINVAL, "the slice [t0, t1) must lie inside the stack and hold more than k_steps frames");
    const int64_t n_start = t1 - k_steps - t0;
    const int blocks = rollout_blocks(n_start * H, W, sm_count());
    void *scr = nullptr;
    int rc = scratch_for(st, sizeof(double) * 4 * (size_t)blocks, &scr);
    if (rc) return rc;
    return launch_ar_rollout(U, H, W, make_consts(d0, d1, dt), term_ids, coef, n_terms, k_steps, t0, n_start, mask, work,
                             (double *)scr, blocks, sums_out, st);
}

int pg_one_step_ss(const double *u_field, const double *ut_pred, int64_t t_max, int64_t frame, double dt, const uint8_t *mask,
                   double *sums_out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!u_field || !ut_pred || !sums_out) PG_FAIL(PG_EINVAL, "null buffer");
    if (t_max < 1 || frame < 1) PG_FAIL(PG_EINVAL, "t_max and frame must be >= 1");
    const int blocks = rollout_blocks(t_max, frame, sm_count());
    void *scr = nullptr;
    int rc = scratch_for(st, sizeof(double) * 2 * (size_t)blocks, &scr);
    if (rc) return rc;
    return launch_one_step(u_field, ut_pred, t_max, frame, dt, mask, (double *)scr, blocks, sums_out, st);
}

int pg_fit_metrics(const double *y_true, const double *y_pred, int64_t n, double *sums_out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n < 1) PG_FAIL(PG_EINVAL, "n must be >= 1");
    if (!y_true || !y_pred || !sums_out) PG_FAIL(PG_EINVAL, "null buffer");
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 4;
    const int blocks = (int)(g > cap ? cap : g);
    void *scr = nullptr;
    int rc = scratch_for(st, sizeof(double) * 5 * (size_t)blocks, &scr);
    if (rc) return rc;
    return launch_fit_metrics(y_true, y_pred, n, (double *)scr, blocks, sums_out, st);
}

int pg_rows_metrics_batched(const double *X, const double *y, const double *coef, int64_t B, int64_t n, int p, int64_t ldx,
                            double *sums_out, double *resid_out, void *stream) {
    if (p < 1 || p > PG_MAX_P) PG_FAIL(PG_EINVAL, "p must be in 1..%d", PG_MAX_P);
    if (B < 0 || n < 1 || ldx < p) PG_FAIL(PG_EINVAL, "bad shape B=%lld n=%lld ldx=%lld", (long long)B, (long long)n, (long long)ldx);
    if (B == 0) return PG_OK;
    if (!X || !y || !coef || !sums_out) PG_FAIL(PG_EINVAL, "null buffer");
    return launch_rows_metrics_batched(X, y, coef, B, n, p, ldx, sums_out, resid_out, (cudaStream_t)stream);
}

int pg_reflect_conv(const void *in, int dtype, int64_t T, int64_t A0, int64_t A1, int axis, const double *weights, int radius,
                    void *out, void *stream) {
    if (!in || !out || !weights) PG_FAIL(PG_EINVAL, "null buffer");
    if (dtype != 0 && dtype != 1) PG_FAIL(PG_EINVAL, "dtype must be 0 (float32) or 1 (float64)");
    if (T < 1 || A0 < 1 || A1 < 1 || radius < 0) PG_FAIL(PG_EINVAL, "bad shape");
    if (axis != 0 && axis != 1) PG_FAIL(PG_EINVAL, "axis must be 0 or 1");
    if (in == out) PG_FAIL(PG_EINVAL, "in-place filtering is not supported");
    return launch_reflect_conv(in, dtype, T, A0, A1, axis, weights, radius, out, (cudaStream_t)stream);
}

int pg_reflect_gauss2d(const void *in, int dtype, int64_t T, int64_t A0, int64_t A1, const double *weights, int radius, void *out,
                       void *stream) {
    if (!in || !out || !weights) PG_FAIL(PG_EINVAL, "null buffer");
    if (dtype != 0 && dtype != 1) PG_FAIL(PG_EINVAL, "dtype must be 0 (float32) or 1 (float64)");
    if (T < 1 || A0 < 1 || A1 < 1 || radius < 0) PG_FAIL(PG_EINVAL, "bad shape");
    if (radius > 32) PG_FAIL(PG_EINVAL, "radius > 32: use two pg_reflect_conv passes");
    if (in == out) PG_FAIL(PG_EINVAL, "in-place filtering is not supported");
    const int rc = launch_reflect_gauss2d(in, dtype, T, A0, A1, weights, radius, out, (cudaStream_t)stream);
    if (rc != PG_OK) PG_FAIL(rc, "stack too large for one launch");
    return rc;
}

int pg_time_moving_average(const double *U, int64_t T, int64_t A0, int64_t A1, int window, double *out, void *stream) {
    if (!U || !out) PG_FAIL(PG_EINVAL, "null buffer");
    if (T < 1 || A0 < 1 || A1 < 1) PG_FAIL(PG_EINVAL, "bad shape");
    if (window < 1 || window % 2 == 0) PG_FAIL(PG_EINVAL, "time smoothing window must be odd");   // ks2d:152-153
    if (window / 2 > T - 1) PG_FAIL(PG_EINVAL, "reflect padding needs window // 2 <= T - 1");      // np.pad(mode='reflect')
    if (U == out) PG_FAIL(PG_EINVAL, "in-place smoothing is not supported");
    return launch_time_moving_average(U, T, A0, A1, window, out, (cudaStream_t)stream);
}

int pg_periodic_conv(const double *in, int64_t T, int64_t A0, int64_t A1, int axis, const int32_t *offsets,
                     const double *weights, int n_taps, double *out, void *stream) {
    if (!in || !out || !offsets || !weights) PG_FAIL(PG_EINVAL, "null buffer");
    if (T < 1 || A0 < 1 || A1 < 1 || n_taps < 1) PG_FAIL(PG_EINVAL, "bad shape");
    if (axis != 0 && axis != 1) PG_FAIL(PG_EINVAL, "axis must be 0 or 1");
    if (in == out) PG_FAIL(PG_EINVAL, "in-place convolution is not supported");
    return launch_periodic_conv(in, T, A0, A1, axis, offsets, weights, n_taps, out, (cudaStream_t)stream);
}

int pg_periodic_gaussian_fft(const double *in, int64_t T, int64_t A0, int64_t A1, double sigma_px, double *out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!in || !out) PG_FAIL(PG_EINVAL, "null buffer");
    if (T < 1 || A0 < 2 || A1 < 2 || A0 > 0x7fffffff || A1 > 0x7fffffff) PG_FAIL(PG_EINVAL, "bad shape");
    if (!(sigma_px > 0.0)) PG_FAIL(PG_EINVAL, "sigma_px must be > 0");
    if (in == out) PG_FAIL(PG_EINVAL, "in-place smoothing is not supported");
    // H(kx, ky) = exp(-sigma^2 (kx^2 + ky^2) / 2) with k = 2 pi fftfreq(n) (ks2d:133-136), split into its two factors
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<double> hx((size_t)A0), hy((size_t)(A1 / 2 + 1));
    for (int64_t i = 0; i < A0; ++i) {
        const double f = (double)(i < (A0 + 1) / 2 ? i : i - A0) / (double)A0, k = two_pi * f;
        hx[(size_t)i] = exp(-0.5 * sigma_px * sigma_px * k * k) / ((double)A0 * (double)A1);
    }
    for (int64_t j = 0; j <= A1 / 2; ++j) {
        const double f = (double)(j < (A1 + 1) / 2 ? j : j - A1) / (double)A1, k = two_pi * f;
        hy[(size_t)j] = exp(-0.5 * sigma_px * sigma_px * k * k);
    }
    int64_t batch = 1;
    const size_t bytes = periodic_gaussian_fft_scratch(T, A0, A1, &batch);
    void *scr = nullptr;
    int rc = scratch_for(st, bytes, &scr);
    if (rc) return rc;
    rc = launch_periodic_gaussian_fft(in, T, A0, A1, hx.data(), hy.data(), out, scr, batch, st);
    if (rc) return rc;
    PG_CUDA(cudaStreamSynchronize(st));     // hx / hy are host vectors copied asynchronously
    return PG_OK;
}

int pg_synth_field(double *U, int64_t T, int64_t A0, int64_t A1, int64_t t_offset, int64_t T_total, uint64_t seed,
                   int kind, double noise, void *stream) {
    if (!U) PG_FAIL(PG_EINVAL, "U is null");
    if (T < 0 || A0 < 1 || A1 < 1) PG_FAIL(PG_EINVAL, "bad shape");
    return launch_synth(U, T, A0, A1, t_offset, T_total, seed, kind, noise, (cudaStream_t)stream);
}

}  // extern "C"
