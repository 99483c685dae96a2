// Generic (any shape, reference-arithmetic) kernels of libpdegram:
//   k1_generic      fused FD + library row + block mean + Gram for any dialect / block / fold layout
//   fd_terms        materialised term stacks, bit-identical to the reference's NumPy arithmetic
//   fd_gather       sampled pointwise rows
//   block_means     build_blockwise_dataset on caller-supplied stacks
//   rows_gram       rows -> statistics (+ per-column min/max)
//   reduce_partials fixed-order compensated reduction of per-warp partial statistics
// The tiled TMA kernels in tiled.cu are the fast path; these cover every other shape, the
// ragged edge blocks the tiled kernels leave out, and serve as their on-device cross-check.
#include <math.h>

#include "common.cuh"
#include "launch.h"

namespace pg {

constexpr int GW = GW_WARPS;

template <int LIB>
__device__ __forceinline__ void eval_point(const K1Params &P, const double *F, int64_t i, int64_t j, PointVals &v) {
    if constexpr (LIB == PG_LIB_BASIC) {
        basic_point(F, P.A1, i + P.off, j + P.off, P.c, v);
    } else if constexpr (LIB == PG_LIB_AR_FULL) {
        slice_point(F, P.A1, i, j, P.c, v);
    } else {
        ks_point<Lib<LIB>::BIH>(F, P.A0, P.A1, i, j, P.c, v);
    }
}

__device__ __forceinline__ void fill_pairs(uint8_t *pa, uint8_t *pb, int p, int S) {
    for (int e = threadIdx.x; e < S; e += blockDim.x) {
        int a, b;
        stats_pair(e, p, a, b);
        pa[e] = (uint8_t)a;
        pb[e] = (uint8_t)b;
    }
}

// One block row (reference arithmetic): the block means th[0..p), y of block `idx` of the launch's index box, its
// fold id and whether it is a row at all (finite, fold in range).  Shared by the statistics kernel and the
// held-out-residual kernel.
template <int LIB>
__device__ __forceinline__ bool generic_block_row(const K1Params &P, int64_t idx, int64_t n1, int64_t n2, double *th, double &y,
                                                  int &fold, unsigned long long &bad_rows, unsigned long long &bad_fold) {
    constexpr int p = Lib<LIB>::P;
    const int64_t frame = P.A0 * P.A1;
    const int64_t Trows = P.T - P.t_halo;
    const int64_t jb = P.i1_lo + idx % n2;
    const int64_t ib = P.i0_lo + (idx / n2) % n1;
    const int64_t tb = P.tb_lo + idx / (n2 * n1);
    int64_t t0 = tb * P.bt, t1 = min(Trows, t0 + P.bt);
    int64_t i0 = ib * P.b0, i1 = min(P.R0, i0 + P.b0);
    int64_t j0 = jb * P.b1, j1 = min(P.R1, j0 + P.b1);
    if (P.rows8) {
        // second stage of the tiled path for (bt, 8m, 8n) blocks: the block mean is the mean of the equally
        // sized (bt, 8, 8) sub-block means the tiled kernel wrote (ragged edge blocks hold fewer of them)
        const int64_t s0 = i0 >> 3, s1 = (i1 + 7) >> 3, c0 = j0 >> 3, c1 = (j1 + 7) >> 3;
        for (int64_t a = s0; a < s1; ++a)
            for (int64_t b = c0; b < c1; ++b) {
                const double *r = P.rows8 + ((tb * P.sub0 + a) * P.sub1 + b) * (p + 1);
                y = __dadd_rn(y, r[0]);
#pragma unroll
                for (int k = 0; k < p; ++k) th[k] = __dadd_rn(th[k], r[1 + k]);
            }
        t1 = t0 + 1; i0 = s0; i1 = s1; j0 = c0; j1 = c1;      // the divisor below: the number of sub-blocks
    } else
    for (int64_t t = t0; t < t1; ++t) {
        const double *F = P.U + t * frame;
        // forward u_t = (U[t+1] - U[t]) / dt (ks2d:1511, basic:46-48); central (U[t+2] - U[t]) / (2 dt) (analyze_results:261)
        const double *Fy = P.Uy ? P.Uy + t * frame : F;          // the stack u_t is taken of (the same one unless Uy is given)
        const double *Fn = Fy + P.t_halo * frame;
        const double tdiv = P.t_halo == 2 ? __dmul_rn(2.0, P.c.dt) : P.c.dt;
        for (int64_t i = i0; i < i1; ++i)
            for (int64_t j = j0; j < j1; ++j) {
                PointVals v;
                eval_point<LIB>(P, F, i, j, v);
                double row[p];
                lib_row<LIB>(v, row);
                const int64_t o = (i + P.off) * P.A1 + (j + P.off);
                y = __dadd_rn(y, __ddiv_rn(__dsub_rn(Fn[o], Fy[o]), tdiv));
#pragma unroll
                for (int k = 0; k < p; ++k) th[k] = __dadd_rn(th[k], row[k]);
            }
    }
    const double cnt = (double)((t1 - t0) * (i1 - i0) * (j1 - j0));
    y = __ddiv_rn(y, cnt);
    bool fin = isfinite(y);
#pragma unroll
    for (int k = 0; k < p; ++k) {
        th[k] = __ddiv_rn(th[k], cnt);
        fin = fin && isfinite(th[k]);
    }
    if (P.fold_of_row) { fold = P.fold_of_row[(tb * P.nB0 + ib) * P.nB1 + jb]; if (fold == 255) fold = -1; }
    else if (P.fold_of_frame) fold = P.fold_of_frame[t0];
    if (!fin) { ++bad_rows; return false; }
    if (fold < 0) return false;                         // excluded on purpose (-1 / 255): not an error
    if (fold >= P.n_folds) { ++bad_fold; return false; }
    return true;
}

template <int LIB>
__global__ void __launch_bounds__(GW * 32) k1_generic_kernel(K1Params P) {
    constexpr int p = Lib<LIB>::P;
    constexpr int S = PG_STATS_LEN(p);
    extern __shared__ double sm[];
    if (P.run_if && *P.run_if == 0) return;                  // conditional fallback launch: nothing to redo
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *wacc_all = sm;                                   // [GW][n_folds][S]
    double *ext_all = wacc_all + GW * P.n_folds * S;         // [GW][32][p+2]
    uint8_t *pa = reinterpret_cast<uint8_t *>(ext_all + GW * 32 * (p + 2));
    uint8_t *pb = pa + S;
    for (int e = threadIdx.x; e < GW * P.n_folds * S; e += blockDim.x) wacc_all[e] = 0.0;
    fill_pairs(pa, pb, p, S);
    __syncthreads();
    double *wacc = wacc_all + warp * P.n_folds * S;
    double *ext = ext_all + warp * 32 * (p + 2);

    const int64_t n0 = P.tb_hi - P.tb_lo, n1 = P.i0_hi - P.i0_lo, n2 = P.i1_hi - P.i1_lo;
    const int64_t total = n0 * n1 * n2;
    const int64_t stride = (int64_t)gridDim.x * GW * 32;
    unsigned long long bad_rows = 0, bad_fold = 0;

    for (int64_t base = ((int64_t)blockIdx.x * GW + warp) * 32; base < total; base += stride) {
        const int64_t idx = base + lane;
        bool valid = idx < total;
        double th[p], y = 0.0;
        int fold = 0;
#pragma unroll
        for (int k = 0; k < p; ++k) th[k] = 0.0;
        if (valid) valid = generic_block_row<LIB>(P, idx, n1, n2, th, y, fold, bad_rows, bad_fold);
        warp_accumulate_rows(wacc, ext, pa, pb, p, S, lane, valid, fold, th, y);
    }
    if (bad_rows) atomicAdd(&P.counters[0], bad_rows);
    if (bad_fold) atomicAdd(&P.counters[1], bad_fold);
    double *out = P.partials + ((int64_t)blockIdx.x * GW + warp) * P.n_folds * S;
    for (int e = lane; e < P.n_folds * S; e += 32) out[e] = wacc[e];
}

// The block-mean rows themselves, [nrows][p + 1] (y first) in reference row order, nothing dropped: the rare path on
// which the caller has to renumber rows after dropping non-finite ones (ks2d:394-395 followed by ks2d:1638-1641).
template <int LIB>
__global__ void __launch_bounds__(GW * 32) k1_generic_rows_kernel(K1Params P, double *__restrict__ rows_out) {
    constexpr int p = Lib<LIB>::P;
    const int64_t n0 = P.tb_hi - P.tb_lo, n1 = P.i0_hi - P.i0_lo, n2 = P.i1_hi - P.i1_lo;
    const int64_t total = n0 * n1 * n2;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        double th[p], y = 0.0;
        int fold = 0;
        unsigned long long b0 = 0, b1 = 0;
#pragma unroll
        for (int k = 0; k < p; ++k) th[k] = 0.0;
        generic_block_row<LIB>(P, idx, n1, n2, th, y, fold, b0, b1);
        double *r = rows_out + idx * (p + 1);
        r[0] = y;
#pragma unroll
        for (int k = 0; k < p; ++k) r[1 + k] = th[k];
    }
}

// Held-out residuals of fitted models straight from the field (second pass; SURVEY 8a21: r2 / rmse of ks2d:29-40 need
// sum (y - theta.c)^2, which the statistics only give up to cancellation when the fit is exact to ~1e-8).  Every
// thread forms its block row like the statistics kernel and adds r_j^2 for each of the J <= 32 coefficient vectors
// when the row belongs to fold `eval_fold` (-1: every row).  partials [gridDim.x][J + 1] (last entry: row count).
constexpr int RESID_MAX_J = 32;

template <int LIB>
__global__ void __launch_bounds__(GW * 32) k1_generic_resid_kernel(K1Params P, const double *__restrict__ coef, int J,
                                                                    int eval_fold, double *__restrict__ partials) {
    constexpr int p = Lib<LIB>::P;
    __shared__ double cs[RESID_MAX_J * PG_MAX_P];
    __shared__ double red[GW][RESID_MAX_J + 1];
    for (int e = threadIdx.x; e < J * p; e += blockDim.x) cs[e] = coef[e];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n0 = P.tb_hi - P.tb_lo, n1 = P.i0_hi - P.i0_lo, n2 = P.i1_hi - P.i1_lo;
    const int64_t total = n0 * n1 * n2;
    const int64_t stride = (int64_t)gridDim.x * GW * 32;
    unsigned long long bad_rows = 0, bad_fold = 0;
    double acc[RESID_MAX_J + 1];
#pragma unroll
    for (int j = 0; j <= RESID_MAX_J; ++j) acc[j] = 0.0;
    for (int64_t idx = ((int64_t)blockIdx.x * GW + warp) * 32 + lane; idx < total; idx += stride) {
        double th[p], y = 0.0;
        int fold = 0;
#pragma unroll
        for (int k = 0; k < p; ++k) th[k] = 0.0;
        if (!generic_block_row<LIB>(P, idx, n1, n2, th, y, fold, bad_rows, bad_fold)) continue;
        if (eval_fold >= 0 && fold != eval_fold) continue;
        acc[RESID_MAX_J] += 1.0;
#pragma unroll
        for (int j = 0; j < RESID_MAX_J; ++j) {
            if (j >= J) break;
            double pred = 0.0;
#pragma unroll
            for (int k = 0; k < p; ++k) pred = fma(th[k], cs[j * p + k], pred);
            const double r = y - pred;
            acc[j] = fma(r, r, acc[j]);
        }
    }
    if (bad_fold) atomicAdd(&P.counters[1], bad_fold);
    // fixed-order reduction: lanes by butterfly, warps in order
#pragma unroll
    for (int j = 0; j <= RESID_MAX_J; ++j) {
        double v = acc[j];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x <= RESID_MAX_J) {
        double v = 0.0;
        for (int w = 0; w < GW; ++w) v += red[w][threadIdx.x];
        if (threadIdx.x < J) partials[(int64_t)blockIdx.x * (J + 1) + threadIdx.x] = v;
        else if (threadIdx.x == RESID_MAX_J) partials[(int64_t)blockIdx.x * (J + 1) + J] = v;
    }
}

// Rows form of the same: X [n][ldx], y [n], optional fold byte per row.
__global__ void __launch_bounds__(GW * 32) rows_resid_kernel(const double *__restrict__ X, const double *__restrict__ y,
                                                             int64_t n, int p, int64_t ldx, const uint8_t *__restrict__ fold_of_row,
                                                             int eval_fold, const double *__restrict__ coef, int J,
                                                             double *__restrict__ partials) {
    __shared__ double cs[RESID_MAX_J * PG_MAX_P];
    __shared__ double red[GW][RESID_MAX_J + 1];
    for (int e = threadIdx.x; e < J * p; e += blockDim.x) cs[e] = coef[e];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[RESID_MAX_J + 1];
#pragma unroll
    for (int j = 0; j <= RESID_MAX_J; ++j) acc[j] = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        if (fold_of_row && eval_fold >= 0 && fold_of_row[r] != eval_fold) continue;
        acc[RESID_MAX_J] += 1.0;
        const double *x = X + r * ldx;
        const double yr = y[r];
#pragma unroll
        for (int j = 0; j < RESID_MAX_J; ++j) {
            if (j >= J) break;
            double pred = 0.0;
            for (int k = 0; k < p; ++k) pred = fma(x[k], cs[j * p + k], pred);
            const double d = yr - pred;
            acc[j] = fma(d, d, acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j <= RESID_MAX_J; ++j) {
        double v = acc[j];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x <= RESID_MAX_J) {
        double v = 0.0;
        for (int w = 0; w < GW; ++w) v += red[w][threadIdx.x];
        if (threadIdx.x < J) partials[(int64_t)blockIdx.x * (J + 1) + threadIdx.x] = v;
        else if (threadIdx.x == RESID_MAX_J) partials[(int64_t)blockIdx.x * (J + 1) + J] = v;
    }
}

// out[e] (+)= sum over parts of partials[part][e].  One CTA per entry: thread k adds parts
// k, k+128, ... (Kahan-compensated), then a fixed-order tree in shared memory, so the result does
// not depend on scheduling (run-to-run bit-identical).
__global__ void __launch_bounds__(128) reduce_partials_kernel(const double *__restrict__ partials, int64_t n_parts,
                                                              int64_t len, double *__restrict__ out, int accumulate,
                                                              const unsigned long long *__restrict__ flag, int64_t n_a,
                                                              int64_t n_b, const unsigned long long *__restrict__ poison) {
    __shared__ double sh[128];
    const int64_t e = blockIdx.x;
    double s = 0.0, comp = 0.0;
    int64_t skip_lo = 0, skip_hi = 0;
    if (flag) {
        if (*flag != 0) skip_hi = n_a;
        else { skip_lo = n_a; skip_hi = n_a + n_b; }
    }
    for (int64_t k = threadIdx.x; k < n_parts; k += 128) {
        if (k >= skip_lo && k < skip_hi) continue;
        const double x = __dsub_rn(partials[k * len + e], comp);
        const double t = __dadd_rn(s, x);
        comp = __dsub_rn(__dsub_rn(t, s), x);
        s = t;
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = 64; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] = __dadd_rn(sh[threadIdx.x], sh[threadIdx.x + w]);
        __syncthreads();
    }
    // rows with an out-of-range fold id (counters[1]) or a halo frame that never arrived (counters[3]) would silently
    // shrink / corrupt the statistics: make them loud instead (no host synchronisation needed to notice)
    if (threadIdx.x == 0) {
        double v = sh[0];
        if (poison && (poison[1] | poison[3])) v = nan("");
        out[e] = accumulate ? __dadd_rn(out[e], v) : v;
    }
}

// ----------------------------------------------------------------------------- term stacks
template <int LIB>
__global__ void fd_terms_ks_kernel(const double *__restrict__ U, int64_t T, int64_t A0, int64_t A1, FdConsts c,
                                   double *__restrict__ out) {
    constexpr int p = Lib<LIB>::P;
    const int64_t frame = A0 * A1, total = T * frame;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = idx / frame, r = idx % frame;
        PointVals v;
        ks_point<Lib<LIB>::BIH>(U + t * frame, A0, A1, r / A1, r % A1, c, v);
        double row[p];
        lib_row<LIB>(v, row);
#pragma unroll
        for (int k = 0; k < p; ++k) out[k * total + idx] = row[k];
    }
}

// compute_derivatives (basic:32-72): out = [u_t, u, u_x, u_y, lap], each [T-1][A0-4][A1-4]
__global__ void fd_terms_basic_kernel(const double *__restrict__ U, int64_t T, int64_t A0, int64_t A1, FdConsts c,
                                      double *__restrict__ out) {
    const int64_t R0 = A0 - 4, R1 = A1 - 4, frame = A0 * A1, total = (T - 1) * R0 * R1;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = idx / (R0 * R1), r = idx % (R0 * R1);
        const int64_t i = r / R1 + 2, j = r % R1 + 2;
        const double *F = U + t * frame;
        PointVals v;
        basic_point(F, A1, i, j, c, v);
        out[0 * total + idx] = __ddiv_rn(__dsub_rn(F[frame + i * A1 + j], F[i * A1 + j]), c.dt);
        out[1 * total + idx] = v.u;
        out[2 * total + idx] = v.g1;
        out[3 * total + idx] = v.g0;
        out[4 * total + idx] = v.lap;
    }
}

template <int LIB>
__global__ void fd_gather_kernel(K1Params P, const int64_t *__restrict__ flat_idx, int64_t n, double *__restrict__ X,
                                 double *__restrict__ y) {
    constexpr int p = Lib<LIB>::P;
    const int64_t frame = P.A0 * P.A1, rs = P.R0 * P.R1, nrows = (P.T - P.t_halo) * rs;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t idx = flat_idx[k];
        if (idx < 0 || idx >= nrows) {  // caller error; poison the row so it cannot pass silently
            for (int c = 0; c < p; ++c) X[k * p + c] = nan("");
            y[k] = nan("");
            atomicAdd(&P.counters[1], 1ull);
            continue;
        }
        const int64_t t = idx / rs, r = idx % rs, i = r / P.R1, j = r % P.R1;
        const double *F = P.U + t * frame;
        PointVals v;
        eval_point<LIB>(P, F, i, j, v);
        double row[p];
        lib_row<LIB>(v, row);
#pragma unroll
        for (int c = 0; c < p; ++c) X[k * p + c] = row[c];
        const int64_t o = (i + P.off) * P.A1 + (j + P.off);
        const double *Fy = P.Uy ? P.Uy + t * frame : F;
        y[k] = __ddiv_rn(__dsub_rn(Fy[P.t_halo * frame + o], Fy[o]), P.t_halo == 2 ? __dmul_rn(2.0, P.c.dt) : P.c.dt);
    }
}

// ----------------------------------------------------------------------------- block means of given stacks
__global__ void block_means_kernel(const double *__restrict__ stack, int k, int64_t T, int64_t A0, int64_t A1, int bt,
                                   int b0, int b1, int64_t nBt, int64_t nB0, int64_t nB1, double *__restrict__ out) {
    const int64_t nrows = nBt * nB0 * nB1, total = nrows * k, vol = T * A0 * A1;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / k;
        const int a = (int)(idx % k);
        const int64_t jb = row % nB1, ib = (row / nB1) % nB0, tb = row / (nB1 * nB0);
        const int64_t t0 = tb * bt, t1 = min(T, t0 + bt), i0 = ib * b0, i1 = min(A0, i0 + b0), j0 = jb * b1,
                      j1 = min(A1, j0 + b1);
        const double *src = stack + (int64_t)a * vol;
        double s = 0.0;
        for (int64_t t = t0; t < t1; ++t)
            for (int64_t i = i0; i < i1; ++i)
                for (int64_t j = j0; j < j1; ++j) s = __dadd_rn(s, src[(t * A0 + i) * A1 + j]);
        out[idx] = __ddiv_rn(s, (double)((t1 - t0) * (i1 - i0) * (j1 - j0)));
    }
}

// ----------------------------------------------------------------------------- rows -> statistics
__global__ void __launch_bounds__(GW * 32) rows_gram_kernel(RowsParams P) {
    const int p = P.p, S = PG_STATS_LEN(p), W = p + 2;
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *wacc_all = sm;                                 // [GW][n_folds][S]
    double *ext_all = wacc_all + GW * P.n_folds * S;       // [GW][32][p+2]
    double *mm_all = ext_all + GW * 32 * W;                // [GW][n_folds][2][p]
    uint8_t *pa = reinterpret_cast<uint8_t *>(mm_all + GW * P.n_folds * 2 * p);
    uint8_t *pb = pa + S;
    for (int e = threadIdx.x; e < GW * P.n_folds * S; e += blockDim.x) wacc_all[e] = 0.0;
    for (int e = threadIdx.x; e < GW * P.n_folds * 2 * p; e += blockDim.x)
        mm_all[e] = ((e / p) & 1) ? -INFINITY : INFINITY;
    fill_pairs(pa, pb, p, S);
    __syncthreads();
    double *wacc = wacc_all + warp * P.n_folds * S;
    double *ext = ext_all + warp * 32 * W;
    double *mm = mm_all + warp * P.n_folds * 2 * p;

    const int64_t b = blockIdx.x / P.chunks;
    const int chunk = (int)(blockIdx.x % P.chunks);
    // weighted mode (bootstrap resamples, ks2d:603-642): every problem reads the same rows with its own multiplicities
    const uint16_t *wb = P.weights ? P.weights + b * P.n : nullptr;
    const double *Xb = P.X + (wb ? 0 : b * P.n * P.ldx);
    const double *yb = P.y + (wb ? 0 : b * P.n);
    const uint8_t *fb = P.fold_of_row ? P.fold_of_row + b * P.n : nullptr;
    const double *sh = P.shift ? P.shift + (wb ? 0 : b * p) : nullptr;
    const int64_t stride = (int64_t)P.chunks * GW * 32;
    unsigned long long bad_fold = 0;
    for (int64_t base = ((int64_t)chunk * GW + warp) * 32; base < P.n; base += stride) {
        const int64_t r = base + lane;
        bool valid = r < P.n;
        int fold = 0;
        double wgt = 1.0;
        if (valid && wb) {
            wgt = (double)wb[r];
            valid = wgt > 0.0;              // rows the resample did not draw do not exist for it (min / max included)
        }
        ext[lane * W] = 1.0;
        if (valid) {
            ext[lane * W + 1] = yb[r];
            for (int j = 0; j < p; ++j) ext[lane * W + 2 + j] = sh ? __dsub_rn(Xb[r * P.ldx + j], sh[j]) : Xb[r * P.ldx + j];
            if (fb) fold = fb[r];
            if (fold == 255) valid = false;                           // excluded on purpose
            else if (fold >= P.n_folds) { valid = false; ++bad_fold; }
        }
        __syncwarp();
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        for (int l = 0; l < 32; ++l) {
            if (!((vmask >> l) & 1u)) continue;
            const int f = __shfl_sync(0xffffffffu, fold, l);
            const double wl = __shfl_sync(0xffffffffu, wgt, l);
            const double *row = ext + l * W;
            double *acc = wacc + f * S;
            for (int e = lane; e < S; e += 32) acc[e] = fma(row[pa[e]] * wl, row[pb[e]], acc[e]);
            if (lane < p) {
                // min/max of the UNSHIFTED column values
                const double xv = sh ? Xb[(base + l) * P.ldx + lane] : row[2 + lane];
                double *m = mm + f * 2 * p;
                m[lane] = fmin(m[lane], xv);
                m[p + lane] = fmax(m[p + lane], xv);
            }
        }
        __syncwarp();
    }
    if (bad_fold) atomicAdd(&P.counters[1], bad_fold);
    const int64_t part = (b * P.chunks + chunk) * GW + warp;
    double *out = P.partials + part * P.n_folds * S;
    for (int e = lane; e < P.n_folds * S; e += 32) out[e] = wacc[e];
    if (P.mm_partials) {
        double *mo = P.mm_partials + part * P.n_folds * 2 * p;
        for (int e = lane; e < P.n_folds * 2 * p; e += 32) mo[e] = mm[e];
    }
}

// per problem: stats[b][e] = sum_k partials[b][k][e];  colminmax likewise with min / max
__global__ void rows_reduce_kernel(const double *__restrict__ partials, const double *__restrict__ mm_partials,
                                   int64_t B, int parts, int len, int mmlen, int p, double *__restrict__ stats,
                                   double *__restrict__ colminmax, const unsigned long long *__restrict__ poison) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < B * len) {
        const int64_t b = idx / len, e = idx % len;
        double s = 0.0, comp = 0.0;
        for (int k = 0; k < parts; ++k) {
            const double x = __dsub_rn(partials[(b * parts + k) * len + e], comp);
            const double t = __dadd_rn(s, x);
            comp = __dsub_rn(__dsub_rn(t, s), x);
            s = t;
        }
        stats[idx] = (poison && poison[1]) ? nan("") : s;   // a fold id >= n_folds is a caller error: loud, not silent
    }
    if (colminmax && idx < B * mmlen) {
        const int64_t b = idx / mmlen, e = idx % mmlen;
        const bool is_max = ((e / p) & 1) != 0;
        double m = is_max ? -INFINITY : INFINITY;
        for (int k = 0; k < parts; ++k) {
            const double x = mm_partials[(b * parts + k) * mmlen + e];
            m = is_max ? fmax(m, x) : fmin(m, x);
        }
        colminmax[idx] = m;
    }
}

// patch_based_sindy.py: per-patch periodic-roll differences and the 11-term library
//   1, u, u_x, u_y, u_xx, u_yy, lap, u^2, u u_x, u u_y, u lap      (sindy:226-270; x = a1, np.roll wraps INSIDE the patch)
// The reference column_stacks the eleven (h, w) term arrays into an (h, 11 w) array and then views it as (h, w, 11)
// (sindy:269, 327-329): the "feature k of pixel (r, c)" it regresses on is therefore element c*11 + k of row r of the
// concatenation, i.e. term (c*11 + k) / w at column (c*11 + k) % w -- eleven consecutive samples of one or two terms,
// not the eleven terms at the pixel.  A faithful port has to reproduce that index map; sindy_term() below is the one
// place where it lives.
__device__ __forceinline__ double sindy_term(const double *__restrict__ F, int64_t ld, int ps, int r, int c, int tid,
                                             const FdConsts &k) {
    const int rm = r == 0 ? ps - 1 : r - 1, rp = r == ps - 1 ? 0 : r + 1, cm = c == 0 ? ps - 1 : c - 1, cp = c == ps - 1 ? 0 : c + 1;
    const double u = F[r * ld + c];
    if (tid == 0) return 1.0;
    if (tid == 1) return u;
    if (tid == 7) return __dmul_rn(u, u);
    const double ue = F[r * ld + cp], uw = F[r * ld + cm], un = F[rp * ld + c], us = F[rm * ld + c];
    const double ux = central_diff(ue, uw, k.two_d1), uy = central_diff(un, us, k.two_d0);
    const double uxx = second_diff(ue, u, uw, k.d1sq), uyy = second_diff(un, u, us, k.d0sq);
    switch (tid) {
        case 2: return ux;
        case 3: return uy;
        case 4: return uxx;
        case 5: return uyy;
        case 6: return __dadd_rn(uxx, uyy);
        case 8: return __dmul_rn(u, ux);
        case 9: return __dmul_rn(u, uy);
        default: return __dmul_rn(u, __dadd_rn(uxx, uyy));
    }
}

// rows of discover_pde_for_patch (sindy:300-340): patch b at origin (oy, ox), frames 1 .. T-2, the masked / subsampled
// pixels in np.where order; X [B][n][11] with the reference's scrambled features (scramble = 1) or the eleven terms
// at the pixel (scramble = 0), y [B][n] = (p[i+1] - p[i-1]) / (2 dt).
__global__ void sindy_rows_kernel(const double *__restrict__ U, int64_t T, int64_t H, int64_t W, const int32_t *__restrict__ origins,
                                  int64_t B, int ps, int skip, int sub, FdConsts k, int scramble, int n_side, int first,
                                  double *__restrict__ X, double *__restrict__ y) {
    const int64_t per_frame = (int64_t)n_side * n_side, n = (T - 2) * per_frame, total = B * n;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = idx / n, q = idx % n, fi = q / per_frame + 1, m = q % per_frame;
        const int r = first + (int)(m / n_side) * sub, c = first + (int)(m % n_side) * sub;
        const double *F = U + fi * H * W + (int64_t)origins[2 * b] * W + origins[2 * b + 1];
        double *xr = X + idx * 11;
#pragma unroll 1
        for (int t = 0; t < 11; ++t) {
            const int f = c * 11 + t;
            xr[t] = scramble ? sindy_term(F, W, ps, r, f % ps, f / ps, k) : sindy_term(F, W, ps, r, c, t, k);
        }
        const int64_t o = (int64_t)r * W + c;
        y[idx] = __ddiv_rn(__dsub_rn(F[H * W + o], F[o - H * W]), __dmul_rn(2.0, k.dt));
    }
}

// build_library of basic_usage (basic:75-101): Theta [N][6] = [1, u, u_x, u_y, lap, u*u] from four flat arrays
__global__ void basic_library_rows_kernel(const double *__restrict__ u, const double *__restrict__ ux,
                                          const double *__restrict__ uy, const double *__restrict__ lap, int64_t n,
                                          double *__restrict__ Theta) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = u[i];
        double *r = Theta + i * 6;
        r[0] = 1.0; r[1] = v; r[2] = ux[i]; r[3] = uy[i]; r[4] = lap[i]; r[5] = __dmul_rn(v, v);
    }
}

// dst[i] += src[i] (statistics of sub-slabs / folds are additive)
__global__ void stats_accumulate_kernel(double *__restrict__ dst, const double *__restrict__ src, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = __dadd_rn(dst[i], src[i]);
}

// ----------------------------------------------------------------------------- host launchers
static int grid_for(int64_t work_items, int per_cta, int max_ctas) {
    int64_t g = (work_items + per_cta - 1) / per_cta;
    if (g < 1) g = 1;
    if (g > max_ctas) g = max_ctas;
    return (int)g;
}

template <int LIB> static int launch_k1_generic_t(const K1Params &P, int n_parts_cta, cudaStream_t st) {
    constexpr int p = Lib<LIB>::P;
    constexpr int S = PG_STATS_LEN(p);
    const size_t smem = sizeof(double) * (GW * P.n_folds * S + GW * 32 * (p + 2)) + 2 * S + 16;
    PG_CUDA(cudaFuncSetAttribute(k1_generic_kernel<LIB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k1_generic_kernel<LIB><<<n_parts_cta, GW * 32, smem, st>>>(P);
    PG_LAUNCHED();
    return PG_OK;
}

int launch_k1_generic(int lib, const K1Params &P, int ctas, cudaStream_t st) {
    switch (lib) {
        case PG_LIB_KS_TRUE: return launch_k1_generic_t<PG_LIB_KS_TRUE>(P, ctas, st);
        case PG_LIB_KS_TRUE_ADV: return launch_k1_generic_t<PG_LIB_KS_TRUE_ADV>(P, ctas, st);
        case PG_LIB_KS_RICH: return launch_k1_generic_t<PG_LIB_KS_RICH>(P, ctas, st);
        case PG_LIB_KS_RICH_NOADV: return launch_k1_generic_t<PG_LIB_KS_RICH_NOADV>(P, ctas, st);
        case PG_LIB_BASIC: return launch_k1_generic_t<PG_LIB_BASIC>(P, ctas, st);
        case PG_LIB_AR_FULL: return launch_k1_generic_t<PG_LIB_AR_FULL>(P, ctas, st);
        default: PG_FAIL(PG_EINVAL, "library %d cannot be used with pg_fd_lib_gram", lib);
    }
}

template <int LIB> static int launch_k1_resid_t(const K1Params &P, const double *coef, int J, int eval_fold, double *partials,
                                                int ctas, cudaStream_t st) {
    k1_generic_resid_kernel<LIB><<<ctas, GW * 32, 0, st>>>(P, coef, J, eval_fold, partials);
    PG_LAUNCHED();
    return PG_OK;
}

int launch_k1_generic_resid(int lib, const K1Params &P, const double *coef, int J, int eval_fold, double *partials, int ctas,
                            cudaStream_t st) {
    switch (lib) {
        case PG_LIB_KS_TRUE: return launch_k1_resid_t<PG_LIB_KS_TRUE>(P, coef, J, eval_fold, partials, ctas, st);
        case PG_LIB_KS_TRUE_ADV: return launch_k1_resid_t<PG_LIB_KS_TRUE_ADV>(P, coef, J, eval_fold, partials, ctas, st);
        case PG_LIB_KS_RICH: return launch_k1_resid_t<PG_LIB_KS_RICH>(P, coef, J, eval_fold, partials, ctas, st);
        case PG_LIB_KS_RICH_NOADV: return launch_k1_resid_t<PG_LIB_KS_RICH_NOADV>(P, coef, J, eval_fold, partials, ctas, st);
        case PG_LIB_BASIC: return launch_k1_resid_t<PG_LIB_BASIC>(P, coef, J, eval_fold, partials, ctas, st);
        case PG_LIB_AR_FULL: return launch_k1_resid_t<PG_LIB_AR_FULL>(P, coef, J, eval_fold, partials, ctas, st);
        default: PG_FAIL(PG_EINVAL, "library %d cannot be used with pg_fd_residual_ss", lib);
    }
}

int launch_sindy_rows(const double *U, int64_t T, int64_t H, int64_t W, const int32_t *origins, int64_t B, int ps, int skip, int sub,
                      const FdConsts &k, int scramble, int n_side, int first, double *X, double *y, cudaStream_t st) {
    const int64_t total = B * (T - 2) * n_side * n_side;
    if (total <= 0) return PG_OK;
    sindy_rows_kernel<<<grid_for(total, 128, 148 * 16), 128, 0, st>>>(U, T, H, W, origins, B, ps, skip, sub, k, scramble, n_side, first,
                                                                     X, y);
    PG_LAUNCHED();
    return PG_OK;
}

int launch_basic_library_rows(const double *u, const double *ux, const double *uy, const double *lap, int64_t n, double *Theta,
                              cudaStream_t st) {
    if (n <= 0) return PG_OK;
    basic_library_rows_kernel<<<grid_for(n, 256, 148 * 16), 256, 0, st>>>(u, ux, uy, lap, n, Theta);
    PG_LAUNCHED();
    return PG_OK;
}

int launch_stats_accumulate(double *dst, const double *src, int64_t n, cudaStream_t st) {
    if (n <= 0) return PG_OK;
    stats_accumulate_kernel<<<grid_for(n, 256, 148 * 4), 256, 0, st>>>(dst, src, n);
    PG_LAUNCHED();
    return PG_OK;
}

int launch_k1_generic_rows(int lib, const K1Params &P, double *rows_out, int ctas, cudaStream_t st) {
    switch (lib) {
#define PG_CASE(L) case L: k1_generic_rows_kernel<L><<<ctas, GW * 32, 0, st>>>(P, rows_out); break;
        PG_CASE(PG_LIB_KS_TRUE) PG_CASE(PG_LIB_KS_TRUE_ADV) PG_CASE(PG_LIB_KS_RICH) PG_CASE(PG_LIB_KS_RICH_NOADV)
        PG_CASE(PG_LIB_BASIC) PG_CASE(PG_LIB_AR_FULL)
#undef PG_CASE
        default: PG_FAIL(PG_EINVAL, "library %d cannot be used with pg_fd_block_rows", lib);
    }
    PG_LAUNCHED();
    return PG_OK;
}

int launch_rows_resid(const double *X, const double *y, int64_t n, int p, int64_t ldx, const uint8_t *fold_of_row,
                      int eval_fold, const double *coef, int J, double *partials, int ctas, cudaStream_t st) {
    rows_resid_kernel<<<ctas, GW * 32, 0, st>>>(X, y, n, p, ldx, fold_of_row, eval_fold, coef, J, partials);
    PG_LAUNCHED();
    return PG_OK;
}

int launch_reduce_partials(const double *partials, int64_t n_parts, int64_t len, double *out, int accumulate,
                           cudaStream_t st, const unsigned long long *flag, int64_t n_a, int64_t n_b,
                           const unsigned long long *poison) {
    if (len <= 0) return PG_OK;
    reduce_partials_kernel<<<(unsigned)len, 128, 0, st>>>(partials, n_parts, len, out, accumulate, flag, n_a, n_b, poison);
    PG_LAUNCHED();
    return PG_OK;
}

int launch_fd_terms(int dialect, int lib, const double *U, int64_t T, int64_t A0, int64_t A1, const FdConsts &c,
                    double *out, cudaStream_t st) {
    if (dialect == PG_FD_BASIC_TRIM) {
        const int64_t total = (T - 1) * (A0 - 4) * (A1 - 4);
        if (total <= 0) return PG_OK;
        fd_terms_basic_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, st>>>(U, T, A0, A1, c, out);
    } else {
        const int64_t total = T * A0 * A1;
        if (total <= 0) return PG_OK;
        const int g = grid_for(total, 256, 148 * 16);
        switch (lib) {
#define PG_CASE(L) case L: fd_terms_ks_kernel<L><<<g, 256, 0, st>>>(U, T, A0, A1, c, out); break;
            PG_CASE(PG_LIB_KS_TRUE) PG_CASE(PG_LIB_KS_TRUE_ADV) PG_CASE(PG_LIB_KS_RICH) PG_CASE(PG_LIB_KS_RICH_NOADV)
            PG_CASE(PG_LIB_KS_GRAD) PG_CASE(PG_LIB_KS_LAP)
#undef PG_CASE
            default: PG_FAIL(PG_EINVAL, "library %d is not a KS-dialect library", lib);
        }
    }
    PG_LAUNCHED();
    return PG_OK;
}

int launch_fd_gather(int lib, const K1Params &P, const int64_t *flat_idx, int64_t n, double *X, double *y,
                     cudaStream_t st) {
    if (n <= 0) return PG_OK;
    const int g = grid_for(n, 128, 148 * 8);
    switch (lib) {
#define PG_CASE(L) case L: fd_gather_kernel<L><<<g, 128, 0, st>>>(P, flat_idx, n, X, y); break;
        PG_CASE(PG_LIB_KS_TRUE) PG_CASE(PG_LIB_KS_TRUE_ADV) PG_CASE(PG_LIB_KS_RICH) PG_CASE(PG_LIB_KS_RICH_NOADV)
        PG_CASE(PG_LIB_BASIC) PG_CASE(PG_LIB_AR_FULL)
#undef PG_CASE
        default: PG_FAIL(PG_EINVAL, "library %d cannot be used with pg_fd_gather_rows", lib);
    }
    PG_LAUNCHED();
    return PG_OK;
}

int launch_block_means(const double *stack, int k, int64_t T, int64_t A0, int64_t A1, int bt, int b0, int b1,
                       double *out, cudaStream_t st) {
    const int64_t nBt = (T + bt - 1) / bt, nB0 = (A0 + b0 - 1) / b0, nB1 = (A1 + b1 - 1) / b1;
    const int64_t total = nBt * nB0 * nB1 * k;
    if (total <= 0) return PG_OK;
    block_means_kernel<<<grid_for(total, 128, 148 * 16), 128, 0, st>>>(stack, k, T, A0, A1, bt, b0, b1, nBt, nB0, nB1,
                                                                      out);
    PG_LAUNCHED();
    return PG_OK;
}

size_t rows_gram_smem(int p, int n_folds) {
    const int S = PG_STATS_LEN(p);
    return sizeof(double) * (GW * n_folds * S + GW * 32 * (p + 2) + GW * n_folds * 2 * p) + 2 * S + 16;
}

// Many small problems (the per-patch fits: thousands of problems of ~100 rows): one WARP per problem, no shared
// memory, no partials.  Lane j < p + 2 loads entry j of the extended row [1, y, x_0 - shift_0, ..]; every
// statistics entry is a product of two of them, fetched by shuffle, with lane e owning entries e, e + 32, ..
// (rows in order, one accumulator per entry: deterministic).
constexpr int RS_NE = (PG_STATS_LEN(PG_MAX_P) + 31) / 32;

__global__ void __launch_bounds__(256) rows_gram_small_kernel(RowsParams P, double *__restrict__ stats,
                                                             double *__restrict__ colminmax) {
    const int p = P.p, S = PG_STATS_LEN(p);
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= P.B) return;
    int ea[RS_NE], eb[RS_NE];
#pragma unroll
    for (int k = 0; k < RS_NE; ++k) {
        ea[k] = eb[k] = 0;
        if (lane + 32 * k < S) stats_pair(lane + 32 * k, p, ea[k], eb[k]);
    }
    const double *Xb = P.X + b * P.n * P.ldx;
    const double *yb = P.y + b * P.n;
    const bool col = lane >= 2 && lane < p + 2;
    const double sh = (P.shift && col) ? P.shift[b * p + lane - 2] : 0.0;
    double acc[RS_NE];
#pragma unroll
    for (int k = 0; k < RS_NE; ++k) acc[k] = 0.0;
    double vmin = INFINITY, vmax = -INFINITY;
    constexpr int UN = 4;    // rows in flight per warp
    for (int64_t r0 = 0; r0 < P.n; r0 += UN) {
        double raw[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int64_t r = r0 + u;
            raw[u] = 1.0;
            if (r < P.n) {
                if (lane == 1) raw[u] = yb[r];
                else if (col) raw[u] = Xb[r * P.ldx + lane - 2];
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            if (r0 + u >= P.n) break;
            if (col) { vmin = fmin(vmin, raw[u]); vmax = fmax(vmax, raw[u]); }
            const double ext = col ? __dsub_rn(raw[u], sh) : raw[u];
#pragma unroll
            for (int k = 0; k < RS_NE; ++k) {
                if (32 * k >= S) break;
                acc[k] = fma(__shfl_sync(0xffffffffu, ext, ea[k]), __shfl_sync(0xffffffffu, ext, eb[k]), acc[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < RS_NE; ++k)
        if (lane + 32 * k < S) stats[b * S + lane + 32 * k] = acc[k];
    if (colminmax && col) {
        colminmax[(b * 2 + 0) * p + lane - 2] = vmin;
        colminmax[(b * 2 + 1) * p + lane - 2] = vmax;
    }
}

int launch_rows_gram(const RowsParams &P, double *stats, double *colminmax, cudaStream_t st) {
    const int S = PG_STATS_LEN(P.p);
    if (P.B >= 256 && P.n <= 4096 && P.n_folds == 1 && !P.fold_of_row && !P.weights) {
        rows_gram_small_kernel<<<(unsigned)((P.B + 7) / 8), 256, 0, st>>>(P, stats, colminmax);
        PG_LAUNCHED();
        return PG_OK;
    }
    const size_t smem = rows_gram_smem(P.p, P.n_folds);
    PG_CUDA(cudaFuncSetAttribute(rows_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rows_gram_kernel<<<(unsigned)(P.chunks * P.B), GW * 32, smem, st>>>(P);
    PG_LAUNCHED();
    const int len = P.n_folds * S, mmlen = P.n_folds * 2 * P.p;
    const int64_t work = P.B * (int64_t)(len > mmlen ? len : mmlen);
    rows_reduce_kernel<<<(unsigned)((work + 127) / 128), 128, 0, st>>>(P.partials, P.mm_partials, P.B, P.chunks * GW,
                                                                       len, mmlen, P.p, stats, colminmax, P.counters);
    PG_LAUNCHED();
    return PG_OK;
}

}  // namespace pg
