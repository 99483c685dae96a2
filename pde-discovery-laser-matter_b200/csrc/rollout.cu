// Rollout check of a discovered KS-dialect PDE (ks2d:1804-1838): explicit Euler with the fitted right-hand
// side, u_hat <- u_hat + DT * sum_k c_k theta_k(u_hat), started from frame 0, and the RMSE of u_hat against
// the observed frame after every step.  The step after the solve; same stencils as K1 (reference arithmetic,
// bit-identical to the NumPy evaluation: products and sums are not contracted, terms are added in library
// order, coefficients with |c| < 1e-12 are skipped as the reference does).
//
// One launch per Euler step (a step needs the whole previous frame: a grid-wide dependency); a frame fits L2
// (33 MB at 2048^2), so the 13-point stencil reads come from L2.  Squared-error partial sums go to
// [blocks][steps] and are reduced once, in a fixed order, at the end.
#include <math.h>

#include "common.cuh"
#include "launch.h"

namespace pg {

constexpr int RO_THREADS = 256;

template <int LIB>
__global__ void __launch_bounds__(RO_THREADS) rollout_step_kernel(const double *__restrict__ u_in, double *__restrict__ u_out,
                                                                  const double *__restrict__ ref, int64_t A0, int64_t A1,
                                                                  FdConsts c, const double *__restrict__ coef, int step,
                                                                  int n_steps, double *__restrict__ partials) {
    constexpr int p = Lib<LIB>::P;
    __shared__ double sh[RO_THREADS];
    double cf[p];
#pragma unroll
    for (int k = 0; k < p; ++k) cf[k] = coef[k];
    const int64_t total = A0 * A1;
    double se = 0.0;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        PointVals v;
        ks_point<Lib<LIB>::BIH>(u_in, A0, A1, idx / A1, idx % A1, c, v);
        double row[p];
        lib_row<LIB>(v, row);
        double rhs = 0.0;   // out = zeros; out += c * v in library order (ks2d:1824-1830)
#pragma unroll
        for (int k = 0; k < p; ++k)
            if (!(fabs(cf[k]) < 1e-12)) rhs = __dadd_rn(rhs, __dmul_rn(cf[k], row[k]));
        const double un = __dadd_rn(v.u, __dmul_rn(c.dt, rhs));
        u_out[idx] = un;
        const double d = __dsub_rn(ref[idx], un);
        se = __dadd_rn(se, __dmul_rn(d, d));
    }
    sh[threadIdx.x] = se;
    __syncthreads();
    for (int w = RO_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] = __dadd_rn(sh[threadIdx.x], sh[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * n_steps + step] = sh[0];
}

__global__ void rollout_finish_kernel(double *__restrict__ rmse, int n_steps, double n_points) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_steps) rmse[k] = sqrt(rmse[k] / n_points);
}

template <int LIB>
static int rollout_t(const double *U, int64_t A0, int64_t A1, const FdConsts &c, const double *coef, int n_steps,
                     double *work, double *partials, int blocks, double *rmse, cudaStream_t st) {
    const int64_t frame = A0 * A1;
    for (int k = 0; k < n_steps; ++k) {
        const double *in = k == 0 ? U : work + (int64_t)((k - 1) & 1) * frame;
        double *out = work + (int64_t)(k & 1) * frame;
        rollout_step_kernel<LIB><<<blocks, RO_THREADS, 0, st>>>(in, out, U + (int64_t)(k + 1) * frame, A0, A1, c, coef, k,
                                                                n_steps, partials);
        PG_LAUNCHED();
    }
    int rc = launch_reduce_partials(partials, blocks, n_steps, rmse, 0, st);
    if (rc) return rc;
    rollout_finish_kernel<<<(n_steps + 127) / 128, 128, 0, st>>>(rmse, n_steps, (double)frame);
    PG_LAUNCHED();
    return PG_OK;
}

// ----------------------------------------------------------------------------- analyze_results rollout (ar:300-395)
// rollout_k_rmse: from EVERY start frame t of a time slice, k explicit-Euler steps u <- u + dt * f(u) with the model's
// terms evaluated by derivs_2d (same-grid central differences through np.pad(mode="reflect"), NOT the fit's
// slice-aligned stencils), then the error against frame t + k.  The start frames are independent: one launch per
// Euler step advances all of them ([n_start][H][W] state, ping-pong), the last step also accumulates
// sum e^2, sum y, sum y^2 and the count over the (optionally masked) targets.
// Terms are added in the caller's order, |c| < 1e-12 skipped, products not contracted (ar:316-345).
__device__ __forceinline__ int64_t reflect1(int64_t i, int64_t n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

__global__ void __launch_bounds__(RO_THREADS) ar_rollout_step_kernel(const double *__restrict__ in, int64_t in_stride,
                                                                     double *__restrict__ out, const double *__restrict__ target,
                                                                     int64_t n_start, int64_t H, int64_t W, FdConsts c,
                                                                     const int32_t *__restrict__ term_ids,
                                                                     const double *__restrict__ coef, int n_terms, int last,
                                                                     const uint8_t *__restrict__ mask, double *__restrict__ partials) {
    __shared__ double sh[RO_THREADS];
    const int64_t frame = H * W, total = n_start * frame;
    double a[4] = {0, 0, 0, 0};
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = idx / frame, r = idx % frame, i = r / W, j = r % W;
        const double *F = in + s * in_stride;
        const double uc = F[i * W + j];
        const double ue = F[i * W + reflect1(j + 1, W)], uw = F[i * W + reflect1(j - 1, W)];
        const double un = F[reflect1(i + 1, H) * W + j], us = F[reflect1(i - 1, H) * W + j];
        const double ux = central_diff(ue, uw, c.two_d1), uy = central_diff(un, us, c.two_d0);
        const double uxx = second_diff(ue, uc, uw, c.d1sq), uyy = second_diff(un, uc, us, c.d0sq);
        double rhs = 0.0;
        for (int k = 0; k < n_terms; ++k) {
            const double cf = coef[k];
            if (fabs(cf) < 1e-12) continue;
            double v;
            switch (term_ids[k]) {
                case 0: v = 1.0; break;
                case 1: v = uc; break;
                case 2: v = ux; break;
                case 3: v = uy; break;
                case 4: v = uxx; break;
                case 5: v = uyy; break;
                case 6: v = __dadd_rn(uxx, uyy); break;
                case 7: v = __dmul_rn(uc, uc); break;
                case 8: v = __dmul_rn(uc, ux); break;
                case 9: v = __dmul_rn(uc, uy); break;
                case 10: v = __dmul_rn(__dmul_rn(uc, uc), uc); break;
                case 11: v = __dmul_rn(ux, ux); break;
                default: v = __dmul_rn(uy, uy); break;
            }
            rhs = __dadd_rn(rhs, __dmul_rn(cf, v));
        }
        const double unew = __dadd_rn(uc, __dmul_rn(c.dt, rhs));
        out[idx] = unew;
        if (last && (!mask || mask[r])) {
            const double y = target[s * frame + r], d = __dsub_rn(y, unew);
            a[0] = fma(d, d, a[0]); a[1] += y; a[2] = fma(y, y, a[2]); a[3] += 1.0;
        }
    }
    if (!last) return;
    for (int q = 0; q < 4; ++q) {
        sh[threadIdx.x] = a[q];
        __syncthreads();
        for (int w = RO_THREADS / 2; w > 0; w >>= 1) {
            if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * 4 + q] = sh[0];
        __syncthreads();
    }
}

int launch_ar_rollout(const double *U, int64_t H, int64_t W, const FdConsts &c, const int32_t *term_ids, const double *coef,
                      int n_terms, int k_steps, int64_t t0, int64_t n_start, const uint8_t *mask, double *work, double *partials,
                      int blocks, double *out4, cudaStream_t st) {
    const int64_t frame = H * W, vol = n_start * frame;
    for (int k = 0; k < k_steps; ++k) {
        const double *in = k == 0 ? U + t0 * frame : work + (int64_t)((k - 1) & 1) * vol;
        double *out = work + (int64_t)(k & 1) * vol;
        ar_rollout_step_kernel<<<blocks, RO_THREADS, 0, st>>>(in, frame, out, U + (t0 + k_steps) * frame, n_start, H, W, c, term_ids,
                                                              coef, n_terms, k == k_steps - 1, mask, partials);
        PG_LAUNCHED();
    }
    return launch_reduce_partials(partials, blocks, 4, out4, 0, st);
}

// one_step_prediction_rmse (ar:150-187): sum over t < t_max of (u[t+1] - (u[t] + dt * ut_pred[t]))^2 (+ count)
__global__ void __launch_bounds__(RO_THREADS) one_step_kernel(const double *__restrict__ u, const double *__restrict__ ut,
                                                              int64_t t_max, int64_t frame, double dt,
                                                              const uint8_t *__restrict__ mask, double *__restrict__ partials) {
    __shared__ double sh[RO_THREADS];
    double a[2] = {0, 0};
    const int64_t total = t_max * frame;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        if (mask && !mask[idx % frame]) continue;
        const double pred = __dadd_rn(u[idx], __dmul_rn(dt, ut[idx]));
        const double d = __dsub_rn(u[idx + frame], pred);
        a[0] = fma(d, d, a[0]); a[1] += 1.0;
    }
    for (int q = 0; q < 2; ++q) {
        sh[threadIdx.x] = a[q];
        __syncthreads();
        for (int w = RO_THREADS / 2; w > 0; w >>= 1) {
            if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * 2 + q] = sh[0];
        __syncthreads();
    }
}

int launch_one_step(const double *u, const double *ut, int64_t t_max, int64_t frame, double dt, const uint8_t *mask,
                    double *partials, int blocks, double *out2, cudaStream_t st) {
    one_step_kernel<<<blocks, RO_THREADS, 0, st>>>(u, ut, t_max, frame, dt, mask, partials);
    PG_LAUNCHED();
    return launch_reduce_partials(partials, blocks, 2, out2, 0, st);
}

// ----------------------------------------------------------------------------- fit metrics (ks2d:29-40, patch:47-65)
// Two passes over (y_true, y_pred): raw sums, then sums centred on the means of the first pass (what np.std,
// np.corrcoef and the reference's r2_score do).  Per-block partial sums, fixed-order reduction.
//   pass 0 entries: sum r, sum r^2, sum |r|, sum y, sum yhat          (r = y - yhat)
//   pass 1 entries: sum (y-my)^2, sum (yhat-mh)^2, sum (y-my)(yhat-mh), sum (r-mr)^2
constexpr int FM_THREADS = 256;

__global__ void __launch_bounds__(FM_THREADS) fit_metrics_kernel(const double *__restrict__ y, const double *__restrict__ yh,
                                                                 int64_t n, int pass, const double *__restrict__ sums0,
                                                                 double *__restrict__ partials) {
    __shared__ double sh[FM_THREADS];
    constexpr int NQ = 5;
    double a[NQ] = {0, 0, 0, 0, 0};
    double my = 0, mh = 0, mr = 0;
    if (pass == 1) { my = sums0[3] / (double)n; mh = sums0[4] / (double)n; mr = sums0[0] / (double)n; }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double yt = y[i], yp = yh[i], r = __dsub_rn(yt, yp);
        if (pass == 0) {
            a[0] += r; a[1] = fma(r, r, a[1]); a[2] += fabs(r); a[3] += yt; a[4] += yp;
        } else {
            const double dy = yt - my, dh = yp - mh, dr = r - mr;
            a[0] = fma(dy, dy, a[0]); a[1] = fma(dh, dh, a[1]); a[2] = fma(dy, dh, a[2]); a[3] = fma(dr, dr, a[3]);
        }
    }
    for (int q = 0; q < NQ; ++q) {
        sh[threadIdx.x] = a[q];
        __syncthreads();
        for (int w = FM_THREADS / 2; w > 0; w >>= 1) {
            if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * NQ + q] = sh[0];
        __syncthreads();
    }
}

// Batched form for the per-patch fits (patch:425-429: regression_metrics of X @ c on every patch's train and test
// rows): one warp per problem, the same two passes; sums_out [B][10] as pg_fit_metrics, resid_out [B][n] optional
// (|resid| medians are taken by the caller).
__global__ void __launch_bounds__(128) rows_metrics_batched_kernel(const double *__restrict__ X, const double *__restrict__ y,
                                                                   const double *__restrict__ coef, int64_t B, int64_t n, int p,
                                                                   int64_t ldx, double *__restrict__ sums, double *__restrict__ resid) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const double *Xb = X + b * n * ldx, *yb = y + b * n, *cb = coef + b * p;
    double a[5] = {0, 0, 0, 0, 0};
    for (int64_t r = lane; r < n; r += 32) {
        double pred = 0.0;     // X @ c: a dot product per row, in column order
        for (int k = 0; k < p; ++k) pred = fma(Xb[r * ldx + k], cb[k], pred);
        const double yt = yb[r], d = __dsub_rn(yt, pred);
        if (resid) resid[b * n + r] = d;
        a[0] += d; a[1] = fma(d, d, a[1]); a[2] += fabs(d); a[3] += yt; a[4] += pred;
    }
#pragma unroll
    for (int q = 0; q < 5; ++q)
#pragma unroll
        for (int o = 16; o; o >>= 1) a[q] += __shfl_xor_sync(0xffffffffu, a[q], o);
    const double nn = (double)n, my = a[3] / nn, mh = a[4] / nn, mr = a[0] / nn;
    double c2[4] = {0, 0, 0, 0};
    for (int64_t r = lane; r < n; r += 32) {
        double pred = 0.0;
        for (int k = 0; k < p; ++k) pred = fma(Xb[r * ldx + k], cb[k], pred);
        const double yt = yb[r], d = __dsub_rn(yt, pred);
        const double dy = yt - my, dh = pred - mh, dr = d - mr;
        c2[0] = fma(dy, dy, c2[0]); c2[1] = fma(dh, dh, c2[1]); c2[2] = fma(dy, dh, c2[2]); c2[3] = fma(dr, dr, c2[3]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int o = 16; o; o >>= 1) c2[q] += __shfl_xor_sync(0xffffffffu, c2[q], o);
    if (lane == 0) {
        double *s = sums + b * 10;
        for (int q = 0; q < 5; ++q) s[q] = a[q];
        for (int q = 0; q < 4; ++q) s[5 + q] = c2[q];
        s[9] = 0.0;
    }
}

int launch_rows_metrics_batched(const double *X, const double *y, const double *coef, int64_t B, int64_t n, int p, int64_t ldx,
                                double *sums_out, double *resid_out, cudaStream_t st) {
    if (B <= 0) return PG_OK;
    rows_metrics_batched_kernel<<<(unsigned)((B + 3) / 4), 128, 0, st>>>(X, y, coef, B, n, p, ldx, sums_out, resid_out);
    PG_LAUNCHED();
    return PG_OK;
}

int launch_fit_metrics(const double *y, const double *yh, int64_t n, double *partials, int blocks, double *out10, cudaStream_t st) {
    fit_metrics_kernel<<<blocks, FM_THREADS, 0, st>>>(y, yh, n, 0, nullptr, partials);
    PG_LAUNCHED();
    int rc = launch_reduce_partials(partials, blocks, 5, out10, 0, st);
    if (rc) return rc;
    fit_metrics_kernel<<<blocks, FM_THREADS, 0, st>>>(y, yh, n, 1, out10, partials);
    PG_LAUNCHED();
    return launch_reduce_partials(partials, blocks, 5, out10 + 5, 0, st);
}

int rollout_blocks(int64_t A0, int64_t A1, int n_sm) {
    int64_t g = (A0 * A1 + RO_THREADS - 1) / RO_THREADS;
    const int64_t cap = (int64_t)n_sm * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int launch_rollout(int lib, const double *U, int64_t A0, int64_t A1, const FdConsts &c, const double *coef, int n_steps,
                   double *work, double *partials, int blocks, double *rmse, cudaStream_t st) {
    switch (lib) {
        case PG_LIB_KS_TRUE: return rollout_t<PG_LIB_KS_TRUE>(U, A0, A1, c, coef, n_steps, work, partials, blocks, rmse, st);
        case PG_LIB_KS_TRUE_ADV: return rollout_t<PG_LIB_KS_TRUE_ADV>(U, A0, A1, c, coef, n_steps, work, partials, blocks, rmse, st);
        case PG_LIB_KS_RICH: return rollout_t<PG_LIB_KS_RICH>(U, A0, A1, c, coef, n_steps, work, partials, blocks, rmse, st);
        case PG_LIB_KS_RICH_NOADV: return rollout_t<PG_LIB_KS_RICH_NOADV>(U, A0, A1, c, coef, n_steps, work, partials, blocks, rmse, st);
        default: PG_FAIL(PG_EINVAL, "library %d does not belong to the KS dialect", lib);
    }
}

}  // namespace pg
