// K3 -- batched STRidge on sufficient statistics.  One warp per (problem, alpha, threshold):
// column normalisation from the Gram diagonal, a p x p ridge solve (LU with partial pivoting as
// np.linalg.solve does for ks2d:60 / basic:125,139; Cholesky as scikit-learn's Ridge does for
// patch:83,92), sequential thresholding, and optional held-out r2/rmse + sweep arg-max.
#include <float.h>
#include <math.h>

#include "common.cuh"
#include "launch.h"

namespace pg {

constexpr int SW = 4;        // warps per CTA
constexpr int LD = PG_MAX_P + 1;  // leading dimension of the per-warp matrices (odd -> no bank conflicts)

struct WarpMem {
    double C[PG_MAX_P * LD];   // standardised normal matrix (without alpha)
    double M[PG_MAX_P * (LD + 1)];  // working system [k][k+1], ld = LD+1
    double r[PG_MAX_P], s[PG_MAX_P], d[PG_MAX_P], c[PG_MAX_P], x[PG_MAX_P];
    int idx[PG_MAX_P];
};
constexpr int LM = LD + 1;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Solve (C[act,act] + alpha I) x = r[act]; writes c (zeros outside act).  All 32 lanes call.
__device__ void solve_active(WarpMem &w, int p, unsigned act, double alpha, bool cholesky, int lane) {
    // compact the active indices (ascending, as boolean indexing does)
    int k = 0;
    for (int j = 0; j < p; ++j)
        if ((act >> j) & 1u) { if (lane == 0) w.idx[k] = j; ++k; }
    __syncwarp();
    for (int e = lane; e < k * (k + 1); e += 32) {
        const int a = e / (k + 1), b = e % (k + 1);
        w.M[a * LM + b] = (b == k) ? w.r[w.idx[a]] : (w.C[w.idx[a] * LD + w.idx[b]] + (a == b ? alpha : 0.0));
    }
    __syncwarp();
    if (!cholesky) {
        for (int c = 0; c < k; ++c) {
            // partial pivoting: first row of maximal |M[a][c]|, a >= c
            double best = (lane >= c && lane < k) ? fabs(w.M[lane * LM + c]) : -1.0;
            if (best != best) best = INFINITY;  // NaN propagates through the pivot like LAPACK's idamax would not; keep going
            int bi = lane;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (bi != c && lane <= k) {
                const double t = w.M[c * LM + lane];
                w.M[c * LM + lane] = w.M[bi * LM + lane];
                w.M[bi * LM + lane] = t;
            }
            __syncwarp();
            if (lane > c && lane < k) {
                const double f = w.M[lane * LM + c] / w.M[c * LM + c];
                for (int b = c + 1; b <= k; ++b) w.M[lane * LM + b] = fma(-f, w.M[c * LM + b], w.M[lane * LM + b]);
            }
            __syncwarp();
        }
        for (int a = k - 1; a >= 0; --a) {
            if (lane == a) w.x[a] = w.M[a * LM + k] / w.M[a * LM + a];
            __syncwarp();
            if (lane < a) w.M[lane * LM + k] = fma(-w.M[lane * LM + a], w.x[a], w.M[lane * LM + k]);
            __syncwarp();
        }
    } else {
        // lower Cholesky in place (columns left to right), then L z = r, L^T x = z
        for (int c = 0; c < k; ++c) {
            if (lane == c) {
                const double dd = w.M[c * LM + c];
                w.M[c * LM + c] = dd > 0.0 ? sqrt(dd) : nan("");
            }
            __syncwarp();
            if (lane > c && lane < k) w.M[lane * LM + c] /= w.M[c * LM + c];
            __syncwarp();
            if (lane > c && lane < k)
                for (int b = c + 1; b <= lane; ++b) w.M[lane * LM + b] = fma(-w.M[lane * LM + c], w.M[b * LM + c], w.M[lane * LM + b]);
            __syncwarp();
        }
        for (int a = 0; a < k; ++a) {  // forward
            if (lane == a) w.x[a] = w.M[a * LM + k] / w.M[a * LM + a];
            __syncwarp();
            if (lane > a && lane < k) w.M[lane * LM + k] = fma(-w.M[lane * LM + a], w.x[a], w.M[lane * LM + k]);
            __syncwarp();
        }
        if (lane < k) w.M[lane * LM + k] = w.x[lane];
        __syncwarp();
        for (int a = k - 1; a >= 0; --a) {  // backward with L^T
            if (lane == a) w.x[a] = w.M[a * LM + k] / w.M[a * LM + a];
            __syncwarp();
            if (lane < a) w.M[lane * LM + k] = fma(-w.M[a * LM + lane], w.x[a], w.M[lane * LM + k]);
            __syncwarp();
        }
    }
    if (lane < p) w.c[lane] = 0.0;
    __syncwarp();
    if (lane < k) w.c[w.idx[lane]] = w.x[lane];
    __syncwarp();
}

// stridge_sign_constrained (ks2d:577-582, 594-598): a coefficient whose sign contradicts signs[j] becomes 0
__device__ __forceinline__ void enforce_signs(WarpMem &w, int p, const int8_t *signs, int lane) {
    if (!signs) return;
    if (lane < p) {
        const int sg = signs[lane];
        const double c = w.c[lane];
        if ((sg == -1 && c > 0.0) || (sg == 1 && c < 0.0)) w.c[lane] = 0.0;
    }
    __syncwarp();
}

__device__ __forceinline__ unsigned big_mask(const WarpMem &w, int p, double thr, int lane) {
    const bool big = lane < p && !(fabs(w.c[lane]) < thr);
    return __ballot_sync(0xffffffffu, big);
}

__global__ void __launch_bounds__(SW * 32) stridge_kernel(StridgeParams P) {
    __shared__ WarpMem wm[SW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t job = (int64_t)blockIdx.x * SW + warp;
    const int64_t njobs = P.B * P.na * P.nt;
    if (job >= njobs) return;
    WarpMem &w = wm[warp];
    const int p = P.p, S = PG_STATS_LEN(p);
    const int it_ = (int)(job % P.nt), ia = (int)((job / P.nt) % P.na);
    const int64_t b = job / ((int64_t)P.nt * P.na);
    const double alpha = P.alphas[ia], thr = P.thrs[it_];
    const double *st = P.stats + b * S;
    const double n = st[0], sy = st[1];
    const double *sx = st + 3, *bb = st + 3 + p, *Gu = st + 3 + 2 * p;
    const double *h = P.shift ? P.shift + b * p : nullptr;

    // full symmetric Gram into C
    for (int e = lane; e < p * p; e += 32) {
        int i = e / p, j = e % p;
        if (i > j) { const int t = i; i = j; j = t; }
        w.C[(e / p) * LD + (e % p)] = Gu[i * p - (i * (i - 1)) / 2 + (j - i)];
    }
    __syncwarp();
    const bool is_const_arg = lane < p && P.const_mask && P.const_mask[lane];
    // ---- optional train-RMS pre-scale (ks2d:1647-1655)
    double dj = 1.0, hj = (h && lane < p) ? h[lane] : 0.0;
    double sxj = lane < p ? sx[lane] : 0.0, bj = lane < p ? bb[lane] : 0.0;
    if ((P.flags & PG_STRIDGE_RMS_PRESCALE) && lane < p) {
        const double raw_gjj = w.C[lane * LD + lane] + 2.0 * hj * sxj + n * hj * hj;
        dj = is_const_arg ? 1.0 : sqrt(raw_gjj / n) + 1e-12;
    }
    if (lane < p) w.d[lane] = dj;
    __syncwarp();
    if (P.flags & PG_STRIDGE_RMS_PRESCALE) {
        for (int e = lane; e < p * p; e += 32) w.C[(e / p) * LD + (e % p)] /= (w.d[e / p] * w.d[e % p]);
        sxj /= dj; bj /= dj; hj /= dj;
        __syncwarp();
    }
    const bool cholesky = P.dialect == PG_STRIDGE_SKLEARN;
    if (P.dialect == PG_STRIDGE_BASIC) {
        if (lane < p) { w.r[lane] = bj; w.s[lane] = 1.0; }
        __syncwarp();
    } else {
        // ---- centre + scale (ks2d:43-52 / StandardScaler): C <- (G - n mu mu^T) / (s s^T), r <- (b - mu*sy)/s
        const double muj = sxj / n;
        if (lane < p) w.x[lane] = muj;
        __syncwarp();
        for (int e = lane; e < p * p; e += 32) {
            const int i = e / p, j = e % p;
            w.C[i * LD + j] = fma(-n * w.x[i], w.x[j], w.C[i * LD + j]);
        }
        __syncwarp();
        bool cst = is_const_arg;
        double sj = 1.0;
        if (lane < p) {
            const double var = w.C[lane * LD + lane] / n;
            if (!(var > 0.0)) cst = true;
            if (P.colminmax && P.colminmax[(b * 2 + 0) * p + lane] == P.colminmax[(b * 2 + 1) * p + lane]) cst = true;
            if (P.dialect == PG_STRIDGE_SKLEARN) {
                const double mean = muj + hj, t = n * mean * DBL_EPSILON;
                if (var <= n * DBL_EPSILON * var + t * t) cst = true;
                if (!cst && sqrt(var) < 10.0 * DBL_EPSILON) cst = true;
            }
            sj = cst ? 1.0 : sqrt(var);
            w.s[lane] = sj;
            w.r[lane] = cst ? 0.0 : (bj - muj * sy) / sj;
        }
        const unsigned cmask = __ballot_sync(0xffffffffu, cst && lane < p);
        __syncwarp();
        for (int e = lane; e < p * p; e += 32) {
            const int i = e / p, j = e % p;
            const bool z = ((cmask >> i) | (cmask >> j)) & 1u;
            w.C[i * LD + j] = z ? 0.0 : w.C[i * LD + j] / (w.s[i] * w.s[j]);
        }
        __syncwarp();
    }

    const unsigned all = p >= 32 ? 0xffffffffu : ((1u << p) - 1u);
    if (P.dialect == PG_STRIDGE_BASIC) {
        if (P.max_iter <= 0) {
            if (lane < p) w.c[lane] = 1.0;  // basic:119 (loop never runs)
            __syncwarp();
        } else {
            solve_active(w, p, all, alpha, false, lane);
            const unsigned act = big_mask(w, p, thr, lane);
            if (act == 0) {
                if (lane < p) w.c[lane] = 0.0;
                __syncwarp();
            } else if (act != all) {
                solve_active(w, p, act, alpha, false, lane);
            } else {
                // basic:136-141 refits on the (full) active set; identical system, identical result
            }
        }
    } else {
        solve_active(w, p, all, alpha, cholesky, lane);
        unsigned prev = all;
        for (int it = 0; it < P.max_iter; ++it) {
            enforce_signs(w, p, P.signs, lane);   // no-op after the first pass (already enforced after the refit)
            const unsigned big = big_mask(w, p, thr, lane);
            if (big == 0) {
                if (lane < p) w.c[lane] = 0.0;
                __syncwarp();
                break;
            }
            if (big == prev && it > 0) break;  // same support -> the refit reproduces w.c
            if (big == all && it == 0) { prev = big; continue; }  // refit of the full system = initial solve
            solve_active(w, p, big, alpha, cholesky, lane);
            enforce_signs(w, p, P.signs, lane);
            prev = big;
        }
    }
    // ---- unscale: c / (s + 1e-12) (ks2d:428, patch:98), then / d (ks2d:1728)
    double cj = 0.0;
    if (lane < p) {
        cj = w.c[lane];
        if (P.dialect != PG_STRIDGE_BASIC) cj = (P.flags & PG_STRIDGE_NO_EPS) ? cj / w.s[lane] : cj / (w.s[lane] + 1e-12);
        if (P.flags & PG_STRIDGE_RMS_PRESCALE) cj = cj / w.d[lane];
        P.coef_out[job * p + lane] = cj;
    }
    if (P.eval_stats && P.metrics_out) {
        // r2 / rmse of X @ c on the held-out rows from their statistics (ks2d:29-40)
        const double *ev = P.eval_stats + b * S;
        const double en = ev[0], esy = ev[1], esyy = ev[2];
        const double *esx = ev + 3, *eb = ev + 3 + p, *eG = ev + 3 + 2 * p;
        const double h_raw = (h && lane < p) ? h[lane] : 0.0;
        const double hc = warp_sum(lane < p ? h_raw * cj : 0.0);
        if (lane < p) w.x[lane] = cj;
        __syncwarp();
        double lin = 0.0, quad = 0.0;
        if (lane < p) {
            lin = cj * (eb[lane] - hc * esx[lane]);
            double rowdot = 0.0;
            for (int j = 0; j < p; ++j) {
                int i0 = lane, j0 = j;
                if (i0 > j0) { const int t = i0; i0 = j0; j0 = t; }
                rowdot = fma(eG[i0 * p - (i0 * (i0 - 1)) / 2 + (j0 - i0)], w.x[j], rowdot);
            }
            quad = cj * rowdot;
        }
        lin = warp_sum(lin);
        quad = warp_sum(quad);
        const double yy = esyy - 2.0 * hc * esy + en * hc * hc;
        double ss_res = yy - 2.0 * lin + quad;
        // Three O(yy) terms: for a fit that is exact to ~1e-8 the difference is rounding noise (|ss_res| <~ 1e-15 yy).
        // relres_out hands the caller the ratio so that it can take the residuals from the rows instead
        // (pg_rows_residual_ss / pg_fd_residual_ss); the clamp only keeps sqrt() defined.
        if (P.relres_out && lane == 0) P.relres_out[job] = ss_res / (yy > 0.0 ? yy : 1.0);
        if (ss_res < 0.0) ss_res = 0.0;
        const double ss_tot = esyy - esy * esy / en;
        if (lane == 0) {
            P.metrics_out[job * 2 + 0] = 1.0 - ss_res / (ss_tot + 1e-18);
            P.metrics_out[job * 2 + 1] = sqrt(ss_res / en);
        }
    }
}

// sweep arg-max (ks2d:1731-1741): key (r2, -n_active, -rmse), strict '>' so the first maximum wins
__global__ void stridge_best_kernel(const double *__restrict__ coef, const double *__restrict__ metrics, int64_t B, int p,
                                    int njob, int32_t *__restrict__ best) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int bi = -1, bna = 0;
    double br2 = 0.0, brm = 0.0;
    for (int k = 0; k < njob; ++k) {
        const double r2 = metrics[(b * njob + k) * 2], rm = metrics[(b * njob + k) * 2 + 1];
        int na = 0;
        for (int j = 0; j < p; ++j) na += fabs(coef[(b * njob + k) * p + j]) > 0.0;
        bool better = bi < 0;
        if (!better) {
            if (r2 != br2) better = r2 > br2;
            else if (na != bna) better = na < bna;
            else better = rm < brm;
        }
        if (better) { bi = k; br2 = r2; bna = na; brm = rm; }
    }
    best[b] = bi;
}

int launch_stridge(const StridgeParams &P, int32_t *best_out, cudaStream_t st) {
    const int64_t njobs = P.B * P.na * P.nt;
    if (njobs <= 0) return PG_OK;
    stridge_kernel<<<(unsigned)((njobs + SW - 1) / SW), SW * 32, 0, st>>>(P);
    PG_LAUNCHED();
    if (best_out && P.metrics_out) {
        stridge_best_kernel<<<(unsigned)((P.B + 127) / 128), 128, 0, st>>>(P.coef_out, P.metrics_out, P.B, P.p,
                                                                          P.na * P.nt, best_out);
        PG_LAUNCHED();
    }
    return PG_OK;
}

}  // namespace pg
