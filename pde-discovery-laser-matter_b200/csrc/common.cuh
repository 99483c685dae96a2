// Shared device helpers for libpdegram (sm_100a).  See include/pdegram.h for the ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pdegram.h"

namespace pg {

// ----------------------------------------------------------------------------- host-side error plumbing
void set_error(const char *fmt, ...);
#define PG_FAIL(code, ...)        \
    do {                          \
        pg::set_error(__VA_ARGS__); \
        return (code);            \
    } while (0)
#define PG_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) PG_FAIL(PG_ECUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

// every kernel launch of the library goes through this: counts it (pg_launch_count) and checks it
void count_launch();
#define PG_LAUNCHED()                \
    do {                             \
        pg::count_launch();          \
        PG_CUDA(cudaGetLastError()); \
    } while (0)

// ----------------------------------------------------------------------------- library traits
template <int LIB> struct Lib;
template <> struct Lib<PG_LIB_KS_TRUE>       { static constexpr int P = 3; static constexpr bool BIH = true;  };
template <> struct Lib<PG_LIB_KS_TRUE_ADV>   { static constexpr int P = 5; static constexpr bool BIH = true;  };
template <> struct Lib<PG_LIB_KS_RICH>       { static constexpr int P = 9; static constexpr bool BIH = true;  };
template <> struct Lib<PG_LIB_KS_RICH_NOADV> { static constexpr int P = 7; static constexpr bool BIH = true;  };
template <> struct Lib<PG_LIB_BASIC>         { static constexpr int P = 6; static constexpr bool BIH = false; };
template <> struct Lib<PG_LIB_KS_GRAD>       { static constexpr int P = 2; static constexpr bool BIH = false; };
template <> struct Lib<PG_LIB_KS_LAP>        { static constexpr int P = 1; static constexpr bool BIH = false; };
template <> struct Lib<PG_LIB_AR_FULL>       { static constexpr int P = 13; static constexpr bool BIH = false; };

inline int library_width(int lib) {
    switch (lib) {
        case PG_LIB_KS_TRUE: return 3;
        case PG_LIB_KS_TRUE_ADV: return 5;
        case PG_LIB_KS_RICH: return 9;
        case PG_LIB_KS_RICH_NOADV: return 7;
        case PG_LIB_BASIC: return 6;
        case PG_LIB_KS_GRAD: return 2;
        case PG_LIB_KS_LAP: return 1;
        case PG_LIB_PATCH_MODEL4: return 6;
        case PG_LIB_PATCH_FULL: return 8;
        case PG_LIB_PATCH_DERIVS: return 6;
        case PG_LIB_AR_FULL: return 13;
        default: return -1;
    }
}

// Values of one grid point that every library row is built from.  g0/g1 are the central
// differences along a0/a1.
struct PointVals {
    double u, g0, g1, lap, bih;
    double d00, d11;   // second differences along a0 / a1 (slice-central dialect only)
};

// Library row from point values.  Products use explicit _rn intrinsics so the compiler cannot
// contract them into FMAs: the materialised terms are then bit-identical to NumPy's.
template <int LIB> __device__ __forceinline__ void lib_row(const PointVals &v, double *th) {
    if constexpr (LIB == PG_LIB_KS_TRUE) {
        th[0] = v.lap; th[1] = v.bih; th[2] = __dadd_rn(__dmul_rn(v.g0, v.g0), __dmul_rn(v.g1, v.g1));
    } else if constexpr (LIB == PG_LIB_KS_TRUE_ADV) {
        th[0] = v.lap; th[1] = v.bih; th[2] = __dadd_rn(__dmul_rn(v.g0, v.g0), __dmul_rn(v.g1, v.g1));
        th[3] = v.g0; th[4] = v.g1;
    } else if constexpr (LIB == PG_LIB_KS_RICH) {
        th[0] = 1.0; th[1] = v.u; th[2] = __dmul_rn(v.u, v.u); th[3] = v.g0; th[4] = v.g1; th[5] = v.lap;
        th[6] = v.bih; th[7] = __dadd_rn(__dmul_rn(v.g0, v.g0), __dmul_rn(v.g1, v.g1));
        th[8] = __dmul_rn(v.u, v.lap);
    } else if constexpr (LIB == PG_LIB_KS_RICH_NOADV) {
        th[0] = 1.0; th[1] = v.u; th[2] = __dmul_rn(v.u, v.u); th[3] = v.lap; th[4] = v.bih;
        th[5] = __dadd_rn(__dmul_rn(v.g0, v.g0), __dmul_rn(v.g1, v.g1)); th[6] = __dmul_rn(v.u, v.lap);
    } else if constexpr (LIB == PG_LIB_BASIC) {
        // basic_usage calls the LAST axis "x": u_x = g1, u_y = g0 (basic:58-59)
        th[0] = 1.0; th[1] = v.u; th[2] = v.g1; th[3] = v.g0; th[4] = v.lap; th[5] = __dmul_rn(v.u, v.u);
    } else if constexpr (LIB == PG_LIB_AR_FULL) {
        // analyze_results:618-623; x = a1.  u**3 is np.power(u, 3) there: u*u*u differs from it by <= 1 ulp
        th[0] = 1.0; th[1] = v.u; th[2] = v.g1; th[3] = v.g0; th[4] = v.d11; th[5] = v.d00; th[6] = v.lap;
        th[7] = __dmul_rn(v.u, v.u); th[8] = __dmul_rn(v.u, v.g1); th[9] = __dmul_rn(v.u, v.g0);
        th[10] = __dmul_rn(__dmul_rn(v.u, v.u), v.u); th[11] = __dmul_rn(v.g1, v.g1); th[12] = __dmul_rn(v.g0, v.g0);
    } else if constexpr (LIB == PG_LIB_KS_GRAD) {
        th[0] = v.g0; th[1] = v.g1;
    } else if constexpr (LIB == PG_LIB_KS_LAP) {
        th[0] = v.lap;
    }
}

// Spacing-derived constants, computed once on the host exactly as the reference does:
// 2*dx, dx**2 (ks2d:65-66,71-72; basic:58-63).
struct FdConsts {
    double two_d0, two_d1, d0sq, d1sq, dt;
};
inline FdConsts make_consts(double d0, double d1, double dt) { return {2.0 * d0, 2.0 * d1, d0 * d0, d1 * d1, dt}; }

// (plus - 2c + minus) / h2 in the reference's evaluation order, no contraction.
__device__ __forceinline__ double second_diff(double plus, double c, double minus, double h2) {
    return __ddiv_rn(__dadd_rn(__dsub_rn(plus, __dmul_rn(2.0, c)), minus), h2);
}
__device__ __forceinline__ double central_diff(double plus, double minus, double two_h) {
    return __ddiv_rn(__dsub_rn(plus, minus), two_h);
}

__device__ __forceinline__ int64_t wrap(int64_t i, int64_t n) {
    i %= n;
    return i < 0 ? i + n : i;
}

// Reference-arithmetic evaluation of one point of frame F ([A0][A1]) with periodic wrap
// (ks2d:63-73; bih = laplacian(laplacian(u)), ks2d:1039-1040).
template <bool BIH>
__device__ __forceinline__ void ks_point(const double *__restrict__ F, int64_t A0, int64_t A1, int64_t i, int64_t j,
                                         const FdConsts &c, PointVals &v) {
    const int64_t im1 = wrap(i - 1, A0), ip1 = wrap(i + 1, A0), jm1 = wrap(j - 1, A1), jp1 = wrap(j + 1, A1);
    const double uc = F[i * A1 + j];
    const double un = F[ip1 * A1 + j], us = F[im1 * A1 + j], ue = F[i * A1 + jp1], uw = F[i * A1 + jm1];
    v.u = uc;
    v.g0 = central_diff(un, us, c.two_d0);
    v.g1 = central_diff(ue, uw, c.two_d1);
    v.lap = __dadd_rn(second_diff(un, uc, us, c.d0sq), second_diff(ue, uc, uw, c.d1sq));
    v.bih = 0.0;
    if constexpr (BIH) {
        const int64_t im2 = wrap(i - 2, A0), ip2 = wrap(i + 2, A0), jm2 = wrap(j - 2, A1), jp2 = wrap(j + 2, A1);
        const double unn = F[ip2 * A1 + j], uss = F[im2 * A1 + j], uee = F[i * A1 + jp2], uww = F[i * A1 + jm2];
        const double une = F[ip1 * A1 + jp1], unw = F[ip1 * A1 + jm1], use_ = F[im1 * A1 + jp1], usw = F[im1 * A1 + jm1];
        const double ln = __dadd_rn(second_diff(unn, un, uc, c.d0sq), second_diff(une, un, unw, c.d1sq));
        const double ls = __dadd_rn(second_diff(uc, us, uss, c.d0sq), second_diff(use_, us, usw, c.d1sq));
        const double le = __dadd_rn(second_diff(une, ue, use_, c.d0sq), second_diff(uee, ue, uc, c.d1sq));
        const double lw = __dadd_rn(second_diff(unw, uw, usw, c.d0sq), second_diff(uc, uw, uww, c.d1sq));
        v.bih = __dadd_rn(second_diff(ln, v.lap, ls, c.d0sq), second_diff(le, v.lap, lw, c.d1sq));
    }
}

// Interior (non-periodic) point of the basic_usage dialect; (i, j) are full-frame indices
// with 1 <= i < A0-1, 1 <= j < A1-1 (basic:56-63).
__device__ __forceinline__ void basic_point(const double *__restrict__ F, int64_t A1, int64_t i, int64_t j,
                                            const FdConsts &c, PointVals &v) {
    const double uc = F[i * A1 + j];
    const double un = F[(i + 1) * A1 + j], us = F[(i - 1) * A1 + j], ue = F[i * A1 + j + 1], uw = F[i * A1 + j - 1];
    v.u = uc;
    v.g0 = central_diff(un, us, c.two_d0);
    v.g1 = central_diff(ue, uw, c.two_d1);
    // lap = u_xx + u_yy with x = a1 (basic:69); addition commutes, so operand order is free
    v.lap = __dadd_rn(second_diff(ue, uc, uw, c.d1sq), second_diff(un, uc, us, c.d0sq));
    v.bih = 0.0;
}

// Slice-aligned point of the analyze_results dialect (analyze_results:257-274): the reference differences slices
// that it then crops to a common ORIGIN, so the "central" differences of row (i, j) reach forward only:
// (i, j), (i+1, j), (i+2, j), (i, j+1), (i, j+2); valid for i < A0-2, j < A1-2.
__device__ __forceinline__ void slice_point(const double *__restrict__ F, int64_t A1, int64_t i, int64_t j,
                                            const FdConsts &c, PointVals &v) {
    const double uc = F[i * A1 + j];
    const double e1 = F[i * A1 + j + 1], e2 = F[i * A1 + j + 2], n1 = F[(i + 1) * A1 + j], n2 = F[(i + 2) * A1 + j];
    v.u = uc;
    v.g1 = central_diff(e2, uc, c.two_d1);             // u_x  (:257)
    v.g0 = central_diff(n2, uc, c.two_d0);             // u_y  (:258)
    v.d11 = second_diff(e2, e1, uc, c.d1sq);           // u_xx (:259)
    v.d00 = second_diff(n2, n1, uc, c.d0sq);           // u_yy (:260)
    v.lap = __dadd_rn(v.d11, v.d00);                   // laplacian = u_xx + u_yy (:274)
    v.bih = 0.0;
}

// ----------------------------------------------------------------------------- statistics entries
// Every entry of the statistics vector is a product ext[a]*ext[b] of the extended row
// ext = [1, y, theta_0 .. theta_{p-1}]:  n = 1*1, sum y = 1*y, sum y^2 = y*y,
// sum theta_j = 1*theta_j, sum theta_j y = y*theta_j, G_ij = theta_i*theta_j.
__host__ __device__ inline void stats_pair(int e, int p, int &a, int &b) {
    if (e == 0) { a = 0; b = 0; return; }
    if (e == 1) { a = 0; b = 1; return; }
    if (e == 2) { a = 1; b = 1; return; }
    if (e < 3 + p) { a = 0; b = 2 + (e - 3); return; }
    if (e < 3 + 2 * p) { a = 1; b = 2 + (e - 3 - p); return; }
    int k = e - 3 - 2 * p;
    int i = 0;
    while (k >= p - i) { k -= p - i; ++i; }
    a = 2 + i; b = 2 + i + k;
}

constexpr int kMaxStats = PG_STATS_LEN(PG_MAX_P);

// Warp-cooperative accumulation of up to 32 rows (one per lane) into a warp-private
// accumulator in shared memory.  Lane L owns entries L, L+32, ... so no two lanes touch
// the same address.  `ext` is the warp's staging area [32][p+2].
__device__ __forceinline__ void warp_accumulate_rows(double *wacc, double *ext, const uint8_t *pair_a,
                                                     const uint8_t *pair_b, int p, int S, int lane, bool valid,
                                                     int fold, const double *th, double y) {
    const int W = p + 2;
    ext[lane * W + 0] = 1.0;
    ext[lane * W + 1] = y;
    for (int j = 0; j < p; ++j) ext[lane * W + 2 + j] = th[j];
    __syncwarp();
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    for (int l = 0; l < 32; ++l) {
        if (!((vmask >> l) & 1u)) continue;
        const int f = __shfl_sync(0xffffffffu, fold, l);
        const double *r = ext + l * W;
        double *acc = wacc + f * S;
        for (int e = lane; e < S; e += 32) acc[e] = fma(r[pair_a[e]], r[pair_b[e]], acc[e]);
    }
    __syncwarp();
}

}  // namespace pg
