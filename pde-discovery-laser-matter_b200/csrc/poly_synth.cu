// K2 -- local-polynomial derivative rows (patch:193-280) as a fixed (2rt+1)(2rs+1)^2-tap stencil,
// and the synthetic-field generator used for the large benchmark stacks.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "launch.h"

namespace pg {

constexpr int PW = 8;  // warps per CTA

template <typename TIn>
__global__ void __launch_bounds__(PW * 32) poly_rows_kernel(const TIn *__restrict__ U, int64_t T, int64_t H, int64_t W,
                                                           const int32_t *__restrict__ pts, int64_t n,
                                                           const double *__restrict__ W6, int rt, int rs, int mode,
                                                           double *__restrict__ X, double *__restrict__ y,
                                                           unsigned long long *counters) {
    extern __shared__ double wsm[];  // [6][nnb]
    const int side = 2 * rs + 1, nnb = (2 * rt + 1) * side * side;
    for (int e = threadIdx.x; e < 6 * nnb; e += blockDim.x) wsm[e] = W6[e];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = mode == 1 ? 8 : 6;  // mode 0 = model4, 1 = full, 2 = raw derivatives
    for (int64_t k = (int64_t)blockIdx.x * PW + warp; k < n; k += (int64_t)gridDim.x * PW) {
        const int64_t t0 = pts[3 * k], y0 = pts[3 * k + 1], x0 = pts[3 * k + 2];
        const bool ok = t0 - rt >= 0 && t0 + rt < T && y0 - rs >= 0 && y0 + rs < H && x0 - rs >= 0 && x0 + rs < W;
        double acc[6] = {0, 0, 0, 0, 0, 0};
        if (ok) {
            for (int e = lane; e < nnb; e += 32) {
                const int ox = e % side, oy = (e / side) % side, ot = e / (side * side);
                const double v = (double)U[((t0 - rt + ot) * H + (y0 - rs + oy)) * W + (x0 - rs + ox)];
#pragma unroll
                for (int q = 0; q < 6; ++q) acc[q] = fma(wsm[q * nnb + e], v, acc[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 6; ++q)
#pragma unroll
            for (int o = 16; o; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        if (lane == 0) {
            if (!ok) {  // the reference would raise on an out-of-range neighbourhood; poison the row
                for (int c = 0; c < p; ++c) X[k * p + c] = nan("");
                y[k] = nan("");
                atomicAdd(&counters[1], 1ull);
            } else {
                const double u = acc[0], ux = acc[2], uy = acc[3], lap = __dadd_rn(acc[4], acc[5]);
                double *r = X + k * p;
                if (mode == 2) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) r[q] = acc[q];
                } else {
                    r[0] = 1.0; r[1] = u; r[2] = ux; r[3] = uy; r[4] = lap; r[5] = __dmul_rn(u, u);
                    if (mode == 1) { r[6] = __dmul_rn(u, ux); r[7] = __dmul_rn(u, uy); }
                }
                y[k] = acc[1];
            }
        }
    }
}

// Fast path for neighbourhoods of at most 256 taps (the script's default rt = 2, rs = 3: 245 taps).  The stencil
// is the same for every point, so each lane keeps its (up to) 8 taps -- offsets and the 6 x 8 weights -- in
// REGISTERS for the whole kernel: per point a lane issues 8 independent gathers and 48 FMAs, with no shared
// memory and no index arithmetic.  NPT points are in flight per warp (NPT x 8 gathers per lane outstanding), and
// the six accumulators are reduced with a packed butterfly (half of the lanes take over half of the values at
// each step: 8 shuffles instead of 30).
constexpr int PNE = 8;   // taps per lane
constexpr int NPT = 4;   // points in flight per warp

__device__ __forceinline__ void reduce6(double (&a)[6], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    // xor 16: bit4 = 0 keeps a0..a2, bit4 = 1 keeps a3..a5
    double k0 = b4 ? a[3] : a[0], k1 = b4 ? a[4] : a[1], k2 = b4 ? a[5] : a[2];
    k0 += __shfl_xor_sync(0xffffffffu, b4 ? a[0] : a[3], 16);
    k1 += __shfl_xor_sync(0xffffffffu, b4 ? a[1] : a[4], 16);
    k2 += __shfl_xor_sync(0xffffffffu, b4 ? a[2] : a[5], 16);
    // xor 8: bit3 = 0 keeps (k0, k1), bit3 = 1 keeps k2
    const double r = __shfl_xor_sync(0xffffffffu, b3 ? k0 : k2, 8);
    const double r2 = __shfl_xor_sync(0xffffffffu, k1, 8);
    double m0 = b3 ? k2 + r : k0 + r, m1 = k1 + r2;
    // xor 4: bit3 = 0: bit2 = 0 keeps m0, bit2 = 1 keeps m1; bit3 = 1: plain add of m0
    const double snd = b3 ? m0 : (b2 ? m0 : m1);
    double m = (b3 ? m0 : (b2 ? m1 : m0)) + __shfl_xor_sync(0xffffffffu, snd, 4);
    m += __shfl_xor_sync(0xffffffffu, m, 2);
    m += __shfl_xor_sync(0xffffffffu, m, 1);
    // value q now lives in lanes {0, 4, 8, 16, 20, 24}[q] (and their low-bit neighbours)
    a[0] = __shfl_sync(0xffffffffu, m, 0);  a[1] = __shfl_sync(0xffffffffu, m, 4);  a[2] = __shfl_sync(0xffffffffu, m, 8);
    a[3] = __shfl_sync(0xffffffffu, m, 16); a[4] = __shfl_sync(0xffffffffu, m, 20); a[5] = __shfl_sync(0xffffffffu, m, 24);
}

// WSMEM = false: weights in registers (8 warps per SM).  WSMEM = true: weights in shared memory, laid out
// [tap][output][lane] so that a warp's read of one weight is one conflict-free 256-byte row; the kernel then
// needs ~110 registers and two CTAs fit an SM, which matters because it is bound by the latency of its
// scattered gathers (ncu: long_scoreboard 4.2 per issue at 8 warps per SM).
template <typename TIn, bool WSMEM>
__global__ void __launch_bounds__(PW * 32, WSMEM ? 3 : 1) poly_rows_reg_kernel(const TIn *__restrict__ U, int64_t T, int64_t H, int64_t W,
                                                               const int32_t *__restrict__ pts, int64_t n,
                                                               const double *__restrict__ W6, int rt, int rs, int mode,
                                                               double *__restrict__ X, double *__restrict__ y,
                                                               unsigned long long *counters) {
    const int side = 2 * rs + 1, nnb = (2 * rt + 1) * side * side;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = mode == 1 ? 8 : 6;
    __shared__ double wsm[WSMEM ? PNE * 6 * 32 : 1];
    constexpr int NP = WSMEM ? 3 : NPT;   // points in flight per warp
    int off[PNE];                          // tap offsets in elements (the launcher checks that they fit 31 bits)
    double w[WSMEM ? 1 : 6][WSMEM ? 1 : PNE];
#pragma unroll
    for (int k = 0; k < PNE; ++k) {
        const int e = lane + 32 * k;
        const bool has = e < nnb;
        const int ee = has ? e : 0;
        const int ox = ee % side, oy = (ee / side) % side, ot = ee / (side * side);
        off[k] = (int)(((int64_t)ot * H + oy) * W + ox);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const double wv = has ? W6[q * nnb + e] : 0.0;   // absent taps: a valid address, weight 0
            if constexpr (WSMEM) { if (warp == 0) wsm[(k * 6 + q) * 32 + lane] = wv; }
            else w[q][k] = wv;
        }
    }
    if constexpr (WSMEM) __syncthreads();
    auto emit = [&](int64_t k, bool ok, double (&acc)[6]) {
        reduce6(acc, lane);
        if (lane != 0) return;
        if (!ok) {  // the reference would raise on an out-of-range neighbourhood; poison the row
            for (int c = 0; c < p; ++c) X[k * p + c] = nan("");
            y[k] = nan("");
            atomicAdd(&counters[1], 1ull);
            return;
        }
        const double u = acc[0], ux = acc[2], uy = acc[3], lap = __dadd_rn(acc[4], acc[5]);
        double *r = X + k * p;
        if (mode == 2) {
#pragma unroll
            for (int q = 0; q < 6; ++q) r[q] = acc[q];
        } else {
            r[0] = 1.0; r[1] = u; r[2] = ux; r[3] = uy; r[4] = lap; r[5] = __dmul_rn(u, u);
            if (mode == 1) { r[6] = __dmul_rn(u, ux); r[7] = __dmul_rn(u, uy); }
        }
        y[k] = acc[1];
    };
    auto base_of = [&](int64_t k, bool &ok) {
        const int64_t t0 = pts[3 * k], y0 = pts[3 * k + 1], x0 = pts[3 * k + 2];
        ok = t0 - rt >= 0 && t0 + rt < T && y0 - rs >= 0 && y0 + rs < H && x0 - rs >= 0 && x0 + rs < W;
        return ok ? ((t0 - rt) * H + (y0 - rs)) * W + (x0 - rs) : (int64_t)0;
    };
    const int64_t stride = (int64_t)gridDim.x * PW;
    for (int64_t k = (int64_t)blockIdx.x * PW + warp; k < n; k += NP * stride) {
        bool ok[NP];
        int64_t kk[NP], base[NP];
        double v[NP][PNE];
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            kk[j] = k + j * stride;
            ok[j] = false;
            base[j] = kk[j] < n ? base_of(kk[j], ok[j]) : 0;
        }
#pragma unroll
        for (int j = 0; j < NP; ++j)
#pragma unroll
            for (int e = 0; e < PNE; ++e) v[j][e] = ok[j] ? (double)U[base[j] + off[e]] : 0.0;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            double acc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int e = 0; e < PNE; ++e)
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    if constexpr (WSMEM) acc[q] = fma(wsm[(e * 6 + q) * 32 + lane], v[j][e], acc[q]);
                    else acc[q] = fma(w[q][e], v[j][e], acc[q]);
                }
            if (kk[j] < n) emit(kk[j], ok[j], acc);
        }
    }
}

int launch_poly_rows(const void *U, int dtype, int64_t T, int64_t H, int64_t W, const int32_t *pts, int64_t n,
                     const double *W6, int rt, int rs, int mode, double *X, double *y, unsigned long long *counters,
                     cudaStream_t st) {
    if (n <= 0) return PG_OK;
    const int side = 2 * rs + 1, nnb = (2 * rt + 1) * side * side;
    const size_t smem = sizeof(double) * 6 * nnb;
    int64_t g = (n + PW - 1) / PW;
    if (g > 148 * 8) g = 148 * 8;
    if (nnb <= 32 * PNE) {   // taps fit the per-lane registers
        const bool wsmem = !getenv("PG_POLY_WREG") && (int64_t)(2 * rt + 1) * H * W < 0x7fffffff;
        const int npt = wsmem ? 3 : NPT;
        int64_t gr = (n + npt * PW - 1) / (npt * PW);
        const int64_t cap = 148 * (wsmem ? 6 : 2);
        if (gr > cap) gr = cap;
        if (dtype == 0) {
            if (wsmem) poly_rows_reg_kernel<float, true><<<(unsigned)gr, PW * 32, 0, st>>>((const float *)U, T, H, W, pts, n, W6, rt, rs, mode, X, y, counters);
            else poly_rows_reg_kernel<float, false><<<(unsigned)gr, PW * 32, 0, st>>>((const float *)U, T, H, W, pts, n, W6, rt, rs, mode, X, y, counters);
        } else {
            if (wsmem) poly_rows_reg_kernel<double, true><<<(unsigned)gr, PW * 32, 0, st>>>((const double *)U, T, H, W, pts, n, W6, rt, rs, mode, X, y, counters);
            else poly_rows_reg_kernel<double, false><<<(unsigned)gr, PW * 32, 0, st>>>((const double *)U, T, H, W, pts, n, W6, rt, rs, mode, X, y, counters);
        }
        PG_LAUNCHED();
        return PG_OK;
    }
    if (dtype == 0) {
        PG_CUDA(cudaFuncSetAttribute(poly_rows_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        poly_rows_kernel<float><<<(unsigned)g, PW * 32, smem, st>>>((const float *)U, T, H, W, pts, n, W6, rt, rs, mode,
                                                                    X, y, counters);
    } else {
        PG_CUDA(cudaFuncSetAttribute(poly_rows_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        poly_rows_kernel<double><<<(unsigned)g, PW * 32, smem, st>>>((const double *)U, T, H, W, pts, n, W6, rt, rs,
                                                                     mode, X, y, counters);
    }
    PG_LAUNCHED();
    return PG_OK;
}

// ----------------------------------------------------------------------------- synthetic field
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void synth_kernel(double *__restrict__ U, int64_t T, int64_t A0, int64_t A1, int64_t t_offset, int64_t T_total,
                             uint64_t seed, int kind, double noise) {
    const int64_t frame = A0 * A1, total = T * frame;
    const double two_pi = 6.283185307179586476925286766559;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tl = idx / frame, r = idx % frame, i = r / A1, j = r % A1;
        const int64_t t = tl + t_offset;
        const double a = two_pi * (double)i / (double)A0, b = two_pi * (double)j / (double)A1;
        const double s = two_pi * (double)t / (double)(T_total > 0 ? T_total : 1);
        // a few travelling waves with integer wave numbers (periodic in a0, a1)
        double v = 0.50 * sin(3.0 * a + 2.0 * b - 5.0 * s) + 0.30 * sin(7.0 * a - 4.0 * b + 3.0 * s + 0.7) +
                   0.15 * sin(13.0 * a + 11.0 * b - 9.0 * s + 1.9) + 0.05 * cos(29.0 * a - 17.0 * b + 2.0 * s);
        const uint64_t hsh = splitmix64(seed ^ splitmix64((uint64_t)t * 0x100000001B3ull + (uint64_t)r));
        const double un = (double)(hsh >> 11) * (1.0 / 9007199254740992.0) - 0.5;  // U(-0.5, 0.5)
        v += noise * un;
        U[idx] = kind == 1 ? 0.5 + 0.4 * v : v;
    }
}

int launch_synth(double *U, int64_t T, int64_t A0, int64_t A1, int64_t t_offset, int64_t T_total, uint64_t seed,
                 int kind, double noise, cudaStream_t st) {
    const int64_t total = T * A0 * A1;
    if (total <= 0) return PG_OK;
    int64_t g = (total + 255) / 256;
    if (g > 148 * 32) g = 148 * 32;
    synth_kernel<<<(unsigned)g, 256, 0, st>>>(U, T, A0, A1, t_offset, T_total, seed, kind, noise);
    PG_LAUNCHED();
    return PG_OK;
}

}  // namespace pg
