// Optional denoising prologue of the ks2d script, the step BEFORE the hot path (ks2d:125-161, 1448-1468):
//   time_smooth_moving_average   reflect-padded moving average along t through a cumulative sum; the kernel
//                                advances two copies of the same sequential cumulative sum (window apart), so
//                                the result is bit-identical to NumPy's cumsum formulation
//   gaussian_smooth_periodic_2d  the reference multiplies the 2-D FFT by exp(-sigma^2 |k|^2 / 2); that is a
//                                separable circular convolution with the periodic Gaussian g = ifft(exp(-sigma^2
//                                k^2 / 2)), whose taps beyond ~9 sigma are below 1e-17 of the peak: two passes of
//                                a short periodic stencil (taps computed on the host) reproduce the FFT result
//                                to rounding without any FFT
#include <dlfcn.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "launch.h"

namespace pg {

// U_pad index q (0 .. T + 2 pad - 1) -> frame of U under np.pad(mode="reflect")
__device__ __forceinline__ int64_t reflect_index(int64_t q, int64_t pad, int64_t T) {
    int64_t t = q - pad;
    if (t < 0) t = -t;
    if (t >= T) t = 2 * (T - 1) - t;
    return t;
}

__global__ void time_moving_average_kernel(const double *__restrict__ U, int64_t T, int64_t frame, int window,
                                           double *__restrict__ out) {
    const int64_t pad = window / 2;
    const double w = (double)window;
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < frame; x += (int64_t)gridDim.x * blockDim.x) {
        // cs[k] = sum_{q<k} U_pad[q] (sequential); out[t] = (cs[t + window] - cs[t]) / window
        double lead = 0.0, lag = 0.0;
        for (int q = 0; q < window; ++q) lead = __dadd_rn(lead, U[reflect_index(q, pad, T) * frame + x]);
        for (int64_t t = 0; t < T; ++t) {
            out[t * frame + x] = __ddiv_rn(__dsub_rn(lead, lag), w);
            if (t + 1 < T) {
                lead = __dadd_rn(lead, U[reflect_index(t + window, pad, T) * frame + x]);
                lag = __dadd_rn(lag, U[reflect_index(t, pad, T) * frame + x]);
            }
        }
    }
}

// The same sums for the usual small windows with the last W padded frames of a column pair kept in registers: the lag
// sum adds ring values instead of re-reading frames that have left L2 (W frames of 2048^2 are 34 W MB), so the stack is
// read once and written once.  Same additions in the same order as above: bit-identical.
template <int W>
__global__ void __launch_bounds__(256) time_moving_average_ring_kernel(const double2 *__restrict__ U, int64_t T, int64_t frame2,
                                                                       double2 *__restrict__ out) {
    constexpr int64_t pad = W / 2;
    const double w = (double)W;
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < frame2; x += (int64_t)gridDim.x * blockDim.x) {
        double2 ring[W], lead = {0.0, 0.0}, lag = {0.0, 0.0};
#pragma unroll
        for (int q = 0; q < W; ++q) {
            ring[q] = U[reflect_index(q, pad, T) * frame2 + x];
            lead.x = __dadd_rn(lead.x, ring[q].x);
            lead.y = __dadd_rn(lead.y, ring[q].y);
        }
        for (int64_t t0 = 0; t0 < T; t0 += W) {
            double2 nu[W];                                    // the next W padded frames: W independent loads in flight
#pragma unroll
            for (int k = 0; k < W; ++k)
                if (t0 + k + 1 < T) nu[k] = U[reflect_index(t0 + k + W, pad, T) * frame2 + x];
#pragma unroll
            for (int k = 0; k < W; ++k) {                     // ring[k] holds U_pad[t] for t = t0 + k
                const int64_t t = t0 + k;
                if (t < T) {
                    out[t * frame2 + x] = make_double2(__ddiv_rn(__dsub_rn(lead.x, lag.x), w), __ddiv_rn(__dsub_rn(lead.y, lag.y), w));
                    if (t + 1 < T) {
                        lead.x = __dadd_rn(lead.x, nu[k].x);
                        lead.y = __dadd_rn(lead.y, nu[k].y);
                        lag.x = __dadd_rn(lag.x, ring[k].x);
                        lag.y = __dadd_rn(lag.y, ring[k].y);
                        ring[k] = nu[k];
                    }
                }
            }
        }
    }
}

// out[t][i][j] = sum_k w[k] in[t][wrap(i - off[k])][j]  (axis 0)  or  in[t][i][wrap(j - off[k])]  (axis 1)
// One thread per output, rows of the stack (t, i) over blockIdx.y / grid-stride, columns over blockIdx.x: no division
// per point.  Taps are accumulated in the order given (the order of the host's tap list fixes the rounding).
__global__ void periodic_conv_kernel(const double *__restrict__ in, int64_t T, int64_t A0, int64_t A1, int axis,
                                     const int32_t *__restrict__ off, const double *__restrict__ w, int n_taps,
                                     double *__restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A1) return;
    for (int64_t row = blockIdx.y; row < T * A0; row += gridDim.y) {
        const int64_t t = row / A0, i = row - t * A0;       // one division per row of threads
        const double *F = in + t * A0 * A1;
        double s = 0.0;
        if (axis == 0) {
            for (int k = 0; k < n_taps; ++k) s = fma(__ldg(w + k), F[wrap(i - __ldg(off + k), A0) * A1 + j], s);
        } else {
            const double *R = F + i * A1;
            for (int k = 0; k < n_taps; ++k) s = fma(__ldg(w + k), R[wrap(j - __ldg(off + k), A1)], s);
        }
        out[row * A1 + j] = s;
    }
}

// scipy.ndimage.gaussian_filter (patch:335,343; analyze_results:222,250): one axis of the separable filter with
// mode="reflect" (half-sample symmetric: d c b a | a b c d | d c b a).  scipy's correlate1d accumulates in double,
// centre tap first, then the symmetric pairs from the outermost inwards, (in[l+j] + in[l-j]) * w[j]; this kernel adds
// in the same order without contraction and rounds to the array's dtype once per axis, so the result is bit-identical
// to scipy for float64 and for float32 stacks.  w [2 r + 1] are scipy's normalised taps (w[r] the centre).
__device__ __forceinline__ int64_t symmetric_index(int64_t i, int64_t n) {
    while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - i - 1;
    return i;
}

template <typename T_>
__global__ void reflect_conv_kernel(const T_ *__restrict__ in, int64_t T, int64_t A0, int64_t A1, int axis,
                                    const double *__restrict__ w, int radius, T_ *__restrict__ out) {
    // columns over blockIdx.x, rows of the stack (t, i) over blockIdx.y / grid-stride: one division per row of threads
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A1) return;
    for (int64_t row = blockIdx.y; row < T * A0; row += gridDim.y) {
        const int64_t t = row / A0, i = row - t * A0;
        const T_ *F = in + t * A0 * A1;
        double acc = __dmul_rn((double)F[i * A1 + j], __ldg(w + radius));
        for (int jj = -radius; jj < 0; ++jj) {
            double a, b;
            if (axis == 0) {
                a = (double)F[symmetric_index(i + jj, A0) * A1 + j];
                b = (double)F[symmetric_index(i - jj, A0) * A1 + j];
            } else {
                a = (double)F[i * A1 + symmetric_index(j + jj, A1)];
                b = (double)F[i * A1 + symmetric_index(j - jj, A1)];
            }
            acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(a, b), __ldg(w + radius + jj)));
        }
        out[row * A1 + j] = (T_)acc;
    }
}

// Both axes of the same filter in ONE pass over the stack (what gaussian_filter(frame, sigma) does per frame: axis 0,
// then axis 1): a CTA loads a (32 + 2 r) x (128 + 2 r) window of a frame with the reflections resolved, filters it
// along axis 0 into a 32 x (128 + 2 r) intermediate in shared memory -- rounded to the stack's dtype like scipy's
// intermediate array -- and along axis 1 into the 32 x 128 output tile.  The additions are those of
// reflect_conv_kernel in the same order, so the result is the same bit for bit; the stack is read once (1.25x with
// the halo, from L2) and written once instead of two reads, two writes and a nine-fold re-read through L2.
#ifndef PG_RG_TH
#define PG_RG_TH 32
#endif
#ifndef PG_RG_NW
#define PG_RG_NW 8
#endif
constexpr int RG_TH = PG_RG_TH, RG_TW = 128, RG_MAX_R = 32, RG_NW = PG_RG_NW, RG_Q = RG_TH / RG_NW;   // RG_Q rows per thread
static_assert(RG_TH % RG_NW == 0, "tile rows must be a multiple of the warps");
template <typename T_>
__global__ void __launch_bounds__(32 * RG_NW) reflect_gauss2d_kernel(const T_ *__restrict__ in, int64_t A0, int64_t A1, int tiles0,
                                                              int tiles1, const double *__restrict__ w, int r,
                                                              T_ *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char rg_smem[];
    const int SW = RG_TW + 2 * r, SH = RG_TH + 2 * r;
    double *wt = reinterpret_cast<double *>(rg_smem);                      // [2 r + 1]
    T_ *tin = reinterpret_cast<T_ *>(wt + ((2 * r + 2) & ~1));             // [SH][SW]
    T_ *mid = tin + SH * SW;                                               // [RG_TH][SW]
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tile = blockIdx.x;
    const int tj = (int)(tile % tiles1), ti = (int)((tile / tiles1) % tiles0);
    const int64_t t = tile / ((int64_t)tiles1 * tiles0);
    const int64_t i0 = (int64_t)ti * RG_TH, j0 = (int64_t)tj * RG_TW;
    const T_ *F = in + t * A0 * A1;
    for (int k = threadIdx.x; k < 2 * r + 1; k += 32 * RG_NW) wt[k] = w[k];
    {
        // a lane's <= 6 window columns are the same for every row: resolve their reflections once, then every row is a
        // batch of independent loads
        constexpr int NC = (RG_TW + 2 * RG_MAX_R + 31) / 32;
        int64_t gj[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) gj[k] = lane + 32 * k < SW ? symmetric_index(j0 - r + lane + 32 * k, A1) : -1;
#pragma unroll 2
        for (int a = wid; a < SH; a += RG_NW) {
            const T_ *R = F + symmetric_index(i0 - r + a, A0) * A1;
            T_ v[NC];
#pragma unroll
            for (int k = 0; k < NC; ++k)
                if (gj[k] >= 0) v[k] = R[gj[k]];
#pragma unroll
            for (int k = 0; k < NC; ++k)
                if (gj[k] >= 0) tin[a * SW + lane + 32 * k] = v[k];
        }
    }
    __syncthreads();
    const double wc = wt[r];
    // axis 0: rows wid, wid + RG_NW, .. of the intermediate (RG_Q independent accumulations per thread)
    for (int b = lane; b < SW; b += 32) {
        double acc[RG_Q];
#pragma unroll
        for (int q = 0; q < RG_Q; ++q) acc[q] = __dmul_rn((double)tin[(wid + RG_NW * q + r) * SW + b], wc);
        for (int jj = -r; jj < 0; ++jj) {
            const double wv = wt[r + jj];
#pragma unroll
            for (int q = 0; q < RG_Q; ++q) {
                const int c = (wid + RG_NW * q + r) * SW + b;
                acc[q] = __dadd_rn(acc[q], __dmul_rn(__dadd_rn((double)tin[c + jj * SW], (double)tin[c - jj * SW]), wv));
            }
        }
#pragma unroll
        for (int q = 0; q < RG_Q; ++q) mid[(wid + RG_NW * q) * SW + b] = (T_)acc[q];
    }
    __syncthreads();
    // axis 1
    for (int b = lane; b < RG_TW; b += 32) {
        if (j0 + b >= A1) break;
        double acc[RG_Q];
#pragma unroll
        for (int q = 0; q < RG_Q; ++q) acc[q] = __dmul_rn((double)mid[(wid + RG_NW * q) * SW + b + r], wc);
        for (int jj = -r; jj < 0; ++jj) {
            const double wv = wt[r + jj];
#pragma unroll
            for (int q = 0; q < RG_Q; ++q) {
                const int c = (wid + RG_NW * q) * SW + b + r;
                acc[q] = __dadd_rn(acc[q], __dmul_rn(__dadd_rn((double)mid[c + jj], (double)mid[c - jj]), wv));
            }
        }
#pragma unroll
        for (int q = 0; q < RG_Q; ++q)
            if (i0 + wid + RG_NW * q < A0) out[(t * A0 + i0 + wid + RG_NW * q) * A1 + j0 + b] = (T_)acc[q];
    }
}

template <typename T_>
static int launch_reflect_gauss2d_t(const T_ *in, int64_t T, int64_t A0, int64_t A1, const double *w, int r, T_ *out, cudaStream_t st) {
    const int64_t tiles0 = (A0 + RG_TH - 1) / RG_TH, tiles1 = (A1 + RG_TW - 1) / RG_TW;
    if (T * tiles0 * tiles1 > 0x7fffffffLL) return PG_EINVAL;
    const size_t smem = 8 * (size_t)((2 * r + 2) & ~1) + sizeof(T_) * (size_t)(RG_TW + 2 * r) * (2 * RG_TH + 2 * r);
    cudaFuncSetAttribute(reflect_gauss2d_kernel<T_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    reflect_gauss2d_kernel<T_><<<(unsigned)(T * tiles0 * tiles1), 32 * RG_NW, smem, st>>>(in, A0, A1, (int)tiles0, (int)tiles1, w, r, out);
    PG_LAUNCHED();
    return PG_OK;
}

// The same for the radii the scripts use (sigma = 1.0, 1.2, 1.5 -> 4, 5, 6; also 2, 3, 8) with the radius a compile-time
// constant, so that the filter windows live in REGISTERS: along axis 0 a thread marches 18 rows down one column of
// the window tile (18 + 2 R shared-memory reads for 18 outputs instead of 9 per output), along axis 1 it forms two
// neighbouring outputs from R + 1 16-byte reads.  Shared memory holds doubles for both dtypes (a float32 stack is
// widened once per loaded value and the intermediate is rounded through float), so no conversion sits in the inner
// loops.  Same additions in the same order as the kernels above.  9 warps: 2 x (128 + 2 R) <= 288 column marches,
// 36 half rows of the 18 x 128 output tile = 4 per warp; the small tile (48 KB of shared memory at R = 4) keeps four
// CTAs per SM in different phases (measured, 256 x 2048^2 float64, sigma = 1: 36-row tiles 5.8 ms, 18-row tiles 4.5 ms).
#ifndef PG_RW_TH
#define PG_RW_TH 18
#endif
constexpr int RW_TH = PG_RW_TH, RW_TW = 128, RW_NW = 9, RW_H1 = RW_TH / 2;
__device__ __forceinline__ int symmetric_index32(int i, int n) {
    while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - i - 1;
    return i;
}

// grid (tiles along a1, tiles along a0, frames of this launch); offsets inside a frame are 32-bit (A0 A1 < 2^31)
template <typename T_, int R>
__global__ void __launch_bounds__(32 * RW_NW) reflect_gauss2d_win_kernel(const T_ *__restrict__ in, int A0, int A1,
                                                                        const double *__restrict__ w, T_ *__restrict__ out) {
    constexpr int SW = RW_TW + 2 * R, SH = RW_TH + 2 * R, NC = (SW + 31) / 32;
    extern __shared__ __align__(16) unsigned char rg_smem[];
    double *tin = reinterpret_cast<double *>(rg_smem);                     // [SH][SW]
    double *mid = tin + SH * SW;                                           // [RW_TH][SW]
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i0 = blockIdx.y * RW_TH, j0 = blockIdx.x * RW_TW;
    const int64_t fo = (int64_t)blockIdx.z * A0 * A1;
    const T_ *F = in + fo;
    double wt[R + 1];                                                      // w[0 .. R]: the taps are symmetric
#pragma unroll
    for (int k = 0; k <= R; ++k) wt[k] = __ldg(w + k);
    {
        int gj[NC];                                                        // the lane's window columns, reflections resolved
#pragma unroll
        for (int k = 0; k < NC; ++k) gj[k] = symmetric_index32(j0 - R + min(lane + 32 * k, SW - 1), A1);
#pragma unroll 2
        for (int a = wid; a < SH; a += RW_NW) {
            const T_ *Rw = F + symmetric_index32(i0 - R + a, A0) * A1;
            T_ v[NC];
#pragma unroll
            for (int k = 0; k < NC; ++k) v[k] = Rw[gj[k]];
#pragma unroll
            for (int k = 0; k < NC; ++k)
                if (32 * (k + 1) <= SW || lane + 32 * k < SW) tin[a * SW + lane + 32 * k] = (double)v[k];
        }
    }
    __syncthreads();
    // axis 0: thread (seg, c) marches rows seg * RW_H1 .. + RW_H1 - 1 of column c
    if (threadIdx.x < 2 * SW) {
        const int seg = threadIdx.x >= SW, c = threadIdx.x - seg * SW;
        const double *col = tin + seg * RW_H1 * SW + c;
        double v[RW_H1 + 2 * R];
#pragma unroll
        for (int k = 0; k < RW_H1 + 2 * R; ++k) v[k] = col[k * SW];
#pragma unroll
        for (int o = 0; o < RW_H1; ++o) {
            double acc = __dmul_rn(v[o + R], wt[R]);
#pragma unroll
            for (int jj = -R; jj < 0; ++jj) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(v[o + R + jj], v[o + R - jj]), wt[R + jj]));
            mid[(seg * RW_H1 + o) * SW + c] = (double)(T_)acc;
        }
    }
    __syncthreads();
    // axis 1: warp wid takes half rows wid, wid + 9, ..; a lane the outputs 2 lane, 2 lane + 1 of the half row
    const bool vec = (A1 & 1) == 0;                                        // rows stay aligned for paired stores
#pragma unroll 2
    for (int h = wid; h < 2 * RW_TH; h += RW_NW) {
        const int a = h >> 1, b0 = (h & 1) * 64 + 2 * lane;
        const double2 *src = reinterpret_cast<const double2 *>(mid + a * SW + b0);
        double v[2 * R + 2];
#pragma unroll
        for (int k = 0; k <= R; ++k) {
            const double2 x = src[k];
            v[2 * k] = x.x;
            v[2 * k + 1] = x.y;
        }
        double acc[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            acc[q] = __dmul_rn(v[q + R], wt[R]);
#pragma unroll
            for (int jj = -R; jj < 0; ++jj) acc[q] = __dadd_rn(acc[q], __dmul_rn(__dadd_rn(v[q + R + jj], v[q + R - jj]), wt[R + jj]));
        }
        const int i = i0 + a, j = j0 + b0;
        if (i < A0) {
            T_ *dst = out + fo + (i * A1 + j);
            if (vec && j + 1 < A1) {
                if constexpr (sizeof(T_) == 8) *reinterpret_cast<double2 *>(dst) = make_double2(acc[0], acc[1]);
                else *reinterpret_cast<float2 *>(dst) = make_float2((float)acc[0], (float)acc[1]);
            } else {
                if (j < A1) dst[0] = (T_)acc[0];
                if (j + 1 < A1) dst[1] = (T_)acc[1];
            }
        }
    }
}

template <typename T_, int R>
static int launch_reflect_gauss2d_win(const T_ *in, int64_t T, int64_t A0, int64_t A1, const double *w, T_ *out, cudaStream_t st) {
    const int64_t tiles0 = (A0 + RW_TH - 1) / RW_TH, tiles1 = (A1 + RW_TW - 1) / RW_TW;
    const size_t smem = 8 * (size_t)(RW_TW + 2 * R) * (2 * RW_TH + 2 * R);
    cudaFuncSetAttribute(reflect_gauss2d_win_kernel<T_, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int64_t t0 = 0; t0 < T; t0 += 65535) {                            // gridDim.z <= 65535
        const int64_t nt = T - t0 < 65535 ? T - t0 : 65535;
        reflect_gauss2d_win_kernel<T_, R><<<dim3((unsigned)tiles1, (unsigned)tiles0, (unsigned)nt), 32 * RW_NW, smem, st>>>(
            in + t0 * A0 * A1, (int)A0, (int)A1, w, out + t0 * A0 * A1);
    }
    PG_LAUNCHED();
    return PG_OK;
}

template <typename T_>
static int launch_reflect_gauss2d_any(const T_ *in, int64_t T, int64_t A0, int64_t A1, const double *w, int r, T_ *out, cudaStream_t st) {
    const bool win_ok = ((uintptr_t)out % 16) == 0 && A0 * A1 < 0x7fffffffLL && (A0 + RW_TH - 1) / RW_TH <= 65535;
    if (win_ok && !getenv("PG_GAUSS_GENERIC")) switch (r) {
        case 2: return launch_reflect_gauss2d_win<T_, 2>(in, T, A0, A1, w, out, st);
        case 3: return launch_reflect_gauss2d_win<T_, 3>(in, T, A0, A1, w, out, st);
        case 4: return launch_reflect_gauss2d_win<T_, 4>(in, T, A0, A1, w, out, st);
        case 5: return launch_reflect_gauss2d_win<T_, 5>(in, T, A0, A1, w, out, st);
        case 6: return launch_reflect_gauss2d_win<T_, 6>(in, T, A0, A1, w, out, st);
        case 8: return launch_reflect_gauss2d_win<T_, 8>(in, T, A0, A1, w, out, st);
        default: break;
    }
    return launch_reflect_gauss2d_t<T_>(in, T, A0, A1, w, r, out, st);
}

int launch_reflect_gauss2d(const void *in, int dtype, int64_t T, int64_t A0, int64_t A1, const double *w, int radius, void *out,
                           cudaStream_t st) {
    if (radius > RG_MAX_R) return PG_EINVAL;
    return dtype == 0 ? launch_reflect_gauss2d_any<float>((const float *)in, T, A0, A1, w, radius, (float *)out, st)
                      : launch_reflect_gauss2d_any<double>((const double *)in, T, A0, A1, w, radius, (double *)out, st);
}

static unsigned grid_of(int64_t items) {
    int64_t g = (items + 255) / 256;
    return (unsigned)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

int launch_time_moving_average(const double *U, int64_t T, int64_t A0, int64_t A1, int window, double *out, cudaStream_t st) {
    const int64_t frame = A0 * A1;
    const bool pairs = frame % 2 == 0 && T >= window && ((uintptr_t)U | (uintptr_t)out) % 16 == 0;
#define PG_TMA_RING(W_)                                                                                                          \
    time_moving_average_ring_kernel<W_><<<(unsigned)((frame / 2 + 255) / 256), 256, 0, st>>>((const double2 *)U, T, frame / 2, (double2 *)out)
    if (pairs && window == 3) PG_TMA_RING(3);
    else if (pairs && window == 5) PG_TMA_RING(5);
    else if (pairs && window == 7) PG_TMA_RING(7);
    else if (pairs && window == 9) PG_TMA_RING(9);
    else time_moving_average_kernel<<<grid_of(frame), 256, 0, st>>>(U, T, frame, window, out);
#undef PG_TMA_RING
    PG_LAUNCHED();
    return PG_OK;
}

int launch_reflect_conv(const void *in, int dtype, int64_t T, int64_t A0, int64_t A1, int axis, const double *w, int radius,
                        void *out, cudaStream_t st) {
    const int64_t rows = T * A0;
    dim3 grid((unsigned)((A1 + 127) / 128), (unsigned)(rows < 148 * 64 ? rows : 148 * 64));
    if (dtype == 0)
        reflect_conv_kernel<float><<<grid, 128, 0, st>>>((const float *)in, T, A0, A1, axis, w, radius, (float *)out);
    else
        reflect_conv_kernel<double><<<grid, 128, 0, st>>>((const double *)in, T, A0, A1, axis, w, radius, (double *)out);
    PG_LAUNCHED();
    return PG_OK;
}

int launch_periodic_conv(const double *in, int64_t T, int64_t A0, int64_t A1, int axis, const int32_t *off, const double *w,
                         int n_taps, double *out, cudaStream_t st) {
    const int64_t rows = T * A0;
    dim3 grid((unsigned)((A1 + 127) / 128), (unsigned)(rows < 148 * 64 ? rows : 148 * 64));
    periodic_conv_kernel<<<grid, 128, 0, st>>>(in, T, A0, A1, axis, off, w, n_taps, out);
    PG_LAUNCHED();
    return PG_OK;
}

// ----------------------------------------------------------------------------- periodic Gaussian through the FFT
// gaussian_smooth_periodic_2d IS an FFT product in the reference (ks2d:125-142).  For sigma below ~2.8 px the periodic
// Gaussian rings and the direct convolution needs ALL n taps per axis (O(n) per point: seconds on a 2048^2 stack), so
// large frames go the reference's own way: batched real-to-complex 2-D transforms, the product with
// H = exp(-sigma^2 (kx^2 + ky^2) / 2) / (A0 A1) in one kernel, and the inverse.  cuFFT is a library call like the
// reference's np.fft (it is resolved with dlopen at first use, so the library has no link-time dependency on it and
// the entry point fails loudly where it is missing); the product kernel is ours.
namespace {
struct Cufft {
    void *h = nullptr;
    int (*PlanMany)(int *, int, int *, int *, int, int, int *, int, int, int, int) = nullptr;
    int (*SetStream)(int, cudaStream_t) = nullptr;
    int (*ExecD2Z)(int, double *, double2 *) = nullptr;
    int (*ExecZ2D)(int, double2 *, double *) = nullptr;
    int (*Destroy)(int) = nullptr;
    bool ok = false, tried = false;
};
Cufft g_cufft;
std::mutex g_cufft_mu;
struct PlanKey {
    int dev;
    int64_t a0, a1, batch;
    bool operator<(const PlanKey &o) const {
        return dev != o.dev ? dev < o.dev : a0 != o.a0 ? a0 < o.a0 : a1 != o.a1 ? a1 < o.a1 : batch < o.batch;
    }
};
std::map<PlanKey, std::pair<int, int>> g_plans;   // (forward D2Z, inverse Z2D)

bool cufft_load() {
    if (g_cufft.tried) return g_cufft.ok;
    g_cufft.tried = true;
    const char *names[] = {"libcufft.so.11", "/usr/local/cuda/lib64/libcufft.so.11", "libcufft.so", "/usr/local/cuda/lib64/libcufft.so"};
    for (const char *n : names) {
        g_cufft.h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (g_cufft.h) break;
    }
    if (!g_cufft.h) return false;
    g_cufft.PlanMany = (decltype(g_cufft.PlanMany))dlsym(g_cufft.h, "cufftPlanMany");
    g_cufft.SetStream = (decltype(g_cufft.SetStream))dlsym(g_cufft.h, "cufftSetStream");
    g_cufft.ExecD2Z = (decltype(g_cufft.ExecD2Z))dlsym(g_cufft.h, "cufftExecD2Z");
    g_cufft.ExecZ2D = (decltype(g_cufft.ExecZ2D))dlsym(g_cufft.h, "cufftExecZ2D");
    g_cufft.Destroy = (decltype(g_cufft.Destroy))dlsym(g_cufft.h, "cufftDestroy");
    g_cufft.ok = g_cufft.PlanMany && g_cufft.SetStream && g_cufft.ExecD2Z && g_cufft.ExecZ2D && g_cufft.Destroy;
    return g_cufft.ok;
}
}  // namespace

void fft_plans_release(int dev) {
    std::lock_guard<std::mutex> lk(g_cufft_mu);
    for (auto it = g_plans.begin(); it != g_plans.end();) {
        if (it->first.dev == dev) {
            if (g_cufft.ok) { g_cufft.Destroy(it->second.first); g_cufft.Destroy(it->second.second); }
            it = g_plans.erase(it);
        } else {
            ++it;
        }
    }
}

// spec[b][i][j] *= hx[i] * hy[j]   (j = 0 .. A1/2; hx carries the 1 / (A0 A1) of the unnormalised inverse)
__global__ void spectrum_scale_kernel(double2 *__restrict__ spec, int64_t rows, int64_t A0, int64_t nc,
                                      const double *__restrict__ hx, const double *__restrict__ hy) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nc) return;
    const double b = hy[j];
    for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
        const double a = hx[row % A0] * b;
        double2 v = spec[row * nc + j];
        v.x *= a; v.y *= a;
        spec[row * nc + j] = v;
    }
}

size_t periodic_gaussian_fft_scratch(int64_t T, int64_t A0, int64_t A1, int64_t *batch_out) {
    const int64_t per = A0 * (A1 / 2 + 1) * (int64_t)sizeof(double2);
    int64_t batch = (int64_t)(512ll << 20) / per;      // spectra of up to 512 MB at a time
    if (batch < 1) batch = 1;
    if (batch > T) batch = T;
    *batch_out = batch;
    return (size_t)(batch * per) + 2 * 16 + sizeof(double) * (size_t)(A0 + A1 / 2 + 1);
}

int launch_periodic_gaussian_fft(const double *in, int64_t T, int64_t A0, int64_t A1, const double *hx_host, const double *hy_host,
                                 double *out, void *scratch, int64_t batch, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_cufft_mu);
    if (!cufft_load()) PG_FAIL(PG_EUNSUPPORTED, "libcufft.so.11 could not be loaded (dlopen): use pg_periodic_conv");
    int dev = 0;
    PG_CUDA(cudaGetDevice(&dev));
    const int64_t nc = A1 / 2 + 1;
    double2 *spec = (double2 *)scratch;
    double *hx = (double *)((char *)scratch + ((batch * A0 * nc * sizeof(double2) + 15) / 16) * 16);
    double *hy = hx + A0;
    PG_CUDA(cudaMemcpyAsync(hx, hx_host, sizeof(double) * A0, cudaMemcpyHostToDevice, st));
    PG_CUDA(cudaMemcpyAsync(hy, hy_host, sizeof(double) * nc, cudaMemcpyHostToDevice, st));
    for (int64_t t0 = 0; t0 < T; t0 += batch) {
        const int64_t nb = T - t0 < batch ? T - t0 : batch;
        auto it = g_plans.find({dev, A0, A1, nb});
        if (it == g_plans.end()) {
            int n[2] = {(int)A0, (int)A1}, pf = 0, pi = 0;
            if (g_cufft.PlanMany(&pf, 2, n, nullptr, 1, 0, nullptr, 1, 0, 0x6a /* CUFFT_D2Z */, (int)nb) != 0 ||
                g_cufft.PlanMany(&pi, 2, n, nullptr, 1, 0, nullptr, 1, 0, 0x6c /* CUFFT_Z2D */, (int)nb) != 0)
                PG_FAIL(PG_ECUDA, "cufftPlanMany failed for %lld x %lld x %lld", (long long)nb, (long long)A0, (long long)A1);
            it = g_plans.emplace(PlanKey{dev, A0, A1, nb}, std::make_pair(pf, pi)).first;
        }
        const int pf = it->second.first, pi = it->second.second;
        if (g_cufft.SetStream(pf, st) != 0 || g_cufft.SetStream(pi, st) != 0) PG_FAIL(PG_ECUDA, "cufftSetStream failed");
        if (g_cufft.ExecD2Z(pf, const_cast<double *>(in) + t0 * A0 * A1, spec) != 0) PG_FAIL(PG_ECUDA, "cufftExecD2Z failed");
        PG_LAUNCHED();
        const int64_t rows = nb * A0;
        dim3 grid((unsigned)((nc + 127) / 128), (unsigned)(rows < 148 * 64 ? rows : 148 * 64));
        spectrum_scale_kernel<<<grid, 128, 0, st>>>(spec, rows, A0, nc, hx, hy);
        PG_LAUNCHED();
        if (g_cufft.ExecZ2D(pi, spec, out + t0 * A0 * A1) != 0) PG_FAIL(PG_ECUDA, "cufftExecZ2D failed");
        PG_LAUNCHED();
    }
    return PG_OK;
}

}  // namespace pg
