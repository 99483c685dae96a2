// Optional denoising prologue of the ks2d script, the step BEFORE the hot path (ks2d:125-161, 1448-1468):
//   time_smooth_moving_average   reflect-padded moving average along t through a cumulative sum; the kernel
//                                advances two copies of the same sequential cumulative sum (window apart), so
//                                the result is bit-identical to NumPy's cumsum formulation
//   gaussian_smooth_periodic_2d  the reference multiplies the 2-D FFT by exp(-sigma^2 |k|^2 / 2); that is a
//                                separable circular convolution with the periodic Gaussian g = ifft(exp(-sigma^2
//                                k^2 / 2)), whose taps beyond ~9 sigma are below 1e-17 of the peak: two passes of
//                                a short periodic stencil (taps computed on the host) reproduce the FFT result
//                                to rounding without any FFT
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

// out[t][i][j] = sum_k w[k] in[t][wrap(i - off[k])][j]  (axis 0)  or  in[t][i][wrap(j - off[k])]  (axis 1)
__global__ void periodic_conv_kernel(const double *__restrict__ in, int64_t T, int64_t A0, int64_t A1, int axis,
                                     const int32_t *__restrict__ off, const double *__restrict__ w, int n_taps,
                                     double *__restrict__ out) {
    const int64_t frame = A0 * A1, total = T * frame;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = idx / frame, r = idx % frame, i = r / A1, j = r % A1;
        const double *F = in + t * frame;
        double s = 0.0;
        if (axis == 0) {
            for (int k = 0; k < n_taps; ++k) s = fma(w[k], F[wrap(i - off[k], A0) * A1 + j], s);
        } else {
            for (int k = 0; k < n_taps; ++k) s = fma(w[k], F[i * A1 + wrap(j - off[k], A1)], s);
        }
        out[idx] = s;
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
    const int64_t frame = A0 * A1, total = T * frame;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = idx / frame, r = idx % frame, i = r / A1, j = r % A1;
        const T_ *F = in + t * frame;
        double acc = __dmul_rn((double)F[r], w[radius]);
        for (int jj = -radius; jj < 0; ++jj) {
            double a, b;
            if (axis == 0) {
                a = (double)F[symmetric_index(i + jj, A0) * A1 + j];
                b = (double)F[symmetric_index(i - jj, A0) * A1 + j];
            } else {
                a = (double)F[i * A1 + symmetric_index(j + jj, A1)];
                b = (double)F[i * A1 + symmetric_index(j - jj, A1)];
            }
            acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(a, b), w[radius + jj]));
        }
        out[idx] = (T_)acc;
    }
}

static unsigned grid_of(int64_t items) {
    int64_t g = (items + 255) / 256;
    return (unsigned)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

int launch_time_moving_average(const double *U, int64_t T, int64_t A0, int64_t A1, int window, double *out, cudaStream_t st) {
    time_moving_average_kernel<<<grid_of(A0 * A1), 256, 0, st>>>(U, T, A0 * A1, window, out);
    PG_LAUNCHED();
    return PG_OK;
}

int launch_reflect_conv(const void *in, int dtype, int64_t T, int64_t A0, int64_t A1, int axis, const double *w, int radius,
                        void *out, cudaStream_t st) {
    if (dtype == 0)
        reflect_conv_kernel<float><<<grid_of(T * A0 * A1), 256, 0, st>>>((const float *)in, T, A0, A1, axis, w, radius, (float *)out);
    else
        reflect_conv_kernel<double><<<grid_of(T * A0 * A1), 256, 0, st>>>((const double *)in, T, A0, A1, axis, w, radius, (double *)out);
    PG_LAUNCHED();
    return PG_OK;
}

int launch_periodic_conv(const double *in, int64_t T, int64_t A0, int64_t A1, int axis, const int32_t *off, const double *w,
                         int n_taps, double *out, cudaStream_t st) {
    periodic_conv_kernel<<<grid_of(T * A0 * A1), 256, 0, st>>>(in, T, A0, A1, axis, off, w, n_taps, out);
    PG_LAUNCHED();
    return PG_OK;
}

}  // namespace pg
