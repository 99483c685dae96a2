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
