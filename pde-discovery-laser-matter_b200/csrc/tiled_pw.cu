// K1 tiled, pointwise rows: fused FD + library + Gram where EVERY grid point is a row (block (1,1,1)).
// Serves the full-grid configurations: basic_usage (basic:32-101,123-124: Theta over all interior points,
// p = 6) and the KS dialect with every point a row and time-holdout folds (SURVEY 8d, C4 i/ii).
//
// Data movement is the one of tiled.cu: a persistent CTA owns a TI x 128 tile, each frame's (TI+4) x 128
// row-halo tile arrives by one 3-D TMA copy, halo columns as 16-byte cells.  Differences:
//   * u_t is pointwise here, so frame t+1 must be resident while frame t is differentiated: a 4-stage ring
//     (current, next, two in flight), TI = 48 (4 x 54 KB), same swizzled stage layout and lane addressing.
//   * every point contributes an outer product, so the kernel is fp64-ISSUE bound, not HBM bound
//     (B200: 64 DFMA/clk/SM; ~32 fp64 ops per point for the p = 3 KS library, ~35 for basic p = 6,
//     against ~23 available per point at the HBM roof).  The design therefore minimises fp64 operations:
//     - all terms are accumulated UNSCALED (differences / stencil sums without their 1/h factors) and the
//       statistics are scaled once when a warp flushes its accumulators;
//     - the '1' column and the duplicates it creates in the statistics vector (sum theta_j == G_0j) are
//       not accumulated twice: the accumulators are the unique pairs of [1, y, non-constant columns];
//     - each lane keeps the whole set of accumulators in registers (no staging, no shuffles per point);
//     - n is counted in integers.
//   * folds are time-holdout folds (one id per frame): a warp accumulates for ONE fold at a time and
//     flushes its registers to its private partial slot when the fold of the next frame differs, so any
//     number of folds costs nothing per point.
//   * rows with a non-finite value: the reference drops them (ks2d:1633-1636).  Testing every point would
//     cost issue slots, so the kernel accumulates unconditionally and raises counters[2] when a flushed
//     accumulator is not finite; the API then lets the generic kernel (exact drop semantics) redo the
//     region, and the reduction ignores this kernel's partials.
//   * basic_usage rows are the interior [2:-2, 2:-2] of each frame: boundary tiles mask rows (warp-uniform)
//     and columns (selects, only instantiated for tiles that touch the left / right border).  The KS
//     dialect wraps periodically; a ragged last tile row (A0 not a multiple of 48) is handled here.
//   * any even width >= 128: when the width is not a multiple of 128 the last tile column is shifted left so that
//     it ends with the frame and masks the columns its neighbour has already counted (second tensor map for
//     the box start inside a 16-column group); odd widths (rows not 16-byte aligned) use the generic kernel.
#include <cuda.h>
#include <math.h>

#include <utility>

#include "common.cuh"
#include "launch.h"
#include "tiled_common.cuh"

namespace pg {

constexpr int PW_NSTAGE = 4;

template <int R_, int NW_> struct GeoPw {
    static constexpr int R = R_, NW = NW_;
    static constexpr int TI = R * NW;                 // tile rows: one R-row band per warp
    static constexpr int HR = TI + 4;
    static constexpr int HOFF = HR * TJ;
    static constexpr int STAGE_BYTES = stage_bytes_for(HR);       // tile + halo-column cells, 1024-byte multiple
    static constexpr int STAGE_DOUBLES = STAGE_BYTES / 8;
    static constexpr int TMA_BYTES = HR * TJ * 8;
    static constexpr int THREADS = 32 * NW;
    static constexpr size_t SMEM = (size_t)PW_NSTAGE * STAGE_BYTES + 128 + 1024;   // + mbarriers + alignment slack
};

struct PwParams {
    const double *U;
    int64_t T, A0, A1;
    double rho, kappa;        // L' = rho*(u[i+1]+u[i-1]) + (u[j+1]+u[j-1]) + kappa*u ; lap = r1*L'
    double r1, q1;            // 1/d1^2 ; 1/(2 d1)^2   (|grad|^2 = q1*(rho*dx^2 + dy^2))
    double h0, h1;            // 1/(2 d0), 1/(2 d1)
    double rdt;
    int n_tiles0, n_tiles1, n_chunks, chunk_frames;
    int64_t n_row_frames;     // T - 1
    const int32_t *fold_of_frame;
    int n_folds;
    double *partials;         // [gridDim.x*NW][n_folds][S]
    unsigned long long *counters;
};

// ----------------------------------------------------------------------------- accumulator layout
// Unique extended row: index 0 = the constant one (implicit), 1 = y, then the library columns in order
// without the '1' column.  Accumulators = upper-triangular pairs (a <= b) except (0,0), which is n.
template <int LIB> struct Pw {
    static constexpr bool KS = LIB != PG_LIB_BASIC;
    static constexpr int P = Lib<LIB>::P;
    static constexpr bool ONE = LIB == PG_LIB_KS_RICH || LIB == PG_LIB_KS_RICH_NOADV || LIB == PG_LIB_BASIC;
    static constexpr int NU = 2 + P - (ONE ? 1 : 0);
    static constexpr int NX = NU - 1;                  // values formed per point
    static constexpr int NACC = NU * (NU + 1) / 2 - 1;
    static constexpr int S = PG_STATS_LEN(P);
};

__host__ __device__ constexpr int pw_slot(int nu, int a, int b) { return a * nu - a * (a - 1) / 2 + (b - a) - 1; }

// statistics entry e -> unique pair; same enumeration as stats_pair (common.cuh)
struct PwPair { int a, b; };
__host__ __device__ constexpr PwPair pw_entry(int e, int p, bool one) {
    int a = 0, b = 0;
    if (e == 0) { a = 0; b = 0; }
    else if (e == 1) { a = 0; b = 1; }
    else if (e == 2) { a = 1; b = 1; }
    else if (e < 3 + p) { a = 0; b = 2 + (e - 3); }
    else if (e < 3 + 2 * p) { a = 1; b = 2 + (e - 3 - p); }
    else {
        int k = e - 3 - 2 * p, i = 0;
        while (k >= p - i) { k -= p - i; ++i; }
        a = 2 + i; b = 2 + i + k;
    }
    // full extended index (0 one, 1 y, 2+k theta_k) -> unique index
    const int ua = a < 2 ? a : (one ? (a == 2 ? 0 : a - 1) : a);
    const int ub = b < 2 ? b : (one ? (b == 2 ? 0 : b - 1) : b);
    return ua <= ub ? PwPair{ua, ub} : PwPair{ub, ua};
}

// Columns that are products of two other columns make a linear sum equal to a pair sum: sum(u^2 * 1) == sum(u * u),
// sum((u L') * 1) == sum(u * L').  pw_dup(b) = the pair (a1, a2) whose accumulator already holds the sum of unique
// column b (a1 = 0: none); such linear sums are not accumulated and pw_emit reads the pair's accumulator instead.
template <int LIB> __host__ __device__ constexpr PwPair pw_dup(int b) {
    if (LIB == PG_LIB_BASIC) return b == 6 ? PwPair{2, 2} : PwPair{0, 0};                                  // u^2
    if (LIB == PG_LIB_KS_RICH) return b == 3 ? PwPair{2, 2} : (b == 9 ? PwPair{2, 6} : PwPair{0, 0});       // u^2, u L'
    if (LIB == PG_LIB_KS_RICH_NOADV) return b == 3 ? PwPair{2, 2} : (b == 7 ? PwPair{2, 4} : PwPair{0, 0});
    return PwPair{0, 0};
}

// KS dialect: the grid is periodic and every time fold holds WHOLE frames, so the sum over a fold of a column that is
// the output of a difference stencil (lap, bih = lap o lap, u_x, u_y) is identically zero: each grid value enters with
// weights that add up to zero.  (The reference's own value is the rounding noise of that sum.)  Such linear sums are
// not accumulated and are emitted as 0.  pw_zero(b): unique column b is one of them.
template <int LIB> __host__ __device__ constexpr bool pw_zero(int b) {
    if (LIB == PG_LIB_KS_TRUE) return b == 2 || b == 3;                    // lap, bih
    if (LIB == PG_LIB_KS_TRUE_ADV) return b == 2 || b == 3 || b == 5 || b == 6;   // + u_x, u_y
    if (LIB == PG_LIB_KS_RICH) return b >= 4 && b <= 7;                    // u_x, u_y, lap, bih
    if (LIB == PG_LIB_KS_RICH_NOADV) return b == 4 || b == 5;              // lap, bih
    // basic_usage: not zero, but a BOUNDARY term of the trimmed interior (pw_external below): not accumulated here
    if (LIB == PG_LIB_BASIC) return b == 3 || b == 4 || b == 5;             // u_x, u_y, lap
    return false;
}

// basic_usage dialect: the rows are the interior [2:-2, 2:-2] of every frame, and some sums over it telescope to boundary
// terms (summation by parts; j along a1, i along a0, W = A1, H = A0, sums over the interior of the other axis):
//   sum (u[j+1] - u[j-1])          = u[W-2] + u[W-3] - u[1] - u[2]                 (u_x; the same along a0 for u_y)
//   sum (u[j+1] - 2 u[j] + u[j-1]) = (u[W-2] - u[W-3]) - (u[2] - u[1])             (the two halves of lap)
//   sum u[j] (u[j+1] - u[j-1])     = u[W-3] u[W-2] - u[1] u[2]                      (u u_x; the same for u u_y)
// So five of the 34 fp64 operations per point (the linear sums of u_x, u_y, lap and the products u u_x, u u_y) are not
// accumulated by the kernel; basic_boundary_kernel forms them from two rows / columns on each side of every frame (1 %
// of its bytes) and basic_external_add_kernel adds them to the reduced statistics.  pw_external(a, b): index 0..4 of
// that quantity for the unique pair (a, b), else -1.
__host__ __device__ constexpr int pw_external_basic(int a, int b) {
    if (a == 0 && b == 3) return 0;   // sum u_x
    if (a == 0 && b == 4) return 1;   // sum u_y
    if (a == 0 && b == 5) return 2;   // sum lap
    if (a == 2 && b == 3) return 3;   // sum u u_x
    if (a == 2 && b == 4) return 4;   // sum u u_y
    return -1;
}
template <int LIB> __host__ __device__ constexpr bool pw_skip_pair(int a, int b) {
    return LIB == PG_LIB_BASIC && pw_external_basic(a, b) >= 3;
}

template <int LIB, int NU, int NACC> __device__ __forceinline__ void pw_accumulate(double (&acc)[NACC], const double (&x)[NU - 1]) {
    int k = 0;
#pragma unroll
    for (int b = 1; b < NU; ++b, ++k)
        if (pw_dup<LIB>(b).a == 0 && !pw_zero<LIB>(b)) acc[k] += x[b - 1];
#pragma unroll
    for (int a = 1; a < NU; ++a)
#pragma unroll
        for (int b = a; b < NU; ++b) {
            if (!pw_skip_pair<LIB>(a, b)) acc[k] = fma(x[a - 1], x[b - 1], acc[k]);
            ++k;
        }
}

// scale of each unique entry (the factor its unscaled accumulation lacks)
template <int LIB> __device__ __forceinline__ void pw_scales(const PwParams &P, double (&sc)[Pw<LIB>::NU]) {
    sc[0] = 1.0;
    sc[1] = P.rdt;
    const double lap = P.r1, bih = P.r1 * P.r1, g = P.q1;
    if constexpr (LIB == PG_LIB_KS_TRUE) { sc[2] = lap; sc[3] = bih; sc[4] = g; }
    else if constexpr (LIB == PG_LIB_KS_TRUE_ADV) { sc[2] = lap; sc[3] = bih; sc[4] = g; sc[5] = P.h0; sc[6] = P.h1; }
    else if constexpr (LIB == PG_LIB_KS_RICH) {
        sc[2] = 1.0; sc[3] = 1.0; sc[4] = P.h0; sc[5] = P.h1; sc[6] = lap; sc[7] = bih; sc[8] = g; sc[9] = lap;
    } else if constexpr (LIB == PG_LIB_KS_RICH_NOADV) {
        sc[2] = 1.0; sc[3] = 1.0; sc[4] = lap; sc[5] = bih; sc[6] = g; sc[7] = lap;
    } else {  // BASIC: 1, u, u_x (a1), u_y (a0), lap, u^2
        sc[2] = 1.0; sc[3] = P.h1; sc[4] = P.h0; sc[5] = lap; sc[6] = 1.0;
    }
}

// ----------------------------------------------------------------------------- one frame of one warp band
// st = current frame's stage, stn = next frame's stage (u_t).  Band rows s = 0..R+3 are stage rows
// band*R + s; the warp's own rows are s = 2..R+1 (bit r of rowmask = own row r is a row of the data set).
// ROWS_ALL: every own row is a row of the data set (no per-row branch: the whole march is one basic block).
// RHO1: square grid cells (d0 == d1, rho == 1): |grad u|^2 needs no multiplication by rho.
template <int LIB, int R, bool MASKED, bool ROWS_ALL, bool RHO1 = false>
__device__ __forceinline__ void march_pw(const double *__restrict__ st, const double *__restrict__ stn, const LaneMap &m,
                                         const PwParams &P, unsigned rowmask, unsigned colmask,
                                         double (&acc)[Pw<LIB>::NACC], unsigned &cnt) {
    using X_ = Pw<LIB>;
    constexpr int NU = X_::NU, NX = X_::NX, NACC = X_::NACC;
    const unsigned ncol = MASKED ? __popc(colmask) : 4u;
    if constexpr (X_::KS) {
        double u[R + 4][8];
        double L[R + 4][6];   // L[s][k]: L' of band row s at window column k+1 (own-1 .. own+4)
#pragma unroll
        for (int s = 0; s < R + 4; ++s) {
            if (s == 0 || s == R + 3) {
                double w4[4];
                load_row4(st, m, s, w4);
#pragma unroll
                for (int c = 0; c < 4; ++c) u[s][c + 2] = w4[c];
            } else {
                load_row8(st, m, s, u[s]);
            }
            if (s >= 2) {
                const int r = s - 1;   // L' of band row r (rows 1 and R+2 only feed the a0-neighbours: own columns)
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    if ((r == 1 || r == R + 2) && (k == 0 || k == 5)) continue;
                    const int q = k + 1;
                    L[r][k] = fma(P.kappa, u[r][q], fma(P.rho, u[r + 1][q] + u[r - 1][q], u[r][q + 1] + u[r][q - 1]));
                }
            }
            if (s >= 4) {
                const int r = s - 2;   // own row r: all of its neighbours' L' are known now
                if (ROWS_ALL || ((rowmask >> (r - 2)) & 1u)) {
                    double nx[4];
                    load_row4(stn, m, r, nx);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int q = c + 2, k = c + 1;
                        const double uc = u[r][q];
                        const double Y = nx[c] - uc;
                        const double dx = u[r + 1][q] - u[r - 1][q];
                        const double dy = u[r][q + 1] - u[r][q - 1];
                        const double Gq = RHO1 ? fma(dx, dx, dy * dy) : fma(P.rho * dx, dx, dy * dy);
                        const double Lc = L[r][k];
                        const double B = fma(P.kappa, Lc, fma(P.rho, L[r + 1][k] + L[r - 1][k], L[r][k + 1] + L[r][k - 1]));
                        double x[NX];
                        if constexpr (LIB == PG_LIB_KS_TRUE) { x[0] = Y; x[1] = Lc; x[2] = B; x[3] = Gq; }
                        else if constexpr (LIB == PG_LIB_KS_TRUE_ADV) { x[0] = Y; x[1] = Lc; x[2] = B; x[3] = Gq; x[4] = dx; x[5] = dy; }
                        else if constexpr (LIB == PG_LIB_KS_RICH) {
                            x[0] = Y; x[1] = uc; x[2] = uc * uc; x[3] = dx; x[4] = dy; x[5] = Lc; x[6] = B; x[7] = Gq; x[8] = uc * Lc;
                        } else {
                            x[0] = Y; x[1] = uc; x[2] = uc * uc; x[3] = Lc; x[4] = B; x[5] = Gq; x[6] = uc * Lc;
                        }
                        if constexpr (MASKED) {
                            const bool ok = (colmask >> c) & 1u;
#pragma unroll
                            for (int k2 = 0; k2 < NX; ++k2) x[k2] = ok ? x[k2] : 0.0;
                        }
                        pw_accumulate<LIB, NU, NACC>(acc, x);
                    }
                    cnt += ncol;
                }
            }
        }
    } else {
        // basic_usage: 5-point stencils only; band rows 1 .. R+2 are needed
        double u[R + 4][8];
#pragma unroll
        for (int s = 1; s < R + 3; ++s) {
            if (s == 1 || s == R + 2) {
                double w4[4];
                load_row4(st, m, s, w4);
#pragma unroll
                for (int c = 0; c < 4; ++c) u[s][c + 2] = w4[c];
            } else {
                load_row8(st, m, s, u[s]);
            }
            if (s >= 3) {
                const int r = s - 1;
                if (ROWS_ALL || ((rowmask >> (r - 2)) & 1u)) {
                    double nx[4];
                    load_row4(stn, m, r, nx);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int q = c + 2;
                        double uc = u[r][q];
                        double Y = nx[c] - uc;
                        double dxj = u[r][q + 1] - u[r][q - 1];          // u_x * 2 d1   (basic:58)
                        double dxi = u[r + 1][q] - u[r - 1][q];          // u_y * 2 d0   (basic:59)
                        double Lc = fma(P.kappa, uc, fma(P.rho, u[r + 1][q] + u[r - 1][q], u[r][q + 1] + u[r][q - 1]));
                        if constexpr (MASKED) {
                            const bool ok = (colmask >> c) & 1u;
                            uc = ok ? uc : 0.0; Y = ok ? Y : 0.0; dxj = ok ? dxj : 0.0; dxi = ok ? dxi : 0.0; Lc = ok ? Lc : 0.0;
                        }
                        const double x[NX] = {Y, uc, dxj, dxi, Lc, uc * uc};
                        pw_accumulate<LIB, NU, NACC>(acc, x);
                    }
                    cnt += ncol;
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------- flush
template <int LIB, int E> __device__ __forceinline__ void pw_emit(const double (&acc)[Pw<LIB>::NACC], const double (&sc)[Pw<LIB>::NU],
                                                                  double n, int lane, double *out) {
    using X_ = Pw<LIB>;
    constexpr PwPair pr = pw_entry(E, X_::P, X_::ONE);
    if (lane == (E & 31)) {
        double v;
        if constexpr (pr.a == 0 && pr.b == 0) v = n;
        else if constexpr (pr.a == 0 && pw_zero<LIB>(pr.b)) v = 0.0;      // periodic sum of a difference stencil / boundary term
        else if constexpr (pw_skip_pair<LIB>(pr.a, pr.b)) v = 0.0;        // boundary term (basic_usage)
        else if constexpr (pr.a == 0 && pw_dup<LIB>(pr.b).a != 0) {
            constexpr PwPair d = pw_dup<LIB>(pr.b);     // linear sum of a product column: held by the pair's accumulator
            v = acc[pw_slot(X_::NU, d.a, d.b)] * (sc[d.a] * sc[d.b]);
        }
        else v = acc[pw_slot(X_::NU, pr.a, pr.b)] * (sc[pr.a] * sc[pr.b]);
        out[E] += v;
    }
}
template <int LIB, int... Es>
__device__ __forceinline__ void pw_emit_all(std::integer_sequence<int, Es...>, const double (&acc)[Pw<LIB>::NACC],
                                            const double (&sc)[Pw<LIB>::NU], double n, int lane, double *out) {
    (pw_emit<LIB, Es>(acc, sc, n, lane, out), ...);
}

// Warp-reduce the lane accumulators, scale, and add them to this warp's partial slot of `fold`.
// Returns true if some accumulator was not finite (a row the reference would have dropped, or overflow).
template <int LIB>
__device__ __forceinline__ bool pw_flush(double (&acc)[Pw<LIB>::NACC], unsigned &cnt, const PwParams &P, int lane, double *out) {
    using X_ = Pw<LIB>;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < X_::NACC; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[k] = v;
        bad = bad || !isfinite(v);
    }
    unsigned c = cnt;
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    double sc[X_::NU];
    pw_scales<LIB>(P, sc);
    pw_emit_all<LIB>(std::make_integer_sequence<int, X_::S>{}, acc, sc, (double)c, lane, out);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < X_::NACC; ++k) acc[k] = 0.0;
    cnt = 0;
    return bad;
}

// The kernel.  Warps are DECOUPLED: there is no block-wide barrier per frame and no producer warp.
// full[s] (TMA complete_tx) says stage s holds its frame; empty[s] counts the NW warps that are done with
// it, and the warp whose arrival completes that phase re-arms the stage with the load four frames ahead,
// so warps drift apart by up to a frame and one warp's bookkeeping overlaps the others' fp64 work.  Each warp
// copies the 16-byte side cells its own window needs (halo columns of its R+4 rows; the periodic wrap
// rows if they fall inside its window) with cp.async, one frame ahead, so they need warp-level visibility
// only.
// WS (warp-specialised, see k1_tiled_b88): a third warpgroup is launched; its first warp is the producer (stage wait,
// TMA load, the halo-column cells of all bands, tied to the `full` barrier), the band warps keep only the wrap rows of
// border tiles and the march.  setmaxnreg moves the producer warpgroup's registers to the consumers.
constexpr int PW_WS_PRODUCER_REGS = 56, PW_WS_CONSUMER_REGS = 224;   // 2 x 128 x 224 + 128 x 56 = 64512 <= 65536

template <int LIB, int R, int NW, bool WS = false>
__global__ void __launch_bounds__(32 * (NW + (WS ? 4 : 0)), 1) k1_tiled_pw(const __grid_constant__ CUtensorMap tmap,
                                                                           const __grid_constant__ CUtensorMap tmap_last,
                                                                           PwParams P) {
    using G_ = GeoPw<R, NW>;
    using X_ = Pw<LIB>;
    constexpr int TI = G_::TI, HOFF = G_::HOFF, STAGE_DOUBLES = G_::STAGE_DOUBLES, NS = PW_NSTAGE;
    constexpr int S = X_::S;
    constexpr bool KS = X_::KS;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem_raw = align1024(smem_dyn);
    double *stages = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + NS * G_::STAGE_BYTES);
    uint64_t *empty = full + NS;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], WS ? 33 : 1); mbar_init(&empty[s], NW); }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    const int64_t frame = P.A0 * P.A1;
    const int n_tiles = P.n_tiles0 * P.n_tiles1;
    const int64_t n_items = (int64_t)n_tiles * P.n_chunks;

    auto geometry = [&](int64_t item, int &i0, int &j0, int &t0, int &nf) {
        const int tile = (int)(item % n_tiles), chunk = (int)(item / n_tiles);
        i0 = (tile / P.n_tiles1) * TI;
        j0 = min((tile % P.n_tiles1) * TJ, (int)P.A1 - TJ);   // a width that is not a multiple of 128: the last tile column is shifted left
        t0 = chunk * P.chunk_frames;
        nf = (int)min((int64_t)P.chunk_frames, P.n_row_frames - t0);
    };

    // ---- loads: no warp is the producer.  Every warp releases a stage when it has read what it needs; the
    // warp whose arrival completes the phase re-arms the stage with the load NS frames ahead.
    auto issue_load = [&](uint32_t s, int i0, int j0, int t) {
        fence_proxy_async();
        mbar_expect_tx(&full[s], G_::TMA_BYTES);
        // a shifted tile column may start inside a 16-column group: tmap_last views the field from column A1 % 16
        tma_load_4d(stages + s * STAGE_DOUBLES, (j0 & 15) ? &tmap_last : &tmap, &full[s], 0, j0 >> 4, i0 - 2, t);
    };
    // coordinates of the load `ahead` frames after frame f of `item`; walks into the following items of this
    // CTA; false when the CTA's stream of frames ends before that
    auto ahead_coords = [&](int64_t item, int i0, int j0, int t0, int nf, int f, int ahead, int &ai0, int &aj0, int &at) {
        int rem = f + ahead;
        while (rem > nf) {
            rem -= nf + 1;
            item += gridDim.x;
            if (item >= n_items) return false;
            geometry(item, i0, j0, t0, nf);
        }
        ai0 = i0; aj0 = j0; at = t0 + rem;
        return true;
    };
    if (!WS && tid == 0 && (int64_t)blockIdx.x < n_items) {
        int i0, j0, t0, nf;
        geometry(blockIdx.x, i0, j0, t0, nf);
        for (int a = 0; a < NS; ++a) {
            int ai0, aj0, at;
            if (ahead_coords(blockIdx.x, i0, j0, t0, nf, 0, a, ai0, aj0, at)) issue_load(a, ai0, aj0, at);
        }
    }
    if constexpr (WS) {
        if (warp >= NW) {
            // ---- producer warpgroup: give the registers back, one warp feeds the ring
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PW_WS_PRODUCER_REGS));
            if (warp > NW) return;
            constexpr int NC = (2 * G_::HR + 31) / 32;     // halo-column cells per lane and frame
            uint32_t Gp = 0, s = 0, ph = 0;
            for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
                int i0, j0, t0, nf;
                geometry(item, i0, j0, t0, nf);
                int hoff[NC];
                int64_t hsrc[NC];          // < 0: outside the frame (basic_usage) -> zero-filled
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    const int c = lane + 32 * k, Rr = c >> 1, side = c & 1;
                    const int C = side ? j0 + TJ : j0 - 2;
                    const int64_t gi = (int64_t)i0 - 2 + Rr;
                    hoff[k] = c < 2 * G_::HR ? HOFF + Rr * 4 + side * 2 : -1;
                    if constexpr (KS) hsrc[k] = wrap(gi, P.A0) * P.A1 + wrap((int64_t)C, P.A1);
                    else hsrc[k] = (gi >= 0 && gi < P.A0 && C >= 0 && C < P.A1) ? gi * P.A1 + C : -1;
                }
                const double *Ft = P.U + (int64_t)t0 * frame;
                for (int f = 0; f <= nf; ++f, ++Gp, Ft += frame) {
                    if (Gp >= NS) mbar_wait(&empty[s], ph ^ 1);   // every band warp has released load Gp - NS
                    if (lane == 0) issue_load(s, i0, j0, t0 + f);
                    if (f < nf) {                                // the frame after a chunk only feeds u_t: own columns
                        double *stg = stages + s * STAGE_DOUBLES;
#pragma unroll
                        for (int k = 0; k < NC; ++k)
                            if (hoff[k] >= 0) cp_async16(stg + hoff[k], hsrc[k] >= 0 ? Ft + hsrc[k] : P.U, hsrc[k] < 0);
                    }
                    cp_async_mbar_arrive_noinc(&full[s]);
                    if (++s == NS) { s = 0; ph ^= 1; }
                }
            }
            cp_async_wait_all();
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PW_WS_CONSUMER_REGS));
    }

    const LaneMap lm = make_lane_map(warp * R, HOFF, lane);

    double acc[X_::NACC];
#pragma unroll
    for (int k = 0; k < X_::NACC; ++k) acc[k] = 0.0;
    unsigned cnt = 0;           // rows accumulated by this lane since the last flush
    int cur_fold = -1;
    bool poisoned = false;
    unsigned long long bad_fold = 0;

    double *slot = P.partials + ((int64_t)blockIdx.x * NW + warp) * P.n_folds * S;
    for (int e = lane; e < P.n_folds * S; e += 32) slot[e] = 0.0;
    __syncwarp();

    uint32_t G = 0;  // consumer load index (stage = G % NS, parity = (G / NS) & 1)
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        int i0, j0, t0, nf;
        geometry(item, i0, j0, t0, nf);
        const int vrows = (int)min((int64_t)TI, P.A0 - i0);      // tile rows inside the frame
        // wb: one of the two halo rows below the tile lies beyond the frame (also A0 = 48 k + 1: the frame ends one row
        // below a whole tile)
        const bool wt = KS && i0 == 0, wb = KS && (int64_t)i0 + vrows + 2 > P.A0;

        // rows / columns of this warp / lane that are rows of the data set
        unsigned rowmask = 0, colmask = 0;
        // a shifted last tile column skips the columns its left neighbour has already counted
        const bool shifted = (j0 & (TJ - 1)) != 0;
        const int64_t jmin = shifted ? (int64_t)(P.n_tiles1 - 1) * TJ : 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t i = (int64_t)i0 + warp * R + r;
            const bool ok = KS ? (i < P.A0) : (i >= 2 && i < P.A0 - 2);
            rowmask |= ok ? (1u << r) : 0u;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int64_t j = (int64_t)j0 + 4 * lm.g + c;
            const bool ok = (KS ? true : (j >= 2 && j < P.A1 - 2)) && j >= jmin;
            colmask |= ok ? (1u << c) : 0u;
        }
        const bool edge_cols = shifted || (!KS && (j0 < 2 || (int64_t)j0 + TJ > P.A1 - 2));   // CTA-uniform
        const unsigned rows_per_frame = (unsigned)__popc(rowmask) * (unsigned)__popc(colmask);

        // Side cells of this warp's window (stage rows warp*R .. warp*R + R+3):
        //   lanes 0 .. 2(R+4)-1: one halo-column cell each (left / right pair of one row);
        //   KS tiles at the top / bottom of the periodic domain: the wrap rows TMA zero-filled (stage rows
        //   0, 1 and vrows+2, vrows+3), 64 cells per row = 2 per lane, if the row lies in the window.
        int h_off = -1;
        int64_t h_src = -1;       // < 0: outside the frame (basic_usage) -> zero-filled
        if (lane < 2 * (R + 4)) {
            const int Rr = warp * R + (lane >> 1), side = lane & 1;
            const int C = side ? j0 + TJ : j0 - 2;
            const int64_t gi = (int64_t)i0 - 2 + Rr;
            h_off = HOFF + Rr * 4 + side * 2;
            if constexpr (KS) h_src = wrap(gi, P.A0) * P.A1 + wrap((int64_t)C, P.A1);
            else if (gi >= 0 && gi < P.A0 && C >= 0 && C < P.A1) h_src = gi * P.A1 + C;
        }
        int w_row[4];             // stage rows to patch (-1: none), warp-uniform
        int64_t w_src[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int Rr = k < 2 ? k : vrows + k;           // 0, 1, vrows+2, vrows+3
            const bool need = (k < 2 ? wt : wb) && Rr >= warp * R && Rr <= warp * R + R + 3;
            w_row[k] = need ? Rr : -1;
            w_src[k] = wrap((int64_t)i0 - 2 + Rr, P.A0) * P.A1 + j0;
        }
        const bool has_wrap = KS && (w_row[0] >= 0 || w_row[1] >= 0 || w_row[2] >= 0 || w_row[3] >= 0);   // warp-uniform
        auto issue_cells = [&](double *stage, int64_t t) {
            const double *Ft = P.U + t * frame;
            if (!WS && h_off >= 0) cp_async16(stage + h_off, h_src >= 0 ? Ft + h_src : P.U, h_src < 0);
            if constexpr (KS) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (w_row[k] >= 0) {
                        cp_async16(stage + w_row[k] * TJ + swz_cell(lane), Ft + w_src[k] + 2 * lane, false);
                        cp_async16(stage + w_row[k] * TJ + swz_cell(lane + 32), Ft + w_src[k] + 64 + 2 * lane, false);
                    }
            }
            cp_async_commit();
        };

        int fold_next = P.fold_of_frame ? __ldg(P.fold_of_frame + t0) : 0;
        mbar_wait(&full[G % NS], (G / NS) & 1);
        if (!WS || has_wrap) issue_cells(stages + (G % NS) * STAGE_DOUBLES, t0);

        for (int f = 0; f <= nf; ++f, ++G) {
            const double *st = stages + (G % NS) * STAGE_DOUBLES;
            // where this frame's stage goes next, known up front so that re-arming it costs no arithmetic
            int n_i0 = i0, n_j0 = j0, n_t = t0 + f + NS;
            bool n_ok = true;
            if constexpr (!WS)
                if (f + NS > nf) n_ok = ahead_coords(item, i0, j0, t0, nf, f, NS, n_i0, n_j0, n_t);
            if (f < nf) {
                double *stn = stages + ((G + 1) % NS) * STAGE_DOUBLES;
                mbar_wait(&full[(G + 1) % NS], ((G + 1) / NS) & 1);   // frame t+1: u_t now, differentiated next
                if (!WS || has_wrap) {
                    if (f + 1 < nf) issue_cells(stn, (int64_t)t0 + f + 1);
                    else cp_async_commit();
                    cp_async_wait<1>();                                // this frame's cells have landed
                }
                __syncwarp();
                const int fold = fold_next;
                if (f + 1 < nf && P.fold_of_frame) fold_next = __ldg(P.fold_of_frame + t0 + f + 1);
                if (fold < 0 || fold >= P.n_folds) {
                    if (fold >= 0) bad_fold += rows_per_frame;          // a negative id excludes the frame on purpose
                } else {
                    if (fold != cur_fold) {
                        if (cur_fold >= 0) poisoned |= pw_flush<LIB>(acc, cnt, P, lane, slot + cur_fold * S);
                        cur_fold = fold;
                    }
                    const bool rows_all = rowmask == (1u << R) - 1u;   // warp-uniform
                    if constexpr (KS) {
                        if (edge_cols) march_pw<LIB, R, true, false>(st, stn, lm, P, rowmask, colmask, acc, cnt);
                        else if (rows_all && P.rho == 1.0) march_pw<LIB, R, false, true, true>(st, stn, lm, P, rowmask, colmask, acc, cnt);
                        else if (rows_all) march_pw<LIB, R, false, true>(st, stn, lm, P, rowmask, colmask, acc, cnt);
                        else march_pw<LIB, R, false, false>(st, stn, lm, P, rowmask, colmask, acc, cnt);
                    } else {
                        if (!edge_cols && rows_all) march_pw<LIB, R, false, true>(st, stn, lm, P, rowmask, colmask, acc, cnt);
                        else if (!edge_cols) march_pw<LIB, R, false, false>(st, stn, lm, P, rowmask, colmask, acc, cnt);
                        else march_pw<LIB, R, true, false>(st, stn, lm, P, rowmask, colmask, acc, cnt);
                    }
                }
            }
            // release the stage: this warp has read everything it needs from load G (only the wrap rows are
            // generic-proxy writes inside the TMA box: their writers fence towards the async proxy)
            if (has_wrap) fence_proxy_async();
            __syncwarp();
            if constexpr (WS) {
                if (lane == 0) mbar_arrive(&empty[G % NS]);
            } else if (lane == 0 && mbar_arrive_pending(&empty[G % NS]) == 1 && n_ok) issue_load(G % NS, n_i0, n_j0, n_t);
        }
    }
    cp_async_wait<0>();
    if (cur_fold >= 0) poisoned |= pw_flush<LIB>(acc, cnt, P, lane, slot + cur_fold * S);
    if (bad_fold) atomicAdd(&P.counters[1], bad_fold);
    if (poisoned && lane == 0) atomicAdd(&P.counters[2], 1ull);
}

// ----------------------------------------------------------------------------- basic_usage boundary terms
// Per frame t (one CTA): the five unscaled boundary sums of pw_external_basic, reduced in a fixed order.
__global__ void __launch_bounds__(256) basic_boundary_kernel(const double *__restrict__ U, int64_t A0, int64_t A1, double rho,
                                                            double *__restrict__ out /* [T-1][5] */) {
    const double *F = U + (int64_t)blockIdx.x * A0 * A1;
    double v[5] = {0, 0, 0, 0, 0};
    // rows i = 2 .. A0-3: columns 1, 2, W-3, W-2
    for (int64_t i = 2 + threadIdx.x; i < A0 - 2; i += blockDim.x) {
        const double *R = F + i * A1;
        const double c1 = R[1], c2 = R[2], c3 = R[A1 - 3], c4 = R[A1 - 2];
        v[0] += (c4 + c3) - (c1 + c2);
        v[2] += (c4 - c3) - (c2 - c1);
        v[3] += c3 * c4 - c1 * c2;
    }
    // columns j = 2 .. A1-3: rows 1, 2, H-3, H-2
    for (int64_t j = 2 + threadIdx.x; j < A1 - 2; j += blockDim.x) {
        const double r1 = F[A1 + j], r2 = F[2 * A1 + j], r3 = F[(A0 - 3) * A1 + j], r4 = F[(A0 - 2) * A1 + j];
        v[1] += (r4 + r3) - (r1 + r2);
        v[2] = fma(rho, (r4 - r3) - (r2 - r1), v[2]);
        v[4] += r3 * r4 - r1 * r2;
    }
    __shared__ double sh[5][256];
#pragma unroll
    for (int k = 0; k < 5; ++k) sh[k][threadIdx.x] = v[k];
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w)
#pragma unroll
            for (int k = 0; k < 5; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x < 5) out[(int64_t)blockIdx.x * 5 + threadIdx.x] = sh[threadIdx.x][0];
}

// One warp per (fold, quantity): sums the frames of the fold in a fixed order and adds scale * sum to every statistics
// entry that holds the quantity.  Skipped when the tiled kernel's result is not used (*skip_if != 0: the exact fallback ran).
struct BasicExternalMap {
    int n_e[5];
    int e[5][4];
    double scale[5];
};
__global__ void basic_external_add_kernel(const double *__restrict__ per_frame, int64_t n_frames, const int32_t *__restrict__ fold_of_frame,
                                          int n_folds, int S, BasicExternalMap map, const unsigned long long *__restrict__ skip_if,
                                          double *__restrict__ stats) {
    if (skip_if && *skip_if != 0) return;
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (warp >= n_folds * 5) return;
    const int fold = warp / 5, k = warp % 5;
    double s = 0.0;
    for (int64_t t = lane; t < n_frames; t += 32)
        if ((fold_of_frame ? fold_of_frame[t] : 0) == fold) s += per_frame[t * 5 + k];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0)
        for (int q = 0; q < map.n_e[k]; ++q) stats[fold * S + map.e[k][q]] += map.scale[k] * s;
}

size_t tiled_pw_external_scratch(const K1Params &P, int lib) {
    return lib == PG_LIB_BASIC ? sizeof(double) * 5 * (size_t)(P.T - 1) : 0;
}

// after the reduction: add the boundary terms the tiled pointwise kernel leaves out (basic_usage library only)
int tiled_pw_external(const K1Params &P, int lib, double *per_frame, const unsigned long long *skip_if, double *stats_out,
                      cudaStream_t st) {
    if (lib != PG_LIB_BASIC) return PG_OK;
    const int64_t nrf = P.T - 1;
    const double rho = P.c.d1sq / P.c.d0sq;
    basic_boundary_kernel<<<(unsigned)nrf, 256, 0, st>>>(P.U, P.A0, P.A1, rho, per_frame);
    PG_LAUNCHED();
    using X_ = Pw<PG_LIB_BASIC>;
    BasicExternalMap map{};
    const double h0 = 1.0 / P.c.two_d0, h1 = 1.0 / P.c.two_d1, r1 = 1.0 / P.c.d1sq;
    const double scale[5] = {h1, h0, r1, h1, h0};      // unique columns: u_x = a1-difference * h1, u_y = a0-difference * h0
    for (int k = 0; k < 5; ++k) map.scale[k] = scale[k];
    for (int e = 0; e < X_::S; ++e) {
        const PwPair pr = pw_entry(e, X_::P, X_::ONE);
        const int k = pw_external_basic(pr.a, pr.b);
        if (k >= 0) map.e[k][map.n_e[k]++] = e;
    }
    const int warps = P.n_folds * 5;
    basic_external_add_kernel<<<(warps * 32 + 127) / 128, 128, 0, st>>>(per_frame, nrf, P.fold_of_frame, P.n_folds, X_::S, map, skip_if, stats_out);
    PG_LAUNCHED();
    return PG_OK;
}

// ----------------------------------------------------------------------------- host side
// Tile geometry: 48 rows as 8 warps x 6 rows (default) or 12 warps x 4 rows (PG_PW_GEO=1, experiments)
static int pw_geo() { return env_int("PG_PW_GEO", 0) == 1 ? 1 : 0; }

bool tiled_pw_plan(const K1Params &P, int lib, int n_sm, TiledPlan &plan) {
    if (P.bt != 1 || P.b0 != 1 || P.b1 != 1) return false;
    if (P.fold_of_row) return false;                        // per-row folds: generic kernel
    const bool ks = P.dialect == PG_FD_KS_PERIODIC;
    if (ks ? !(lib == PG_LIB_KS_TRUE || lib == PG_LIB_KS_TRUE_ADV || lib == PG_LIB_KS_RICH || lib == PG_LIB_KS_RICH_NOADV)
           : lib != PG_LIB_BASIC)
        return false;
    if (P.A1 % 2 != 0 || (reinterpret_cast<uintptr_t>(P.U) & 15)) return false;   // 16-byte aligned rows for TMA / cp.async
    if (P.T < 2 || P.T > 0x7fffffff || P.A0 > 0x7fffffff || P.A1 > 0x7fffffff) return false;
    if (P.A0 < 4 || P.A1 < TJ) return false;
    if (!encode_fn()) return false;
    constexpr int TI = 48;
    const int geo = pw_geo();
    const int NWg = geo == 1 ? 12 : 8;
    const int64_t nt0 = (P.A0 + TI - 1) / TI;
    const int64_t nt1 = (P.A1 + TJ - 1) / TJ;   // a last tile column that is not whole is shifted left and masked
    const int64_t n_tiles = nt0 * nt1, nrf = P.T - 1;
    // number of frame chunks: balance the persistent CTAs; every item pays one extra frame + pipeline fill
    int64_t best_c = 1;
    double best_cost = 1e300;
    for (int64_t c = 1; c <= nrf && c <= 4096; ++c) {
        const int64_t cf = (nrf + c - 1) / c, cc = (nrf + cf - 1) / cf;
        const int64_t rounds = (n_tiles * cc + n_sm - 1) / n_sm;
        const double cost = (double)rounds * ((double)cf + 2.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best_c = cc; }
    }
    const int64_t cf = (nrf + best_c - 1) / best_c;
    plan = TiledPlan{};
    plan.nbt = nrf;
    plan.nb0 = P.R0;
    plan.nb1 = P.R1;
    plan.chunk_t = (int)cf;
    plan.n_chunks = (nrf + cf - 1) / cf;
    plan.n_tiles0 = nt0; plan.n_tiles1 = nt1;
    const int64_t items = n_tiles * plan.n_chunks;
    plan.grid = (int)(items < n_sm ? items : n_sm);
    plan.n_parts = (int64_t)plan.grid * NWg;
    plan.extra_scratch = 0;
    plan.kernel_id = 100 + geo;
    plan.tile0 = TI; plan.tile1 = TJ;
    return true;
}

// Measured (256 x 2048^2, B200): basic_usage p = 6 2.81 -> 2.69 ms (2.97 -> 2.72 with two folds), KS advection 3.02 -> 2.99,
// KS p = 3 2.36 -> 2.39 (the march is 95 % of that kernel and it is fp64-issue bound either way), rich p = 9 slower
// (it needs more than the 224 registers a consumer gets).  Default (-1): the basic_usage dialect and KS advection only.
#ifndef PG_PW_WS_DEFAULT
#define PG_PW_WS_DEFAULT -1
#endif
template <int LIB, int R, int NW, bool WS = false>
static int launch_pw_g(const CUtensorMap (&map)[2], const PwParams &pp, int grid, cudaStream_t st) {
    using G_ = GeoPw<R, NW>;
    PG_CUDA(cudaFuncSetAttribute(k1_tiled_pw<LIB, R, NW, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_::SMEM));
    k1_tiled_pw<LIB, R, NW, WS><<<grid, G_::THREADS + (WS ? 128 : 0), G_::SMEM, st>>>(map[0], map[1], pp);
    PG_LAUNCHED();
    return PG_OK;
}
template <int LIB> static int launch_pw_t(const CUtensorMap (&map)[2], const PwParams &pp, int grid, int geo, cudaStream_t st) {
    if (geo == 1) return launch_pw_g<LIB, 4, 12>(map, pp, grid, st);
    const int ws = env_int("PG_PW_WS", PG_PW_WS_DEFAULT);
    if (ws > 0 || (ws < 0 && (LIB == PG_LIB_BASIC || LIB == PG_LIB_KS_TRUE_ADV))) return launch_pw_g<LIB, 6, 8, true>(map, pp, grid, st);
    return launch_pw_g<LIB, 6, 8>(map, pp, grid, st);
}

int tiled_pw_launch(const K1Params &P, int lib, const TiledPlan &plan, double *partials, cudaStream_t st) {
    CUtensorMap map[2];
    for (int k = 0; k < 2; ++k) {
        const CUresult r = encode_field_map(&map[k], P.U, P.T, P.A0, P.A1, 48 + 4, k ? P.A1 % 16 : 0);
        if (r != CUDA_SUCCESS) PG_FAIL(PG_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    }
    PwParams pp{};
    pp.U = P.U; pp.T = P.T; pp.A0 = P.A0; pp.A1 = P.A1;
    pp.rho = P.c.d1sq / P.c.d0sq;
    pp.kappa = -2.0 * (1.0 + pp.rho);
    pp.r1 = 1.0 / P.c.d1sq;
    pp.q1 = 1.0 / (P.c.two_d1 * P.c.two_d1);
    pp.h0 = 1.0 / P.c.two_d0; pp.h1 = 1.0 / P.c.two_d1;
    pp.rdt = 1.0 / P.c.dt;
    pp.n_tiles0 = (int)plan.n_tiles0; pp.n_tiles1 = (int)plan.n_tiles1; pp.n_chunks = (int)plan.n_chunks;
    pp.chunk_frames = plan.chunk_t;
    pp.n_row_frames = P.T - 1;
    pp.fold_of_frame = P.fold_of_frame; pp.n_folds = P.n_folds;
    pp.partials = partials; pp.counters = P.counters;
    switch (lib) {
        case PG_LIB_KS_TRUE: return launch_pw_t<PG_LIB_KS_TRUE>(map, pp, plan.grid, plan.kernel_id - 100, st);
        case PG_LIB_KS_TRUE_ADV: return launch_pw_t<PG_LIB_KS_TRUE_ADV>(map, pp, plan.grid, plan.kernel_id - 100, st);
        case PG_LIB_KS_RICH: return launch_pw_t<PG_LIB_KS_RICH>(map, pp, plan.grid, plan.kernel_id - 100, st);
        case PG_LIB_KS_RICH_NOADV: return launch_pw_t<PG_LIB_KS_RICH_NOADV>(map, pp, plan.grid, plan.kernel_id - 100, st);
        case PG_LIB_BASIC: return launch_pw_t<PG_LIB_BASIC>(map, pp, plan.grid, plan.kernel_id - 100, st);
        default: PG_FAIL(PG_EUNSUPPORTED, "no tiled pointwise kernel for library %d", lib);
    }
}

}  // namespace pg
