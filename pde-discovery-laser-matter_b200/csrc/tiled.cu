// K1 tiled: fused FD + library + block mean + Gram for the KS dialect with (bt, 8, 8) blocks.
//
// One CTA owns a 64 x 128 spatial tile and marches through a chunk of frames.  Each frame's
// (64+4) x (128+4) halo tile is brought into shared memory ONCE by a 3-D TMA tensor copy
// (cp.async.bulk.tensor, mbarrier completion) into a 3-stage ring: stage f is the frame being
// differentiated, stages f+1 and f+2 are in flight (two loads ahead: a single outstanding 72 KB
// load per SM cannot cover the HBM latency-bandwidth product).  The forward u_t never touches
// frame t+1 pointwise: over a t-block it telescopes to (sum u(t0+bt) - sum u(t0)) / dt, and every
// frame's block sum of u is formed while that frame is current.  Periodic wrap: TMA zero-fills the out-of-bounds halo of a
// border tile; the wrapped values are prefetched with plain loads one iteration ahead and stored
// over the zero fill after the copy has landed.
//
// Inside a frame, warp w owns rows [8w, 8w+8) (one block row) and lane l owns columns
// [4l, 4l+4) (half a block), marching down 12 tile rows with a register sliding window: four
// conflict-free LDS.128 per row give u at columns own-2 .. own+5, with no exchange between
// lanes.  Block sums of the linear terms (lap, bih, u_x, u_y, u, u_t) reduce by the discrete
// divergence theorem to per-row boundary scalars (see march_frame), so only the nonlinear
// terms cost per-point fp64 work: about 8 fp64 ops per grid point for the true library.  At the end
// of a t-block the two lanes of a block pair-reduce by shuffle, form the block-mean row and the
// warp adds its 16 rows to lane-owned Gram entries held in registers.
//
// No tensor cores: p <= 9 and the kernel is HBM / fp64-issue bound.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "launch.h"

namespace pg {

constexpr int TJ = 128;                     // tile columns (32 lanes x 4)
constexpr int HC = TJ + 4;                  // halo tile columns
constexpr int NSTAGE = 3;

// Tile geometry for NW warps per CTA.  KC = columns per lane, BANDS = 8-row block bands per tile.
//   NW = 16: 64 x 128 tile, KC = 2 (each band is split between two warps), one CTA per SM,
//            <= 128 registers: 16 resident warps hide the fp64 / shared-memory latencies
//   NW = 8 : 64 x 128 tile, KC = 4, one CTA per SM
//   NW = 4 : 32 x 128 tile, KC = 4, two CTAs per SM
template <int NW> struct Geo {
    static constexpr int KC = NW == 16 ? 2 : 4;
    static constexpr int BANDS = NW == 16 ? 8 : NW;
    static constexpr int TI = 8 * BANDS;
    static constexpr int HR = TI + 4;
    static constexpr int STAGE_DOUBLES = HR * HC;
    static constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
    static constexpr int THREADS = 32 * NW;
    static constexpr int MAXWRAP = (4 * HR + 4 * HC + THREADS - 1) / THREADS;   // wrap cells per thread
    static constexpr int CTAS_PER_SM = NW == 4 ? 2 : 1;
    static constexpr int LPB = 8 / KC;          // lanes per 8-column block
    static constexpr int RPW = 32 / LPB;        // block rows a warp emits per t-block
    // block rows staged per Gram-update batch (shared memory is the scarce resource at 2 CTAs/SM)
    __host__ __device__ static constexpr int slots(int p) { return NW == 4 ? (p <= 5 ? 4 : 2) : 8; }
    __host__ __device__ static constexpr size_t smem(int p) { return (size_t)NSTAGE * STAGE_BYTES + 64 + sizeof(double) * NW * slots(p) * (p + 2); }
};

struct TiledParams {
    const double *U;
    int64_t T, A0, A1;
    double rho, kappa;        // L' = rho*(u[i+1]+u[i-1]) + (u[j+1]+u[j-1]) + kappa*u ; lap = r1*L'
    double r1;                // 1/d1^2
    double q0, q1;            // 1/(4 d0^2), 1/(4 d1^2)   (squares of the central-difference scales)
    double h0, h1;            // 1/(2 d0), 1/(2 d1)
    double rdt;               // 1/dt
    int bt;
    int n_tiles0, n_tiles1, n_chunks, chunk_tb;
    int64_t nbt;              // t-blocks covered
    int64_t nB0, nB1;         // block counts of the whole row space (row numbering)
    const uint8_t *fold_of_row;
    const int32_t *fold_of_frame;
    int n_folds;
    double *partials;         // [gridDim.x*NW][n_folds][S]
    unsigned long long *counters;
};

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ----------------------------------------------------------------------------- per-lane block sums
// Unscaled sums over the lane's half block (4 columns x 8 rows x frames); which ones are live
// depends on the library.  FrameSums holds one frame's contribution, Acc the running t-block.
struct Sums {
    double SL = 0, SE1 = 0, SE2 = 0, SGx = 0, SGy = 0;   // all libraries
    double SU = 0;                                      // sum of u: rich column, and u_t by telescoping
    double SDx = 0, SDy = 0;                            // advection columns
    double SU2 = 0, SUL = 0;                            // rich
};

template <int LIB> constexpr bool kNeedAdv = (LIB == PG_LIB_KS_TRUE_ADV || LIB == PG_LIB_KS_RICH);
template <int LIB> constexpr bool kRich = (LIB == PG_LIB_KS_RICH || LIB == PG_LIB_KS_RICH_NOADV);

// 16-byte chunk pair (k, k+1) of a lane's row segment.  Lanes are 32 B apart, so a plain LDS.128
// would hit every bank group twice per quarter-warp; lanes with bit 2 set fetch the two chunks
// in the opposite order (conflict-free) and swap them back.
__device__ __forceinline__ void load_pair(const double2 *src, int k, int sw, double2 &lo, double2 &hi) {
    const double2 x = src[k + sw], y = src[k + (sw ^ 1)];
    lo = sw ? y : x;
    hi = sw ? x : y;
}

// One frame of one warp band: 12 tile rows march through a 3-row register window.
//   cur : halo tile of frame t (row pitch HC), pointing at the warp's first tile row, lane's first column
//
// With w[0..7] the lane's row segment (columns own-2 .. own+5, own = q 2..5) and band rows
// s = 0..11 (outputs are rows 2..9), every block sum of a LINEAR term reduces, by the discrete
// divergence theorem, to a few per-row scalars:
//   rsU(s) = w2+w3+w4+w5                      D(s) = (w1-w2) + (w6-w5)        e(s) = (w0-w3) + (w7-w4)
//   rsL(s) = sum_own L'(s,.) = rho*(rsU(s+1) + rsU(s-1) - 2 rsU(s)) + D(s)
//   SL  = sum_{2..9} rsL           = rho*((rsU10 - rsU9) - (rsU2 - rsU1)) + sum_{2..9} D
//   SE1 = a0-flux of L' through the block = rsL1 - rsL2 + rsL10 - rsL9
//   SE2 = a1-flux of L' through the block = -3 sum_{2..9} D + rho*(D10 + D1 - D2 - D9) + sum_{2..9} e
// (L' = rho*(u[i+1]+u[i-1]) + (u[j+1]+u[j-1]) + kappa*u with kappa = -2(1+rho); lap = r1*L',
// block sum of bih = r1^2*(rho*SE1 + SE2)).  Only the nonlinear terms (|grad u|^2, and for the
// rich library u^2 and u*L') cost per-point fp64 work.
template <int LIB>
__device__ __forceinline__ void march_frame(const double *__restrict__ cur, const TiledParams &P, Sums &F, const int sw) {
    double wp[4], wc[8], wn[8];              // own columns of row s-2; rows s-1 and s (all 8 columns)
    double rsU[12], D[12];
    // per-column partial sums of the nonlinear terms: four independent FMA chains per quantity
    double gx[4] = {0, 0, 0, 0}, gy[4] = {0, 0, 0, 0}, u2[4] = {0, 0, 0, 0}, ul[4] = {0, 0, 0, 0};
    double sD = 0, sE = 0, sU = 0, sDy = 0;
#pragma unroll
    for (int s = 0; s < 12; ++s) {
        const double2 *src = reinterpret_cast<const double2 *>(cur + s * HC);
        double2 a0, a1, a2, a3;
        load_pair(src, 0, sw, a0, a1);
        load_pair(src, 2, sw, a2, a3);
        wn[0] = a0.x; wn[1] = a0.y; wn[2] = a1.x; wn[3] = a1.y; wn[4] = a2.x; wn[5] = a2.y; wn[6] = a3.x; wn[7] = a3.y;
        rsU[s] = (wn[2] + wn[3]) + (wn[4] + wn[5]);
        D[s] = (wn[1] - wn[2]) + (wn[6] - wn[5]);
        if (s >= 2 && s <= 9) {
            sD += D[s];
            sE += (wn[0] - wn[3]) + (wn[7] - wn[4]);
            sU += rsU[s];
            if constexpr (kNeedAdv<LIB>) sDy += (wn[6] + wn[5]) - (wn[2] + wn[1]);
        }
        if (s >= 3 && s <= 10) {
            // ---- nonlinear terms of output row s-1 (u in wc); rows s-2 (wp) and s (wn) are its a0-neighbours
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int q = c + 2;
                const double dx = wn[q] - wp[c];
                const double dy = wc[q + 1] - wc[q - 1];
                gx[c] = fma(dx, dx, gx[c]);
                gy[c] = fma(dy, dy, gy[c]);
                if constexpr (kRich<LIB>) {
                    const double Lq = fma(P.kappa, wc[q], fma(P.rho, wn[q] + wp[c], wc[q + 1] + wc[q - 1]));
                    u2[c] = fma(wc[q], wc[q], u2[c]);
                    ul[c] = fma(wc[q], Lq, ul[c]);
                }
            }
        }
        // ---- rotate the window
#pragma unroll
        for (int c = 0; c < 4; ++c) wp[c] = wc[c + 2];
#pragma unroll
        for (int q = 0; q < 8; ++q) wc[q] = wn[q];
    }
    const double rsL1 = fma(P.rho, (rsU[2] + rsU[0]) - 2.0 * rsU[1], D[1]);
    const double rsL2 = fma(P.rho, (rsU[3] + rsU[1]) - 2.0 * rsU[2], D[2]);
    const double rsL9 = fma(P.rho, (rsU[10] + rsU[8]) - 2.0 * rsU[9], D[9]);
    const double rsL10 = fma(P.rho, (rsU[11] + rsU[9]) - 2.0 * rsU[10], D[10]);
    F.SL = fma(P.rho, (rsU[10] - rsU[9]) - (rsU[2] - rsU[1]), sD);
    F.SE1 = (rsL1 - rsL2) + (rsL10 - rsL9);
    F.SE2 = fma(P.rho, (D[10] + D[1]) - (D[2] + D[9]), fma(-3.0, sD, sE));
    F.SU = sU;
    F.SGx = (gx[0] + gx[1]) + (gx[2] + gx[3]);
    F.SGy = (gy[0] + gy[1]) + (gy[2] + gy[3]);
    if constexpr (kNeedAdv<LIB>) {
        F.SDy = sDy;
        F.SDx = (rsU[10] + rsU[9]) - (rsU[2] + rsU[1]);   // sum_{2..9} (rsU(s+1) - rsU(s-1))
    }
    if constexpr (kRich<LIB>) {
        F.SU2 = (u2[0] + u2[1]) + (u2[2] + u2[3]);
        F.SUL = (ul[0] + ul[1]) + (ul[2] + ul[3]);
    }
}

// KC = 2 variant: the lane owns two columns (w[2], w[3] of the six-value segment own-2 .. own+3);
// lanes are 16 B apart, so the three LDS.128 per row are conflict-free as they stand.  Same
// scalars as above with D(s) = (w1-w2) + (w4-w3), e(s) = (w0-w3) + (w5-w2), rsU(s) = w2+w3.
template <int LIB>
__device__ __forceinline__ void march_frame2(const double *__restrict__ cur, const TiledParams &P, Sums &F) {
    double wp[2], wc[6], wn[6];
    double rsU[12], D[12];
    double gx[2] = {0, 0}, gy[2] = {0, 0}, u2[2] = {0, 0}, ul[2] = {0, 0};
    double sD = 0, sE = 0, sU = 0, sDy = 0;
#pragma unroll
    for (int s = 0; s < 12; ++s) {
        const double2 *src = reinterpret_cast<const double2 *>(cur + s * HC);
        const double2 a0 = src[0], a1 = src[1], a2 = src[2];
        wn[0] = a0.x; wn[1] = a0.y; wn[2] = a1.x; wn[3] = a1.y; wn[4] = a2.x; wn[5] = a2.y;
        rsU[s] = wn[2] + wn[3];
        D[s] = (wn[1] - wn[2]) + (wn[4] - wn[3]);
        if (s >= 2 && s <= 9) {
            sD += D[s];
            sE += (wn[0] - wn[3]) + (wn[5] - wn[2]);
            sU += rsU[s];
            if constexpr (kNeedAdv<LIB>) sDy += (wn[3] + wn[4]) - (wn[1] + wn[2]);
        }
        if (s >= 3 && s <= 10) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int q = c + 2;
                const double dx = wn[q] - wp[c];
                const double dy = wc[q + 1] - wc[q - 1];
                gx[c] = fma(dx, dx, gx[c]);
                gy[c] = fma(dy, dy, gy[c]);
                if constexpr (kRich<LIB>) {
                    const double Lq = fma(P.kappa, wc[q], fma(P.rho, wn[q] + wp[c], wc[q + 1] + wc[q - 1]));
                    u2[c] = fma(wc[q], wc[q], u2[c]);
                    ul[c] = fma(wc[q], Lq, ul[c]);
                }
            }
        }
        wp[0] = wc[2]; wp[1] = wc[3];
#pragma unroll
        for (int q = 0; q < 6; ++q) wc[q] = wn[q];
    }
    const double rsL1 = fma(P.rho, (rsU[2] + rsU[0]) - 2.0 * rsU[1], D[1]);
    const double rsL2 = fma(P.rho, (rsU[3] + rsU[1]) - 2.0 * rsU[2], D[2]);
    const double rsL9 = fma(P.rho, (rsU[10] + rsU[8]) - 2.0 * rsU[9], D[9]);
    const double rsL10 = fma(P.rho, (rsU[11] + rsU[9]) - 2.0 * rsU[10], D[10]);
    F.SL = fma(P.rho, (rsU[10] - rsU[9]) - (rsU[2] - rsU[1]), sD);
    F.SE1 = (rsL1 - rsL2) + (rsL10 - rsL9);
    F.SE2 = fma(P.rho, (D[10] + D[1]) - (D[2] + D[9]), fma(-3.0, sD, sE));
    F.SU = sU;
    F.SGx = gx[0] + gx[1];
    F.SGy = gy[0] + gy[1];
    if constexpr (kNeedAdv<LIB>) {
        F.SDy = sDy;
        F.SDx = (rsU[10] + rsU[9]) - (rsU[2] + rsU[1]);
    }
    if constexpr (kRich<LIB>) {
        F.SU2 = u2[0] + u2[1];
        F.SUL = ul[0] + ul[1];
    }
}

__device__ __forceinline__ double sum_frame_u2(const double *__restrict__ own) {
    double s0 = 0, s1 = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const double2 a = *reinterpret_cast<const double2 *>(own + r * HC);
        s0 += a.x;
        s1 += a.y;
    }
    return s0 + s1;
}

// Sum of u over the lane's own 8 rows x 4 columns (the frame after a chunk only feeds u_t).
//   own : halo tile, pointing at the warp's first OUTPUT row and the lane's first OWN column
__device__ __forceinline__ double sum_frame_u(const double *__restrict__ own, const int sw) {
    double s0 = 0, s1 = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        double2 a, b;
        load_pair(reinterpret_cast<const double2 *>(own + r * HC), 0, sw, a, b);
        s0 += a.x + a.y;
        s1 += b.x + b.y;
    }
    return s0 + s1;
}

template <int LIB, int NF, int NW>
__global__ void __launch_bounds__(32 * NW, Geo<NW>::CTAS_PER_SM) k1_tiled_b88(const __grid_constant__ CUtensorMap tmap,
                                                                              TiledParams P) {
    using G_ = Geo<NW>;
    constexpr int TI = G_::TI, HR = G_::HR, STAGE_DOUBLES = G_::STAGE_DOUBLES, STAGE_BYTES = G_::STAGE_BYTES;
    constexpr int MAXWRAP = G_::MAXWRAP, THREADS = G_::THREADS;
    constexpr int p = Lib<LIB>::P;
    constexpr int S = PG_STATS_LEN(p);
    constexpr int W = p + 2;
    constexpr int NE = (S + 31) / 32;   // lane-owned statistics entries
    constexpr int SB = G_::slots(p);    // block rows per staging batch
    constexpr int KC = G_::KC, LPB = G_::LPB, RPW = G_::RPW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stages = reinterpret_cast<double *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + NSTAGE * STAGE_BYTES);
    double *ext_all = reinterpret_cast<double *>(smem_raw + NSTAGE * STAGE_BYTES + 64);  // [NW][SB][W]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sw = (lane >> 2) & 1;
    const int band = KC == 4 ? warp : warp >> 1;                   // 8-row block band of the tile
    const int col0 = (KC == 4 ? 0 : (warp & 1) * 64) + lane * KC;   // tile column where the lane's segment starts
    double *ext = ext_all + warp * SB * W;

    int ea[NE], eb[NE];
    bool ev[NE];
#pragma unroll
    for (int k = 0; k < NE; ++k) {
        const int e = lane + 32 * k;
        ev[k] = e < S;
        ea[k] = eb[k] = 0;
        if (ev[k]) stats_pair(e, p, ea[k], eb[k]);
    }
    double acc[NF][NE];
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int k = 0; k < NE; ++k) acc[f][k] = 0.0;
    // Small libraries (p <= 5): every lane that owns a block row keeps a PRIVATE copy of the whole
    // statistics vector in registers (S <= 33 FMAs per emitted row, no staging, no shuffles);
    // larger ones stage rows in shared memory and spread the S entries over the lanes.
    constexpr bool PRIV = S * NF <= 36 && NW != 16;   // register budget: true library (S = 18), or p = 5 with one fold
    constexpr int SP = PRIV ? S : 1;
    double pacc[NF][SP];
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int e = 0; e < SP; ++e) pacc[f][e] = 0.0;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&bars[s], 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    const int64_t frame = P.A0 * P.A1;
    const int n_tiles = P.n_tiles0 * P.n_tiles1;
    const int64_t n_items = (int64_t)n_tiles * P.n_chunks;
    unsigned long long bad_rows = 0, bad_fold = 0;

    // item -> (tile origin, first frame, number of row frames)
    auto geometry = [&](int64_t item, int &i0, int &j0, int64_t &tb0, int &nf) {
        const int tile = (int)(item % n_tiles), chunk = (int)(item / n_tiles);
        i0 = (tile / P.n_tiles1) * TI;
        j0 = (tile % P.n_tiles1) * TJ;
        tb0 = (int64_t)chunk * P.chunk_tb;
        nf = (int)((min(P.nbt, tb0 + P.chunk_tb) - tb0) * P.bt);
    };

    // ---- producer (thread 0): one continuous stream of frame loads over all items of this CTA,
    // always two loads ahead of the consumer, so the pipeline never drains between items.  The
    // cursor is advanced incrementally: the divisions of geometry() run once per item, not per frame.
    int64_t p_item = blockIdx.x;
    int p_i0 = 0, p_j0 = 0, p_t = 0, p_left = 0;
    uint32_t p_g = 0;
    if (tid == 0 && p_item < n_items) {
        int nf;
        int64_t tb0;
        geometry(p_item, p_i0, p_j0, tb0, nf);
        p_t = (int)(tb0 * P.bt);
        p_left = nf + 1;
    }
    auto produce = [&]() {
        if (p_left == 0) return;
        uint64_t *bar = &bars[p_g % NSTAGE];
        fence_proxy_async();
        mbar_expect_tx(bar, STAGE_BYTES);
        tma_load_3d(stages + (p_g % NSTAGE) * STAGE_DOUBLES, &tmap, bar, p_j0 - 2, p_i0 - 2, p_t);
        ++p_g;
        ++p_t;
        if (--p_left == 0) {
            p_item += gridDim.x;
            if (p_item < n_items) {
                int nf;
                int64_t tb0;
                geometry(p_item, p_i0, p_j0, tb0, nf);
                p_t = (int)(tb0 * P.bt);
                p_left = nf + 1;
            }
        }
    };
    if (tid == 0) { produce(); produce(); }

    uint32_t G = 0;  // consumer load index (stage = G % 3, parity = (G / 3) & 1)
    Sums A;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        int i0, j0, nf;
        int64_t tb0;
        geometry(item, i0, j0, tb0, nf);
        const int64_t t0 = tb0 * P.bt;   // frames t0 .. t0+nf are loaded; the last one only feeds u_t
        const bool wl = j0 == 0, wr = j0 + TJ == P.A1, wt = i0 == 0, wb = i0 + TI == P.A0;
        const bool border = wl || wr || wt || wb;

        // periodic wrap cells of this tile handled by this thread: tile offset and offset inside a frame
        int w_off[MAXWRAP];
        int64_t w_src[MAXWRAP];
#pragma unroll
        for (int k = 0; k < MAXWRAP; ++k) w_off[k] = -1;
        if (border) {
#pragma unroll
            for (int k = 0; k < MAXWRAP; ++k) {
                int c = tid + k * THREADS, R = -1, C = -1;
                if (wl) { if (c >= 0 && c < 2 * HR) { R = c >> 1; C = c & 1; } c -= 2 * HR; }
                if (wr) { if (c >= 0 && c < 2 * HR) { R = c >> 1; C = HC - 2 + (c & 1); } c -= 2 * HR; }
                if (wt) { if (c >= 0 && c < 2 * HC) { R = c / HC; C = c % HC; } c -= 2 * HC; }
                if (wb) { if (c >= 0 && c < 2 * HC) { R = HR - 2 + c / HC; C = c % HC; } c -= 2 * HC; }
                if (R >= 0) {
                    w_off[k] = R * HC + C;
                    w_src[k] = wrap((int64_t)i0 - 2 + R, P.A0) * P.A1 + wrap((int64_t)j0 - 2 + C, P.A1);
                }
            }
        }
        double w_val[MAXWRAP];
        auto wrap_fetch = [&](int64_t t) {
#pragma unroll
            for (int k = 0; k < MAXWRAP; ++k)
                if (w_off[k] >= 0) w_val[k] = __ldg(P.U + t * frame + w_src[k]);
        };
        auto wrap_store = [&](double *stage) {
#pragma unroll
            for (int k = 0; k < MAXWRAP; ++k)
                if (w_off[k] >= 0) stage[w_off[k]] = w_val[k];
        };
        if (border) wrap_fetch(t0);

        int fold = 0;
        double su_first = 0.0;
        const int64_t ib = (int64_t)(i0 >> 3) + band, jb = (int64_t)((j0 + col0) >> 3);
        for (int f = 0; f <= nf; ++f, ++G) {
            double *st = stages + (G % NSTAGE) * STAGE_DOUBLES;
            mbar_wait(&bars[G % NSTAGE], (G / NSTAGE) & 1);
            if (border) wrap_store(st);
            __syncthreads();  // wrap stores visible; every warp finished the previous frame, whose stage is free
            if (tid == 0) produce();                       // load G+2 -> the stage just freed
            if (border && f < nf) wrap_fetch(t0 + f + 1);  // consumed after the next barrier wait

            Sums F;
            if constexpr (KC == 4) {
                if (f < nf) march_frame<LIB>(st + (band * 8) * HC + col0, P, F, sw);
                else F.SU = sum_frame_u(st + (band * 8 + 2) * HC + col0 + 2, sw);
            } else {
                if (f < nf) march_frame2<LIB>(st + (band * 8) * HC + col0, P, F);
                else F.SU = sum_frame_u2(st + (band * 8 + 2) * HC + col0 + 2);
            }

            if (f % P.bt == 0 && f > 0) {
                // ---- the t-block that ended at frame f-1: u_t telescopes to (sum u(f) - sum u(f-bt)) / dt.
                // The LPB lanes of an 8-column block hold one block's sums.
                double SY = F.SU - su_first;
#define PG_PAIR(x)                                                  \
    do {                                                            \
        x += __shfl_xor_sync(0xffffffffu, x, 1);                    \
        if constexpr (LPB == 4) x += __shfl_xor_sync(0xffffffffu, x, 2); \
    } while (0)
                PG_PAIR(A.SL); PG_PAIR(A.SE1); PG_PAIR(A.SE2); PG_PAIR(A.SGx); PG_PAIR(A.SGy); PG_PAIR(SY);
                if constexpr (kNeedAdv<LIB>) { PG_PAIR(A.SDx); PG_PAIR(A.SDy); }
                if constexpr (kRich<LIB>) { PG_PAIR(A.SU); PG_PAIR(A.SU2); PG_PAIR(A.SUL); }
#undef PG_PAIR
                const double invN = 1.0 / (64.0 * (double)P.bt);
                const double lap = P.r1 * A.SL * invN;
                const double bih = P.r1 * P.r1 * fma(P.rho, A.SE1, A.SE2) * invN;
                const double gsq = fma(P.q0, A.SGx, P.q1 * A.SGy) * invN;
                const double y = SY * P.rdt * invN;
                double th[p];
                if constexpr (LIB == PG_LIB_KS_TRUE) {
                    th[0] = lap; th[1] = bih; th[2] = gsq;
                } else if constexpr (LIB == PG_LIB_KS_TRUE_ADV) {
                    th[0] = lap; th[1] = bih; th[2] = gsq; th[3] = P.h0 * A.SDx * invN; th[4] = P.h1 * A.SDy * invN;
                } else if constexpr (LIB == PG_LIB_KS_RICH) {
                    th[0] = 1.0; th[1] = A.SU * invN; th[2] = A.SU2 * invN; th[3] = P.h0 * A.SDx * invN;
                    th[4] = P.h1 * A.SDy * invN; th[5] = lap; th[6] = bih; th[7] = gsq; th[8] = P.r1 * A.SUL * invN;
                } else {
                    th[0] = 1.0; th[1] = A.SU * invN; th[2] = A.SU2 * invN; th[3] = lap; th[4] = bih; th[5] = gsq;
                    th[6] = P.r1 * A.SUL * invN;
                }
                A = Sums();
                bool fin = isfinite(y);
#pragma unroll
                for (int k = 0; k < p; ++k) fin = fin && isfinite(th[k]);
                bool valid = (lane % LPB) == 0;
                if (valid && !fin) { valid = false; ++bad_rows; }
                else if (valid && (fold < 0 || fold >= NF)) { valid = false; ++bad_fold; }
                if constexpr (PRIV) {
                    if (valid) {
                        double m[NF];
#pragma unroll
                        for (int ff = 0; ff < NF; ++ff) m[ff] = (NF == 1 || fold == ff) ? 1.0 : 0.0;
                        auto add = [&](int e, double v) {
#pragma unroll
                            for (int ff = 0; ff < NF; ++ff) pacc[ff][e] = fma(v, m[ff], pacc[ff][e]);
                        };
                        add(0, 1.0);
                        add(1, y);
                        add(2, y * y);
                        int e = 3 + 2 * p;
#pragma unroll
                        for (int i = 0; i < p; ++i) {
                            add(3 + i, th[i]);
                            add(3 + p + i, th[i] * y);
#pragma unroll
                            for (int j = i; j < p; ++j) add(e++, th[i] * th[j]);
                        }
                    }
                } else {
#pragma unroll
                for (int h = 0; h < RPW / SB; ++h) {
                    const bool mine = valid && (lane / LPB) / SB == h;
                    if (mine) {
                        double *r = ext + ((lane / LPB) % SB) * W;
                        r[0] = 1.0; r[1] = y;
#pragma unroll
                        for (int k = 0; k < p; ++k) r[2 + k] = th[k];
                    }
                    __syncwarp();
                    const unsigned vm = __ballot_sync(0xffffffffu, mine);
#pragma unroll
                    for (int slot = 0; slot < SB; ++slot) {
                        const int src = (h * SB + slot) * LPB;
                        if (!((vm >> src) & 1u)) continue;
                        const int fr = __shfl_sync(0xffffffffu, fold, src);
                        const double *r = ext + slot * W;
#pragma unroll
                        for (int k = 0; k < NE; ++k) {
                            if (!ev[k]) continue;
                            const double prod = r[ea[k]] * r[eb[k]];
#pragma unroll
                            for (int ff = 0; ff < NF; ++ff) acc[ff][k] += (fr == ff) ? prod : 0.0;
                        }
                    }
                    __syncwarp();
                }
                }
            }
            if (f < nf) {
                if (f % P.bt == 0) {
                    // a t-block starts at this frame: remember sum u, fetch its fold id (used at its end)
                    su_first = F.SU;
                    const int64_t tbs = tb0 + f / P.bt;
                    if (P.fold_of_row) fold = P.fold_of_row[(tbs * P.nB0 + ib) * P.nB1 + jb];
                    else if (P.fold_of_frame) fold = P.fold_of_frame[tbs * P.bt];
                }
                A.SL += F.SL; A.SE1 += F.SE1; A.SE2 += F.SE2; A.SGx += F.SGx; A.SGy += F.SGy;
                if constexpr (kNeedAdv<LIB>) { A.SDx += F.SDx; A.SDy += F.SDy; }
                if constexpr (kRich<LIB>) { A.SU += F.SU; A.SU2 += F.SU2; A.SUL += F.SUL; }
            }
        }
    }
    if (bad_rows) atomicAdd(&P.counters[0], bad_rows);
    if (bad_fold) atomicAdd(&P.counters[1], bad_fold);
    double *out = P.partials + ((int64_t)blockIdx.x * NW + warp) * NF * S;
    if constexpr (PRIV) {
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int e = 0; e < S; ++e) {
                double v = pacc[f][e];
#pragma unroll
                for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == (e & 31)) out[f * S + e] = v;
            }
    } else {
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int k = 0; k < NE; ++k)
                if (ev[k]) out[f * S + lane + 32 * k] = acc[f][k];
    }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}
// warps per CTA (see Geo): 16, 8 or 4; PG_TILED_WARPS overrides the default for experiments
static int tiled_warps() {
    const int w = env_int("PG_TILED_WARPS", 8);
    return w == 4 || w == 16 ? w : 8;
}

bool tiled_plan(const K1Params &P, int lib, int64_t nBt, int n_sm, TiledPlan &plan) {
    if (P.dialect != PG_FD_KS_PERIODIC) return false;
    if (lib != PG_LIB_KS_TRUE && lib != PG_LIB_KS_TRUE_ADV && lib != PG_LIB_KS_RICH && lib != PG_LIB_KS_RICH_NOADV)
        return false;
    if (P.b0 != 8 || P.b1 != 8) return false;
    if (P.n_folds > 2) return false;
    if (P.A1 % 2 != 0 || (reinterpret_cast<uintptr_t>(P.U) & 15)) return false;   // TMA: 16-byte strides / base
    const int NW = tiled_warps();
    const int TI = NW == 4 ? 32 : 64;
    const int workers = n_sm * (NW == 4 ? 2 : 1);
    const int64_t nt0 = P.A0 / TI, nt1 = P.A1 / TJ;
    const int64_t nbt = (P.T - 1) / P.bt;   // full t-blocks only; a ragged last one goes to the generic kernel
    if (nt0 < 1 || nt1 < 1 || nbt < 1) return false;
    if (P.T > 0x7fffffff || P.A0 > 0x7fffffff || P.A1 > 0x7fffffff) return false;
    if (!encode_fn()) return false;
    (void)nBt;
    const int64_t n_tiles = nt0 * nt1;
    // choose the number of frame chunks: balance the persistent CTAs, pay one extra frame + pipeline refill per item
    int64_t best_c = 1;
    double best_cost = 1e300;
    for (int64_t c = 1; c <= nbt && c <= 4096; ++c) {
        const int64_t ctb = (nbt + c - 1) / c;
        const int64_t cc = (nbt + ctb - 1) / ctb;           // chunks actually produced
        const int64_t rounds = (n_tiles * cc + workers - 1) / workers;
        const double cost = (double)rounds * ((double)ctb * P.bt + 4.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best_c = cc; }
    }
    const int64_t ctb = (nbt + best_c - 1) / best_c;
    plan.nbt = nbt; plan.nb0 = nt0 * (TI / 8); plan.nb1 = nt1 * (TJ / 8);
    plan.chunk_t = (int)ctb; plan.n_chunks = (nbt + ctb - 1) / ctb;
    plan.n_tiles0 = nt0; plan.n_tiles1 = nt1;
    const int64_t items = n_tiles * plan.n_chunks;
    plan.grid = (int)(items < workers ? items : workers);
    plan.n_parts = (int64_t)plan.grid * NW;
    plan.extra_scratch = 0;
    plan.kernel_id = NW;
    plan.tile0 = TI; plan.tile1 = TJ;
    return true;
}

template <int LIB, int NF, int NW> static int launch_tiled_k(const CUtensorMap &map, const TiledParams &tp, int grid,
                                                            cudaStream_t st) {
    const size_t smem = Geo<NW>::smem(Lib<LIB>::P);
    PG_CUDA(cudaFuncSetAttribute(k1_tiled_b88<LIB, NF, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k1_tiled_b88<LIB, NF, NW><<<grid, 32 * NW, smem, st>>>(map, tp);
    PG_LAUNCHED();
    return PG_OK;
}

template <int LIB> static int launch_tiled_t(const CUtensorMap &map, const TiledParams &tp, int n_folds, int nw, int grid,
                                             cudaStream_t st) {
    if (nw == 16) return n_folds == 1 ? launch_tiled_k<LIB, 1, 16>(map, tp, grid, st) : launch_tiled_k<LIB, 2, 16>(map, tp, grid, st);
    if (nw == 8) return n_folds == 1 ? launch_tiled_k<LIB, 1, 8>(map, tp, grid, st) : launch_tiled_k<LIB, 2, 8>(map, tp, grid, st);
    return n_folds == 1 ? launch_tiled_k<LIB, 1, 4>(map, tp, grid, st) : launch_tiled_k<LIB, 2, 4>(map, tp, grid, st);
}

int tiled_launch(const K1Params &P, int lib, const TiledPlan &plan, double *partials, char *, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) PG_FAIL(PG_EUNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)P.A1, (cuuint64_t)P.A0, (cuuint64_t)P.T};
    const cuuint64_t gstr[2] = {(cuuint64_t)P.A1 * 8, (cuuint64_t)P.A0 * (cuuint64_t)P.A1 * 8};
    const int NW = plan.kernel_id;
    const cuuint32_t box[3] = {HC, (cuuint32_t)(plan.tile0 + 4), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(P.U), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)env_int("PG_TMA_L2PROMO", 3),
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) PG_FAIL(PG_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    TiledParams tp{};
    tp.U = P.U; tp.T = P.T; tp.A0 = P.A0; tp.A1 = P.A1;
    const double d0sq = P.c.d0sq, d1sq = P.c.d1sq;
    tp.rho = d1sq / d0sq;
    tp.kappa = -2.0 * (1.0 + tp.rho);
    tp.r1 = 1.0 / d1sq;
    tp.q0 = 1.0 / (P.c.two_d0 * P.c.two_d0); tp.q1 = 1.0 / (P.c.two_d1 * P.c.two_d1);
    tp.h0 = 1.0 / P.c.two_d0; tp.h1 = 1.0 / P.c.two_d1;
    tp.rdt = 1.0 / P.c.dt;
    tp.bt = P.bt;
    tp.n_tiles0 = (int)plan.n_tiles0; tp.n_tiles1 = (int)plan.n_tiles1; tp.n_chunks = (int)plan.n_chunks;
    tp.chunk_tb = plan.chunk_t;
    tp.nbt = plan.nbt; tp.nB0 = P.nB0; tp.nB1 = P.nB1;
    tp.fold_of_row = P.fold_of_row; tp.fold_of_frame = P.fold_of_frame; tp.n_folds = P.n_folds;
    tp.partials = partials; tp.counters = P.counters;
    switch (lib) {
        case PG_LIB_KS_TRUE: return launch_tiled_t<PG_LIB_KS_TRUE>(map, tp, P.n_folds, NW, plan.grid, st);
        case PG_LIB_KS_TRUE_ADV: return launch_tiled_t<PG_LIB_KS_TRUE_ADV>(map, tp, P.n_folds, NW, plan.grid, st);
        case PG_LIB_KS_RICH: return launch_tiled_t<PG_LIB_KS_RICH>(map, tp, P.n_folds, NW, plan.grid, st);
        case PG_LIB_KS_RICH_NOADV: return launch_tiled_t<PG_LIB_KS_RICH_NOADV>(map, tp, P.n_folds, NW, plan.grid, st);
        default: PG_FAIL(PG_EUNSUPPORTED, "no tiled kernel for library %d", lib);
    }
}

}  // namespace pg
