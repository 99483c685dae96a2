// K1 tiled: fused FD + library + block mean + Gram for the KS dialect with (bt, 8, 8) blocks.
//
// One persistent CTA (8 warps) owns a 64 x 128 spatial tile and marches through a chunk of frames.  Each
// frame's (64+4) x 128 row-halo tile is brought into shared memory ONCE by a 4-D TMA tensor copy
// (cp.async.bulk.tensor, mbarrier completion, 128-byte swizzle: tiled_common.cuh) into a 3-stage ring.
// The TMA box is exactly 128 columns = whole 128-byte lines; the two halo columns on either side are
// fetched separately as one 16-byte cell per row and side (a 32-byte sector instead of a 128/256-byte
// line: this is what keeps DRAM traffic near the algorithmic 8 B/point), which also gives the periodic
// wrap along a1 for free.  Wrap along a0: TMA zero-fills the out-of-bounds rows of a border tile and the
// wrapped rows are copied over the zero fill after the load has landed.  A width that is not a multiple of
// 128: the last tile column is shifted left over its neighbour and skips the blocks already counted.
//
// The warps are DECOUPLED (no block-wide barrier per frame, no producer warp): see k1_tiled_b88 below.
//
// The forward u_t never touches frame t+1 pointwise: over a t-block it telescopes to
// (sum u(t0+bt) - sum u(t0)) / dt, and every frame's block sum of u is formed while that frame is current.
//
// Inside a frame, warp w owns rows [8w, 8w+8) (one block row) and lane l owns the 4 columns of group
// g(l) (half a block), marching down 12 tile rows with a register sliding window: four conflict-free
// LDS.128 per row give u at columns own-2 .. own+5 in natural register order (the swizzle does the bank
// spreading), with no exchange between lanes.  Block sums of the linear terms (lap, bih, u_x, u_y, u, u_t)
// reduce by the discrete divergence theorem to per-row boundary scalars (see march_frame), so only the
// nonlinear terms cost per-point fp64 work: about 8 fp64 ops per grid point for the true library.  At the
// end of a t-block the two lanes of a block (l, l ^ 8) pair-reduce by shuffle and form the block-mean row.
//
// No tensor cores: p <= 9 and the kernel is HBM / fp64-issue bound.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "launch.h"
#include "tiled_common.cuh"

namespace pg {

#ifndef PG_NSTAGE
#define PG_NSTAGE 3
#endif
constexpr int NSTAGE = PG_NSTAGE;

// Tile geometry: NW warps per CTA, one 8-row block band per warp (NW = 8: 64 x 128 tile, one CTA per SM).
// Stage layout: swizzled tile [HR][TJ] written by TMA (tiled_common.cuh), then hcol [HR][4] = (left2, right2).
template <int NW> struct Geo {
    static constexpr int TI = 8 * NW;
    static constexpr int HR = TI + 4;
    static constexpr int HOFF = HR * TJ;                        // start of hcol inside a stage
    static constexpr int STAGE_BYTES = stage_bytes_for(HR);     // 1024-byte multiple
    static constexpr int STAGE_DOUBLES = STAGE_BYTES / 8;
    static constexpr int TMA_BYTES = HR * TJ * 8;
    static constexpr int THREADS = 32 * NW;
    // block rows staged per Gram-update batch
    __host__ __device__ static constexpr int slots(int p) { return 16; }   // all 16 blocks of a band in one batch
    __host__ __device__ static constexpr size_t smem(int p) {
        return (size_t)NSTAGE * STAGE_BYTES + 64 + sizeof(double) * NW * slots(p) * (p + 2) + 1024;   // + alignment slack
    }
};

struct TiledParams {
    const double *U;
    int64_t T, A0, A1;
    int A1c;                  // columns covered by tiles: the whole blocks, (A1 / 8) * 8
    double rho, kappa;        // L' = rho*(u[i+1]+u[i-1]) + (u[j+1]+u[j-1]) + kappa*u ; lap = r1*L'
    // Rows are accumulated UNSCALED (block sums without 1/h^k, 1/dt and 1/(64 bt)); sc[] holds the factor each
    // entry of the extended row [1, y, theta_0 ..] lacks, applied once per flush to the Gram products.
    double sc[PG_MAX_P + 2];
    int bt;
    // Pacing (see k1_tiled_b88): epoch_done[k] counts the CTAs that have consumed their k-th group of 2^epoch_shift
    // frames; a CTA loads frames of epoch k only once every CTA is through epoch k - epoch_lead.
    unsigned int *epoch_done;
    int n_epochs, epoch_shift, epoch_lead;
    int n_tiles0, n_tiles1, n_chunks, chunk_tb;
    int64_t nbt;              // t-blocks covered (the last one may hold fewer than bt frames)
    int64_t n_row_frames;     // T - 1
    int64_t nB0, nB1;         // block counts of the whole row space (row numbering)
    const uint8_t *fold_of_row;
    const int32_t *fold_of_frame;
    int n_folds;
    double *partials;         // [gridDim.x*NW][n_folds][S]
    unsigned long long *counters;
    double *rows8;            // EMIT kernels: block-mean rows [nbt][A0/8][A1c/8][p+1] (y first) instead of statistics
    const double *tail_means; // nullable: block means [A0/8][A1c/8] of the frame after the last row frame (it then is a placeholder)
    // nullable: frame T-1 of U is still being written (a copy engine pulling it from the next rank's slab);
    // it may be loaded only once *halo_flag has reached halo_epoch (pg_fd_lib_gram_halo)
    const unsigned int *halo_flag;
    unsigned int halo_epoch;
    // nullable: the time derivative is taken of ANOTHER stack (pg_fd_lib_gram_two).  Over a t-block it telescopes to a
    // difference of (8, 8) block sums of that stack's frames min(k bt, T - 1), k = 0 .. nbt, which a small kernel has
    // put here as [nbt + 1][ysum_nb0][A1c / 8]; the block sums of u of THIS stack then only feed the rich library's column.
    const double *y_sums;
    int64_t ysum_nb0;
};

// ----------------------------------------------------------------------------- per-lane block sums
// Unscaled sums over the lane's half block (4 columns x 8 rows x frames); which ones are live
// depends on the library.
struct Sums {
    double SL = 0, SE1 = 0, SE2 = 0, SGx = 0, SGy = 0;   // all libraries
    double SU = 0;                                      // sum of u: rich column, and u_t by telescoping
    double SDx = 0, SDy = 0;                            // advection columns
    double SU2 = 0, SUL = 0;                            // rich
};

template <int LIB> constexpr bool kNeedAdv = (LIB == PG_LIB_KS_TRUE_ADV || LIB == PG_LIB_KS_RICH);
template <int LIB> constexpr bool kRich = (LIB == PG_LIB_KS_RICH || LIB == PG_LIB_KS_RICH_NOADV);

// Lane addressing: LaneMap / make_lane_map / load_row8 in tiled_common.cuh (swizzled tile, natural register
// order).  A block (8 columns) is the two column groups 2m, 2m+1, owned by lanes l and l ^ 8; the lane with
// bit 3 clear emits the block's row.
__device__ __forceinline__ int block_lane(int b) { return 8 * ((2 * b) & 3) + (b >> 1); }   // lane owning group 2b

// One frame of one warp band: 12 tile rows march through a 3-row register window.
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
__device__ __forceinline__ void march_frame(const double *__restrict__ st, const LaneMap &m, const TiledParams &P, Sums &F) {
    double wp[4], wc[8], wn[8], nx[8];       // own columns of row s-2; rows s-1 and s (all 8 columns); row s+1 in flight
    double rsU[12], D[12];                   // only rows 0-3, 8-11 (rsU) and 1, 2, 9, 10 (D) are ever formed
    // per-column partial sums of the nonlinear terms: four independent FMA chains per quantity
    double gx[4] = {0, 0, 0, 0}, gy[4] = {0, 0, 0, 0}, u2[4] = {0, 0, 0, 0};
    // rich library, block sum of u*L': with the block sum of u^2 already at hand it needs only the products of
    // neighbouring points, sum u(s,q) (u(s+1,q) + u(s-1,q)) = [pairs (1,2), (9,10)] + 2 [pairs (2,3) .. (8,9)] and the
    // same along the row, i.e. ~2.4 FMA per point instead of forming L' at every point (5 operations):
    //   SUL = kappa SU2 + rho (vb + 2 vi) + (he + 2 hi)
    // (same cancellation as u * lap itself: the terms are O(sum u^2), the result O((k h)^2 sum u^2)).
    double vi[2] = {0, 0}, hi[2] = {0, 0}, vb = 0, he = 0;
    // Column-pair sums over the output rows 2..9.  The row scalars D(s) and e(s) are only needed SUMMED over those
    // rows (individually just at rows 1, 2, 9, 10), so the rows are added up per column pair first:
    //   sum D = S16 - Sa,  sum e = S07 - Sb,  sum rsU = Sa + Sb      (8 additions per row instead of 12)
    // (the advection libraries keep w1, w6, w2, w5 apart: sum (w6+w5) - (w2+w1) is their a1-difference column).
    double S07 = 0, Sb = 0, S16 = 0, Sa = 0, S1 = 0, S6 = 0, S2 = 0, S5 = 0;
    load_row8(st, m, 0, nx);
#pragma unroll
    for (int s = 0; s < 12; ++s) {
#pragma unroll
        for (int q = 0; q < 8; ++q) wn[q] = nx[q];
        if (s < 11) load_row8(st, m, s + 1, nx);          // one row ahead: the shared-memory latency hides under this row's arithmetic
        const bool inner = s >= 2 && s <= 9, needU = s <= 3 || s >= 8, needD = s == 1 || s == 2 || s == 9 || s == 10;
        const double b = wn[3] + wn[4];
        double a = 0.0, o = 0.0;
        if (needU || (inner && !kNeedAdv<LIB>)) a = wn[2] + wn[5];
        if (needD || (inner && !kNeedAdv<LIB>)) o = wn[1] + wn[6];
        if (needU) rsU[s] = a + b;
        if (needD) D[s] = o - a;
        if (inner) {
            if constexpr (kNeedAdv<LIB>) { S1 += wn[1]; S6 += wn[6]; S2 += wn[2]; S5 += wn[5]; }
            else { S16 += o; Sa += a; }
            S07 += wn[0] + wn[7];
            Sb += b;
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
                if constexpr (kRich<LIB>) u2[c] = fma(wc[q], wc[q], u2[c]);
            }
            if constexpr (kRich<LIB>) {
                he = fma(wc[1], wc[2], he);
                hi[0] = fma(wc[2], wc[3], hi[0]);
                hi[1] = fma(wc[3], wc[4], hi[1]);
                hi[0] = fma(wc[4], wc[5], hi[0]);
                he = fma(wc[5], wc[6], he);
            }
        }
        if constexpr (kRich<LIB>) {
            if (s >= 2 && s <= 10) {     // rows (s-1, s)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (s == 2 || s == 10) vb = fma(wc[c + 2], wn[c + 2], vb);
                    else vi[c & 1] = fma(wc[c + 2], wn[c + 2], vi[c & 1]);
                }
            }
        }
        // ---- rotate the window
#pragma unroll
        for (int c = 0; c < 4; ++c) wp[c] = wc[c + 2];
#pragma unroll
        for (int q = 0; q < 8; ++q) wc[q] = wn[q];
    }
    if constexpr (kNeedAdv<LIB>) { S16 = S1 + S6; Sa = S2 + S5; }
    const double sD = S16 - Sa, sE = S07 - Sb, sU = Sa + Sb;
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
        F.SDy = (S6 + S5) - (S2 + S1);
        F.SDx = (rsU[10] + rsU[9]) - (rsU[2] + rsU[1]);   // sum_{2..9} (rsU(s+1) - rsU(s-1))
    }
    if constexpr (kRich<LIB>) {
        F.SU2 = (u2[0] + u2[1]) + (u2[2] + u2[3]);
        F.SUL = fma(P.kappa, F.SU2, fma(P.rho, fma(2.0, vi[0] + vi[1], vb), fma(2.0, hi[0] + hi[1], he)));
    }
}

// Sum of u over the lane's own 8 rows x 4 columns (the frame after a chunk only feeds u_t).
__device__ __forceinline__ double sum_frame_u(const double *__restrict__ st, const LaneMap &m) {
    double s0 = 0, s1 = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        double w[4];
        load_row4(st, m, r + 2, w);
        s0 += w[0] + w[1];
        s1 += w[2] + w[3];
    }
    return s0 + s1;
}

// ----------------------------------------------------------------------------- decoupled kernel
// k1_tiled_b88: the warps are DECOUPLED.
//   * no warp is "the producer": every warp releases a stage when it has read what it needs (arrive on
//     empty[s]), and the warp whose arrival completes the phase re-arms it at once with the TMA load
//     of the frame three ahead, so the ring refills as early as possible and nobody waits for anybody's
//     arithmetic (a ninth producer warp would cost a third warp on one SM sub-partition: 170 registers);
//   * there is no block-wide barrier per frame: each consumer warp copies the 16-byte side cells of ITS
//     window (halo columns of its 12 rows; the periodic wrap rows for the top / bottom band of a border
//     tile) with cp.async one frame ahead, so they need warp-level visibility only, and waits on full[s];
//     warps drift by up to a frame, so one warp's t-block epilogue (shuffles, row, Gram update: long
//     dependent chains) overlaps the others' stencil work instead of idling the SM;
//   * PACING: neighbouring tiles share their halo rows / halo cells through L2, but only while the CTAs that
//     own them work on nearby frames; persistent CTAs drift apart over a long chunk and the halo is then
//     re-read from DRAM (measured: DRAM traffic 1.04x the algorithmic bytes over 256 frames, 1.12x over 1024).
//     So the CTA-local frame stream is cut into epochs of 2^epoch_shift frames: the warp that re-arms a stage
//     only issues a load of epoch k once every CTA has consumed epoch k - epoch_lead (one relaxed poll per
//     epoch; the slowest CTA never waits, the others would only have finished early).  The poll gives up after
//     a bounded number of tries: pacing is an optimisation, never a correctness requirement.
//   * TIMEFOLD: folds given per frame (time-holdout folds) or no folds: a warp accumulates for ONE fold
//     at a time in registers and flushes into its partial slot when the fold changes, so any number of
//     folds runs at the single-fold cost.  !TIMEFOLD: fold_of_row, NF <= 2 masked accumulators as before.
#ifndef PG_DNW
#define PG_DNW 8
#endif
#ifndef PG_TILED_WS_DEFAULT
#define PG_TILED_WS_DEFAULT 1
#endif
constexpr int DNW = PG_DNW;   // consumer warps
// warp-specialised kernels: 12 warps are launched at <= 168 registers; the producer warpgroup shrinks to 72 and the two
// consumer warpgroups grow to 216 (2 x 128 x 216 + 128 x 72 = 64512 <= 65536)
constexpr int WS_PRODUCER_REGS = 72, WS_CONSUMER_REGS = 216;

//   * EMIT: the (scaled) block-mean rows are written to P.rows8 instead of being accumulated: first stage of the path
//     for (bt, 8m, 8n) blocks, whose rows are means of these sub-block rows (api.cu).
//   * WS (warp-specialised): a third warpgroup is launched whose first warp is the PRODUCER: it walks the CTA's frame
//     stream, waits for a stage to be released, paces, issues the TMA load and copies the halo-column cells of ALL
//     bands (completion of those copies is tied to the same `full` barrier with cp.async.mbarrier.arrive), so the
//     consumer warps no longer execute any of that control code (ncu: it was a quarter of their non-waiting time,
//     low-IPC branches and address arithmetic repeated by all eight warps).  setmaxnreg hands the producer
//     warpgroup's registers to the consumers, which keep their ~230 registers although 12 warps are resident.
template <int LIB, int NF, bool TIMEFOLD, bool EMIT = false, bool WS = false>
__global__ void __launch_bounds__(32 * (DNW + (WS ? 4 : 0)), 1) k1_tiled_b88(const __grid_constant__ CUtensorMap tmap,
                                                                             const __grid_constant__ CUtensorMap tmap_last,
                                                                             TiledParams P) {
    constexpr int NW = DNW;
    using G_ = Geo<NW>;
    constexpr int TI = G_::TI, HOFF = G_::HOFF, STAGE_DOUBLES = G_::STAGE_DOUBLES;
    constexpr int p = Lib<LIB>::P;
    constexpr int S = PG_STATS_LEN(p);
    constexpr int W = p + 2;
    constexpr bool PRIV = S * NF <= 36;   // every lane keeps the whole statistics vector in registers
    // Lane-owned statistics entries (the path that spreads the vector over the lanes).  The rich libraries' first
    // column is the constant 1, so the p + 2 entries it takes part in (sum theta_0 = G_00 = n, sum theta_0 y = sum y,
    // G_0j = sum theta_j) duplicate other entries: they are not accumulated, the flush copies them.
    constexpr bool SKIPDUP = kRich<LIB> && !PRIV;
    constexpr int NE = ((SKIPDUP ? S - (p + 2) : S) + 31) / 32;
    constexpr int SB = G_::slots(p);    // block rows per staging batch
    static_assert(!TIMEFOLD || NF == 1, "time folds accumulate one fold at a time");
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem_raw = align1024(smem_dyn);
    double *stages = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + NSTAGE * G_::STAGE_BYTES);
    uint64_t *empty = full + NSTAGE;
    volatile int *pace_off = reinterpret_cast<volatile int *>(empty + NSTAGE);   // set when a pacing poll timed out
    double *ext_all = reinterpret_cast<double *>(smem_raw + NSTAGE * G_::STAGE_BYTES + 64);  // [NW][SB][W]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0 && blockIdx.x == 0) {   // diagnostics: SM cycle counter / wall clock at the start of CTA 0 (effective SM clock)
        P.counters[4] = (unsigned long long)clock64();
        P.counters[5] = global_ns();
    }
    if (tid == 0) {
        *pace_off = 0;
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], WS ? 33 : 1); mbar_init(&empty[s], NW); }
        fence_barrier_init();
        fence_proxy_async();
    }
    __syncthreads();

    const int64_t frame = P.A0 * P.A1;
    const int n_tiles = P.n_tiles0 * P.n_tiles1;
    const int64_t n_items = (int64_t)n_tiles * P.n_chunks;
    auto geometry = [&](int64_t item, int &i0, int &j0, int64_t &tb0, int &nf) {
        const int tile = (int)(item % n_tiles), chunk = (int)(item / n_tiles);
        i0 = (tile / P.n_tiles1) * TI;
        j0 = min((tile % P.n_tiles1) * TJ, P.A1c - TJ);   // a width that is not a multiple of 128: the last tile column is shifted left
        tb0 = (int64_t)chunk * P.chunk_tb;
        nf = (int)(min(P.n_row_frames, (tb0 + P.chunk_tb) * P.bt) - tb0 * P.bt);   // a ragged last t-block is shorter
    };

    // load index g of this CTA -> (tile origin, frame); g counts every frame of every item in order
    // gl = CTA-local index of the load (pacing); one lane calls this
    // what a load may have to wait for besides its stage: the pacing epoch and the arrival of the halo frame
    auto load_gate = [&](int t, uint32_t gl) {
        if (P.epoch_done && (gl & ((1u << P.epoch_shift) - 1u)) == 0 && *pace_off == 0) {
            const int k = (int)(gl >> P.epoch_shift) - P.epoch_lead;
            if (k >= 0 && k < P.n_epochs) {
                const volatile unsigned int *c = P.epoch_done + k;
                int tries = 0;
                while (*c < gridDim.x && ++tries < (1 << 14)) __nanosleep(64);
                if (tries >= (1 << 14)) *pace_off = 1;   // some CTA is not making progress (shared GPU?): stop pacing this CTA
            }
        }
        if (P.halo_flag && t == (int)P.T - 1 && !wait_flag_reached(P.halo_flag, P.halo_epoch))
            atomicAdd(&P.counters[3], 1ull);        // the frame never arrived: the API poisons the statistics
    };
    auto issue_tma = [&](uint32_t s, int i0, int j0, int t) {
        fence_proxy_async();
        mbar_expect_tx(&full[s], G_::TMA_BYTES);
        // the shifted tile column of a width with A1 % 16 == 8 starts inside a 16-column group: tmap_last views the field from column 8
        tma_load_4d(stages + s * STAGE_DOUBLES, (j0 & 15) ? &tmap_last : &tmap, &full[s], 0, j0 >> 4, i0 - 2, t);
    };
    auto issue_load = [&](uint32_t s, int i0, int j0, int t, uint32_t gl) {
        load_gate(t, gl);
        issue_tma(s, i0, j0, t);
    };
    // coordinates of the load `ahead` frames after frame f of `item` (geometry i0, j0, t0, nf); walks into the
    // following items of this CTA; false when the CTA's stream of frames ends before that
    auto ahead_coords = [&](int64_t item, int i0, int j0, int t0, int nf, int f, int ahead, int &ai0, int &aj0, int &at) {
        int rem = f + ahead;
        while (rem > nf) {
            rem -= nf + 1;
            item += gridDim.x;
            if (item >= n_items) return false;
            int64_t tb0;
            geometry(item, i0, j0, tb0, nf);
            t0 = (int)(tb0 * P.bt);
        }
        ai0 = i0; aj0 = j0; at = t0 + rem;
        return true;
    };
    if (!WS && tid == 0 && (int64_t)blockIdx.x < n_items) {
        int i0, j0, nf;
        int64_t tb0;
        geometry(blockIdx.x, i0, j0, tb0, nf);
        for (int a = 0; a < NSTAGE; ++a) {
            int ai0, aj0, at;
            if (ahead_coords(blockIdx.x, i0, j0, (int)(tb0 * P.bt), nf, 0, a, ai0, aj0, at)) issue_load(a, ai0, aj0, at, (uint32_t)a);
        }
    }
    if constexpr (WS) {
        if (warp >= NW) {
            // ---- producer warpgroup: give the registers back, one warp feeds the ring
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_PRODUCER_REGS));
            if (warp > NW) return;
            auto count_epoch = [&](uint32_t g) {     // load g has been consumed by every warp
                if (P.epoch_done && ((g + 1) & ((1u << P.epoch_shift) - 1u)) == 0 && (int)(g >> P.epoch_shift) < P.n_epochs)
                    atomicAdd(P.epoch_done + (g >> P.epoch_shift), 1u);
            };
            constexpr int NC = (2 * G_::HR + 31) / 32;     // halo-column cells per lane and frame
            uint32_t Gp = 0, s = 0, ph = 0;
            for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
                int i0, j0, nf;
                int64_t tb0;
                geometry(item, i0, j0, tb0, nf);
                int hoff[NC];
                int64_t hsrc[NC];
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    const int c = lane + 32 * k, Rr = c >> 1, side = c & 1;
                    hoff[k] = c < 2 * G_::HR ? HOFF + Rr * 4 + side * 2 : -1;
                    hsrc[k] = wrap((int64_t)i0 - 2 + Rr, P.A0) * P.A1 + wrap((int64_t)(side ? j0 + TJ : j0 - 2), P.A1);
                }
                const int t0 = (int)(tb0 * P.bt);
                const double *Ft = P.U + (int64_t)t0 * frame;
                for (int f = 0; f <= nf; ++f, ++Gp, Ft += frame) {
                    // the pacing poll (a global load) and the halo-frame flag come BEFORE the wait for the stage: the
                    // producer idles there anyway, so their latency never delays a load
                    if (lane == 0) load_gate(t0 + f, Gp);
                    __syncwarp();
                    if (Gp >= NSTAGE) mbar_wait(&empty[s], ph ^ 1);   // every consumer warp has released load Gp - NSTAGE
                    if (lane == 0) issue_tma(s, i0, j0, t0 + f);
                    if (f < nf) {                                // the frame after a chunk only feeds u_t: own columns
                        double *stg = stages + s * STAGE_DOUBLES;
#pragma unroll
                        for (int k = 0; k < NC; ++k)
                            if (hoff[k] >= 0) cp_async16(stg + hoff[k], Ft + hsrc[k]);
                    }
                    cp_async_mbar_arrive_noinc(&full[s]);
                    if (lane == 0 && Gp >= NSTAGE) count_epoch(Gp - NSTAGE);
                    if (++s == NSTAGE) { s = 0; ph ^= 1; }
                }
            }
            cp_async_wait_all();
            if (lane == 0 && P.epoch_done) {
                // the last loads are still being consumed: count their epochs as they are released, then every epoch
                // this CTA will never reach
                for (uint32_t g = Gp > NSTAGE ? Gp - NSTAGE : 0; g < Gp; ++g) {
                    mbar_wait(&empty[g % NSTAGE], (g / NSTAGE) & 1);
                    count_epoch(g);
                }
                for (int k = (int)(Gp >> P.epoch_shift); k < P.n_epochs; ++k) atomicAdd(P.epoch_done + k, 1u);
            }
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WS_CONSUMER_REGS));
    }

    const LaneMap lm = make_lane_map(warp * 8, HOFF, lane);
    double *ext = ext_all + warp * SB * W;
    int ea[NE], eb[NE], ee[NE];
    bool ev[NE];
#pragma unroll
    for (int k = 0; k < NE; ++k) {
        // the (lane + 32 k)-th entry that is accumulated (static register indices: one scan per k, once per kernel)
        const int want = lane + 32 * k;
        int cnt = 0, fe = -1, fa = 0, fb = 0;
        for (int e = 0; e < S; ++e) {
            int a, b;
            stats_pair(e, p, a, b);
            if (SKIPDUP && (a == 2 || b == 2)) continue;       // an entry with theta_0 = 1: duplicate
            if (cnt == want) { fe = e; fa = a; fb = b; }
            ++cnt;
        }
        ev[k] = fe >= 0; ee[k] = fe < 0 ? 0 : fe; ea[k] = fa; eb[k] = fb;
    }
    double acc[NF][NE];
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int k = 0; k < NE; ++k) acc[f][k] = 0.0;
    constexpr int SP = PRIV ? S : 1;
    double pacc[NF][SP];
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int e = 0; e < SP; ++e) pacc[f][e] = 0.0;

    // this warp's partial slot [n_folds][S]: zeroed here, flushes add into it
    double *slot = P.partials + ((int64_t)blockIdx.x * NW + warp) * P.n_folds * S;
    for (int e = lane; e < P.n_folds * S; e += 32) slot[e] = 0.0;
    __syncwarp();
    // add the register accumulators of mask-fold f into slot[fold] and clear them (warp-collective)
    auto flush = [&](int f, int fold) {
        double *out = slot + fold * S;
        if constexpr (PRIV) {
#pragma unroll
            for (int e = 0; e < S; ++e) {
                double v = pacc[f][e];
#pragma unroll
                for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                int ea_, eb_;
                stats_pair(e, p, ea_, eb_);
                if (lane == (e & 31)) out[e] = fma(v, P.sc[ea_] * P.sc[eb_], out[e]);
                pacc[f][e] = 0.0;
            }
        } else {
#pragma unroll
            for (int k = 0; k < NE; ++k) {
                if (ev[k]) out[ee[k]] = fma(acc[f][k], P.sc[ea[k]] * P.sc[eb[k]], out[ee[k]]);
                acc[f][k] = 0.0;
            }
            if constexpr (SKIPDUP) {
                // entries of the constant column: sum theta_0 = n, sum theta_0 y = sum y, G_0j = (j == 0 ? n : sum theta_j)
                __syncwarp();
                if (lane == 0) out[3] = out[0];
                else if (lane == 1) out[3 + p] = out[1];
                else if (lane - 2 < p) out[3 + 2 * p + (lane - 2)] = lane == 2 ? out[0] : out[3 + (lane - 2)];
            }
        }
        __syncwarp();
    };

    unsigned long long bad_rows = 0, bad_fold = 0;
    int cur_fold = -1;
    uint32_t G = 0;  // consumer load index; cs = G % NSTAGE and cph = (G / NSTAGE) & 1 are kept incrementally
    uint32_t cs = 0, cph = 0;
    Sums A;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        int i0, j0, nf;
        int64_t tb0;
        geometry(item, i0, j0, tb0, nf);
        const int64_t t0 = tb0 * P.bt;
        // a ragged last tile row (A0 a multiple of 8 but not of 64): only the first vrows / 8 bands hold blocks; the
        // periodic wrap rows then sit inside the TMA box, right below the last valid band
        const int vrows = (int)min((int64_t)TI, P.A0 - i0);
        const bool band_ok = warp * 8 < vrows;
        // (i0 + vrows + 2 > A0: one of the two halo rows below the tile lies beyond the frame -- also when the frame ends
        // one row below a whole tile, A0 = 64 k + 1, whose remaining rows belong to the generic kernel)
        const bool need_top = i0 == 0 && warp == 0, need_bot = (int64_t)i0 + vrows + 2 > P.A0 && warp == (vrows >> 3) - 1;

        // side cells of this warp's window (stage rows 8*warp .. 8*warp+11): lanes 0..23 one halo-column cell
        // each; the wrap rows TMA zero-filled (stage rows 0,1 / TI+2,TI+3) belong to the first / last band.
        int h_off = -1;
        int64_t h_src = 0;
        if (lane < 24) {
            const int Rr = warp * 8 + (lane >> 1), side = lane & 1;
            h_off = HOFF + Rr * 4 + side * 2;
            h_src = wrap((int64_t)i0 - 2 + Rr, P.A0) * P.A1 + wrap((int64_t)(side ? j0 + TJ : j0 - 2), P.A1);
        }
        const int64_t top_src = wrap((int64_t)i0 - 2, P.A0) * P.A1 + j0;      // rows i0-2, i0-1 (consecutive mod A0 when A0 >= 2)
        const int64_t top_src1 = wrap((int64_t)i0 - 1, P.A0) * P.A1 + j0;
        const int64_t bot_src = wrap((int64_t)i0 + vrows, P.A0) * P.A1 + j0;
        const int64_t bot_src1 = wrap((int64_t)i0 + vrows + 1, P.A0) * P.A1 + j0;
        // the side cell of this lane in the frame whose copy is issued next: a running pointer (one 64-bit add per
        // frame instead of t * frame + h_src every time)
        const double *hp = P.U + t0 * frame + h_src;
        auto issue_halo = [&](double *stage) {
            if (h_off >= 0) cp_async16(stage + h_off, hp);
            hp += frame;
        };
        auto issue_wrap = [&](double *stage, int64_t t) {
            const double *Ft = P.U + t * frame;
            if (need_top) {
                cp_async16(stage + swz_cell(lane), Ft + top_src + 2 * lane);
                cp_async16(stage + swz_cell(lane + 32), Ft + top_src + 64 + 2 * lane);
                cp_async16(stage + TJ + swz_cell(lane), Ft + top_src1 + 2 * lane);
                cp_async16(stage + TJ + swz_cell(lane + 32), Ft + top_src1 + 64 + 2 * lane);
            }
            if (need_bot) {
                cp_async16(stage + (vrows + 2) * TJ + swz_cell(lane), Ft + bot_src + 2 * lane);
                cp_async16(stage + (vrows + 2) * TJ + swz_cell(lane + 32), Ft + bot_src + 64 + 2 * lane);
                cp_async16(stage + (vrows + 3) * TJ + swz_cell(lane), Ft + bot_src1 + 2 * lane);
                cp_async16(stage + (vrows + 3) * TJ + swz_cell(lane + 32), Ft + bot_src1 + 64 + 2 * lane);
            }
        };
        // first frame of the item: its stage has landed => every warp released the stage's previous frame
        if (!WS || need_top || need_bot) mbar_wait(&full[cs], cph);
        if constexpr (!WS) issue_halo(stages + cs * STAGE_DOUBLES);
        issue_wrap(stages + cs * STAGE_DOUBLES, t0);
        cp_async_commit();

        int fold = 0;
        double su_first = 0.0;
        const int64_t ib = (int64_t)(i0 >> 3) + warp, jb = (int64_t)(j0 >> 3) + (lm.g >> 1);
        // blocks of a shifted last tile column that the tile column before it has already counted
        const bool col_ok = (j0 & (TJ - 1)) == 0 || jb >= (int64_t)(P.n_tiles1 - 1) * (TJ / 8);
        int fb = 0;               // f % bt, kept incrementally (no integer division in the frame loop)
        int64_t tbs = tb0 - 1;    // t-block that starts at the latest frame with fb == 0
        // One frame.  FBc carries f % bt as a COMPILE-TIME constant (>= 0) for the frames of whole t-blocks of three, which
        // the loop below unrolls by three: those frames are never the tail frame, only the first of a block runs the
        // t-block epilogue and the block-start bookkeeping, and none of that is tested per frame.  FBc < 0: f % bt is the
        // run-time counter fb (other block lengths, a ragged last t-block, the tail frame).
        auto frame_step = [&](auto FBc, const int f) {
            constexpr int FB = decltype(FBc)::value;
            constexpr bool ST = FB >= 0;
            double *st = stages + cs * STAGE_DOUBLES;
            // where the stage of this frame goes next (known before the frame is even waited for, so the warp
            // that releases it last can re-arm it without any arithmetic in between)
            [[maybe_unused]] int n_i0 = i0, n_j0 = j0, n_t = (int)t0 + f + NSTAGE;
            [[maybe_unused]] bool n_ok = true;
            if constexpr (!WS) {
                if (f + NSTAGE > nf) n_ok = ahead_coords(item, i0, j0, (int)t0, nf, f, NSTAGE, n_i0, n_j0, n_t);
                if (f + 1 < nf) {
                    // side cells of the next frame, one frame ahead.  Its stage may still hold the frame two back:
                    // wait until every warp released that one (what the producer waits for as well).
                    const uint32_t g1 = G + 1, s1 = cs + 1 == NSTAGE ? 0 : cs + 1, ph1 = s1 == 0 ? cph ^ 1 : cph;
                    if (g1 >= NSTAGE) mbar_wait(&empty[s1], ph1 ^ 1);
                    issue_halo(stages + s1 * STAGE_DOUBLES);
                    if (need_top || need_bot) {
                        mbar_wait(&full[s1], ph1);     // the wrap rows lie inside the TMA box
                        issue_wrap(stages + s1 * STAGE_DOUBLES, t0 + f + 1);
                    }
                }
                cp_async_commit();
                mbar_wait(&full[cs], cph);
                cp_async_wait<1>();
            } else if (need_top || need_bot) {
                // only the wrap rows of a border tile are still copied by the band that reads them, one frame ahead,
                // once the TMA box they lie in has landed (the producer warp has loaded it: its stage was free)
                if (f + 1 < nf) {
                    const uint32_t s1 = cs + 1 == NSTAGE ? 0 : cs + 1, ph1 = s1 == 0 ? cph ^ 1 : cph;
                    mbar_wait(&full[s1], ph1);
                    issue_wrap(stages + s1 * STAGE_DOUBLES, t0 + f + 1);
                }
                cp_async_commit();
                mbar_wait(&full[cs], cph);
                cp_async_wait<1>();
            } else {
                mbar_wait(&full[cs], cph);
            }
            __syncwarp();

            Sums F;
            if (band_ok) {
                if (ST || f < nf) march_frame<LIB>(st, lm, P, F);
                else if (P.tail_means && t0 + nf == P.n_row_frames)   // the slab's trailing frame lives on another GPU
                    F.SU = (lane & 8) ? 0.0 : 64.0 * P.tail_means[ib * (int64_t)(P.A1c >> 3) + jb];
                else F.SU = sum_frame_u(st, lm);
            }

            // release the stage (this warp has read everything it needs from load G); the warp whose arrival
            // completes the phase re-arms it with the load NSTAGE ahead.  Only the wrap rows are generic-proxy
            // writes inside the TMA box, so only their writers need the cross-proxy fence.
            if (need_top || need_bot) fence_proxy_async();
            __syncwarp();
            if constexpr (WS) {
                if (lane == 0) mbar_arrive(&empty[cs]);
            } else if (lane == 0 && mbar_arrive_pending(&empty[cs]) == 1) {
                // every warp has consumed load G: count the epoch it closes, then re-arm the stage
                if (P.epoch_done && ((G + 1) & ((1u << P.epoch_shift) - 1u)) == 0 && (int)(G >> P.epoch_shift) < P.n_epochs)
                    atomicAdd(P.epoch_done + (G >> P.epoch_shift), 1u);
                if (n_ok) issue_load(cs, n_i0, n_j0, n_t, G + NSTAGE);
            }
            if (++cs == NSTAGE) { cs = 0; cph ^= 1; }

            // a t-block ends after bt frames, or with the stack (ragged last block: fewer frames, ks2d:384-389)
            if ((ST ? (FB == 0 && f > 0) : ((fb == 0 && f > 0) || (f == nf && fb != 0))) && band_ok) {
                double SY = F.SU - su_first;
                if (P.y_sums) {
                    const double *ys = P.y_sums + (tbs * P.ysum_nb0 + ib) * (int64_t)(P.A1c >> 3) + jb;
                    SY = (lane & 8) ? 0.0 : ys[P.ysum_nb0 * (int64_t)(P.A1c >> 3)] - ys[0];
                }
#define PG_PAIR(x) x += __shfl_xor_sync(0xffffffffu, x, 8)
                PG_PAIR(A.SL); PG_PAIR(A.SE1); PG_PAIR(A.SE2); PG_PAIR(A.SGx); PG_PAIR(A.SGy); PG_PAIR(SY);
                if constexpr (kNeedAdv<LIB>) { PG_PAIR(A.SDx); PG_PAIR(A.SDy); }
                if constexpr (kRich<LIB>) { PG_PAIR(A.SU); PG_PAIR(A.SU2); PG_PAIR(A.SUL); }
#undef PG_PAIR
                // unscaled block-mean row (scales: TiledParams::sc); |grad u|^2 = q1 (rho SGx + SGy), bih = r1^2 (rho SE1 + SE2)
                const double y = SY;
                const double bih = fma(P.rho, A.SE1, A.SE2), gsq = fma(P.rho, A.SGx, A.SGy);
                double th[p];
                if constexpr (LIB == PG_LIB_KS_TRUE) {
                    th[0] = A.SL; th[1] = bih; th[2] = gsq;
                } else if constexpr (LIB == PG_LIB_KS_TRUE_ADV) {
                    th[0] = A.SL; th[1] = bih; th[2] = gsq; th[3] = A.SDx; th[4] = A.SDy;
                } else if constexpr (LIB == PG_LIB_KS_RICH) {
                    th[0] = 1.0; th[1] = A.SU; th[2] = A.SU2; th[3] = A.SDx; th[4] = A.SDy; th[5] = A.SL; th[6] = bih;
                    th[7] = gsq; th[8] = A.SUL;
                } else {
                    th[0] = 1.0; th[1] = A.SU; th[2] = A.SU2; th[3] = A.SL; th[4] = bih; th[5] = gsq; th[6] = A.SUL;
                }
                A = Sums();
                double y_ = y;
                if (!ST && fb != 0) {
                    // ragged block of fb frames: the common scale assumes bt frames per block
                    const double ratio = (double)P.bt / (double)fb;
                    y_ *= ratio;
#pragma unroll
                    for (int k = 0; k < p; ++k)
                        if (!(kRich<LIB> && k == 0)) th[k] *= ratio;
                }
                if constexpr (EMIT) {
                    if ((lane & 8) == 0 && col_ok) {
                        double *r = P.rows8 + ((tbs * (P.A0 >> 3) + ib) * (int64_t)(P.A1c >> 3) + jb) * (p + 1);
                        r[0] = y_ * P.sc[1];
#pragma unroll
                        for (int k = 0; k < p; ++k) r[1 + k] = th[k] * P.sc[2 + k];
                    }
                }
                bool fin = isfinite(y_);
#pragma unroll
                for (int k = 0; k < p; ++k) fin = fin && isfinite(th[k]);
                bool valid = (lane & 8) == 0 && col_ok && !EMIT;
                if (valid && !fin) { valid = false; ++bad_rows; }
                else if (valid && fold < 0) valid = false;            // excluded on purpose (-1 / 255): not an error
                else if (valid && fold >= P.n_folds) { valid = false; ++bad_fold; }
                if constexpr (TIMEFOLD) {
                    // fold is warp-uniform (one id per t-block): switch the accumulation target when it changes
                    if (fold >= 0 && fold < P.n_folds && fold != cur_fold) {
                        if (cur_fold >= 0) flush(0, cur_fold);
                        cur_fold = fold;
                    }
                }
                if constexpr (PRIV) {
                    if (valid) {
                        double m[NF];
#pragma unroll
                        for (int ff = 0; ff < NF; ++ff) m[ff] = (NF == 1 || fold == ff) ? 1.0 : 0.0;
                        auto add = [&](int e, double v) {
#pragma unroll
                            for (int ff = 0; ff < NF; ++ff) pacc[ff][e] = NF == 1 ? pacc[ff][e] + v : fma(v, m[ff], pacc[ff][e]);
                        };
                        add(0, 1.0);
                        add(1, y_);
                        add(2, y_ * y_);
                        int e = 3 + 2 * p;
#pragma unroll
                        for (int i = 0; i < p; ++i) {
                            add(3 + i, th[i]);
                            add(3 + p + i, th[i] * y_);
#pragma unroll
                            for (int j = i; j < p; ++j) add(e++, th[i] * th[j]);
                        }
                    }
                } else {
#pragma unroll
                    for (int h = 0; h < 16 / SB; ++h) {
                        // The lane that holds the row of block h * SB + slot stages it; a row that does not count (not
                        // finite, excluded fold, already counted by the neighbouring tile column) is staged as ZEROS, so
                        // the product loop below has no branches: the compiler issues its loads together instead of one
                        // LDS -> DFMA dependency per product (ncu: that serialisation was a quarter of the rich kernel).
                        if ((lane & 8) == 0 && (lm.g >> 1) / SB == h) {
                            double *r = ext + ((lm.g >> 1) % SB) * W;
                            r[0] = valid ? 1.0 : 0.0; r[1] = valid ? y_ : 0.0;
#pragma unroll
                            for (int k = 0; k < p; ++k) r[2 + k] = valid ? th[k] : 0.0;
                        }
                        __syncwarp();
#pragma unroll
                        for (int slot_i = 0; slot_i < SB; ++slot_i) {
                            const double *r = ext + slot_i * W;
                            if constexpr (NF == 1) {
#pragma unroll
                                for (int k = 0; k < NE; ++k) acc[0][k] = fma(r[ea[k]], r[eb[k]], acc[0][k]);
                            } else {
                                const int fr = __shfl_sync(0xffffffffu, fold, block_lane(h * SB + slot_i));
#pragma unroll
                                for (int k = 0; k < NE; ++k) {
                                    const double prod = r[ea[k]] * r[eb[k]];
#pragma unroll
                                    for (int ff = 0; ff < NF; ++ff) acc[ff][k] = fma(prod, fr == ff ? 1.0 : 0.0, acc[ff][k]);
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            if ((ST || f < nf) && band_ok) {
                if (ST ? FB == 0 : fb == 0) {
                    su_first = F.SU;
                    ++tbs;
                    if constexpr (TIMEFOLD) fold = P.fold_of_frame ? P.fold_of_frame[t0 + f] : 0;
                    else { fold = P.fold_of_row[(tbs * P.nB0 + ib) * P.nB1 + jb]; if (fold == 255) fold = -1; }
                }
                A.SL += F.SL; A.SE1 += F.SE1; A.SE2 += F.SE2; A.SGx += F.SGx; A.SGy += F.SGy;
                if constexpr (kNeedAdv<LIB>) { A.SDx += F.SDx; A.SDy += F.SDy; }
                if constexpr (kRich<LIB>) { A.SU += F.SU; A.SU2 += F.SU2; A.SUL += F.SUL; }
            }
            if constexpr (!ST)
                if (++fb == P.bt) fb = 0;
            ++G;
        };
        int f = 0;
        if constexpr (WS) {
            if (P.bt == 3)
                for (; f + 3 <= nf; f += 3) {
                    frame_step(std::integral_constant<int, 0>{}, f);
                    frame_step(std::integral_constant<int, 1>{}, f + 1);
                    frame_step(std::integral_constant<int, 2>{}, f + 2);
                }
        }
        for (; f <= nf; ++f) frame_step(std::integral_constant<int, -1>{}, f);
    }
    cp_async_wait<0>();
    if (tid == 0 && blockIdx.x == 0) {
        P.counters[6] = (unsigned long long)clock64();
        P.counters[7] = global_ns();
    }
    if (!WS && P.epoch_done && warp == 0 && lane == 0)
        for (int k = (int)(G >> P.epoch_shift); k < P.n_epochs; ++k) atomicAdd(P.epoch_done + k, 1u);
    if (bad_rows) atomicAdd(&P.counters[0], bad_rows);
    if (bad_fold) atomicAdd(&P.counters[1], bad_fold);
    if constexpr (TIMEFOLD) {
        if (cur_fold >= 0) flush(0, cur_fold);
    } else {
#pragma unroll
        for (int f = 0; f < NF; ++f)
            if (f < P.n_folds) flush(f, f);
    }
}

// (8, 8) block sums of the frames min(k bt, T - 1), k = 0 .. nbt, of a stack: the y side of the two-stack path.
// One thread per block, 16-byte loads (rows are 16-byte aligned: A1 even); a warp reads 2 KB of each of its 8 rows.
__global__ void frame_block_sums8_kernel(const double *__restrict__ Uy, int64_t T, int64_t A0, int64_t A1, int bt, int64_t nbt,
                                         int64_t nb0, int64_t nb1, double *__restrict__ out) {
    const int64_t n = (nbt + 1) * nb0 * nb1;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t jb = idx % nb1, ib = (idx / nb1) % nb0, k = idx / (nb1 * nb0);
        const int64_t t = min(k * bt, T - 1);
        double s = 0.0;
        if (ib * 8 + 8 <= A0) {
            const double *F = Uy + t * A0 * A1 + ib * 8 * A1 + jb * 8;
            double r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double2 *q = reinterpret_cast<const double2 *>(F + i * A1);
                const double2 a = q[0], b = q[1], c = q[2], d = q[3];
                r[i] = ((a.x + a.y) + (b.x + b.y)) + ((c.x + c.y) + (d.x + d.y));
            }
            s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        }
        out[idx] = s;
    }
}

int launch_frame_block_sums8(const double *Uy, int64_t T, int64_t A0, int64_t A1, int bt, int64_t nbt, double *out, cudaStream_t st) {
    const int64_t nb0 = (A0 + 7) / 8, nb1 = A1 / 8, n = (nbt + 1) * nb0 * nb1;
    int64_t g = (n + 255) / 256;
    frame_block_sums8_kernel<<<(unsigned)(g < 1 ? 1 : (g > 148 * 32 ? 148 * 32 : g)), 256, 0, st>>>(Uy, T, A0, A1, bt, nbt, nb0, nb1, out);
    PG_LAUNCHED();
    return PG_OK;
}

// ----------------------------------------------------------------------------- host side
// eight consumer warps per CTA (DNW): one 64 x 128 tile, one CTA per SM

bool tiled_plan(const K1Params &P, int lib, int64_t nBt, int n_sm, TiledPlan &plan) {
    if (P.dialect != PG_FD_KS_PERIODIC) return false;
    if (lib != PG_LIB_KS_TRUE && lib != PG_LIB_KS_TRUE_ADV && lib != PG_LIB_KS_RICH && lib != PG_LIB_KS_RICH_NOADV)
        return false;
    if (P.b0 != 8 || P.b1 != 8) return false;
    if (P.n_folds > 2 && P.fold_of_row) return false;   // per-row folds: two masked accumulator sets; time folds: any number
    // 16-byte aligned rows for TMA / cp.async; at least one 128-column tile.  The tiles cover the whole blocks along a1
    // (a last tile column that does not start at a multiple of 128 is shifted left over its neighbour and skips the
    // blocks already counted); a ragged last block column (A1 % 8 != 0) is left to the generic kernel.
    if (P.A1 % 2 != 0 || P.A1 < TJ || (reinterpret_cast<uintptr_t>(P.U) & 15)) return false;
    const int64_t A1c = P.A1 / 8 * 8;
    const int NW = DNW;
    const int TI = 8 * NW;
    const int workers = n_sm;
    // tile rows: a ragged last one is handled in-kernel when it still consists of whole blocks
    const int64_t nt0 = P.A0 % 8 == 0 ? (P.A0 + TI - 1) / TI : P.A0 / TI, nt1 = (A1c + TJ - 1) / TJ;
    const int64_t nbt = (P.T - 1 + P.bt - 1) / P.bt;   // a ragged last t-block (fewer frames) is handled in-kernel
    if (nt0 < 1 || nt1 < 1 || nbt < 1) return false;
    if (P.T > 0x7fffffff || P.A0 > 0x7fffffff || P.A1 > 0x7fffffff) return false;
    if (!encode_fn()) return false;
    (void)nBt;
    const int64_t n_tiles = nt0 * nt1;
    // choose the number of frame chunks: balance the persistent CTAs, pay one extra frame per item
    int64_t best_c = 1;
    double best_cost = 1e300;
    for (int64_t c = 1; c <= nbt && c <= 4096; ++c) {
        const int64_t ctb = (nbt + c - 1) / c;
        const int64_t cc = (nbt + ctb - 1) / ctb;           // chunks actually produced
        const int64_t rounds = (n_tiles * cc + workers - 1) / workers;
        const double cost = (double)rounds * ((double)ctb * P.bt + 2.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best_c = cc; }
    }
    const int64_t ctb = (nbt + best_c - 1) / best_c;
    plan.nbt = nbt; plan.nb0 = P.A0 % 8 == 0 ? P.A0 / 8 : nt0 * (TI / 8); plan.nb1 = A1c / 8;
    plan.chunk_t = (int)ctb; plan.n_chunks = (nbt + ctb - 1) / ctb;
    plan.n_tiles0 = nt0; plan.n_tiles1 = nt1;
    const int64_t items = n_tiles * plan.n_chunks;
    plan.grid = (int)(items < workers ? items : workers);
    plan.n_parts = (int64_t)plan.grid * NW;
    // pacing epochs: the longest CTA-local frame stream is rounds x (chunk frames + 1)
    {
        const int64_t rounds = (items + workers - 1) / workers;
        const int64_t max_loads = rounds * ((int64_t)plan.chunk_t * P.bt + 1);
        plan.extra_scratch = sizeof(unsigned int) * (size_t)((max_loads >> 2) + 2);
    }
    plan.kernel_id = 1;
    plan.tile0 = TI; plan.tile1 = TJ;
    return true;
}

template <int LIB, int NF, bool TIMEFOLD, bool EMIT = false>
static int launch_tiled_d(const CUtensorMap (&map)[2], const TiledParams &tp, int grid, cudaStream_t st) {
    const size_t smem = Geo<DNW>::smem(Lib<LIB>::P);
    if (env_int("PG_TILED_WS", PG_TILED_WS_DEFAULT)) {
        PG_CUDA(cudaFuncSetAttribute(k1_tiled_b88<LIB, NF, TIMEFOLD, EMIT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k1_tiled_b88<LIB, NF, TIMEFOLD, EMIT, true><<<grid, 32 * (DNW + 4), smem, st>>>(map[0], map[1], tp);
    } else {
        PG_CUDA(cudaFuncSetAttribute(k1_tiled_b88<LIB, NF, TIMEFOLD, EMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k1_tiled_b88<LIB, NF, TIMEFOLD, EMIT><<<grid, 32 * DNW, smem, st>>>(map[0], map[1], tp);
    }
    PG_LAUNCHED();
    return PG_OK;
}

template <int LIB> static int launch_tiled_t(const CUtensorMap (&map)[2], const TiledParams &tp, int n_folds, int nw, int grid,
                                             cudaStream_t st) {
    // time folds / no folds at the single-fold cost; per-row folds: two masked accumulator sets
    (void)n_folds; (void)nw;
    if (tp.rows8) return launch_tiled_d<LIB, 1, true, true>(map, tp, grid, st);
    if (!tp.fold_of_row) return launch_tiled_d<LIB, 1, true>(map, tp, grid, st);
    return launch_tiled_d<LIB, 2, false>(map, tp, grid, st);
}

int tiled_launch(const K1Params &P, int lib, const TiledPlan &plan, double *partials, char *extra, cudaStream_t st,
                 double *rows8) {
    if (!encode_fn()) PG_FAIL(PG_EUNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap map[2];
    const int NW = plan.kernel_id;
    for (int k = 0; k < 2; ++k) {
        const CUresult r = encode_field_map(&map[k], P.U, P.T, P.A0, P.A1, plan.tile0 + 4, k ? (P.A1 / 8 * 8) % 16 : 0);
        if (r != CUDA_SUCCESS) PG_FAIL(PG_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    }
    TiledParams tp{};
    tp.U = P.U; tp.T = P.T; tp.A0 = P.A0; tp.A1 = P.A1;
    tp.A1c = (int)(P.A1 / 8 * 8);
    const double d0sq = P.c.d0sq, d1sq = P.c.d1sq;
    tp.rho = d1sq / d0sq;
    tp.kappa = -2.0 * (1.0 + tp.rho);
    {
        const double invN = 1.0 / (64.0 * (double)P.bt), r1 = 1.0 / d1sq, q1 = 1.0 / (P.c.two_d1 * P.c.two_d1);
        const double h0 = 1.0 / P.c.two_d0, h1 = 1.0 / P.c.two_d1;
        const double lapS = r1 * invN, bihS = r1 * r1 * invN, gS = q1 * invN;
        double *sc = tp.sc;
        for (int k = 0; k < PG_MAX_P + 2; ++k) sc[k] = 1.0;
        sc[1] = invN / P.c.dt;
        switch (lib) {
            case PG_LIB_KS_TRUE: sc[2] = lapS; sc[3] = bihS; sc[4] = gS; break;
            case PG_LIB_KS_TRUE_ADV: sc[2] = lapS; sc[3] = bihS; sc[4] = gS; sc[5] = h0 * invN; sc[6] = h1 * invN; break;
            case PG_LIB_KS_RICH:
                sc[3] = invN; sc[4] = invN; sc[5] = h0 * invN; sc[6] = h1 * invN; sc[7] = lapS; sc[8] = bihS; sc[9] = gS;
                sc[10] = lapS;
                break;
            case PG_LIB_KS_RICH_NOADV: sc[3] = invN; sc[4] = invN; sc[5] = lapS; sc[6] = bihS; sc[7] = gS; sc[8] = lapS; break;
            default: break;
        }
    }
    tp.bt = P.bt;
    // Pacing: epochs of 2^shift frames; a CTA may run `lead` epochs ahead of the slowest (PG_TILED_LEAD < 0: no pacing).
    // Measured at C4 with the decoupled kernel: DRAM traffic 1.11x the algorithmic bytes without pacing, 1.01x with
    // epochs of 4 frames and lead 2 (but the waits cost more than they save), lead 4 is the fastest.  The
    // warp-specialised kernel is faster on the SM side, its CTAs stay closer together by themselves (1.06x without
    // pacing) and it is the WAITS that cost: epochs of 8 frames with a lead of 16-24 are its optimum (5.10-5.15 ms
    // against 5.27 for (4 frames, lead 16), 5.44 without pacing, 5.5-5.7 for (4 frames, lead 4); tools/k1_ab.py).
    // The lead must cover the ring depth (a CTA waits for its own epochs too): >= 2.
    const bool ws = env_int("PG_TILED_WS", PG_TILED_WS_DEFAULT) != 0;
    tp.epoch_shift = env_int("PG_TILED_ESHIFT", ws ? 3 : 2);
    if (tp.epoch_shift < 2) tp.epoch_shift = 2;           // the counters are sized for epochs of >= 4 frames
    tp.epoch_lead = env_int("PG_TILED_LEAD", ws ? 20 : 4);
    if (tp.epoch_lead >= 0 && tp.epoch_lead < 2) tp.epoch_lead = 2;
    tp.n_epochs = (int)(plan.extra_scratch / sizeof(unsigned int)) - 1;
    tp.epoch_done = nullptr;
    if (tp.epoch_lead >= 0 && plan.grid > 1 && extra) {
        tp.epoch_done = reinterpret_cast<unsigned int *>(extra);
        PG_CUDA(cudaMemsetAsync(extra, 0, plan.extra_scratch, st));
    }
    tp.n_tiles0 = (int)plan.n_tiles0; tp.n_tiles1 = (int)plan.n_tiles1; tp.n_chunks = (int)plan.n_chunks;
    tp.chunk_tb = plan.chunk_t;
    tp.nbt = plan.nbt; tp.n_row_frames = P.T - 1; tp.nB0 = P.nB0; tp.nB1 = P.nB1;
    tp.fold_of_row = P.fold_of_row; tp.fold_of_frame = P.fold_of_frame; tp.n_folds = P.n_folds;
    tp.partials = partials; tp.counters = P.counters;
    tp.rows8 = rows8;
    tp.tail_means = P.tail_means;
    tp.halo_flag = P.halo_flag; tp.halo_epoch = P.halo_epoch;
    tp.y_sums = P.y_sums; tp.ysum_nb0 = (P.A0 + 7) / 8;
    switch (lib) {
        case PG_LIB_KS_TRUE: return launch_tiled_t<PG_LIB_KS_TRUE>(map, tp, P.n_folds, NW, plan.grid, st);
        case PG_LIB_KS_TRUE_ADV: return launch_tiled_t<PG_LIB_KS_TRUE_ADV>(map, tp, P.n_folds, NW, plan.grid, st);
        case PG_LIB_KS_RICH: return launch_tiled_t<PG_LIB_KS_RICH>(map, tp, P.n_folds, NW, plan.grid, st);
        case PG_LIB_KS_RICH_NOADV: return launch_tiled_t<PG_LIB_KS_RICH_NOADV>(map, tp, P.n_folds, NW, plan.grid, st);
        default: PG_FAIL(PG_EUNSUPPORTED, "no tiled kernel for library %d", lib);
    }
}

}  // namespace pg
