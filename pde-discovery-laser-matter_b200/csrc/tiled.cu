// K1 tiled: fused FD + library + block mean + Gram for the KS dialect with (bt, 8, 8) blocks.
//
// One CTA owns a 64 x 128 spatial tile and marches through a chunk of frames.  Each frame's
// (64+4) x (128+4) halo tile is brought into shared memory ONCE by a 3-D TMA tensor copy
// (cp.async.bulk.tensor, mbarrier completion) into a 3-stage ring: stage f is the frame being
// differentiated, stage f+1 supplies u(t+1) for the forward u_t and becomes "current" next
// iteration, stage f+2 is in flight.  Periodic wrap: TMA zero-fills the out-of-bounds halo of a
// border tile; the wrapped values are prefetched with plain loads one iteration ahead and stored
// over the zero fill after the copy has landed.
//
// Inside a frame, warp w owns rows [8w, 8w+8) (one block row) and lane l owns columns
// [4l, 4l+4) (half a block), marching down 12 tile rows with a register sliding window: two
// LDS.128 per row give u at columns own-2 .. own+5, so the unscaled Laplacian L' is formed at
// own-1 .. own+4 without any exchange between lanes.  Block sums of the linear terms use row
// sums and the discrete divergence theorem (sum over a block of lap(L') = differences of L'
// across the block boundary), so only the nonlinear terms cost per-point fp64 work.  At the end
// of a t-block the two lanes of a block pair-reduce by shuffle, form the block-mean row and the
// warp adds its 16 rows to lane-owned Gram entries held in registers.
//
// No tensor cores: p <= 9 and the kernel is HBM / fp64-issue bound.
#include <cuda.h>
#include <math.h>

#include "common.cuh"
#include "launch.h"

namespace pg {

constexpr int TI = 64, TJ = 128;            // tile rows / cols
constexpr int HR = TI + 4, HC = TJ + 4;     // halo tile
constexpr int NSTAGE = 3;
constexpr int STAGE_DOUBLES = HR * HC;      // 8976
constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
constexpr int TW = 8;                       // warps per CTA
constexpr int MAXWRAP = 4;                  // wrap cells per thread (<= 4*256 >= 2*2*68 + 2*2*132)

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
    double *partials;         // [gridDim.x*TW][n_folds][S]
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
template <int LIB> struct Acc {
    // unscaled sums over the lane's half block; which ones exist depends on the library
    double SL = 0, SE1 = 0, SE2 = 0, SGx = 0, SGy = 0, SY = 0;   // all libraries
    double SDx = 0, SDy = 0;                                     // advection columns
    double SU = 0, SU2 = 0, SUL = 0;                             // rich
    __device__ __forceinline__ void reset() { SL = SE1 = SE2 = SGx = SGy = SY = SDx = SDy = SU = SU2 = SUL = 0.0; }
};

template <int LIB> constexpr bool kNeedAdv = (LIB == PG_LIB_KS_TRUE_ADV || LIB == PG_LIB_KS_RICH);
template <int LIB> constexpr bool kRich = (LIB == PG_LIB_KS_RICH || LIB == PG_LIB_KS_RICH_NOADV);

// One frame of one warp band: 12 tile rows march through the register window.
//   cur : halo tile of frame t (row pitch HC), pointing at the warp's first tile row, lane's first column
//   nxt : halo tile of frame t+1, pointing at the warp's first OUTPUT row, lane's first own column
template <int LIB>
__device__ __forceinline__ void march_frame(const double *__restrict__ cur, const double *__restrict__ nxt,
                                            const TiledParams &P, Acc<LIB> &A, const int sw) {
    double uA[8], uB[8], uC[8], uZ[4];       // rows s-2, s-1, s (cols own-2..own+5), row s-3 (own cols)
    double Lc[6], Ln[6];                     // L' rows s-2, s-1 at cols own-1..own+4 (index q-1, q = 1..6)
    double rsM = 0, rsC = 0, rsN = 0;        // row sums of L' over the own columns
    double rsU_m1 = 0;                       // row sum of u over own columns, previous output-side row
    (void)rsU_m1;
#pragma unroll
    for (int s = 0; s < 12; ++s) {
        // ---- load tile row s: 8 doubles as four 16-byte chunks.  Lanes are 32 B apart, so a plain
        // LDS.128 would hit every bank group twice per quarter-warp; lanes with bit 2 set fetch the
        // two chunks of each pair in the opposite order (conflict-free) and swap them back.
        const double2 *src = reinterpret_cast<const double2 *>(cur + s * HC);
        const double2 p0 = src[sw], p1 = src[sw ^ 1], p2 = src[2 + sw], p3 = src[2 + (sw ^ 1)];
        const double2 a0 = sw ? p1 : p0, a1 = sw ? p0 : p1, a2 = sw ? p3 : p2, a3 = sw ? p2 : p3;
        uC[0] = a0.x; uC[1] = a0.y; uC[2] = a1.x; uC[3] = a1.y; uC[4] = a2.x; uC[5] = a2.y; uC[6] = a3.x; uC[7] = a3.y;
        if (s >= 2) {
            // ---- L' of tile row s-1 at q = 1..6
#pragma unroll
            for (int q = 1; q <= 6; ++q) {
                const double v = uC[q] + uA[q];
                const double h = uB[q + 1] + uB[q - 1];
                Ln[q - 1] = fma(P.kappa, uB[q], fma(P.rho, v, h));
            }
            rsN = (Ln[1] + Ln[2]) + (Ln[3] + Ln[4]);
        }
        if (s == 3) {
            // L rows 1 (in Lc) and 2 (in Ln) exist: top boundary flux of the block, d/da0 direction
            A.SE1 += rsC - rsN;
        }
        if (s >= 4) {
            // ---- output tile row s-2 (u in uA, L' in Lc), neighbours: rows s-3 (uZ, rsM) and s-1 (uB, Ln)
            A.SL += rsC;
            A.SE2 += (Lc[0] - Lc[1]) + (Lc[5] - Lc[4]);
            const double2 *ns = reinterpret_cast<const double2 *>(nxt + (s - 4) * HC);
            const double2 m0 = ns[sw], m1 = ns[sw ^ 1];
            const double2 n0 = sw ? m1 : m0, n1 = sw ? m0 : m1;
            const double un[4] = {n0.x, n0.y, n1.x, n1.y};
            const double rsU = (uA[2] + uA[3]) + (uA[4] + uA[5]);
            A.SY += ((un[0] + un[1]) + (un[2] + un[3])) - rsU;
            if constexpr (kRich<LIB>) A.SU += rsU;
            if constexpr (kNeedAdv<LIB>) A.SDy += (uA[6] + uA[5]) - (uA[2] + uA[1]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int q = c + 2;
                const double dx = uB[q] - uZ[c];
                const double dy = uA[q + 1] - uA[q - 1];
                A.SGx = fma(dx, dx, A.SGx);
                A.SGy = fma(dy, dy, A.SGy);
                if constexpr (kNeedAdv<LIB>) A.SDx += dx;
                if constexpr (kRich<LIB>) {
                    A.SU2 = fma(uA[q], uA[q], A.SU2);
                    A.SUL = fma(uA[q], Lc[q - 1], A.SUL);
                }
            }
        }
        if (s == 11) {
            // L rows 9 (in Lc) and 10 (in Ln): bottom boundary flux
            A.SE1 += rsN - rsC;
        }
        // ---- rotate the window
#pragma unroll
        for (int c = 0; c < 4; ++c) uZ[c] = uA[c + 2];
#pragma unroll
        for (int q = 0; q < 8; ++q) { uA[q] = uB[q]; uB[q] = uC[q]; }
#pragma unroll
        for (int q = 0; q < 6; ++q) Lc[q] = Ln[q];
        rsM = rsC; rsC = rsN;
    }
    (void)rsM;
}

template <int LIB, int NF>
__global__ void __launch_bounds__(TW * 32, 1) k1_tiled_b88(const __grid_constant__ CUtensorMap tmap, TiledParams P) {
    constexpr int p = Lib<LIB>::P;
    constexpr int S = PG_STATS_LEN(p);
    constexpr int W = p + 2;
    constexpr int NE = (S + 31) / 32;   // lane-owned statistics entries
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stages = reinterpret_cast<double *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + NSTAGE * STAGE_BYTES);
    double *ext_all = reinterpret_cast<double *>(smem_raw + NSTAGE * STAGE_BYTES + 64);  // [TW][8][W]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double *ext = ext_all + warp * 8 * W;

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
    uint32_t G = 0;  // loads issued so far by this CTA (stage = G % 3, parity = (G / 3) & 1)
    Acc<LIB> A;

    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = (int)(item % n_tiles), chunk = (int)(item / n_tiles);
        const int tj = tile % P.n_tiles1, ti = tile / P.n_tiles1;
        const int64_t tb0 = (int64_t)chunk * P.chunk_tb;
        const int64_t tb1 = min(P.nbt, tb0 + P.chunk_tb);
        const int64_t t0 = tb0 * P.bt;
        const int nf = (int)((tb1 - tb0) * P.bt);   // row frames; frames t0 .. t0+nf are loaded
        const int i0 = ti * TI, j0 = tj * TJ;
        const bool wl = j0 == 0, wr = j0 + TJ == P.A1, wt = i0 == 0, wb = i0 + TI == P.A0;
        const bool border = wl || wr || wt || wb;

        // wrap cells of this tile handled by this thread: tile offset (row*HC+col) and global offset in a frame
        int w_off[MAXWRAP];
        int64_t w_src[MAXWRAP];
#pragma unroll
        for (int k = 0; k < MAXWRAP; ++k) w_off[k] = -1;
        if (border) {
#pragma unroll
            for (int k = 0; k < MAXWRAP; ++k) {
                int c = tid + k * (TW * 32), R = -1, C = -1;
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
        auto issue = [&](uint32_t g, int64_t t) {   // thread 0 only
            uint64_t *bar = &bars[g % NSTAGE];
            fence_proxy_async();
            mbar_expect_tx(bar, STAGE_BYTES);
            tma_load_3d(stages + (g % NSTAGE) * STAGE_DOUBLES, &tmap, bar, j0 - 2, i0 - 2, (int)t);
        };

        __syncthreads();  // every warp has left the previous item: its stages may be overwritten
        const uint32_t G0 = G;
        if (tid == 0) {
            issue(G0, t0);
            issue(G0 + 1, t0 + 1);
        }
        if (border) {
            wrap_fetch(t0);
            mbar_wait(&bars[G0 % NSTAGE], (G0 / NSTAGE) & 1);
            wrap_store(stages + (G0 % NSTAGE) * STAGE_DOUBLES);
            wrap_fetch(t0 + 1);
        } else {
            mbar_wait(&bars[G0 % NSTAGE], (G0 / NSTAGE) & 1);
        }

        for (int f = 0; f < nf; ++f) {
            const uint32_t gc = G0 + f, gn = gc + 1;
            double *st_c = stages + (gc % NSTAGE) * STAGE_DOUBLES;
            double *st_n = stages + (gn % NSTAGE) * STAGE_DOUBLES;
            mbar_wait(&bars[gn % NSTAGE], (gn / NSTAGE) & 1);
            if (border) wrap_store(st_n);
            __syncthreads();  // wrap stores visible; every warp finished frame f-1 (stage (gc+2)%3 is free)
            if (f + 2 <= nf) {
                if (tid == 0) issue(gc + 2, t0 + f + 2);
                if (border) wrap_fetch(t0 + f + 2);
            }
            march_frame<LIB>(st_c + (warp * 8) * HC + lane * 4, st_n + (warp * 8 + 2) * HC + lane * 4 + 2, P, A,
                             (lane >> 2) & 1);

            if ((f + 1) % P.bt == 0) {
                // ---- end of a t-block: the lane pair (2m, 2m+1) holds one block's sums
                const int64_t tb = tb0 + f / P.bt;
#define PG_PAIR(x) x += __shfl_xor_sync(0xffffffffu, x, 1)
                PG_PAIR(A.SL); PG_PAIR(A.SE1); PG_PAIR(A.SE2); PG_PAIR(A.SGx); PG_PAIR(A.SGy); PG_PAIR(A.SY);
                if constexpr (kNeedAdv<LIB>) { PG_PAIR(A.SDx); PG_PAIR(A.SDy); }
                if constexpr (kRich<LIB>) { PG_PAIR(A.SU); PG_PAIR(A.SU2); PG_PAIR(A.SUL); }
#undef PG_PAIR
                const double invN = 1.0 / (64.0 * (double)P.bt);
                const double lap = P.r1 * A.SL * invN;
                const double bih = P.r1 * P.r1 * fma(P.rho, A.SE1, A.SE2) * invN;
                const double gsq = fma(P.q0, A.SGx, P.q1 * A.SGy) * invN;
                const double y = A.SY * P.rdt * invN;
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
                A.reset();
                bool fin = isfinite(y);
#pragma unroll
                for (int k = 0; k < p; ++k) fin = fin && isfinite(th[k]);
                const int64_t ib = (int64_t)(i0 >> 3) + warp, jb = (int64_t)(j0 >> 3) + (lane >> 1);
                int fold = 0;
                if (P.fold_of_row) fold = P.fold_of_row[(tb * P.nB0 + ib) * P.nB1 + jb];
                else if (P.fold_of_frame) fold = P.fold_of_frame[tb * P.bt];
                bool valid = (lane & 1) == 0;
                if (valid && !fin) { valid = false; ++bad_rows; }
                else if (valid && (fold < 0 || fold >= NF)) { valid = false; ++bad_fold; }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const bool mine = valid && (lane >> 4) == h;
                    if (mine) {
                        double *r = ext + ((lane >> 1) & 7) * W;
                        r[0] = 1.0; r[1] = y;
#pragma unroll
                        for (int k = 0; k < p; ++k) r[2 + k] = th[k];
                    }
                    __syncwarp();
                    const unsigned vm = __ballot_sync(0xffffffffu, mine);
#pragma unroll
                    for (int slot = 0; slot < 8; ++slot) {
                        const int src = h * 16 + slot * 2;
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
        G = G0 + nf + 1;
    }
    if (bad_rows) atomicAdd(&P.counters[0], bad_rows);
    if (bad_fold) atomicAdd(&P.counters[1], bad_fold);
    double *out = P.partials + ((int64_t)blockIdx.x * TW + warp) * NF * S;
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int k = 0; k < NE; ++k)
            if (ev[k]) out[f * S + lane + 32 * k] = acc[f][k];
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

static size_t tiled_smem(int p) { return (size_t)NSTAGE * STAGE_BYTES + 64 + sizeof(double) * TW * 8 * (p + 2); }

bool tiled_plan(const K1Params &P, int lib, int64_t nBt, int n_sm, TiledPlan &plan) {
    if (P.dialect != PG_FD_KS_PERIODIC) return false;
    if (lib != PG_LIB_KS_TRUE && lib != PG_LIB_KS_TRUE_ADV && lib != PG_LIB_KS_RICH && lib != PG_LIB_KS_RICH_NOADV)
        return false;
    if (P.b0 != 8 || P.b1 != 8) return false;
    if (P.n_folds > 2) return false;
    if (P.A1 % 2 != 0 || (reinterpret_cast<uintptr_t>(P.U) & 15)) return false;   // TMA: 16-byte strides / base
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
        const int64_t rounds = (n_tiles * cc + n_sm - 1) / n_sm;
        const double cost = (double)rounds * ((double)ctb * P.bt + 4.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best_c = cc; }
    }
    const int64_t ctb = (nbt + best_c - 1) / best_c;
    plan.nbt = nbt; plan.nb0 = nt0 * (TI / 8); plan.nb1 = nt1 * (TJ / 8);
    plan.chunk_t = (int)ctb; plan.n_chunks = (nbt + ctb - 1) / ctb;
    plan.n_tiles0 = nt0; plan.n_tiles1 = nt1;
    const int64_t items = n_tiles * plan.n_chunks;
    plan.grid = (int)(items < n_sm ? items : n_sm);
    plan.n_parts = (int64_t)plan.grid * TW;
    plan.extra_scratch = 0;
    plan.kernel_id = lib;
    plan.tile0 = TI; plan.tile1 = TJ;
    return true;
}

template <int LIB> static int launch_tiled_t(const CUtensorMap &map, const TiledParams &tp, int n_folds, int grid,
                                             cudaStream_t st) {
    const size_t smem = tiled_smem(Lib<LIB>::P);
    if (n_folds == 1) {
        PG_CUDA(cudaFuncSetAttribute(k1_tiled_b88<LIB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k1_tiled_b88<LIB, 1><<<grid, TW * 32, smem, st>>>(map, tp);
    } else {
        PG_CUDA(cudaFuncSetAttribute(k1_tiled_b88<LIB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k1_tiled_b88<LIB, 2><<<grid, TW * 32, smem, st>>>(map, tp);
    }
    PG_LAUNCHED();
    return PG_OK;
}

int tiled_launch(const K1Params &P, int lib, const TiledPlan &plan, double *partials, char *, cudaStream_t st) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) PG_FAIL(PG_EUNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)P.A1, (cuuint64_t)P.A0, (cuuint64_t)P.T};
    const cuuint64_t gstr[2] = {(cuuint64_t)P.A1 * 8, (cuuint64_t)P.A0 * (cuuint64_t)P.A1 * 8};
    const cuuint32_t box[3] = {HC, HR, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(P.U), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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
        case PG_LIB_KS_TRUE: return launch_tiled_t<PG_LIB_KS_TRUE>(map, tp, P.n_folds, plan.grid, st);
        case PG_LIB_KS_TRUE_ADV: return launch_tiled_t<PG_LIB_KS_TRUE_ADV>(map, tp, P.n_folds, plan.grid, st);
        case PG_LIB_KS_RICH: return launch_tiled_t<PG_LIB_KS_RICH>(map, tp, P.n_folds, plan.grid, st);
        case PG_LIB_KS_RICH_NOADV: return launch_tiled_t<PG_LIB_KS_RICH_NOADV>(map, tp, P.n_folds, plan.grid, st);
        default: PG_FAIL(PG_EUNSUPPORTED, "no tiled kernel for library %d", lib);
    }
}

}  // namespace pg
