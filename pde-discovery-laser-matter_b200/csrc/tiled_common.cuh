// Pieces shared by the tiled TMA kernels (tiled.cu: block means, tiled_pw.cu: pointwise rows):
// tile constants, PTX wrappers for mbarrier / TMA, and the tensor-map encoder lookup.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace pg {

constexpr int TJ = 128;                     // tile columns (32 lanes x 4) = TMA box width
constexpr int PITCH = TJ + 4;               // doubles per tile row in a stage incl. its 4 halo-column slots

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
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// arrive (release.cta) and return how many arrivals the phase was still waiting for BEFORE this one
// (1 => this arrival completed the phase)
__device__ __forceinline__ uint32_t mbar_arrive_pending(uint64_t *bar) {
    uint64_t state;
    uint32_t pending;
    asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(state) : "r"(smem_u32(bar)) : "memory");
    asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(pending) : "l"(state));
    return pending;
}
// 16-byte asynchronous global -> shared copy (generic proxy); zero = true writes zeros instead (src is not read)
__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool zero = false) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(zero ? 0 : 16) : "memory");
}
// arrive on `bar` once every cp.async this thread has issued so far has completed (the barrier's expected count
// includes this arrival: .noinc)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Wait (one thread) until the 32-bit flag, written by a stream memory operation behind a copy-engine transfer, has
// reached `epoch` (wrap-safe).  Bounded: ~4 s of polling, then false -- a kernel must never hang on a transfer
// that was not queued.
__device__ __forceinline__ bool wait_flag_reached(const unsigned int *flag, unsigned int epoch) {
    unsigned long long t0 = 0;
    for (unsigned int spins = 0;; ++spins) {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - epoch) >= 0) return true;
        if ((spins & 1023u) == 1023u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) return false;
        }
        __nanosleep(200);
    }
}

// ----------------------------------------------------------------------------- swizzled tile layout
// A stage holds a [rows][128] fp64 tile written by TMA through a 4-D view of the field
// (16 doubles, A1/16 groups, A0, T) with CU_TENSOR_MAP_SWIZZLE_128B: inside every 128-byte segment the
// 16-byte cells are XOR-permuted with the segment index (mod 8).  Lane l owns column group
// g = (l >> 3) + 4 (l & 7) (columns 4g .. 4g+3 = cells 2g, 2g+1): the eight lanes of a quarter-warp then
// sit in eight different segments at the same in-segment cell, which the swizzle spreads over all 32
// banks -- every LDS.128 of "cell 2g + d" is conflict-free with the registers in natural order (no
// operand swaps).  A stage is followed by its halo-column cells [rows][4] (left pair, right pair).
__device__ __forceinline__ int lane_group(int lane) { return (lane >> 3) + 4 * (lane & 7); }
// element offset, inside a row, of 16-byte cell `cell` (0..63)
__host__ __device__ __forceinline__ int swz_cell(int cell) { return ((cell >> 3) << 4) + ((((cell & 7) ^ (cell >> 3)) & 7) << 1); }

// Element offsets (inside a stage) of a lane's four cells 2g-1 .. 2g+2 in band row 0 and their row strides.
// Cell -1 / 64 live in the halo-column array (stride 4) instead of the tile (stride TJ).
struct LaneMap {
    int a[4], s[4];
    int g;
};
__device__ __forceinline__ LaneMap make_lane_map(int row0, int hoff, int lane) {
    LaneMap m;
    m.g = lane_group(lane);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int cell = 2 * m.g - 1 + k;
        if (cell < 0) { m.a[k] = hoff + row0 * 4; m.s[k] = 4; }
        else if (cell > 63) { m.a[k] = hoff + row0 * 4 + 2; m.s[k] = 4; }
        else { m.a[k] = row0 * TJ + swz_cell(cell); m.s[k] = TJ; }
    }
    return m;
}
// columns own-2 .. own+5 of band row s
__device__ __forceinline__ void load_row8(const double *__restrict__ st, const LaneMap &m, int s, double (&w)[8]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int off = m.a[k] + s * m.s[k];
#ifndef PG_NO_OPAQUE_ROW_OFFSET
        // cells -1 / 64 of the two edge lanes live in the halo-column array, so the row stride of k = 0 and k = 3 is a
        // per-lane value.  Left alone, the compiler strength-reduces a[k] + s * stride into ONE register that it
        // increments in place row after row -- and every increment then waits (short scoreboard, write-after-read) for
        // the LDS that is still reading the register.  Making each row's offset opaque forces a fresh IMAD per row.
        if (k == 0 || k == 3) asm("" : "+r"(off));
#endif
        const double2 x = *reinterpret_cast<const double2 *>(st + off);
        w[2 * k] = x.x; w[2 * k + 1] = x.y;
    }
}
// own columns only
__device__ __forceinline__ void load_row4(const double *__restrict__ st, const LaneMap &m, int s, double (&w)[4]) {
    const double2 x = *reinterpret_cast<const double2 *>(st + m.a[1] + s * TJ);
    const double2 y = *reinterpret_cast<const double2 *>(st + m.a[2] + s * TJ);
    w[0] = x.x; w[1] = x.y; w[2] = y.x; w[3] = y.y;
}
// bytes of a stage: tile + halo-column cells, padded so that every stage starts 1024-byte aligned (swizzle atom)
__host__ __device__ constexpr int stage_bytes_for(int rows) { return ((rows * (TJ + 4) * 8 + 1023) / 1024) * 1024; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
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

inline int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// 4-D tensor map over U[T][A0][A1] (fp64) viewed as (16, (A1 - col0)/16, A0, T) from column col0 on, box
// (16, 8, box_rows, 1), 128-byte swizzle; out-of-bounds elements are zero-filled.  A box starts at a column
// j0 with j0 % 16 == col0 (coordinate j0 >> 4); only the whole 16-column groups of a row are visible.
// Requires A1 and col0 even (16-byte aligned rows) and A1 - col0 >= 16.
inline CUresult encode_field_map(CUtensorMap *map, const double *U, int64_t T, int64_t A0, int64_t A1, int box_rows,
                                 int64_t col0 = 0) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return CUDA_ERROR_NOT_SUPPORTED;
    U += col0;
    const cuuint64_t gdim[4] = {16, (cuuint64_t)((A1 - col0) / 16), (cuuint64_t)A0, (cuuint64_t)T};
    const cuuint64_t gstr[3] = {128, (cuuint64_t)A1 * 8, (cuuint64_t)A0 * (cuuint64_t)A1 * 8};
    const cuuint32_t box[4] = {16, TJ / 16, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double *>(U), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               (CUtensorMapL2promotion)env_int("PG_TMA_L2PROMO", 2), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// first 1024-byte aligned address of the dynamic shared memory window (allocate 1024 bytes of slack)
__device__ __forceinline__ unsigned char *align1024(unsigned char *p) {
    const uint32_t a = smem_u32(p);
    return p + (((a + 1023u) & ~1023u) - a);
}

}  // namespace pg
