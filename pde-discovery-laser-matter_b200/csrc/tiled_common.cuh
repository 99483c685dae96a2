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
// arrive (release.cta) and return how many arrivals the phase was still waiting for BEFORE this one
// (1 => this arrival completed the phase)
__device__ __forceinline__ uint32_t mbar_arrive_pending(uint64_t *bar) {
    uint64_t state;
    uint32_t pending;
    asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(state) : "r"(smem_u32(bar)) : "memory");
    asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(pending) : "l"(state));
    return pending;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

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

// 3-D tensor map over U[T][A0][A1] (fp64) with a (1, box_rows, TJ) box; out-of-bounds elements are zero-filled.
inline CUresult encode_field_map(CUtensorMap *map, const double *U, int64_t T, int64_t A0, int64_t A1, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return CUDA_ERROR_NOT_SUPPORTED;
    const cuuint64_t gdim[3] = {(cuuint64_t)A1, (cuuint64_t)A0, (cuuint64_t)T};
    const cuuint64_t gstr[2] = {(cuuint64_t)A1 * 8, (cuuint64_t)A0 * (cuuint64_t)A1 * 8};
    const cuuint32_t box[3] = {TJ, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(U), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               (CUtensorMapL2promotion)env_int("PG_TMA_L2PROMO", 2), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace pg
