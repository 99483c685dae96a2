// Developer probe (not part of the product): what does the memory system deliver for the K1 tile
// access pattern, with NO compute?  Persistent CTAs march tiles through frames exactly like
// k1_tiled_b88 (3-D TMA box per frame into a ring, one __syncthreads per frame) and only touch one
// value per thread.  Sweeps box shape / halo alignment / ring depth to separate "HBM bound" from
// "L2-to-SM request amplification" from "TMA latency".
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/tma_probe.cu -o tools/tma_probe && tools/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Probe {
    int n_tiles0, n_tiles1, n_chunks, chunk_frames, ti, tj, off0, off1, nstage, stage_bytes, ahead;
    double *sink;
};

__global__ void __launch_bounds__(256, 1) probe_kernel(const __grid_constant__ CUtensorMap tmap, Probe P) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)P.nstage * P.stage_bytes);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < P.nstage; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bars[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int n_tiles = P.n_tiles0 * P.n_tiles1;
    const long n_items = (long)n_tiles * P.n_chunks;
    long p_item = blockIdx.x;
    int p_f = 0;
    uint32_t p_g = 0;
    auto produce = [&]() {
        if (p_item >= n_items) return;
        const int tile = (int)(p_item % n_tiles), chunk = (int)(p_item / n_tiles);
        const int i0 = (tile / P.n_tiles1) * P.ti, j0 = (tile % P.n_tiles1) * P.tj;
        uint64_t *bar = &bars[p_g % P.nstage];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(P.stage_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                         s32(smem + (size_t)(p_g % P.nstage) * P.stage_bytes)),
                     "l"(&tmap), "r"(s32(bar)), "r"(j0 + P.off1), "r"(i0 + P.off0), "r"(chunk * P.chunk_frames + p_f)
                     : "memory");
        ++p_g;
        if (++p_f >= P.chunk_frames) { p_f = 0; p_item += gridDim.x; }
    };
    if (tid == 0) for (int k = 0; k < P.ahead; ++k) produce();
    uint32_t G = 0;
    double acc = 0;
    for (long item = blockIdx.x; item < n_items; item += gridDim.x) {
        for (int f = 0; f < P.chunk_frames; ++f, ++G) {
            uint64_t *bar = &bars[G % P.nstage];
            const uint32_t parity = (G / P.nstage) & 1;
            asm volatile(
                "{\n.reg .pred p;\nW_L:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra.uni W_D;\nbra.uni W_L;\nW_D:\n}\n" ::"r"(
                    s32(bar)),
                "r"(parity)
                : "memory");
            __syncthreads();
            if (tid == 0) produce();
            acc += reinterpret_cast<const double *>(smem + (size_t)(G % P.nstage) * P.stage_bytes)[tid];
        }
    }
    if (acc == 12345.678) P.sink[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    const long A = 2048, T = argc > 1 ? atol(argv[1]) : 384;
    double *U, *sink;
    CK(cudaMalloc(&U, sizeof(double) * T * A * A));
    CK(cudaMemset(U, 0, sizeof(double) * T * A * A));
    CK(cudaMalloc(&sink, 8));
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    int n_sm = 0;
    CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0));
    struct Case { const char *name; int ti, tj, halo0, halo1, nstage, ahead; };
    const Case cases[] = {
        {"64x128 halo2x2 s3 a2 (K1 now)", 64, 128, 2, 2, 3, 2}, {"64x128 halo2x2 s3 a1", 64, 128, 2, 2, 3, 1},
        {"64x128 halo2x0 s3 a2 (no col halo)", 64, 128, 2, 0, 3, 2}, {"64x128 halo0x0 s3 a2 (no halo)", 64, 128, 0, 0, 3, 2},
        {"64x128 halo2x16 s2 a1 (line-aligned)", 64, 128, 2, 16, 2, 1}, {"32x128 halo2x2 s5 a4", 32, 128, 2, 2, 5, 4},
        {"32x128 halo2x0 s5 a4", 32, 128, 2, 0, 5, 4}, {"32x252 halo2x2 s3 a2", 32, 252, 2, 2, 3, 2},
        {"16x252 halo2x2 s5 a4", 16, 252, 2, 2, 5, 4}, {"128x64 halo2x2 s3 a2", 128, 64, 2, 2, 3, 2},
        {"32x256 halo0x0 s3 a2", 32, 256, 0, 0, 3, 2}, {"8x256 halo0x0 s12 a10", 8, 256, 0, 0, 12, 10},
    };
    for (const Case &c : cases) {
        const int R = c.ti + 2 * c.halo0, C = c.tj + 2 * c.halo1;
        if (C > 256) { printf("%-40s skipped (box cols > 256)\n", c.name); continue; }
        CUtensorMap map;
        const cuuint64_t gdim[3] = {(cuuint64_t)A, (cuuint64_t)A, (cuuint64_t)T};
        const cuuint64_t gstr[2] = {(cuuint64_t)A * 8, (cuuint64_t)A * A * 8};
        const cuuint32_t box[3] = {(cuuint32_t)C, (cuuint32_t)R, 1}, estr[3] = {1, 1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, U, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%-40s encode failed %d\n", c.name, (int)r); continue; }
        Probe P{};
        P.ti = c.ti; P.tj = c.tj; P.off0 = -c.halo0; P.off1 = -c.halo1;
        P.n_tiles0 = (int)(A / c.ti); P.n_tiles1 = (int)(A / c.tj);
        P.chunk_frames = 96; P.n_chunks = (int)(T / 96);
        P.nstage = c.nstage; P.ahead = c.ahead; P.stage_bytes = R * C * 8; P.sink = sink;
        const size_t smem = (size_t)c.nstage * P.stage_bytes + 256;
        if (smem > 227 * 1024) { printf("%-40s skipped (smem %zu)\n", c.name, smem); continue; }
        CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int it = 0; it < 4; ++it) {
            CK(cudaEventRecord(e0));
            probe_kernel<<<n_sm, 256, smem>>>(map, P);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it > 0 && ms < best) best = ms;
        }
        const double alg = (double)P.n_tiles0 * c.ti * P.n_tiles1 * c.tj * 8.0 * P.n_chunks * 96;
        const double req = (double)P.n_tiles0 * P.n_tiles1 * (double)P.stage_bytes * P.n_chunks * 96;
        printf("%-40s %7.3f ms  algorithmic %7.1f GB/s  box bytes %7.1f GB/s  (box/alg %.3f)\n", c.name, best,
               alg / best / 1e6, req / best / 1e6, req / alg);
    }
    return 0;
}
