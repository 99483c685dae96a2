// Multi-GPU plumbing of the sharded path over NVLink peer memory (include/pdegram.h "pg_comm_*").
//
// One process per GPU.  Every rank owns one WORKSPACE of pg_comm_workspace_bytes() bytes that every other rank
// has mapped into its own address space (CUDA VMM allocations exchanged between the processes: torch symmetric
// memory in pde_b200.slabs, cuMemExportToShareableHandle in a C host).  The library never allocates or maps
// peer memory itself: it gets plain device pointers.
//
//   pg_comm_barrier      one warp: rank r stores the epoch into bar[r] of every peer, then waits for its own bar[*]
//   pg_allreduce_stats   ONE launch, one CTA: the rank's vector is stored into slot[epoch & 1][rank] of every peer
//                        (P2P stores over NVSwitch), a release flag follows, the CTA waits for every peer's flag and
//                        sums the slots in RANK ORDER -- the same bits on every rank and from run to run, which a
//                        ring / tree all-reduce does not promise; len <= PG_COMM_MAX_LEN doubles, so latency is
//                        the only cost (2 x 18 doubles for the true library)
//   pg_halo_exchange     copy engine: the next rank's first frame -> this rank's trailing halo frame, then a stream
//                        memory operation (cuStreamWriteValue32: front-end, no SM) publishes the epoch in a LOCAL
//                        flag that the persistent K1 polls before it loads that frame (pg_fd_lib_gram_halo).  No
//                        kernel is involved, because K1 leaves no SM free for one.
//
// Workspace layout (zero-initialised by the owner before the first barrier):
//   [0, 64)      uint32 bar[16]     barrier epochs, one word per sender
//   [64, 128)    uint32 arf[16]     all-reduce epochs, one word per sender
//   [128, ...)   double slot[2][16][PG_COMM_MAX_LEN]
// Two slot sets alternate by epoch parity: a rank can be at most one all-reduce ahead of the slowest one (it needs
// everybody's flag of epoch e to finish epoch e), so the set of epoch e - 1 may still be read while e is written.
#include <cuda.h>
#include <math.h>
#include <new>

#include "common.cuh"
#include "launch.h"

namespace pg {

constexpr size_t WS_BAR = 0, WS_ARF = 64, WS_SLOT = 128;
constexpr size_t WS_BYTES = WS_SLOT + sizeof(double) * 2 * PG_COMM_MAX_RANKS * PG_COMM_MAX_LEN;

struct PeerSet {
    char *ws[PG_COMM_MAX_RANKS];
};

struct Comm {
    int rank, world, device;
    PeerSet peers;
    unsigned int bar_epoch, ar_epoch, halo_epoch;
    unsigned int *halo_flag;        // local device word
    unsigned long long *errors;     // local device counter: waits that timed out
};

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// bounded wait (~4 s): false when the peer never arrived
__device__ __forceinline__ bool wait_epoch(const unsigned int *p, unsigned int epoch) {
    unsigned long long t0 = 0;
    for (unsigned int spins = 0;; ++spins) {
        if ((int)(ld_acquire_sys(p) - epoch) >= 0) return true;
        if ((spins & 1023u) == 1023u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) return false;
        }
        __nanosleep(100);
    }
}

__global__ void __launch_bounds__(32) comm_barrier_kernel(PeerSet peers, int rank, int world, unsigned int epoch,
                                                          unsigned long long *errors) {
    const int q = threadIdx.x;
    if (q < world) {
        __threadfence_system();      // everything this GPU wrote before the barrier is visible to the peers
        st_release_sys(reinterpret_cast<unsigned int *>(peers.ws[q] + WS_BAR) + rank, epoch);
        if (!wait_epoch(reinterpret_cast<const unsigned int *>(peers.ws[rank] + WS_BAR) + q, epoch)) atomicAdd(errors, 1ull);
    }
}

__global__ void __launch_bounds__(256) comm_allreduce_kernel(PeerSet peers, int rank, int world, unsigned int epoch,
                                                            double *__restrict__ stats, int len,
                                                            unsigned long long *errors) {
    __shared__ int failed;
    const int par = epoch & 1u;
    if (threadIdx.x == 0) failed = 0;
    // 1. my vector into slot[par][rank] of every rank (my own included)
    for (int i = threadIdx.x; i < len * world; i += blockDim.x) {
        const int q = i / len, e = i - q * len;
        double *dst = reinterpret_cast<double *>(peers.ws[q] + WS_SLOT) + ((size_t)par * PG_COMM_MAX_RANKS + rank) * PG_COMM_MAX_LEN;
        dst[e] = stats[e];
    }
    __threadfence_system();
    __syncthreads();
    // 2. flags out, flags in
    if (threadIdx.x < world) {
        const int q = threadIdx.x;
        st_release_sys(reinterpret_cast<unsigned int *>(peers.ws[q] + WS_ARF) + rank, epoch);
        if (!wait_epoch(reinterpret_cast<const unsigned int *>(peers.ws[rank] + WS_ARF) + q, epoch)) failed = 1;
    }
    __syncthreads();
    // 3. rank-ordered sum of the slots that landed in MY workspace
    const volatile double *mine = reinterpret_cast<const volatile double *>(peers.ws[rank] + WS_SLOT) + (size_t)par * PG_COMM_MAX_RANKS * PG_COMM_MAX_LEN;
    for (int e = threadIdx.x; e < len; e += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < world; ++q) s = __dadd_rn(s, mine[(size_t)q * PG_COMM_MAX_LEN + e]);
        stats[e] = failed ? nan("") : s;
    }
    if (failed && threadIdx.x == 0) atomicAdd(errors, 1ull);
}

typedef CUresult (*WriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

static WriteValue32Fn write_value_fn() {
    static WriteValue32Fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (WriteValue32Fn)ptr;
    }
    return fn;
}

}  // namespace pg

using namespace pg;

extern "C" {

size_t pg_comm_workspace_bytes(void) { return WS_BYTES; }

int pg_comm_init(int rank, int world, void *const *peer_workspaces_host, void **comm_out) {
    if (!comm_out) PG_FAIL(PG_EINVAL, "comm_out is null");
    *comm_out = nullptr;
    if (world < 1 || world > PG_COMM_MAX_RANKS || rank < 0 || rank >= world) PG_FAIL(PG_EINVAL, "bad rank %d / world %d (at most %d ranks)", rank, world, PG_COMM_MAX_RANKS);
    if (!peer_workspaces_host) PG_FAIL(PG_EINVAL, "peer_workspaces_host is null");
    Comm *c = new (std::nothrow) Comm();
    if (!c) PG_FAIL(PG_ENOMEM, "out of host memory");
    c->rank = rank; c->world = world;
    for (int q = 0; q < PG_COMM_MAX_RANKS; ++q) c->peers.ws[q] = nullptr;
    for (int q = 0; q < world; ++q) {
        if (!peer_workspaces_host[q]) { delete c; PG_FAIL(PG_EINVAL, "workspace pointer of rank %d is null", q); }
        c->peers.ws[q] = (char *)peer_workspaces_host[q];
    }
    c->bar_epoch = c->ar_epoch = c->halo_epoch = 0;
    if (cudaGetDevice(&c->device) != cudaSuccess) { delete c; PG_FAIL(PG_ECUDA, "cudaGetDevice failed"); }
    void *mem = nullptr;
    if (cudaMalloc(&mem, 128) != cudaSuccess) { cudaGetLastError(); delete c; PG_FAIL(PG_ENOMEM, "cannot allocate the halo flag"); }
    if (cudaMemset(mem, 0, 128) != cudaSuccess) { cudaFree(mem); delete c; PG_FAIL(PG_ECUDA, "cudaMemset failed"); }
    c->halo_flag = (unsigned int *)mem;
    c->errors = (unsigned long long *)((char *)mem + 64);
    *comm_out = c;
    return PG_OK;
}

int pg_comm_destroy(void *comm) {
    Comm *c = (Comm *)comm;
    if (!c) return PG_OK;
    if (c->halo_flag) cudaFree(c->halo_flag);
    delete c;
    return PG_OK;
}

int pg_comm_barrier(void *comm, void *stream) {
    Comm *c = (Comm *)comm;
    if (!c) PG_FAIL(PG_EINVAL, "comm is null");
    if (c->world == 1) return PG_OK;
    comm_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(c->peers, c->rank, c->world, ++c->bar_epoch, c->errors);
    PG_LAUNCHED();
    return PG_OK;
}

int pg_allreduce_stats(void *comm, double *stats, int len, void *stream) {
    Comm *c = (Comm *)comm;
    if (!c) PG_FAIL(PG_EINVAL, "comm is null");
    if (!stats) PG_FAIL(PG_EINVAL, "stats is null");
    if (len < 0 || len > PG_COMM_MAX_LEN) PG_FAIL(PG_EINVAL, "len must be in 0..%d", PG_COMM_MAX_LEN);
    if (c->world == 1 || len == 0) return PG_OK;
    comm_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(c->peers, c->rank, c->world, ++c->ar_epoch, stats, len, c->errors);
    PG_LAUNCHED();
    return PG_OK;
}

int pg_halo_exchange(void *comm, double *halo_dst, const double *peer_first_frame, size_t bytes, void *copy_stream,
                     const uint32_t **halo_flag_out, uint32_t *halo_epoch_out) {
    Comm *c = (Comm *)comm;
    if (!c) PG_FAIL(PG_EINVAL, "comm is null");
    if (!halo_dst || !peer_first_frame) PG_FAIL(PG_EINVAL, "null frame pointer");
    WriteValue32Fn wv = write_value_fn();
    if (!wv) PG_FAIL(PG_EUNSUPPORTED, "cuStreamWriteValue32 is not available from this driver");
    cudaStream_t cs = (cudaStream_t)copy_stream;
    PG_CUDA(cudaMemcpyAsync(halo_dst, peer_first_frame, bytes, cudaMemcpyDeviceToDevice, cs));
    const CUresult r = wv((CUstream)cs, (CUdeviceptr)(uintptr_t)c->halo_flag, ++c->halo_epoch, 0);
    if (r != CUDA_SUCCESS) PG_FAIL(PG_ECUDA, "cuStreamWriteValue32 failed with CUresult %d", (int)r);
    if (halo_flag_out) *halo_flag_out = c->halo_flag;
    if (halo_epoch_out) *halo_epoch_out = c->halo_epoch;
    return PG_OK;
}

int64_t pg_comm_errors(void *comm) {
    Comm *c = (Comm *)comm;
    if (!c) return -1;
    unsigned long long v = 0;
    if (cudaMemcpy(&v, c->errors, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)v;
}

}  // extern "C"
