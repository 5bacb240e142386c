// peer.cu -- tiny all-reduce over peer memory (NVLink, CUDA IPC) for the scalar exchanges of the sharded CVs.
//
// Reference: the sharded CVs exchange a handful of doubles per step with a host-synchronous MPI_Allreduce -- the 2 n_q
// Fourier modes of LamellarOrderParameter (LamellarOrderParameterGPU.cc:70-77), the potential energy of
// WellTemperedEnsemble (WellTemperedEnsemble.cc:58-64) and CollectiveWrapper (CollectiveWrapper.cc:63-69), computeSigma's
// matrix (IntegratorMetaDynamics.cc:1265-1274).  A library all-reduce of 48 bytes costs a kernel launch, a protocol
// round trip and, through a framework, a stream hop; here it is ONE 256-thread kernel on the caller's stream:
//   publish   my values go into row [rank] of a table in EVERY peer's arena (plain peer stores),
//   barrier   release my epoch in every peer's flag row, acquire everybody's epoch in mine,
//   reduce    sum the rows of my table in rank order (deterministic, identical on every rank) back into the caller's buffer.
// The table is double-buffered on the epoch parity: a rank that races ahead into the next all-reduce writes the other
// half while slower ranks still read this one (it cannot get two epochs ahead: the barrier in between needs everybody).
// The epoch is a device-side counter, so the launch can be captured in a CUDA graph and replayed.
#include "common.cuh"

namespace metad {
namespace peer {
constexpr int kMaxPeers = 8;
constexpr int kMaxValues = 32;
constexpr size_t kTableBytes = 2 * kMaxPeers * kMaxValues * sizeof(double);
constexpr size_t kArenaBytes = kTableBytes + 256;

struct Peers { char* arena[kMaxPeers]; unsigned n, rank; };

// mode 0: the whole all-reduce.  Single-process emulation of the ranks (tests; the launch order replaces the flag barrier):
// mode 1 = publish only (every rank first), mode 2 = reduce only.
__global__ void __launch_bounds__(256) allreduce_kernel(Peers pt, unsigned* __restrict__ d_epoch, double* __restrict__ data, unsigned nvals,
                                                        int mode, unsigned* __restrict__ status) {
    pdl_wait(); pdl_trigger();
    const unsigned epoch = *d_epoch + 1u, slot = epoch & 1u;
    const unsigned r = threadIdx.x / kMaxValues, k = threadIdx.x % kMaxValues;
    if (mode != 2 && r < pt.n && k < nvals)
        reinterpret_cast<double*>(pt.arena[r])[((size_t)slot * kMaxPeers + pt.rank) * kMaxValues + k] = data[k];
    if (mode == 1) return;
    __threadfence_system();
    __syncthreads();
    if (mode == 0 && pt.n > 1 && threadIdx.x < pt.n) {
        unsigned* theirs = reinterpret_cast<unsigned*>(pt.arena[threadIdx.x] + kTableBytes) + pt.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
        const unsigned* mine = reinterpret_cast<const unsigned*>(pt.arena[pt.rank] + kTableBytes) + threadIdx.x;
        const long long t0 = clock64();
        unsigned v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
            if ((int)(v - epoch) >= 0) break;
            if (clock64() - t0 > 8000000000LL) { atomicExch(status, 1u); break; }       // a peer never arrived: do not hang the GPU
            __nanosleep(32);
        } while (true);
    }
    __syncthreads();
    if (threadIdx.x < nvals) {
        const double* table = reinterpret_cast<const double*>(pt.arena[pt.rank]) + (size_t)slot * kMaxPeers * kMaxValues;
        double s = 0.0;
        for (unsigned q = 0; q < pt.n; ++q) s += __ldcv(table + q * kMaxValues + threadIdx.x);
        data[threadIdx.x] = s;
    }
    if (threadIdx.x == 0) *d_epoch = epoch;
}
}  // namespace peer
}  // namespace metad

using namespace metad;

struct metad_peer {
    peer::Peers pt = {};
    char* arena = nullptr;
    bool mapped[peer::kMaxPeers] = {};
    bool ready = false, local = false;
    unsigned* d_epoch = nullptr;
    unsigned* d_status = nullptr;
};

extern "C" int metad_peer_create(metad_peer** out, unsigned n_ranks, unsigned rank) {
    METAD_REQUIRE(out, "metad_peer_create: null argument");
    METAD_REQUIRE(n_ranks >= 1 && n_ranks <= (unsigned)peer::kMaxPeers && rank < n_ranks, "metad_peer_create: 1..8 ranks");
    auto* p = new metad_peer();
    p->pt.n = n_ranks; p->pt.rank = rank;
    cudaError_t e = cudaMalloc(&p->arena, peer::kArenaBytes);
    if (e == cudaSuccess) e = cudaMemset(p->arena, 0, peer::kArenaBytes);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_epoch, 2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(p->d_epoch, 0, 2 * sizeof(unsigned));
    if (e != cudaSuccess) { cudaFree(p->arena); cudaFree(p->d_epoch); delete p; return cuda_fail(e, "metad_peer_create", __FILE__, __LINE__); }
    p->d_status = p->d_epoch + 1;
    p->pt.arena[rank] = p->arena;
    if (n_ranks == 1) p->ready = true;
    *out = p;
    return METAD_OK;
}

extern "C" int metad_peer_destroy(metad_peer* p) {
    if (!p) return METAD_OK;
    for (unsigned r = 0; r < (unsigned)peer::kMaxPeers; ++r)
        if (p->mapped[r]) cudaIpcCloseMemHandle(p->pt.arena[r]);
    cudaFree(p->arena); cudaFree(p->d_epoch);
    delete p;
    return METAD_OK;
}

extern "C" int metad_peer_handle(metad_peer* p, void* handle_out64) {
    METAD_REQUIRE(p && handle_out64, "metad_peer_handle: null argument");
    cudaIpcMemHandle_t h;
    METAD_CUDA(cudaIpcGetMemHandle(&h, p->arena));
    memcpy(handle_out64, &h, sizeof h);
    return METAD_OK;
}

extern "C" int metad_peer_connect(metad_peer* p, const void* handles) {
    METAD_REQUIRE(p && handles, "metad_peer_connect: null argument");
    for (unsigned r = 0; r < p->pt.n; ++r) {
        if (r == p->pt.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * (size_t)r, sizeof h);
        void* ptr = nullptr;
        METAD_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        p->pt.arena[r] = (char*)ptr;
        p->mapped[r] = true;
    }
    p->ready = true;
    return METAD_OK;
}

// all ranks in one process on one GPU (tests): the launch order replaces the flag barrier
extern "C" int metad_peer_connect_local(metad_peer* p, metad_peer* const* all) {
    METAD_REQUIRE(p && all, "metad_peer_connect_local: null argument");
    for (unsigned r = 0; r < p->pt.n; ++r) {
        METAD_REQUIRE(all[r] && all[r]->pt.n == p->pt.n && all[r]->pt.rank == r, "metad_peer_connect_local: handles must be in rank order");
        p->pt.arena[r] = all[r]->arena;
    }
    p->ready = true; p->local = true;
    return METAD_OK;
}

// d_data[0..n) <- sum over the ranks, in rank order.  phase (local emulation only): 0 = publish, 1 = reduce; -1 = both
// (real multi-process use; with connect_local and more than one rank the caller runs phase 0 for every rank, then phase 1).
extern "C" int metad_peer_allreduce_sum(metad_peer* p, double* d_data, unsigned n, int phase, metad_stream_t stream) {
    METAD_REQUIRE(p && d_data, "metad_peer_allreduce_sum: null argument");
    METAD_REQUIRE(p->ready, "metad_peer_allreduce_sum: connect the peers first");
    METAD_REQUIRE(n >= 1 && n <= (unsigned)peer::kMaxValues, "metad_peer_allreduce_sum: 1..32 values");
    int mode = 0;
    if (p->local && p->pt.n > 1) {
        METAD_REQUIRE(phase == 0 || phase == 1, "metad_peer_allreduce_sum: local emulation runs phase 0 on every rank, then phase 1");
        mode = phase == 0 ? 1 : 2;
    }
    METAD_CUDA(launch_pdl(false, peer::allreduce_kernel, 1, 256, 0, stream, p->pt, p->d_epoch, d_data, n, mode, p->d_status));
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}

extern "C" int metad_peer_status(metad_peer* p, unsigned* timed_out) {
    METAD_REQUIRE(p && timed_out, "metad_peer_status: null argument");
    METAD_CUDA(cudaDeviceSynchronize());
    METAD_CUDA(cudaMemcpy(timed_out, p->d_status, sizeof(unsigned), cudaMemcpyDeviceToHost));
    return METAD_OK;
}
