// mesh_p2p.cuh -- peer-memory (NVLink) plumbing of the z-slab sharded mesh path: no NCCL call inside a step.
//
// Every rank owns one "arena" allocation that its peers map through CUDA IPC.  A step uses it as follows
// (reference for the decomposition: HOOMD domain decomposition + CommunicatorGrid + dfft, OrderParameterMesh.cc:263-315,
// 659-746; the reference moves the same data with MPI messages and three host-synchronous MPI_Allreduce calls):
//   * halo planes of the fixed-point density, of Re IFFT(G), and the per-rank partial sums are PUSHED into the
//     neighbours' / all peers' arenas by a small copy kernel (plain P2P stores);
//   * the two transposes of the distributed FFT are fused into the FFT sweeps themselves (mesh_fft_kernels.cuh, PeerOut);
//   * ranks synchronise with a flag barrier in peer memory (one 32-thread kernel), which also performs the tiny
//     all-reduces (sum of the P partial results, in rank order: deterministic and identical on every rank).
#pragma once
#include "common.cuh"
#include "mesh_fft_kernels.cuh"

namespace metad {
namespace p2p {

constexpr int kMaxPeers = fft::kMaxPeers;

// layout of an arena (all offsets in bytes, 256-byte aligned)
struct ArenaLayout {
    size_t pencil, recv, ghost_rho, ghost_inv, sums, cv, flags, total;
};
inline ArenaLayout arena_layout(size_t m_local /* local mesh cells */, size_t plane) {
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    ArenaLayout a;
    size_t o = 0;
    a.pencil = o; o = up(o + m_local * sizeof(float));                 // [nz][ny][kxl] complex
    a.recv = o; o = up(o + m_local * sizeof(float));                   // [src rank][local plane][y][kxl] complex
    a.ghost_rho = o; o = up(o + 2 * (plane + 4) * sizeof(int));        // two halo messages of the fixed-point density
    a.ghost_inv = o; o = up(o + 2 * plane * sizeof(float));            // planes z0-1 and z0+nz of Re IFFT(G)
    a.sums = o; o = up(o + kMaxPeers * 4 * sizeof(double));            // per-rank {sum a^2, sum a, outside, -}
    a.cv = o; o = up(o + kMaxPeers * sizeof(double));                  // per-rank CV partials
    a.flags = o; o = up(o + 8 * kMaxPeers * sizeof(unsigned));         // flags[4 phases][peer] of the fused synchronisation, then [peer] of the barrier kernel
    a.total = o;
    return a;
}

struct PeerTable {
    char* arena[kMaxPeers];
    unsigned n, rank;
};

struct Publish { const double* src; size_t table_offset; unsigned per_rank, n; };

// ---- copy kernel: up to 4 segments of 16-byte units, destination in any rank's memory ------------------
struct PushJob {
    int4* dst[4];
    const int4* src[4];
    unsigned n16[4];      // 16-byte units per segment (0 = unused)
};
// fused mode: block 0 also publishes `pub.n` doubles into row [rank] of a table in every arena, and the last CTA signals
// the phase (fft::peer_signal)
__global__ void __launch_bounds__(256) push_kernel(PushJob job, Publish pub, const __grid_constant__ fft::PeerSync sync) {
    const unsigned stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait(); pdl_trigger();
    fft::peer_wait(sync);
#pragma unroll
    for (int s = 0; s < 4; ++s)
        for (unsigned i = t0; i < job.n16[s]; i += stride) job.dst[s][i] = job.src[s][i];
    if (pub.n && blockIdx.x == 0) {
        const unsigned r = threadIdx.x / 4, k = threadIdx.x % 4;
        if (r < sync.n && k < pub.n) reinterpret_cast<double*>(sync.arena[r] + pub.table_offset)[sync.rank * pub.per_rank + k] = pub.src[k];
    }
    fft::peer_signal(sync);
}
// broadcast of a few doubles into slot [rank] of every peer's table
__global__ void push_scalars_kernel(PeerTable pt, size_t table_offset, unsigned per_rank, const double* __restrict__ src, unsigned n) {
    const unsigned r = threadIdx.x / 8, k = threadIdx.x % 8;
    pdl_wait(); pdl_trigger();
    if (r < pt.n && k < n) reinterpret_cast<double*>(pt.arena[r] + table_offset)[pt.rank * per_rank + k] = src[k];
}

// ---- flag barrier + rank-ordered reduction ----------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// All ranks launch this kernel the same number of times; the epoch is a device-side counter that every launch advances
// (so the launch can sit in a CUDA graph).  Steps of one launch:
//   1. publish: `n_pub` doubles of this rank go into row [rank] of a table in EVERY peer's arena (tiny all-gather);
//   2. barrier: lane r publishes the new epoch in rank r's flag slot of THIS rank and waits until rank r has published it
//      here; everything the peers stored before their barrier launch (earlier kernels of their streams, step 1) is visible;
//   3. reduce: rows [0, n_ranks) x `width` doubles of a table are summed in rank order into out[0..width) -- the tiny
//      all-reduces of the path, deterministic and identical on every rank.
// wait == 0: steps 1 and 3 only (single-process emulation of the ranks, where the launch order already orders the data).
// status[0] is set to 1 if a peer did not arrive within a few seconds (a crashed rank must not hang the GPU).
struct Reduce { const double* table; unsigned per_rank, width; double* out; };
// executed by one warp
__device__ __forceinline__ void barrier_body(const PeerTable& pt, size_t flags_offset, unsigned* __restrict__ d_epoch, int wait,
                                             const Publish& pub, const Reduce& red, unsigned* __restrict__ status) {
    const unsigned lane = threadIdx.x & 31;
    if (pub.n) {
        const unsigned r = lane / 4, k = lane % 4;          // up to 8 ranks x 4 values
        if (r < pt.n && k < pub.n) reinterpret_cast<double*>(pt.arena[r] + pub.table_offset)[pt.rank * pub.per_rank + k] = pub.src[k];
    }
    if (wait) {
        const unsigned epoch = *d_epoch + 1u;
        __threadfence_system();
        __syncwarp();
        if (lane < pt.n) {
            st_release_sys(reinterpret_cast<unsigned*>(pt.arena[lane] + flags_offset) + pt.rank, epoch);
            const unsigned* mine = reinterpret_cast<const unsigned*>(pt.arena[pt.rank] + flags_offset) + lane;
            const long long t0 = clock64();
            // epochs are compared as signed distances so that the counter may wrap
            while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
                if (clock64() - t0 > 8000000000LL) { atomicExch(status, 1u); break; }
                __nanosleep(32);
            }
        }
        __syncwarp();
        if (lane == 0) *d_epoch = epoch;
    }
    if (red.out && lane < red.width) {
        double s = 0.0;
        for (unsigned r = 0; r < pt.n; ++r) s += __ldcv(red.table + r * red.per_rank + lane);
        red.out[lane] = s;
    }
}
__global__ void barrier_kernel(PeerTable pt, size_t flags_offset, unsigned* __restrict__ d_epoch, int wait, Publish pub, Reduce red,
                               unsigned* __restrict__ status) {
    pdl_wait(); pdl_trigger();
    barrier_body(pt, flags_offset, d_epoch, wait, pub, red, status);
}

// halo push and the barrier that follows it in one launch: every CTA copies its share, fences its peer stores at system
// scope and takes a ticket; the last one runs the barrier (publish, flags, wait, reduce) in its first warp.  With 32 CTAs
// the per-CTA fence is cheap (it is what made the same idea too slow inside the 4096-CTA FFT sweeps).
__global__ void __launch_bounds__(256) push_barrier_kernel(PushJob job, PeerTable pt, size_t flags_offset, unsigned* __restrict__ d_epoch,
                                                           Publish pub, Reduce red, unsigned* __restrict__ status,
                                                           unsigned* __restrict__ ticket) {
    const unsigned stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ bool is_last;
    pdl_wait(); pdl_trigger();
#pragma unroll
    for (int s = 0; s < 4; ++s)
        for (unsigned i = t0; i < job.n16[s]; i += stride) job.dst[s][i] = job.src[s][i];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    if (threadIdx.x == 0) *ticket = 0;
    if (threadIdx.x < 32) barrier_body(pt, flags_offset, d_epoch, 1, pub, red, status);
}

}  // namespace p2p
}  // namespace metad
