// mesh.cu -- OrderParameterMesh plan object and C ABI (see include/metad_b200.h).
//
// Step structure of metad_mesh_cv (reference: OrderParameterMesh::getCurrentValue, OrderParameterMesh.cc:925-968):
//   [tile order: bin -> scan -> place]   only every `period` calls, or when particles drifted (amortised; see below)
//   spread                               assignParticles (:517-640) into the fixed-point mesh; also sum a^2 (m_mode_sq), sum a
//   x fwd (+ int -> float, DC removal), y fwd, z fused (+plane0), y inv, x inv   updateMeshes (:642-747) + computeCV (:866-923)
// metad_mesh_forces: gather              interpolateForces (:749-864)
//
// Tile order.  The spread and gather kernels visit the particles tile by tile through a permutation.  Because the
// density is accumulated in integers and every particle's cell is recomputed from its current position, ANY permutation
// gives bitwise the same result; the order only decides speed (particles that left their padded tile take a slow path).
// It is therefore rebuilt lazily: on the first call, when N changes, every `period` calls (metad_mesh_set key 0,
// default 32), and as soon as the device reports drifted particles or a cell near the fixed-point range (counters
// copied asynchronously to pinned host memory after every spread; no host synchronisation anywhere in a step).
//
// DC removal: the x pass subtracts the mean density (sum a / M) before the transforms.  The k = 0 mode is
// excluded from the CV (:892) and a constant offset of IFFT(G) cannot produce a force (the TSC derivative
// weights of the 27 taps sum to zero), so results are unchanged -- but without it the fp32 transforms carry a
// DC term ~sqrt(N) times larger than every other mode and the force mesh loses several digits.
//
// Triclinic boxes: the same kernels with sheared coordinates in the stencil (instantiations kSpTri / TRI), forces through
// the reciprocal lattice vectors; with the reference's literal in-cell offsets (knob 16, Geom::tq) the derivative weights
// do not sum to zero and the k = 0 mode is put back before the convolution (ConvParams::dc_restore).
// Mesh sizes the tiled kernels do not take (not a power of two, or outside 32 <= nx <= 1024, 16 <= ny, nz <= 512) run
// through the general path (mesh_general.cuh; general_create / general_cv / general_forces below).
//
// z-slab sharding (metad_mesh_slab_*): rank r owns the planes [r nz/P, (r+1) nz/P) and the particles inside them
// (reference: HOOMD domain decomposition + CommunicatorGrid ghost exchange + dfft, OrderParameterMesh.cc:263-315,
// 659-746).  Per step: local spread (the integer mesh carries one ghost plane per side), halo exchange of the two
// integer ghost planes, x FFT written directly in the layout of the slab -> kx-pencil all-to-all, y and fused z passes
// on the pencil (the packed kx = 0 column lives entirely on rank 0), all-to-all back, x inverse, halo-fill of one plane
// per side, local gather.  The collectives themselves (2 all-to-alls, 2 neighbour exchanges, 2 tiny all-reduces) are
// issued by the caller (NCCL via torch.distributed in sharded.MeshSlab); this file provides the five compute stages
// between them.  Integer halo addition makes the sharded density bitwise equal to the single-GPU one.
#include "mesh_kernels.cuh"
#include "mesh_fft_kernels.cuh"
#include "mesh_fft_xy.cuh"
#include "mesh_general.cuh"
#include "mesh_p2p.cuh"

#include <cuda.h>            // CUtensorMap + cuTensorMapEncodeTiled (resolved through cudaGetDriverEntryPoint, no -lcuda)
#include <algorithm>
#include <cmath>
#include <map>
#include <vector>

using namespace metad;
using namespace metad::mesh;
using namespace metad::fft;

// stages timed when profiling is on: 0 tile order (bin + scan + place; 0 ms on calls that reuse the order), 1 spread,
// 2 fft x fwd, 3 fft y fwd, 4 fft z fused, 5 fft y inv, 6 fft x inv, 7 gather
constexpr int kNumStages = 8;

struct metad_mesh {
    Geom g;                         // LOCAL geometry (slab: nz = planes of this rank)
    unsigned n_ranks = 1, rank = 0; // z-slab sharding
    unsigned nzg = 0;               // global planes
    unsigned kxl = 0;               // kx pencil width (complex) = nx/2/n_ranks
    int ntypes = 0;
    float* d_mode = nullptr;
    float amax = 0.f;               // largest |mode coefficient|
    // tile order
    unsigned cap = 0;
    // d_keys: cell key per particle; d_ranks: arrival rank inside the cell during a rebuild, afterwards THE TILE ORDER
    // (layer order, see mesh_layer_order_kernel); d_perm: cell-sorted permutation (intermediate)
    unsigned *d_keys = nullptr, *d_ranks = nullptr, *d_perm = nullptr;
    float4* d_cache4 = nullptr;     // particle cache spread -> gather (tile order): offsets + amplitude
    unsigned *d_count = nullptr, *d_start = nullptr, *d_block_sums = nullptr, *d_tstart = nullptr, *d_max_count = nullptr;
    bool order_valid = false;
    unsigned order_N = 0, calls_since_rebuild = 0, period = 32;
    unsigned long long n_rebuilds = 0;
    // mesh
    int* d_mesh_alloc = nullptr;    // integer density: nz planes (+ one ghost plane on each side in slab mode)
    int* d_mesh_i = nullptr;        // local plane 0 inside d_mesh_alloc
    long long* d_mesh64 = nullptr;  // 64-bit density (wide accumulation; allocated when first needed)
    // accumulator width: the 32-bit density shares its range between the resolution of one tap and the total of a cell, so
    // beyond ~11 |a|max of load in one cell the scale -- and with it the precision of the CV -- would have to drop; then
    // the plan switches to 64-bit accumulation (kSpWide).  Decided at a rebuild of the tile order from the largest cell
    // load: h_mode[0] (pinned) = 1: 32 bits suffice, 2: wide needed, 3: more particles per cell than the wide tile can
    // count.  The first rebuild waits for the answer; later ones read it one call late.
    bool wide = false;
    unsigned* h_mode = nullptr;
    // knob 9: the spread writes a per-particle cache for the gather.  Measured at C4 (profiles/r02_*): spread 0.219 + gather
    // 0.182 ms with the cache, 0.202 + 0.234 ms without (the gather then re-reads the positions through the tile order)
    bool cache = true;
    bool tma_flush = true;          // knob 10: flush the spread tile with 3-D tensor-map reductions
    int spread_debug = 0;           // knob 12: timing experiments (wrong results)
    // knob 15: x and y sweeps fused on whole z planes by thread-block clusters (mesh_fft_xy.cuh): three sweeps instead of five.
    // 0 = never, 1 = where it was measured faster (default), 2 = wherever an instantiation exists.  Measured on B200
    // (profiles/r02_notes.md): 128 x 128 planes (one CTA per plane) 0.0229 -> 0.0145 ms forward, 0.0244 -> 0.0148 ms inverse at C3;
    // 256 x 256 planes (clusters of 2 / 4 / 8 CTAs over distributed shared memory) 0.084 -> 0.087..0.107 ms forward, 0.070 ->
    // 0.089..0.111 ms inverse at C4: correct, but slower than the separate sweeps -- one or two CTAs per SM with long serial phases.
    int fuse_xy = 1;
    // knob 16: triclinic boxes -- 1 (default): the in-cell offsets carry the constant the reference's own code gives them
    // (Geom::tq in mesh_kernels.cuh: makeFraction(shift + lo) shears `lo` as well, OrderParameterMesh.cc:571-573); weight
    // is then lost exactly where the reference loses it.  0: geometrically correct assignment (partition of unity).
    bool tilt_literal = true;
    // knob 13: epilogues of the fused z sweep -- arg-max of |f_k|^2 (q*_max / sq_max log quantities, computeQmax) and the
    // k-space virial sums (computeVirial) with the tabulated kernel derivative of metad_mesh_set_table
    bool extras = false;
    bool use_table = false;
    unsigned n_table = 0;
    float* d_table_d = nullptr;
    double k_min = 0, k_max = 0;
    double* d_vir_partials = nullptr;
    unsigned long long* d_amax_key = nullptr;
    double* d_extras_out = nullptr;
    double box_L[3] = {0, 0, 0}, box_tilt[3] = {0, 0, 0};
    double box_b[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};      // reciprocal lattice vectors (rows), without 2 pi
    unsigned extras_N_global = 0;
    bool tma_gather = true;         // knob 11: load the gather tile with one 3-D tensor-map copy when it does not wrap
    alignas(64) CUtensorMap tmap_mesh = {};     // integer mesh (incl. ghost planes), box = padded tile
    alignas(64) CUtensorMap tmap_inv = {};      // Re IFFT(G) (d_buf), box = padded tile
    bool have_tmaps = false;
    float* d_buf = nullptr;         // M_local floats: packed half spectrum -> Re IFFT(G)
    float* d_rho_keep = nullptr;    // optional copy of rho (introspection)
    float2 *d_twx = nullptr, *d_twy = nullptr, *d_twz = nullptr;
    float* d_fx = nullptr;          // {scale, 1/scale} of the fixed-point density (+ a 16-byte copy of 1/scale at [4])
    double* d_sums = nullptr;       // [0] sum a^2  [1] sum a  [2] particles outside the slab
    double* d_tile_sums = nullptr;
    unsigned* d_counters = nullptr; // [0] ticket [1] drifted particles [2] particles outside the slab [3] cells near the range limit
    unsigned* h_counters = nullptr; // pinned copy of d_counters, refreshed asynchronously after every spread
    double* d_partials = nullptr;
    unsigned* d_ticket = nullptr;
    unsigned n_partials = 0;
    // peer-memory mode of the slab path (mesh_p2p.cuh): arena of this rank, the peers' arenas mapped through CUDA IPC
    char* arena = nullptr;
    p2p::ArenaLayout lay = {};
    p2p::PeerTable peers = {};
    bool peers_mapped[p2p::kMaxPeers] = {};     // opened with cudaIpcOpenMemHandle (to be closed)
    bool p2p_ready = false;
    unsigned* d_epoch = nullptr;                // barrier epoch (device-side counter: launches can be replayed from a graph)
    bool last_cv_fused = false;                 // the last metad_mesh_slab_p2p_cv published phase 3 for the gather to wait on
    // signal / wait folded into the producer / consumer kernels instead of barrier launches (knob 5).  Off by default:
    // measured slower on B200 (2 GPUs, C4: 0.506 vs 0.464 ms/step) -- every producer CTA has to fence its peer stores at
    // system scope before taking the ticket, which stalls the CTA tails that otherwise drain asynchronously.
    bool fused_sync = false;
    bool pdl = true;                            // programmatic dependent launch of the per-step kernels (knob 7)
    int order_kind = 1;                         // order inside a tile: 0 = layer order, 1 = bank order (knob 6)
    unsigned* d_sync = nullptr;                 // [4] phase epochs + [4] CTA tickets of the fused synchronisation, [8] ticket of push_barrier_kernel
    // halo push and the barrier after it in one launch (knob 8).  Off: measured slower (2 GPUs: C4 0.392 vs 0.381 ms, C3 0.127 vs
    // 0.118 ms) -- the per-CTA system-scope fence + ticket costs more than the kernel boundary it removes.
    bool merge_push = false;
    // CUDA-graph replay of the per-call kernel sequence (metad_mesh_set key 4): everything a call enqueues after the
    // (eager) tile-order decision is captured once per argument signature and replayed with one launch
    bool graph_mode = false;
    struct GraphKey { const void* postype; unsigned N, N_global; double L[3], tilt[3]; const void* d_cv; cudaStream_t stream; int kind; bool keep_rho, keep_cells; int variant; };
    GraphKey gkey = {};
    int gwarm = 0;
    cudaGraphExec_t gexec = nullptr;
    cudaStream_t capture_stream = nullptr;
    unsigned long long n_graph_launches = 0;
    double* d_sums_global = nullptr;            // [4] sums over all ranks
    double* d_cv_partial = nullptr;
    unsigned* d_p2p_status = nullptr;
    // general path (mesh_general.cuh): mesh sizes the tiled kernels do not take (not a power of two, or outside their range)
    bool general = false;
    float2* d_spec = nullptr;                   // complex mesh: density -> spectrum -> G -> IFFT(G)
    int* d_cells = nullptr;                     // knob 3: reported cell of every particle, int[3 N]
    unsigned cells_cap = 0;
    double* d_gen_sums = nullptr;               // per-block partials of sum a^2, sum a
    float2* d_gtw[3] = {nullptr, nullptr, nullptr};
    meshgen::Radices grad[3] = {};
    // state
    bool have_cv = false;
    unsigned last_N = 0;
    bool keep_rho = false, keep_cells = false;
    // optional per-stage timing (CUDA events on the caller's stream): ev[i] = start of stage i for i < 7, ev[7] = end of the
    // cv pipeline, ev[8]/ev[9] = start/end of the gather
    bool profile = false;
    cudaEvent_t evp[13] = {};                   // peer-memory step, profiling: boundaries of its 12 segments (metad_mesh_get 8)
    cudaEvent_t ev[kNumStages + 2] = {};
    size_t M() const { return (size_t)g.nx * g.ny * g.nz; }
};

namespace {

bool is_pow2(unsigned n) { return n && !(n & (n - 1)); }
unsigned ilog2(unsigned n) { unsigned l = 0; while ((1u << l) < n) ++l; return l; }

int upload_twiddles(float2** dst, unsigned n) {
    std::vector<float2> t(n);
    for (unsigned k = 0; k < n; ++k) {
        const double ph = -2.0 * M_PI * (double)k / (double)n;
        t[k] = make_float2((float)cos(ph), (float)sin(ph));
    }
    METAD_CUDA(cudaMalloc(dst, sizeof(float2) * n));
    METAD_CUDA(cudaMemcpy(*dst, t.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
    return METAD_OK;
}

// opt-in dynamic shared memory, once per kernel (the attribute call must not sit inside a stream capture)
template <class K> int set_smem(K kernel, size_t bytes) {
    static std::map<const void*, size_t> configured;
    if (bytes <= 48 * 1024) return METAD_OK;
    size_t& have = configured[(const void*)kernel];
    if (have < bytes) {
        METAD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        have = bytes;
    }
    return METAD_OK;
}

// ---- FFT launchers -----------------------------------------------------------------------------------
// x pass over the local rows.  io: nullptr = the plan's own buffer; otherwise the packed all-to-all buffer (output of
// the forward pass / input of the inverse pass), [part][row][kx in part] with parts of width kxl.
// Forward: consumes (and clears) the integer density; d_sums = (global) sums, d_ghost = received halo planes (slab).
template <int LC> int run_x(metad_mesh* p, bool inverse, float2* io, const double* d_sums, const int* d_ghost, cudaStream_t st,
                            const PeerOut* peer_out = nullptr, const PeerSync* sync = nullptr, const double* table = nullptr,
                            double* d_out = nullptr) {
    PeerSync ps;
    memset(&ps, 0, sizeof ps);
    if (sync) ps = *sync;
    const size_t smem = sizeof(float2) * (LayoutRow::size(LC) + 2 * LC);
    const unsigned rows = p->g.ny * p->g.nz;
    float2* buf = reinterpret_cast<float2*>(p->d_buf);
    const unsigned lg_part = (io || peer_out) ? ilog2(p->kxl) : ilog2(LC);
    if (!inverse) {
        int rc = set_smem(fft_x_fwd_kernel<LC, false>, smem); if (rc) return rc;
        DensityIn in;
        in.mesh = reinterpret_cast<const int2*>(p->d_mesh_i);
        in.d_fx = p->d_fx;
        in.d_sums = d_sums;
        in.inv_cells = 1.0 / ((double)p->g.nx * (double)p->g.ny * (double)p->nzg);
        in.ghost = reinterpret_cast<const int2*>(d_ghost);
        in.lgy = p->g.lgy; in.nz = p->g.nz;
        in.sums_table = sync ? table : nullptr;
        in.sums_out = d_out;
        in.rho_keep = p->keep_rho ? reinterpret_cast<float2*>(p->d_rho_keep) : nullptr;
        // the kernel clears the rows it consumes (and, in peer-memory mode, this rank's two ghost planes, which were
        // pushed to the neighbours before): the accumulator is empty again for the next spread
        in.zero = reinterpret_cast<int4*>(p->d_mesh_i);
        in.zero_lo = peer_out ? reinterpret_cast<int4*>(p->d_mesh_alloc) : nullptr;
        in.zero_hi = peer_out ? reinterpret_cast<int4*>(p->d_mesh_alloc + ((size_t)p->g.nz + 1) * p->g.nx * p->g.ny) : nullptr;
        in.mesh64 = reinterpret_cast<const longlong2*>(p->d_mesh64);
        in.zero64 = reinterpret_cast<int4*>(p->d_mesh64);
        in.range_counter = p->d_counters + 6;
        in.h_range = p->h_counters + 3;
        PeerOut po;
        memset(&po, 0, sizeof po);
        if (peer_out) po = *peer_out;
        if (p->wide) {
            rc = set_smem(fft_x_fwd_kernel<LC, true>, smem); if (rc) return rc;
            METAD_CUDA(launch_pdl(p->pdl, fft_x_fwd_kernel<LC, true>, rows / kLines, kLines * LC / kE, smem, st, in, p->d_twx, io ? io : buf, lg_part, rows, po, ps));
        } else {
            METAD_CUDA(launch_pdl(p->pdl, fft_x_fwd_kernel<LC, false>, rows / kLines, kLines * LC / kE, smem, st, in, p->d_twx, io ? io : buf, lg_part, rows, po, ps));
        }
        METAD_LAUNCH_CHECK();
    } else {
        int rc = set_smem(fft_x_inv_kernel<LC>, smem); if (rc) return rc;
        METAD_CUDA(launch_pdl(p->pdl, fft_x_inv_kernel<LC>, rows / kLines, kLines * LC / kE, smem, st, buf, p->d_twx, io ? io : buf, lg_part, rows, ps, sync ? table : nullptr, d_out));
    }
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}
// y pass on buf = [nz_rows][ny][row_len]
template <int L> int run_y(metad_mesh* p, bool inverse, float2* buf, unsigned row_len, unsigned nz_rows, cudaStream_t st,
                           const PeerOut* peer_out = nullptr, const PeerSync* sync = nullptr) {
    PeerSync ps;
    memset(&ps, 0, sizeof ps);
    if (sync) ps = *sync;
    PeerOut po;
    memset(&po, 0, sizeof po);
    if (peer_out) po = *peer_out;
    const unsigned lg_planes = ilog2(p->g.nz);
    // Two groups of 16 lines per tile (256-byte rows) are implemented (G = 2) but measured SLOWER on B200 (C4: 48.5 / 39.1 us
    // against 38.9 / 34.8 us for the forward / inverse sweep: 1024-thread CTAs, two per SM), so they stay off.
    constexpr bool can_wide = false;
    if (can_wide && row_len % (2 * kLines) == 0) {
        constexpr int G = can_wide ? 2 : 1;
        const size_t smem = sizeof(float2) * (LayoutColWide<G>::size(L) + L);
        dim3 grid(row_len / (G * kLines), nz_rows);
        if (!inverse) {
            int rc = set_smem(fft_y_kernel<L, -1, G>, smem); if (rc) return rc;
            METAD_CUDA(launch_pdl(p->pdl, fft_y_kernel<L, -1, G>, grid, G * kLines * L / kE, smem, st, buf, p->d_twy, row_len, po, lg_planes, ps));
        } else {
            int rc = set_smem(fft_y_kernel<L, +1, G>, smem); if (rc) return rc;
            METAD_CUDA(launch_pdl(p->pdl, fft_y_kernel<L, +1, G>, grid, G * kLines * L / kE, smem, st, buf, p->d_twy, row_len, po, lg_planes, ps));
        }
    } else {
        const size_t smem = sizeof(float2) * (LayoutColWide<1>::size(L) + L);
        dim3 grid(row_len / kLines, nz_rows);
        if (!inverse) {
            int rc = set_smem(fft_y_kernel<L, -1, 1>, smem); if (rc) return rc;
            METAD_CUDA(launch_pdl(p->pdl, fft_y_kernel<L, -1, 1>, grid, kLines * L / kE, smem, st, buf, p->d_twy, row_len, po, lg_planes, ps));
        } else {
            int rc = set_smem(fft_y_kernel<L, +1, 1>, smem); if (rc) return rc;
            METAD_CUDA(launch_pdl(p->pdl, fft_y_kernel<L, +1, 1>, grid, kLines * L / kE, smem, st, buf, p->d_twy, row_len, po, lg_planes, ps));
        }
    }
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}
// fused z pass on buf = [nzg][ny][row_len]; d_sums: (global) sum a^2; d_cv receives 0.5 * (local) energy sum
template <int L> int run_z(metad_mesh* p, float2* buf, unsigned row_len, unsigned kx_off, const double* d_sums, unsigned N_global,
                           double* d_cv, cudaStream_t st, bool publish_cv = false) {
    const size_t smem = sizeof(float2) * (LayoutCol::size(L) + L);
    const unsigned ny = p->g.ny;
    ConvParams cp;
    cp.nx = p->g.nx; cp.ny = ny; cp.nz = p->nzg;
    cp.row_len = row_len; cp.kx_off = kx_off;
    cp.inv_n = (float)(1.0 / (double)N_global);
    cp.n_global = (double)N_global;
    cp.d_mode_sq = d_sums;
    cp.dc_restore = (p->g.tri && (p->g.tq[0] != 0.f || p->g.tq[1] != 0.f)) ? 1 : 0;
    cp.inv_cells = 1.0 / ((double)p->g.nx * (double)p->g.ny * (double)p->nzg);
    cp.partials = p->d_partials;
    cp.ticket = p->d_ticket;
    cp.n_blocks_plane0 = kx_off == 0 ? (ny / 2 + 1 + kLines / 2 - 1) / (kLines / 2) : 0;
    cp.d_cv = d_cv;
    memset(cp.cv_arena, 0, sizeof cp.cv_arena);
    cp.cv_off = 0; cp.cv_n = 0; cp.cv_rank = 0;
    if (publish_cv) {
        for (unsigned r = 0; r < p->n_ranks; ++r) cp.cv_arena[r] = p->peers.arena[r];
        cp.cv_off = p->lay.cv; cp.cv_n = p->n_ranks; cp.cv_rank = p->rank;
    }
    cp.extras = p->extras ? 1 : 0;
    cp.use_table = (p->use_table && p->d_table_d && p->n_table >= 2) ? 1 : 0;
    cp.n_table = p->n_table; cp.table_d = p->d_table_d;
    cp.k_min = (float)p->k_min; cp.k_max = (float)p->k_max;
    cp.delta_k = p->n_table >= 2 ? (float)((p->k_max - p->k_min) / (double)(p->n_table - 1)) : 1.0f;
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c) cp.bk[3 * i + c] = (float)(2.0 * M_PI * p->box_b[i][c]);
    cp.vir_partials = p->d_vir_partials; cp.amax_key = p->d_amax_key; cp.extras_out = p->d_extras_out;
    if (p->extras) {
        if (!p->d_vir_partials) {
            METAD_CUDA(cudaMalloc(&p->d_vir_partials, sizeof(double) * 6 * p->n_partials));
            METAD_CUDA(cudaMalloc(&p->d_amax_key, sizeof(unsigned long long)));
            METAD_CUDA(cudaMalloc(&p->d_extras_out, sizeof(double) * 8));
            cp.vir_partials = p->d_vir_partials; cp.amax_key = p->d_amax_key; cp.extras_out = p->d_extras_out;
        }
        METAD_CUDA(cudaMemsetAsync(p->d_amax_key, 0, sizeof(unsigned long long), st));
        p->extras_N_global = N_global;
    }
    const unsigned nblocks = cp.n_blocks_plane0 + (row_len / kLines) * ny;
    // experiment (METAD_Z_CTAS=4): four CTAs per SM (32 registers) instead of three (40 registers) for 512-thread CTAs
    static const bool four = getenv("METAD_Z_CTAS") && atoi(getenv("METAD_Z_CTAS")) == 4;
    if (four && kLines * L / kE == 256) {
        int rc4 = set_smem(fft_z_fused_kernel<L, 4>, smem); if (rc4) return rc4;
        METAD_CUDA(launch_pdl(p->pdl, fft_z_fused_kernel<L, 4>, nblocks, kLines * L / kE, smem, st, buf, p->d_twz, cp));
        METAD_LAUNCH_CHECK();
        return METAD_OK;
    }
    if (p->extras) {
        constexpr int MB = (kLines * L / kE) <= 256 ? 2 : 1;
        int rcx = set_smem(fft_z_fused_kernel<L, MB, true>, smem); if (rcx) return rcx;
        METAD_CUDA(launch_pdl(p->pdl, fft_z_fused_kernel<L, MB, true>, nblocks, kLines * L / kE, smem, st, buf, p->d_twz, cp));
        METAD_LAUNCH_CHECK();
        return METAD_OK;
    }
    int rc = set_smem(fft_z_fused_kernel<L>, smem); if (rc) return rc;
    METAD_CUDA(launch_pdl(p->pdl, fft_z_fused_kernel<L>, nblocks, kLines * L / kE, smem, st, buf, p->d_twz, cp));
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}

#define METAD_DISPATCH_LEN(n, EXPR)                                       \
    switch (n) {                                                          \
        case 16: { constexpr int LL = 16; rc = EXPR; } break;             \
        case 32: { constexpr int LL = 32; rc = EXPR; } break;             \
        case 64: { constexpr int LL = 64; rc = EXPR; } break;             \
        case 128: { constexpr int LL = 128; rc = EXPR; } break;           \
        case 256: { constexpr int LL = 256; rc = EXPR; } break;           \
        case 512: { constexpr int LL = 512; rc = EXPR; } break;           \
        default: set_error("cv.mesh: unsupported mesh dimension"); rc = METAD_ERR_UNSUPPORTED; \
    }

int mark(metad_mesh* p, int i, cudaStream_t st) {
    if (!p->profile) return METAD_OK;
    if (!p->ev[i]) METAD_CUDA(cudaEventCreate(&p->ev[i]));
    METAD_CUDA(cudaEventRecord(p->ev[i], st));
    return METAD_OK;
}

int markp(metad_mesh* p, int i, cudaStream_t st) {
    if (!p->profile) return METAD_OK;
    if (!p->evp[i]) METAD_CUDA(cudaEventCreate(&p->evp[i]));
    METAD_CUDA(cudaEventRecord(p->evp[i], st));
    return METAD_OK;
}

// fused x + y sweeps of one unsharded plan; inverse = false: density -> [z][y][kx] spectrum, true: spectrum -> real rows in place
template <int LC, int LY, int C, int NT> int run_xy(metad_mesh* p, bool inverse, cudaStream_t st) {
    using P = XYPlan<LC, LY, C, NT>;
    float2* buf = reinterpret_cast<float2*>(p->d_buf);
    const dim3 grid(C, p->g.nz);
    if (inverse) {
        int rc = set_smem(fft_xy_inv_kernel<LC, LY, C, NT>, P::smem_bytes); if (rc) return rc;
        METAD_CUDA(launch_cluster_pdl(p->pdl, C, fft_xy_inv_kernel<LC, LY, C, NT>, grid, P::NT, P::smem_bytes, st, buf, p->d_twx, p->d_twy));
        METAD_LAUNCH_CHECK();
        return METAD_OK;
    }
    DensityIn in;
    memset(&in, 0, sizeof in);
    in.mesh = reinterpret_cast<const int2*>(p->d_mesh_i);
    in.d_fx = p->d_fx;
    in.d_sums = p->d_sums;
    in.inv_cells = 1.0 / ((double)p->g.nx * (double)p->g.ny * (double)p->nzg);
    in.lgy = p->g.lgy; in.nz = p->g.nz;
    in.zero = reinterpret_cast<int4*>(p->d_mesh_i);
    in.range_counter = p->d_counters + 6;
    in.h_range = p->h_counters + 3;
    int rc = set_smem(fft_xy_fwd_kernel<LC, LY, C, NT>, P::smem_bytes); if (rc) return rc;
    METAD_CUDA(launch_cluster_pdl(p->pdl, C, fft_xy_fwd_kernel<LC, LY, C, NT>, grid, P::NT, P::smem_bytes, st, in, p->d_twx, p->d_twy, buf));
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}
// plane shapes with a fused instantiation: (nx/2, ny) -> cluster size
bool can_fuse_xy(const metad_mesh* p) {
    if (!p->fuse_xy || p->g.slab || p->keep_rho || p->wide) return false;
    const unsigned lc = p->g.nx / 2, ly = p->g.ny;
    if (lc == 64 && ly == 128) return true;
    return p->fuse_xy >= 2 && ((lc == 128 && ly == 256) || (lc == 256 && ly == 512));
}
int dispatch_xy(metad_mesh* p, bool inverse, cudaStream_t st) {
    const unsigned lc = p->g.nx / 2;
    // cluster size / CTA size per plane shape (METAD_XY_VARIANT selects the alternatives for experiments)
    static const int variant = getenv("METAD_XY_VARIANT") ? atoi(getenv("METAD_XY_VARIANT")) : 0;
    if (lc == 128) {        // default: the fastest of the measured shapes (profiles/r02_notes.md)
        if (variant == 1) return run_xy<128, 256, 2, 512>(p, inverse, st);
        if (variant == 2) return run_xy<128, 256, 8, 256>(p, inverse, st);
        if (variant == 3) return run_xy<128, 256, 4, 512>(p, inverse, st);
        return run_xy<128, 256, 4, 256>(p, inverse, st);
    }
    if (lc == 64) {
        if (variant == 1) return run_xy<64, 128, 2, 256>(p, inverse, st);
        if (variant == 2) return run_xy<64, 128, 2, 128>(p, inverse, st);
        return run_xy<64, 128, 1, 512>(p, inverse, st);
    }
    return run_xy<256, 512, 8, 512>(p, inverse, st);
}

int fft_pipeline(metad_mesh* p, unsigned N_global, double* d_cv, cudaStream_t st) {
    int rc = METAD_OK;
    float2* buf = reinterpret_cast<float2*>(p->d_buf);
    const unsigned nxh = p->g.nx / 2;
    if (p->keep_rho && !p->d_rho_keep) METAD_CUDA(cudaMalloc(&p->d_rho_keep, sizeof(float) * p->M()));
    if (can_fuse_xy(p)) {       // three sweeps: xy forward, fused z, xy inverse (stage timers: the y slots stay empty)
        rc = mark(p, 2, st); if (rc) return rc;
        rc = dispatch_xy(p, false, st); if (rc) return rc;
        rc = mark(p, 3, st); if (rc) return rc;
        rc = mark(p, 4, st); if (rc) return rc;
        METAD_DISPATCH_LEN(p->g.nz, (run_z<LL>(p, buf, nxh, 0, p->d_sums, N_global, d_cv, st))); if (rc) return rc;
        rc = mark(p, 5, st); if (rc) return rc;
        rc = mark(p, 6, st); if (rc) return rc;
        rc = dispatch_xy(p, true, st); if (rc) return rc;
        return mark(p, 7, st);
    }
    rc = mark(p, 2, st); if (rc) return rc;
    METAD_DISPATCH_LEN(nxh, (run_x<LL>(p, false, nullptr, p->d_sums, nullptr, st))); if (rc) return rc;
    rc = mark(p, 3, st); if (rc) return rc;
    METAD_DISPATCH_LEN(p->g.ny, (run_y<LL>(p, false, buf, nxh, p->g.nz, st))); if (rc) return rc;
    rc = mark(p, 4, st); if (rc) return rc;
    METAD_DISPATCH_LEN(p->g.nz, (run_z<LL>(p, buf, nxh, 0, p->d_sums, N_global, d_cv, st))); if (rc) return rc;
    rc = mark(p, 5, st); if (rc) return rc;
    METAD_DISPATCH_LEN(p->g.ny, (run_y<LL>(p, true, buf, nxh, p->g.nz, st))); if (rc) return rc;
    rc = mark(p, 6, st); if (rc) return rc;
    METAD_DISPATCH_LEN(nxh, (run_x<LL>(p, true, nullptr, nullptr, nullptr, st))); if (rc) return rc;
    rc = mark(p, 7, st); if (rc) return rc;
    return METAD_OK;
}

int ensure_capacity(metad_mesh* p, unsigned N) {
    if (N <= p->cap) return METAD_OK;
    cudaFree(p->d_keys); cudaFree(p->d_ranks); cudaFree(p->d_perm); cudaFree(p->d_cache4);
    p->d_keys = p->d_ranks = p->d_perm = nullptr; p->d_cache4 = nullptr; p->cap = 0;
    p->order_valid = false;
    const unsigned cap = N + N / 16 + 1024;
    METAD_CUDA(cudaMalloc(&p->d_keys, sizeof(unsigned) * cap));
    METAD_CUDA(cudaMalloc(&p->d_ranks, sizeof(unsigned) * cap));
    METAD_CUDA(cudaMalloc(&p->d_perm, sizeof(unsigned) * cap));
    if (p->cache) {
        METAD_CUDA(cudaMalloc(&p->d_cache4, sizeof(float4) * cap));
    }
    p->cap = cap;
    return METAD_OK;
}

// reciprocal lattice vectors b_i = (a_j x a_k) / V of the box (rows of `b`, without 2 pi), a_1 = (Lx, 0, 0), a_2 = (xy Ly, Ly, 0),
// a_3 = (xz Lz, yz Lz, Lz)  (OrderParameterMesh.cc:362-369, 761-769)
void reciprocal_vectors(const double* L, const double* tilt, double (&b)[3][3]) {
    const double a1[3] = {L[0], 0.0, 0.0}, a2[3] = {tilt[0] * L[1], L[1], 0.0}, a3[3] = {tilt[1] * L[2], tilt[2] * L[2], L[2]};
    const double V = L[0] * L[1] * L[2];
    auto cross = [&](const double* u, const double* v, double* o) {
        o[0] = (u[1] * v[2] - u[2] * v[1]) / V; o[1] = (u[2] * v[0] - u[0] * v[2]) / V; o[2] = (u[0] * v[1] - u[1] * v[0]) / V;
    };
    cross(a2, a3, b[0]); cross(a3, a1, b[1]); cross(a1, a2, b[2]);
}

// n_a b_a of the force interpolation (ForceParams, mesh_kernels.cuh); an orthorhombic box keeps the plain quotients n / L
void set_force_matrix(ForceParams& fp, const Geom& g, const metad_box* box) {
    memset(fp.nb1, 0, sizeof fp.nb1); memset(fp.nb2, 0, sizeof fp.nb2); memset(fp.nb3, 0, sizeof fp.nb3);
    if (box->tilt[0] == 0.0 && box->tilt[1] == 0.0 && box->tilt[2] == 0.0) {
        fp.nb1[0] = (float)((double)g.nx / box->L[0]);
        fp.nb2[1] = (float)((double)g.ny / box->L[1]);
        fp.nb3[2] = (float)((double)g.nzg / box->L[2]);
        return;
    }
    double b[3][3];
    reciprocal_vectors(box->L, box->tilt, b);
    for (int c = 0; c < 3; ++c) {
        fp.nb1[c] = (float)((double)g.nx * b[0][c]);
        fp.nb2[c] = (float)((double)g.ny * b[1][c]);
        fp.nb3[c] = (float)((double)g.nzg * b[2][c]);
    }
}

int set_box(metad_mesh* p, const metad_box* box) {
    for (int i = 0; i < 3; ++i) METAD_REQUIRE(box->L[i] > 0.0, "cv.mesh: box lengths must be positive");
    geom_set_box(p->g, box->L, box->tilt, p->tilt_literal);          // the box is the GLOBAL box
    for (int i = 0; i < 3; ++i) { p->box_L[i] = box->L[i]; p->box_tilt[i] = box->tilt[i]; }
    reciprocal_vectors(box->L, box->tilt, p->box_b);
    return METAD_OK;
}

// counting sort of the particles by tile-major cell key -> perm, tstart; fixed-point scale for the following calls
int rebuild_order(metad_mesh* p, const float* d_postype, unsigned N, cudaStream_t stream, bool sync_mode) {
    const Geom& g = p->g;
    const size_t M = p->M();
    const int sms = device_sm_count();
    METAD_CUDA(cudaMemsetAsync(p->d_max_count, 0, sizeof(unsigned), stream));
    long nbp = ((long)N + kBinThreads * 4L - 1) / (kBinThreads * 4L);
    if (nbp > sms * 16L) nbp = sms * 16L;
    if (nbp < 1) nbp = 1;
    mesh_bin_kernel<<<(int)nbp, kBinThreads, 0, stream>>>((const float4*)d_postype, N, g, p->d_mode, p->ntypes, p->d_keys, p->d_ranks,
                                                         p->d_count, p->d_max_count);
    METAD_LAUNCH_CHECK();
    if (M % (16 * kScanThreads) == 0) {
        const unsigned nb_scan = (unsigned)(M / (16 * kScanThreads));
        scan_reduce_kernel<4><<<nb_scan, kScanThreads, 0, stream>>>((const uint4*)p->d_count, p->d_block_sums);
        METAD_LAUNCH_CHECK();
        scan_offsets_kernel<<<1, kScanThreads, 0, stream>>>(p->d_block_sums, nb_scan);
        METAD_LAUNCH_CHECK();
        scan_apply_kernel<4><<<nb_scan, kScanThreads, 0, stream>>>((uint4*)p->d_count, p->d_block_sums, p->d_start, (unsigned)M);
        METAD_LAUNCH_CHECK();
    } else {
        const unsigned nb_scan = (unsigned)(M / (4 * kScanThreads));
        scan_reduce_kernel<1><<<nb_scan, kScanThreads, 0, stream>>>((const uint4*)p->d_count, p->d_block_sums);
        METAD_LAUNCH_CHECK();
        scan_offsets_kernel<<<1, kScanThreads, 0, stream>>>(p->d_block_sums, nb_scan);
        METAD_LAUNCH_CHECK();
        scan_apply_kernel<1><<<nb_scan, kScanThreads, 0, stream>>>((uint4*)p->d_count, p->d_block_sums, p->d_start, (unsigned)M);
        METAD_LAUNCH_CHECK();
    }
    mesh_place_kernel<<<(int)nbp, kBinThreads, 0, stream>>>(N, p->d_keys, p->d_ranks, p->d_start, p->d_perm, num_tiles(g), 3 * g.lgT,
                                                           p->d_tstart);
    METAD_LAUNCH_CHECK();
    // layer order inside every tile; the arrival ranks are no longer needed, their buffer receives the final order
    if (p->order_kind == 1) {
        if (g.lgT == 4) mesh_bank_order_kernel<4><<<num_tiles(g), kLayerThreads, 0, stream>>>(p->d_start, p->d_perm, p->d_ranks);
        else mesh_bank_order_kernel<3><<<num_tiles(g), kLayerThreads, 0, stream>>>(p->d_start, p->d_perm, p->d_ranks);
    } else {
        if (g.lgT == 4) mesh_layer_order_kernel<4><<<num_tiles(g), kLayerThreads, 0, stream>>>(p->d_start, p->d_perm, p->d_ranks);
        else mesh_layer_order_kernel<3><<<num_tiles(g), kLayerThreads, 0, stream>>>(p->d_start, p->d_perm, p->d_ranks);
    }
    METAD_LAUNCH_CHECK();
    // accumulator width for the calls until the next rebuild (see metad_mesh::wide).  The first rebuild of a particle set
    // waits for the device's answer (nothing is in flight that early); afterwards the answer of rebuild k is read by the
    // host before rebuild k+1 at the latest (prepare_order), so a step never synchronises.
    const bool can_wide = !p->g.slab;
    if (sync_mode) {
        mesh_fx_mode_kernel<<<1, 1, 0, stream>>>(p->d_max_count, p->amax, p->h_mode);
        METAD_LAUNCH_CHECK();
        METAD_CUDA(cudaStreamSynchronize(stream));
        if (p->h_mode[0] == 3 && can_wide) {
            set_error("cv.mesh: more than 38 000 particles in one mesh cell -- beyond the range of the fixed-point density");
            return METAD_ERR_UNSUPPORTED;
        }
        p->wide = can_wide && p->h_mode[0] >= 2;
    }
    if (p->wide && !p->d_mesh64) {
        METAD_CUDA(cudaMalloc(&p->d_mesh64, sizeof(long long) * M));
        METAD_CUDA(cudaMemsetAsync(p->d_mesh64, 0, sizeof(long long) * M, stream));
    }
    mesh_fx_scale_kernel<<<1, 1, 0, stream>>>(p->d_max_count, p->amax, p->wide ? 1 : 0, p->d_fx, p->h_mode);
    METAD_LAUNCH_CHECK();
    p->order_valid = true;
    p->order_N = N;
    p->calls_since_rebuild = 0;
    ++p->n_rebuilds;
    return METAD_OK;
}

template <int LGT> size_t tile_smem_bytes() {
    return sizeof(int) * (size_t)((1 << LGT) + 2 * kHaloX) * ((1 << LGT) + 2 * kHalo) * ((1 << LGT) + 2 * kHalo);
}
// gather: the tile plus two staging buffers of the particle cache (float4 + uint2 per thread)
template <int LGT> size_t spread_smem_bytes(int ntypes, int flags) {
    return sizeof(int) * (size_t)spread_tile_words<LGT>(flags) + kSpreadStages * kSpreadThreads * sizeof(float4) + sizeof(float) * ntypes;
}
// staging buffers of the gather: cache entries (float4 + uint2 per thread and stage), or positions + the mode table
size_t gather_stage_bytes(int threads, bool cache, int ntypes) {
    return (size_t)kGatherStages * threads * (sizeof(float4) + (cache ? sizeof(unsigned) : 0)) + sizeof(float) * ntypes;
}

template <int LGT, int FLAGS>
int launch_spread(metad_mesh* p, const float* d_postype, const SpreadOut& out, cudaStream_t stream) {
    const Geom& g = p->g;
    int rc = set_smem(mesh_spread_kernel<LGT, FLAGS>, spread_smem_bytes<LGT>(kSpreadModes, FLAGS)); if (rc) return rc;
    METAD_CUDA(launch_pdl(p->pdl, mesh_spread_kernel<LGT, FLAGS>, num_tiles(g), kSpreadThreads, spread_smem_bytes<LGT>(p->ntypes, FLAGS), stream,
                          (const float4*)d_postype, p->d_ranks, p->d_tstart, g, p->d_mode, p->ntypes, p->d_fx, out));
    return METAD_OK;
}
template <int LGT>
int dispatch_spread(metad_mesh* p, int flags, const float* d_postype, const SpreadOut& out, cudaStream_t stream) {
    switch (flags) {
#define METAD_SP(F) case F: return launch_spread<LGT, F>(p, d_postype, out, stream);
        METAD_SP(0) METAD_SP(1) METAD_SP(2) METAD_SP(3) METAD_SP(4) METAD_SP(5) METAD_SP(6) METAD_SP(7)
        METAD_SP(8) METAD_SP(9) METAD_SP(10) METAD_SP(11)
        // triclinic boxes (kSpTri): row flush only
        METAD_SP(16) METAD_SP(17) METAD_SP(18) METAD_SP(19) METAD_SP(20) METAD_SP(21) METAD_SP(22) METAD_SP(23)
#undef METAD_SP
        default: set_error("spread: unknown kernel variant"); return METAD_ERR_INVALID;
    }
}

// tensor maps of the integer mesh and of Re IFFT(G) with the padded tile as box (the driver entry point is resolved at run
// time: the library does not link libcuda)
int ensure_tmaps(metad_mesh* p) {
    if (p->have_tmaps) return METAD_OK;
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    METAD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return METAD_ERR_CUDA; }
    const Geom& g = p->g;
    const unsigned T = 1u << g.lgT;
    const cuuint32_t box[3] = {T + 2 * kHaloX, T + 2 * kHalo, T + 2 * kHalo}, estr[3] = {1, 1, 1};
    const cuuint64_t planes = (cuuint64_t)g.nz + (g.slab ? 2 : 0);
    const cuuint64_t dims_mesh[3] = {g.nx, g.ny, planes}, dims_inv[3] = {g.nx, g.ny, g.nz};
    const cuuint64_t strides[2] = {(cuuint64_t)g.nx * 4, (cuuint64_t)g.nx * g.ny * 4};
    CUresult r1 = ((EncodeFn)fn)(&p->tmap_mesh, CU_TENSOR_MAP_DATA_TYPE_INT32, 3, p->d_mesh_alloc, dims_mesh, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = ((EncodeFn)fn)(&p->tmap_inv, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p->d_buf, dims_inv, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed for the mesh tiles"); return METAD_ERR_CUDA; }
    p->have_tmaps = true;
    return METAD_OK;
}

// tile order of this call (rebuilt if needed) + spread into the integer mesh; sums -> p->d_sums
// host-side decision + (rare) rebuild of the tile order: never part of a captured graph
int prepare_order(metad_mesh* p, const float* d_postype, unsigned N, cudaStream_t stream) {
    int rc = ensure_capacity(p, N); if (rc) return rc;
    rc = mark(p, 0, stream); if (rc) return rc;
    if (N == 0) return METAD_OK;
    // drifted particles / range warnings reported by an earlier spread (asynchronous copy: may lag by a call)
    const bool drift = p->h_counters[1] > N / 256u || p->h_counters[3] > 0;
    // accumulator width asked for by the last rebuild (read one or more calls late): a change re-runs the rebuild
    const unsigned mode = p->h_mode[0];
    bool mode_change = false;
    if (p->order_valid && mode != 0 && !p->g.slab) {
        if (mode == 3) { set_error("cv.mesh: more than 38 000 particles in one mesh cell -- beyond the range of the fixed-point density"); return METAD_ERR_UNSUPPORTED; }
        mode_change = (mode == 2) != p->wide;
        if (mode_change) p->wide = mode == 2;
    }
    if (!p->order_valid || p->order_N != N || p->calls_since_rebuild >= p->period || drift || mode_change) {
        const bool first = !p->order_valid || p->order_N != N;
        rc = rebuild_order(p, d_postype, N, stream, first); if (rc) return rc;
        p->h_counters[1] = p->h_counters[3] = 0;
    }
    ++p->calls_since_rebuild;
    return METAD_OK;
}

// spread into the integer mesh through the current tile order; sums -> p->d_sums
int enqueue_spread(metad_mesh* p, const float* d_postype, unsigned N, cudaStream_t stream) {
    const Geom& g = p->g;
    int rc = mark(p, 1, stream); if (rc) return rc;
    if (N == 0) {
        METAD_CUDA(cudaMemsetAsync(p->d_sums, 0, 4 * sizeof(double), stream));
        return METAD_OK;
    }
    SpreadOut out;
    memset(&out, 0, sizeof out);
    out.mesh = p->d_mesh_i;
    out.mesh64 = p->d_mesh64;
    out.tile_sums = p->d_tile_sums;
    out.sums = p->d_sums;
    out.counters = p->d_counters;
    out.h_counters = p->h_counters;
    out.keys = p->keep_cells ? p->d_keys : nullptr;
    out.cache4 = p->d_cache4;
    out.debug = p->spread_debug;
    int flags = (p->keep_cells ? kSpKeys : 0) | (p->cache ? kSpCache : 0);
    if (g.tri) flags |= kSpTri;
    if (p->wide) flags |= kSpWide;
    else if (p->tma_flush && !g.slab && !g.tri) {
        rc = ensure_tmaps(p); if (rc) return rc;
        flags |= kSpTma;
        out.tmap = p->tmap_mesh;
        out.tmap_z0 = 0;
    }
    rc = g.lgT == 4 ? dispatch_spread<4>(p, flags, d_postype, out, stream) : dispatch_spread<3>(p, flags, d_postype, out, stream);
    if (rc) return rc;
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}

int order_and_spread(metad_mesh* p, const float* d_postype, unsigned N, cudaStream_t stream) {
    int rc = prepare_order(p, d_postype, N, stream); if (rc) return rc;
    return enqueue_spread(p, d_postype, N, stream);
}

// Run `body` (which only enqueues work on `stream`) either directly or -- in graph mode -- through a CUDA graph that is
// captured the second time the same argument signature is seen and replayed afterwards.
template <class Body>
int run_captured(metad_mesh* p, const metad_mesh::GraphKey& key, cudaStream_t stream, Body body) {
    if (!p->graph_mode || p->profile) return body(stream);
    const bool same = memcmp(&key, &p->gkey, sizeof key) == 0;
    if (same && p->gexec) {
        METAD_CUDA(cudaGraphLaunch(p->gexec, stream));
        ++p->n_graph_launches;
        return METAD_OK;
    }
    if (!same) {
        if (p->gexec) { cudaGraphExecDestroy(p->gexec); p->gexec = nullptr; }
        p->gkey = key;
        p->gwarm = 1;
        return body(stream);            // first call with this signature: eager (lazy allocations, kernel attributes)
    }
    // capture on a private stream (the caller's may be the legacy default stream, which cannot be captured); the
    // instantiated graph is then launched into the caller's stream
    if (!p->capture_stream) METAD_CUDA(cudaStreamCreateWithFlags(&p->capture_stream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    METAD_CUDA(cudaStreamBeginCapture(p->capture_stream, cudaStreamCaptureModeRelaxed));
    const int rc = body(p->capture_stream);
    const cudaError_t e = cudaStreamEndCapture(p->capture_stream, &graph);
    if (rc != METAD_OK || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        if (rc == METAD_OK) return cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
        return rc;
    }
    const cudaError_t ei = cudaGraphInstantiate(&p->gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) { p->gexec = nullptr; return cuda_fail(ei, "cudaGraphInstantiate", __FILE__, __LINE__); }
    METAD_CUDA(cudaGraphLaunch(p->gexec, stream));
    ++p->n_graph_launches;
    return METAD_OK;
}

metad_mesh::GraphKey make_key(metad_mesh* p, const void* postype, unsigned N, unsigned N_global, const metad_box* box, const void* d_cv,
                              cudaStream_t stream, int kind) {
    metad_mesh::GraphKey k;
    memset(&k, 0, sizeof k);
    k.postype = postype; k.N = N; k.N_global = N_global;
    for (int i = 0; i < 3; ++i) { k.L[i] = box->L[i]; k.tilt[i] = box->tilt[i]; }
    k.d_cv = d_cv; k.stream = stream; k.kind = kind; k.keep_rho = p->keep_rho; k.keep_cells = p->keep_cells;
    k.variant = (p->tilt_literal ? (1 << 30) : 0) | (p->fuse_xy << 6) | (p->wide ? 1 : 0) | (p->cache ? 2 : 0) | (p->tma_flush ? 4 : 0) | (p->tma_gather ? 8 : 0) | (p->extras ? 16 : 0) | (p->use_table ? 32 : 0) |
                (int)(p->n_table << 8);
    return k;
}

int launch_gather(metad_mesh* p, const float* d_postype, const float* d_ghost, float* d_force, unsigned N_global, const metad_box* box,
                  const double* d_bias, cudaStream_t stream, const PeerSync* sync = nullptr) {
    PeerSync ps;
    memset(&ps, 0, sizeof ps);
    if (sync) ps = *sync;
    const Geom& g = p->g;
    // reciprocal lattice vectors of the box without 2 pi, times the mesh dimensions (:761-769, :852-854)
    ForceParams fp;
    memset(&fp, 0, sizeof fp);
    set_force_matrix(fp, g, box);
    fp.two_over_n = 2.0 / (double)N_global;
    int rc = mark(p, 8, stream); if (rc) return rc;
    GatherIn gin;
    memset(&gin, 0, sizeof gin);
    gin.postype = (const float4*)d_postype;
    gin.order = p->d_ranks;
    gin.cache4 = p->d_cache4;
    gin.mode = p->d_mode;
    gin.ntypes = p->ntypes;
    if (p->tma_gather && !g.slab) {
        rc = ensure_tmaps(p); if (rc) return rc;
        gin.tmap = p->tmap_inv;
        gin.use_tmap = 1;
    }
    if (g.tri) {
        // triclinic box: general TSC weights (mesh_gather_kernel<..., TRI>); one CTA shape per tile size
        if (g.lgT == 4) {
            const size_t sm = tile_smem_bytes<4>() + gather_stage_bytes(192, p->cache, p->ntypes);
            if (p->cache) {
                rc = set_smem(mesh_gather_kernel<4, 192, 3, true, true>, tile_smem_bytes<4>() + gather_stage_bytes(192, true, kSpreadModes)); if (rc) return rc;
                METAD_CUDA(launch_pdl(p->pdl, mesh_gather_kernel<4, 192, 3, true, true>, num_tiles(g), 192, sm, stream, gin, p->d_tstart, g, p->d_buf, d_ghost, fp,
                                      d_bias, (float4*)d_force, ps));
            } else {
                rc = set_smem(mesh_gather_kernel<4, 192, 3, false, true>, tile_smem_bytes<4>() + gather_stage_bytes(192, false, kSpreadModes)); if (rc) return rc;
                METAD_CUDA(launch_pdl(p->pdl, mesh_gather_kernel<4, 192, 3, false, true>, num_tiles(g), 192, sm, stream, gin, p->d_tstart, g, p->d_buf, d_ghost, fp,
                                      d_bias, (float4*)d_force, ps));
            }
        } else {
            const size_t sm = tile_smem_bytes<3>() + gather_stage_bytes(kGatherThreads, p->cache, p->ntypes);
            if (p->cache)
                METAD_CUDA(launch_pdl(p->pdl, mesh_gather_kernel<3, kGatherThreads, 3, true, true>, num_tiles(g), kGatherThreads, sm, stream, gin, p->d_tstart, g, p->d_buf,
                                      d_ghost, fp, d_bias, (float4*)d_force, ps));
            else
                METAD_CUDA(launch_pdl(p->pdl, mesh_gather_kernel<3, kGatherThreads, 3, false, true>, num_tiles(g), kGatherThreads, sm, stream, gin, p->d_tstart, g, p->d_buf,
                                      d_ghost, fp, d_bias, (float4*)d_force, ps));
        }
    } else if (g.lgT == 4) {
        // CTA size / residency of the gather.  Measured on B200 (C4, ms per launch): 256 threads x 3 CTAs/SM (80 registers, the loop
        // state spills) 0.238; 256 x 2 0.210; 192 x 3 (96 registers, no spill) 0.201; 128 x 5 0.216; 128 x 4 0.212; 192 x 4 0.247.
        // METAD_GATHER_VARIANT selects the others for experiments.
        // With at most 5 tiles per SM (C3, or a rank's slab on 8 GPUs) 128 x 5 keeps every CTA resident in one round:
        // C3 0.0266 -> 0.0245 ms.
        static const int forced = getenv("METAD_GATHER_VARIANT") ? atoi(getenv("METAD_GATHER_VARIANT")) : -1;
        const int variant = forced >= 0 ? forced : (num_tiles(g) <= 5u * (unsigned)device_sm_count() ? 3 : 2);
#define METAD_GATHER_LAUNCH_C(T_, B_, C_)                                                                                         \
    {                                                                                                                             \
        const size_t sm = tile_smem_bytes<4>() + gather_stage_bytes(T_, C_, p->ntypes);                                           \
        rc = set_smem(mesh_gather_kernel<4, T_, B_, C_>, tile_smem_bytes<4>() + gather_stage_bytes(T_, C_, kSpreadModes)); if (rc) return rc; \
        METAD_CUDA(launch_pdl(p->pdl, mesh_gather_kernel<4, T_, B_, C_>, num_tiles(g), T_, sm, stream, gin, p->d_tstart, g, p->d_buf, d_ghost, fp, \
                              d_bias, (float4*)d_force, ps));                                                                     \
    }
#define METAD_GATHER_LAUNCH(T_, B_) { if (p->cache) METAD_GATHER_LAUNCH_C(T_, B_, true) else METAD_GATHER_LAUNCH_C(T_, B_, false) }
        switch (variant) {
            case 1: METAD_GATHER_LAUNCH(256, 2) break;
            case 2: METAD_GATHER_LAUNCH(192, 3) break;
            case 3: METAD_GATHER_LAUNCH(128, 5) break;
            case 4: METAD_GATHER_LAUNCH(128, 6) break;
            case 5: METAD_GATHER_LAUNCH(192, 4) break;
            case 6: METAD_GATHER_LAUNCH(128, 4) break;
            case 7: METAD_GATHER_LAUNCH(224, 3) break;
            case 8: METAD_GATHER_LAUNCH(160, 4) break;
            case 9: METAD_GATHER_LAUNCH(160, 3) break;
            default: METAD_GATHER_LAUNCH(256, 3) break;
        }
#undef METAD_GATHER_LAUNCH
#undef METAD_GATHER_LAUNCH_C
    } else {
        const size_t sm = tile_smem_bytes<3>() + gather_stage_bytes(kGatherThreads, p->cache, p->ntypes);
        if (p->cache)
            METAD_CUDA(launch_pdl(p->pdl, mesh_gather_kernel<3, kGatherThreads, 3, true>, num_tiles(g), kGatherThreads, sm, stream, gin, p->d_tstart, g, p->d_buf,
                                  d_ghost, fp, d_bias, (float4*)d_force, ps));
        else
            METAD_CUDA(launch_pdl(p->pdl, mesh_gather_kernel<3, kGatherThreads, 3, false>, num_tiles(g), kGatherThreads, sm, stream, gin, p->d_tstart, g, p->d_buf,
                                  d_ghost, fp, d_bias, (float4*)d_force, ps));
    }
    METAD_LAUNCH_CHECK();
    return mark(p, 9, stream);
}

// ---- general path (mesh_general.cuh) ------------------------------------------------------------------
bool tiled_path_takes(unsigned nx, unsigned ny, unsigned nz) {
    return is_pow2(nx) && is_pow2(ny) && is_pow2(nz) && nx >= 32 && nx <= 1024 && ny >= 16 && ny <= 512 && nz >= 16 && nz <= 512;
}

int general_create(metad_mesh** out, unsigned nx, unsigned ny, unsigned nz, int ntypes, const double* mode) {
    using namespace meshgen;
    if (nx < 1 || ny < 1 || nz < 1 || nx > kMaxLen || ny > kMaxLen || nz > kMaxLen) {
        set_error("cv.mesh: supported mesh sizes are 1 <= nx, ny, nz <= 1024");
        return METAD_ERR_UNSUPPORTED;
    }
    auto* p = new metad_mesh();
    p->general = true;
    geom_set_dims_general(p->g, nx, ny, nz);
    p->n_ranks = 1; p->rank = 0; p->nzg = nz; p->kxl = 0;
    p->ntypes = ntypes;
    std::vector<float> m(ntypes);
    p->amax = 0.f;
    for (int i = 0; i < ntypes; ++i) { m[i] = (float)mode[i]; p->amax = fmaxf(p->amax, fabsf(m[i])); }
    const size_t M = (size_t)nx * ny * nz;
    p->n_partials = kMaxConvBlocks;
    int rc = METAD_OK;
    auto fail = [&](cudaError_t e, const char* what) { rc = cuda_fail(e, what, __FILE__, __LINE__); };
    cudaError_t e;
#define TRY(call) if (rc == METAD_OK && (e = (call)) != cudaSuccess) fail(e, #call)
    TRY(cudaMalloc(&p->d_mode, sizeof(float) * ntypes));
    TRY(cudaMemcpy(p->d_mode, m.data(), sizeof(float) * ntypes, cudaMemcpyHostToDevice));
    TRY(cudaMalloc(&p->d_mesh64, sizeof(long long) * M));
    TRY(cudaMemset(p->d_mesh64, 0, sizeof(long long) * M));
    TRY(cudaMalloc(&p->d_spec, sizeof(float2) * M));
    TRY(cudaMalloc(&p->d_sums, sizeof(double) * 4));
    TRY(cudaMemset(p->d_sums, 0, sizeof(double) * 4));
    TRY(cudaMalloc(&p->d_gen_sums, sizeof(double) * 2 * kMaxConvBlocks));
    TRY(cudaMalloc(&p->d_partials, sizeof(double) * p->n_partials));
    TRY(cudaMalloc(&p->d_ticket, sizeof(unsigned)));
    TRY(cudaMemset(p->d_ticket, 0, sizeof(unsigned)));
    TRY(cudaMallocHost(&p->h_counters, sizeof(unsigned) * 4));
    TRY(cudaMallocHost(&p->h_mode, sizeof(unsigned) * 4));
#undef TRY
    if (rc == METAD_OK) {
        memset(p->h_counters, 0, sizeof(unsigned) * 4);
        memset(p->h_mode, 0, sizeof(unsigned) * 4);
    }
    const unsigned dims[3] = {nx, ny, nz};
    for (int a = 0; a < 3 && rc == METAD_OK; ++a) {
        p->grad[a] = factorize(dims[a]);
        rc = upload_twiddles(&p->d_gtw[a], dims[a]);
    }
    if (rc != METAD_OK) { metad_mesh_destroy(p); return rc; }
    *out = p;
    return METAD_OK;
}

int general_fft(metad_mesh* p, int axis, float sign, cudaStream_t st) {
    using namespace meshgen;
    LineMap lm;
    lm.nx = p->g.nx; lm.ny = p->g.ny; lm.nz = p->g.nz; lm.axis = axis;
    const unsigned n = axis == 0 ? lm.nx : (axis == 1 ? lm.ny : lm.nz);
    if (n == 1) return METAD_OK;                                    // a transform of length one is the identity
    unsigned threads = (n + 31u) / 32u * 32u;
    if (threads > (unsigned)kFftThreads) threads = kFftThreads;
    gen_fft_kernel<<<line_count(lm), threads, sizeof(float2) * 3 * n, st>>>(p->d_spec, lm, p->grad[axis], p->d_gtw[axis], sign);
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}

int general_cv(metad_mesh* p, const float* d_postype, unsigned N, unsigned N_global, double* d_cv, cudaStream_t st) {
    using namespace meshgen;
    const Geom& g = p->g;
    const size_t M = p->M();
    int rc = mark(p, 0, st); if (rc) return rc;
    rc = mark(p, 1, st); if (rc) return rc;
    if (p->keep_cells && p->cells_cap < N) {
        cudaFree(p->d_cells); p->d_cells = nullptr; p->cells_cap = 0;
        METAD_CUDA(cudaMalloc(&p->d_cells, sizeof(int) * 3 * (size_t)(N + N / 16 + 1024)));
        p->cells_cap = N + N / 16 + 1024;
    }
    if (p->keep_rho && !p->d_rho_keep) METAD_CUDA(cudaMalloc(&p->d_rho_keep, sizeof(float) * M));
    // fixed-point scale: one tap below 2^22; the 64-bit accumulators leave the cell totals alone
    const float scale = fx_scale_for(p->amax, p->amax);
    unsigned nb = (N + kParticleThreads - 1) / kParticleThreads;
    if (nb > (unsigned)kMaxConvBlocks) nb = kMaxConvBlocks;
    if (N > 0) {
        int* cells = p->keep_cells ? p->d_cells : nullptr;
        if (g.tri)
            gen_spread_kernel<true><<<nb, kParticleThreads, 0, st>>>((const float4*)d_postype, N, g, p->d_mode, scale,
                                                                    (unsigned long long*)p->d_mesh64, cells, p->d_gen_sums);
        else
            gen_spread_kernel<false><<<nb, kParticleThreads, 0, st>>>((const float4*)d_postype, N, g, p->d_mode, scale,
                                                                     (unsigned long long*)p->d_mesh64, cells, p->d_gen_sums);
        METAD_LAUNCH_CHECK();
    }
    gen_sums_kernel<<<1, 32, 0, st>>>(p->d_gen_sums, N > 0 ? nb : 0u, p->d_sums);
    METAD_LAUNCH_CHECK();
    unsigned cb = (unsigned)((M + kConvThreads - 1) / kConvThreads);
    if (cb > (unsigned)kMaxConvBlocks) cb = kMaxConvBlocks;
    const double inv_cells = 1.0 / (double)M;
    gen_density_kernel<<<cb, kConvThreads, 0, st>>>(p->d_mesh64, p->d_spec, p->keep_rho ? p->d_rho_keep : nullptr, M, 1.0f / scale, p->d_sums, inv_cells);
    METAD_LAUNCH_CHECK();
    rc = mark(p, 2, st); if (rc) return rc;
    rc = general_fft(p, 0, -1.f, st); if (rc) return rc;
    rc = mark(p, 3, st); if (rc) return rc;
    rc = general_fft(p, 1, -1.f, st); if (rc) return rc;
    rc = mark(p, 4, st); if (rc) return rc;
    rc = general_fft(p, 2, -1.f, st); if (rc) return rc;
    fft::ConvParams cp;
    memset(&cp, 0, sizeof cp);
    cp.nx = g.nx; cp.ny = g.ny; cp.nz = g.nz;
    cp.inv_n = (float)(1.0 / (double)N_global);
    cp.n_global = (double)N_global;
    cp.d_mode_sq = p->d_sums;
    cp.dc_restore = (g.tri && (g.tq[0] != 0.f || g.tq[1] != 0.f)) ? 1 : 0;
    cp.inv_cells = inv_cells;
    cp.partials = p->d_partials;
    cp.ticket = p->d_ticket;
    cp.d_cv = d_cv;
    cp.extras = p->extras ? 1 : 0;
    cp.use_table = (p->use_table && p->d_table_d && p->n_table >= 2) ? 1 : 0;
    cp.n_table = p->n_table; cp.table_d = p->d_table_d;
    cp.k_min = (float)p->k_min; cp.k_max = (float)p->k_max;
    cp.delta_k = p->n_table >= 2 ? (float)((p->k_max - p->k_min) / (double)(p->n_table - 1)) : 1.0f;
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c) cp.bk[3 * i + c] = (float)(2.0 * M_PI * p->box_b[i][c]);
    if (p->extras) {
        if (!p->d_vir_partials) {
            METAD_CUDA(cudaMalloc(&p->d_vir_partials, sizeof(double) * 6 * p->n_partials));
            METAD_CUDA(cudaMalloc(&p->d_amax_key, sizeof(unsigned long long)));
            METAD_CUDA(cudaMalloc(&p->d_extras_out, sizeof(double) * 8));
        }
        METAD_CUDA(cudaMemsetAsync(p->d_amax_key, 0, sizeof(unsigned long long), st));
        p->extras_N_global = N_global;
    }
    cp.vir_partials = p->d_vir_partials; cp.amax_key = p->d_amax_key; cp.extras_out = p->d_extras_out;
    if (p->extras) gen_conv_kernel<true><<<cb, kConvThreads, 0, st>>>(p->d_spec, M, cp);
    else gen_conv_kernel<false><<<cb, kConvThreads, 0, st>>>(p->d_spec, M, cp);
    METAD_LAUNCH_CHECK();
    rc = general_fft(p, 2, +1.f, st); if (rc) return rc;
    rc = mark(p, 5, st); if (rc) return rc;
    rc = general_fft(p, 1, +1.f, st); if (rc) return rc;
    rc = mark(p, 6, st); if (rc) return rc;
    rc = general_fft(p, 0, +1.f, st); if (rc) return rc;
    return mark(p, 7, st);
}

int general_forces(metad_mesh* p, const float* d_postype, float* d_force, unsigned N, unsigned N_global, const metad_box* box,
                   const double* d_bias, cudaStream_t st) {
    using namespace meshgen;
    const Geom& g = p->g;
    ForceParams fp;
    memset(&fp, 0, sizeof fp);
    set_force_matrix(fp, g, box);
    fp.two_over_n = 2.0 / (double)N_global;
    int rc = mark(p, 8, st); if (rc) return rc;
    unsigned nb = (N + kParticleThreads - 1) / kParticleThreads;
    if (nb > 64u * (unsigned)device_sm_count()) nb = 64u * (unsigned)device_sm_count();
    if (g.tri)
        gen_gather_kernel<true><<<nb, kParticleThreads, 0, st>>>((const float4*)d_postype, N, g, p->d_mode, p->d_spec, fp, d_bias, (float4*)d_force);
    else
        gen_gather_kernel<false><<<nb, kParticleThreads, 0, st>>>((const float4*)d_postype, N, g, p->d_mode, p->d_spec, fp, d_bias, (float4*)d_force);
    METAD_LAUNCH_CHECK();
    return mark(p, 9, st);
}

int create_common(metad_mesh** out, unsigned nx, unsigned ny, unsigned nzg, unsigned n_ranks, unsigned rank, int ntypes,
                  const double* mode) {
    METAD_REQUIRE(out && mode, "metad_mesh_create: null argument");
    METAD_REQUIRE(ntypes > 0, "Number of modes unequal number of particle types.");
    if (ntypes > kSpreadModes) { set_error("cv.mesh: at most 1024 particle types"); return METAD_ERR_UNSUPPORTED; }
    // one GPU: every mesh size works -- the tiled kernels for powers of two in their range, the general path otherwise
    if (n_ranks == 1 && (!tiled_path_takes(nx, ny, nzg) || (getenv("METAD_MESH_GENERAL") && atoi(getenv("METAD_MESH_GENERAL")) == 1)))
        return general_create(out, nx, ny, nzg, ntypes, mode);
    // z slabs: like the reference under domain decomposition (OrderParameterMesh.cc:70-79), powers of two only
    if (!is_pow2(nx) || !is_pow2(ny) || !is_pow2(nzg)) {
        set_error("cv.mesh: the number of mesh points along every direction must be a power of two when the mesh is sharded");
        return METAD_ERR_UNSUPPORTED;
    }
    if (nx < 32 || nx > 1024 || ny < 16 || ny > 512 || nzg < 16 || nzg > 512) {
        set_error("cv.mesh: sharded meshes need 32 <= nx <= 1024, 16 <= ny,nz <= 512");
        return METAD_ERR_UNSUPPORTED;
    }
    METAD_REQUIRE(n_ranks >= 1 && rank < n_ranks && is_pow2(n_ranks), "cv.mesh: the number of ranks must be a power of two");
    const bool slab = n_ranks > 1;
    if (slab && (nzg / n_ranks < 8 || nx / 2 / n_ranks < (unsigned)kLines)) {
        set_error("cv.mesh: z-slab sharding needs nz/ranks >= 8 and nx/2/ranks >= 16");
        return METAD_ERR_UNSUPPORTED;
    }
    auto* p = new metad_mesh();
    Geom& g = p->g;
    memset(&g, 0, sizeof g);
    const unsigned nzl = nzg / n_ranks;
    const size_t M = (size_t)nx * ny * nzl;
    // 16^3 tiles once there are enough of them to fill the GPU twice, 8^3 otherwise
    unsigned lgT = (M / 4096 >= (size_t)2 * device_sm_count()) ? 4 : 3;
    if (nzl < 16) lgT = 3;
    if (const char* e = getenv("METAD_TILE_LOG2")) {            // experiments: force 8^3 or 16^3 tiles
        if (e[0] == '3') lgT = 3;
        if (e[0] == '4' && nzl >= 16) lgT = 4;
    }
    geom_set_dims(g, nx, ny, nzl, lgT);
    g.nzg = nzg; g.z0 = rank * nzl; g.slab = slab ? 1 : 0;
    p->n_ranks = n_ranks; p->rank = rank; p->nzg = nzg; p->kxl = nx / 2 / n_ranks;
    p->ntypes = ntypes;
    std::vector<float> m(ntypes);
    for (int i = 0; i < ntypes; ++i) m[i] = (float)mode[i];
    p->amax = 0.f;
    for (float v : m) p->amax = fmaxf(p->amax, fabsf(v));
    const unsigned nb_scan = (unsigned)(M / (4 * kScanThreads));
    const unsigned nby = (ny / 2 + 1 + kLines / 2 - 1) / (kLines / 2);
    p->n_partials = nby + (p->kxl / kLines) * ny;
    const size_t plane = (size_t)nx * ny;
    const size_t mesh_ints = M + (slab ? 2 * plane : 0);
    const float fx_init[8] = {1.0f, 1.0f, 0.f, 0.f, 1.0f, 0.f, 0.f, 0.f};
    int rc = METAD_OK;
    auto fail = [&](cudaError_t e, const char* what) { rc = cuda_fail(e, what, __FILE__, __LINE__); };
    cudaError_t e;
#define TRY(call) if (rc == METAD_OK && (e = (call)) != cudaSuccess) fail(e, #call)
    TRY(cudaMalloc(&p->d_mode, sizeof(float) * ntypes));
    TRY(cudaMemcpy(p->d_mode, m.data(), sizeof(float) * ntypes, cudaMemcpyHostToDevice));
    TRY(cudaMalloc(&p->d_count, sizeof(unsigned) * M));
    TRY(cudaMemset(p->d_count, 0, sizeof(unsigned) * M));
    TRY(cudaMalloc(&p->d_start, sizeof(unsigned) * (M + 4)));
    TRY(cudaMalloc(&p->d_block_sums, sizeof(unsigned) * (nb_scan + 1)));
    TRY(cudaMalloc(&p->d_tstart, sizeof(unsigned) * (num_tiles(g) + 1)));
    TRY(cudaMemset(p->d_tstart, 0, sizeof(unsigned) * (num_tiles(g) + 1)));
    TRY(cudaMalloc(&p->d_max_count, sizeof(unsigned)));
    TRY(cudaMalloc(&p->d_mesh_alloc, sizeof(int) * mesh_ints));
    TRY(cudaMemset(p->d_mesh_alloc, 0, sizeof(int) * mesh_ints));
    TRY(cudaMalloc(&p->d_buf, sizeof(float) * M));
    TRY(cudaMalloc(&p->d_fx, sizeof(float) * 8));      // {scale, 1/scale, -, -, 1/scale, 0, 0, 0}: [4..7] = trailer of a halo message
    TRY(cudaMemcpy(p->d_fx, fx_init, sizeof fx_init, cudaMemcpyHostToDevice));
    TRY(cudaMalloc(&p->d_sums, sizeof(double) * 4));
    TRY(cudaMemset(p->d_sums, 0, sizeof(double) * 4));
    TRY(cudaMalloc(&p->d_tile_sums, sizeof(double) * 2 * num_tiles(g)));
    TRY(cudaMalloc(&p->d_counters, sizeof(unsigned) * 8));
    TRY(cudaMemset(p->d_counters, 0, sizeof(unsigned) * 8));
    TRY(cudaMallocHost(&p->h_counters, sizeof(unsigned) * 4));
    TRY(cudaMallocHost(&p->h_mode, sizeof(unsigned) * 4));
    TRY(cudaMalloc(&p->d_partials, sizeof(double) * p->n_partials));
    TRY(cudaMalloc(&p->d_ticket, sizeof(unsigned)));
    TRY(cudaMemset(p->d_ticket, 0, sizeof(unsigned)));
    TRY(cudaMalloc(&p->d_sums_global, sizeof(double) * 4));
    TRY(cudaMemset(p->d_sums_global, 0, sizeof(double) * 4));
    TRY(cudaMalloc(&p->d_cv_partial, sizeof(double)));
    TRY(cudaMalloc(&p->d_p2p_status, sizeof(unsigned)));
    TRY(cudaMemset(p->d_p2p_status, 0, sizeof(unsigned)));
    TRY(cudaMalloc(&p->d_epoch, sizeof(unsigned)));
    TRY(cudaMemset(p->d_epoch, 0, sizeof(unsigned)));
    TRY(cudaMalloc(&p->d_sync, sizeof(unsigned) * 12));
    TRY(cudaMemset(p->d_sync, 0, sizeof(unsigned) * 12));
#undef TRY
    if (rc == METAD_OK) {
        memset(p->h_counters, 0, sizeof(unsigned) * 4);
        memset(p->h_mode, 0, sizeof(unsigned) * 4);
        p->d_mesh_i = p->d_mesh_alloc + (slab ? plane : 0);
    }
    if (rc == METAD_OK) rc = upload_twiddles(&p->d_twx, nx);
    if (rc == METAD_OK) rc = upload_twiddles(&p->d_twy, ny);
    if (rc == METAD_OK) rc = upload_twiddles(&p->d_twz, nzg);
    if (rc != METAD_OK) { metad_mesh_destroy(p); return rc; }
    *out = p;
    return METAD_OK;
}

}  // namespace

extern "C" int metad_mesh_create(metad_mesh** out, unsigned nx, unsigned ny, unsigned nz, int ntypes, const double* mode) {
    return create_common(out, nx, ny, nz, 1, 0, ntypes, mode);
}

extern "C" int metad_mesh_slab_create(metad_mesh** out, unsigned nx, unsigned ny, unsigned nz, unsigned n_ranks, unsigned rank, int ntypes,
                                      const double* mode) {
    METAD_REQUIRE(n_ranks >= 2, "metad_mesh_slab_create: use metad_mesh_create for a single rank");
    return create_common(out, nx, ny, nz, n_ranks, rank, ntypes, mode);
}

extern "C" int metad_mesh_destroy(metad_mesh* p) {
    if (!p) return METAD_OK;
    cudaFree(p->d_mode); cudaFree(p->d_keys); cudaFree(p->d_ranks); cudaFree(p->d_perm); cudaFree(p->d_cache4);
    cudaFree(p->d_count); cudaFree(p->d_start); cudaFree(p->d_block_sums); cudaFree(p->d_tstart); cudaFree(p->d_max_count);
    cudaFree(p->d_mesh_alloc); cudaFree(p->d_buf); cudaFree(p->d_fx); cudaFree(p->d_tile_sums); cudaFree(p->d_counters);
    if (p->h_counters) cudaFreeHost(p->h_counters);
    if (p->h_mode) cudaFreeHost(p->h_mode);
    cudaFree(p->d_mesh64); cudaFree(p->d_table_d); cudaFree(p->d_vir_partials); cudaFree(p->d_amax_key); cudaFree(p->d_extras_out);
    cudaFree(p->d_rho_keep); cudaFree(p->d_twx); cudaFree(p->d_twy); cudaFree(p->d_twz); cudaFree(p->d_sums);
    cudaFree(p->d_partials); cudaFree(p->d_ticket);
    cudaFree(p->d_spec); cudaFree(p->d_cells); cudaFree(p->d_gen_sums);
    for (auto& t : p->d_gtw) cudaFree(t);
    cudaFree(p->d_sums_global); cudaFree(p->d_cv_partial); cudaFree(p->d_p2p_status); cudaFree(p->d_epoch); cudaFree(p->d_sync);
    if (p->gexec) cudaGraphExecDestroy(p->gexec);
    if (p->capture_stream) cudaStreamDestroy(p->capture_stream);
    for (unsigned r = 0; r < p2p::kMaxPeers; ++r)
        if (p->peers_mapped[r]) cudaIpcCloseMemHandle(p->peers.arena[r]);
    cudaFree(p->arena);
    for (auto& e : p->ev) if (e) cudaEventDestroy(e);
    for (auto& e : p->evp) if (e) cudaEventDestroy(e);
    delete p;
    return METAD_OK;
}

extern "C" int metad_mesh_cv(metad_mesh* p, const float* d_postype, unsigned N, unsigned N_global, const metad_box* box,
                             double* d_cv, metad_stream_t stream) {
    METAD_REQUIRE(p && box && d_cv, "metad_mesh_cv: null argument");
    METAD_REQUIRE(!p->g.slab, "metad_mesh_cv: this plan is a z-slab shard, use the metad_mesh_slab_* stages");
    METAD_REQUIRE(N == 0 || d_postype, "metad_mesh_cv: null positions");
    METAD_REQUIRE(N_global > 0, "metad_mesh_cv: N_global must be positive");
    int rc = set_box(p, box); if (rc) return rc;
    p->have_cv = false;
    if (p->general) {
        rc = general_cv(p, d_postype, N, N_global, d_cv, stream); if (rc) return rc;
        p->have_cv = true;
        p->last_N = N;
        return METAD_OK;
    }
    rc = prepare_order(p, d_postype, N, stream); if (rc) return rc;
    rc = run_captured(p, make_key(p, d_postype, N, N_global, box, d_cv, stream, 0), stream, [&](cudaStream_t st) -> int {
        int r = enqueue_spread(p, d_postype, N, st); if (r) return r;
        return fft_pipeline(p, N_global, d_cv, st);
    });
    if (rc) return rc;
    p->have_cv = true;
    p->last_N = N;
    return METAD_OK;
}

extern "C" int metad_mesh_forces(metad_mesh* p, const float* d_postype, float* d_force, unsigned N, unsigned N_global,
                                 const metad_box* box, const double* d_bias, metad_stream_t stream) {
    METAD_REQUIRE(p && box && d_bias, "metad_mesh_forces: null argument");
    METAD_REQUIRE(!p->g.slab, "metad_mesh_forces: this plan is a z-slab shard, use metad_mesh_slab_forces");
    METAD_REQUIRE(N_global > 0, "metad_mesh_forces: N_global must be positive");
    if (!p->have_cv || p->last_N != N) {
        set_error("metad_mesh_forces: call metad_mesh_cv for the same particles first");
        return METAD_ERR_STATE;
    }
    if (N == 0) return METAD_OK;
    METAD_REQUIRE(d_postype && d_force, "metad_mesh_forces: null particle arrays");
    if (p->general) return general_forces(p, d_postype, d_force, N, N_global, box, d_bias, stream);
    return launch_gather(p, d_postype, nullptr, d_force, N_global, box, d_bias, stream);
}

// ---- z-slab stages ------------------------------------------------------------------------------------
extern "C" int metad_mesh_slab_spread(metad_mesh* p, const float* d_postype, unsigned N_local, const metad_box* global_box,
                                      double* d_sums, int* d_ghost_send, metad_stream_t stream) {
    METAD_REQUIRE(p && global_box && d_sums && d_ghost_send, "metad_mesh_slab_spread: null argument");
    METAD_REQUIRE(p->g.slab, "metad_mesh_slab_spread: not a slab plan");
    METAD_REQUIRE(N_local == 0 || d_postype, "metad_mesh_slab_spread: null positions");
    int rc = set_box(p, global_box); if (rc) return rc;
    p->have_cv = false;
    rc = order_and_spread(p, d_postype, N_local, stream); if (rc) return rc;
    // the two ghost planes (global planes z0-1 and z0+nz) leave for the neighbours and are cleared for the next call
    const size_t plane = (size_t)p->g.nx * p->g.ny;
    int* below = p->d_mesh_alloc;
    int* above = p->d_mesh_alloc + plane * (p->g.nz + 1);
    const size_t msg = plane + 4;        // a halo message: the plane, then 4 ints of which the first carries the bits of 1/scale
    METAD_CUDA(cudaMemcpyAsync(d_ghost_send, below, plane * sizeof(int), cudaMemcpyDeviceToDevice, stream));
    METAD_CUDA(cudaMemcpyAsync(d_ghost_send + plane, p->d_fx + 1, sizeof(float), cudaMemcpyDeviceToDevice, stream));
    METAD_CUDA(cudaMemcpyAsync(d_ghost_send + msg, above, plane * sizeof(int), cudaMemcpyDeviceToDevice, stream));
    METAD_CUDA(cudaMemcpyAsync(d_ghost_send + msg + plane, p->d_fx + 1, sizeof(float), cudaMemcpyDeviceToDevice, stream));
    METAD_CUDA(cudaMemsetAsync(below, 0, plane * sizeof(int), stream));
    METAD_CUDA(cudaMemsetAsync(above, 0, plane * sizeof(int), stream));
    METAD_CUDA(cudaMemcpyAsync(d_sums, p->d_sums, 3 * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    p->last_N = N_local;
    return METAD_OK;
}

extern "C" int metad_mesh_slab_fft_x(metad_mesh* p, const int* d_ghost_recv, const double* d_sums_global, float* d_send,
                                     metad_stream_t stream) {
    METAD_REQUIRE(p && d_ghost_recv && d_sums_global && d_send, "metad_mesh_slab_fft_x: null argument");
    METAD_REQUIRE(p->g.slab, "metad_mesh_slab_fft_x: not a slab plan");
    if (p->keep_rho && !p->d_rho_keep) METAD_CUDA(cudaMalloc(&p->d_rho_keep, sizeof(float) * p->M()));
    int rc = METAD_OK;
    METAD_DISPATCH_LEN(p->g.nx / 2, (run_x<LL>(p, false, reinterpret_cast<float2*>(d_send), d_sums_global, d_ghost_recv, stream)));
    return rc;
}

extern "C" int metad_mesh_slab_fft_yz(metad_mesh* p, float* d_pencil, const double* d_sums_global, unsigned N_global,
                                      double* d_cv_partial, metad_stream_t stream) {
    METAD_REQUIRE(p && d_pencil && d_sums_global && d_cv_partial, "metad_mesh_slab_fft_yz: null argument");
    METAD_REQUIRE(p->g.slab, "metad_mesh_slab_fft_yz: not a slab plan");
    METAD_REQUIRE(N_global > 0, "metad_mesh_slab_fft_yz: N_global must be positive");
    float2* buf = reinterpret_cast<float2*>(d_pencil);
    int rc = METAD_OK;
    METAD_DISPATCH_LEN(p->g.ny, (run_y<LL>(p, false, buf, p->kxl, p->nzg, stream))); if (rc) return rc;
    METAD_DISPATCH_LEN(p->nzg, (run_z<LL>(p, buf, p->kxl, p->rank * p->kxl, d_sums_global, N_global, d_cv_partial, stream))); if (rc) return rc;
    METAD_DISPATCH_LEN(p->g.ny, (run_y<LL>(p, true, buf, p->kxl, p->nzg, stream)));
    return rc;
}

extern "C" int metad_mesh_slab_fft_x_inv(metad_mesh* p, const float* d_recv, float* d_planes_out, metad_stream_t stream) {
    METAD_REQUIRE(p && d_recv, "metad_mesh_slab_fft_x_inv: null argument");
    METAD_REQUIRE(p->g.slab, "metad_mesh_slab_fft_x_inv: not a slab plan");
    int rc = METAD_OK;
    METAD_DISPATCH_LEN(p->g.nx / 2, (run_x<LL>(p, true, reinterpret_cast<float2*>(const_cast<float*>(d_recv)), nullptr, nullptr, stream)));
    if (rc) return rc;
    const size_t plane = (size_t)p->g.nx * p->g.ny;
    if (d_planes_out) {     // first and last local plane of Re IFFT(G): the neighbours' halo planes
        METAD_CUDA(cudaMemcpyAsync(d_planes_out, p->d_buf, plane * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        METAD_CUDA(cudaMemcpyAsync(d_planes_out + plane, p->d_buf + plane * (p->g.nz - 1), plane * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    }
    p->have_cv = true;
    return METAD_OK;
}

extern "C" int metad_mesh_slab_forces(metad_mesh* p, const float* d_ghost_inv, const float* d_postype, float* d_force, unsigned N_local,
                                      unsigned N_global, const metad_box* global_box, const double* d_bias, metad_stream_t stream) {
    METAD_REQUIRE(p && d_ghost_inv && global_box && d_bias, "metad_mesh_slab_forces: null argument");
    METAD_REQUIRE(p->g.slab, "metad_mesh_slab_forces: not a slab plan");
    METAD_REQUIRE(N_global > 0, "metad_mesh_slab_forces: N_global must be positive");
    if (!p->have_cv || p->last_N != N_local) {
        set_error("metad_mesh_slab_forces: run the spread/fft stages for the same particles first");
        return METAD_ERR_STATE;
    }
    if (N_local == 0) return METAD_OK;
    METAD_REQUIRE(d_postype && d_force, "metad_mesh_slab_forces: null particle arrays");
    return launch_gather(p, d_postype, d_ghost_inv, d_force, N_global, global_box, d_bias, stream);
}

// ---- z-slab stages over peer memory (NVLink): no library collective inside a step ----------------------------
namespace {

int ensure_arena(metad_mesh* p) {
    if (p->arena) return METAD_OK;
    METAD_REQUIRE(p->g.slab, "peer-memory mode: not a slab plan");
    METAD_REQUIRE(p->n_ranks <= (unsigned)p2p::kMaxPeers, "peer-memory mode supports up to 8 ranks");
    p->lay = p2p::arena_layout(p->M(), (size_t)p->g.nx * p->g.ny);
    METAD_CUDA(cudaMalloc(&p->arena, p->lay.total));
    METAD_CUDA(cudaMemset(p->arena, 0, p->lay.total));
    return METAD_OK;
}

int p2p_barrier(metad_mesh* p, int wait, p2p::Publish pub, p2p::Reduce red, cudaStream_t st) {
    METAD_CUDA(launch_pdl(p->pdl, p2p::barrier_kernel, 1, 32, 0, st, p->peers, p->lay.flags + 4 * p2p::kMaxPeers * sizeof(unsigned), p->d_epoch, wait, pub, red,
                                           p->d_p2p_status));
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}

int p2p_push_barrier(metad_mesh* p, const p2p::PushJob& job, p2p::Publish pub, p2p::Reduce red, cudaStream_t st) {
    METAD_CUDA(launch_pdl(p->pdl, p2p::push_barrier_kernel, 32, 256, 0, st, job, p->peers, p->lay.flags + 4 * p2p::kMaxPeers * sizeof(unsigned),
                          p->d_epoch, pub, red, p->d_p2p_status, p->d_sync + 8));
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}

PeerSync make_sync(metad_mesh* p, int wait_k, int signal_k) {
    PeerSync ps;
    memset(&ps, 0, sizeof ps);
    for (unsigned r = 0; r < p->n_ranks; ++r) ps.arena[r] = p->peers.arena[r];
    ps.n = p->n_ranks; ps.rank = p->rank;
    ps.flags_off = p->lay.flags;
    ps.d_epoch = p->d_sync; ps.ticket = p->d_sync + 4;
    ps.status = p->d_p2p_status;
    ps.wait_k = wait_k; ps.signal_k = signal_k;
    return ps;
}

int p2p_stage(metad_mesh* p, int stage, int wait, const float* d_postype, unsigned N_local, unsigned N_global, const metad_box* box,
              double* d_cv, cudaStream_t st) {
    const bool fused = wait && p->fused_sync;      // signal / wait inside the kernels instead of barrier launches
    const Geom& g = p->g;
    const size_t plane = (size_t)g.nx * g.ny;
    const unsigned P = p->n_ranks, r = p->rank, down = (r + P - 1) % P, up = (r + 1) % P;
    char* mine = p->arena;
    int rc = METAD_OK;
    switch (stage) {
        case 0: {   // local spread; halo planes of the density (+ their scale) to the neighbours, partial sums to everyone
            rc = markp(p, 0, st); if (rc) return rc;
            rc = enqueue_spread(p, d_postype, N_local, st); if (rc) return rc;
            rc = markp(p, 1, st); if (rc) return rc;
            int* below = p->d_mesh_alloc;
            int* above = p->d_mesh_alloc + plane * (g.nz + 1);
            const size_t msg = (plane + 4) * sizeof(int);
            p2p::PushJob job;
            memset(&job, 0, sizeof job);
            // what goes down arrives as the lower rank's message [1] (added to its last plane), what goes up as message [0]
            char* gd = p->peers.arena[down] + p->lay.ghost_rho + msg;
            char* gu = p->peers.arena[up] + p->lay.ghost_rho;
            job.dst[0] = (int4*)gd; job.src[0] = (const int4*)below; job.n16[0] = (unsigned)(plane * sizeof(int) / 16);
            job.dst[1] = (int4*)(gd + plane * sizeof(int)); job.src[1] = (const int4*)(p->d_fx + 4); job.n16[1] = 1;
            job.dst[2] = (int4*)gu; job.src[2] = (const int4*)above; job.n16[2] = (unsigned)(plane * sizeof(int) / 16);
            job.dst[3] = (int4*)(gu + plane * sizeof(int)); job.src[3] = (const int4*)(p->d_fx + 4); job.n16[3] = 1;
            const bool merged = wait && !fused && p->merge_push;
            if (merged) {       // push + barrier 1 (+ all-reduce of the partial sums) in one launch
                rc = p2p_push_barrier(p, job, p2p::Publish{p->d_sums, p->lay.sums, 4, 3},
                                      p2p::Reduce{(const double*)(mine + p->lay.sums), 4, 3, p->d_sums_global}, st);
                if (rc) return rc;
                rc = markp(p, 2, st); if (rc) return rc;
                return markp(p, 3, st);
            }
            if (fused) {
                METAD_CUDA(launch_pdl(p->pdl, p2p::push_kernel, 32, 256, 0, st, job, p2p::Publish{p->d_sums, p->lay.sums, 4, 3}, make_sync(p, -1, 0)));
            } else {
                PeerSync none;
                memset(&none, 0, sizeof none);
                METAD_CUDA(launch_pdl(p->pdl, p2p::push_kernel, 32, 256, 0, st, job, p2p::Publish{nullptr, 0, 0, 0}, none));
            }
            METAD_LAUNCH_CHECK();
            if (!wait) {    // emulation: the partial sums must be in place before ANY rank reduces them in stage 1
                METAD_CUDA(launch_pdl(p->pdl, p2p::push_scalars_kernel, 1, 64, 0, st, p->peers, p->lay.sums, 4, p->d_sums, 3));
                METAD_LAUNCH_CHECK();
            }
            return markp(p, 2, st);
        }
        case 1: {   // [barrier: halos and sums have arrived] x forward pass, every kx pencil stored into its owner's memory
            if (!fused && !(wait && p->merge_push)) {
                rc = p2p_barrier(p, wait, p2p::Publish{p->d_sums, p->lay.sums, 4, wait ? 3u : 0u},
                                 p2p::Reduce{(const double*)(mine + p->lay.sums), 4, 3, p->d_sums_global}, st);
                if (rc) return rc;
                rc = markp(p, 3, st); if (rc) return rc;
            }
            if (p->keep_rho && !p->d_rho_keep) METAD_CUDA(cudaMalloc(&p->d_rho_keep, sizeof(float) * p->M()));
            PeerOut po;
            memset(&po, 0, sizeof po);
            po.n = P; po.rank = r;
            for (unsigned q = 0; q < P; ++q) po.ptr[q] = (float2*)(p->peers.arena[q] + p->lay.pencil);
            if (fused) {
                const PeerSync ps = make_sync(p, 0, 1);
                METAD_DISPATCH_LEN(g.nx / 2, (run_x<LL>(p, false, nullptr, p->d_sums_global, (const int*)(mine + p->lay.ghost_rho), st, &po, &ps,
                                                        (const double*)(mine + p->lay.sums), p->d_sums_global)));
            } else {
                METAD_DISPATCH_LEN(g.nx / 2, (run_x<LL>(p, false, nullptr, p->d_sums_global, (const int*)(mine + p->lay.ghost_rho), st, &po)));
            }
            if (rc) return rc;
            return markp(p, 4, st);
        }
        case 2: {   // [barrier: the pencil is complete] y, fused z, inverse y with every plane stored into its owner's memory
            float2* pen = (float2*)(mine + p->lay.pencil);
            if (fused) {
                const PeerSync ps = make_sync(p, 1, -1);
                METAD_DISPATCH_LEN(g.ny, (run_y<LL>(p, false, pen, p->kxl, p->nzg, st, nullptr, &ps))); if (rc) return rc;
            } else {
                rc = p2p_barrier(p, wait, p2p::Publish{nullptr, 0, 0, 0}, p2p::Reduce{nullptr, 0, 0, nullptr}, st); if (rc) return rc;
                rc = markp(p, 5, st); if (rc) return rc;
                METAD_DISPATCH_LEN(g.ny, (run_y<LL>(p, false, pen, p->kxl, p->nzg, st))); if (rc) return rc;
            }
            rc = markp(p, 6, st); if (rc) return rc;
            METAD_DISPATCH_LEN(p->nzg, (run_z<LL>(p, pen, p->kxl, r * p->kxl, p->d_sums_global, N_global, p->d_cv_partial, st, fused))); if (rc) return rc;
            rc = markp(p, 7, st); if (rc) return rc;
            PeerOut po;
            memset(&po, 0, sizeof po);
            po.n = P; po.rank = r;
            for (unsigned q = 0; q < P; ++q) po.ptr[q] = (float2*)(p->peers.arena[q] + p->lay.recv);
            if (fused) {
                const PeerSync ps = make_sync(p, -1, 2);
                METAD_DISPATCH_LEN(g.ny, (run_y<LL>(p, true, pen, p->kxl, p->nzg, st, &po, &ps))); if (rc) return rc;
                return METAD_OK;
            }
            METAD_DISPATCH_LEN(g.ny, (run_y<LL>(p, true, pen, p->kxl, p->nzg, st, &po))); if (rc) return rc;
            rc = markp(p, 8, st); if (rc) return rc;
            if (!wait) {    // emulation: see stage 0
                METAD_CUDA(launch_pdl(p->pdl, p2p::push_scalars_kernel, 1, 64, 0, st, p->peers, p->lay.cv, 1, p->d_cv_partial, 1));
                METAD_LAUNCH_CHECK();
            }
            return METAD_OK;
        }
        case 3: {   // [barrier: planes and CV partials have arrived] CV, inverse x pass, halo planes of Re IFFT(G) to the neighbours
            if (fused) {
                const PeerSync ps = make_sync(p, 2, -1);
                METAD_DISPATCH_LEN(g.nx / 2, (run_x<LL>(p, true, (float2*)(mine + p->lay.recv), nullptr, nullptr, st, nullptr, &ps,
                                                        (const double*)(mine + p->lay.cv), d_cv)));
                if (rc) return rc;
            } else {
                rc = p2p_barrier(p, wait, p2p::Publish{p->d_cv_partial, p->lay.cv, 1, wait ? 1u : 0u},
                                 p2p::Reduce{(const double*)(mine + p->lay.cv), 1, 1, d_cv}, st);
                if (rc) return rc;
                rc = markp(p, 9, st); if (rc) return rc;
                METAD_DISPATCH_LEN(g.nx / 2, (run_x<LL>(p, true, (float2*)(mine + p->lay.recv), nullptr, nullptr, st))); if (rc) return rc;
            }
            rc = markp(p, 10, st); if (rc) return rc;
            p2p::PushJob job;
            memset(&job, 0, sizeof job);
            // my first plane is the lower rank's plane z0+nz (its ghost [1]); my last plane the upper rank's plane z0-1 (ghost [0])
            job.dst[0] = (int4*)(p->peers.arena[down] + p->lay.ghost_inv + plane * sizeof(float));
            job.src[0] = (const int4*)p->d_buf; job.n16[0] = (unsigned)(plane * sizeof(float) / 16);
            job.dst[1] = (int4*)(p->peers.arena[up] + p->lay.ghost_inv);
            job.src[1] = (const int4*)(p->d_buf + plane * (g.nz - 1)); job.n16[1] = (unsigned)(plane * sizeof(float) / 16);
            if (wait && !fused && p->merge_push) {      // push + barrier 4 in one launch
                rc = p2p_push_barrier(p, job, p2p::Publish{nullptr, 0, 0, 0}, p2p::Reduce{nullptr, 0, 0, nullptr}, st);
                if (rc) return rc;
                rc = markp(p, 11, st); if (rc) return rc;
                return markp(p, 12, st);
            }
            if (fused) {
                METAD_CUDA(launch_pdl(p->pdl, p2p::push_kernel, 32, 256, 0, st, job, p2p::Publish{nullptr, 0, 0, 0}, make_sync(p, -1, 3)));
            } else {
                PeerSync none;
                memset(&none, 0, sizeof none);
                METAD_CUDA(launch_pdl(p->pdl, p2p::push_kernel, 32, 256, 0, st, job, p2p::Publish{nullptr, 0, 0, 0}, none));
            }
            METAD_LAUNCH_CHECK();
            return markp(p, 11, st);
        }
        case 4:     // [barrier: the halo planes of Re IFFT(G) have arrived]; fused mode: the gather waits for phase 3 itself
            if (fused || (wait && p->merge_push)) return METAD_OK;
            rc = p2p_barrier(p, wait, p2p::Publish{nullptr, 0, 0, 0}, p2p::Reduce{nullptr, 0, 0, nullptr}, st); if (rc) return rc;
            return markp(p, 12, st);
        default:
            set_error("metad_mesh_slab_p2p_cv: stage must be -1 (whole step) or 0..4");
            return METAD_ERR_INVALID;
    }
}

}  // namespace

extern "C" int metad_mesh_slab_p2p_arena(metad_mesh* p, void* handle_out, unsigned long long* bytes_out) {
    METAD_REQUIRE(p && handle_out, "metad_mesh_slab_p2p_arena: null argument");
    int rc = ensure_arena(p); if (rc) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    METAD_CUDA(cudaIpcGetMemHandle(&h, p->arena));
    memcpy(handle_out, &h, sizeof h);
    if (bytes_out) *bytes_out = p->lay.total;
    return METAD_OK;
}

extern "C" int metad_mesh_slab_p2p_connect(metad_mesh* p, const void* handles) {
    METAD_REQUIRE(p && handles, "metad_mesh_slab_p2p_connect: null argument");
    int rc = ensure_arena(p); if (rc) return rc;
    p->peers.n = p->n_ranks; p->peers.rank = p->rank;
    for (unsigned r = 0; r < p->n_ranks; ++r) {
        if (r == p->rank) { p->peers.arena[r] = p->arena; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * (size_t)r, sizeof h);
        void* ptr = nullptr;
        METAD_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        p->peers.arena[r] = (char*)ptr;
        p->peers_mapped[r] = true;
    }
    p->p2p_ready = true;
    return METAD_OK;
}

extern "C" int metad_mesh_slab_p2p_connect_local(metad_mesh* p, metad_mesh* const* plans) {
    METAD_REQUIRE(p && plans, "metad_mesh_slab_p2p_connect_local: null argument");
    int rc = ensure_arena(p); if (rc) return rc;
    p->peers.n = p->n_ranks; p->peers.rank = p->rank;
    for (unsigned r = 0; r < p->n_ranks; ++r) {
        METAD_REQUIRE(plans[r] && plans[r]->n_ranks == p->n_ranks && plans[r]->rank == r, "metad_mesh_slab_p2p_connect_local: plans must be in rank order");
        rc = ensure_arena(plans[r]); if (rc) return rc;
        p->peers.arena[r] = plans[r]->arena;
    }
    p->p2p_ready = true;
    return METAD_OK;
}

extern "C" int metad_mesh_slab_p2p_cv(metad_mesh* p, const float* d_postype, unsigned N_local, unsigned N_global, const metad_box* box,
                                      double* d_cv, int stage, metad_stream_t stream) {
    METAD_REQUIRE(p && box && d_cv, "metad_mesh_slab_p2p_cv: null argument");
    METAD_REQUIRE(p->g.slab && p->p2p_ready, "metad_mesh_slab_p2p_cv: connect the peers first (metad_mesh_slab_p2p_connect)");
    METAD_REQUIRE(N_local == 0 || d_postype, "metad_mesh_slab_p2p_cv: null positions");
    METAD_REQUIRE(N_global > 0, "metad_mesh_slab_p2p_cv: N_global must be positive");
    if (stage <= 0) {       // host side of a step: geometry, tile-order decision (never captured)
        int rc = set_box(p, box); if (rc) return rc;
        p->have_cv = false;
        rc = prepare_order(p, d_postype, N_local, stream); if (rc) return rc;
        p->last_N = N_local;
    }
    if (stage >= 0) {
        const int rc = p2p_stage(p, stage, 0, d_postype, N_local, N_global, box, d_cv, stream);
        if (rc == METAD_OK && stage == 4) { p->have_cv = true; p->last_cv_fused = false; }
        return rc;
    }
    const int rc = run_captured(p, make_key(p, d_postype, N_local, N_global, box, d_cv, stream, (p->fused_sync ? 2 : 1) + (p->merge_push ? 4 : 0)), stream, [&](cudaStream_t st) -> int {
        for (int s = 0; s <= 4; ++s) {
            const int r = p2p_stage(p, s, 1, d_postype, N_local, N_global, box, d_cv, st);
            if (r) return r;
        }
        return METAD_OK;
    });
    if (rc) return rc;
    p->have_cv = true;
    p->last_cv_fused = p->fused_sync;
    return METAD_OK;
}

extern "C" int metad_mesh_slab_p2p_forces(metad_mesh* p, const float* d_postype, float* d_force, unsigned N_local, unsigned N_global,
                                          const metad_box* box, const double* d_bias, metad_stream_t stream) {
    METAD_REQUIRE(p && box && d_bias, "metad_mesh_slab_p2p_forces: null argument");
    METAD_REQUIRE(p->g.slab && p->p2p_ready, "metad_mesh_slab_p2p_forces: connect the peers first");
    if (!p->have_cv || p->last_N != N_local) {
        set_error("metad_mesh_slab_p2p_forces: run metad_mesh_slab_p2p_cv for the same particles first");
        return METAD_ERR_STATE;
    }
    if (N_local == 0) return METAD_OK;
    METAD_REQUIRE(d_postype && d_force, "metad_mesh_slab_p2p_forces: null particle arrays");
    if (p->last_cv_fused) {
        const PeerSync ps = make_sync(p, 3, -1);
        return launch_gather(p, d_postype, (const float*)(p->arena + p->lay.ghost_inv), d_force, N_global, box, d_bias, stream, &ps);
    }
    return launch_gather(p, d_postype, (const float*)(p->arena + p->lay.ghost_inv), d_force, N_global, box, d_bias, stream);
}

extern "C" int metad_mesh_set_table(metad_mesh* p, const double* dK, unsigned n, double k_min, double k_max, int use_table) {
    METAD_REQUIRE(p, "metad_mesh_set_table: null plan");
    if (n) {
        METAD_REQUIRE(dK, "metad_mesh_set_table: null table");
        METAD_REQUIRE(n >= 2 && k_min >= 0.0 && k_max > k_min, "cv.mesh kmin, kmax is invalid");
        std::vector<float> t(n);
        for (unsigned i = 0; i < n; ++i) t[i] = (float)dK[i];
        METAD_CUDA(cudaDeviceSynchronize());
        cudaFree(p->d_table_d); p->d_table_d = nullptr;
        METAD_CUDA(cudaMalloc(&p->d_table_d, sizeof(float) * (n + 1)));
        METAD_CUDA(cudaMemset(p->d_table_d, 0, sizeof(float) * (n + 1)));
        METAD_CUDA(cudaMemcpy(p->d_table_d, t.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
        p->n_table = n; p->k_min = k_min; p->k_max = k_max;
    }
    p->use_table = use_table != 0;
    return METAD_OK;
}

extern "C" int metad_mesh_get(metad_mesh* p, int which, void* h_out) {
    METAD_REQUIRE(p && h_out, "metad_mesh_get: null argument");
    METAD_CUDA(cudaDeviceSynchronize());
    const size_t M = p->M();
    switch (which) {
        case 0: {
            if (p->last_N == 0) return METAD_OK;
            if (!p->keep_cells) { set_error("metad_mesh_get: enable metad_mesh_set(p, 3, 1) before the spread to keep the cell indices"); return METAD_ERR_STATE; }
            if (p->general) {
                METAD_CUDA(cudaMemcpy(h_out, p->d_cells, sizeof(int) * 3 * (size_t)p->last_N, cudaMemcpyDeviceToHost));
                return METAD_OK;
            }
            std::vector<unsigned> keys(p->last_N);
            METAD_CUDA(cudaMemcpy(keys.data(), p->d_keys, sizeof(unsigned) * p->last_N, cudaMemcpyDeviceToHost));
            int* out = (int*)h_out;
            for (unsigned i = 0; i < p->last_N; ++i) {
                unsigned ix, iy, iz;
                cell_of_key(keys[i], p->g, ix, iy, iz);
                out[3 * (size_t)i] = (int)ix; out[3 * (size_t)i + 1] = (int)iy; out[3 * (size_t)i + 2] = (int)(iz + p->g.z0);
            }
            return METAD_OK;
        }
        case 1:
            if (!p->d_rho_keep) { set_error("metad_mesh_get: enable metad_mesh_set(p, 1, 1) before the spread to keep rho"); return METAD_ERR_STATE; }
            METAD_CUDA(cudaMemcpy(h_out, p->d_rho_keep, sizeof(float) * M, cudaMemcpyDeviceToHost));
            return METAD_OK;
        case 2:
            if (!p->have_cv) { set_error("metad_mesh_get: no inverse mesh yet"); return METAD_ERR_STATE; }
            if (p->general) {          // real part of the complex mesh
                std::vector<float2> h(M);
                METAD_CUDA(cudaMemcpy(h.data(), p->d_spec, sizeof(float2) * M, cudaMemcpyDeviceToHost));
                for (size_t c = 0; c < M; ++c) ((float*)h_out)[c] = h[c].x;
                return METAD_OK;
            }
            METAD_CUDA(cudaMemcpy(h_out, p->d_buf, sizeof(float) * M, cudaMemcpyDeviceToHost));
            return METAD_OK;
        case 3:
            METAD_CUDA(cudaMemcpy(h_out, p->d_sums, sizeof(double), cudaMemcpyDeviceToHost));
            return METAD_OK;
        case 4: {
            // per-stage milliseconds of the last metad_mesh_cv + metad_mesh_forces pair (float[kNumStages])
            if (!p->profile || !p->ev[0] || !p->ev[7] || !p->ev[9]) { set_error("metad_mesh_get: profiling is off (metad_mesh_set(p, 2, 1))"); return METAD_ERR_STATE; }
            float* out = (float*)h_out;
            for (int i = 0; i < 7; ++i) METAD_CUDA(cudaEventElapsedTime(out + i, p->ev[i], p->ev[i + 1]));
            METAD_CUDA(cudaEventElapsedTime(out + 7, p->ev[8], p->ev[9]));
            return METAD_OK;
        }
        case 8: {
            // peer-memory step (barrier launches), profiling on: milliseconds of its 12 segments (float[12]): spread, halo push,
            // barrier 1, x forward, barrier 2, y forward, z fused, y inverse, barrier 3, x inverse, halo push, barrier 4
            float* out = (float*)h_out;
            for (int i = 0; i < 12; ++i) {
                if (!p->profile || !p->evp[i] || !p->evp[i + 1]) { set_error("metad_mesh_get: no profiled peer-memory step yet"); return METAD_ERR_STATE; }
                METAD_CUDA(cudaEventElapsedTime(out + i, p->evp[i], p->evp[i + 1]));
            }
            return METAD_OK;
        }
        case 5: {
            // statistics, double[6]: rebuilds of the tile order so far; of the LAST spread: particles that took the direct
            // path, particles outside the slab, cells past half of the fixed-point range; fixed-point scale; calls since rebuild
            double* out = (double*)h_out;
            if (p->general) {          // no tile order, no drift, no range to watch
                out[0] = out[1] = out[2] = out[3] = out[5] = 0.0;
                out[4] = (double)fx_scale_for(p->amax, p->amax);
                return METAD_OK;
            }
            unsigned c[8];
            float fx[2];
            METAD_CUDA(cudaMemcpy(c, p->d_counters, sizeof c, cudaMemcpyDeviceToHost));
            METAD_CUDA(cudaMemcpy(fx, p->d_fx, sizeof fx, cudaMemcpyDeviceToHost));
            out[0] = (double)p->n_rebuilds; out[1] = c[4]; out[2] = c[5]; out[3] = c[6]; out[4] = fx[0]; out[5] = p->calls_since_rebuild;
            return METAD_OK;
        }
        case 10: {  // epilogues of the last metad_mesh_cv with knob 13 on, double[12]: k-space virial sums xx, xy, xz, yy, yz, zz (times the
                    // bias factor = external virial), q_max x, y, z, sq_max (computeQmax: arg-max over ALL k including k = 0, times N),
                    // flat index of the arg-max, its |f|^2
            if (!p->extras || !p->d_extras_out || !p->have_cv) { set_error("metad_mesh_get: no epilogue results (metad_mesh_set(p, 13, 1) before metad_mesh_cv)"); return METAD_ERR_STATE; }
            if (p->g.slab) { set_error("metad_mesh_get: q_max / virial epilogues are available on unsharded plans only"); return METAD_ERR_UNSUPPORTED; }
            double* out = (double*)h_out;
            unsigned long long key = 0;
            double sums[4];
            METAD_CUDA(cudaMemcpy(out, p->d_extras_out, sizeof(double) * 6, cudaMemcpyDeviceToHost));
            METAD_CUDA(cudaMemcpy(&key, p->d_amax_key, sizeof key, cudaMemcpyDeviceToHost));
            METAD_CUDA(cudaMemcpy(sums, p->d_sums, sizeof(double) * 3, cudaMemcpyDeviceToHost));
            unsigned vb = (unsigned)(key >> 32), flat = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFu);
            float amp;
            memcpy(&amp, &vb, 4);
            // the mean density is removed before the transforms: the k = 0 mode, f_0 = sum a / N, competes analytically; it wins
            // ties (flat index 0 comes first in the reference's scan)
            const double ng = (double)p->extras_N_global, f0 = sums[1] / ng;
            double a = (double)amp;
            // (literal triclinic offsets: the kernel restored f_0 itself, ConvParams::dc_restore, and it is no longer sum a / N)
            const bool dc_in_kernel = p->g.tri && (p->g.tq[0] != 0.f || p->g.tq[1] != 0.f);
            if (!dc_in_kernel && (key == 0 || f0 * f0 >= a)) { a = f0 * f0; flat = 0; }
            if (!(a > 0.0)) flat = 0;
            const unsigned nx = p->g.nx, ny = p->g.ny, nz = p->nzg;
            const unsigned kx = flat % nx, ky = (flat / nx) % ny, kz = flat / (nx * ny);
            const double m3[3] = {(double)miller(kx, nx), (double)miller(ky, ny), (double)miller(kz, nz)};
            for (int c = 0; c < 3; ++c)
                out[6 + c] = a > 0.0 ? 2.0 * M_PI * (m3[0] * p->box_b[0][c] + m3[1] * p->box_b[1][c] + m3[2] * p->box_b[2][c]) : 0.0;
            out[9] = a * ng;
            out[10] = (double)flat;
            out[11] = a;
            return METAD_OK;
        }
        case 9: {   // accumulator width, unsigned[2]: {in use: 0 = 32-bit, 1 = 64-bit (wide); asked for by the last rebuild: 1 / 2 / 3, see metad_mesh::wide}
            unsigned* out = (unsigned*)h_out;
            out[0] = p->wide ? 1u : 0u; out[1] = p->h_mode[0];
            return METAD_OK;
        }
        case 7: *(unsigned long long*)h_out = p->n_graph_launches; return METAD_OK;
        case 6: {   // peer-memory mode: unsigned[2] = {a barrier timed out (a peer never arrived), sum over ranks of particles outside their slab}
            unsigned* out = (unsigned*)h_out;
            double sg[4];
            METAD_CUDA(cudaMemcpy(out, p->d_p2p_status, sizeof(unsigned), cudaMemcpyDeviceToHost));
            METAD_CUDA(cudaMemcpy(sg, p->d_sums_global, sizeof sg, cudaMemcpyDeviceToHost));
            out[1] = (unsigned)sg[2];
            return METAD_OK;
        }
        default:
            set_error("metad_mesh_get: unknown selector");
            return METAD_ERR_INVALID;
    }
}

extern "C" int metad_mesh_set(metad_mesh* p, int key, long value) {
    METAD_REQUIRE(p, "metad_mesh_set: null plan");
    switch (key) {
        case 0:                                        // rebuild period of the tile order (calls); 0 = rebuild now
            if (value <= 0) p->order_valid = false;
            else p->period = (unsigned)value;
            return METAD_OK;
        case 1: p->keep_rho = value != 0; return METAD_OK;
        case 2: p->profile = value != 0; return METAD_OK;
        case 3: p->keep_cells = value != 0; return METAD_OK;
        case 4: p->graph_mode = value != 0; return METAD_OK;
        case 5: p->fused_sync = value != 0; return METAD_OK;
        case 6: p->order_kind = value != 0 ? 1 : 0; p->order_valid = false; return METAD_OK;
        case 7: p->pdl = value != 0; return METAD_OK;
        case 8: p->merge_push = value != 0; return METAD_OK;
        case 9:                                         // particle cache spread -> gather (round-1 data flow; default off)
            if (p->cache != (value != 0)) { p->cache = value != 0; p->cap = 0; }
            return METAD_OK;
        case 10: p->tma_flush = value != 0; return METAD_OK;
        case 11: p->tma_gather = value != 0; return METAD_OK;
        case 12: p->spread_debug = (int)value; return METAD_OK;
        case 13: p->extras = value != 0; return METAD_OK;
        case 15: p->fuse_xy = (int)value; return METAD_OK;
        case 16: p->tilt_literal = value != 0; return METAD_OK;
        case 14: p->use_table = value != 0; return METAD_OK;
        default: set_error("metad_mesh_set: unknown key"); return METAD_ERR_INVALID;
    }
}
