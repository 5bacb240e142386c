// mesh_fft_kernels.cuh -- the five FFT sweeps of the mesh CV (see mesh_fft.cuh for the engine).
//
// Buffer layout: one array of M = nx*ny*nz floats, reinterpreted as M/2 float2 "packed half spectrum"
// [z][y][kx], kx < nx/2.  Slot kx = 0 holds X[0] + i X[nx/2] of the row (both real after the x pass); the y and
// z passes treat it as an ordinary complex column, and the convolve step untangles it (plane0 kernel).
//
// Convolution restated from OrderParameterMesh.cc:689-708 (updateMeshes) and :887-905 (computeCV), folded onto
// the half spectrum: with f = F/N, d = 0.5*mode_sq/N^2 and chi_k = [all Miller indices of k >= 0]
// (= interp_k^2 of the reference, OrderParameterMesh.cc:448, SURVEY 8a note a4):
//     reference   G_k = f_k (|f_k|^2 - d chi_k),  CV = 1/2 sum_{k != 0} (|f_k|^4 - 2 d chi_k |f_k|^2),
//                 forces use Re(IFFT(G)).
//     here        G^H_k = f_k (|f_k|^2 - d chis_k),  chis_k = (chi_k + chi_{-k})/2   (Hermitian part of G:
//                 IFFT(G^H) = Re IFFT(G) because f_{-k} = conj f_k), and
//                 CV = 1/2 sum_{stored k != 0} w_k (|f_k|^4 - 2 d chis_k |f_k|^2), w_k = 2 for 0 < kx < nx/2 else 1.
#pragma once
#include "common.cuh"
#include "mesh_fft.cuh"
#ifdef __CUDACC__
#include <cuda_pipeline.h>
#endif

namespace metad {
namespace fft {

constexpr int kMaxPeers = 8;

struct ConvParams {
    unsigned nx, ny, nz;          // GLOBAL mesh dimensions
    unsigned row_len;             // complex elements per (z,y) row of the buffer: nx/2, or the kx pencil width when sharded
    unsigned kx_off;              // global kx of local column 0 (0 unless sharded)
    float inv_n;                  // 1 / N_global
    double n_global;
    const double* d_mode_sq;      // device: sum_j a_j^2 of this step; [1] = sum_j a_j
    // Triclinic boxes with the reference's literal offsets (Geom::tq): the derivative weights of a particle no longer sum to
    // zero, so the k = 0 mode of G matters for the forces.  The forward x sweep removed mean = fl(sum a / M) from every cell;
    // dc_restore puts M * mean / N back into f_0 before the convolution (the reference convolves k = 0 like every other mode).
    int dc_restore;
    double inv_cells;             // 1 / M (global number of cells), the value the forward x sweep uses
    double* partials;             // per-block energy partial sums
    unsigned* ticket;
    unsigned n_blocks_plane0;     // plane0 blocks write partials[0 .. n_blocks_plane0)
    double* d_cv;                 // device output: CV
    // fused peer mode: the last block also stores the partial into slot [rank] of every rank's CV table
    char* cv_arena[kMaxPeers];
    size_t cv_off;
    unsigned cv_n, cv_rank;
    // optional epilogues on the spectrum f = F/N while it is in shared memory (metad_mesh_set key 13; off by default):
    //   * arg-max of |f_k|^2 over all k (computeQmax, OrderParameterMesh.cc:1108-1179: the q*_max / sq_max log quantities)
    //   * the k-space virial sums (computeVirial, :970-1050) with the tabulated derivative of the convolution kernel
    int extras;
    int use_table;
    unsigned n_table;
    const float* table_d;         // device: derivative table dK(k), n_table entries on [k_min, k_max]
    float k_min, k_max, delta_k;
    float bk[9];                  // 2 pi b_i (rows): k = mx b_1 + my b_2 + mz b_3 (diagonal 2 pi / L for an orthorhombic box)
    double* vir_partials;         // [blocks][6]
    unsigned long long* amax_key; // packed {|f|^2 bits, ~flat index}: atomicMax over the blocks
    double* extras_out;           // [6] virial sums (without the bias factor)
};

// accumulators of the optional epilogues, one per thread
struct ExtraAcc { double v[6]; unsigned long long key; };
MHD void extra_init(ExtraAcc& a) { for (int i = 0; i < 6; ++i) a.v[i] = 0.0; a.key = 0ull; }
MHD int miller(unsigned k, unsigned n) { return (int)k - (k >= n / 2 + n % 2 ? (int)n : 0); }
// one stored mode (kx, ky, kz) with |f|^2 = val; weight = 2 when its mirror image -k is not stored
MHD void extra_add(ExtraAcc& a, const ConvParams& cp, float val, unsigned kx, unsigned ky, unsigned kz, float weight) {
    // arg-max: ties between k and -k (equal by symmetry) go to the smaller flat index x + nx (y + ny z), like the
    // reference's ascending scan with a strict comparison
    const unsigned fk = kx + cp.nx * (ky + cp.ny * kz);
    const unsigned fm = (cp.nx - kx) % cp.nx + cp.nx * ((cp.ny - ky) % cp.ny + cp.ny * ((cp.nz - kz) % cp.nz));
    const unsigned flat = weight > 1.5f ? (fk < fm ? fk : fm) : fk;
    unsigned vb;
#ifdef __CUDA_ARCH__
    vb = __float_as_uint(val);
#else
    memcpy(&vb, &val, 4);
#endif
    const unsigned long long key = ((unsigned long long)vb << 32) | (unsigned long long)(0xFFFFFFFFu - flat);
    if (key > a.key) a.key = key;
    if (cp.use_table && fk != 0) {
        const float mx = (float)miller(kx, cp.nx), my = (float)miller(ky, cp.ny), mz = (float)miller(kz, cp.nz);
        const float k0 = mx * cp.bk[0] + my * cp.bk[3] + mz * cp.bk[6], k1 = mx * cp.bk[1] + my * cp.bk[4] + mz * cp.bk[7],
                    k2 = mx * cp.bk[2] + my * cp.bk[5] + mz * cp.bk[8];
        const float knorm = sqrtf(k0 * k0 + k1 * k1 + k2 * k2);
        if (knorm >= cp.k_min && knorm < cp.k_max) {
            const float vf = (knorm - cp.k_min) / cp.delta_k;
            const unsigned vi = (unsigned)vf;
            const float d0 = cp.table_d[vi], d1 = cp.table_d[vi + 1];
            const float val_D = d0 + (vf - (float)vi) * (d1 - d0);
            // rhog = |f|^4 / N^2 (the reference divides the already normalised f by N twice more), kfac = dK / (2 |k|)
            const double rk = (double)weight * (double)val * (double)val * (double)cp.inv_n * (double)cp.inv_n * (double)val_D / (2.0 * (double)knorm);
            a.v[0] += rk * k0 * k0; a.v[1] += rk * k0 * k1; a.v[2] += rk * k0 * k2;
            a.v[3] += rk * k1 * k1; a.v[4] += rk * k1 * k2; a.v[5] += rk * k2 * k2;
        }
    }
}

// chi for one dimension: [k < ceil(n/2)] (non-negative Miller index), OrderParameterMesh.cc:417-422
MHD bool nonneg(unsigned k, unsigned n) { return k < n / 2 + n % 2; }

// pointwise convolution of one stored mode with 0 < kx < nx/2 (weight 2 in the CV sum):
// chi_k = [ky>=0][kz>=0], chi_{-k} = 0 -> chis = chi_k/2.  Returns G^H, adds the CV summand to e.
MHD float2 conv_general(float2 F, float inv_n, float d, bool chi_k, double& e) {
    const float fr = F.x * inv_n, fi = F.y * inv_n;
    const float val = fr * fr + fi * fi;
    const float chis = chi_k ? 0.5f : 0.0f;
    const float g = val - d * chis;
    const double v = (double)val;
    e += 2.0 * v * (v - 2.0 * (double)d * (double)chis);
    return make_float2(fr * g, fi * g);
}
// pointwise convolution of the packed kx = 0 slot: z1 = Z(ky,kz), z2 = Z(-ky,-kz) (not yet conjugated).
// A = (z1 + conj z2)/2 is the kx = 0 mode, B = (z1 - conj z2)/(2i) the kx = nx/2 mode; returns A' + i B'.
MHD float2 conv_plane0(float2 z1, float2 z2raw, float inv_n, float d, unsigned ky, unsigned kz, unsigned ny,
                       unsigned nz, bool counted, double& e, float dc = 0.f /* added to f_0 (ky = kz = 0 only), see ConvParams::dc_restore */) {
    const float2 z2 = cconj(z2raw);
    const float2 A = make_float2(0.5f * (z1.x + z2.x), 0.5f * (z1.y + z2.y));
    const float2 Dm = make_float2(0.5f * (z1.x - z2.x), 0.5f * (z1.y - z2.y));
    const float2 B = make_float2(Dm.y, -Dm.x);
    // kx = 0: chi_k = [ky>=0][kz>=0], chi_{-k} = [-ky>=0][-kz>=0]
    const float chis = 0.5f * ((nonneg(ky, ny) && nonneg(kz, nz) ? 1.0f : 0.0f) +
                               (nonneg((ny - ky) % ny, ny) && nonneg((nz - kz) % nz, nz) ? 1.0f : 0.0f));
    const float ar = A.x * inv_n + ((ky == 0 && kz == 0) ? dc : 0.f), ai = A.y * inv_n;
    const float va = ar * ar + ai * ai;
    const float ga = va - d * chis;
    const float2 GA = make_float2(ar * ga, ai * ga);
    // kx = nx/2: negative Miller index in x for both k and -k -> chis = 0
    const float br = B.x * inv_n, bi = B.y * inv_n;
    const float vb = br * br + bi * bi;
    const float2 GB = make_float2(br * vb, bi * vb);
    if (counted) {
        const double dva = (double)va, dvb = (double)vb;
        if (!(ky == 0 && kz == 0)) e += dva * (dva - 2.0 * (double)d * (double)chis);   // flat index 0 excluded (:892)
        e += dvb * dvb;
    }
    return make_float2(GA.x - GB.y, GA.y + GB.x);      // A' + i B'
}

// cooperative copy of the forward twiddle table exp(-2 pi i k/TWL) into shared memory
template <int TWL> __device__ __forceinline__ void load_twiddles(float2* s_tw, const float2* __restrict__ g_tw) {
    for (int i = threadIdx.x; i < TWL; i += blockDim.x) s_tw[i] = __ldg(g_tw + i);
}

// ---------------------------------------------------------------------------------------------------
// x pass, forward (R2C): 16 rows per CTA, row = y + ny*z.  The input is the fixed-point density accumulated by the
// spread (mesh_kernels.cuh): it is converted to float and the mean density is removed (DC removal, see mesh.cu) inside
// this sweep.
// ---------------------------------------------------------------------------------------------------
// Peer-memory output of a sweep (multi-GPU, NVLink): the all-to-all transposes of the slab decomposition are not a
// separate collective -- the x forward pass stores every kx pencil straight into the owning rank's pencil buffer, and the
// inverse y pass stores every plane straight into the owning rank's receive buffer (P2P stores over NVLink, overlapped
// with the transform of the other tiles).  n == 0: single buffer (unsharded, or staged for a library all-to-all).
struct PeerOut {
    float2* ptr[kMaxPeers];   // ptr[r] = destination buffer in rank r's memory
    unsigned n;               // number of ranks (0 = not used)
    unsigned rank;            // this rank
};

// Fused inter-rank synchronisation (peer-memory mode): instead of separate barrier launches, the LAST CTA of a producer
// kernel publishes "rank r has finished phase k" in every peer's flag table (release at system scope, after all CTAs of
// the kernel fenced their peer stores), and every CTA of the consumer kernel starts by waiting until all ranks have
// published phase k (acquire).  Epochs are device-side counters, so the kernels replay from a CUDA graph.  Phases:
// 0 halos + partial sums pushed, 1 pencils stored (x forward), 2 planes + CV partials stored (y inverse), 3 halo planes
// of Re IFFT(G) pushed.  n == 0: no synchronisation (single GPU, staged path, single-process emulation).
struct PeerSync {
    char* arena[kMaxPeers];
    unsigned n, rank;
    size_t flags_off;          // unsigned flags[4][kMaxPeers] in every arena
    unsigned* d_epoch;         // [4] phases completed by THIS rank
    unsigned* ticket;          // [4] CTA tickets of the producer kernels
    unsigned* status;          // [0] set if a wait timed out
    int wait_k, signal_k;      // phase this kernel waits for / publishes (-1: none)
};
#ifdef __CUDACC__
__device__ __forceinline__ void peer_wait(const PeerSync& ps) {
    if (ps.n == 0 || ps.wait_k < 0) return;
    if (threadIdx.x < ps.n) {
        const unsigned epoch = *(volatile unsigned*)(ps.d_epoch + ps.wait_k);
        const unsigned* flag = reinterpret_cast<const unsigned*>(ps.arena[ps.rank] + ps.flags_off) + ps.wait_k * kMaxPeers + threadIdx.x;
        const long long t0 = clock64();
        unsigned v;
        do {        // relaxed polling (every CTA of the kernel does this), one acquire fence once the flag is there
            asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if ((int)(v - epoch) >= 0) break;
            if (clock64() - t0 > 8000000000LL) { atomicExch(ps.status, 1u); break; }
            __nanosleep(32);
        } while (true);
        asm volatile("fence.acq_rel.sys;" ::: "memory");
    }
    __syncthreads();
}
// call at the very end of a producer kernel, by all threads of every CTA
__device__ __forceinline__ void peer_signal(const PeerSync& ps) {
    if (ps.n == 0 || ps.signal_k < 0) return;
    __shared__ bool last_cta;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                                 // this CTA's peer stores, made visible before the ticket
        last_cta = atomicAdd(ps.ticket + ps.signal_k, 1u) == gridDim.x * gridDim.y - 1;
    }
    __syncthreads();
    if (!last_cta) return;
    const unsigned epoch = *(volatile unsigned*)(ps.d_epoch + ps.signal_k) + 1u;
    __threadfence_system();
    if (threadIdx.x < ps.n) {
        unsigned* flag = reinterpret_cast<unsigned*>(ps.arena[threadIdx.x] + ps.flags_off) + ps.signal_k * kMaxPeers + ps.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) { ps.d_epoch[ps.signal_k] = epoch; ps.ticket[ps.signal_k] = 0; }
}
#endif

struct DensityIn {
    const int2* mesh;       // integer density of the local planes, row-major [z][y][x] (the caller clears it afterwards)
    const float* d_fx;      // device: {scale, 1/scale} of the fixed-point density
    const double* d_sums;   // device: [1] = (global) sum of the mode coefficients -> mean density
    double inv_cells;       // 1 / (global number of mesh cells)
    const int2* ghost;      // z slab: the two halo messages received from the neighbours, each ny*nx ints of fixed-point density
                            // followed by 4 ints of which the first holds the bits of the sender's 1/scale; message [0] is
                            // added to the first local plane, [1] to the last; nullptr if unsharded
    const double* sums_table;   // fused peer mode: per-rank {sum a^2, sum a, outside, -} rows of this rank's arena (else nullptr)
    double* sums_out;           //   ... block 0 stores the rank-ordered totals here for the later sweeps
    unsigned lgy, nz;       // log2(ny), local planes
    float2* rho_keep;       // optional: float copy of the density (before the mean is removed), or nullptr
    int4* zero;             // the same rows, writable: every thread clears what it has read, so that the accumulator is empty
                            // for the next spread without a separate memset on the critical path (nullptr: leave it)
    int4* zero_lo;          // peer-memory mode: this rank's own ghost planes z = -1 and z = nz (already pushed to the
    int4* zero_hi;          //   neighbours), cleared by the CTAs of the first / last local plane; else nullptr
    const longlong2* mesh64;    // WIDE instantiation: the density in 64-bit fixed point (unsharded plans only), cleared through zero64
    int4* zero64;
    unsigned* range_counter;    // 32-bit density: incremented when a cell has passed half of the range (|v| > 2^30) ...
    unsigned* h_range;          //   ... and the pinned host word that makes the next call rebuild with a smaller scale
};

// density of mesh cell pair `v` (+ ghost contribution) as float: value = v / scale
MHD float2 density_to_float(int2 v, float inv_scale) { return make_float2((float)v.x * inv_scale, (float)v.y * inv_scale); }

template <int LC, bool WIDE = false>
__global__ void __launch_bounds__(kLines * LC / kE)
fft_x_fwd_kernel(DensityIn in, const float2* __restrict__ g_tw /* length 2*LC */, float2* out,
                 unsigned lg_part /* log2 of the kx pencil width */, unsigned rows_total, const __grid_constant__ PeerOut peers,
                 const __grid_constant__ PeerSync sync) {
    // lg_part = log2(LC): out is the plain [row][kx] buffer.  Sharded: out = send buffer laid out [part][row][kx in part],
    // i.e. already packed for the slab -> pencil all-to-all.
    extern __shared__ float2 smem[];
    float2* tile = smem;
    float2* s_tw = smem + LayoutRow::size(LC);
    const int nthr = kLines * LC / kE;
    const size_t row0 = (size_t)blockIdx.x * kLines;
    load_twiddles<2 * LC>(s_tw, g_tw);
    pdl_wait(); pdl_trigger();
    peer_wait(sync);                                         // halos and partial sums of every rank have arrived
    const float inv_scale = __ldg(in.d_fx + 1);
    __shared__ double s_sum_a;
    if (threadIdx.x == 0) {
        if (in.sums_table) {                                 // rank-ordered totals (identical on every rank)
            double tot[3] = {0.0, 0.0, 0.0};
            for (unsigned r = 0; r < sync.n; ++r)
                for (int k = 0; k < 3; ++k) tot[k] += in.sums_table[4 * r + k];
            s_sum_a = tot[1];
            if (blockIdx.x == 0) { in.sums_out[0] = tot[0]; in.sums_out[1] = tot[1]; in.sums_out[2] = tot[2]; }
        } else {
            s_sum_a = in.d_sums[1];
        }
    }
    __syncthreads();
    const float mean = (float)(s_sum_a * in.inv_cells);
    const unsigned ny = 1u << in.lgy;
    // 128-bit loads: kE/2 per thread (twice as many for the 64-bit density), all issued before the first use
    int4 vin[kE / 2];
    longlong2 win[WIDE ? kE : 1];
#pragma unroll
    for (int q = 0; q < kE / 2; ++q) {
        const int idx = threadIdx.x + q * nthr;          // pair index: row w = idx / (LC/2), columns 2*(idx % (LC/2)) and +1
        if (WIDE) {
            const longlong2* src = in.mesh64 + (row0 + idx / (LC / 2)) * LC + 2 * (idx % (LC / 2));
            win[(2 * q) % (WIDE ? kE : 1)] = __ldcs(src);
            win[(2 * q + 1) % (WIDE ? kE : 1)] = __ldcs(src + 1);
        } else {
            vin[q] = __ldcs(reinterpret_cast<const int4*>(in.mesh + (row0 + idx / (LC / 2)) * LC) + idx % (LC / 2));
        }
    }
    if (!WIDE && in.range_counter) {
        // the spread accumulates in 32 bits: a cell past half of the range asks for a rebuild with a smaller scale before
        // anything can wrap (the scale leaves a factor 4 of headroom over the largest cell load at the last rebuild)
        int m = 0;
#pragma unroll
        for (int q = 0; q < kE / 2; ++q) m = max(m, max(max(abs(vin[q].x), abs(vin[q].y)), max(abs(vin[q].z), abs(vin[q].w))));
        if (m > (1 << 30)) { atomicAdd(in.range_counter, 1u); if (in.h_range) *reinterpret_cast<volatile unsigned*>(in.h_range) = 1u; }
    }
#pragma unroll
    for (int q2 = 0; q2 < kE; ++q2) {
        const int idx = threadIdx.x + (q2 >> 1) * nthr;
        const int w = idx / (LC / 2), l = 2 * (idx % (LC / 2)) + (q2 & 1);
        const size_t row = row0 + w;
        int2 v = (q2 & 1) ? make_int2(vin[q2 >> 1].z, vin[q2 >> 1].w) : make_int2(vin[q2 >> 1].x, vin[q2 >> 1].y);
        float2 r;
        if (WIDE) {
            const longlong2 v64 = win[q2 % (WIDE ? kE : 1)];
            r = make_float2(__ll2float_rn(v64.x) * inv_scale, __ll2float_rn(v64.y) * inv_scale);
        } else if (in.ghost) {
            // ranks choose their fixed-point scales independently: equal scales (the common case) add as integers, which
            // reproduces the unsharded density bit for bit; different scales add as floats
            const unsigned z = (unsigned)(row >> in.lgy), y = (unsigned)row & (ny - 1);
            const size_t msg = (size_t)ny * LC + 2;          // int2 elements per message
            float2 extra = make_float2(0.f, 0.f);
            if (z == 0) {
                const int2 a = __ldg(in.ghost + (size_t)y * LC + l);
                const float gs = __int_as_float(__ldg(&in.ghost[(size_t)ny * LC].x));
                if (gs == inv_scale) { v.x += a.x; v.y += a.y; } else { extra.x += (float)a.x * gs; extra.y += (float)a.y * gs; }
            }
            if (z == in.nz - 1) {
                const int2 a = __ldg(in.ghost + msg + (size_t)y * LC + l);
                const float gs = __int_as_float(__ldg(&in.ghost[msg + (size_t)ny * LC].x));
                if (gs == inv_scale) { v.x += a.x; v.y += a.y; } else { extra.x += (float)a.x * gs; extra.y += (float)a.y * gs; }
            }
            r = density_to_float(v, inv_scale);
            r.x += extra.x; r.y += extra.y;
        } else {
            r = density_to_float(v, inv_scale);
        }
        if (in.rho_keep) in.rho_keep[row * LC + l] = r;
        r.x -= mean; r.y -= mean;
        tile[LayoutRow::addr(w, l, LC)] = r;
    }
    if (WIDE) {
        const int4 z4 = make_int4(0, 0, 0, 0);
#pragma unroll
        for (int q = 0; q < kE / 2; ++q) {
            const int idx = threadIdx.x + q * nthr;
            int4* dst = in.zero64 + (row0 + idx / (LC / 2)) * LC + 2 * (idx % (LC / 2));
            dst[0] = z4; dst[1] = z4;
        }
    } else if (in.zero) {
        const int4 z4 = make_int4(0, 0, 0, 0);
#pragma unroll
        for (int q = 0; q < kE / 2; ++q) {
            const int idx = threadIdx.x + q * nthr;
            const size_t row = row0 + idx / (LC / 2);
            in.zero[row * (LC / 2) + idx % (LC / 2)] = z4;
            if (in.zero_lo) {
                const unsigned z = (unsigned)(row >> in.lgy), y = (unsigned)row & (ny - 1);
                if (z == 0) in.zero_lo[(size_t)y * (LC / 2) + idx % (LC / 2)] = z4;
                if (z == in.nz - 1) in.zero_hi[(size_t)y * (LC / 2) + idx % (LC / 2)] = z4;
            }
        }
    }
    __syncthreads();
    const int w = threadIdx.x & (kLines - 1), t = threadIdx.x / kLines;
    line_fft<LC, -1, 2 * LC, LayoutRow>(tile, w, t, s_tw);
    for (int idx = threadIdx.x; idx < kLines * (LC / 2 + 1); idx += nthr) {
        const int ww = idx & (kLines - 1), k = idx / kLines;
        float2& zk = tile[LayoutRow::addr(ww, k, LC)];
        float2& zp = tile[LayoutRow::addr(ww, (LC - k) % LC, LC)];
        r2c_pair(zk, zp, k, LC, s_tw[k]);
    }
    __syncthreads();
    const unsigned part_len = 1u << lg_part;
    if (peers.n) {
        // rank (l >> lg_part) owns this kx range; in its pencil [z global][y][kx] our rows start at rank * rows_total
        for (int idx = threadIdx.x; idx < kLines * LC; idx += nthr) {
            const int ww = idx / LC, l = idx % LC;
            const size_t dst = ((size_t)peers.rank * rows_total + (row0 + ww)) * part_len + (l & (part_len - 1));
            peers.ptr[l >> lg_part][dst] = tile[LayoutRow::addr(ww, l, LC)];
        }
        peer_signal(sync);                                   // phase 1: my share of every pencil is stored
        return;
    }
#pragma unroll
    for (int q = 0; q < kE / 2; ++q) {          // 128-bit stores: two coefficients of one part
        const int idx = threadIdx.x + q * nthr, ww = idx / (LC / 2), l = 2 * (idx % (LC / 2));
        const size_t dst = ((size_t)(l >> lg_part) * rows_total + (row0 + ww)) * part_len + (l & (part_len - 1));
        const float2 a = tile[LayoutRow::addr(ww, l, LC)], b = tile[LayoutRow::addr(ww, l + 1, LC)];
        *reinterpret_cast<float4*>(out + dst) = make_float4(a.x, a.y, b.x, b.y);
    }
}

// x pass, inverse (C2R)
template <int LC>
__global__ void __launch_bounds__(kLines * LC / kE)
fft_x_inv_kernel(float2* buf, const float2* __restrict__ g_tw, const float2* in, unsigned lg_part,
                 unsigned rows_total, const __grid_constant__ PeerSync sync, const double* cv_table, double* d_cv) {
    // in == buf, lg_part = log2(LC): plain in-place transform.  Sharded: in = receive buffer of the pencil -> slab
    // all-to-all, laid out [part][row][kx in part].
    extern __shared__ float2 smem[];
    float2* tile = smem;
    float2* s_tw = smem + LayoutRow::size(LC);
    const int nthr = kLines * LC / kE;
    const size_t row0 = (size_t)blockIdx.x * kLines;
    load_twiddles<2 * LC>(s_tw, g_tw);
    pdl_wait(); pdl_trigger();
    peer_wait(sync);                                         // planes and CV partials of every rank have arrived
    if (cv_table && blockIdx.x == 0 && threadIdx.x == 0) {   // the CV: rank-ordered sum of the partials
        double cv = 0.0;
        for (unsigned r = 0; r < sync.n; ++r) cv += __ldcv(cv_table + r);
        *d_cv = cv;
    }
    const unsigned part_len = 1u << lg_part;
    // 128-bit loads (two coefficients of one part: parts are at least 16 wide), all issued before the first use
    float4 vin[kE / 2];
#pragma unroll
    for (int q = 0; q < kE / 2; ++q) {
        const int idx = threadIdx.x + q * nthr, w = idx / (LC / 2), l = 2 * (idx % (LC / 2));
        const size_t src = ((size_t)(l >> lg_part) * rows_total + (row0 + w)) * part_len + (l & (part_len - 1));
        vin[q] = *reinterpret_cast<const float4*>(in + src);
    }
#pragma unroll
    for (int q = 0; q < kE / 2; ++q) {
        const int idx = threadIdx.x + q * nthr, w = idx / (LC / 2), l = 2 * (idx % (LC / 2));
        tile[LayoutRow::addr(w, l, LC)] = make_float2(vin[q].x, vin[q].y);
        tile[LayoutRow::addr(w, l + 1, LC)] = make_float2(vin[q].z, vin[q].w);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < kLines * (LC / 2 + 1); idx += nthr) {
        const int ww = idx & (kLines - 1), k = idx / kLines;
        float2& xk = tile[LayoutRow::addr(ww, k, LC)];
        float2& xp = tile[LayoutRow::addr(ww, (LC - k) % LC, LC)];
        c2r_pair(xk, xp, k, LC, s_tw[k]);
    }
    __syncthreads();
    const int w = threadIdx.x & (kLines - 1), t = threadIdx.x / kLines;
    line_fft<LC, +1, 2 * LC, LayoutRow>(tile, w, t, s_tw);
#pragma unroll
    for (int q = 0; q < kE / 2; ++q) {
        const int idx = threadIdx.x + q * nthr, ww = idx / (LC / 2), l = 2 * (idx % (LC / 2));
        const float2 a = tile[LayoutRow::addr(ww, l, LC)], b = tile[LayoutRow::addr(ww, l + 1, LC)];
        *reinterpret_cast<float4*>(buf + (row0 + ww) * LC + l) = make_float4(a.x, a.y, b.x, b.y);
    }
}

// ---------------------------------------------------------------------------------------------------
// y pass: tile = 16 consecutive kx  x  all y, fixed z.  grid = (nxh/16, nz)
// ---------------------------------------------------------------------------------------------------
// G groups of 16 lines side by side: a row of the tile is 16 G consecutive kx = 128 G contiguous bytes in memory.
// Measured on B200 (profiles/r01m_notes.md): G = 2 is slower than G = 1 (coarser CTAs), and a persistent, double-buffered
// (cp.async) variant of the G = 1 kernel gained nothing -- the sweep runs at ~3.7 TB/s of L2 traffic either way.
template <int L, int SIGN, int G>
__global__ void __launch_bounds__(G * kLines * L / kE)
fft_y_kernel(float2* __restrict__ buf, const float2* __restrict__ g_tw, unsigned nxh, const __grid_constant__ PeerOut peers,
             unsigned lg_planes /* peers.n != 0: log2 of the planes per rank */, const __grid_constant__ PeerSync sync) {
    extern __shared__ __align__(16) float2 smem[];
    using Lay = LayoutColWide<G>;
    constexpr int gthr = kLines * L / kE, nthr = G * gthr, W = kLines * G, PAIRS = W * L / 2;
    float2* tile = smem;
    float2* s_tw = smem + Lay::size(L);
    const size_t base = (size_t)blockIdx.y * L * nxh + (size_t)blockIdx.x * W;
    load_twiddles<L>(s_tw, g_tw);
    pdl_wait(); pdl_trigger();
    peer_wait(sync);                                         // forward sweep: every rank's share of my pencil has arrived
    float4* tile4 = reinterpret_cast<float4*>(tile);
#pragma unroll
    for (int q = 0; q < PAIRS / nthr; ++q) {
        const int p = threadIdx.x + q * nthr, l = p / (W / 2), c2 = p % (W / 2);
        tile4[p] = *reinterpret_cast<const float4*>(buf + base + (size_t)l * nxh + 2 * c2);
    }
    __syncthreads();
    const int g = threadIdx.x / gthr, lt = threadIdx.x % gthr;
    const int w = lt & (kLines - 1), t = lt / kLines;
    line_fft<L, SIGN, L, Lay>(tile + kLines * g, w, t, s_tw);
    float2* out;
    if (peers.n) {
        // plane z = blockIdx.y of the pencil belongs to rank z >> lg_planes; its receive buffer is laid out
        // [source rank][local plane][y][kx in the source's pencil] (what the inverse x pass unpacks)
        const unsigned z = blockIdx.y, dest = z >> lg_planes, zl = z & ((1u << lg_planes) - 1);
        out = peers.ptr[dest] + (((size_t)peers.rank << lg_planes) + zl) * L * nxh + (size_t)blockIdx.x * W;
    } else {
        out = buf + base;
    }
#pragma unroll
    for (int q = 0; q < PAIRS / nthr; ++q) {
        const int p = threadIdx.x + q * nthr, l = p / (W / 2), c2 = p % (W / 2);
        *reinterpret_cast<float4*>(out + (size_t)l * nxh + 2 * c2) = tile4[p];
    }
    peer_signal(sync);                                       // inverse sweep: phase 2, my share of every slab is stored
}

// energy partial -> per-block slot; the last block of the LAST kernel (main z pass) sums all slots in order
template <bool EXTRAS = false>
__device__ __forceinline__ void energy_block_finish(double e, const ConvParams& cp, const ExtraAcc* xa = nullptr) {
    __shared__ double red[32];
    __shared__ unsigned long long kred[32];
    __shared__ bool is_last;
    const unsigned slot = blockIdx.x, total_slots = gridDim.x;
    if (EXTRAS && xa) {
        for (int c = 0; c < 6; ++c) {
            const double r = block_sum(xa->v[c], red);
            if (threadIdx.x == 0) cp.vir_partials[6 * (size_t)slot + c] = r;
        }
        unsigned long long k = xa->key;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long y = __shfl_xor_sync(0xffffffffu, k, o); k = y > k ? y : k; }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) kred[threadIdx.x >> 5] = k;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (unsigned w = 1; w < (blockDim.x + 31) / 32; ++w) k = kred[w] > k ? kred[w] : k;
            atomicMax(cp.amax_key, k);
        }
    }
    const double r = block_sum(e, red);
    if (threadIdx.x == 0) cp.partials[slot] = r;
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = (atomicAdd(cp.ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double s = 0.0;
    for (unsigned b = threadIdx.x; b < total_slots; b += blockDim.x) s += __ldcg(cp.partials + b);
    s = block_sum(s, red);
    if (EXTRAS) {
        for (int c = 0; c < 6; ++c) {
            double v = 0.0;
            for (unsigned b = threadIdx.x; b < total_slots; b += blockDim.x) v += __ldcg(cp.vir_partials + 6 * (size_t)b + c);
            v = block_sum(v, red);
            if (threadIdx.x == 0) cp.extras_out[c] = v;
        }
    }
    if (threadIdx.x == 0) {
        *cp.d_cv = 0.5 * s;      // sum *= 1/2, OrderParameterMesh.cc:905
        *cp.ticket = 0;
    }
    if (threadIdx.x == 0)
        for (unsigned r = 0; r < cp.cv_n; ++r) reinterpret_cast<double*>(cp.cv_arena[r] + cp.cv_off)[cp.cv_rank] = 0.5 * s;
}

// ---------------------------------------------------------------------------------------------------
// z pass fused with the convolution and the CV energy: forward z FFT -> G^H -> inverse z FFT.
// tile = 16 consecutive kx  x  all z, fixed y.  Column kx = 0 is left to the plane0 blocks of the same launch.
// ---------------------------------------------------------------------------------------------------
template <int L, bool EXTRAS>
__device__ __forceinline__ void z_general_body(float2* __restrict__ buf, const float2* __restrict__ g_tw, const ConvParams& cp,
                                               unsigned bx, unsigned by, float2* smem) {
    float2* tile = smem;
    float2* s_tw = smem + LayoutCol::size(L);
    const int nthr = kLines * L / kE;
    const unsigned nxh = cp.row_len, ny = cp.ny;
    const unsigned kx0 = bx * kLines, ky = by;
    const size_t base = (size_t)ky * nxh + kx0;
    const size_t zstride = (size_t)ny * nxh;
    load_twiddles<L>(s_tw, g_tw);
    float4* tile4 = reinterpret_cast<float4*>(tile);
#pragma unroll
    for (int q = 0; q < kE / 2; ++q) {
        const int p = threadIdx.x + q * nthr, l = p / (kLines / 2), w2 = p % (kLines / 2);
        tile4[p] = *reinterpret_cast<const float4*>(buf + base + (size_t)l * zstride + 2 * w2);
    }
    __syncthreads();
    const int w = threadIdx.x & (kLines - 1), t = threadIdx.x / kLines;
    line_fft<L, -1, L, LayoutCol>(tile, w, t, s_tw);

    const double nd = cp.n_global;
    const float d = (float)(0.5 * (*cp.d_mode_sq) / nd / nd);
    const bool ky_nonneg = nonneg(ky, ny);
    double e = 0.0;
    ExtraAcc xa;
    extra_init(xa);
    for (int idx = threadIdx.x; idx < kLines * L; idx += nthr) {
        const int ww = idx & (kLines - 1);
        const unsigned kz = idx / kLines;
        if (cp.kx_off + kx0 + ww == 0) continue;
        if (EXTRAS) {
            const float fr = tile[idx].x * cp.inv_n, fi = tile[idx].y * cp.inv_n;
            extra_add(xa, cp, fr * fr + fi * fi, cp.kx_off + kx0 + ww, ky, kz, 2.0f);
        }
        tile[idx] = conv_general(tile[idx], cp.inv_n, d, ky_nonneg && nonneg(kz, L), e);
    }
    __syncthreads();
    line_fft<L, +1, L, LayoutCol>(tile, w, t, s_tw);
    const bool has_col0 = cp.kx_off + kx0 == 0;          // column kx = 0 belongs to the plane0 blocks
#pragma unroll
    for (int q = 0; q < kE / 2; ++q) {
        const int p = threadIdx.x + q * nthr, l = p / (kLines / 2), w2 = p % (kLines / 2);
        float2* dst = buf + base + (size_t)l * zstride + 2 * w2;
        const float4 v = tile4[p];
        if (has_col0 && w2 == 0) dst[1] = make_float2(v.z, v.w);
        else *reinterpret_cast<float4*>(dst) = v;
    }
    energy_block_finish<EXTRAS>(e, cp, &xa);
}

// ---------------------------------------------------------------------------------------------------
// plane kx = 0 (packed DC + Nyquist planes).  Each CTA handles 8 line pairs (ky, -ky): column 2p holds line ky,
// column 2p+1 line (ny-ky)%ny.  After the z FFT:  A(ky,kz) = (Z(ky,kz) + conj Z(-ky,-kz))/2  is the kx = 0
// spectrum, B = (Z(ky,kz) - conj Z(-ky,-kz))/(2i) the kx = nx/2 spectrum; both are convolved, re-packed as
// A' + i B' and transformed back.  grid = ceil((ny/2+1)/8)
// ---------------------------------------------------------------------------------------------------
template <int L, bool EXTRAS>
__device__ __forceinline__ void z_plane0_body(float2* __restrict__ buf, const float2* __restrict__ g_tw, const ConvParams& cp,
                                              unsigned blk, float2* smem) {
    float2* tile = smem;
    float2* s_tw = smem + LayoutCol::size(L);
    constexpr int nthr = kLines * L / kE;
    constexpr int per_thread = kLines * L / nthr;   // = 8
    const unsigned nxh = cp.row_len, ny = cp.ny;
    const size_t zstride = (size_t)ny * nxh;
    load_twiddles<L>(s_tw, g_tw);
    for (int idx = threadIdx.x; idx < kLines * L; idx += nthr) {
        const int w = idx & (kLines - 1), l = idx / kLines;
        const unsigned kyp = blk * (kLines / 2) + (w >> 1);
        float2 v = make_float2(0.f, 0.f);
        if (kyp <= ny / 2) {
            const unsigned ky = (w & 1) ? (ny - kyp) % ny : kyp;
            v = buf[(size_t)ky * nxh + (size_t)l * zstride];
        }
        tile[idx] = v;
    }
    __syncthreads();
    const int w = threadIdx.x & (kLines - 1), t = threadIdx.x / kLines;
    line_fft<L, -1, L, LayoutCol>(tile, w, t, s_tw);

    const double nd = cp.n_global;
    const float d = (float)(0.5 * (*cp.d_mode_sq) / nd / nd);
    double e = 0.0;
    ExtraAcc xa;
    extra_init(xa);
    float dc = 0.f;
    if (cp.dc_restore) {
        const float mean = (float)(cp.d_mode_sq[1] * cp.inv_cells);       // the value the forward x sweep subtracted from every cell
        dc = (float)((double)mean / cp.inv_cells * (double)cp.inv_n);
    }
    float2 outv[per_thread];
#pragma unroll
    for (int it = 0; it < per_thread; ++it) {
        const int idx = threadIdx.x + it * nthr;
        const int ww = idx & (kLines - 1);
        const unsigned kz = idx / kLines;
        const unsigned kyp = blk * (kLines / 2) + (ww >> 1);
        const unsigned pky = (ny - kyp) % ny;
        const unsigned ky = (ww & 1) ? pky : kyp;
        const bool valid = kyp <= ny / 2;
        const bool counted = valid && (!(ww & 1) || pky != kyp);
        if (EXTRAS && counted) {
            // the kx = 0 mode A and the kx = nx/2 mode B of this (ky, kz), untangled as in conv_plane0
            const float2 z1 = tile[idx], z2 = cconj(tile[((L - kz) % L) * kLines + (ww ^ 1)]);
            const float ar = 0.5f * (z1.x + z2.x) * cp.inv_n + ((ky == 0 && kz == 0) ? dc : 0.f), ai = 0.5f * (z1.y + z2.y) * cp.inv_n;
            const float br = 0.5f * (z1.y - z2.y) * cp.inv_n, bi = -0.5f * (z1.x - z2.x) * cp.inv_n;
            extra_add(xa, cp, ar * ar + ai * ai, 0u, ky, kz, 1.0f);
            extra_add(xa, cp, br * br + bi * bi, cp.nx / 2, ky, kz, 1.0f);
        }
        outv[it] = conv_plane0(tile[idx], tile[((L - kz) % L) * kLines + (ww ^ 1)], cp.inv_n, d, ky, kz, ny, L, counted, e, dc);
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < per_thread; ++it) tile[threadIdx.x + it * nthr] = outv[it];
    __syncthreads();
    line_fft<L, +1, L, LayoutCol>(tile, w, t, s_tw);
    for (int idx = threadIdx.x; idx < kLines * L; idx += nthr) {
        const int ww = idx & (kLines - 1), l = idx / kLines;
        const unsigned kyp = blk * (kLines / 2) + (ww >> 1);
        const unsigned pky = (ny - kyp) % ny;
        if (kyp > ny / 2) continue;
        if ((ww & 1) && pky == kyp) continue;
        const unsigned ky = (ww & 1) ? pky : kyp;
        buf[(size_t)ky * nxh + (size_t)l * zstride] = tile[idx];
    }
    energy_block_finish<EXTRAS>(e, cp, &xa);
}

// one launch: blocks [0, n_blocks_plane0) untangle the kx = 0 slot, the rest are (kx tile, ky) blocks of the general case
// EXTRAS: with the q_max / virial epilogues (a separate instantiation: their accumulators cost registers the default step
// should not pay for -- measured 0.059 -> 0.068 ms at C4 when they were a run-time branch)
template <int L, int MINB = ((kLines * L / kE) <= 256 ? 3 : 1), bool EXTRAS = false>
__global__ void __launch_bounds__(kLines * L / kE, MINB)
fft_z_fused_kernel(float2* __restrict__ buf, const float2* __restrict__ g_tw, ConvParams cp) {
    extern __shared__ float2 smem[];
    pdl_wait(); pdl_trigger();
    if (blockIdx.x < cp.n_blocks_plane0) {
        z_plane0_body<L, EXTRAS>(buf, g_tw, cp, blockIdx.x, smem);
    } else {
        const unsigned b = blockIdx.x - cp.n_blocks_plane0, ntx = cp.row_len / kLines;
        z_general_body<L, EXTRAS>(buf, g_tw, cp, b % ntx, b / ntx, smem);
    }
}

}  // namespace fft
}  // namespace metad
