// mesh_kernels.cuh -- particle <-> mesh kernels of the OrderParameterMesh path (sm_100a).
//
// Reference behaviour restated (CPU path = parity target): OrderParameterMesh.cc:517-640 (assignParticles),
// :457-483 (TSC weights), :749-864 (interpolateForces).  Reference GPU kernels replaced:
// gpu_bin_particles_kernel / gpu_assign_binned_particles_to_scratch_kernel / gpu_reduce_scratch_kernel /
// gpu_compute_forces_kernel (OrderParameterMeshGPU.cu:90-364, 566-769): non-deterministic atomicInc binning with
// an overflow-retry loop, a 27x scratch mesh, texture gathers.
//
// Here: particles are counting-sorted (stable: ties keep the input order) by a TILE-MAJOR cell key (tile of T^3
// cells, T = 8 or 16), so that
//   * spreading is atomics-free: one thread per cell COLUMN of a tile walks the column in z with the 3x9 partial
//     sums of three planes in registers; columns exchange their x,y taps through nine write-once "replica" planes
//     in shared memory (no read-modify-write, one barrier per plane); every tile is flushed as one contiguous
//     padded tile and a merge pass sums the <= 8 overlapping padded tiles per cell in a fixed order.  Summation
//     orders are fixed, so results are bitwise reproducible;
//   * force interpolation is thread-per-particle over a contiguous particle range per tile, reading a
//     shared-memory tile of Re(IFFT(G)) with halo.
// Cell indices are computed with non-contracted IEEE fp32 operations and are bit-exact against the reference's
// single-precision arithmetic; in-cell offsets are evaluated in fp64 (they only need to be accurate).
#pragma once
#include "common.cuh"

#ifndef MHD
#define MHD __host__ __device__ __forceinline__
#endif

namespace metad {
namespace mesh {

struct Geom {
    unsigned nx, ny, nz;        // mesh points of the LOCAL mesh (powers of two); nz = planes of this z slab
    unsigned lgx, lgy, lgz;     // log2 of the above
    unsigned nzg, z0;           // global number of z planes and first global plane of this slab (nzg = nz, z0 = 0 if unsharded)
    unsigned slab;              // 1: z is not periodic locally, the planes z0-1 and z0+nz belong to the neighbour ranks
    unsigned lgT;               // log2 of the tile edge T (3 or 4)
    unsigned ntx, nty, ntz;     // tiles per dimension (powers of two)
    unsigned lgtx, lgty, lgtz;  // log2 of the above
    float lo[3], L[3];          // single-precision box (HOOMD SINGLE_PRECISION BoxDim)
    double dlo[3], dscale[3];   // fp64: lo and n/L for the in-cell offset
};

MHD void geom_set_dims(Geom& g, unsigned nx, unsigned ny, unsigned nz, unsigned lgT) {
    auto lg = [](unsigned n) { unsigned l = 0; while ((1u << l) < n) ++l; return l; };
    g.nx = nx; g.ny = ny; g.nz = nz;
    g.lgx = lg(nx); g.lgy = lg(ny); g.lgz = lg(nz);
    g.nzg = nz; g.z0 = 0; g.slab = 0;
    g.lgT = lgT;
    g.lgtx = g.lgx - lgT; g.lgty = g.lgy - lgT; g.lgtz = g.lgz - lgT;
    g.ntx = 1u << g.lgtx; g.nty = 1u << g.lgty; g.ntz = 1u << g.lgtz;
}

MHD unsigned tile_edge(const Geom& g) { return 1u << g.lgT; }
MHD unsigned cells_per_tile(const Geom& g) { return 1u << (3 * g.lgT); }
MHD unsigned padded_edge(const Geom& g) { return (1u << g.lgT) + 2; }
MHD unsigned num_tiles(const Geom& g) { return 1u << (g.lgtx + g.lgty + g.lgtz); }
MHD unsigned tile_index(unsigned tx, unsigned ty, unsigned tz, const Geom& g) { return (((tz << g.lgty) + ty) << g.lgtx) + tx; }
MHD void tile_coords(unsigned tile, const Geom& g, unsigned& tx, unsigned& ty, unsigned& tz) {
    tx = tile & (g.ntx - 1);
    ty = (tile >> g.lgtx) & (g.nty - 1);
    tz = tile >> (g.lgtx + g.lgty);
}

// non-contracted IEEE single-precision helpers (host: plain ops, the emulation is built without FMA contraction)
MHD float f_sub(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
MHD float f_div(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}
MHD float f_mul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}

// cell coordinate along one axis: OrderParameterMesh.cc:543-561 with BoxDim::makeFraction = (x - lo)/L.
//   f = (x - lo)/L ; r = f*n ; i = (int) r (truncation) ; i == n -> 0
// Out-of-box input (which HOOMD never hands over) is folded back periodically instead of indexing out of range.
MHD int cell_coord(float x, float lo, float L, unsigned n) {
    const float f = f_div(f_sub(x, lo), L);
    const float r = f_mul(f, (float)n);
    int i = (int)r;
    if (i == (int)n) i = 0;
    if (i < 0 || i > (int)n) {
        i %= (int)n;
        if (i < 0) i += (int)n;
    }
    return i;
}

// tile-major key: tile index * T^3 + local cell index (x fastest inside the tile)
MHD unsigned key_of(unsigned ix, unsigned iy, unsigned iz, const Geom& g) {
    const unsigned T1 = (1u << g.lgT) - 1;
    const unsigned tile = tile_index(ix >> g.lgT, iy >> g.lgT, iz >> g.lgT, g);
    const unsigned local = ((((iz & T1) << g.lgT) + (iy & T1)) << g.lgT) + (ix & T1);
    return (tile << (3 * g.lgT)) + local;
}
MHD void cell_of_key(unsigned key, const Geom& g, unsigned& ix, unsigned& iy, unsigned& iz) {
    const unsigned T1 = (1u << g.lgT) - 1;
    const unsigned local = key & ((1u << (3 * g.lgT)) - 1);
    unsigned tx, ty, tz;
    tile_coords(key >> (3 * g.lgT), g, tx, ty, tz);
    ix = (tx << g.lgT) + (local & T1);
    iy = (ty << g.lgT) + ((local >> g.lgT) & T1);
    iz = (tz << g.lgT) + (local >> (2 * g.lgT));
}

// in-cell offset in cell units, s in [-1/2, 1/2] (OrderParameterMesh.cc:565-573: minimum-image distance to the
// cell centre through makeCoordinates/minImage/makeFraction; evaluated here directly in fp64)
MHD float cell_shift(float x, unsigned i, int axis, const Geom& g) {
    // i is the GLOBAL cell coordinate (for z: z0 + local plane)
    const unsigned n = axis == 0 ? g.nx : (axis == 1 ? g.ny : g.nzg);
    double s = ((double)x - g.dlo[axis]) * g.dscale[axis] - ((double)i + 0.5);
    const double half = 0.5 * (double)n;
    if (s > half) s -= (double)n;
    else if (s < -half) s += (double)n;
    return (float)s;
}

// TSC weights of the three taps i = -1, 0, +1 for offset s (assignTSC, OrderParameterMesh.cc:457-468, with
// d = s - i):  W(s+1) = (1/2)(1/2 - s)^2,  W(s) = 3/4 - s^2,  W(s-1) = (1/2)(1/2 + s)^2
MHD void tsc(float s, float (&w)[3]) {
    const float a = 0.5f - s, b = 0.5f + s;
    w[0] = 0.5f * a * a;
    w[1] = 0.75f - s * s;
    w[2] = 0.5f * b * b;
}
// derivative weights (assignTSCderiv, :470-483):  W'(s+1) = s - 1/2,  W'(s) = -2 s,  W'(s-1) = s + 1/2
MHD void tsc_deriv(float s, float (&w)[3]) {
    w[0] = s - 0.5f;
    w[1] = -2.0f * s;
    w[2] = s + 0.5f;
}

// ---------------------------------------------------------------------------------------------------
// per-thread bodies shared by the kernels and the CPU emulation (tests/cpu_emul/mesh_emul.cu)
// ---------------------------------------------------------------------------------------------------

// index inside the padded tile (edge P = T+2) of tap (i,j,k) in {0,1,2}^3 of local cell (lx,ly,lz)
MHD unsigned padded_index(unsigned lx, unsigned ly, unsigned lz, int i, int j, int k, unsigned P) {
    return ((lz + k) * P + (ly + j)) * P + (lx + i);
}

// Separable TSC weights of one particle relative to its cell: w[0..2] = a*Wx(tap -1,0,+1), w[3..5] = Wy, w[6..8] = Wz
MHD void spread_weights(float4 p /* x,y,z,a */, unsigned ix, unsigned iy, unsigned iz, const Geom& g, float (&w)[9]) {
    float wx[3], wy[3], wz[3];
    tsc(cell_shift(p.x, ix, 0, g), wx);
    tsc(cell_shift(p.y, iy, 1, g), wy);
    tsc(cell_shift(p.z, iz, 2, g), wz);
#pragma unroll
    for (int i = 0; i < 3; ++i) { w[i] = p.w * wx[i]; w[3 + i] = wy[i]; w[6 + i] = wz[i]; }
}
// Rolling accumulators of a cell column: acc[k*9 + i*3 + j], k = z tap (plane lz-1+k), i = x tap, j = y tap
MHD void spread_accumulate9(const float (&w)[9], float (&acc)[27]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float axy = w[i] * w[3 + j];
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[k * 9 + i * 3 + j] = fmaf(axy, w[6 + k], acc[k * 9 + i * 3 + j]);
        }
}
// Exchange in x,y without atomics or read-modify-write: column (lx,ly) stores its nine (i,j) partial sums of a finished
// plane into nine "replica" planes at padded position (lx+i, ly+j); replica r = i*3+j is written at most once per
// position.  The value of padded position (px,py) is the sum over the replicas whose source column exists.
MHD unsigned replica_index(int r, unsigned px, unsigned py, unsigned P) { return (r * P + py) * P + px; }
MHD float reduce_replicas(const float* rep, unsigned px, unsigned py, unsigned T) {
    const unsigned P = T + 2;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            // source column (px - i, py - j) must lie inside the tile
            if (px >= (unsigned)i && px - i < T && py >= (unsigned)j && py - j < T) sum += rep[replica_index(i * 3 + j, px, py, P)];
        }
    return sum;
}

// merge: value of mesh cell (x,y,z) = sum over the <= 8 padded tiles that cover it, in fixed (z,y,x) order.
// merge_plane sums the x,y candidates of padded plane pz of tile row tz.
MHD float merge_plane(const float* __restrict__ scratch, unsigned x, unsigned y, unsigned tz, unsigned pz, const Geom& g) {
    const unsigned T = 1u << g.lgT, P = T + 2, P3 = P * P * P;
    const unsigned tx = x >> g.lgT, ty = y >> g.lgT;
    const unsigned lx = x & (T - 1), ly = y & (T - 1);
    unsigned ctx[2], cpx[2], cty[2], cpy[2];
    int nxc = 1, nyc = 1;
    ctx[0] = tx; cpx[0] = lx + 1;
    if (lx == 0) { ctx[1] = (tx + g.ntx - 1) & (g.ntx - 1); cpx[1] = T + 1; nxc = 2; }
    else if (lx == T - 1) { ctx[1] = (tx + 1) & (g.ntx - 1); cpx[1] = 0; nxc = 2; }
    cty[0] = ty; cpy[0] = ly + 1;
    if (ly == 0) { cty[1] = (ty + g.nty - 1) & (g.nty - 1); cpy[1] = T + 1; nyc = 2; }
    else if (ly == T - 1) { cty[1] = (ty + 1) & (g.nty - 1); cpy[1] = 0; nyc = 2; }
    float sum = 0.f;
    for (int b = 0; b < nyc; ++b)
        for (int a = 0; a < nxc; ++a) {
            const unsigned tile = tile_index(ctx[a], cty[b], tz, g);
            sum += scratch[(size_t)tile * P3 + (pz * P + cpy[b]) * P + cpx[a]];
        }
    return sum;
}
// In slab mode the halo planes below the first / above the last local tile row are NOT wrapped around: they are
// extracted by merge_plane(..., tz = 0, pz = 0) / (tz = ntz-1, pz = T+1) and added by the neighbour rank.
MHD float merge_cell(const float* __restrict__ scratch, unsigned x, unsigned y, unsigned z, const Geom& g) {
    const unsigned T = 1u << g.lgT;
    const unsigned tz = z >> g.lgT, lz = z & (T - 1);
    float sum = merge_plane(scratch, x, y, tz, lz + 1, g);
    if (lz == 0 && !(g.slab && tz == 0)) sum += merge_plane(scratch, x, y, (tz + g.ntz - 1) & (g.ntz - 1), T + 1, g);
    else if (lz == T - 1 && !(g.slab && tz == g.ntz - 1)) sum += merge_plane(scratch, x, y, (tz + 1) & (g.ntz - 1), 0, g);
    return sum;
}

// force on one particle from the padded tile of Re(IFFT(G)) (interpolateForces, OrderParameterMesh.cc:812-860):
//   F = -(a) * sum_taps inv * [ nb1 W'x Wy Wz + nb2 Wx W'y Wz + nb3 Wx Wy W'z ],  nb_a = n_a * b_a (no 2 pi)
// evaluated as three separable contractions; returns the three scalar sums (Sx,Sy,Sz).
MHD void gather_sums(const float* tile, unsigned lx, unsigned ly, unsigned lz, unsigned P, const float (&wx)[3],
                     const float (&wy)[3], const float (&wz)[3], const float (&dx)[3], const float (&dy)[3],
                     const float (&dz)[3], float& Sx, float& Sy, float& Sz) {
    Sx = 0.f; Sy = 0.f; Sz = 0.f;
    const float* base = tile + (lz * P + ly) * P + lx;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float tw = 0.f, td = 0.f, tz = 0.f;   // sum_j {Wy, W'y, Wy} * sum_k {Wz, Wz, W'z} inv
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float u = 0.f, v = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float val = base[(k * P + j) * P + i];
                u = fmaf(wz[k], val, u);
                v = fmaf(dz[k], val, v);
            }
            tw = fmaf(wy[j], u, tw);
            td = fmaf(dy[j], u, td);
            tz = fmaf(wy[j], v, tz);
        }
        Sx = fmaf(dx[i], tw, Sx);
        Sy = fmaf(wx[i], td, Sy);
        Sz = fmaf(wx[i], tz, Sz);
    }
}

struct ForceParams {
    float nb1[3], nb2[3], nb3[3];   // n_a * b_a, b_a = reciprocal lattice vectors of the box without 2 pi (:761-769)
    double two_over_n;              // 2 / N_global  (:858)
};

// (lx,ly,lz): local cell inside the tile; (ix,iy,iz): global cell; scale = (2/N) * bias rounded to float
MHD float4 gather_force(float4 p /* x,y,z,a */, unsigned ix, unsigned iy, unsigned iz, unsigned lx, unsigned ly, unsigned lz,
                        const float* tile, const Geom& g, const ForceParams& fp, float scale) {
    const unsigned P = (1u << g.lgT) + 2;
    const float sx = cell_shift(p.x, ix, 0, g), sy = cell_shift(p.y, iy, 1, g), sz = cell_shift(p.z, iz, 2, g);
    float wx[3], wy[3], wz[3], dx[3], dy[3], dz[3];
    tsc(sx, wx); tsc(sy, wy); tsc(sz, wz);
    tsc_deriv(sx, dx); tsc_deriv(sy, dy); tsc_deriv(sz, dz);
    float Sx, Sy, Sz;
    gather_sums(tile, lx, ly, lz, P, wx, wy, wz, dx, dy, dz, Sx, Sy, Sz);
    const float m = -p.w * scale;
    return make_float4(m * (fp.nb1[0] * Sx + fp.nb2[0] * Sy + fp.nb3[0] * Sz), m * (fp.nb1[1] * Sx + fp.nb2[1] * Sy + fp.nb3[1] * Sz),
                       m * (fp.nb1[2] * Sx + fp.nb2[2] * Sy + fp.nb3[2] * Sz), 0.f);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
constexpr int kBinThreads = 256;

// bin: key + rank (arrival order inside the cell) per particle, per-cell counts, sum a^2 and sum a
__global__ void __launch_bounds__(kBinThreads)
mesh_bin_kernel(const float4* __restrict__ postype, unsigned N, Geom g, const float* __restrict__ mode,
                unsigned* __restrict__ keys, unsigned* __restrict__ ranks, unsigned* __restrict__ count,
                double* __restrict__ sums /* [0] sum a^2, [1] sum a, [2] misplaced particles */) {
    double sq = 0.0, s1 = 0.0;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const float4 p = ld_stream(postype + i);
        const unsigned ix = cell_coord(p.x, g.lo[0], g.L[0], g.nx);
        const unsigned iy = cell_coord(p.y, g.lo[1], g.L[1], g.ny);
        // global plane -> local plane of this slab; a particle outside the slab is a caller error: it is folded into
        // the slab (keeps memory safe) and counted in sums[2]
        unsigned iz = (unsigned)cell_coord(p.z, g.lo[2], g.L[2], g.nzg) - g.z0;
        if (iz >= g.nz) { iz &= (g.nz - 1); atomicAdd(sums + 2, 1.0); }
        const unsigned key = key_of(ix, iy, iz, g);
        keys[i] = key;
        ranks[i] = atomicAdd(count + key, 1u);
        const float a = __ldg(mode + __float_as_int(p.w));
        sq += (double)a * (double)a;        // m_mode_sq, OrderParameterMesh.cc:623
        s1 += (double)a;
    }
    __shared__ double red[32];
    const double tsq = block_sum(sq, red);
    const double ts1 = block_sum(s1, red);
    if (threadIdx.x == 0) { atomicAdd(sums, tsq); atomicAdd(sums + 1, ts1); }
}

// ---- exclusive scan of count[0..n) -> start[0..n].  Three launches: per-block sums, scan of the block sums (single
// block), apply.  Each thread owns V consecutive uint4 (n must be a multiple of 4096*V).  The apply pass also
// clears count[] for the next step.
constexpr int kScanThreads = 1024;

// block-wide exclusive scan of one value per thread (blockDim.x == 1024); total returned to every thread
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned* sm /* >= 33 */, unsigned& total) {
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += y;
    }
    if (lane == 31) sm[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const unsigned w = sm[lane];
        unsigned wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (unsigned)o) wi += y;
        }
        sm[lane] = wi - w;
        if (lane == 31) sm[32] = wi;
    }
    __syncthreads();
    const unsigned excl = sm[wid] + incl - v;
    total = sm[32];
    __syncthreads();
    return excl;
}

template <int V>
__global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(const uint4* __restrict__ count4, unsigned* __restrict__ block_sums) {
    __shared__ unsigned sm[33];
    const uint4* src = count4 + ((size_t)blockIdx.x * kScanThreads + threadIdx.x) * V;
    uint4 v[V];
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = src[k];
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < V; ++k) s += v[k].x + v[k].y + v[k].z + v[k].w;
    unsigned total;
    block_excl_scan(s, sm, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads)
scan_offsets_kernel(unsigned* __restrict__ block_sums, unsigned nb) {
    __shared__ unsigned sm[33];
    unsigned carry = 0;
    for (unsigned base = 0; base < nb; base += kScanThreads) {
        const unsigned i = base + threadIdx.x;
        const unsigned v = i < nb ? block_sums[i] : 0u;
        unsigned total;
        const unsigned excl = block_excl_scan(v, sm, total);
        if (i < nb) block_sums[i] = carry + excl;
        carry += total;
    }
    if (threadIdx.x == 0) block_sums[nb] = carry;
}

template <int V>
__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(uint4* __restrict__ count4, const unsigned* __restrict__ block_sums, unsigned* __restrict__ start,
                  unsigned n) {
    __shared__ unsigned sm[33];
    const size_t i4 = ((size_t)blockIdx.x * kScanThreads + threadIdx.x) * V;
    uint4 v[V];
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = count4[i4 + k];
#pragma unroll
    for (int k = 0; k < V; ++k) count4[i4 + k] = make_uint4(0u, 0u, 0u, 0u);
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < V; ++k) s += v[k].x + v[k].y + v[k].z + v[k].w;
    unsigned total;
    unsigned run = block_excl_scan(s, sm, total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        uint4 o;
        o.x = run; o.y = o.x + v[k].x; o.z = o.y + v[k].y; o.w = o.z + v[k].z;
        run = o.w + v[k].w;
        reinterpret_cast<uint4*>(start)[i4 + k] = o;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) start[n] = run;
}

// place: slot[start[key] + arrival rank] = particle index.  The arrival rank comes from atomics and is not
// reproducible; the reorder pass below turns it into the stable rank (ascending particle index inside a cell).
__global__ void __launch_bounds__(kBinThreads)
mesh_place_kernel(unsigned N, const unsigned* __restrict__ keys, const unsigned* __restrict__ ranks,
                  const unsigned* __restrict__ start, unsigned* __restrict__ slot) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) slot[__ldg(start + keys[i]) + ranks[i]] = i;
}

// reorder (one thread per slot): stable position inside the cell = number of cell mates with a smaller particle
// index; sorted[...] = {x, y, z, a(type)}, perm[...] = particle index, skey[...] = key.
__global__ void __launch_bounds__(kBinThreads)
mesh_reorder_kernel(const float4* __restrict__ postype, unsigned N, const float* __restrict__ mode,
                    const unsigned* __restrict__ keys, const unsigned* __restrict__ start, const unsigned* __restrict__ slot,
                    float4* __restrict__ sorted, unsigned* __restrict__ perm, unsigned* __restrict__ skey) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += stride) {
        const unsigned i = slot[j];
        const unsigned key = __ldg(keys + i);
        const unsigned s = __ldg(start + key), e = __ldg(start + key + 1);
        unsigned dst = s;
        for (unsigned m = s; m < e; ++m) dst += (__ldg(slot + m) < i) ? 1u : 0u;
        float4 p = __ldg(postype + i);
        p.w = __ldg(mode + __float_as_int(p.w));
        sorted[dst] = p;
        perm[dst] = i;
        skey[dst] = key;
    }
}

// spread: one CTA per tile, one thread per cell COLUMN (lx,ly); the thread walks its column in z keeping the
// 3x9 partial sums of the planes lz-1, lz, lz+1 in registers.  Per plane:
//   phase 1  (thread per particle, balanced): separable weights of the plane's particles -> shared memory
//   phase 2  (thread per cell): accumulate the cell's particles in slot order (= ascending particle index)
//   flush    the finished plane lz-1: nine replica stores per column, one barrier, replica reduction -> padded tile
// No atomics, no shared-memory read-modify-write.
constexpr int kSpreadCap = 512;     // particles per phase-1 chunk
template <int LGT>
__global__ void __launch_bounds__(1 << (2 * LGT), LGT == 4 ? 3 : 8)
mesh_spread_kernel(const float4* __restrict__ sorted, const unsigned* __restrict__ skey, const unsigned* __restrict__ start,
                   Geom g, float* __restrict__ scratch) {
    constexpr unsigned T = 1u << LGT, P = T + 2, PP = P * P, NT = T * T;
    __shared__ float wbuf[9 * kSpreadCap];
    __shared__ float rep[2][9 * PP];
    const unsigned tid = threadIdx.x, lx = tid & (T - 1), ly = tid >> LGT;
    const unsigned tile_id = blockIdx.x;
    unsigned tx, ty, tz;
    tile_coords(tile_id, g, tx, ty, tz);
    float* out = scratch + (size_t)tile_id * PP * P;
    float acc[27];
#pragma unroll
    for (int r = 0; r < 27; ++r) acc[r] = 0.f;
    int buf = 0;
    for (unsigned lz = 0; lz < T + 2; ++lz) {
        if (lz < T) {
            const unsigned key0 = (tile_id << (3 * LGT)) + lz * NT;
            const unsigned s_plane = __ldg(start + key0), e_plane = __ldg(start + key0 + NT);
            const unsigned s = __ldg(start + key0 + tid), e = __ldg(start + key0 + tid + 1);
            for (unsigned c0 = s_plane; c0 < e_plane; c0 += kSpreadCap) {
                const unsigned c1 = min(c0 + (unsigned)kSpreadCap, e_plane);
                for (unsigned j = c0 + tid; j < c1; j += NT) {
                    const unsigned local = __ldg(skey + j) & (NT - 1);     // cell inside the plane
                    float w[9];
                    spread_weights(sorted[j], (tx << LGT) + (local & (T - 1)), (ty << LGT) + (local >> LGT), g.z0 + (tz << LGT) + lz, g, w);
#pragma unroll
                    for (int c = 0; c < 9; ++c) wbuf[c * kSpreadCap + (j - c0)] = w[c];
                }
                __syncthreads();
                const unsigned a = max(s, c0), b = min(e, c1);
                for (unsigned j = a; j < b; ++j) {
                    float w[9];
#pragma unroll
                    for (int c = 0; c < 9; ++c) w[c] = wbuf[c * kSpreadCap + (j - c0)];
                    spread_accumulate9(w, acc);
                }
                __syncthreads();
            }
        }
        // flush the finished plane: padded z index lz (= tile plane lz - 1)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) rep[buf][replica_index(i * 3 + j, lx + i, ly + j, P)] = acc[i * 3 + j];
        __syncthreads();
        for (unsigned idx = tid; idx < PP; idx += NT) out[(size_t)lz * PP + idx] = reduce_replicas(rep[buf], idx % P, idx / P, T);
        buf ^= 1;
#pragma unroll
        for (int r = 0; r < 9; ++r) { acc[r] = acc[9 + r]; acc[9 + r] = acc[18 + r]; acc[18 + r] = 0.f; }
    }
}

// merge: mesh[x + nx (y + ny z)] = sum of covering padded tiles - mean (DC removal, see mesh.cu)
__global__ void __launch_bounds__(256)
mesh_merge_kernel(const float* __restrict__ scratch, Geom g, const double* __restrict__ sums, float* __restrict__ rho,
                  float* __restrict__ rho_keep) {
    const size_t M = (size_t)g.nx * g.ny * g.nz;
    // slab mode: the mean needs the GLOBAL sum, it is subtracted later (mesh_add_ghost_kernel)
    const float mean = g.slab ? 0.f : (float)(sums[1] / (double)M);
    const unsigned T = 1u << g.lgT, P = T + 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < M; c += stride) {
        const unsigned x = (unsigned)(c & (g.nx - 1)), y = (unsigned)((c >> g.lgx) & (g.ny - 1)), z = (unsigned)(c >> (g.lgx + g.lgy));
        const unsigned lx = x & (T - 1), ly = y & (T - 1), lz = z & (T - 1);
        float v;
        if (lx != 0 && lx != T - 1 && ly != 0 && ly != T - 1 && lz != 0 && lz != T - 1) {
            // interior cell of its tile: a single contribution
            const unsigned tile = tile_index(x >> g.lgT, y >> g.lgT, z >> g.lgT, g);
            v = __ldg(scratch + (size_t)tile * (P * P * P) + ((lz + 1) * P + (ly + 1)) * P + (lx + 1));
        } else {
            v = merge_cell(scratch, x, y, z, g);
        }
        if (rho_keep) rho_keep[c] = v;
        rho[c] = v - mean;
    }
}

// slab mode: the two halo planes that belong to the neighbour ranks: ghost[0] = plane z0-1, ghost[1] = plane z0+nz
__global__ void __launch_bounds__(256)
mesh_ghost_extract_kernel(const float* __restrict__ scratch, Geom g, float* __restrict__ ghost) {
    const unsigned plane = g.nx * g.ny, T = 1u << g.lgT;
    for (unsigned c = blockIdx.x * blockDim.x + threadIdx.x; c < 2 * plane; c += gridDim.x * blockDim.x) {
        const unsigned which = c >= plane, cc = which ? c - plane : c;
        const unsigned x = cc & (g.nx - 1), y = cc >> g.lgx;
        ghost[c] = which ? merge_plane(scratch, x, y, g.ntz - 1, T + 1, g) : merge_plane(scratch, x, y, 0, 0, g);
    }
}
// slab mode: add the planes received from the neighbours (recv[0] -> first local plane, recv[1] -> last local plane)
// and subtract the global mean density (DC removal)
__global__ void __launch_bounds__(256)
mesh_add_ghost_kernel(float* __restrict__ rho, Geom g, const float* __restrict__ recv, const double* __restrict__ sums_global,
                      float* __restrict__ rho_keep) {
    const size_t M = (size_t)g.nx * g.ny * g.nz, plane = (size_t)g.nx * g.ny;
    const float mean = (float)(sums_global[1] / ((double)plane * (double)g.nzg));
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < M; c += stride) {
        float v = rho[c];
        if (c < plane) v += recv[c];
        if (c >= M - plane) v += recv[plane + (c - (M - plane))];
        if (rho_keep) rho_keep[c] = v;
        rho[c] = v - mean;
    }
}

// gather: one CTA per tile; shared tile of Re(IFFT(G)) with halo; one thread per particle of the tile
constexpr int kGatherThreads = 256;
template <int LGT>
__global__ void __launch_bounds__(kGatherThreads)
mesh_gather_kernel(const float4* __restrict__ sorted, const unsigned* __restrict__ perm, const unsigned* __restrict__ skey,
                   const unsigned* __restrict__ start, Geom g, const float* __restrict__ inv,
                   const float* __restrict__ ghost /* slab mode: planes z0-1 and z0+nz of Re IFFT(G) */, ForceParams fp,
                   const double* __restrict__ d_bias, float4* __restrict__ force) {
    constexpr unsigned T = 1u << LGT, P = T + 2, P3 = P * P * P;
    __shared__ float tile[P3];
    const unsigned tile_id = blockIdx.x;
    const unsigned s = __ldg(start + (tile_id << (3 * LGT))), e = __ldg(start + ((tile_id + 1) << (3 * LGT)));
    if (e == s) return;                                   // empty tile: nothing to interpolate
    unsigned tx, ty, tz;
    tile_coords(tile_id, g, tx, ty, tz);
    // padded tile, one row (P floats along x) per warp iteration
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned row = warp; row < P * P; row += kGatherThreads / 32) {
        const unsigned py = row % P, pz = row / P;
        const unsigned y = ((ty << LGT) + py + g.ny - 1) & (g.ny - 1);
        const int zl = (int)((tz << LGT) + pz) - 1;                  // local plane, -1 and nz are halo planes
        const float* src;
        if (g.slab && zl < 0) src = ghost + (size_t)g.nx * y;
        else if (g.slab && zl >= (int)g.nz) src = ghost + (size_t)g.nx * (g.ny + y);
        else src = inv + (size_t)g.nx * (y + (size_t)g.ny * ((unsigned)(zl + (int)g.nz) & (g.nz - 1)));
        if (lane < P) {
            const unsigned x = ((tx << LGT) + lane + g.nx - 1) & (g.nx - 1);
            tile[row * P + lane] = __ldg(src + x);
        }
    }
    __syncthreads();
    const float scale = (float)(fp.two_over_n * *d_bias);
    for (unsigned j = s + threadIdx.x; j < e; j += kGatherThreads) {
        const unsigned local = __ldg(skey + j) & ((1u << (3 * LGT)) - 1);
        const unsigned lx = local & (T - 1), ly = (local >> LGT) & (T - 1), lz = local >> (2 * LGT);
        force[__ldg(perm + j)] = gather_force(sorted[j], (tx << LGT) + lx, (ty << LGT) + ly, g.z0 + (tz << LGT) + lz, lx, ly, lz, tile, g, fp, scale);
    }
}
#endif  // __CUDACC__

}  // namespace mesh
}  // namespace metad
