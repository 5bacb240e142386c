// mesh_kernels.cuh -- particle <-> mesh kernels of the OrderParameterMesh path (sm_100a).
//
// Reference behaviour restated (CPU path = parity target): OrderParameterMesh.cc:517-640 (assignParticles),
// :457-483 (TSC weights), :749-864 (interpolateForces).  Reference GPU kernels replaced:
// gpu_bin_particles_kernel / gpu_assign_binned_particles_to_scratch_kernel / gpu_reduce_scratch_kernel /
// gpu_compute_forces_kernel (OrderParameterMeshGPU.cu:90-364, 566-769): non-deterministic atomicInc binning with
// an overflow-retry loop, a 27x scratch mesh, texture gathers.
//
// Here: particles are counting-sorted by a TILE-MAJOR cell key (tile of T^3 cells, T = 8 or 16), so that
//   * spreading is a thread-per-cell, atomics-free accumulation into a shared-memory tile with a one-cell halo
//     (27 conflict-free "shift" rounds), flushed as one contiguous padded tile; a merge pass sums the <= 8
//     overlapping padded tiles per cell in a fixed order (deterministic, no float atomics anywhere);
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
    unsigned nx, ny, nz;        // mesh points (powers of two)
    unsigned lgx, lgy, lgz;     // log2 of the above
    unsigned lgT;               // log2 of the tile edge T (3 or 4)
    unsigned ntx, nty, ntz;     // tiles per dimension
    float lo[3], L[3];          // single-precision box (HOOMD SINGLE_PRECISION BoxDim)
    double dlo[3], dscale[3];   // fp64: lo and n/L for the in-cell offset
};

MHD unsigned tile_edge(const Geom& g) { return 1u << g.lgT; }
MHD unsigned cells_per_tile(const Geom& g) { return 1u << (3 * g.lgT); }
MHD unsigned padded_edge(const Geom& g) { return (1u << g.lgT) + 2; }
MHD unsigned num_tiles(const Geom& g) { return g.ntx * g.nty * g.ntz; }

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
    const unsigned tx = ix >> g.lgT, ty = iy >> g.lgT, tz = iz >> g.lgT;
    const unsigned tile = (tz * g.nty + ty) * g.ntx + tx;
    const unsigned local = ((((iz & T1) << g.lgT) + (iy & T1)) << g.lgT) + (ix & T1);
    return (tile << (3 * g.lgT)) + local;
}
MHD void cell_of_key(unsigned key, const Geom& g, unsigned& ix, unsigned& iy, unsigned& iz) {
    const unsigned T1 = (1u << g.lgT) - 1;
    const unsigned local = key & ((1u << (3 * g.lgT)) - 1), tile = key >> (3 * g.lgT);
    const unsigned tx = tile % g.ntx, ty = (tile / g.ntx) % g.nty, tz = tile / (g.ntx * g.nty);
    ix = (tx << g.lgT) + (local & T1);
    iy = (ty << g.lgT) + ((local >> g.lgT) & T1);
    iz = (tz << g.lgT) + (local >> (2 * g.lgT));
}

// in-cell offset in cell units, s in [-1/2, 1/2] (OrderParameterMesh.cc:565-573: minimum-image distance to the
// cell centre through makeCoordinates/minImage/makeFraction; evaluated here directly in fp64)
MHD float cell_shift(float x, unsigned i, int axis, const Geom& g) {
    const unsigned n = axis == 0 ? g.nx : (axis == 1 ? g.ny : g.nz);
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

// accumulate one particle into the 27 per-cell partial sums; acc index = (i*3 + j)*3 + k with i the x tap
MHD void spread_accumulate(float4 p /* x,y,z,a */, unsigned ix, unsigned iy, unsigned iz, const Geom& g, float (&acc)[27]) {
    float wx[3], wy[3], wz[3];
    tsc(cell_shift(p.x, ix, 0, g), wx);
    tsc(cell_shift(p.y, iy, 1, g), wy);
    tsc(cell_shift(p.z, iz, 2, g), wz);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float ax = p.w * wx[i];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float axy = ax * wy[j];
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[(i * 3 + j) * 3 + k] += axy * wz[k];
        }
    }
}

// index inside the padded tile (edge P = T+2) of tap (i,j,k) in {0,1,2}^3 of local cell (lx,ly,lz)
MHD unsigned padded_index(unsigned lx, unsigned ly, unsigned lz, int i, int j, int k, unsigned P) {
    return ((lz + k) * P + (ly + j)) * P + (lx + i);
}

// merge: value of mesh cell (x,y,z) = sum over the <= 8 padded tiles that cover it, in fixed (z,y,x) order
MHD float merge_cell(const float* __restrict__ scratch, unsigned x, unsigned y, unsigned z, const Geom& g) {
    const unsigned T = 1u << g.lgT, P = T + 2, P3 = P * P * P;
    const unsigned tx = x >> g.lgT, ty = y >> g.lgT, tz = z >> g.lgT;
    const unsigned lx = x & (T - 1), ly = y & (T - 1), lz = z & (T - 1);
    // candidates per axis: (tile, padded coordinate)
    unsigned ctx[2], cpx[2], cty[2], cpy[2], ctz[2], cpz[2];
    int nxc = 1, nyc = 1, nzc = 1;
    ctx[0] = tx; cpx[0] = lx + 1;
    if (lx == 0) { ctx[1] = (tx + g.ntx - 1) % g.ntx; cpx[1] = T + 1; nxc = 2; }
    else if (lx == T - 1) { ctx[1] = (tx + 1) % g.ntx; cpx[1] = 0; nxc = 2; }
    cty[0] = ty; cpy[0] = ly + 1;
    if (ly == 0) { cty[1] = (ty + g.nty - 1) % g.nty; cpy[1] = T + 1; nyc = 2; }
    else if (ly == T - 1) { cty[1] = (ty + 1) % g.nty; cpy[1] = 0; nyc = 2; }
    ctz[0] = tz; cpz[0] = lz + 1;
    if (lz == 0) { ctz[1] = (tz + g.ntz - 1) % g.ntz; cpz[1] = T + 1; nzc = 2; }
    else if (lz == T - 1) { ctz[1] = (tz + 1) % g.ntz; cpz[1] = 0; nzc = 2; }
    float sum = 0.f;
    for (int c = 0; c < nzc; ++c)
        for (int b = 0; b < nyc; ++b)
            for (int a = 0; a < nxc; ++a) {
                const unsigned tile = (ctz[c] * g.nty + cty[b]) * g.ntx + ctx[a];
                sum += scratch[(size_t)tile * P3 + (cpz[c] * P + cpy[b]) * P + cpx[a]];
            }
    return sum;
}

// force on one particle from the padded tile of Re(IFFT(G)) (interpolateForces, OrderParameterMesh.cc:812-860):
//   F = -(a) * sum_taps inv * [ nb1 W'x Wy Wz + nb2 Wx W'y Wz + nb3 Wx Wy W'z ],  nb_a = n_a * b_a (no 2 pi)
// evaluated as three separable contractions; returns the three scalar sums (Sx,Sy,Sz).
MHD void gather_sums(const float* tile, unsigned lx, unsigned ly, unsigned lz, unsigned P, const float (&wx)[3],
                     const float (&wy)[3], const float (&wz)[3], const float (&dx)[3], const float (&dy)[3],
                     const float (&dz)[3], float& Sx, float& Sy, float& Sz) {
    Sx = 0.f; Sy = 0.f; Sz = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float tw = 0.f, td = 0.f, tz = 0.f;   // sum_j {Wy, W'y, Wy} * sum_k {Wz, Wz, W'z} inv
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float u = 0.f, v = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float val = tile[padded_index(lx, ly, lz, i, j, k, P)];
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

MHD float4 gather_force(float4 p /* x,y,z,a */, unsigned ix, unsigned iy, unsigned iz, const float* tile, const Geom& g,
                        const ForceParams& fp, double bias) {
    const unsigned T1 = (1u << g.lgT) - 1, P = (1u << g.lgT) + 2;
    const float sx = cell_shift(p.x, ix, 0, g), sy = cell_shift(p.y, iy, 1, g), sz = cell_shift(p.z, iz, 2, g);
    float wx[3], wy[3], wz[3], dx[3], dy[3], dz[3];
    tsc(sx, wx); tsc(sy, wy); tsc(sz, wz);
    tsc_deriv(sx, dx); tsc_deriv(sy, dy); tsc_deriv(sz, dz);
    float Sx, Sy, Sz;
    gather_sums(tile, ix & T1, iy & T1, iz & T1, P, wx, wy, wz, dx, dy, dz, Sx, Sy, Sz);
    const float fx = -p.w * (fp.nb1[0] * Sx + fp.nb2[0] * Sy + fp.nb3[0] * Sz);
    const float fy = -p.w * (fp.nb1[1] * Sx + fp.nb2[1] * Sy + fp.nb3[1] * Sz);
    const float fz = -p.w * (fp.nb1[2] * Sx + fp.nb2[2] * Sy + fp.nb3[2] * Sz);
    const double sc = fp.two_over_n * bias;
    return make_float4((float)((double)fx * sc), (float)((double)fy * sc), (float)((double)fz * sc), 0.f);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
constexpr int kBinThreads = 256;
constexpr int kTileThreads = 512;

// bin: key + rank (slot inside the cell) per particle, per-cell counts, sum a^2 and sum a
__global__ void __launch_bounds__(kBinThreads)
mesh_bin_kernel(const float4* __restrict__ postype, unsigned N, Geom g, const float* __restrict__ mode,
                unsigned* __restrict__ keys, unsigned* __restrict__ ranks, unsigned* __restrict__ count,
                double* __restrict__ sums /* [0] sum a^2, [1] sum a */) {
    double sq = 0.0, s1 = 0.0;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const float4 p = ld_stream(postype + i);
        const unsigned ix = cell_coord(p.x, g.lo[0], g.L[0], g.nx);
        const unsigned iy = cell_coord(p.y, g.lo[1], g.L[1], g.ny);
        const unsigned iz = cell_coord(p.z, g.lo[2], g.L[2], g.nz);
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

// ---- exclusive scan of count[0..n) -> start[0..n]; n is a multiple of 4096.  Three launches: per-block sums,
// scan of the block sums (single block), apply.  The apply pass also clears count[] for the next step.
constexpr int kScanThreads = 1024;
constexpr int kScanBlockItems = 4 * kScanThreads;

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

__global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(const uint4* __restrict__ count4, unsigned* __restrict__ block_sums) {
    __shared__ unsigned sm[33];
    const uint4 v = count4[(size_t)blockIdx.x * kScanThreads + threadIdx.x];
    unsigned total;
    block_excl_scan(v.x + v.y + v.z + v.w, sm, total);
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

__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(uint4* __restrict__ count4, const unsigned* __restrict__ block_sums, unsigned* __restrict__ start,
                  unsigned n) {
    __shared__ unsigned sm[33];
    const size_t i4 = (size_t)blockIdx.x * kScanThreads + threadIdx.x;
    const uint4 v = count4[i4];
    count4[i4] = make_uint4(0u, 0u, 0u, 0u);
    unsigned total;
    const unsigned excl = block_excl_scan(v.x + v.y + v.z + v.w, sm, total) + block_sums[blockIdx.x];
    uint4 o;
    o.x = excl; o.y = o.x + v.x; o.z = o.y + v.y; o.w = o.z + v.z;
    reinterpret_cast<uint4*>(start)[i4] = o;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) start[n] = o.w + v.w;
}

// reorder: sorted[start[key] + rank] = {x, y, z, a(type)}, perm[...] = original index
__global__ void __launch_bounds__(kBinThreads)
mesh_reorder_kernel(const float4* __restrict__ postype, unsigned N, const float* __restrict__ mode,
                    const unsigned* __restrict__ keys, const unsigned* __restrict__ ranks,
                    const unsigned* __restrict__ start, float4* __restrict__ sorted, unsigned* __restrict__ perm) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        float4 p = ld_stream(postype + i);
        p.w = __ldg(mode + __float_as_int(p.w));
        const unsigned dst = __ldg(start + keys[i]) + ranks[i];
        sorted[dst] = p;
        perm[dst] = i;
    }
}

// spread: one CTA per tile, one thread per cell (batches of kTileThreads cells), atomics-free.
// Particles of a cell are visited in ascending original index so the fp32 sum order is reproducible.
template <int LGT>
__global__ void __launch_bounds__(kTileThreads)
mesh_spread_kernel(const float4* __restrict__ sorted, const unsigned* __restrict__ perm,
                   const unsigned* __restrict__ start, Geom g, float* __restrict__ scratch) {
    constexpr unsigned T = 1u << LGT, P = T + 2, P3 = P * P * P, NC = T * T * T;
    __shared__ float tile[P3];
    for (unsigned i = threadIdx.x; i < P3; i += kTileThreads) tile[i] = 0.f;
    __syncthreads();
    const unsigned tile_id = blockIdx.x;
    for (unsigned lc0 = 0; lc0 < NC; lc0 += kTileThreads) {
        const unsigned lc = lc0 + threadIdx.x;          // NC is a multiple of kTileThreads
        const unsigned key = (tile_id << (3 * LGT)) + lc;
        unsigned ix, iy, iz;
        cell_of_key(key, g, ix, iy, iz);
        const unsigned lx = lc & (T - 1), ly = (lc >> LGT) & (T - 1), lz = lc >> (2 * LGT);
        const unsigned s = __ldg(start + key), e = __ldg(start + key + 1);
        float acc[27];
#pragma unroll
        for (int r = 0; r < 27; ++r) acc[r] = 0.f;
        if (e - s == 1) {
            spread_accumulate(sorted[s], ix, iy, iz, g, acc);
        } else if (e > s) {
            // selection by ascending original index (cells hold few particles)
            long long last = -1;
            for (unsigned it = s; it < e; ++it) {
                unsigned best = 0xffffffffu, bj = s;
                for (unsigned j = s; j < e; ++j) {
                    const unsigned pj = __ldg(perm + j);
                    if ((long long)pj > last && pj < best) { best = pj; bj = j; }
                }
                last = (long long)best;
                spread_accumulate(sorted[bj], ix, iy, iz, g, acc);
            }
        }
        const bool any = e > s;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    // in one round every thread targets a distinct address (its cell shifted by the same offset)
                    if (any) tile[padded_index(lx, ly, lz, i, j, k, P)] += acc[(i * 3 + j) * 3 + k];
                    __syncthreads();
                }
    }
    float* out = scratch + (size_t)tile_id * P3;
    for (unsigned i = threadIdx.x; i < P3; i += kTileThreads) out[i] = tile[i];
}

// merge: mesh[x + nx (y + ny z)] = sum of covering padded tiles - mean (DC removal, see mesh.cu)
__global__ void __launch_bounds__(256)
mesh_merge_kernel(const float* __restrict__ scratch, Geom g, const double* __restrict__ sums, float* __restrict__ rho,
                  float* __restrict__ rho_keep) {
    const size_t M = (size_t)g.nx * g.ny * g.nz;
    const float mean = (float)(sums[1] / (double)M);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < M; c += stride) {
        const unsigned x = (unsigned)(c & (g.nx - 1)), y = (unsigned)((c >> g.lgx) & (g.ny - 1)), z = (unsigned)(c >> (g.lgx + g.lgy));
        const float v = merge_cell(scratch, x, y, z, g);
        if (rho_keep) rho_keep[c] = v;
        rho[c] = v - mean;
    }
}

// gather: one CTA per tile; shared tile of Re(IFFT(G)) with halo; one thread per particle of the tile
template <int LGT>
__global__ void __launch_bounds__(kTileThreads)
mesh_gather_kernel(const float4* __restrict__ sorted, const unsigned* __restrict__ perm,
                   const unsigned* __restrict__ start, Geom g, const float* __restrict__ inv, ForceParams fp,
                   const double* __restrict__ d_bias, float4* __restrict__ force) {
    constexpr unsigned T = 1u << LGT, P = T + 2, P3 = P * P * P;
    __shared__ float tile[P3];
    const unsigned tile_id = blockIdx.x;
    const unsigned tx = tile_id % g.ntx, ty = (tile_id / g.ntx) % g.nty, tz = tile_id / (g.ntx * g.nty);
    for (unsigned i = threadIdx.x; i < P3; i += kTileThreads) {
        const unsigned px = i % P, py = (i / P) % P, pz = i / (P * P);
        const unsigned x = ((tx << LGT) + px + g.nx - 1) & (g.nx - 1);
        const unsigned y = ((ty << LGT) + py + g.ny - 1) & (g.ny - 1);
        const unsigned z = ((tz << LGT) + pz + g.nz - 1) & (g.nz - 1);
        tile[i] = __ldg(inv + (size_t)x + (size_t)g.nx * (y + (size_t)g.ny * z));
    }
    __syncthreads();
    const double bias = *d_bias;
    const unsigned s = __ldg(start + (tile_id << (3 * LGT))), e = __ldg(start + ((tile_id + 1) << (3 * LGT)));
    for (unsigned j = s + threadIdx.x; j < e; j += kTileThreads) {
        const float4 p = sorted[j];
        const unsigned ix = cell_coord(p.x, g.lo[0], g.L[0], g.nx);
        const unsigned iy = cell_coord(p.y, g.lo[1], g.L[1], g.ny);
        const unsigned iz = cell_coord(p.z, g.lo[2], g.L[2], g.nz);
        force[__ldg(perm + j)] = gather_force(p, ix, iy, iz, tile, g, fp, bias);
    }
}
#endif  // __CUDACC__

}  // namespace mesh
}  // namespace metad
