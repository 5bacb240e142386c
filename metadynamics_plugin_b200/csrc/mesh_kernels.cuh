// mesh_kernels.cuh -- particle <-> mesh kernels of the OrderParameterMesh path (sm_100a).
//
// Reference behaviour restated (CPU path = parity target): OrderParameterMesh.cc:517-640 (assignParticles),
// :457-483 (TSC weights), :749-864 (interpolateForces).  Reference GPU kernels replaced:
// gpu_bin_particles_kernel / gpu_assign_binned_particles_to_scratch_kernel / gpu_reduce_scratch_kernel /
// gpu_compute_forces_kernel (OrderParameterMeshGPU.cu:90-364, 566-769): per-step atomicInc binning with an
// overflow-retry loop, a 27x float scratch mesh, texture gathers.
//
// Design (measured on B200, bench_micro/spread_bench.cu + profiles/):
//   * The density is accumulated in 32-bit FIXED POINT.  Integer addition is associative, so the mesh is bitwise
//     independent of the order in which particles are processed (any particle order, any tile assignment, any
//     number of ranks) -- and sm_100a has a native shared-memory integer atomic (ATOMS.ADD), which costs nothing
//     next to the per-particle arithmetic, whereas float atomics are CAS loops (3x slower for the whole kernel).
//   * Particles are visited through a TILE ORDER: a permutation that lists the particles tile by tile (T^3 cells,
//     T = 8 or 16).  The order is rebuilt only every few calls (counting sort, amortised): a CTA owns one tile plus a
//     halo of H = 2 cells in shared memory, and every particle's cell is recomputed from its CURRENT position each
//     call, so a stale order is still exact -- a particle that drifted out of its padded tile falls back to global
//     atomics and is counted (the count triggers the next rebuild).
//   * The padded tile is flushed with red.global.add.s32 into the integer mesh; the x FFT pass converts it to float,
//     removes the mean density and clears it for the next call.
//   * Force interpolation: one CTA per tile, shared-memory tile of Re IFFT(G) with halo, one thread per particle.
// Cell indices are computed with non-contracted IEEE fp32 operations and are bit-exact against the reference's
// single-precision arithmetic; in-cell offsets are evaluated in compensated fp32 (error < 1e-7 cell) against the
// double-precision box.
#pragma once
#include "common.cuh"
#include "mesh_fft_kernels.cuh"
#ifdef __CUDACC__
#include <cuda.h>            // CUtensorMap
#include <cuda_pipeline.h>
#include <cuda/ptx>
#endif

#ifndef MHD
#define MHD __host__ __device__ __forceinline__
#endif

namespace metad {
namespace mesh {

// halo cells of a padded tile: 1 for the TSC stencil + drift tolerance.  The x halo is 4 cells so that every row of the
// padded tile starts on a 16-byte boundary of the mesh: rows then move with one bulk asynchronous copy each (TMA engine:
// cp.reduce.async.bulk for the spread flush, cp.async.bulk for the gather tile).
constexpr int kHalo = 2;
constexpr int kHaloX = 4;

struct Geom {
    unsigned nx, ny, nz;        // mesh points of the LOCAL mesh (powers of two on the tiled path; any size on the general path, which
                                // uses only nx, ny, nz, nzg and the box members); nz = planes of this z slab
    unsigned lgx, lgy, lgz;     // log2 of the above
    unsigned nzg, z0;           // global number of z planes and first global plane of this slab (nzg = nz, z0 = 0 if unsharded)
    unsigned slab;              // 1: z is not periodic locally, the planes z0-1 and z0+nz belong to the neighbour ranks
    unsigned lgT;               // log2 of the tile edge T (3 or 4)
    unsigned ntx, nty, ntz;     // tiles per dimension (powers of two)
    unsigned lgtx, lgty, lgtz;  // log2 of the above
    float lo[3], L[3];          // single-precision box (HOOMD SINGLE_PRECISION BoxDim)
    double dlo[3], dscale[3];   // fp64: lo and n/L of the double-precision box (reference form of the in-cell offset)
    // the same two numbers as unevaluated float pairs: -lo = hl_hi + hl_lo, n/L = c_hi + c_lo
    float hl_hi[3], hl_lo[3], c_hi[3], c_lo[3];
    float fn[3];                // (float) global mesh dimensions
    float rcpL[3];              // 1/L in single precision (HOOMD BoxDim::m_Linv)
    // triclinic box (HOOMD tilt factors xy, xz, yz; BoxDim::makeFraction shears x and y before the division by L):
    //   x' = x - (xz - yz xy) z - xy y,   y' = y - yz z,   z' = z
    unsigned tri;               // 1: at least one tilt factor is non-zero
    float t_xy, t_a, t_yz;      // single precision, rounded like a SINGLE_PRECISION BoxDim: xy, fl(xz - fl(yz xy)), yz
    double d_xy, d_a, d_yz;     // fp64: xy, xz - yz xy, yz
    // Reference behaviour in a triclinic box, restated literally: the in-cell offset is taken as
    // makeFraction(shift_cart + lo) (OrderParameterMesh.cc:571-573, 806-808), and makeFraction applies its shear to the
    // whole argument, lo included, so every offset carries the constant  -shear(lo)/L * n:
    //   x: n_x ((xz - yz xy) Lz + xy Ly) / (2 Lx),   y: n_y yz Lz / (2 Ly),   z: 0     (in cells)
    // The TSC weights are then evaluated at (offset + tq - tap) with their general, compactly supported form, exactly as the
    // reference does (weight is lost where |.| > 3/2).  tq = 0 gives the geometrically correct assignment (knob 16).
    float tq[3];
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
// box: L[] in double (the GLOBAL box), tilt[] = xy, xz, yz (may be null); n[] = global mesh dimensions
inline void geom_set_box(Geom& g, const double* Ld, const double* tilt = nullptr, bool literal_offset = true) {
    {
        const double xy = tilt ? tilt[0] : 0.0, xz = tilt ? tilt[1] : 0.0, yz = tilt ? tilt[2] : 0.0;
        g.tri = (xy != 0.0 || xz != 0.0 || yz != 0.0) ? 1u : 0u;
        g.d_xy = xy; g.d_yz = yz; g.d_a = xz - yz * xy;
        g.t_xy = (float)xy; g.t_yz = (float)yz;
        volatile float fxz = (float)xz, prod = g.t_yz * g.t_xy;      // two roundings, as in (m_xz - m_yz*m_xy) without contraction
        volatile float ta = fxz - prod;
        g.t_a = ta;
        g.tq[0] = g.tq[1] = g.tq[2] = 0.f;
        if (g.tri && literal_offset) {
            g.tq[0] = (float)((double)g.nx * (g.d_a * Ld[2] + xy * Ld[1]) / (2.0 * Ld[0]));
            g.tq[1] = (float)((double)g.ny * yz * Ld[2] / (2.0 * Ld[1]));
        }
    }
    const unsigned n[3] = {g.nx, g.ny, g.nzg};
    for (int i = 0; i < 3; ++i) {
        g.L[i] = (float)Ld[i];
        g.lo[i] = -(g.L[i] / 2.0f);
        g.dlo[i] = -Ld[i] / 2.0;
        g.dscale[i] = (double)n[i] / Ld[i];
        const double hl = Ld[i] / 2.0;
        g.hl_hi[i] = (float)hl; g.hl_lo[i] = (float)(hl - (double)g.hl_hi[i]);
        g.c_hi[i] = (float)g.dscale[i]; g.c_lo[i] = (float)(g.dscale[i] - (double)g.c_hi[i]);
        g.fn[i] = (float)n[i];
    }
    // HOOMD's BoxDim stores m_Linv = Scalar(1.0)/(hi - lo), a single-precision IEEE division
    for (int i = 0; i < 3; ++i) {
        volatile float one = 1.0f, Lf = g.L[i];
        g.rcpL[i] = one / Lf;
    }
}

MHD unsigned tile_edge(const Geom& g) { return 1u << g.lgT; }
MHD unsigned cells_per_tile(const Geom& g) { return 1u << (3 * g.lgT); }
MHD unsigned padded_cells(const Geom& g) { const unsigned T = 1u << g.lgT; return (T + 2 * kHaloX) * (T + 2 * kHalo) * (T + 2 * kHalo); }
MHD unsigned num_tiles(const Geom& g) { return 1u << (g.lgtx + g.lgty + g.lgtz); }
MHD unsigned tile_index(unsigned tx, unsigned ty, unsigned tz, const Geom& g) { return (((tz << g.lgty) + ty) << g.lgtx) + tx; }
MHD void tile_coords(unsigned tile, const Geom& g, unsigned& tx, unsigned& ty, unsigned& tz) {
    tx = tile & (g.ntx - 1);
    ty = (tile >> g.lgtx) & (g.nty - 1);
    tz = tile >> (g.lgtx + g.lgty);
}

MHD int f2i_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int i; memcpy(&i, &f, 4); return i;
#endif
}

// non-contracted IEEE single-precision helpers (host: plain ops, the emulation is built without FMA contraction)
MHD float f_sub(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
MHD float f_add(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
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
MHD float f_fma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

#ifdef __CUDACC__
// Packed fp32 pairs (sm_100: fma/mul.rn.f32x2 -> FFMA2 / FMUL2, two IEEE operations per issue slot).  The spread and the
// gather are bound by instruction issue, not by the FMA pipe, so pairing halves the cost of their tap arithmetic.  Every
// lane performs the same correctly rounded operation as the scalar code, so results are bit-identical.  A pair built from
// the same scalar twice costs nothing: the instruction takes a scalar register as a broadcast operand.
struct F2 { unsigned long long r; };
__device__ __forceinline__ F2 f2_pack(float lo, float hi) { F2 o; asm("mov.b64 %0, {%1, %2};" : "=l"(o.r) : "f"(lo), "f"(hi)); return o; }
__device__ __forceinline__ F2 f2_dup(float v) { return f2_pack(v, v); }
__device__ __forceinline__ void f2_unpack(F2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v.r)); }
__device__ __forceinline__ float f2_lo(F2 v) { float lo, hi; f2_unpack(v, lo, hi); return lo; }
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) { F2 o; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(o.r) : "l"(a.r), "l"(b.r), "l"(c.r)); return o; }
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) { F2 o; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(o.r) : "l"(a.r), "l"(b.r)); return o; }
#endif

// cell coordinate along one axis: OrderParameterMesh.cc:543-561 with HOOMD's BoxDim::makeFraction, which multiplies by
// the stored reciprocal m_Linv = Scalar(1)/(hi - lo) (it does NOT divide; the two round differently and the cell index
// of a particle next to a cell face depends on it -- confirmed against the reference's own assignParticles compiled
// from its sources, tests/test_reference_build.py):
//   f = (x - lo) * Linv ; r = f*n ; i = (int) r (truncation) ; i == n -> 0
// Reference form (C truncation).  Out-of-box input (which HOOMD never hands over) is folded into the mesh instead of
// indexing out of range.
MHD int cell_coord_ref(float x, float lo, float Linv, unsigned n) {
    const float f = f_mul(f_sub(x, lo), Linv);
    const float r = f_mul(f, (float)n);
    int i = (r >= 0.f) ? (r < 4194303.f ? (int)r : 4194303) : 0;
    if (i >= (int)n) i = 0;
    return i;
}
// Hot form: same value for every input, no conversion-pipe instruction.
// `raw` receives the index before the upper-edge wrap (n for a particle on the upper face): the in-cell offset is
// measured from that cell, which is the periodic image of cell 0 the particle actually sits in.
// delta = the component of BoxDim::makeFraction's `delta` before the multiplication by 1/L (x - lo for an orthorhombic box)
MHD int cell_coord_delta(float delta, int axis, const Geom& g, int& raw) {
    const float f = f_mul(delta, g.rcpL[axis]);
    float r = f_mul(f, g.fn[axis]);
    const int n = (int)(axis == 0 ? g.nx : (axis == 1 ? g.ny : g.nzg));
#ifdef __CUDA_ARCH__
    // truncation: for 0 <= r < 2^22, RZ(r + 2^23) = 2^23 + trunc(r) exactly.  Out-of-box input (which HOOMD never hands
    // over) is clamped: below the box (and NaN) -> cell 0, beyond it -> wrapped to 0 below
    r = fminf(fmaxf(r, 0.f), 4194303.f);
    int i = __float_as_int(__fadd_rz(r, 8388608.f)) & 0x7fffff;
#else
    int i = (r >= 0.f) ? (r < 4194303.f ? (int)r : 4194303) : 0;
#endif
    raw = i <= n ? i : 0;
    if (i >= n) i = 0;
    return i;
}
MHD int cell_coord(float x, int axis, const Geom& g, int& raw) { return cell_coord_delta(f_sub(x, g.lo[axis]), axis, g, raw); }
MHD int cell_coord(float x, int axis, const Geom& g) { int raw; return cell_coord(x, axis, g, raw); }

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
// cell centre through makeCoordinates/minImage/makeFraction).  Reference form in fp64:
MHD float cell_shift_f64(float x, unsigned i, int axis, const Geom& g) {
    // i is the GLOBAL cell coordinate (for z: z0 + local plane)
    const unsigned n = axis == 0 ? g.nx : (axis == 1 ? g.ny : g.nzg);
    double s = ((double)x - g.dlo[axis]) * g.dscale[axis] - ((double)i + 0.5);
    const double half = 0.5 * (double)n;
    if (s > half) s -= (double)n;
    else if (s < -half) s += (double)n;
    return (float)s;
}
// The same quantity in compensated single precision (no conversion-pipe or fp64 instructions): with
// -lo = hl_hi + hl_lo and n/L = c_hi + c_lo,  x - lo = d + e exactly (Fast2Sum, |x| <= L/2), (d + e)(c_hi + c_lo)
// is evaluated as fma(d, c_hi, -(i + 1/2)) -- one rounding of a number of magnitude <= 1/2 -- plus the two small terms.
// |cell_shift - cell_shift_f64| < 1e-7 for in-box particles (tests/cpu_emul/mesh_emul.cu checks it).
MHD float cell_shift(float x, unsigned i /* cell index BEFORE the upper-edge wrap (cell_coord's raw) */, int axis, const Geom& g) {
    const float hl = g.hl_hi[axis], ch = g.c_hi[axis];
    const float d = f_add(hl, x);
    const float e = f_add(f_sub(x, f_sub(d, hl)), g.hl_lo[axis]);
    // (float)i + 0.5 without the conversion pipe: i < 2^22
#ifdef __CUDA_ARCH__
    const float ci = __int_as_float(0x4B000000 | (int)i) - 8388607.5f;
#else
    const float ci = (float)i + 0.5f;
#endif
    // d*ch - (i + 1/2) in ONE rounding (the product is exact inside the FMA; |result| <= 1/2), plus the two small terms
    return f_add(f_fma(d, ch, -ci), f_fma(e, ch, f_mul(d, g.c_lo[axis])));
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
// The same weights for an arbitrary offset (triclinic boxes: the reference's offsets carry a constant, Geom::tq), i.e.
// assignTSC / assignTSCderiv as written: W(x) = 3/4 - x^2 (|x| <= 1/2), (3/2 - |x|)^2 / 2 (|x| <= 3/2), 0;
// W'(x) = -2x, -(3/2 - |x|) sign(x), 0.  Taps i = -1, 0, +1 sit at x = s + 1, s, s - 1.
MHD float tsc_w(float x) {
    const float ax = fabsf(x), r = 1.5f - ax;
    return ax <= 0.5f ? 0.75f - x * x : (ax <= 1.5f ? 0.5f * r * r : 0.f);
}
MHD float tsc_wd(float x) {
    const float ax = fabsf(x), r = 1.5f - ax;
    return ax <= 0.5f ? -2.0f * x : (ax <= 1.5f ? (x < 0.f ? r : -r) : 0.f);
}
template <bool GENERAL> MHD void tsc_any(float s, float (&w)[3]) {
    if (GENERAL) { w[0] = tsc_w(s + 1.0f); w[1] = tsc_w(s); w[2] = tsc_w(s - 1.0f); }
    else tsc(s, w);
}
template <bool GENERAL> MHD void tsc_deriv_any(float s, float (&w)[3]) {
    if (GENERAL) { w[0] = tsc_wd(s + 1.0f); w[1] = tsc_wd(s); w[2] = tsc_wd(s - 1.0f); }
    else tsc_deriv(s, w);
}

// ---------------------------------------------------------------------------------------------------
// fixed point
// ---------------------------------------------------------------------------------------------------
// round(a*b) as an integer without the conversion pipe: fma(a, b, 1.5*2^23) has ulp 1, its mantissa bits are
// 0x4B400000 + rint(a*b) for |a*b| < 2^22 (single rounding, round to nearest even).
constexpr unsigned kWideMaxCount = (1u << 20) / 27u;    // particles per cell the wide (64-bit) spread can count, see kSpWide
constexpr float kFxMagic = 12582912.0f;
constexpr int kFxMagicBits = 0x4B400000;
MHD int fx_round(float a, float b) { return f2i_bits(f_fma(a, b, kFxMagic)) - kFxMagicBits; }

// Scale of the fixed-point density: a power of two such that (i) one tap, |a| W^3 <= 0.421875 |a|max, stays below
// 2^22 (fx_round) and (ii) the total of a cell stays below 2^31 / 4: a cell collects at most
// sum_offsets Wmax(offset) = (1/2 + 3/4 + 1/2)^3 = 5.36 times the largest per-cell load of its 27 neighbours.
// max_cell_load = max over cells of sum |a| at the last rebuild of the tile order (the factor 4 is the headroom for
// density changes until the next rebuild; the flush raises a flag once any cell passes 2^30).
MHD float fx_scale_for(float amax, float max_cell_load) {
    if (!(amax > 0.f)) return 1.0f;
    const float tap = 4194304.0f / (0.421875f * amax);
    const float load = max_cell_load > amax ? max_cell_load : amax;
    const float tot = 2147483648.0f / (4.0f * 5.359375f * load);
    const float lim = tap < tot ? tap : tot;
    // largest power of two <= lim
    int b = f2i_bits(lim);
    b &= 0x7f800000;
    float p;
#ifdef __CUDA_ARCH__
    p = __int_as_float(b);
#else
    memcpy(&p, &b, 4);
#endif
    return p;
}

// ---------------------------------------------------------------------------------------------------
// per-particle bodies shared by the kernels and the CPU emulation (tests/cpu_emul/mesh_emul.cu)
// ---------------------------------------------------------------------------------------------------
struct Cell {
    int ix, iy, iz;      // global cell (bit-exact reference rule); iz is the GLOBAL plane
    int rx, ry, rz;      // the same before the upper-edge wrap (n instead of 0 for a particle on an upper face)
    bool owned;          // slab mode: the plane belongs to this rank
};
MHD Cell particle_cell(float4 p, const Geom& g) {
    Cell c;
    if (g.tri) {
        // BoxDim::makeFraction of a SINGLE_PRECISION build, operation by operation (no contraction):
        //   delta = v - lo;  delta.x -= (xz - yz*xy)*v.z + xy*v.y;  delta.y -= yz*v.z;
        const float dx = f_sub(f_sub(p.x, g.lo[0]), f_add(f_mul(g.t_a, p.z), f_mul(g.t_xy, p.y)));
        const float dy = f_sub(f_sub(p.y, g.lo[1]), f_mul(g.t_yz, p.z));
        c.ix = cell_coord_delta(dx, 0, g, c.rx);
        c.iy = cell_coord_delta(dy, 1, g, c.ry);
    } else {
        c.ix = cell_coord(p.x, 0, g, c.rx);
        c.iy = cell_coord(p.y, 1, g, c.ry);
    }
    c.iz = cell_coord(p.z, 2, g, c.rz);
    c.owned = (unsigned)(c.iz - (int)g.z0) < g.nz;
    return c;
}
// coordinates of the cell inside the padded tile with origin (ox,oy,oz) = tile origin - halo (local planes in z);
// returns true if all 27 taps lie inside the padded tile
MHD bool padded_coords(const Cell& c, int ox, int oy, int oz, const Geom& g, int PX, int PY, int PZ, unsigned& lx, unsigned& ly,
                       unsigned& lz) {
    lx = (unsigned)(c.ix - ox) & (g.nx - 1);
    ly = (unsigned)(c.iy - oy) & (g.ny - 1);
    const int zl = c.iz - (int)g.z0;
    lz = g.slab ? (unsigned)(zl - oz) : ((unsigned)(zl - oz) & (g.nz - 1));
    return (lx - 1u) < (unsigned)(PX - 2) && (ly - 1u) < (unsigned)(PY - 2) && (lz - 1u) < (unsigned)(PZ - 2);
}
// Row (py, pz) of a padded tile in the mesh: plane (local; a slab also has planes -1 and nz) and row y; returns false if the
// plane does not exist.  The x range [ox, ox + PX) wraps periodically: `first` cells lie before the wrap.
struct TileRow { int z; unsigned y; unsigned x0; int first; };
MHD bool tile_row(int ox, int oy, int oz, int py, int pz, int PX, const Geom& g, TileRow& r) {
    r.y = (unsigned)(oy + py) & (g.ny - 1);
    r.z = oz + pz;
    if (g.slab) { if (r.z < -1 || r.z > (int)g.nz) return false; }
    else r.z = (int)((unsigned)r.z & (g.nz - 1));
    r.x0 = (unsigned)ox & (g.nx - 1);
    const int room = (int)g.nx - (int)r.x0;
    r.first = room < PX ? room : PX;
    return true;
}
// in-cell offsets of a particle (cell units)
MHD float3 particle_shift(float4 p, const Cell& c, const Geom& g) {
    return make_float3(cell_shift(p.x, c.rx, 0, g), cell_shift(p.y, c.ry, 1, g), cell_shift(p.z, c.rz, 2, g));
}
// Stencil base.  The cell of particle_cell() follows the reference's SINGLE-precision rule bit for bit (that is what is
// reported as "the cell of the particle"), but for a particle within ~n 2^-23 cells of a face it can differ from the
// cell that holds the particle in exact arithmetic -- the cell a double-precision build of the reference picks.  The
// offset measured from it then lies outside [-1/2, 1/2] by that much, and because the TSC derivative weights are only
// piecewise linear (W'' jumps at |x| = 1/2, OrderParameterMesh.cc:470-483) extending the inner polynomials would put a
// FIRST-order error of 3 (|s| - 1/2) into the force of that particle (measured: 1.6e-5 of max|F| for 512 cells per
// axis, 3.9e-5 for 1024 -- the parity tolerance is 1e-5).  So the 27 taps are placed around the cell the accurate offset
// points to: s > 1/2 moves the base one cell up, s < -1/2 one cell down (periodic).  A z slab keeps its own base when
// the move would leave the slab (its ghost planes hold one layer only).
MHD void rebase_axis(int& i, float& s, unsigned n) {
    const int adj = (s > 0.5f ? 1 : 0) - (s < -0.5f ? 1 : 0);
    i = (int)((unsigned)(i + adj) & (n - 1));
    s -= (float)adj;
}
MHD void particle_rebase(Cell& c, float3& s, const Geom& g) {
    rebase_axis(c.ix, s.x, g.nx);
    rebase_axis(c.iy, s.y, g.ny);
    int iz = c.iz;
    float sz = s.z;
    rebase_axis(iz, sz, g.nzg);
    if (!g.slab || (unsigned)(iz - (int)g.z0) < g.nz) { c.iz = iz; s.z = sz; }
}
// Hot form of "cell + offset + re-base" in one go (what the spread and the gather run per particle): the cell that holds
// the particle in (nearly) exact arithmetic and the offset from its centre, |s| <= 1/2.  With x - lo = d + e exactly
// (Fast2Sum, as in cell_shift) and n/L = c_hi + c_lo, the cell coordinate is r = rh + rl + e c_hi + d c_lo where
// rh = fl(d c_hi) and rl = fma(d, c_hi, -rh) is its rounding error, exactly.  trunc(rh) comes from RZ(rh + 2^23) (no
// conversion-pipe instruction), rh - trunc(rh) and the subtraction of 1/2 are exact (or rounded once below 1/4), the small
// terms are added last: |s - exact| < 1e-7.  Because rh is itself rounded, trunc(rh) can be one cell off at a face; the
// offset then leaves [-1/2, 1/2] and the final rint() step moves cell and offset back together (the re-base above).
// Agrees with particle_cell + particle_shift + particle_rebase (tests/cpu_emul/mesh_emul.cu compares the two on every
// particle); 20 instead of 34 instructions per axis.
// LOW: the coordinate is the unevaluated sum x + xl (a sheared coordinate of a triclinic box, |xl| <= ulp(x)/2).
// POW2 = false: n is any positive number (general mesh path, mesh_general.cuh): the periodic wrap is a remainder.
template <bool LOW = false, bool POW2 = true>
MHD void axis_stencil(float x, int axis, const Geom& g, unsigned n, int& i, float& s, float xl = 0.f) {
    const float hl = g.hl_hi[axis], ch = g.c_hi[axis];
    const float d = f_add(hl, x);
    float e = f_add(f_sub(x, f_sub(d, hl)), g.hl_lo[axis]);
    if (LOW) e = f_add(e, xl);
    const float rh = f_mul(d, ch);
    const float rl = f_fma(d, ch, -rh);
    const float small = f_add(rl, f_fma(e, ch, f_mul(d, g.c_lo[axis])));
#ifdef __CUDA_ARCH__
    const float u = __fadd_rz(fmaxf(rh, 0.f), 8388608.f);                 // 2^23 + trunc(rh)
    const float t = __fsub_rn(u, 8388608.f);
    const float sv = __fadd_rn(__fadd_rn(__fsub_rn(rh, t), -0.5f), small);
    const float ur = __fadd_rn(sv, kFxMagic);                             // 1.5 2^23 + rint(sv)
    const int iv = __float_as_int(u) + __float_as_int(ur) - (0x4B000000 + kFxMagicBits);
    if (POW2) i = (int)((unsigned)iv & (n - 1));
    else { const int m = iv % (int)n; i = m < 0 ? m + (int)n : m; }
    s = __fsub_rn(sv, __fsub_rn(ur, kFxMagic));
#else
    const float rc = rh > 0.f ? rh : 0.f;
    const int i0 = rc < 4194303.f ? (int)rc : 4194303;
    const float t = (float)i0;
    const float sv = f_add(f_add(f_sub(rh, t), -0.5f), small);
    const float ur = f_add(sv, kFxMagic);
    const int iv = i0 + (f2i_bits(ur) - kFxMagicBits);
    if (POW2) i = (int)((unsigned)iv & (n - 1));
    else { const int m = iv % (int)n; i = m < 0 ? m + (int)n : m; }
    s = f_sub(sv, f_sub(ur, kFxMagic));
#endif
}
// stencil base (global cell, z = global plane) and offsets of a particle; `owned` = the float-rule plane of the particle
// belongs to this rank's slab (always true for an unsharded plan).  A slab keeps the base inside its planes (see
// particle_rebase): the particle's taps must not reach beyond the one ghost layer.
// TRI: triclinic box.  The sheared coordinates x' = x - (xz - yz xy) z - xy y and y' = y - yz z are evaluated in fp64 (two
// fused multiply-adds; the products of single-precision positions with the tilt factors do not fit single precision) and
// handed on as float pairs, so the offsets keep the < 1e-7 cell accuracy of the orthorhombic path.  A separate
// instantiation: the orthorhombic kernels do not carry the fp64 instructions.
template <bool TRI = false, bool POW2 = true>
MHD void particle_stencil(float4 p, const Geom& g, Cell& c, float3& s) {
    if (TRI) {
        const double xd = (double)p.x - (g.d_a * (double)p.z + g.d_xy * (double)p.y);
        const double yd = (double)p.y - g.d_yz * (double)p.z;
        const float xh = (float)xd, yh = (float)yd;
        axis_stencil<true, POW2>(xh, 0, g, g.nx, c.ix, s.x, (float)(xd - (double)xh));
        axis_stencil<true, POW2>(yh, 1, g, g.ny, c.iy, s.y, (float)(yd - (double)yh));
        s.x += g.tq[0];          // the reference's constant (see Geom::tq); the weights take their general form from here on
        s.y += g.tq[1];
    } else {
        axis_stencil<false, POW2>(p.x, 0, g, g.nx, c.ix, s.x);
        axis_stencil<false, POW2>(p.y, 1, g, g.ny, c.iy, s.y);
    }
    axis_stencil<false, POW2>(p.z, 2, g, g.nzg, c.iz, s.z);
    c.owned = true;
    if (POW2 && g.slab) {
        const int izf = cell_coord(p.z, 2, g);
        c.owned = (unsigned)(izf - (int)g.z0) < g.nz;
        if (c.owned && (unsigned)(c.iz - (int)g.z0) >= g.nz) {
            // the float rule (which assigns particles to ranks) and the accurate cell disagree across the slab face
            const float dlt = (((unsigned)(izf - c.iz)) & (g.nzg - 1)) == 1u ? 1.0f : -1.0f;
            c.iz = izf;
            s.z -= dlt;
        }
    }
}
// separable weights: w[0..2] = Wx(tap -1,0,+1), w[3..5] = Wy, w[6..8] = amp * Wz
template <bool TRI = false>
MHD void spread_weights(float3 s, float amp, float (&w)[9]) {
    float wx[3], wy[3], wz[3];
    tsc_any<TRI>(s.x, wx);
    tsc_any<TRI>(s.y, wy);
    tsc(s.z, wz);
#pragma unroll
    for (int i = 0; i < 3; ++i) { w[i] = wx[i]; w[3 + i] = wy[i]; w[6 + i] = amp * wz[i]; }
}
// Particle cache written by the spread for the gather of the same call pair (tile order), 16 bytes per particle:
// {sx, sy, sz, code} with the code word = padded-tile coordinates of the stencil base (5 bits each) | particle type << 15
// (10 bits) | kCacheInside (all taps inside the padded tile) | kCacheOwned (the plane belongs to this rank).  The gather then
// needs neither the positions nor the cell arithmetic; the particle index comes from the tile order (4 bytes, sequential).
constexpr unsigned kCacheInside = 1u << 30, kCacheOwned = 1u << 31;
MHD unsigned cache_code(unsigned lx, unsigned ly, unsigned lz, unsigned type, bool inside, bool owned) {
    return inside ? (lx | (ly << 5) | (lz << 10) | (type << 15) | kCacheInside | kCacheOwned) : ((type << 15) | (owned ? kCacheOwned : 0u));
}
// fixed-point value of tap (i,j,k) in {0,1,2}^3: the SAME expression on every path (tile, stray, emulation)
MHD int tap_value(const float (&w)[9], int i, int j, int k) { return fx_round(w[i], f_mul(w[3 + j], w[6 + k])); }

// mesh index of tap (i,j,k) of cell c for the direct (stray) path; returns false if the plane is outside the slab + ghosts.
// `mesh` points at local plane 0; in slab mode the ghost planes are plane -1 and plane nz of the same allocation.
MHD bool tap_index(const Cell& c, int i, int j, int k, const Geom& g, long long& idx) {
    const unsigned x = (unsigned)(c.ix + i - 1) & (g.nx - 1), y = (unsigned)(c.iy + j - 1) & (g.ny - 1);
    int z = c.iz - (int)g.z0 + k - 1;
    if (g.slab) { if (z < -1 || z > (int)g.nz) return false; }
    else z = (int)((unsigned)z & (g.nz - 1));
    idx = (long long)x + (long long)g.nx * ((long long)y + (long long)g.ny * (long long)z);
    return true;
}

// force on one particle from a tile of Re(IFFT(G)) (interpolateForces, OrderParameterMesh.cc:812-860):
//   F = -(a) * sum_taps inv * [ nb1 W'x Wy Wz + nb2 Wx W'y Wz + nb3 Wx Wy W'z ],  nb_a = n_a * b_a (no 2 pi)
// evaluated as three separable contractions; returns the three scalar sums (Sx,Sy,Sz).  base = address of tap
// (0,0,0); sx/sy/sz = element strides.
MHD void gather_sums(const float* base, long long sy_, long long sz_, const float (&wx)[3],
                     const float (&wy)[3], const float (&wz)[3], const float (&dx)[3], const float (&dy)[3],
                     const float (&dz)[3], float& Sx, float& Sy, float& Sz) {
    Sx = 0.f; Sy = 0.f; Sz = 0.f;
    // plane by plane (9 values live at a time): a_j = sum_i Wx_i v_ij, b_j = sum_i W'x_i v_ij, then the y and z contractions
#ifdef __CUDA_ARCH__
    // the same operations, two per instruction: {a_j, b_j} from {Wx_i, W'x_i} * v (v broadcast), {p_k, q_k} += Wy_j * {a_j, b_j},
    // {Sz, Sx} += {W'z_k, Wz_k} * {p_k, q_k}
    const F2 wdx0 = f2_pack(wx[0], dx[0]), wdx1 = f2_pack(wx[1], dx[1]), wdx2 = f2_pack(wx[2], dx[2]);
    F2 Szx = f2_pack(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        F2 pq = f2_pack(0.f, 0.f);
        float rk = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float* row = base + k * sz_ + j * sy_;
            const float v0 = row[0], v1 = row[1], v2 = row[2];
            const F2 ab = f2_fma(wdx2, f2_dup(v2), f2_fma(wdx1, f2_dup(v1), f2_mul(wdx0, f2_dup(v0))));
            pq = f2_fma(f2_dup(wy[j]), ab, pq);
            rk = fmaf(dy[j], f2_lo(ab), rk);
        }
        Szx = f2_fma(f2_pack(dz[k], wz[k]), pq, Szx);
        Sy = fmaf(wz[k], rk, Sy);
    }
    f2_unpack(Szx, Sz, Sx);
    return;
#endif
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float pk = 0.f, qk = 0.f, rk = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float* row = base + k * sz_ + j * sy_;
            const float v0 = row[0], v1 = row[1], v2 = row[2];
            const float a = fmaf(wx[2], v2, fmaf(wx[1], v1, wx[0] * v0));
            const float b = fmaf(dx[2], v2, fmaf(dx[1], v1, dx[0] * v0));
            pk = fmaf(wy[j], a, pk);
            qk = fmaf(wy[j], b, qk);
            rk = fmaf(dy[j], a, rk);
        }
        Sx = fmaf(wz[k], qk, Sx);
        Sy = fmaf(wz[k], rk, Sy);
        Sz = fmaf(dz[k], pk, Sz);
    }
}

struct ForceParams {
    float nb1[3], nb2[3], nb3[3];   // n_a * b_a, b_a = reciprocal lattice vectors of the box without 2 pi (:761-769)
    double two_over_n;              // 2 / N_global  (:858)
};

struct GatherWeights { float wx[3], wy[3], wz[3], dx[3], dy[3], dz[3]; };
template <bool TRI = false>
MHD void gather_weights(float3 s, GatherWeights& w) {
    tsc_any<TRI>(s.x, w.wx); tsc_any<TRI>(s.y, w.wy); tsc(s.z, w.wz);
    tsc_deriv_any<TRI>(s.x, w.dx); tsc_deriv_any<TRI>(s.y, w.dy); tsc_deriv(s.z, w.dz);
}
// amp = a(type); scale = (2/N) * bias rounded to float
MHD float4 force_from_sums(float Sx, float Sy, float Sz, float amp, const ForceParams& fp, float scale) {
    const float m = -amp * scale;
    return make_float4(m * (fp.nb1[0] * Sx + fp.nb2[0] * Sy + fp.nb3[0] * Sz), m * (fp.nb1[1] * Sx + fp.nb2[1] * Sy + fp.nb3[1] * Sz),
                       m * (fp.nb1[2] * Sx + fp.nb2[2] * Sy + fp.nb3[2] * Sz), 0.f);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------
// kernels: tile order (rebuilt every few calls)
// ---------------------------------------------------------------------------------------------------
constexpr int kBinThreads = 256;

// bin: key + arrival rank inside the cell per particle, per-cell counts, largest per-cell load sum |a|
__global__ void __launch_bounds__(kBinThreads)
mesh_bin_kernel(const float4* __restrict__ postype, unsigned N, Geom g, const float* __restrict__ mode, int ntypes,
                unsigned* __restrict__ keys, unsigned* __restrict__ ranks, unsigned* __restrict__ count,
                unsigned* __restrict__ max_count) {
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned mx = 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const float4 p = ld_stream(postype + i);
        const Cell c = particle_cell(p, g);
        // a particle outside the slab is a caller error: it is listed in a tile of the slab (the spread skips and counts it)
        const unsigned iz = (unsigned)(c.iz - (int)g.z0) & (g.nz - 1);
        const unsigned key = key_of(c.ix, c.iy, iz, g);
        keys[i] = key;
        const unsigned r = atomicAdd(count + key, 1u);
        ranks[i] = r;
        mx = max(mx, r + 1);
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) atomicMax(max_count, mx);
}

// fixed-point scale of the density for the calls until the next rebuild: d_fx = {scale, 1/scale}
// wide: the accumulators are 64 bits wide, only the tap limit applies.  h_mode (pinned host word, may be null) receives
// what the cell loads of this rebuild ask for: 1 = 32-bit accumulation keeps the full tap resolution, 2 = 64-bit
// accumulation needed for that, 3 = more particles in one cell than the split 32-bit tiles of the wide spread can count
MHD unsigned fx_mode_for(float amax, unsigned max_count) {
    if (max_count > kWideMaxCount) return 3u;
    return fx_scale_for(amax, amax * (float)max_count) < fx_scale_for(amax, amax) ? 2u : 1u;
}
__global__ void mesh_fx_mode_kernel(const unsigned* __restrict__ max_count, float amax, unsigned* __restrict__ h_mode) {
    *h_mode = fx_mode_for(amax, *max_count);
}
__global__ void mesh_fx_scale_kernel(const unsigned* __restrict__ max_count, float amax, int wide, float* __restrict__ d_fx,
                                     unsigned* __restrict__ h_mode) {
    const float s = wide ? fx_scale_for(amax, amax) : fx_scale_for(amax, amax * (float)*max_count);
    d_fx[0] = s;
    d_fx[1] = 1.0f / s;
    d_fx[4] = 1.0f / s;          // 16-byte aligned copy: trailer of the halo messages (peer-memory mode)
    if (h_mode) *h_mode = fx_mode_for(amax, *max_count);
}

// ---- exclusive scan of count[0..n) -> start[0..n].  Three launches: per-block sums, scan of the block sums (single
// block), apply.  Each thread owns V consecutive uint4 (n must be a multiple of 4096*V).  The apply pass also
// clears count[] for the next rebuild.
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

// place: perm[start[key] + arrival rank] = particle index; tstart[tile] = first slot of the tile.  The order inside
// a cell is arbitrary -- results do not depend on it (integer accumulation).
__global__ void __launch_bounds__(kBinThreads)
mesh_place_kernel(unsigned N, const unsigned* __restrict__ keys, const unsigned* __restrict__ ranks,
                  const unsigned* __restrict__ start, unsigned* __restrict__ perm, unsigned ntiles, unsigned lg_cells,
                  unsigned* __restrict__ tstart) {
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (unsigned i = t0; i < N; i += stride) perm[__ldg(start + keys[i]) + ranks[i]] = i;
    for (unsigned t = t0; t <= ntiles; t += stride) tstart[t] = __ldg(start + ((size_t)t << lg_cells));
}

// layer order inside a tile: first one particle of every occupied cell (cells ascending, x fastest), then the second
// particles, and so on.  Consecutive lanes of a warp then work on DIFFERENT, mostly x-adjacent cells, so the 27
// shared-memory atomics of a warp hit distinct addresses in distinct banks (particles of one cell in neighbouring lanes
// would serialise: measured 4.05 wavefronts per ATOMS with the plain cell order).  One CTA per tile; amortised.
constexpr int kLayerThreads = 256;
template <int LGT>
__global__ void __launch_bounds__(kLayerThreads)
mesh_layer_order_kernel(const unsigned* __restrict__ start, const unsigned* __restrict__ perm, unsigned* __restrict__ order) {
    constexpr int CELLS = 1 << (3 * LGT), CPT = CELLS / kLayerThreads;      // cells per thread (16 or 2)
    __shared__ unsigned s_start[CELLS + 1];
    __shared__ unsigned s_scan[kLayerThreads / 32 + 1];
    __shared__ unsigned s_max;
    const size_t key0 = (size_t)blockIdx.x << (3 * LGT);
    for (int i = threadIdx.x; i <= CELLS; i += kLayerThreads) s_start[i] = __ldg(start + key0 + i);
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    const unsigned tile_begin = s_start[0];
    if (s_start[CELLS] == tile_begin) return;
    const int c0 = threadIdx.x * CPT;
    unsigned mx = 0;
#pragma unroll
    for (int k = 0; k < CPT; ++k) mx = max(mx, s_start[c0 + k + 1] - s_start[c0 + k]);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) atomicMax(&s_max, mx);
    __syncthreads();
    const unsigned layers = s_max;
    unsigned out = tile_begin;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (unsigned r = 0; r < layers; ++r) {
        unsigned cnt = 0;
#pragma unroll
        for (int k = 0; k < CPT; ++k) cnt += (s_start[c0 + k + 1] - s_start[c0 + k] > r) ? 1u : 0u;
        // block exclusive scan of cnt
        unsigned incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += y;
        }
        if (lane == 31) s_scan[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const unsigned w = lane < kLayerThreads / 32 ? s_scan[lane] : 0u;
            unsigned wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= (unsigned)o) wi += y;
            }
            if (lane < kLayerThreads / 32) s_scan[lane] = wi - w;
            if (lane == 31) s_scan[kLayerThreads / 32] = wi;
        }
        __syncthreads();
        unsigned pos = out + s_scan[wid] + incl - cnt;
        const unsigned total = s_scan[kLayerThreads / 32];
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const unsigned b = s_start[c0 + k];
            if (s_start[c0 + k + 1] - b > r) order[pos++] = __ldg(perm + b + r);
        }
        out += total;
        __syncthreads();
    }
}

// bank order inside a tile.  All 27 taps of a particle sit at a fixed offset from the shared-memory word of its base
// cell, so whether two lanes of a warp collide in a bank is decided by the BANK CLASS of their base cells alone
// (word index of the cell in the padded tile, modulo 32).  The order interleaves the 32 classes: slot 32 r + b holds the
// r-th particle of class b (cells ascending inside a class), so that lane b of every warp works on bank b for every tap:
// no bank conflict and no same-address serialisation, for the spread's atomics and the gather's loads alike.  Once the
// smallest classes run out, the remaining ones close ranks (at most two particles per bank and warp while more than half
// of the classes are alive).  One CTA per tile; amortised like the layer order it replaces.
template <int LGT> MHD unsigned bank_class(unsigned local_cell) {
    constexpr unsigned T = 1u << LGT, PX = T + 2 * kHaloX, PY = T + 2 * kHalo;
    const unsigned lx = local_cell & (T - 1), ly = (local_cell >> LGT) & (T - 1), lz = local_cell >> (2 * LGT);
    return (lx + PX * (ly + PY * lz)) & 31u;
}
// slot (relative to the tile) of the r-th particle of class b: all particles of lower rank, then the lower classes of
// rank r
MHD unsigned bank_order_slot(unsigned r, unsigned b, const unsigned* class_count) {
    unsigned slot = 0;
    for (unsigned c = 0; c < 32; ++c) {
        const unsigned n = class_count[c];
        slot += (n < r ? n : r) + ((c < b && n > r) ? 1u : 0u);
    }
    return slot;
}
// x coordinate (inside the tile) of the one cell of bank class `cls` in tile row (ly, lz), or >= T if the row has none
template <int LGT> MHD unsigned bank_class_cell_x(unsigned cls, unsigned row /* ly + T * lz */) {
    constexpr unsigned T = 1u << LGT, PX = T + 2 * kHaloX, PY = T + 2 * kHalo;
    const unsigned ly = row & (T - 1), lz = row >> LGT;
    return (cls - PX * (ly + PY * lz)) & 31u;
}
constexpr int kBankTable = 1024;      // ranks covered by the slot table (larger class populations use bank_order_slot directly)
template <int LGT>
__global__ void __launch_bounds__(kLayerThreads)
mesh_bank_order_kernel(const unsigned* __restrict__ start, const unsigned* __restrict__ perm, unsigned* __restrict__ order) {
    constexpr int T = 1 << LGT, CELLS = 1 << (3 * LGT), CPT = CELLS / kLayerThreads, SEGS = kLayerThreads / 32, ROWS = T * T / SEGS;
    static_assert(T <= 32, "one cell per class and tile row");
    __shared__ unsigned s_start[CELLS + 1];
    __shared__ unsigned s_base[CELLS];              // rank inside its class of the first particle of a cell
    __shared__ unsigned s_seg[SEGS][32];
    __shared__ unsigned s_count[32];
    __shared__ unsigned s_below[kBankTable];        // particles of rank < r, all classes: sum_c min(n_c, r)
    __shared__ unsigned s_alive[kBankTable];        // bit c: class c has a particle of rank r
    const size_t key0 = (size_t)blockIdx.x << (3 * LGT);
    for (int i = threadIdx.x; i <= CELLS; i += kLayerThreads) s_start[i] = __ldg(start + key0 + i);
    __syncthreads();
    const unsigned tile_begin = s_start[0];
    if (s_start[CELLS] == tile_begin) return;
    // class totals and the per-cell base ranks: thread (segment, class) walks the tile rows of its segment; a row holds
    // at most one cell of its class
    const unsigned cls = threadIdx.x & 31, seg = threadIdx.x >> 5;
    unsigned cnt = 0;
    for (unsigned row = seg * ROWS; row < (seg + 1) * ROWS; ++row) {
        const unsigned lx = bank_class_cell_x<LGT>(cls, row);
        if (lx < (unsigned)T) { const unsigned c = (row << LGT) + lx; cnt += s_start[c + 1] - s_start[c]; }
    }
    s_seg[seg][cls] = cnt;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned run = 0;
        for (int k = 0; k < SEGS; ++k) { const unsigned t = s_seg[k][threadIdx.x]; s_seg[k][threadIdx.x] = run; run += t; }
        s_count[threadIdx.x] = run;
    }
    __syncthreads();
    unsigned run = s_seg[seg][cls];
    for (unsigned row = seg * ROWS; row < (seg + 1) * ROWS; ++row) {
        const unsigned lx = bank_class_cell_x<LGT>(cls, row);
        if (lx < (unsigned)T) { const unsigned c = (row << LGT) + lx; s_base[c] = run; run += s_start[c + 1] - s_start[c]; }
    }
    unsigned rmax = 0;
    for (int c = 0; c < 32; ++c) rmax = max(rmax, s_count[c]);
    for (unsigned r = threadIdx.x; r < min(rmax, (unsigned)kBankTable); r += kLayerThreads) {
        unsigned below = 0, alive = 0;
        for (unsigned c = 0; c < 32; ++c) {
            const unsigned n = s_count[c];
            below += min(n, r);
            alive |= (n > r ? 1u : 0u) << c;
        }
        s_below[r] = below; s_alive[r] = alive;
    }
    __syncthreads();
    for (int k = 0; k < CPT; ++k) {
        const unsigned c = threadIdx.x + k * kLayerThreads;
        const unsigned b = s_start[c], n = s_start[c + 1] - b, r0 = s_base[c], cb = bank_class<LGT>(c);
        for (unsigned i = 0; i < n; ++i) {
            const unsigned r = r0 + i;
            const unsigned slot = r < (unsigned)kBankTable ? s_below[r] + __popc(s_alive[r] & ((1u << cb) - 1u)) : bank_order_slot(r, cb, s_count);
            order[tile_begin + slot] = __ldg(perm + b + i);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// spread
// ---------------------------------------------------------------------------------------------------
constexpr int kSpreadThreads = 256;
constexpr int kSpreadL2Lead = 4;      // iterations between the L2 prefetch of a permutation index and its load
constexpr int kSpreadStages = 3;       // staging buffers of the positions (kSpreadStages - 1 particles in flight per thread)
constexpr int kSpreadModes = 1024;     // most particle types supported (their mode coefficients are staged in shared memory)
// kernel variants (template FLAGS of the spread, CACHE also of the gather)
constexpr int kSpKeys = 1;             // store the tile-major cell key of every particle (introspection, tests)
constexpr int kSpCache = 2;            // write the particle cache {offsets, amplitude, code word, index} for the gather
constexpr int kSpWide = 4;             // 64-bit accumulation: the density of a cell no longer shares 32 bits with the resolution
constexpr int kSpTma = 8;              // flush interior tiles with one 3-D tensor-map reduction (cp.reduce.async.bulk.tensor)
constexpr int kSpTri = 16;             // triclinic box: sheared coordinates (particle_stencil<true>); instantiated without kSpTma
// Wide accumulation.  sm_100a has no native 64-bit shared-memory atomic add (atom.shared.add.u64 compiles to a
// ATOMS.CAST.SPIN loop), so a tap v (|v| < 2^22) is split as v = hi * 2^12 + lo, 0 <= lo < 2^12, and the two parts are
// added to two 32-bit tiles with the native ATOMS.ADD; a tile cell can take 2^20 taps (lo) / 2^21 taps (hi), i.e. about
// 38 000 particles per mesh cell.  The flush adds hi * 2^12 + lo to a 64-bit global mesh (RED.E.ADD.64, native).
constexpr int kWideLoBits = 12;
// counters[] (device, unsigned): [0] ticket, [1] particles handled by the direct path (drifted out of their padded
// tile), [2] particles outside the slab (caller error), [3] unused; [4..5] = [1..2] of the last finished spread,
// [6] cells past half of the 32-bit range seen by the x sweep that consumed the density of the last spread
struct SpreadOut {
    int* mesh;               // integer density, local plane 0 (slab: ghost planes at -1 and nz)
    long long* mesh64;       // the same in 64 bits (wide accumulation; unsharded plans only)
    double* tile_sums;       // [ntiles][2] partial sum a^2, sum a
    double* sums;            // [0] sum a^2 (m_mode_sq, OrderParameterMesh.cc:623), [1] sum a, [2] particles outside the slab
    unsigned* counters;      // see above
    unsigned* h_counters;    // pinned host words [1..3] (device-visible address), or nullptr
    unsigned* keys;          // kSpKeys: tile-major cell key per particle
    float4* cache4;          // kSpCache: particle cache for the gather, tile order: {sx, sy, sz, code word (cache_code())}
    // kSpTma: tensor map of the integer mesh (dims nx, ny, planes incl. ghosts; box = padded tile).  It travels inside this
    // __grid_constant__ kernel parameter: the TMA unit fetches descriptors through its own cache, parameter space is the
    // one place that needs no tensormap proxy fence
    alignas(64) CUtensorMap tmap;
    int tmap_z0;             //   ... plane index of local plane 0 inside that tensor (1 for a slab, else 0)
    int debug;               // timing experiments only (knob 12; results are WRONG): 1 = no flush, 2 = no tile atomics
};

template <int LGT> MHD constexpr int spread_tile_words(int flags) {
    return ((1 << LGT) + 2 * kHaloX) * ((1 << LGT) + 2 * kHalo) * ((1 << LGT) + 2 * kHalo) * ((flags & kSpWide) ? 2 : 1);
}

template <int LGT, int FLAGS>
__global__ void __launch_bounds__(kSpreadThreads)
mesh_spread_kernel(const float4* __restrict__ postype, const unsigned* __restrict__ perm, const unsigned* __restrict__ tstart,
                   const __grid_constant__ Geom g, const float* __restrict__ mode, int ntypes, const float* __restrict__ d_fx,
                   const __grid_constant__ SpreadOut out) {
    constexpr bool KEYS = FLAGS & kSpKeys, CACHE = FLAGS & kSpCache, WIDE = FLAGS & kSpWide, TMA = (FLAGS & kSpTma) && !WIDE;
    constexpr bool TRI = FLAGS & kSpTri;
    constexpr int T = 1 << LGT, PX = T + 2 * kHaloX, PY = T + 2 * kHalo, PZ = PY, P3 = PX * PY * PZ;
    constexpr int TW = spread_tile_words<LGT>(FLAGS);
    extern __shared__ __align__(128) int tile[];      // P3 words (WIDE: the low parts, then P3 words of high parts)
    __shared__ double red[32];
    __shared__ bool is_last;
    // mode coefficients in shared memory: a global load here shares a scoreboard with the position prefetch of the NEXT
    // particle (ptxas puts all three loads of the loop on one), so its first consumer waited for that prefetch in every
    // iteration (measured: 35 % of all stall samples)
    float4* s_pos = reinterpret_cast<float4*>(tile + TW);         // [kSpreadStages][kSpreadThreads] staged positions, behind the tile
    float* s_mode = reinterpret_cast<float*>(s_pos + kSpreadStages * kSpreadThreads);       // [ntypes]
    // the tile is cleared before the programmatic-launch wait: this part overlaps the tail of the previous kernel
    for (int i = threadIdx.x; i < TW / 4; i += kSpreadThreads) reinterpret_cast<int4*>(tile)[i] = make_int4(0, 0, 0, 0);
    pdl_wait(); pdl_trigger();
    for (int i = threadIdx.x; i < ntypes; i += kSpreadThreads) s_mode[i] = __ldg(mode + i);
    const unsigned s = __ldg(tstart + blockIdx.x), e = __ldg(tstart + blockIdx.x + 1);
    unsigned tx, ty, tz;
    tile_coords(blockIdx.x, g, tx, ty, tz);
    const int ox = (int)(tx << LGT) - kHaloX, oy = (int)(ty << LGT) - kHalo, oz = (int)(tz << LGT) - kHalo;
    double sq = 0.0, s1 = 0.0;
    if (e > s) {
        __syncthreads();
        const float scale = __ldg(d_fx);
        unsigned strays = 0, foreign = 0;
        int outer = 0;          // a tap of this thread landed in the outermost y / z layer of the padded tile
        // software pipeline: positions are staged through shared memory with asynchronous copies, kSpreadStages - 1
        // particles ahead (one iteration is shorter than the loaded DRAM latency: with one position in flight 27 % of
        // all stall samples sat on its first use); the index of the particle after those is in flight in a register
        constexpr int D = kSpreadStages - 1, PA = 2;      // positions D particles ahead, indices PA particles further
        unsigned j = s + threadIdx.x;
        unsigned nq[D + PA];
        int buf = 0;
#pragma unroll
        for (int d = 0; d < D + PA; ++d) {
            const unsigned jd = j + d * kSpreadThreads;
            nq[d] = jd < e ? __ldg(perm + jd) : 0u;
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            if (j + d * kSpreadThreads < e) __pipeline_memcpy_async(s_pos + d * kSpreadThreads + threadIdx.x, postype + nq[d], sizeof(float4));
            __pipeline_commit();
        }
        while (j < e) {
            const unsigned jn = j + D * kSpreadThreads;
            int bn = buf + D;
            if (bn >= kSpreadStages) bn -= kSpreadStages;
            if (jn < e) __pipeline_memcpy_async(s_pos + bn * kSpreadThreads + threadIdx.x, postype + nq[D], sizeof(float4));
            __pipeline_commit();
            const unsigned n_far = (jn + PA * kSpreadThreads < e) ? __ldg(perm + jn + PA * kSpreadThreads) : 0u;
            // the register ring shifts at the end of the iteration, so this load must land within one iteration: an L2
            // prefetch a few iterations earlier (one lane per warp: a warp reads one 128-byte line) turns it into an L2 hit
            if ((threadIdx.x & 31) == 0 && jn + (PA + kSpreadL2Lead) * kSpreadThreads < e)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(perm + jn + (PA + kSpreadL2Lead) * kSpreadThreads));
            __pipeline_wait_prior(D);
            const float4 p = s_pos[buf * kSpreadThreads + threadIdx.x];
            const unsigned n = nq[0];
            if (++buf == kSpreadStages) buf = 0;
            const float a = s_mode[__float_as_int(p.w)];
            if (KEYS) {          // the reported cell: the reference's single-precision rule, bit for bit
                const Cell cf = particle_cell(p, g);
                out.keys[n] = key_of(cf.ix, cf.iy, (unsigned)(cf.iz - (int)g.z0) & (g.nz - 1), g);
            }
            Cell c;
            float3 sh;
            particle_stencil<TRI>(p, g, c, sh);
            unsigned lx = 0, ly = 0, lz = 0;
            const bool inside = c.owned && padded_coords(c, ox, oy, oz, g, PX, PY, PZ, lx, ly, lz);
            if (CACHE) out.cache4[j] = make_float4(sh.x, sh.y, sh.z, __uint_as_float(cache_code(lx, ly, lz, (unsigned)__float_as_int(p.w), inside, c.owned)));
            if (c.owned) {
                sq += (double)a * (double)a;
                s1 += (double)a;
                float w[9];
                spread_weights<TRI>(sh, a * scale, w);
                if (inside && (out.debug & 2)) {
                } else if (inside) {
                    // the flush skips the outermost y / z layers of the padded tile unless a particle has drifted that far
                    outer |= (int)((ly - 2u) >= (unsigned)(PY - 4)) | (int)((lz - 2u) >= (unsigned)(PZ - 4));
                    int* base = tile + ((lz - 1) * PY + (ly - 1)) * PX + (lx - 1);
                    // tap (i, jj, k) = fx_round(w[i], w[3 + jj] * w[6 + k]) as on every other path, two per instruction
                    const F2 w01 = f2_pack(w[0], w[1]), wy01 = f2_pack(w[3], w[4]), magic = f2_dup(kFxMagic);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        float wyz[3];
                        f2_unpack(f2_mul(wy01, f2_dup(w[6 + k])), wyz[0], wyz[1]);
                        wyz[2] = f_mul(w[5], w[6 + k]);
#pragma unroll
                        for (int jj = 0; jj < 3; ++jj) {
                            float t0, t1;
                            f2_unpack(f2_fma(w01, f2_dup(wyz[jj]), magic), t0, t1);
                            int* row = base + (k * PY + jj) * PX;
                            const int v0 = __float_as_int(t0) - kFxMagicBits, v1 = __float_as_int(t1) - kFxMagicBits, v2 = fx_round(w[2], wyz[jj]);
                            if (WIDE) {
                                constexpr int LO = (1 << kWideLoBits) - 1;
                                atomicAdd(row, v0 & LO); atomicAdd(row + P3, v0 >> kWideLoBits);
                                atomicAdd(row + 1, v1 & LO); atomicAdd(row + 1 + P3, v1 >> kWideLoBits);
                                atomicAdd(row + 2, v2 & LO); atomicAdd(row + 2 + P3, v2 >> kWideLoBits);
                            } else {
                                atomicAdd(row, v0);
                                atomicAdd(row + 1, v1);
                                atomicAdd(row + 2, v2);
                            }
                        }
                    }
                } else {
                    ++strays;
                    for (int k = 0; k < 3; ++k)
                        for (int jj = 0; jj < 3; ++jj)
                            for (int i = 0; i < 3; ++i) {
                                long long idx;
                                if (tap_index(c, i, jj, k, g, idx)) {
                                    if (WIDE) atomicAdd(reinterpret_cast<unsigned long long*>(out.mesh64 + idx), (unsigned long long)(long long)tap_value(w, i, jj, k));
                                    else atomicAdd(out.mesh + idx, tap_value(w, i, jj, k));
                                }
                            }
                }
            } else {
                ++foreign;
            }
            j += kSpreadThreads;
#pragma unroll
            for (int d = 0; d < D + PA - 1; ++d) nq[d] = nq[d + 1];
            nq[D + PA - 1] = n_far;
        }
        if (strays) atomicAdd(out.counters + 1, strays);
        if (foreign) atomicAdd(out.counters + 2, foreign);
        const int any_outer = __syncthreads_or(outer);
        const size_t plane = (size_t)g.nx * g.ny;
        if (out.debug & 1) {
        } else if (WIDE) {
            // flush into the 64-bit mesh, one thread per row of the padded tile
            for (int row = threadIdx.x; row < PY * PZ; row += kSpreadThreads) {
                TileRow r;
                if (!tile_row(ox, oy, oz, row % PY, row / PY, PX, g, r)) continue;
                long long* dst = out.mesh64 + (long long)r.z * (long long)plane + (size_t)r.y * g.nx;
                const int* lo = tile + row * PX;
#pragma unroll 4
                for (int i = 0; i < PX; ++i) {
                    const long long v = ((long long)lo[i + P3] << kWideLoBits) + (long long)(unsigned)lo[i];
                    if (v != 0) atomicAdd(reinterpret_cast<unsigned long long*>(dst + ((r.x0 + i) & (g.nx - 1))), (unsigned long long)v);
                }
            }
        } else if (TMA && !g.slab && ox >= 0 && oy >= 0 && oz >= 0 && ox + PX <= (int)g.nx && oy + PY <= (int)g.ny && oz + PZ <= (int)g.nz) {
            // flush of a tile that does not wrap periodically: ONE 3-D tensor-map reduction adds the whole padded tile to the
            // mesh (cp.reduce.async.bulk.tensor, UTMAREDG in SASS).  Measured on B200 (tools/micro/tma_reduce_test.cu): the
            // reduction form of the instruction faults with "illegal instruction" when the box leaves the tensor (negative
            // or overhanging coordinates are fine for loads, not for reductions), so boundary tiles take the row path below.
            cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);      // the tile was written through the generic proxy
            if (threadIdx.x == 0) {
                const int32_t crd[3] = {ox, oy, oz};
                cuda::ptx::cp_reduce_async_bulk_tensor(cuda::ptx::space_global, cuda::ptx::space_shared, cuda::ptx::op_add, &out.tmap, crd, tile);
                cuda::ptx::cp_async_bulk_commit_group();
                cuda::ptx::cp_async_bulk_wait_group_read(cuda::ptx::n32_t<0>());       // the tile must outlive the reads
            }
        } else {
            // flush: one thread per row of the padded tile, ONE bulk asynchronous reduction per row (cp.reduce.async.bulk
            // .add.s32, executed by the TMA engine / L2), two if it wraps in x.  The outermost y / z layers are skipped
            // unless a particle of this tile has drifted that far.
            cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);      // the tile was written through the generic proxy
            const unsigned x0 = (unsigned)ox & (g.nx - 1);
            const int room = (int)g.nx - (int)x0, first = room < PX ? room : PX;
            const unsigned tile_s = (unsigned)__cvta_generic_to_shared(tile);
            for (int row = threadIdx.x; row < PY * PZ; row += kSpreadThreads) {
                const int py = row % PY, pz = row / PY;
                const bool inner = (unsigned)(py - 1) < (unsigned)(PY - 2) && (unsigned)(pz - 1) < (unsigned)(PZ - 2);
                int z = oz + pz;
                bool ok = inner || any_outer;
                if (g.slab) ok = ok && z >= -1 && z <= (int)g.nz;
                else z = (int)((unsigned)z & (g.nz - 1));
                if (ok) {
                    const unsigned y = (unsigned)(oy + py) & (g.ny - 1);
                    int* dst = out.mesh + (long long)z * (long long)plane + (size_t)y * g.nx;
                    const unsigned src = tile_s + (unsigned)(row * PX * sizeof(int));
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.s32 [%0], [%1], %2;"
                                 ::"l"(dst + x0), "r"(src), "r"((unsigned)(first * sizeof(int))) : "memory");
                    if (first < PX)
                        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.s32 [%0], [%1], %2;"
                                     ::"l"(dst), "r"(src + (unsigned)(first * sizeof(int))), "r"((unsigned)((PX - first) * sizeof(int))) : "memory");
                }
            }
            cuda::ptx::cp_async_bulk_commit_group();
            cuda::ptx::cp_async_bulk_wait_group_read(cuda::ptx::n32_t<0>());       // the tile must outlive the reads
        }
    }
    // deterministic sums: per-tile partials, the last CTA adds them in tile order
    const double tsq = block_sum(sq, red);
    const double ts1 = block_sum(s1, red);
    if (threadIdx.x == 0) {
        out.tile_sums[2 * blockIdx.x] = tsq;
        out.tile_sums[2 * blockIdx.x + 1] = ts1;
        __threadfence();
        is_last = (atomicAdd(out.counters, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double a2 = 0.0, a1 = 0.0;
    for (unsigned t = threadIdx.x; t < gridDim.x; t += kSpreadThreads) { a2 += __ldcg(out.tile_sums + 2 * t); a1 += __ldcg(out.tile_sums + 2 * t + 1); }
    a2 = block_sum(a2, red);
    a1 = block_sum(a1, red);
    if (threadIdx.x == 0) {
        // every other CTA has taken its ticket, i.e. finished its counter updates: publish the counters of this spread
        // (device snapshot for metad_mesh_get, pinned host words for the drift decision of a later call -- a plain store
        // to mapped host memory instead of a copy node on the critical path) and clear them for the next one.  The range
        // counter [6] / host word [3] belongs to the x sweep that consumes this density (fft_x_fwd_kernel).
        const unsigned c1 = __ldcg(out.counters + 1), c2 = __ldcg(out.counters + 2);
        out.sums[0] = a2;
        out.sums[1] = a1;
        out.sums[2] = (double)c2;
        out.counters[4] = c1; out.counters[5] = c2; out.counters[6] = 0;
        if (out.h_counters) { out.h_counters[1] = c1; out.h_counters[2] = c2; }
        out.counters[1] = 0; out.counters[2] = 0; out.counters[3] = 0;
        out.counters[0] = 0;
    }
}

// ---------------------------------------------------------------------------------------------------
// gather: one CTA per tile; shared tile of Re(IFFT(G)) with halo; one thread per particle of the tile
// ---------------------------------------------------------------------------------------------------
constexpr int kGatherThreads = 256;
constexpr int kGatherStages = 4;       // staging buffers of the particle data (kGatherStages - 1 entries in flight per thread)

// slow path of a particle that drifted out of its padded tile: the 27 taps come from global memory.  Not inlined and
// fed by value, so that the fast path keeps its weights in registers; recomputes cell and weights from the position.
struct GatherDirectArgs { const float* inv; const float* ghost; const Geom* g; };
template <bool TRI>
__device__ __noinline__ float3 gather_direct(float4 p, GatherDirectArgs a) {
    const Geom g = *a.g;
    Cell c;
    float3 sh;
    particle_stencil<TRI>(p, g, c, sh);
    GatherWeights w;
    gather_weights<TRI>(sh, w);
    const size_t plane = (size_t)g.nx * g.ny;
    float t27[27];
    for (int k = 0; k < 3; ++k)
        for (int jj = 0; jj < 3; ++jj)
            for (int i = 0; i < 3; ++i) {
                const unsigned x = (unsigned)(c.ix + i - 1) & (g.nx - 1), y = (unsigned)(c.iy + jj - 1) & (g.ny - 1);
                const int z = c.iz - (int)g.z0 + k - 1;
                float v;
                if (g.slab) v = z == -1 ? a.ghost[(size_t)g.nx * y + x] : (z == (int)g.nz ? a.ghost[plane + (size_t)g.nx * y + x] : a.inv[(size_t)g.nx * y + x + plane * (size_t)z]);
                else v = a.inv[(size_t)g.nx * y + x + plane * (size_t)((unsigned)z & (g.nz - 1))];
                t27[(k * 3 + jj) * 3 + i] = v;
            }
    float3 S;
    gather_sums(t27, 3, 9, w.wx, w.wy, w.wz, w.dx, w.dy, w.dz, S.x, S.y, S.z);
    return S;
}

// what the gather reads per particle: CACHE -- the entry the spread wrote (16 B, tile order) and the tile order (4 B), both
// sequential; otherwise the tile order and the position (16 B, through the order), from which cell and offsets are
// recomputed (the spread then writes nothing per particle, but the gather is slower: profiles/r02_notes.md).
struct GatherIn {
    const float4* postype;
    const unsigned* order;          // tile order (no cache)
    const float4* cache4;           // particle cache
    const float* mode;              // mode coefficients per type (no cache)
    int ntypes;
    int use_tmap;                   // tmap is valid
    alignas(64) CUtensorMap tmap;   // tensor map of Re IFFT(G) (dims nx, ny, nz; box = padded tile), inside the __grid_constant__ parameter
};

template <int LGT, int THREADS = kGatherThreads, int MINB = 3, bool CACHE = true, bool TRI = false>
__global__ void __launch_bounds__(THREADS, MINB)
mesh_gather_kernel(const __grid_constant__ GatherIn in, const unsigned* __restrict__ tstart, const __grid_constant__ Geom g,
                   const float* __restrict__ inv, const float* __restrict__ ghost /* slab mode: planes z0-1 and z0+nz of Re IFFT(G) */,
                   ForceParams fp, const double* __restrict__ d_bias, float4* __restrict__ force, const __grid_constant__ fft::PeerSync sync) {
    constexpr int T = 1 << LGT, PX = T + 2 * kHaloX, PY = T + 2 * kHalo, PZ = PY, P3 = PX * PY * PZ;
    extern __shared__ __align__(128) float ftile[];          // P3 floats, then the staging buffers
    pdl_wait(); pdl_trigger();
    fft::peer_wait(sync);                                   // fused peer mode: the neighbours' halo planes of Re IFFT(G) have arrived
    float4* s_q = reinterpret_cast<float4*>(ftile + P3);    // [kGatherStages][THREADS]: cache entries / positions
    unsigned* s_c = reinterpret_cast<unsigned*>(s_q + kGatherStages * THREADS);  // CACHE: [kGatherStages][THREADS] particle indices
    float* s_mode = reinterpret_cast<float*>(s_c + (CACHE ? kGatherStages * THREADS : 0));       // [ntypes] mode coefficients
    __shared__ uint64_t bar;
    const unsigned s = __ldg(tstart + blockIdx.x), e = __ldg(tstart + blockIdx.x + 1);
    if (e == s) return;                                   // empty tile: nothing to interpolate
    unsigned tx, ty, tz;
    tile_coords(blockIdx.x, g, tx, ty, tz);
    const int ox = (int)(tx << LGT) - kHaloX, oy = (int)(ty << LGT) - kHalo, oz = (int)(tz << LGT) - kHalo;
    const size_t plane = (size_t)g.nx * g.ny;
    if (threadIdx.x == 0) cuda::ptx::mbarrier_init(&bar, THREADS);
    for (int i = threadIdx.x; i < in.ntypes; i += THREADS) s_mode[i] = __ldg(in.mode + i);
    __syncthreads();
    // padded tile of Re IFFT(G).  A tile that does not wrap periodically comes with ONE 3-D tensor-map copy
    // (cp.async.bulk.tensor, UTMALDG); otherwise one bulk asynchronous copy per row (two if the row wraps in x).
    // Completion is counted in bytes on an mbarrier.
    const bool interior = in.use_tmap && !g.slab && ox >= 0 && oy >= 0 && oz >= 0 && ox + PX <= (int)g.nx && oy + PY <= (int)g.ny && oz + PZ <= (int)g.nz;
    if (interior) {
        if (threadIdx.x == 0) {
            const int32_t crd[3] = {ox, oy, oz};
            cuda::ptx::cp_async_bulk_tensor(cuda::ptx::space_cluster, cuda::ptx::space_global, ftile, &in.tmap, crd, &bar);
        }
        cuda::ptx::mbarrier_arrive_expect_tx(cuda::ptx::sem_release, cuda::ptx::scope_cta, cuda::ptx::space_shared, &bar,
                                             threadIdx.x == 0 ? (unsigned)(P3 * sizeof(float)) : 0u);
    } else {
        unsigned bytes = 0;
        for (int row = threadIdx.x; row < PY * PZ; row += THREADS) {
            TileRow r;
            float* dst = ftile + row * PX;
            if (tile_row(ox, oy, oz, row % PY, row / PY, PX, g, r)) {
                const float* src;
                if (g.slab && r.z == -1) src = ghost;
                else if (g.slab && r.z == (int)g.nz) src = ghost + plane;
                else src = inv + plane * (size_t)r.z;
                src += (size_t)r.y * g.nx;
                cuda::ptx::cp_async_bulk(cuda::ptx::space_cluster, cuda::ptx::space_global, dst, src + r.x0,
                                         (unsigned)(r.first * sizeof(float)), &bar);
                if (r.first < PX)
                    cuda::ptx::cp_async_bulk(cuda::ptx::space_cluster, cuda::ptx::space_global, dst + r.first, src,
                                             (unsigned)((PX - r.first) * sizeof(float)), &bar);
                bytes += PX * sizeof(float);
            } else {
                for (int i = 0; i < PX; ++i) dst[i] = 0.f;          // plane outside the slab and its ghosts: never read by owned particles
            }
        }
        cuda::ptx::mbarrier_arrive_expect_tx(cuda::ptx::sem_release, cuda::ptx::scope_cta, cuda::ptx::space_shared, &bar, bytes);
    }
    const float scale = (float)(fp.two_over_n * *d_bias);
    unsigned j = s + threadIdx.x;
    int buf = 0;
    if constexpr (CACHE) {
        // one thread per particle of the tile; offsets, amplitude, padded-tile cell and particle index come from the cache
        // the spread wrote.  The entries are staged through shared memory with asynchronous copies, kGatherStages - 1 in
        // flight per thread: one iteration of this loop is shorter than the loaded DRAM latency
#pragma unroll
        for (int d = 0; d < kGatherStages - 1; ++d) {
            const unsigned jd = j + d * THREADS;
            if (jd < e) {
                __pipeline_memcpy_async(s_q + d * THREADS + threadIdx.x, in.cache4 + jd, sizeof(float4));
                __pipeline_memcpy_async(s_c + d * THREADS + threadIdx.x, in.order + jd, sizeof(unsigned));
            }
            __pipeline_commit();
        }
        while (!cuda::ptx::mbarrier_try_wait_parity(&bar, 0)) {}
        __syncthreads();                                       // zero-filled rows (generic stores) are visible too
        for (; j < e; j += THREADS) {
            const unsigned jn = j + (kGatherStages - 1) * THREADS;
            int bn = buf + kGatherStages - 1;
            if (bn >= kGatherStages) bn -= kGatherStages;
            if (jn < e) {
                __pipeline_memcpy_async(s_q + bn * THREADS + threadIdx.x, in.cache4 + jn, sizeof(float4));
                __pipeline_memcpy_async(s_c + bn * THREADS + threadIdx.x, in.order + jn, sizeof(unsigned));
            }
            __pipeline_commit();
            __pipeline_wait_prior(kGatherStages - 1);          // everything but the newest kGatherStages - 1 groups has landed
            const float4 q = s_q[buf * THREADS + threadIdx.x];
            const unsigned n = s_c[buf * THREADS + threadIdx.x];
            const unsigned code = __float_as_uint(q.w);
            if (++buf == kGatherStages) buf = 0;
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (code & kCacheOwned) {
                float Sx, Sy, Sz;
                if (code & kCacheInside) {
                    GatherWeights w;
                    gather_weights<TRI>(make_float3(q.x, q.y, q.z), w);
                    const unsigned lx = code & 31u, ly = (code >> 5) & 31u, lz = (code >> 10) & 31u;
                    gather_sums(ftile + ((lz - 1) * PY + (ly - 1)) * PX + (lx - 1), PX, PX * PY, w.wx, w.wy, w.wz, w.dx, w.dy, w.dz, Sx, Sy, Sz);
                } else {
                    GatherDirectArgs da;
                    da.inv = inv; da.ghost = ghost; da.g = &g;
                    const float3 S = gather_direct<TRI>(__ldg(in.postype + n), da);
                    Sx = S.x; Sy = S.y; Sz = S.z;
                }
                f = force_from_sums(Sx, Sy, Sz, s_mode[(code >> 15) & 1023u], fp, scale);
            }
            force[n] = f;
        }
    } else {
        // no cache: positions come through the tile order (index two stages ahead of its position, like the spread), cell
        // and offsets are recomputed with the expressions of the spread
        constexpr int D = kGatherStages - 1, PA = 2;
        unsigned nq[D + PA];
#pragma unroll
        for (int d = 0; d < D + PA; ++d) {
            const unsigned jd = j + d * THREADS;
            nq[d] = jd < e ? __ldg(in.order + jd) : 0u;
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            if (j + d * THREADS < e) __pipeline_memcpy_async(s_q + d * THREADS + threadIdx.x, in.postype + nq[d], sizeof(float4));
            __pipeline_commit();
        }
        while (!cuda::ptx::mbarrier_try_wait_parity(&bar, 0)) {}
        __syncthreads();
        for (; j < e; j += THREADS) {
            const unsigned jn = j + D * THREADS;
            int bn = buf + D;
            if (bn >= kGatherStages) bn -= kGatherStages;
            if (jn < e) __pipeline_memcpy_async(s_q + bn * THREADS + threadIdx.x, in.postype + nq[D], sizeof(float4));
            __pipeline_commit();
            const unsigned n_far = (jn + PA * THREADS < e) ? __ldg(in.order + jn + PA * THREADS) : 0u;
            if ((threadIdx.x & 31) == 0 && jn + (PA + kSpreadL2Lead) * THREADS < e)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(in.order + jn + (PA + kSpreadL2Lead) * THREADS));
            __pipeline_wait_prior(D);
            const float4 p = s_q[buf * THREADS + threadIdx.x];
            const unsigned n = nq[0];
            if (++buf == kGatherStages) buf = 0;
            Cell c;
            float3 sh;
            particle_stencil<TRI>(p, g, c, sh);
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c.owned) {
                unsigned lx, ly, lz;
                float Sx, Sy, Sz;
                if (padded_coords(c, ox, oy, oz, g, PX, PY, PZ, lx, ly, lz)) {
                    GatherWeights w;
                    gather_weights<TRI>(sh, w);
                    gather_sums(ftile + ((lz - 1) * PY + (ly - 1)) * PX + (lx - 1), PX, PX * PY, w.wx, w.wy, w.wz, w.dx, w.dy, w.dz, Sx, Sy, Sz);
                } else {
                    GatherDirectArgs da;
                    da.inv = inv; da.ghost = ghost; da.g = &g;
                    const float3 S = gather_direct<TRI>(p, da);
                    Sx = S.x; Sy = S.y; Sz = S.z;
                }
                f = force_from_sums(Sx, Sy, Sz, s_mode[__float_as_int(p.w)], fp, scale);
            }
            force[n] = f;
#pragma unroll
            for (int d = 0; d < D + PA - 1; ++d) nq[d] = nq[d + 1];
            nq[D + PA - 1] = n_far;
        }
    }
}
#endif  // __CUDACC__

}  // namespace mesh
}  // namespace metad
