// mesh_general.cuh -- OrderParameterMesh for mesh sizes the tiled path does not take: any number of points per direction
// (1 <= n <= 1024; the reference restricts the mesh to powers of two only under domain decomposition,
// OrderParameterMesh.cc:70-79 -- kiss_fftnd / cuFFT transform any length), orthorhombic or triclinic box, one GPU.
//
// Same arithmetic as the tiled path wherever that matters for parity -- the stencil (particle_stencil: accurate cell +
// offset), the TSC weights, the fixed-point taps (tap_value) accumulated in integers (here: 64-bit global atomics, so the
// density is bitwise independent of the particle order and there is no range to watch), mean removal before the
// transforms, fp64 energy partials summed in a fixed order -- but none of its machinery: no tile order, no shared-memory
// tiles, a complex-to-complex transform of the full mesh.  It is the "works for every mesh" path, not the fast one
// (DESIGN.md section 4): one particle per thread for spread and gather, one line per CTA for the transforms.
//
//   gen_spread_kernel     assignParticles                      OrderParameterMesh.cc:517-640
//   gen_density_kernel    integer density -> complex mesh (clears the accumulator)
//   gen_fft_kernel x 3    forward transform                    :650-657 (kiss_fftnd)
//   gen_conv_kernel       updateMeshes + computeCV (+ computeQmax / computeVirial epilogues)   :659-747, 866-923, 970-1050, 1108-1179
//   gen_fft_kernel x 3    inverse transform                    :737-745
//   gen_gather_kernel     interpolateForces                    :749-864
//
// Transform: Stockham autosort, mixed radix.  The length n is factored into radices 4, 2 and the odd primes; a stage of
// radix R maps x -> y with, for butterfly j in [0, n/R) and output q in [0, R):
//     k = j mod Ns,   y[(j div Ns) Ns R + k + q Ns] = sum_{r<R} x[j + r n/R] w_n^(r (k n/(Ns R) + q n/R)),   w_n = exp(-+2 pi i/n)
// (Ns = product of the radices of the earlier stages).  One table of the n-th roots of unity serves every stage; an output
// costs R complex multiply-adds whatever R is, so a prime length is an O(n^2) transform per line -- correct, not fast.
#pragma once
#include "mesh_kernels.cuh"
#include "mesh_fft_kernels.cuh"

namespace metad {
namespace meshgen {

using namespace metad::mesh;

constexpr unsigned kMaxLen = 1024;        // longest line: three arrays of n complex numbers in shared memory (24 KB)
constexpr int kMaxStages = 12;
constexpr int kFftThreads = 256;
constexpr int kParticleThreads = 256;
constexpr int kConvThreads = 256;
constexpr int kMaxConvBlocks = 1024;

struct Radices { int n, count; int r[kMaxStages]; };
inline Radices factorize(unsigned n) {
    Radices f; memset(&f, 0, sizeof f); f.n = (int)n;
    unsigned m = n;
    while (m % 4 == 0) { f.r[f.count++] = 4; m /= 4; }
    while (m % 2 == 0) { f.r[f.count++] = 2; m /= 2; }
    for (unsigned p = 3; m > 1; p += 2)
        while (m % p == 0) { f.r[f.count++] = (int)p; m /= p; }
    return f;
}

// one output of one stage (header); tw[m] = exp(-2 pi i m/n), sign = -1 forward / +1 inverse (conjugate table)
MHD float2 stage_output(const float2* x, const float2* tw, int n, int R, int Ns, int w, float sign, int& out_index) {
    const int m = n / R, j = w % m, q = w / m, k = j % Ns;
    out_index = (j / Ns) * Ns * R + k + q * Ns;
    const int inc = (k * (n / (Ns * R)) + q * m) % n;
    float ar = 0.f, ai = 0.f;
    int idx = 0;
    for (int r = 0; r < R; ++r) {
        const float2 v = x[j + r * m];
        const float tr = tw[idx].x, ti = -sign * tw[idx].y;
        ar += v.x * tr - v.y * ti;
        ai += v.x * ti + v.y * tr;
        idx += inc;
        if (idx >= n) idx -= n;
    }
    return make_float2(ar, ai);
}

// lines of the [z][y][x] mesh along one axis: element e of line l sits at base(l) + e * stride
struct LineMap { unsigned nx, ny, nz; int axis; };
MHD void line_of(const LineMap& lm, unsigned line, size_t& base, size_t& stride, unsigned& n) {
    const size_t plane = (size_t)lm.nx * lm.ny;
    if (lm.axis == 0) { base = (size_t)line * lm.nx; stride = 1; n = lm.nx; }
    else if (lm.axis == 1) { base = (size_t)(line / lm.nx) * plane + line % lm.nx; stride = lm.nx; n = lm.ny; }
    else { base = line; stride = plane; n = lm.nz; }
}
MHD unsigned line_count(const LineMap& lm) {
    return lm.axis == 0 ? lm.ny * lm.nz : (lm.axis == 1 ? lm.nx * lm.nz : lm.nx * lm.ny);
}

// periodic neighbour index (n need not be a power of two; c in [0, n), d in {-1, 0, 1})
MHD unsigned wrap1(int c, int d, unsigned n) {
    const int v = c + d;
    return (unsigned)(v < 0 ? v + (int)n : (v >= (int)n ? v - (int)n : v));
}
MHD size_t tap_cell(const Cell& c, int i, int j, int k, const Geom& g) {
    return (size_t)wrap1(c.ix, i - 1, g.nx) + (size_t)g.nx * ((size_t)wrap1(c.iy, j - 1, g.ny) + (size_t)g.ny * (size_t)wrap1(c.iz, k - 1, g.nz));
}

inline void geom_set_dims_general(Geom& g, unsigned nx, unsigned ny, unsigned nz) {
    memset(&g, 0, sizeof g);
    g.nx = nx; g.ny = ny; g.nz = nz; g.nzg = nz; g.z0 = 0; g.slab = 0;
}

// pointwise convolution of one mode of the FULL spectrum (updateMeshes :659-735 and the summand of computeCV :880-903):
// f = F/N, G = f (|f|^2 - d chi_k), e += |f|^2 (|f|^2 - 2 d chi_k) for flat != 0; d = mode_sq / (2 N^2); chi_k = the
// interpolation factor squared, which is the indicator of the non-negative Miller octant (fft::nonneg).  dc = what the
// mean removal took out of f_0 when k = 0 matters for the forces (literal triclinic offsets), else 0: a constant in
// Re IFFT(G) does not move a force whose derivative weights sum to zero, and carrying it through single-precision
// transforms would only add rounding noise of its size to the part that does.
struct ConvGeom { unsigned nx, ny, nz; float inv_n, d, dc; };
MHD float2 conv_mode(float2 F, size_t flat, const ConvGeom& c, double& e, float& val, unsigned& kx, unsigned& ky, unsigned& kz) {
    kx = (unsigned)(flat % c.nx); ky = (unsigned)((flat / c.nx) % c.ny); kz = (unsigned)(flat / ((size_t)c.nx * c.ny));
    const float fr = F.x * c.inv_n + (flat == 0 ? c.dc : 0.f), fi = F.y * c.inv_n;
    val = fr * fr + fi * fi;
    const float dk = (fft::nonneg(kx, c.nx) && fft::nonneg(ky, c.ny) && fft::nonneg(kz, c.nz)) ? c.d : 0.f;
    const float g = val - dk;
    if (flat != 0) e += (double)val * ((double)val - 2.0 * (double)dk);
    return make_float2(fr * g, fi * g);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
// spread: one particle per thread, 27 integer taps into the 64-bit mesh; per-block partial sums of a^2 and a
template <bool TRI>
__global__ void __launch_bounds__(kParticleThreads)
gen_spread_kernel(const float4* __restrict__ postype, unsigned N, const __grid_constant__ Geom g, const float* __restrict__ mode, float scale,
                  unsigned long long* __restrict__ mesh64, int* __restrict__ cells, double* __restrict__ block_sums) {
    __shared__ double red[32];
    double sq = 0.0, s1 = 0.0;
    for (unsigned n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float4 p = ld_stream(postype + n);
        const float a = __ldg(mode + __float_as_int(p.w));
        if (cells) {          // the reported cell: the reference's single-precision rule, bit for bit
            const Cell cf = particle_cell(p, g);
            cells[3 * (size_t)n] = cf.ix; cells[3 * (size_t)n + 1] = cf.iy; cells[3 * (size_t)n + 2] = cf.iz;
        }
        Cell c;
        float3 sh;
        particle_stencil<TRI, false>(p, g, c, sh);
        float w[9];
        spread_weights<TRI>(sh, a * scale, w);
        for (int k = 0; k < 3; ++k)
            for (int j = 0; j < 3; ++j)
                for (int i = 0; i < 3; ++i) {
                    const int v = tap_value(w, i, j, k);
                    if (v != 0) atomicAdd(mesh64 + tap_cell(c, i, j, k, g), (unsigned long long)(long long)v);
                }
        sq += (double)a * (double)a;
        s1 += (double)a;
    }
    const double rsq = block_sum(sq, red);
    const double rs1 = block_sum(s1, red);
    if (threadIdx.x == 0) { block_sums[2 * blockIdx.x] = rsq; block_sums[2 * blockIdx.x + 1] = rs1; }
}
// fixed-order sum of the per-block partials -> sums[0] = sum a^2, sums[1] = sum a, sums[2] = 0
__global__ void gen_sums_kernel(const double* __restrict__ block_sums, unsigned nblocks, double* __restrict__ sums) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (unsigned i = 0; i < nblocks; ++i) { a += block_sums[2 * i]; b += block_sums[2 * i + 1]; }
        sums[0] = a; sums[1] = b; sums[2] = 0.0;
    }
}
// integer density -> complex mesh with the mean removed; clears the accumulator for the next call
__global__ void __launch_bounds__(kConvThreads)
gen_density_kernel(long long* __restrict__ mesh64, float2* __restrict__ spec, float* __restrict__ rho_keep, size_t M, float inv_scale,
                   const double* __restrict__ d_sums, double inv_cells) {
    const float mean = (float)(d_sums[1] * inv_cells);
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < M; c += (size_t)gridDim.x * blockDim.x) {
        const float r = __ll2float_rn(mesh64[c]) * inv_scale;
        mesh64[c] = 0;
        if (rho_keep) rho_keep[c] = r;
        spec[c] = make_float2(r - mean, 0.f);
    }
}
// one line per CTA: load (strided), all stages in shared memory, store
__global__ void __launch_bounds__(kFftThreads)
gen_fft_kernel(float2* __restrict__ data, LineMap lm, Radices rad, const float2* __restrict__ g_tw, float sign) {
    extern __shared__ float2 gsm[];
    size_t base, stride;
    unsigned n;
    line_of(lm, blockIdx.x, base, stride, n);
    float2 *x = gsm, *y = gsm + n, *tw = gsm + 2 * n;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) { x[i] = data[base + i * stride]; tw[i] = __ldg(g_tw + i); }
    __syncthreads();
    int Ns = 1;
    for (int s = 0; s < rad.count; ++s) {
        const int R = rad.r[s];
        for (unsigned w = threadIdx.x; w < n; w += blockDim.x) {
            int o;
            const float2 v = stage_output(x, tw, (int)n, R, Ns, (int)w, sign, o);
            y[o] = v;
        }
        __syncthreads();
        float2* t = x; x = y; y = t;
        Ns *= R;
    }
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) data[base + i * stride] = x[i];
}
// convolution + CV energy (+ epilogues), full spectrum; the last block sums the partials in order (energy_block_finish)
template <bool EXTRAS>
__global__ void __launch_bounds__(kConvThreads)
gen_conv_kernel(float2* __restrict__ spec, size_t M, fft::ConvParams cp) {
    ConvGeom cg;
    cg.nx = cp.nx; cg.ny = cp.ny; cg.nz = cp.nz; cg.inv_n = cp.inv_n;
    const double nd = cp.n_global;
    cg.d = (float)(0.5 * (*cp.d_mode_sq) / nd / nd);
    cg.dc = 0.f;
    if (cp.dc_restore) {          // literal triclinic offsets only (fft::ConvParams::dc_restore); otherwise k = 0 stays out, as in the tiled path
        const float mean = (float)(cp.d_mode_sq[1] * cp.inv_cells);
        cg.dc = (float)((double)mean / cp.inv_cells * (double)cp.inv_n);
    }
    double e = 0.0;
    fft::ExtraAcc xa;
    fft::extra_init(xa);
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < M; c += (size_t)gridDim.x * blockDim.x) {
        float val;
        unsigned kx, ky, kz;
        spec[c] = conv_mode(spec[c], c, cg, e, val, kx, ky, kz);
        if (EXTRAS) fft::extra_add(xa, cp, val, kx, ky, kz, 1.0f);
    }
    fft::energy_block_finish<EXTRAS>(e, cp, &xa);
}
// gather: one particle per thread, 27 values of Re IFFT(G) from global memory
template <bool TRI>
__global__ void __launch_bounds__(kParticleThreads)
gen_gather_kernel(const float4* __restrict__ postype, unsigned N, const __grid_constant__ Geom g, const float* __restrict__ mode,
                  const float2* __restrict__ spec, ForceParams fp, const double* __restrict__ d_bias, float4* __restrict__ force) {
    const float scale = (float)(fp.two_over_n * *d_bias);
    for (unsigned n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float4 p = ld_stream(postype + n);
        Cell c;
        float3 sh;
        particle_stencil<TRI, false>(p, g, c, sh);
        GatherWeights w;
        gather_weights<TRI>(sh, w);
        float t27[27];
        for (int k = 0; k < 3; ++k)
            for (int j = 0; j < 3; ++j)
                for (int i = 0; i < 3; ++i) t27[(k * 3 + j) * 3 + i] = __ldg(&spec[tap_cell(c, i, j, k, g)].x);
        float Sx, Sy, Sz;
        gather_sums(t27, 3, 9, w.wx, w.wy, w.wz, w.dx, w.dy, w.dz, Sx, Sy, Sz);
        force[n] = force_from_sums(Sx, Sy, Sz, __ldg(mode + __float_as_int(p.w)), fp, scale);
    }
}
#endif  // __CUDACC__

}  // namespace meshgen
}  // namespace metad
