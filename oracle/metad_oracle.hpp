// metad_oracle.hpp -- CPU restatement of the reference's CV + bias-force hot path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing under oracle/ is part of the product.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load it, and only as the checker or the CPU baseline -- never as the thing shipped.
//
// *** PARITY: PINNED TO THE REFERENCE'S OWN CODE (HOOMD's BoxDim / kiss_fft, which are not in the reference tree, excepted). ***
// The reference (jglaser/metadynamics-plugin) ships no golden vectors and no unit tests, and its build needs
// HOOMD-blue 2.x, which is neither installed nor vendored.  This file restates the reference's *CPU* code path
// formula by formula, each function citing the reference file:line it follows (paths relative to
// /root/reference/metadynamics/).  It is pinned in two ways:
//   (1) against the reference's own classes: CollectiveVariable.cc, LamellarOrderParameter.cc, OrderParameterMesh.cc,
//       AspectRatio.cc, Density.cc, IndexGrid.cc, IntegratorMetaDynamics.cc and WellTemperedEnsemble.cc are compiled UNMODIFIED from /root/reference against a
//       HOOMD stand-in (oracle/ref_shim/, oracle/ref_capi.cc, `make -C oracle ref` -> oracle/_ref/) and run on seeded
//       inputs (tests/golden/make_ref_golden.py -> tests/golden/ref_golden.npz; tests/test_reference_build.py).  The
//       density mesh of this file equals the reference's BIT FOR BIT in the float and the double build (every cell
//       index, weight and summation order); CV values and forces agree to 1e-12 in double (the FFTs differ); Lamellar
//       modes / CV / forces to 1e-13; the bias grid (bias factors after every step, grid, reweighted estimator, weights,
//       sigma grid, histograms) BIT FOR BIT for 1-3 CVs, standard and well-tempered, both builds; umbrella, aspect
//       ratio and IndexGrid exactly;
//   (2) against independent numpy restatements (numpy.fft, closed-form TSC, analytic single-particle / lattice cases,
//       finite differences) in tests/test_oracle.py.
//       WellTemperedEnsemble.cc likewise (its header needs ENABLE_CUDA to compile: oracle/ref_shim_cuda/ provides inert
//       stand-ins, the CPU branch runs): CV and the scaled arrays BIT FOR BIT, both builds; Density.cc and computeQmax too.
// NOT pinned by an executable: the two pieces of HOOMD itself that the stand-in has to restate as well --
// BoxDim (lo/hi/L/Linv, makeFraction = (v - lo) * Linv, branching minImage) and kiss_fftnd (unnormalised DFT, dims
// slowest first).  HOOMD's version is not pinned by the reference (no submodule).
//
// Everything is templated on S = float (HOOMD SINGLE_PRECISION build) or double (HOOMD
// default build).  The double instance is the "truth" for floating-point tolerances, the
// float instance is the single-precision CPU build used for bit-exact integer checks
// (cell indices, histogram bins, IndexGrid).
#pragma once
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace oracle {

// ----------------------------------------------------------------------------------------
// small vector helper (HOOMD Scalar3 semantics: component-wise ops)
// ----------------------------------------------------------------------------------------
template <class S> struct V3 { S x, y, z; };
template <class S> inline V3<S> v3(S x, S y, S z) { return V3<S>{x, y, z}; }
template <class S> inline V3<S> operator+(V3<S> a, V3<S> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class S> inline V3<S> operator-(V3<S> a, V3<S> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class S> inline V3<S> operator*(V3<S> a, V3<S> b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
template <class S> inline V3<S> operator/(V3<S> a, V3<S> b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
template <class S> inline V3<S> operator*(V3<S> a, S s) { return {a.x * s, a.y * s, a.z * s}; }
template <class S> inline V3<S> operator*(S s, V3<S> a) { return {s * a.x, s * a.y, s * a.z}; }
template <class S> inline V3<S> operator/(V3<S> a, S s) { return {a.x / s, a.y / s, a.z / s}; }
template <class S> inline V3<S> operator-(V3<S> a) { return {-a.x, -a.y, -a.z}; }
template <class S> inline S dot(V3<S> a, V3<S> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// ----------------------------------------------------------------------------------------
// BoxDim -- restates the HOOMD-blue 2.x hoomd/BoxDim.h members the path calls
// (call sites: OrderParameterMesh.cc:543,570-573,783,807-810,362-369;
//  LamellarOrderParameter.cc:151-159).  lo = -L/2, hi = +L/2.
//   makeFraction(v)    = ((v - lo) - tilt terms) * Linv, Linv = 1/(hi - lo)   (ghost width 0)
//   makeCoordinates(f) = lo + f*L, then x += xy*y + xz*z, y += yz*z
//   minImage(v)        = host (branching) variant: one box length per direction
//   getLatticeVector   = (Lx,0,0), (Ly*xy,Ly,0), (Lz*xz,Lz*yz,Lz)
// ----------------------------------------------------------------------------------------
template <class S> struct Box {
    V3<S> lo, hi, L, Linv;
    S xy = 0, xz = 0, yz = 0;
    bool periodic[3] = {true, true, true};

    static Box make(double Lx, double Ly, double Lz, double xy_ = 0, double xz_ = 0, double yz_ = 0) {
        Box b;
        b.L = {(S)Lx, (S)Ly, (S)Lz};
        b.hi = b.L / S(2.0);
        b.lo = -b.hi;
        b.Linv = {S(1.0) / b.L.x, S(1.0) / b.L.y, S(1.0) / b.L.z};
        b.xy = (S)xy_; b.xz = (S)xz_; b.yz = (S)yz_;
        return b;
    }
    V3<S> makeFraction(V3<S> v) const {
        V3<S> d = v - lo;
        d.x -= (xz - yz * xy) * v.z + xy * v.y;
        d.y -= yz * v.z;
        return d * Linv;          // HOOMD multiplies by the stored reciprocal m_Linv = 1/(hi - lo); a division would round differently
    }
    V3<S> makeCoordinates(V3<S> f) const {
        V3<S> v = lo + f * L;
        v.x += xy * v.y + xz * v.z;
        v.y += yz * v.z;
        return v;
    }
    V3<S> minImage(V3<S> v) const {
        V3<S> w = v;
        if (periodic[2]) {
            if (w.z >= hi.z) { w.z -= L.z; w.y -= L.z * yz; w.x -= L.z * xz; }
            else if (w.z < lo.z) { w.z += L.z; w.y += L.z * yz; w.x += L.z * xz; }
        }
        if (periodic[1]) {
            if (w.y >= hi.y) { w.y -= L.y; w.x -= L.y * xy; }
            else if (w.y < lo.y) { w.y += L.y; w.x += L.y * xy; }
        }
        if (periodic[0]) {
            if (w.x >= hi.x) w.x -= L.x;
            else if (w.x < lo.x) w.x += L.x;
        }
        return w;
    }
    V3<S> latticeVector(int i) const {
        if (i == 0) return {L.x, S(0), S(0)};
        if (i == 1) return {L.y * xy, L.y, S(0)};
        return {L.z * xz, L.z * yz, L.z};
    }
    S volume() const { return L.x * L.y * L.z; }
};

// reciprocal lattice vectors b_i = (a_j x a_k)/V, optionally times 2*pi
// (OrderParameterMesh.cc:362-369 with 2pi, :761-769 without; LamellarOrderParameter.cc:151-159)
template <class S> inline void reciprocal(const Box<S>& box, bool two_pi, V3<S>& b1, V3<S>& b2, V3<S>& b3) {
    V3<S> a1 = box.latticeVector(0), a2 = box.latticeVector(1), a3 = box.latticeVector(2);
    S V = box.volume();
    b1 = v3<S>(a2.y * a3.z - a2.z * a3.y, a2.z * a3.x - a2.x * a3.z, a2.x * a3.y - a2.y * a3.x);
    b2 = v3<S>(a3.y * a1.z - a3.z * a1.y, a3.z * a1.x - a3.x * a1.z, a3.x * a1.y - a3.y * a1.x);
    b3 = v3<S>(a1.y * a2.z - a1.z * a2.y, a1.z * a2.x - a1.x * a2.z, a1.x * a2.y - a1.y * a2.x);
    if (two_pi) {
        b1 = S(2.0 * M_PI) * b1 / V; b2 = S(2.0 * M_PI) * b2 / V; b3 = S(2.0 * M_PI) * b3 / V;
    } else {
        b1 = b1 / V; b2 = b2 / V; b3 = b3 / V;
    }
}

inline unsigned type_of(const float* postype4) {
    // __scalar_as_int(postype.w): the type id is stored as raw bits in the .w float
    int32_t t; std::memcpy(&t, postype4 + 3, 4); return (unsigned)t;
}

// ----------------------------------------------------------------------------------------
// unnormalised 3-D complex DFT with kiss_fftnd semantics (dims = {nz,ny,nx}; forward
// e^{-i}, inverse e^{+i}; neither normalised): OrderParameterMesh.cc:319-325,655,719.
// kiss_fft itself lives in HOOMD (hoomd/extern/kiss_fftnd.h), not in the reference tree;
// any exact DFT is an admissible restatement up to rounding.  Power-of-two lines use an
// iterative radix-2 transform, other lengths a direct O(n^2) DFT.  Twiddles are evaluated
// in double and rounded to S (as kiss_fft does).
// ----------------------------------------------------------------------------------------
template <class S> struct LineFFT {
    unsigned n; int sign; bool pow2;
    std::vector<std::complex<S>> tw;   // tw[k] = exp(sign*2*pi*i*k/n), k<n
    std::vector<unsigned> rev;
    std::vector<std::complex<S>> tmp;
    LineFFT(unsigned n_, int sign_) : n(n_), sign(sign_) {
        pow2 = n && !(n & (n - 1));
        tw.resize(n);
        for (unsigned k = 0; k < n; ++k) {
            double ph = sign * 2.0 * M_PI * (double)k / (double)n;
            tw[k] = std::complex<S>((S)std::cos(ph), (S)std::sin(ph));
        }
        if (pow2) {
            rev.resize(n);
            unsigned lg = 0; while ((1u << lg) < n) ++lg;
            for (unsigned i = 0; i < n; ++i) {
                unsigned r = 0;
                for (unsigned b = 0; b < lg; ++b) if (i & (1u << b)) r |= 1u << (lg - 1 - b);
                rev[i] = r;
            }
        }
        tmp.resize(n);
    }
    // in-place transform of a strided line
    void run(std::complex<S>* d, size_t stride) {
        if (n == 1) return;
        for (unsigned i = 0; i < n; ++i) tmp[i] = d[i * stride];
        if (pow2) {
            for (unsigned i = 0; i < n; ++i) if (rev[i] > i) std::swap(tmp[i], tmp[rev[i]]);
            for (unsigned len = 2; len <= n; len <<= 1) {
                unsigned half = len >> 1, step = n / len;
                for (unsigned s = 0; s < n; s += len)
                    for (unsigned j = 0; j < half; ++j) {
                        std::complex<S> u = tmp[s + j], t = tmp[s + j + half] * tw[j * step];
                        tmp[s + j] = u + t; tmp[s + j + half] = u - t;
                    }
            }
            for (unsigned i = 0; i < n; ++i) d[i * stride] = tmp[i];
        } else {
            std::vector<std::complex<S>> out(n);
            for (unsigned k = 0; k < n; ++k) {
                std::complex<S> acc(0, 0);
                for (unsigned j = 0; j < n; ++j) acc += tmp[j] * tw[(size_t)((uint64_t)j * k % n)];
                out[k] = acc;
            }
            for (unsigned i = 0; i < n; ++i) d[i * stride] = out[i];
        }
    }
};

template <class S>
inline void fft3d(const std::vector<std::complex<S>>& in, std::vector<std::complex<S>>& out,
                  unsigned nx, unsigned ny, unsigned nz, int sign) {
    out = in;
    LineFFT<S> fx(nx, sign), fy(ny, sign), fz(nz, sign);
    for (unsigned z = 0; z < nz; ++z)
        for (unsigned y = 0; y < ny; ++y) fx.run(&out[(size_t)nx * (y + (size_t)ny * z)], 1);
    for (unsigned z = 0; z < nz; ++z)
        for (unsigned x = 0; x < nx; ++x) fy.run(&out[x + (size_t)nx * ny * z], nx);
    for (unsigned y = 0; y < ny; ++y)
        for (unsigned x = 0; x < nx; ++x) fz.run(&out[x + (size_t)nx * y], (size_t)nx * ny);
}

// ----------------------------------------------------------------------------------------
// OrderParameterMesh (CPU path, single domain => no ghost cells, grid_dim = mesh_points)
// ----------------------------------------------------------------------------------------
template <class S> struct MeshCV {
    unsigned nx, ny, nz, M;
    std::vector<S> mode;            // per-type mode coefficient a(type)
    Box<S> box;                     // local box == global box (single domain)
    unsigned N_global = 0;
    S mode_sq = 0, cv = 0;
    std::vector<std::complex<S>> mesh, fourier, fourier_G, inv;
    std::vector<S> interp;          // m_interpolation_f
    std::vector<int> cells;         // (ix,iy,iz) per particle from the last assign()

    MeshCV(unsigned nx_, unsigned ny_, unsigned nz_, const std::vector<double>& mode_, const Box<S>& b)
        : nx(nx_), ny(ny_), nz(nz_), M(nx_ * ny_ * nz_), box(b) {
        for (double m : mode_) mode.push_back((S)m);
        mesh.assign(M, {0, 0}); fourier.assign(M, {0, 0}); fourier_G.assign(M, {0, 0}); inv.assign(M, {0, 0});
        compute_interpolation();
    }

    // assignTSC: OrderParameterMesh.cc:457-468
    static S W(S x) {
        S xsq = x * x;
        S xabs = std::sqrt(xsq);
        if (xsq <= S(1.0 / 4.0)) return S(3.0 / 4.0) - xsq;
        else if (xsq <= S(9.0 / 4.0)) return S(1.0 / 2.0) * (S(3.0 / 2.0) - xabs) * (S(3.0 / 2.0) - xabs);
        return S(0.0);
    }
    // assignTSCderiv: OrderParameterMesh.cc:470-483.  The reference writes |x| as copysignf(x, 1) even when
    // Scalar = double (SURVEY 8a note 3), i.e. a double build rounds |x| to float.  In a SINGLE_PRECISION build
    // copysignf is exact.  `literal_copysignf` selects the literal restatement (default); with it switched off
    // the double instance evaluates the single-precision build's formula (|x| exact) in double, which is the
    // tolerance target for forces: the float-rounded |x| of the double build multiplies the DC term of the
    // inverse mesh and perturbs ideal-gas-like forces at the 1e-2 level (tests/test_oracle.py demonstrates it).
    bool literal_copysignf = true;
    S Wd(S x) const {
        S xsq = x * x;
        S xabs = literal_copysignf ? (S)copysignf((float)x, 1.0f) : std::fabs(x);
        S fac = (S(3.0 / 2.0) - xabs);
        S ret(0.0);
        if (xsq <= S(1.0 / 4.0)) ret = -S(2.0) * x;
        else if (xsq <= S(9.0 / 4.0)) ret = -fac * x / xabs;
        return ret;
    }
    // assignTSCfourier: OrderParameterMesh.cc:487-511
    static S Wk(S x) {
        const S c[] = {S(1.0), S(-1.0 / 6.0), S(1.0 / 120.0), S(-1.0 / 5040.0), S(1.0 / 362880.0), S(-1.0 / 39916800.0)};
        S sinc = 0;
        if (x * x <= S(1.0)) {
            S term = S(1.0);
            for (unsigned i = 0; i < 6; ++i) { sinc += c[i] * term; term *= x * x; }
        } else sinc = std::sin(x) / x;
        return sinc * sinc * sinc;
    }

    // computeInfluenceFunction, the part that feeds CV/forces: OrderParameterMesh.cc:388-450.
    // Quirk restated literally: n.x/global_dim.x is int/unsigned => unsigned integer division
    // (:448), so interp == 1 for non-negative Miller indices and ~1e-27 otherwise.
    void compute_interpolation() {
        interp.assign(M, S(0));
        for (unsigned cell = 0; cell < M; ++cell) {
            unsigned wz = cell / (ny * nx);
            unsigned wy = (cell - wz * nx * ny) / nx;
            unsigned wx = cell % nx;
            int n_x = (int)wx, n_y = (int)wy, n_z = (int)wz;
            if (n_x >= (int)(nx / 2 + nx % 2)) n_x -= (int)nx;
            if (n_y >= (int)(ny / 2 + ny % 2)) n_y -= (int)ny;
            if (n_z >= (int)(nz / 2 + nz % 2)) n_z -= (int)nz;
            unsigned qx = (unsigned)n_x / nx, qy = (unsigned)n_y / ny, qz = (unsigned)n_z / nz;
            S kHx = S(M_PI * 2.0) * (S)qx, kHy = S(M_PI * 2.0) * (S)qy, kHz = S(M_PI * 2.0) * (S)qz;
            interp[cell] = Wk(kHx) * Wk(kHy) * Wk(kHz);
        }
    }

    // cell index + in-cell shift shared by assignParticles (:543-573) and interpolateForces (:783-810)
    inline void locate(V3<S> pos, int& ix, int& iy, int& iz, V3<S>& shift) const {
        V3<S> f = box.makeFraction(pos);
        V3<S> r = v3<S>(f.x * (S)nx, f.y * (S)ny, f.z * (S)nz);
        ix = (int)r.x; iy = (int)r.y; iz = (int)r.z;
        if (ix == (int)nx) ix = 0;
        if (iy == (int)ny) iy = 0;
        if (iz == (int)nz) iz = 0;
        V3<S> center = v3<S>((S)ix + S(0.5), (S)iy + S(0.5), (S)iz + S(0.5));
        V3<S> dims = v3<S>((S)nx, (S)ny, (S)nz);
        V3<S> c_cart = box.makeCoordinates(center / dims);
        V3<S> shift_cart = box.minImage(pos - c_cart);
        // As written (:571-573, :806-808): makeFraction shears its whole argument, `lo` included, so in a triclinic box the
        // offset carries the constant -shear(lo)/L (x: ((xz - yz xy) Lz + xy Ly)/(2 Lx), y: yz Lz/(2 Ly), times the mesh
        // dimensions) and TSC weight is lost where |offset - tap| > 3/2.  Restated literally by default (pinned to the
        // reference's own output, vectors t0-t2); literal_tilt_offset = false removes the constant.
        V3<S> shift_f = box.makeFraction(shift_cart + box.lo);
        if (!literal_tilt_offset) {
            shift_f.x += ((box.xz - box.yz * box.xy) * box.lo.z + box.xy * box.lo.y) * box.Linv.x;
            shift_f.y += box.yz * box.lo.z * box.Linv.y;
        }
        shift = shift_f * dims;
    }
    bool literal_tilt_offset = true;
    static inline int wrap(int i, int n) { if (i == n) return 0; if (i < 0) return i + n; return i; }

    // assignParticles: OrderParameterMesh.cc:517-640
    void assign(const float* postype, unsigned N) {
        std::fill(mesh.begin(), mesh.end(), std::complex<S>(0, 0));
        cells.resize((size_t)3 * N);
        mode_sq = S(0.0);
        for (unsigned p = 0; p < N; ++p) {
            const float* pt = postype + 4 * (size_t)p;
            V3<S> pos = v3<S>((S)pt[0], (S)pt[1], (S)pt[2]);
            unsigned type = type_of(pt);
            int ix, iy, iz; V3<S> s;
            locate(pos, ix, iy, iz, s);
            cells[3 * (size_t)p] = ix; cells[3 * (size_t)p + 1] = iy; cells[3 * (size_t)p + 2] = iz;
            for (int i = -1; i <= 1; ++i)
                for (int j = -1; j <= 1; ++j)
                    for (int k = -1; k <= 1; ++k) {
                        int ni = wrap(ix + i, (int)nx), nj = wrap(iy + j, (int)ny), nk = wrap(iz + k, (int)nz);
                        V3<S> d = s - v3<S>((S)i, (S)j, (S)k);
                        S frac = W(d.x) * W(d.y) * W(d.z);
                        size_t idx = (size_t)ni + (size_t)nx * ((size_t)nj + (size_t)ny * nk);
                        mesh[idx] += std::complex<S>(mode[type] * frac, 0);
                    }
            mode_sq += mode[type] * mode[type];
        }
    }

    // updateMeshes (non-MPI branch): OrderParameterMesh.cc:642-747
    void update_meshes() {
        fft3d(mesh, fourier, nx, ny, nz, -1);
        for (unsigned k = 0; k < M; ++k) {
            S fr = fourier[k].real(), fi = fourier[k].imag();
            fr /= (S)N_global; fi /= (S)N_global;
            S val = fr * fr + fi * fi;
            S gr = fr * val, gi = fi * val;
            S diag = S(0.5) * interp[k] * interp[k] * mode_sq / (S)N_global / (S)N_global;
            gr -= fr * diag; gi -= fi * diag;
            fourier_G[k] = {gr, gi};
            fourier[k] = {fr, fi};
        }
        fft3d(fourier_G, inv, nx, ny, nz, +1);
    }

    // computeCV: OrderParameterMesh.cc:866-923 (exclude flat index 0 only)
    S compute_cv() {
        S sum(0.0);
        for (unsigned k = 1; k < M; ++k) {
            sum += fourier_G[k].real() * fourier[k].real() + fourier_G[k].imag() * fourier[k].imag();
            S norm2 = fourier[k].real() * fourier[k].real() + fourier[k].imag() * fourier[k].imag();
            S diag = S(0.5) * norm2 * interp[k] * interp[k] * mode_sq / (S)N_global / (S)N_global;
            sum -= diag;
        }
        sum *= S(1.0 / 2.0);
        return sum;
    }

    // getCurrentValue without the per-timestep cache: OrderParameterMesh.cc:925-968
    S current_value(const float* postype, unsigned N) {
        assign(postype, N);
        update_meshes();
        cv = compute_cv();
        return cv;
    }

    // interpolateForces: OrderParameterMesh.cc:749-864 (force.w = 0)
    void forces(const float* postype, unsigned N, S bias, S* out4) const {
        V3<S> b1, b2, b3;
        reciprocal(box, false, b1, b2, b3);
        for (unsigned p = 0; p < N; ++p) {
            const float* pt = postype + 4 * (size_t)p;
            V3<S> pos = v3<S>((S)pt[0], (S)pt[1], (S)pt[2]);
            S a = mode[type_of(pt)];
            int ix, iy, iz; V3<S> s;
            locate(pos, ix, iy, iz, s);
            V3<S> force = v3<S>(0, 0, 0);
            for (int i = -1; i <= 1; ++i)
                for (int j = -1; j <= 1; ++j)
                    for (int k = -1; k <= 1; ++k) {
                        int ni = wrap(ix + i, (int)nx), nj = wrap(iy + j, (int)ny), nk = wrap(iz + k, (int)nz);
                        V3<S> d = s - v3<S>((S)i, (S)j, (S)k);
                        size_t idx = (size_t)ni + (size_t)nx * ((size_t)nj + (size_t)ny * nk);
                        S re = inv[idx].real();
                        force = force + (-(S)nx) * b1 * a * Wd(d.x) * W(d.y) * W(d.z) * re;
                        force = force + (-(S)ny) * b2 * a * W(d.x) * Wd(d.y) * W(d.z) * re;
                        force = force + (-(S)nz) * b3 * a * W(d.x) * W(d.y) * Wd(d.z) * re;
                    }
            force = force * (S(2.0) / (S)N_global * bias);
            out4[4 * (size_t)p] = force.x; out4[4 * (size_t)p + 1] = force.y;
            out4[4 * (size_t)p + 2] = force.z; out4[4 * (size_t)p + 3] = S(0.0);
        }
    }

    // computeQmax: OrderParameterMesh.cc:1108-1179 (log quantities q*_max, sq_max)
    // computeVirial: OrderParameterMesh.cc:970-1050.  k-space virial of the bias: only the derivative table of the
    // convolution kernel enters (val_D = 0 without a table, or outside [k_min, k_max)); `fourier` is f = F/N here (it is
    // divided by N once more, twice, as written); flat index 0 excluded; result = bias * sum.
    void virial(const std::vector<S>& table_d, S k_min, S k_max, bool use_table, S bias, S out6[6]) const {
        V3<S> b1, b2, b3;
        reciprocal(box, true, b1, b2, b3);
        const S delta_k = (k_max - k_min) / S(table_d.size() - 1);
        S vir[6] = {0, 0, 0, 0, 0, 0};
        for (unsigned cell = 1; cell < M; ++cell) {
            unsigned wz = cell / (ny * nx), wy = (cell - wz * nx * ny) / nx, wx = cell % nx;
            int n_x = (int)wx, n_y = (int)wy, n_z = (int)wz;
            if (n_x >= (int)(nx / 2 + nx % 2)) n_x -= (int)nx;
            if (n_y >= (int)(ny / 2 + ny % 2)) n_y -= (int)ny;
            if (n_z >= (int)(nz / 2 + nz % 2)) n_z -= (int)nz;
            const V3<S> k = (S)n_x * b1 + (S)n_y * b2 + (S)n_z * b3;
            const S ksq = dot(k, k), knorm = std::sqrt(ksq);
            S kfac = S(1.0) / S(2.0) / knorm;
            S val_D(0.0);
            if (use_table && knorm >= k_min && knorm < k_max) {
                const S value_f = (knorm - k_min) / delta_k;
                const unsigned value_i = (unsigned)value_f;
                const S dK0 = table_d[value_i], dK1 = table_d[value_i + 1];
                const S f = value_f - S(value_i);
                val_D = dK0 + f * (dK1 - dK0);
            }
            kfac *= val_D;
            const S a2 = fourier[cell].real() * fourier[cell].real() + fourier[cell].imag() * fourier[cell].imag();
            const S val = a2 / (S)N_global;
            const S rhog = a2 * val / (S)N_global;
            vir[0] += rhog * kfac * k.x * k.x; vir[1] += rhog * kfac * k.x * k.y; vir[2] += rhog * kfac * k.x * k.z;
            vir[3] += rhog * kfac * k.y * k.y; vir[4] += rhog * kfac * k.y * k.z; vir[5] += rhog * kfac * k.z * k.z;
        }
        for (int i = 0; i < 6; ++i) out6[i] = bias * vir[i];
    }
    void qmax(S out4[4]) const {
        V3<S> b1, b2, b3;
        reciprocal(box, true, b1, b2, b3);
        S max_amp(0.0); V3<S> q = v3<S>(0, 0, 0);
        for (unsigned cell = 0; cell < M; ++cell) {
            S a = fourier[cell].real() * fourier[cell].real() + fourier[cell].imag() * fourier[cell].imag();
            if (a > max_amp) {
                unsigned wz = cell / (ny * nx), wy = (cell - wz * nx * ny) / nx, wx = cell % nx;
                int n_x = (int)wx, n_y = (int)wy, n_z = (int)wz;
                if (n_x >= (int)(nx / 2 + nx % 2)) n_x -= (int)nx;
                if (n_y >= (int)(ny / 2 + ny % 2)) n_y -= (int)ny;
                if (n_z >= (int)(nz / 2 + nz % 2)) n_z -= (int)nz;
                q = (S)n_x * b1 + (S)n_y * b2 + (S)n_z * b3;
                max_amp = a;
            }
        }
        out4[0] = q.x; out4[1] = q.y; out4[2] = q.z; out4[3] = max_amp * (S)N_global;
    }
};

// ----------------------------------------------------------------------------------------
// LamellarOrderParameter (CPU path)
// ----------------------------------------------------------------------------------------
template <class S> struct Lamellar {
    std::vector<S> mode;
    std::vector<int> lattice;     // 3 ints per wave vector
    Box<S> box;                   // global box
    unsigned N_global = 0;
    std::vector<S> fourier;       // (re, im) per wave vector
    S cv = 0;

    V3<S> q_of(unsigned k) const {
        V3<S> b1, b2, b3;
        reciprocal(box, true, b1, b2, b3);
        return b1 * (S)lattice[3 * k] + b2 * (S)lattice[3 * k + 1] + b3 * (S)lattice[3 * k + 2];
    }
    // calculateFourierModes: LamellarOrderParameter.cc:143-179
    void fourier_modes(const float* postype, unsigned N) {
        unsigned nw = (unsigned)lattice.size() / 3;
        fourier.assign(2 * nw, S(0));
        for (unsigned k = 0; k < nw; ++k) {
            V3<S> q = q_of(k);
            for (unsigned p = 0; p < N; ++p) {
                const float* pt = postype + 4 * (size_t)p;
                V3<S> pos = v3<S>((S)pt[0], (S)pt[1], (S)pt[2]);
                S a = mode[type_of(pt)];
                S dp = dot(q, pos);
                fourier[2 * k] += a * std::cos(dp);
                fourier[2 * k + 1] += a * std::sin(dp);
            }
        }
    }
    // computeCV: LamellarOrderParameter.cc:42-74
    S compute_cv(const float* postype, unsigned N) {
        fourier_modes(postype, N);
        S sum = 0.0;
        for (size_t k = 0; k < fourier.size() / 2; ++k) sum += fourier[2 * k];
        sum /= (S)N_global;
        cv = sum;
        return cv;
    }
    // computeBiasForces: LamellarOrderParameter.cc:77-140 (factor 2 is the reference's, :120)
    void forces(const float* postype, unsigned N, S bias, S* out4) const {
        unsigned nw = (unsigned)lattice.size() / 3;
        S denom = (S)N_global;
        std::vector<V3<S>> qs(nw);
        for (unsigned k = 0; k < nw; ++k) qs[k] = q_of(k);
        for (unsigned p = 0; p < N; ++p) {
            const float* pt = postype + 4 * (size_t)p;
            V3<S> pos = v3<S>((S)pt[0], (S)pt[1], (S)pt[2]);
            S a = mode[type_of(pt)];
            S fx = 0, fy = 0, fz = 0;
            for (unsigned k = 0; k < nw; ++k) {
                S dp = dot(pos, qs[k]);
                S f = S(2.0) * a * std::sin(dp);
                fx += qs[k].x * f; fy += qs[k].y * f; fz += qs[k].z * f;
            }
            fx *= bias; fy *= bias; fz *= bias;
            fx /= denom; fy /= denom; fz /= denom;
            out4[4 * (size_t)p] = fx; out4[4 * (size_t)p + 1] = fy; out4[4 * (size_t)p + 2] = fz;
            out4[4 * (size_t)p + 3] = S(0);
        }
    }
};

// ----------------------------------------------------------------------------------------
// CollectiveVariable umbrella logic: CollectiveVariable.cc:22-66 (bias increment applied in
// computeForces before computeBiasForces) and :68-106 (getUmbrellaPotential)
// ----------------------------------------------------------------------------------------
enum Umbrella { no_umbrella = 0, linear = 1, harmonic = 2, wall = 3, gaussian = 4 };

template <class S> struct UmbrellaParams { int kind = no_umbrella; S cv0 = 0, kappa = 1, width_flat = 0, scale = 1; };

template <class S> inline S umbrella_bias(const UmbrellaParams<S>& u, S val, S bias_in) {
    if (u.kind == no_umbrella) return bias_in;
    if ((val < u.cv0 + u.width_flat / S(2.0)) && (val > u.cv0 - u.width_flat / S(2.0))) return bias_in;
    S delta(0.0);
    if (val > u.cv0) delta = val - u.cv0 - u.width_flat / S(2.0);
    else delta = val - u.cv0 + u.width_flat / S(2.0);
    if (u.kind == linear) return bias_in + u.scale * S(1.0);
    if (u.kind == harmonic) return bias_in + u.kappa * delta;
    if (u.kind == wall) return bias_in + u.scale * S(12.0) * std::pow(delta / u.kappa, S(11.0)) / u.kappa;
    if (u.kind == gaussian)
        return bias_in - u.scale * (val - u.cv0) * std::exp(-(val - u.cv0) * (val - u.cv0) / u.kappa / u.kappa / S(2.0));
    return bias_in;
}
template <class S> inline S umbrella_potential(const UmbrellaParams<S>& u, S val) {
    if (u.kind == no_umbrella) return S(0.0);
    if ((val < u.cv0 + u.width_flat / S(2.0)) && (val > u.cv0 - u.width_flat / S(2.0))) return S(0.0);
    S delta(0.0);
    if (val > u.cv0) delta = val - u.cv0 - u.width_flat / S(2.0);
    else if (val < u.cv0) delta = val - u.cv0 + u.width_flat / S(2.0);
    if (u.kind == linear) return u.scale * delta;
    if (u.kind == harmonic) return S(1.0 / 2.0) * delta * delta * u.kappa;
    if (u.kind == wall) return u.scale * std::pow(delta / u.kappa, S(12.0));
    if (u.kind == gaussian)
        return u.scale * std::exp(-(val - u.cv0) * (val - u.cv0) / u.kappa / u.kappa / S(2.0)) - u.scale;
    return S(0.0);
}

// ----------------------------------------------------------------------------------------
// IndexGrid: IndexGrid.cc:20-67 (first CV fastest-varying)
// ----------------------------------------------------------------------------------------
struct IndexGrid {
    std::vector<unsigned> lengths, factors;
    void setLengths(const std::vector<unsigned>& l) {
        lengths = l; factors.resize(l.size());
        for (size_t i = 0; i < l.size(); ++i) factors[i] = (i == 0) ? 1 : lengths[i - 1] * factors[i - 1];
    }
    unsigned getIndex(const std::vector<unsigned>& c) const {
        unsigned idx = 0;
        for (size_t i = 0; i < lengths.size(); ++i) idx += c[i] * factors[i];
        return idx;
    }
    void getCoordinates(unsigned idx, std::vector<unsigned>& c) const {
        unsigned rest = idx;
        for (int i = (int)lengths.size() - 1; i >= 0; --i) { c[i] = rest / factors[i]; rest -= c[i] * factors[i]; }
    }
    unsigned getNumElements() const { unsigned r = 1; for (unsigned l : lengths) r *= l; return r; }
    unsigned dim() const { return (unsigned)lengths.size(); }
};

// ----------------------------------------------------------------------------------------
// IntegratorMetaDynamics, grid mode (the only mode the Python API enables, integrate.py:266-267)
// ----------------------------------------------------------------------------------------
template <class S> struct MetaGrid {
    struct Var { S sigma, cv_min, cv_max; unsigned num_points; std::string name; };
    std::vector<Var> vars;
    S W = 1, T_shift = 1, temp = 1;
    unsigned stride = 1;
    bool add_bias = true, well_tempered = false;
    unsigned num_gaussians = 0;
    S curr_bias_potential = 0, curr_reweight = 1;
    IndexGrid gi;
    std::vector<S> grid, grid_delta, grid_reweighted, grid_weight, sigma_grid, sigma_grid_delta, sigma_inv;
    std::vector<unsigned> hist, hist_delta, hist_gauss, hist_gauss_delta;
    unsigned n_out_of_bounds = 0;   // count of "out of bounds" warnings (:677-683)

    // prepRun :142-182 + setupGrid :590-661
    void setup() {
        size_t d = vars.size();
        sigma_inv.assign(d * d, S(0));
        std::vector<unsigned> len(d);
        for (size_t i = 0; i < d; ++i) { sigma_inv[i * d + i] = S(1.0) / vars[i].sigma; len[i] = vars[i].num_points; }
        gi.setLengths(len);
        size_t G = gi.getNumElements();
        grid.assign(G, 0); grid_delta.assign(G, 0); grid_reweighted.assign(G, 0); grid_weight.assign(G, S(1.0));
        sigma_grid.assign(G, 0); sigma_grid_delta.assign(G, 0);
        hist.assign(G, 0); hist_delta.assign(G, 0); hist_gauss.assign(G, 0); hist_gauss_delta.assign(G, 0);
    }
    S delta_of(size_t i) const { return (vars[i].cv_max - vars[i].cv_min) / (S)(vars[i].num_points - 1); }

    // interpolateGrid: IntegratorMetaDynamics.cc:663-736
    S interpolate(const std::vector<S>& val, bool reweight) {
        size_t d = vars.size();
        std::vector<unsigned> lower_idx(d), upper_idx(d);
        std::vector<S> rel(d);
        for (size_t i = 0; i < d; ++i) {
            S delta = (vars[i].cv_max - vars[i].cv_min) / (vars[i].num_points - 1);
            if (val[i] < vars[i].cv_min || val[i] >= vars[i].cv_max) { ++n_out_of_bounds; return S(0.0); }
            int lower = (int)((val[i] - vars[i].cv_min) / delta);
            int upper = lower + 1;
            if (upper >= (int)vars[i].num_points) { lower--; upper--; }
            S lower_bound = vars[i].cv_min + delta * lower;
            S upper_bound = vars[i].cv_min + delta * upper;
            lower_idx[i] = lower; upper_idx[i] = upper;
            rel[i] = (val[i] - lower_bound) / (upper_bound - lower_bound);
        }
        unsigned n_term = 1u << d;
        S res(0.0);
        std::vector<unsigned> coords(d);
        for (unsigned bits = 0; bits < n_term; ++bits) {
            S term(1.0);
            for (size_t i = 0; i < d; ++i) {
                if (bits & (1u << i)) { coords[i] = lower_idx[i]; term *= (S(1.0) - rel[i]); }
                else { coords[i] = upper_idx[i]; term *= rel[i]; }
            }
            unsigned idx = gi.getIndex(coords);
            term *= (reweight ? grid_weight[idx] : grid[idx]);
            res += term;
        }
        return res;
    }
    // biasPotentialDerivative: IntegratorMetaDynamics.cc:738-776
    S derivative(unsigned cv, const std::vector<S>& val) {
        S delta = (vars[cv].cv_max - vars[cv].cv_min) / (S)(vars[cv].num_points - 1);
        if (val[cv] - delta < vars[cv].cv_min) {
            std::vector<S> v2 = val; v2[cv] += delta;
            S y2 = interpolate(v2, false), y1 = interpolate(val, false);
            return (y2 - y1) / delta;
        } else if (val[cv] + delta > vars[cv].cv_max) {
            std::vector<S> v2 = val; v2[cv] -= delta;
            S y1 = interpolate(v2, false), y2 = interpolate(val, false);
            return (y2 - y1) / delta;
        }
        std::vector<S> v1 = val, v2 = val;
        v1[cv] -= delta; v2[cv] += delta;
        S y1 = interpolate(v1, false), y2 = interpolate(v2, false);
        return (y2 - y1) / (S(2.0) * delta);
    }
    // histogram bin shared by updateHistogram :1092-1119 and updateSigmaGrid :1122-1155
    // Scalar -> unsigned conversion.  For a negative quotient the conversion is undefined behaviour in C++; the reference
    // as compiled for x86-64 (cvttsd2si to 64 bit, low word kept) gives bin 0 for -1 < q < 0 and a huge value, i.e.
    // off-grid, for q <= -1.  That is what the reference's own code does here (tests/test_reference_build.py) and what is
    // restated: a CV value less than one grid spacing below cv_min is counted in the first bin.
    bool bin_of(const std::vector<S>& val, unsigned& idx) const {
        size_t d = vars.size();
        std::vector<unsigned> c(d);
        bool on = true;
        for (size_t i = 0; i < d; ++i) {
            S delta = (vars[i].cv_max - vars[i].cv_min) / (vars[i].num_points - 1);
            S q = (val[i] - vars[i].cv_min) / delta;
            if (!(q > S(-1.0)) || q >= S(4294967296.0)) { on = false; c[i] = 0; continue; }
            c[i] = q < S(0) ? 0u : (unsigned)q;
            if (c[i] >= vars[i].num_points) on = false;
        }
        if (on) idx = gi.getIndex(c);
        return on;
    }
    // sigmaDeterminant: IntegratorMetaDynamics.cc:1296-1313 (Eigen determinant -> plain elimination)
    S sigma_det() const {
        size_t d = vars.size();
        std::vector<S> m(sigma_inv);
        S det = 1;
        for (size_t c = 0; c < d; ++c) {
            size_t piv = c;
            for (size_t r = c + 1; r < d; ++r) if (std::fabs(m[r * d + c]) > std::fabs(m[piv * d + c])) piv = r;
            if (m[piv * d + c] == S(0)) return S(0);
            if (piv != c) { for (size_t k = 0; k < d; ++k) std::swap(m[piv * d + k], m[c * d + k]); det = -det; }
            det *= m[c * d + c];
            for (size_t r = c + 1; r < d; ++r) {
                S f = m[r * d + c] / m[c * d + c];
                for (size_t k = c; k < d; ++k) m[r * d + k] -= f * m[c * d + k];
            }
        }
        return det;
    }
    // updateGrid (CPU): IntegratorMetaDynamics.cc:1002-1047 -- d_i and gauss are double even in the float build
    void deposit(const std::vector<S>& cur, S scal) {
        size_t d = vars.size();
        unsigned len = gi.getNumElements();
        std::vector<unsigned> coords(d);
        for (unsigned g = 0; g < len; ++g) {
            gi.getCoordinates(g, coords);
            S gauss_exp(0.0);
            for (size_t i = 0; i < d; ++i) {
                S delta_i = (vars[i].cv_max - vars[i].cv_min) / (vars[i].num_points - 1);
                S val_i = vars[i].cv_min + coords[i] * delta_i;
                double d_i = val_i - cur[i];
                for (size_t j = 0; j < d; ++j) {
                    S delta_j = (vars[j].cv_max - vars[j].cv_min) / (vars[j].num_points - 1);
                    S val_j = vars[j].cv_min + coords[j] * delta_j;
                    double d_j = val_j - cur[j];
                    S sij = sigma_inv[i * d + j];
                    gauss_exp += d_i * d_j * S(1.0 / 2.0) * (sij * sij);
                }
            }
            double gauss = std::exp(-gauss_exp);
            grid_delta[g] = W * scal * gauss;
        }
    }
    // updateReweightedEstimator: IntegratorMetaDynamics.cc:1053-1090
    void reweight() {
        unsigned len = gi.getNumElements();
        S avg(0.0), norm(0.0);
        for (unsigned g = 0; g < len; ++g) {
            grid_reweighted[g] += (S)hist_delta[g];
            avg += grid_reweighted[g] * grid_delta[g];
            norm += grid_reweighted[g];
        }
        avg /= norm;
        for (unsigned g = 0; g < len; ++g) {
            double dV = grid_delta[g];
            S fac = std::exp(-(dV - avg) / temp);
            grid_reweighted[g] *= fac;
            grid_weight[g] /= fac;
        }
    }
    // computeSigma: IntegratorMetaDynamics.cc:1205-1294 (adaptive Gaussians).  force[i] = derivative array of CV i as left in
    // its force array by computeDerivatives (Scalar4 per particle, bias factor 1), or null if the CV cannot compute
    // derivatives (then its diagonal entry is sigma_i^2 and its off-diagonal entries are 0).  sigma_inv = inverse of the
    // matrix of element-wise square roots -- a negative sum of products gives NaN, as in the reference.  Accumulation in
    // Scalar, particle order.  (Eigen's inverse -> Gauss-Jordan with partial pivoting.)
    void compute_sigma(const std::vector<const float*>& force, unsigned N, S sigma_g) {
        const size_t d = vars.size();
        std::vector<S> sigmasq(d * d, S(0));
        for (size_t i = 0; i < d; ++i)
            for (size_t j = 0; j < d; ++j) {
                if (force[i] && force[j]) {
                    for (unsigned n = 0; n < N; ++n) {
                        const float* fi = force[i] + 4 * (size_t)n; const float* fj = force[j] + 4 * (size_t)n;
                        sigmasq[i * d + j] += sigma_g * sigma_g * ((S)fi[0] * (S)fj[0] + (S)fi[1] * (S)fj[1] + (S)fi[2] * (S)fj[2]);
                    }
                } else if (i == j) sigmasq[i * d + j] = vars[i].sigma * vars[i].sigma;
            }
        std::vector<S> m(d * d), inv(d * d, S(0));
        for (size_t k = 0; k < d * d; ++k) m[k] = std::sqrt(sigmasq[k]);
        for (size_t i = 0; i < d; ++i) inv[i * d + i] = S(1);
        for (size_t c = 0; c < d; ++c) {
            size_t piv = c;
            for (size_t r = c + 1; r < d; ++r) if (std::fabs(m[r * d + c]) > std::fabs(m[piv * d + c])) piv = r;
            if (piv != c) for (size_t k = 0; k < d; ++k) { std::swap(m[piv * d + k], m[c * d + k]); std::swap(inv[piv * d + k], inv[c * d + k]); }
            const S p = m[c * d + c];
            for (size_t k = 0; k < d; ++k) { m[c * d + k] /= p; inv[c * d + k] /= p; }
            for (size_t r = 0; r < d; ++r) {
                if (r == c) continue;
                const S f = m[r * d + c];
                for (size_t k = 0; k < d; ++k) { m[r * d + k] -= f * m[c * d + k]; inv[r * d + k] -= f * inv[c * d + k]; }
            }
        }
        sigma_inv = inv;
    }
    // updateBiasPotential, grid branch, root rank, no file output: IntegratorMetaDynamics.cc:314-588.  The two halves on
    // either side of the multiple-walker all-reduce of the four delta arrays (:392-410) are separate functions so that
    // tests can put the sum over walkers in between; update() = both, one walker.
    void update_deposit(unsigned timestep, const std::vector<S>& cur) {
        unsigned idx;
        if (bin_of(cur, idx)) hist_delta[idx]++;                       // updateHistogram
        if (add_bias && (timestep % stride == 0)) {
            if (bin_of(cur, idx)) { sigma_grid_delta[idx] += sigma_det(); hist_gauss_delta[idx]++; }   // updateSigmaGrid
            S scal = S(1.0);
            if (well_tempered) { S V = interpolate(cur, false); scal = std::exp(-V / T_shift); }
            deposit(cur, scal);
        }
    }
    void update_merge(unsigned timestep, const std::vector<S>& cur, std::vector<S>& bias) {
        size_t d = vars.size();
        bias.assign(d, S(0.0));
        if (add_bias && (timestep % stride == 0)) {
            reweight();
            for (size_t g = 0; g < grid.size(); ++g) {
                grid[g] += grid_delta[g]; sigma_grid[g] += sigma_grid_delta[g];
                hist[g] += hist_delta[g]; hist_gauss[g] += hist_gauss_delta[g];
                grid_delta[g] = S(0.0); sigma_grid_delta[g] = S(0.0); hist_delta[g] = 0; hist_gauss_delta[g] = 0;
            }
            num_gaussians++;
        }
        for (unsigned i = 0; i < d; ++i) bias[i] = derivative(i, cur);
        curr_bias_potential = interpolate(cur, false);
        curr_reweight = interpolate(cur, true);
    }
    void update(unsigned timestep, const std::vector<S>& cur, std::vector<S>& bias) {
        size_t d = vars.size();
        bias.assign(d, S(0.0));
        unsigned idx;
        if (bin_of(cur, idx)) hist_delta[idx]++;                       // updateHistogram
        if (add_bias && (timestep % stride == 0)) {
            if (bin_of(cur, idx)) { sigma_grid_delta[idx] += sigma_det(); hist_gauss_delta[idx]++; }   // updateSigmaGrid
            S scal = S(1.0);
            if (well_tempered) { S V = interpolate(cur, false); scal = std::exp(-V / T_shift); }
            deposit(cur, scal);
            reweight();
            for (size_t g = 0; g < grid.size(); ++g) {
                grid[g] += grid_delta[g]; sigma_grid[g] += sigma_grid_delta[g];
                hist[g] += hist_delta[g]; hist_gauss[g] += hist_gauss_delta[g];
                grid_delta[g] = S(0.0); sigma_grid_delta[g] = S(0.0); hist_delta[g] = 0; hist_gauss_delta[g] = 0;
            }
            num_gaussians++;
        }
        for (unsigned i = 0; i < d; ++i) bias[i] = derivative(i, cur);
        curr_bias_potential = interpolate(cur, false);
        curr_reweight = interpolate(cur, true);
    }
    // resetHistogram: IntegratorMetaDynamics.cc:1195-1203
    void reset_histogram() { std::fill(hist.begin(), hist.end(), 0u); std::fill(hist_delta.begin(), hist_delta.end(), 0u); }

    // writeGrid: IntegratorMetaDynamics.cc:831-926 (file name = filename + "_" + timestep)
    void write_grid(const std::string& filename, unsigned timestep) const {
        std::ofstream file((filename + "_" + std::to_string(timestep)).c_str(), std::ios_base::out);
        file << "#n_cv: " << gi.dim() << std::endl;
        file << "#dim: ";
        for (unsigned i = 0; i < gi.dim(); i++) file << " " << gi.lengths[i];
        file << std::endl;
        file << "#num_gaussians: " << num_gaussians << std::endl;
        for (size_t i = 0; i < vars.size(); i++) file << vars[i].name << "\t";
        file << "grid_value" << "\t" << "det_sigma" << "\t" << "num_gaussians" << "\t" << "hist" << "\t"
             << "hist_reweight" << "\t" << "weight" << std::endl;
        unsigned len = gi.getNumElements();
        std::vector<unsigned> coords(gi.dim());
        for (unsigned g = 0; g < len; g++) {
            gi.getCoordinates(g, coords);
            for (size_t i = 0; i < vars.size(); ++i) {
                S delta = (vars[i].cv_max - vars[i].cv_min) / (vars[i].num_points - 1);
                S val = vars[i].cv_min + coords[i] * delta;
                file << std::setprecision(10) << val << "\t";
            }
            file << std::setprecision(10) << grid[g];
            S val = hist_gauss[g] > 0 ? sigma_grid[g] / (S)hist_gauss[g] : S(0.0);
            file << "\t" << std::setprecision(10) << val;
            file << "\t" << hist_gauss[g] << "\t" << hist[g];
            file << "\t" << std::setprecision(10) << grid_reweighted[g];
            file << "\t" << std::setprecision(10) << grid_weight[g] << std::endl;
        }
    }
    // readGrid: IntegratorMetaDynamics.cc:928-1000
    void read_grid(const std::string& filename) {
        std::ifstream file(filename.c_str());
        if (!file.good()) throw std::runtime_error("Error reading grid.");
        std::string line, tmp;
        std::getline(file, line); std::getline(file, line);
        std::getline(file, line);
        { std::istringstream iss(line); iss >> tmp >> num_gaussians; }
        std::getline(file, line);
        unsigned len = gi.getNumElements();
        for (unsigned g = 0; g < len; g++) {
            if (!file.good()) throw std::runtime_error("Error reading grid.");
            std::getline(file, line);
            std::istringstream iss(line);
            for (size_t i = 0; i < vars.size(); i++) iss >> tmp;
            iss >> grid[g] >> sigma_grid[g] >> hist_gauss[g] >> hist[g];
            sigma_grid[g] *= hist_gauss[g];
            iss >> grid_reweighted[g] >> grid_weight[g];
        }
    }
};

// ----------------------------------------------------------------------------------------
// WellTemperedEnsemble (CPU path): WellTemperedEnsemble.cc:30-68 (pe), :135-188 (scale)
// ----------------------------------------------------------------------------------------
template <class S> inline S wte_potential_energy(const float* net_force4, unsigned N, S external_energy) {
    S pe(0.0);
    for (unsigned i = 0; i < N; ++i) pe += (S)net_force4[4 * (size_t)i + 3];
    pe += external_energy;
    return pe;
}
// scales net force xyz, net torque xyzw, six virial rows (pitch-strided) by fac = 1 + bias
template <class S>
inline void wte_scale(S* net_force4, S* net_torque4, S* net_virial, unsigned pitch, unsigned N, S bias, S ext_virial[6]) {
    S fac = S(1.0) + bias;
    for (unsigned i = 0; i < N; ++i) {
        net_force4[4 * (size_t)i] *= fac; net_force4[4 * (size_t)i + 1] *= fac; net_force4[4 * (size_t)i + 2] *= fac;
        net_torque4[4 * (size_t)i] *= fac; net_torque4[4 * (size_t)i + 1] *= fac;
        net_torque4[4 * (size_t)i + 2] *= fac; net_torque4[4 * (size_t)i + 3] *= fac;
        for (unsigned r = 0; r < 6; ++r) net_virial[i + (size_t)r * pitch] *= fac;
    }
    for (unsigned r = 0; r < 6; ++r) ext_virial[r] = fac * ext_virial[r];
}

// ----------------------------------------------------------------------------------------
// AspectRatio: AspectRatio.cc:24-57 (value; dir2==0 assigns length1 -- the reference's bug,
// restated), :59-130 (external virial)
// ----------------------------------------------------------------------------------------
template <class S> inline S aspect_ratio_value(const Box<S>& box, unsigned dir1, unsigned dir2) {
    S L[3] = {box.L.x, box.L.y, box.L.z};
    S length1(0.0), length2(0.0);
    length1 = L[dir1];
    if (dir2 == 0) length1 = L[0];
    else length2 = L[dir2];
    return length1 / length2;
}
template <class S> inline void aspect_ratio_virial(const Box<S>& box, unsigned dir1, unsigned dir2, S bias, S out[6]) {
    S Lx = box.L.x, Ly = box.L.y, Lz = box.L.z;
    S dx(0.0), dy(0.0), dz(0.0);
    if (dir1 == 0 && dir2 == 1) { dx = S(1.0) / Ly; dy = -Lx / Ly / Ly; }
    else if (dir1 == 0 && dir2 == 2) { dx = S(1.0) / Lz; dz = -Lx / Lz / Lz; }
    else if (dir1 == 1 && dir2 == 0) { dx = -Ly / Lx / Lx; dy = S(1.0) / Lx; }
    else if (dir1 == 1 && dir2 == 2) { dy = S(1.0) / Lz; dz = -Ly / Lz / Lz; }
    else if (dir1 == 2 && dir2 == 0) { dx = -Lz / Lx / Lx; dz = S(1.0) / Lx; }
    else if (dir1 == 2 && dir2 == 1) { dy = -Lz / Ly / Ly; dz = S(1.0) / Ly; }
    out[0] = -bias * dx * Lx;
    out[1] = -bias * dx * (Ly * box.xy);
    out[2] = -bias * dx * (Lz * box.xz);
    out[3] = -bias * dy * Ly;
    out[4] = -bias * dy * (Lz * box.yz);
    out[5] = -bias * dz * Lz;
}

// Density: Density.cc:20-54 (host scalar CV used by the test_2d.py scenario)
template <class S> inline S density_value(const Box<S>& box, unsigned N_group) { return (S)N_group / box.volume(); }
template <class S> inline void density_virial(const Box<S>& box, unsigned N_group, S bias, S out[6]) {
    S V = box.volume();
    S fac = -(S)N_group / (V * V);
    S v = -bias * fac * box.L.x * box.L.y * box.L.z;
    out[0] = v; out[1] = 0; out[2] = 0; out[3] = v; out[4] = 0; out[5] = v;
}

}  // namespace oracle
