// Stand-in for HOOMD's bundled kiss_fft (hoomd/extern/kiss_fftnd.h), used ONLY to compile the reference's own sources for
// the oracle cross-check (oracle/Makefile, target _ref).  Same API and semantics as kiss_fftnd: unnormalised, forward
// exp(-i), inverse exp(+i), dims slowest first.  The transform itself is a plain separable DFT (radix-2 where the length
// is a power of two, O(n^2) otherwise) in double precision -- any correct DFT is admissible here.
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cmath>
#include <complex>
#include <cstdlib>
#include <vector>
#include <hoomd/ForceCompute.h>      // Scalar

#define kiss_fft_scalar Scalar
typedef struct { kiss_fft_scalar r; kiss_fft_scalar i; } kiss_fft_cpx;
struct kiss_fftnd_state { int dims[3]; int ndims; int inverse; };
typedef kiss_fftnd_state* kiss_fftnd_cfg;

inline kiss_fftnd_cfg kiss_fftnd_alloc(const int* dims, int ndims, int inverse_fft, void*, size_t*) {
    kiss_fftnd_cfg c = (kiss_fftnd_cfg)malloc(sizeof(kiss_fftnd_state));
    c->ndims = ndims; c->inverse = inverse_fft;
    for (int i = 0; i < 3; ++i) c->dims[i] = i < ndims ? dims[i] : 1;
    return c;
}
inline void kiss_fft_cleanup() {}
#define kiss_fft_free free

namespace ref_shim_fft {
inline void line(std::vector<std::complex<double>>& a, int sign) {
    const size_t n = a.size();
    if (n & (n - 1)) {       // general length
        std::vector<std::complex<double>> o(n);
        for (size_t k = 0; k < n; ++k) {
            std::complex<double> s = 0;
            for (size_t j = 0; j < n; ++j) s += a[j] * std::polar(1.0, sign * 2.0 * M_PI * (double)((k * j) % n) / (double)n);
            o[k] = s;
        }
        a.swap(o);
        return;
    }
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const std::complex<double> wl = std::polar(1.0, sign * 2.0 * M_PI / (double)len);
        for (size_t i = 0; i < n; i += len) {
            std::complex<double> w = 1;
            for (size_t k = 0; k < len / 2; ++k) {
                const std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v; a[i + k + len / 2] = u - v;
                w *= wl;
            }
        }
    }
}
}  // namespace ref_shim_fft

inline void kiss_fftnd(kiss_fftnd_cfg cfg, const kiss_fft_cpx* fin, kiss_fft_cpx* fout) {
    const int n0 = cfg->dims[0], n1 = cfg->dims[1], n2 = cfg->dims[2];      // n2 fastest
    const int sign = cfg->inverse ? +1 : -1;
    const size_t M = (size_t)n0 * n1 * n2;
    std::vector<std::complex<double>> a(M);
    for (size_t i = 0; i < M; ++i) a[i] = std::complex<double>(fin[i].r, fin[i].i);
    std::vector<std::complex<double>> l;
    l.resize(n2);
    for (int i0 = 0; i0 < n0; ++i0) for (int i1 = 0; i1 < n1; ++i1) {
        std::complex<double>* p = &a[((size_t)i0 * n1 + i1) * n2];
        for (int k = 0; k < n2; ++k) l[k] = p[k];
        ref_shim_fft::line(l, sign);
        for (int k = 0; k < n2; ++k) p[k] = l[k];
    }
    l.resize(n1);
    for (int i0 = 0; i0 < n0; ++i0) for (int i2 = 0; i2 < n2; ++i2) {
        for (int k = 0; k < n1; ++k) l[k] = a[((size_t)i0 * n1 + k) * n2 + i2];
        ref_shim_fft::line(l, sign);
        for (int k = 0; k < n1; ++k) a[((size_t)i0 * n1 + k) * n2 + i2] = l[k];
    }
    l.resize(n0);
    for (int i1 = 0; i1 < n1; ++i1) for (int i2 = 0; i2 < n2; ++i2) {
        for (int k = 0; k < n0; ++k) l[k] = a[((size_t)k * n1 + i1) * n2 + i2];
        ref_shim_fft::line(l, sign);
        for (int k = 0; k < n0; ++k) a[((size_t)k * n1 + i1) * n2 + i2] = l[k];
    }
    for (size_t i = 0; i < M; ++i) { fout[i].r = (kiss_fft_scalar)a[i].real(); fout[i].i = (kiss_fft_scalar)a[i].imag(); }
}
