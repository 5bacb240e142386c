#pragma once
// emul_fft.h -- CPU emulation of the mesh FFT sweeps (host-only program, built with nvcc, runs without a GPU).
// It executes the SAME __host__ __device__ phase functions the kernels in csrc/mesh_fft_kernels.cuh are made of
// (stage_load / stage_compute / stage_store, r2c_pair / c2r_pair, conv_general / conv_plane0), thread by thread,
// with the kernels' tile/index mapping, and dumps the result for comparison against numpy (tests/test_fft_emul.py).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../metadynamics_plugin_b200/csrc/mesh_fft_kernels.cuh"

using namespace metad::fft;

template <int L, int SIGN, int TWL, class Lay> void emul_line_fft(float2* tile, const float2* tw) {
    using P = Plan<L>;
    constexpr int T = L / kE;
    static float2 regs[kLines][L / kE][kE];
#define FORALL for (int w = 0; w < kLines; ++w) for (int t = 0; t < T; ++t)
    FORALL { stage_load<L, Lay>(regs[w][t], tile, w, t); stage_compute<L, P::r0, 1, SIGN, TWL>(regs[w][t], t, tw); }
    FORALL stage_store<L, P::r0, 1, Lay>(regs[w][t], tile, w, t);
    if (P::n >= 2) {
        FORALL { stage_load<L, Lay>(regs[w][t], tile, w, t); stage_compute<L, P::r1, P::r0, SIGN, TWL>(regs[w][t], t, tw); }
        FORALL stage_store<L, P::r1, P::r0, Lay>(regs[w][t], tile, w, t);
    }
    if (P::n >= 3) {
        FORALL { stage_load<L, Lay>(regs[w][t], tile, w, t); stage_compute<L, P::r2, P::r0 * P::r1, SIGN, TWL>(regs[w][t], t, tw); }
        FORALL stage_store<L, P::r2, P::r0 * P::r1, Lay>(regs[w][t], tile, w, t);
    }
#undef FORALL
}

static std::vector<float2> twiddles(int n) {
    std::vector<float2> t(n);
    for (int k = 0; k < n; ++k) { double ph = -2.0 * M_PI * k / n; t[k] = make_float2((float)cos(ph), (float)sin(ph)); }
    return t;
}

template <int LC> void x_fwd(float2* buf, unsigned rows) {
    auto tw = twiddles(2 * LC);
    std::vector<float2> tile(LayoutRow::size(LC));
    for (unsigned row0 = 0; row0 < rows; row0 += kLines) {
        for (int idx = 0; idx < kLines * LC; ++idx) { int w = idx / LC, l = idx % LC; tile[LayoutRow::addr(w, l, LC)] = buf[(size_t)(row0 + w) * LC + l]; }
        emul_line_fft<LC, -1, 2 * LC, LayoutRow>(tile.data(), tw.data());
        for (int idx = 0; idx < kLines * (LC / 2 + 1); ++idx) {
            int ww = idx & (kLines - 1), k = idx / kLines;
            r2c_pair(tile[LayoutRow::addr(ww, k, LC)], tile[LayoutRow::addr(ww, (LC - k) % LC, LC)], k, LC, tw[k]);
        }
        for (int idx = 0; idx < kLines * LC; ++idx) { int w = idx / LC, l = idx % LC; buf[(size_t)(row0 + w) * LC + l] = tile[LayoutRow::addr(w, l, LC)]; }
    }
}
template <int LC> void x_inv(float2* buf, unsigned rows) {
    auto tw = twiddles(2 * LC);
    std::vector<float2> tile(LayoutRow::size(LC));
    for (unsigned row0 = 0; row0 < rows; row0 += kLines) {
        for (int idx = 0; idx < kLines * LC; ++idx) { int w = idx / LC, l = idx % LC; tile[LayoutRow::addr(w, l, LC)] = buf[(size_t)(row0 + w) * LC + l]; }
        for (int idx = 0; idx < kLines * (LC / 2 + 1); ++idx) {
            int ww = idx & (kLines - 1), k = idx / kLines;
            c2r_pair(tile[LayoutRow::addr(ww, k, LC)], tile[LayoutRow::addr(ww, (LC - k) % LC, LC)], k, LC, tw[k]);
        }
        emul_line_fft<LC, +1, 2 * LC, LayoutRow>(tile.data(), tw.data());
        for (int idx = 0; idx < kLines * LC; ++idx) { int w = idx / LC, l = idx % LC; buf[(size_t)(row0 + w) * LC + l] = tile[LayoutRow::addr(w, l, LC)]; }
    }
}
template <int L, int SIGN> void y_pass(float2* buf, unsigned nxh, unsigned nz) {
    auto tw = twiddles(L);
    std::vector<float2> tile(LayoutCol::size(L));
    for (unsigned bz = 0; bz < nz; ++bz)
        for (unsigned bx = 0; bx < nxh / kLines; ++bx) {
            size_t base = (size_t)bz * L * nxh + (size_t)bx * kLines;
            for (int idx = 0; idx < kLines * L; ++idx) { int w = idx & (kLines - 1), l = idx / kLines; tile[idx] = buf[base + (size_t)l * nxh + w]; }
            emul_line_fft<L, SIGN, L, LayoutCol>(tile.data(), tw.data());
            for (int idx = 0; idx < kLines * L; ++idx) { int w = idx & (kLines - 1), l = idx / kLines; buf[base + (size_t)l * nxh + w] = tile[idx]; }
        }
}
template <int L> double z_fused(float2* buf, unsigned nx, unsigned ny, float inv_n, float d) {
    auto tw = twiddles(L);
    std::vector<float2> tile(LayoutCol::size(L));
    const unsigned nxh = nx / 2;
    const size_t zstride = (size_t)ny * nxh;
    double etot = 0.0;
    for (unsigned ky = 0; ky < ny; ++ky)
        for (unsigned bx = 0; bx < nxh / kLines; ++bx) {
            unsigned kx0 = bx * kLines;
            size_t base = (size_t)ky * nxh + kx0;
            for (int idx = 0; idx < kLines * L; ++idx) { int w = idx & (kLines - 1), l = idx / kLines; tile[idx] = buf[base + (size_t)l * zstride + w]; }
            emul_line_fft<L, -1, L, LayoutCol>(tile.data(), tw.data());
            double e = 0.0;
            for (int idx = 0; idx < kLines * L; ++idx) {
                int ww = idx & (kLines - 1); unsigned kz = idx / kLines;
                if (kx0 + ww == 0) continue;
                tile[idx] = conv_general(tile[idx], inv_n, d, nonneg(ky, ny) && nonneg(kz, L), e);
            }
            emul_line_fft<L, +1, L, LayoutCol>(tile.data(), tw.data());
            for (int idx = 0; idx < kLines * L; ++idx) {
                int ww = idx & (kLines - 1), l = idx / kLines;
                if (kx0 + ww == 0) continue;
                buf[base + (size_t)l * zstride + ww] = tile[idx];
            }
            etot += e;
        }
    return etot;
}
template <int L> double z_plane0(float2* buf, unsigned nx, unsigned ny, float inv_n, float d, float dc = 0.f) {
    auto tw = twiddles(L);
    std::vector<float2> tile(LayoutCol::size(L)), outv(LayoutCol::size(L));
    const unsigned nxh = nx / 2;
    const size_t zstride = (size_t)ny * nxh;
    double etot = 0.0;
    const unsigned nblocks = (ny / 2 + 1 + kLines / 2 - 1) / (kLines / 2);
    for (unsigned b = 0; b < nblocks; ++b) {
        for (int idx = 0; idx < kLines * L; ++idx) {
            int w = idx & (kLines - 1), l = idx / kLines;
            unsigned kyp = b * (kLines / 2) + (w >> 1);
            float2 v = make_float2(0.f, 0.f);
            if (kyp <= ny / 2) { unsigned ky = (w & 1) ? (ny - kyp) % ny : kyp; v = buf[(size_t)ky * nxh + (size_t)l * zstride]; }
            tile[idx] = v;
        }
        emul_line_fft<L, -1, L, LayoutCol>(tile.data(), tw.data());
        double e = 0.0;
        for (int idx = 0; idx < kLines * L; ++idx) {
            int ww = idx & (kLines - 1); unsigned kz = idx / kLines;
            unsigned kyp = b * (kLines / 2) + (ww >> 1), pky = (ny - kyp) % ny, ky = (ww & 1) ? pky : kyp;
            bool valid = kyp <= ny / 2, counted = valid && (!(ww & 1) || pky != kyp);
            outv[idx] = conv_plane0(tile[idx], tile[((L - kz) % L) * kLines + (ww ^ 1)], inv_n, d, ky, kz, ny, L, counted, e, dc);
        }
        tile = outv;
        emul_line_fft<L, +1, L, LayoutCol>(tile.data(), tw.data());
        for (int idx = 0; idx < kLines * L; ++idx) {
            int ww = idx & (kLines - 1), l = idx / kLines;
            unsigned kyp = b * (kLines / 2) + (ww >> 1), pky = (ny - kyp) % ny;
            if (kyp > ny / 2) continue;
            if ((ww & 1) && pky == kyp) continue;
            unsigned ky = (ww & 1) ? pky : kyp;
            buf[(size_t)ky * nxh + (size_t)l * zstride] = tile[idx];
        }
        etot += e;
    }
    return etot;
}

#define DISPATCH(n, CALL)                                  \
    switch (n) {                                           \
        case 16: { constexpr int LL = 16; CALL; } break;   \
        case 32: { constexpr int LL = 32; CALL; } break;   \
        case 64: { constexpr int LL = 64; CALL; } break;   \
        case 128: { constexpr int LL = 128; CALL; } break; \
        case 256: { constexpr int LL = 256; CALL; } break; \
        case 512: { constexpr int LL = 512; CALL; } break; \
        default: fprintf(stderr, "unsupported length %d\n", (int)(n)); exit(2); \
    }

