// mesh_fft.cuh -- hand-written radix FFT stages for the OrderParameterMesh path (sm_100a).
//
// Replaces the reference's cuFFT C2C plans (OrderParameterMeshGPU.cc:118,256,322-325; cufftExecC2C forward
// and inverse, 6 sweeps of 16 B/cell) with a real-to-complex pipeline of 5 sweeps of 8 B/cell:
//   x forward (R2C, packed) -> y forward -> z forward + convolve + CV energy + z inverse (fused, G never
//   reaches HBM) -> y inverse -> x inverse (C2R).
// Semantics are those of the reference's transforms (kiss_fftnd on the CPU path, OrderParameterMesh.cc:319-325,
// 655,719): unnormalised, forward e^{-i}, inverse e^{+i}, index x + nx*(y + ny*z).
//
// Engine: a line of L complex points (L = 16..512, power of two) is transformed by T = L/16 threads, 16 points per
// thread in registers, as a Stockham autosort sequence of radix-16/8/4/2 stages (256 points: two radix-16 stages, ONE
// exchange; round 1 used 8 points per thread and three stages -- the sweeps were bound by shared-memory traffic); between stages the points are
// exchanged through a shared-memory tile.  16 lines are processed side by side so that every global and shared
// access of a half-warp is one contiguous 128-byte segment (16 x float2).
//
// All index/phase logic lives in __host__ __device__ "phase" functions (one per barrier interval) so the same
// code is validated on the CPU by tests/cpu_emul/fft_emul.cu, which runs the phases thread by thread.
#pragma once
#include <cuda_runtime.h>

#ifndef MHD
#define MHD __host__ __device__ __forceinline__
#endif

namespace metad {
namespace fft {

constexpr int kE = 16;       // points per thread (16: a 256-point line is two radix-16 stages, ONE exchange through shared memory)
constexpr int kLines = 16;   // lines per tile

MHD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
MHD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
MHD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
MHD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -i (SIGN=-1, forward) or +i (SIGN=+1, inverse)
template <int SIGN> MHD float2 mul_i(float2 a) { return SIGN < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x); }
// table holds forward twiddles exp(-2 pi i k / L); the inverse transform uses their conjugates
template <int SIGN> MHD float2 twiddle(float2 a, float2 w) { return SIGN < 0 ? cmul(a, w) : cmul(a, cconj(w)); }

template <int SIGN> MHD void dft2(float2& a, float2& b) {
    const float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}
template <int SIGN> MHD void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_i<SIGN>(csub(a1, a3));
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
}
template <int SIGN> MHD void dft8(float2 (&v)[8]) {
    // decimation in frequency: X[2k] = DFT4(v[n]+v[n+4]), X[2k+1] = DFT4((v[n]-v[n+4]) w8^n)
    const float h = 0.70710678118654752440f;
    float2 e0 = cadd(v[0], v[4]), e1 = cadd(v[1], v[5]), e2 = cadd(v[2], v[6]), e3 = cadd(v[3], v[7]);
    float2 o0 = csub(v[0], v[4]), o1 = csub(v[1], v[5]), o2 = csub(v[2], v[6]), o3 = csub(v[3], v[7]);
    // w8^1 = (1 -+ i)/sqrt2, w8^2 = -+i, w8^3 = (-1 -+ i)/sqrt2   (upper sign: forward)
    o1 = SIGN < 0 ? make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x)) : make_float2(h * (o1.x - o1.y), h * (o1.y + o1.x));
    o2 = mul_i<SIGN>(o2);
    o3 = SIGN < 0 ? make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y)) : make_float2(-h * (o3.x + o3.y), h * (o3.x - o3.y));
    dft4<SIGN>(e0, e1, e2, e3);
    dft4<SIGN>(o0, o1, o2, o3);
    v[0] = e0; v[2] = e1; v[4] = e2; v[6] = e3;
    v[1] = o0; v[3] = o1; v[5] = o2; v[7] = o3;
}

template <int SIGN> MHD void dft16(float2 (&v)[16]) {
    // n = n1 + 4 n2, k = 4 k1 + k2:  X[4 k1 + k2] = sum_n1 w4^(n1 k1) [ w16^(n1 k2) sum_n2 w4^(n2 k2) v[n1 + 4 n2] ]
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) dft4<SIGN>(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);       // -> y[n1][k2] in v[n1 + 4 k2]
    // twiddles w16^(n1 k2), forward: cos - i sin; inverse: cos + i sin
    auto rot = [](float2 a, float c, float sn) { return SIGN < 0 ? make_float2(a.x * c + a.y * sn, a.y * c - a.x * sn) : make_float2(a.x * c - a.y * sn, a.y * c + a.x * sn); };
    v[1 + 4 * 1] = rot(v[1 + 4 * 1], c1, s1);          // m = 1
    v[1 + 4 * 2] = rot(v[1 + 4 * 2], h, h);            // m = 2
    v[1 + 4 * 3] = rot(v[1 + 4 * 3], s1, c1);          // m = 3
    v[2 + 4 * 1] = rot(v[2 + 4 * 1], h, h);            // m = 2
    v[2 + 4 * 2] = mul_i<SIGN>(v[2 + 4 * 2]);          // m = 4
    v[2 + 4 * 3] = rot(v[2 + 4 * 3], -h, h);           // m = 6
    v[3 + 4 * 1] = rot(v[3 + 4 * 1], s1, c1);          // m = 3
    v[3 + 4 * 2] = rot(v[3 + 4 * 2], -h, h);           // m = 6
    v[3 + 4 * 3] = rot(v[3 + 4 * 3], -c1, -s1);        // m = 9
    float2 o[16];
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
        float2 a0 = v[0 + 4 * k2], a1 = v[1 + 4 * k2], a2 = v[2 + 4 * k2], a3 = v[3 + 4 * k2];
        dft4<SIGN>(a0, a1, a2, a3);                    // over n1 -> k1
        o[k2] = a0; o[4 + k2] = a1; o[8 + k2] = a2; o[12 + k2] = a3;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = o[k];
}

// radix sequence per line length
template <int L> struct Plan;
template <> struct Plan<16>  { static constexpr int n = 1; static constexpr int r0 = 16, r1 = 1, r2 = 1; };
template <> struct Plan<32>  { static constexpr int n = 2; static constexpr int r0 = 16, r1 = 2, r2 = 1; };
template <> struct Plan<64>  { static constexpr int n = 2; static constexpr int r0 = 16, r1 = 4, r2 = 1; };
template <> struct Plan<128> { static constexpr int n = 2; static constexpr int r0 = 16, r1 = 8, r2 = 1; };
template <> struct Plan<256> { static constexpr int n = 2; static constexpr int r0 = 16, r1 = 16, r2 = 1; };
template <> struct Plan<512> { static constexpr int n = 3; static constexpr int r0 = 16, r1 = 16, r2 = 2; };

// shared-memory tile layouts: element (line w, index l)
struct LayoutCol {   // y/z passes: tile[l][w], w fastest (16 float2 = 128 B per l)
    MHD static int addr(int w, int l, int /*L*/) { return l * kLines + w; }
    MHD static int size(int L) { return L * kLines; }
};
template <int G> struct LayoutColWide {   // y pass: G side-by-side groups of 16 lines, tile[l][16 G]; a group starts at column 16 g
    MHD static int addr(int w, int l, int /*L*/) { return l * (kLines * G) + w; }
    MHD static int size(int L) { return L * kLines * G; }
};
struct LayoutRow {   // x pass: tile[w][l] with one float2 of padding per line (odd stride: conflict-free)
    MHD static int addr(int w, int l, int L) { return w * (L + 1) + l; }
    MHD static int size(int L) { return kLines * (L + 1); }
};

// a[q] <- tile(w, t + T q): the register set of thread t is always the residue class t mod T
template <int L, class Lay> MHD void stage_load(float2 (&a)[kE], const float2* tile, int w, int t) {
    constexpr int T = L / kE;
#pragma unroll
    for (int q = 0; q < kE; ++q) a[q] = tile[Lay::addr(w, t + T * q, L)];
}

// one Stockham stage of radix R with NS = product of the previous radices.  Butterfly m (m < 8/R) of thread t
// is virtual thread j = t + m T; its inputs are a[m + r*(8/R)], r < R (indices j + r L/R).
// tw: forward twiddle table exp(-2 pi i k / TWL), k < TWL, TWL a multiple of L.
template <int L, int R, int NS, int SIGN, int TWL> MHD void stage_compute(float2 (&a)[kE], int t, const float2* tw) {
    constexpr int T = L / kE, B = kE / R;
#pragma unroll
    for (int m = 0; m < B; ++m) {
        const int j = t + m * T;
        const int k = j & (NS - 1);
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = a[m + r * B];
        if (NS > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = twiddle<SIGN>(v[r], tw[k * r * (TWL / (NS * R))]);
        }
        if (R == 16) {
            float2 u[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) u[r] = v[r % R];
            dft16<SIGN>(u);
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = u[r];
        } else if (R == 8) {
            float2 u[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) u[r] = v[r % R];
            dft8<SIGN>(u);
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = u[r];
        } else if (R == 4) {
            dft4<SIGN>(v[0], v[1 % R], v[2 % R], v[3 % R]);
        } else if (R == 2) {
            dft2<SIGN>(v[0], v[1 % R]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) a[m + r * B] = v[r];
    }
}

// scatter the stage outputs: out[(j/NS) NS R + (j%NS) + r NS]
template <int L, int R, int NS, class Lay> MHD void stage_store(const float2 (&a)[kE], float2* tile, int w, int t) {
    constexpr int T = L / kE, B = kE / R;
#pragma unroll
    for (int m = 0; m < B; ++m) {
        const int j = t + m * T;
        const int k = j & (NS - 1);
        const int base = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) tile[Lay::addr(w, base + r * NS, L)] = a[m + r * B];
    }
}

// index of the output that register slot (m, r) of thread t holds after the LAST stage (radix R, stride NS)
template <int L, int R, int NS> MHD int out_index(int t, int m, int r) {
    constexpr int T = L / kE;
    const int j = t + m * T;
    const int k = j & (NS - 1);
    return (j - k) * R + k + r * NS;
}

#ifdef __CUDACC__
// Whole line transform on the device.  On entry the tile holds the input (natural order) and has been
// synchronised; on exit it holds the output (natural order) and has been synchronised.
template <int L, int SIGN, int TWL, class Lay>
__device__ __forceinline__ void line_fft(float2* tile, int w, int t, const float2* tw) {
    using P = Plan<L>;
    float2 a[kE];
    stage_load<L, Lay>(a, tile, w, t);
    stage_compute<L, P::r0, 1, SIGN, TWL>(a, t, tw);
    __syncthreads();
    stage_store<L, P::r0, 1, Lay>(a, tile, w, t);
    __syncthreads();
    if (P::n >= 2) {
        stage_load<L, Lay>(a, tile, w, t);
        stage_compute<L, P::r1, P::r0, SIGN, TWL>(a, t, tw);
        __syncthreads();
        stage_store<L, P::r1, P::r0, Lay>(a, tile, w, t);
        __syncthreads();
    }
    if (P::n >= 3) {
        stage_load<L, Lay>(a, tile, w, t);
        stage_compute<L, P::r2, P::r0 * P::r1, SIGN, TWL>(a, t, tw);
        __syncthreads();
        stage_store<L, P::r2, P::r0 * P::r1, Lay>(a, tile, w, t);
        __syncthreads();
    }
}
#endif

// ---- real <-> half-complex packing of the x pass ------------------------------------------------------
// A real row x[0..N) is viewed as Lc = N/2 complex points z[n] = x[2n] + i x[2n+1]; Z = DFT_Lc(z).
// forward:  X[k] = (Z[k] + conj Z[Lc-k])/2 - (i/2) w^k (Z[k] - conj Z[Lc-k]),  w = exp(-2 pi i / N)
//           stored packed: P[0] = X[0] + i X[N/2] (both real), P[k] = X[k] for 0 < k < Lc.
// Each call handles the pair (k, Lc-k), 0 <= k <= Lc/2, in place.
MHD void r2c_pair(float2& zk, float2& zp, int k, int Lc, float2 wk /* exp(-2 pi i k/N) */) {
    if (k == 0) {
        zk = make_float2(zk.x + zk.y, zk.x - zk.y);
        return;
    }
    if (2 * k == Lc) {
        zk = cconj(zk);
        return;
    }
    const float2 a = zk, b = cconj(zp);
    const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y + b.y));    // (Z[k] + conj Z[Lc-k])/2
    const float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y - b.y));    // (Z[k] - conj Z[Lc-k])/2
    const float2 o = cmul(make_float2(d.y, -d.x), wk);                      // -i d w^k
    zk = cadd(e, o);
    // X[Lc-k] = conj(e) - conj(o) ... derived from the same two inputs: X[Lc-k] = conj(E[k]) + w^(Lc-k) conj(O[k])
    zp = make_float2(e.x - o.x, -(e.y - o.y));
}
// inverse:  Z[k] = (X[k] + conj X[Lc-k]) + i (X[k] - conj X[Lc-k]) conj(w^k);  z = IDFT_Lc(Z) (unnormalised)
//           gives x[2n] = Re z[n], x[2n+1] = Im z[n] with x the unnormalised length-N inverse.
MHD void c2r_pair(float2& xk, float2& xp, int k, int Lc, float2 wk /* exp(-2 pi i k/N) */) {
    if (k == 0) {
        xk = make_float2(xk.x + xk.y, xk.x - xk.y);
        return;
    }
    if (2 * k == Lc) {
        xk = make_float2(2.0f * xk.x, -2.0f * xk.y);
        return;
    }
    const float2 a = xk, b = cconj(xp);
    const float2 e = cadd(a, b);                         // X[k] + conj X[Lc-k]
    const float2 d = csub(a, b);                         // X[k] - conj X[Lc-k]
    const float2 o = cmul(d, cconj(wk));                 // ... times exp(+2 pi i k/N)
    xk = make_float2(e.x - o.y, e.y + o.x);              // e + i o
    // Z[Lc-k] = conj(e) + i * conj(-o) ... = conj(e - i o) with the sign bookkeeping below
    xp = make_float2(e.x + o.y, -(e.y - o.x));
}

}  // namespace fft
}  // namespace metad
