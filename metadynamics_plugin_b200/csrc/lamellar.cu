// lamellar.cu -- LamellarOrderParameter CV + bias force for sm_100a.
//
// Reference behaviour (CPU path = parity target): LamellarOrderParameter.cc:42-74 (CV), :77-140 (force),
// :143-179 (Fourier modes).  Reference GPU drivers replaced: LamellarOrderParameterGPU.cu:8-147 (one pass
// over all positions PER wave vector, float accumulation, host-side final sum) and :149-236.
//
// Design (HBM-bound, 16 B read per particle for the CV pass, 16 B read + 16 B write for the force pass):
//   * one coalesced 128-bit streaming load per particle, all n_wave modes evaluated from that one load;
//   * phases are evaluated in *turns* in fp64 (t = (q_k/2pi).r), reduced exactly to [-1/2, 1/2] and fed to an fp32
//     polynomial sin/cos(pi u) (no conversion- or transcendental-pipe instruction per term), so the result tracks
//     the double-precision CPU build to ~1e-7 although the transcendental itself is fp32;
//   * per-thread fp64 accumulators -> warp shuffles -> one partial per block -> the last block to finish
//     (ticket counter) sums the block partials in block order: deterministic, no second launch, no host sum;
//   * the CV and the bias factor stay in device memory (double), the force pass reads dV/ds from there.
#include "common.cuh"

#include <vector>

namespace metad {

constexpr int kLamMaxWave = 8;     // wave vectors handled per pass (register accumulators)
constexpr int kLamThreads = 256;

template <int NW> struct WaveSet {
    double qt[NW][3];   // q_k / (2 pi): turns per unit length
    float q[NW][3];     // q_k in rad per unit length (force prefactor)
};

// ---- phase arithmetic without the conversion / transcendental pipe ------------------------------------------------
// (measured: with fp64 rint, float<->double conversions of every term and sincospif the kernels ran at 70-76 % of the
// XU pipe and 2x above their HBM time; profiles/r01i_lamellar_c5_metrics_before.csv)
// fractional part of a phase given in turns, in [-1/2, 1/2]: rint by the fp64 magic constant 1.5 * 2^52 (two DADDs)
__device__ __forceinline__ float frac_turns(double t) {
    const double r = (t + 6755399441055744.0) - 6755399441055744.0;
    return (float)(t - r);
}
// sin(pi u), cos(pi u) for |u| <= 1: quadrant k = rint(2u) by the fp32 magic constant, then Taylor polynomials in
// v = u - k/2, |v| <= 1/4 (truncation error < 3e-9), all on the FMA pipe
__device__ __forceinline__ void sincospi_small(float u, float& s, float& c) {
    const float km = fmaf(2.0f, u, 12582912.0f);
    const int q = __float_as_int(km) & 3;                 // low bits of the integer k (two's complement: valid for k < 0)
    const float k = km - 12582912.0f;
    const float v = fmaf(k, -0.5f, u);
    const float v2 = v * v;
    float sp = fmaf(v2, 0.0821458866f, -0.599264529f);
    sp = fmaf(sp, v2, 2.55016404f);
    sp = fmaf(sp, v2, -5.16771278f);
    sp = fmaf(sp, v2, 3.14159265f);
    sp *= v;
    float cp = fmaf(v2, -0.0258068913f, 0.235330630f);
    cp = fmaf(cp, v2, -1.33526277f);
    cp = fmaf(cp, v2, 4.05871213f);
    cp = fmaf(cp, v2, -4.93480220f);
    cp = fmaf(cp, v2, 1.0f);
    // rotate by k quarter turns: k=1: (c,-s) ... stated as sin/cos of pi(v + k/2)
    const float s1 = (q & 1) ? cp : sp, c1 = (q & 1) ? sp : cp;
    s = (q & 2) ? -s1 : s1;
    c = ((q + 1) & 2) ? -c1 : c1;
}

template <int NW>
__global__ void __launch_bounds__(kLamThreads)
lamellar_modes_kernel(const float4* __restrict__ postype, unsigned N, WaveSet<NW> ws, const float* __restrict__ mode,
                      double* __restrict__ partials, unsigned* __restrict__ ticket, double* __restrict__ d_modes,
                      int k0, int finalize, double n_global, double* __restrict__ d_cv) {
    double accr[NW], acci[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) { accr[k] = 0.0; acci[k] = 0.0; }

    // terms are summed in fp32 over short runs (8 particles: error ~1e-7 of a run) and the runs in fp64
    float runr[NW], runi[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) { runr[k] = 0.f; runi[k] = 0.f; }
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned in_run = 0;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const float4 p = ld_stream(postype + i);
        const float a = __ldg(mode + __float_as_int(p.w));
        const double x = (double)p.x, y = (double)p.y, z = (double)p.z;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const double t = fma(ws.qt[k][0], x, fma(ws.qt[k][1], y, ws.qt[k][2] * z));
            float s, c;
            sincospi_small(2.0f * frac_turns(t), s, c);
            runr[k] = fmaf(a, c, runr[k]);
            runi[k] = fmaf(a, s, runi[k]);
        }
        if (++in_run == 8) {
#pragma unroll
            for (int k = 0; k < NW; ++k) { accr[k] += (double)runr[k]; acci[k] += (double)runi[k]; runr[k] = 0.f; runi[k] = 0.f; }
            in_run = 0;
        }
    }
#pragma unroll
    for (int k = 0; k < NW; ++k) { accr[k] += (double)runr[k]; acci[k] += (double)runi[k]; }

    __shared__ double red[32];
    __shared__ bool is_last;
    double* mine = partials + (size_t)blockIdx.x * (2 * NW);
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        double r = block_sum(accr[k], red);
        double im = block_sum(acci[k], red);
        if (threadIdx.x == 0) { mine[2 * k] = r; mine[2 * k + 1] = im; }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // last block: fixed-order sum over block partials, one (mode, re/im) component per warp lane group
    for (int c = threadIdx.x >> 5; c < 2 * NW; c += (blockDim.x >> 5)) {
        double s = 0.0;
        for (unsigned b = threadIdx.x & 31; b < gridDim.x; b += 32) s += __ldcg(&partials[(size_t)b * (2 * NW) + c]);
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) d_modes[2 * k0 + c] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *ticket = 0;
        if (finalize) {
            double sum = 0.0;
            for (int k = 0; k < NW; ++k) sum += d_modes[2 * (k0 + k)];
            *d_cv = sum / n_global;
        }
    }
}

__global__ void lamellar_finalize_kernel(const double* __restrict__ d_modes, int n_wave, double n_global,
                                         double* __restrict__ d_cv) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double sum = 0.0;
        for (int k = 0; k < n_wave; ++k) sum += d_modes[2 * k];
        *d_cv = sum / n_global;
    }
}

template <int NW>
__global__ void __launch_bounds__(kLamThreads)
lamellar_force_kernel(const float4* __restrict__ postype, float4* __restrict__ force, unsigned N, WaveSet<NW> ws,
                      const float* __restrict__ mode, const double* __restrict__ d_bias, double n_global,
                      int accumulate) {
    const float scale = (float)(*d_bias / n_global);
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const float4 p = ld_stream(postype + i);
        const float a = __ldg(mode + __float_as_int(p.w));
        const double x = (double)p.x, y = (double)p.y, z = (double)p.z;
        float fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const double t = fma(ws.qt[k][0], x, fma(ws.qt[k][1], y, ws.qt[k][2] * z));
            float sn, cs;
            sincospi_small(2.0f * frac_turns(t), sn, cs);
            const float f = 2.0f * a * sn;
            fx = fmaf(ws.q[k][0], f, fx);
            fy = fmaf(ws.q[k][1], f, fy);
            fz = fmaf(ws.q[k][2], f, fz);
        }
        float4 out = make_float4(fx * scale, fy * scale, fz * scale, 0.f);
        if (accumulate) {
            const float4 old = force[i];
            out.x += old.x; out.y += old.y; out.z += old.z;
        }
        st_stream(force + i, out);
    }
}

}  // namespace metad

using namespace metad;

struct metad_lamellar {
    int n_wave = 0, ntypes = 0;
    std::vector<int> lattice;
    float* d_mode = nullptr;
    double* d_partials = nullptr;
    unsigned* d_ticket = nullptr;
    int max_blocks = 0;
};

namespace {
// q_k = n_kx b1 + n_ky b2 + n_kz b3, b_i = 2 pi (a_j x a_k)/V of the global box (LamellarOrderParameter.cc:151-159)
void wave_vectors(const metad_box* box, const int* lv, int n, double (*q_turns)[3]) {
    const double Lx = box->L[0], Ly = box->L[1], Lz = box->L[2];
    const double xy = box->tilt[0], xz = box->tilt[1], yz = box->tilt[2];
    const double a1[3] = {Lx, 0, 0}, a2[3] = {Ly * xy, Ly, 0}, a3[3] = {Lz * xz, Lz * yz, Lz};
    const double V = Lx * Ly * Lz;
    const double b1[3] = {(a2[1] * a3[2] - a2[2] * a3[1]) / V, (a2[2] * a3[0] - a2[0] * a3[2]) / V, (a2[0] * a3[1] - a2[1] * a3[0]) / V};
    const double b2[3] = {(a3[1] * a1[2] - a3[2] * a1[1]) / V, (a3[2] * a1[0] - a3[0] * a1[2]) / V, (a3[0] * a1[1] - a3[1] * a1[0]) / V};
    const double b3[3] = {(a1[1] * a2[2] - a1[2] * a2[1]) / V, (a1[2] * a2[0] - a1[0] * a2[2]) / V, (a1[0] * a2[1] - a1[1] * a2[0]) / V};
    for (int k = 0; k < n; ++k)
        for (int c = 0; c < 3; ++c) q_turns[k][c] = lv[3 * k] * b1[c] + lv[3 * k + 1] * b2[c] + lv[3 * k + 2] * b3[c];
}

int lam_blocks(unsigned N, int cap) {
    long b = ((long)N + kLamThreads * 4L - 1) / (kLamThreads * 4L);
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (int)b;
}

template <int NW> WaveSet<NW> make_waveset(const double (*qt)[3], int k0, int n_wave) {
    WaveSet<NW> ws;
    for (int k = 0; k < NW; ++k)
        for (int c = 0; c < 3; ++c) {
            const double v = (k0 + k < n_wave) ? qt[k0 + k][c] : 0.0;
            ws.qt[k][c] = v;
            ws.q[k][c] = (float)(2.0 * M_PI * v);
        }
    return ws;
}

template <int NW>
int launch_modes(metad_lamellar* p, const float* d_postype, unsigned N, unsigned N_global, const double (*qt)[3],
                 int k0, double* d_modes, int finalize, double* d_cv, cudaStream_t st) {
    auto ws = make_waveset<NW>(qt, k0, p->n_wave);
    const int blocks = lam_blocks(N, p->max_blocks);
    lamellar_modes_kernel<NW><<<blocks, kLamThreads, 0, st>>>((const float4*)d_postype, N, ws, p->d_mode, p->d_partials,
                                                              p->d_ticket, d_modes, k0, finalize, (double)N_global, d_cv);
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}
template <int NW>
int launch_force(metad_lamellar* p, const float* d_postype, float* d_force, unsigned N, unsigned N_global,
                 const double (*qt)[3], int k0, const double* d_bias, int accumulate, cudaStream_t st) {
    auto ws = make_waveset<NW>(qt, k0, p->n_wave);
    const int blocks = lam_blocks(N, p->max_blocks);
    lamellar_force_kernel<NW><<<blocks, kLamThreads, 0, st>>>((const float4*)d_postype, (float4*)d_force, N, ws, p->d_mode,
                                                              d_bias, (double)N_global, accumulate);
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}
}  // namespace

extern "C" int metad_lamellar_create(metad_lamellar** out, int n_wave, const int* lattice_vectors, int ntypes,
                                     const double* mode) {
    METAD_REQUIRE(out && lattice_vectors && mode, "metad_lamellar_create: null argument");
    METAD_REQUIRE(n_wave > 0, "cv.lamellar: List of supplied lattice vectors is empty.");
    METAD_REQUIRE(ntypes > 0, "cv.lamellar: Number of mode parameters has to equal the number of particle types!");
    auto* p = new metad_lamellar();
    p->n_wave = n_wave;
    p->ntypes = ntypes;
    p->lattice.assign(lattice_vectors, lattice_vectors + 3 * n_wave);
    p->max_blocks = device_sm_count() * 8;
    std::vector<float> m(ntypes);
    for (int i = 0; i < ntypes; ++i) m[i] = (float)mode[i];
    cudaError_t e;
    if ((e = cudaMalloc(&p->d_mode, sizeof(float) * ntypes)) != cudaSuccess ||
        (e = cudaMemcpy(p->d_mode, m.data(), sizeof(float) * ntypes, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMalloc(&p->d_partials, sizeof(double) * 2 * kLamMaxWave * p->max_blocks)) != cudaSuccess ||
        (e = cudaMalloc(&p->d_ticket, sizeof(unsigned))) != cudaSuccess ||
        (e = cudaMemset(p->d_ticket, 0, sizeof(unsigned))) != cudaSuccess) {
        metad_lamellar_destroy(p);
        return cuda_fail(e, "metad_lamellar_create", __FILE__, __LINE__);
    }
    *out = p;
    return METAD_OK;
}

extern "C" int metad_lamellar_destroy(metad_lamellar* p) {
    if (!p) return METAD_OK;
    cudaFree(p->d_mode);
    cudaFree(p->d_partials);
    cudaFree(p->d_ticket);
    delete p;
    return METAD_OK;
}

extern "C" int metad_lamellar_modes(metad_lamellar* p, const float* d_postype, unsigned N, unsigned N_global,
                                    const metad_box* global_box, double* d_modes, int finalize, double* d_cv,
                                    metad_stream_t stream) {
    METAD_REQUIRE(p && global_box && d_modes, "metad_lamellar_modes: null argument");
    METAD_REQUIRE(!finalize || d_cv, "metad_lamellar_modes: finalize requested without d_cv");
    METAD_REQUIRE(N == 0 || d_postype, "metad_lamellar_modes: null positions");
    METAD_REQUIRE(N_global > 0, "metad_lamellar_modes: N_global must be positive");
    std::vector<double> qtv(3 * (size_t)p->n_wave);
    auto qt = reinterpret_cast<double (*)[3]>(qtv.data());
    wave_vectors(global_box, p->lattice.data(), p->n_wave, qt);
    const bool single = p->n_wave <= kLamMaxWave;
    for (int k0 = 0; k0 < p->n_wave; k0 += kLamMaxWave) {
        const int rem = p->n_wave - k0;
        const int fin = (single && finalize) ? 1 : 0;
        int rc;
        switch (rem >= kLamMaxWave ? kLamMaxWave : rem) {
            case 1: rc = launch_modes<1>(p, d_postype, N, N_global, qt, k0, d_modes, fin, d_cv, stream); break;
            case 2: rc = launch_modes<2>(p, d_postype, N, N_global, qt, k0, d_modes, fin, d_cv, stream); break;
            case 3: rc = launch_modes<3>(p, d_postype, N, N_global, qt, k0, d_modes, fin, d_cv, stream); break;
            case 4: rc = launch_modes<4>(p, d_postype, N, N_global, qt, k0, d_modes, fin, d_cv, stream); break;
            case 5: rc = launch_modes<5>(p, d_postype, N, N_global, qt, k0, d_modes, fin, d_cv, stream); break;
            case 6: rc = launch_modes<6>(p, d_postype, N, N_global, qt, k0, d_modes, fin, d_cv, stream); break;
            case 7: rc = launch_modes<7>(p, d_postype, N, N_global, qt, k0, d_modes, fin, d_cv, stream); break;
            default: rc = launch_modes<8>(p, d_postype, N, N_global, qt, k0, d_modes, fin, d_cv, stream); break;
        }
        if (rc != METAD_OK) return rc;
    }
    if (finalize && !single) return metad_lamellar_finalize(p, d_modes, N_global, d_cv, stream);
    return METAD_OK;
}

extern "C" int metad_lamellar_finalize(metad_lamellar* p, const double* d_modes, unsigned N_global, double* d_cv,
                                       metad_stream_t stream) {
    METAD_REQUIRE(p && d_modes && d_cv, "metad_lamellar_finalize: null argument");
    METAD_REQUIRE(N_global > 0, "metad_lamellar_finalize: N_global must be positive");
    lamellar_finalize_kernel<<<1, 32, 0, stream>>>(d_modes, p->n_wave, (double)N_global, d_cv);
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}

extern "C" int metad_lamellar_forces(metad_lamellar* p, const float* d_postype, float* d_force, unsigned N,
                                     unsigned N_global, const metad_box* global_box, const double* d_bias,
                                     metad_stream_t stream) {
    METAD_REQUIRE(p && global_box && d_bias, "metad_lamellar_forces: null argument");
    METAD_REQUIRE(N_global > 0, "metad_lamellar_forces: N_global must be positive");
    if (N == 0) return METAD_OK;
    METAD_REQUIRE(d_postype && d_force, "metad_lamellar_forces: null particle arrays");
    std::vector<double> qtv(3 * (size_t)p->n_wave);
    auto qt = reinterpret_cast<double (*)[3]>(qtv.data());
    wave_vectors(global_box, p->lattice.data(), p->n_wave, qt);
    for (int k0 = 0; k0 < p->n_wave; k0 += kLamMaxWave) {
        const int rem = p->n_wave - k0;
        const int acc = k0 > 0;
        int rc;
        switch (rem >= kLamMaxWave ? kLamMaxWave : rem) {
            case 1: rc = launch_force<1>(p, d_postype, d_force, N, N_global, qt, k0, d_bias, acc, stream); break;
            case 2: rc = launch_force<2>(p, d_postype, d_force, N, N_global, qt, k0, d_bias, acc, stream); break;
            case 3: rc = launch_force<3>(p, d_postype, d_force, N, N_global, qt, k0, d_bias, acc, stream); break;
            case 4: rc = launch_force<4>(p, d_postype, d_force, N, N_global, qt, k0, d_bias, acc, stream); break;
            case 5: rc = launch_force<5>(p, d_postype, d_force, N, N_global, qt, k0, d_bias, acc, stream); break;
            case 6: rc = launch_force<6>(p, d_postype, d_force, N, N_global, qt, k0, d_bias, acc, stream); break;
            case 7: rc = launch_force<7>(p, d_postype, d_force, N, N_global, qt, k0, d_bias, acc, stream); break;
            default: rc = launch_force<8>(p, d_postype, d_force, N, N_global, qt, k0, d_bias, acc, stream); break;
        }
        if (rc != METAD_OK) return rc;
    }
    return METAD_OK;
}
