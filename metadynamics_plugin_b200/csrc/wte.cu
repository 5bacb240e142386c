// wte.cu -- WellTemperedEnsemble reduce/scale kernels and the device-side umbrella of CollectiveVariable.
//
// Reference behaviour restated (CPU path = parity target): WellTemperedEnsemble.cc:30-68 (pe = sum of
// net_force.w + external energy), :135-188 (net force xyz, net torque xyzw, six virial rows *= 1+bias);
// CollectiveVariable.cc:22-66 (umbrella bias increment), :68-106 (umbrella energy).
// Reference GPU drivers replaced: gpu_reduce_potential_energy / gpu_scale_netforce
// (WellTemperedEnsemble.cu:19-243; atomicCAS double add, autotuned block size, managed scratch).
#include "common.cuh"
#include <map>
#include <mutex>
#include <utility>

namespace metad {

constexpr int kWteThreads = 256;

__global__ void __launch_bounds__(kWteThreads)
wte_reduce_kernel(const float4* __restrict__ net_force, unsigned N, double external_energy, double* __restrict__ partials,
                  unsigned* __restrict__ ticket, double* __restrict__ d_pe) {
    double acc = 0.0;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) acc += (double)ld_stream(net_force + i).w;
    __shared__ double red[32];
    __shared__ bool is_last;
    const double r = block_sum(acc, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = r;
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double s = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) s += __ldcg(partials + b);
    s = block_sum(s, red);
    if (threadIdx.x == 0) { *d_pe = s + external_energy; *ticket = 0; }
}

// sum_n f_i(n) . f_j(n) over the xyz components of two force arrays (computeSigma, IntegratorMetaDynamics.cc:1238-1247:
// the products of the CV derivatives); fp64 accumulation, per-block partials, last block adds them in order
__global__ void __launch_bounds__(kWteThreads)
force_dot_kernel(const float4* __restrict__ fi, const float4* __restrict__ fj, unsigned N, double scale, double* __restrict__ partials,
                 unsigned* __restrict__ ticket, double* __restrict__ d_out) {
    double acc = 0.0;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const float4 a = ld_stream(fi + i), b = ld_stream(fj + i);
        acc += (double)a.x * (double)b.x + (double)a.y * (double)b.y + (double)a.z * (double)b.z;
    }
    __shared__ double red[32];
    __shared__ bool is_last;
    const double r = block_sum(acc, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = r;
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double s = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) s += __ldcg(partials + b);
    s = block_sum(s, red);
    if (threadIdx.x == 0) { *d_out = scale * s; *ticket = 0; }
}

__global__ void __launch_bounds__(kWteThreads)
wte_scale_kernel(float4* __restrict__ net_force, float4* __restrict__ net_torque, float* __restrict__ net_virial,
                 unsigned pitch, unsigned N, const double* __restrict__ d_bias) {
    const float fac = (float)(1.0 + *d_bias);
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        float4 f = net_force[i];
        f.x *= fac; f.y *= fac; f.z *= fac;
        net_force[i] = f;
        if (net_torque) {
            float4 t = net_torque[i];
            t.x *= fac; t.y *= fac; t.z *= fac; t.w *= fac;
            net_torque[i] = t;
        }
        if (net_virial) {
#pragma unroll
            for (int r = 0; r < 6; ++r) net_virial[i + (size_t)r * pitch] *= fac;
        }
    }
}

__global__ void umbrella_kernel(int kind, double cv0, double kappa, double width_flat, double scale,
                                const double* __restrict__ d_cv, const double* __restrict__ d_bias_in,
                                double* __restrict__ d_bias_out, double* __restrict__ d_energy_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double val = *d_cv;
    double bias = d_bias_in ? *d_bias_in : 0.0;
    double energy = 0.0;
    const bool flat = (val < cv0 + width_flat / 2.0) && (val > cv0 - width_flat / 2.0);
    if (kind != 0 && !flat) {
        double delta = 0.0;
        if (val > cv0) delta = val - cv0 - width_flat / 2.0;
        else delta = val - cv0 + width_flat / 2.0;
        // energy uses delta = 0 when val == cv0 exactly (CollectiveVariable.cc:80-84)
        const double edelta = (val > cv0 || val < cv0) ? delta : 0.0;
        if (kind == 1) { bias += scale * 1.0; energy = scale * edelta; }
        else if (kind == 2) { bias += kappa * delta; energy = 0.5 * edelta * edelta * kappa; }
        else if (kind == 3) { bias += scale * 12.0 * pow(delta / kappa, 11.0) / kappa; energy = scale * pow(edelta / kappa, 12.0); }
        else if (kind == 4) {
            const double g = exp(-(val - cv0) * (val - cv0) / kappa / kappa / 2.0);
            bias -= scale * (val - cv0) * g;
            energy = scale * g - scale;
        }
    }
    if (d_bias_out) *d_bias_out = bias;
    if (d_energy_out) *d_energy_out = energy;
}

__global__ void set_double_kernel(double* __restrict__ dst, double v) { *dst = v; }

__global__ void __launch_bounds__(kWteThreads)
accumulate_force_kernel(float4* __restrict__ net, const float4* __restrict__ f, unsigned N, int init) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        float4 a = f[i];
        if (!init) { const float4 b = net[i]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
        net[i] = a;
    }
}

struct WteScratch {
    double* partials = nullptr;
    unsigned* ticket = nullptr;
    int blocks = 0;
};
// One scratch (partials + ticket of the last-block finalisation) per device AND stream: two reduces in flight on different
// streams, or one process driving several devices, must not share a ticket.  The handful of (device, stream) pairs a process
// uses live until it exits.
static std::mutex g_wte_mutex;
static std::map<std::pair<int, cudaStream_t>, WteScratch> g_wte;

static int wte_scratch(cudaStream_t stream, WteScratch** out) {
    int dev = 0;
    METAD_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_wte_mutex);
    WteScratch& w = g_wte[std::make_pair(dev, stream)];
    if (!w.partials) {
        // the first use of a stream may happen while that stream is being captured into a CUDA graph (the capture stream of a
        // framework is a fresh stream): allocations are legal there in relaxed capture mode, and the ticket is cleared by a
        // memset ON the stream (captured as a node: it then also runs, harmlessly, at every replay)
        cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
        METAD_CUDA(cudaThreadExchangeStreamCaptureMode(&mode));
        w.blocks = device_sm_count() * 8;
        cudaError_t e1 = cudaMalloc(&w.partials, sizeof(double) * w.blocks);
        cudaError_t e2 = cudaMalloc(&w.ticket, sizeof(unsigned));
        METAD_CUDA(cudaThreadExchangeStreamCaptureMode(&mode));
        METAD_CUDA(e1);
        METAD_CUDA(e2);
        METAD_CUDA(cudaMemsetAsync(w.ticket, 0, sizeof(unsigned), stream));
    }
    *out = &w;
    return METAD_OK;
}

}  // namespace metad

using namespace metad;

// A host scalar into device memory WITHOUT a staging buffer or a synchronisation: the value travels as a kernel argument
// (copied at launch).  What the host-side CVs (box ratios, densities) use to hand their value to the device-resident step.
extern "C" int metad_set_double(double* d_dst, double value, metad_stream_t stream) {
    METAD_REQUIRE(d_dst, "metad_set_double: null destination");
    set_double_kernel<<<1, 1, 0, stream>>>(d_dst, value);
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}

extern "C" int metad_force_dot(const float* d_force_i, const float* d_force_j, unsigned N, double scale, double* d_out,
                               metad_stream_t stream) {
    METAD_REQUIRE(d_out, "metad_force_dot: null output");
    METAD_REQUIRE(N == 0 || (d_force_i && d_force_j), "metad_force_dot: null force array");
    WteScratch* w = nullptr;
    int rc = wte_scratch(stream, &w);
    if (rc) return rc;
    long b = ((long)N + kWteThreads * 8L - 1) / (kWteThreads * 8L);
    if (b < 1) b = 1;
    if (b > w->blocks) b = w->blocks;
    force_dot_kernel<<<(int)b, kWteThreads, 0, stream>>>((const float4*)d_force_i, (const float4*)d_force_j, N, scale, w->partials, w->ticket, d_out);
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}

extern "C" int metad_wte_reduce(const float* d_net_force, unsigned N, double external_energy, double* d_pe,
                                metad_stream_t stream) {
    METAD_REQUIRE(d_pe, "metad_wte_reduce: null output");
    METAD_REQUIRE(N == 0 || d_net_force, "metad_wte_reduce: null net force array");
    WteScratch* w = nullptr;
    int rc = wte_scratch(stream, &w);
    if (rc) return rc;
    long b = ((long)N + kWteThreads * 8L - 1) / (kWteThreads * 8L);
    if (b < 1) b = 1;
    if (b > w->blocks) b = w->blocks;
    wte_reduce_kernel<<<(int)b, kWteThreads, 0, stream>>>((const float4*)d_net_force, N, external_energy, w->partials,
                                                         w->ticket, d_pe);
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}

extern "C" int metad_wte_scale(float* d_net_force, float* d_net_torque, float* d_net_virial, unsigned pitch, unsigned N,
                               const double* d_bias, metad_stream_t stream) {
    METAD_REQUIRE(d_bias, "metad_wte_scale: null bias");
    if (N == 0) return METAD_OK;
    METAD_REQUIRE(d_net_force, "metad_wte_scale: null net force array");
    METAD_REQUIRE(!d_net_virial || pitch >= N, "metad_wte_scale: virial pitch smaller than N");
    long b = ((long)N + kWteThreads * 4L - 1) / (kWteThreads * 4L);
    const long cap = device_sm_count() * 8L;
    if (b > cap) b = cap;
    wte_scale_kernel<<<(int)b, kWteThreads, 0, stream>>>((float4*)d_net_force, (float4*)d_net_torque, d_net_virial, pitch, N,
                                                        d_bias);
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}

extern "C" int metad_accumulate_force(float* d_net_force, const float* d_force, unsigned N, int init, metad_stream_t stream) {
    if (N == 0) return METAD_OK;
    METAD_REQUIRE(d_net_force && d_force, "metad_accumulate_force: null array");
    long b = ((long)N + kWteThreads * 4L - 1) / (kWteThreads * 4L);
    const long cap = device_sm_count() * 8L;
    if (b > cap) b = cap;
    accumulate_force_kernel<<<(int)b, kWteThreads, 0, stream>>>((float4*)d_net_force, (const float4*)d_force, N, init);
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}

extern "C" int metad_umbrella_apply(int kind, double cv0, double kappa, double width_flat, double scale,
                                    const double* d_cv, const double* d_bias_in, double* d_bias_out,
                                    double* d_energy_out, metad_stream_t stream) {
    METAD_REQUIRE(d_cv, "metad_umbrella_apply: null cv");
    METAD_REQUIRE(kind >= 0 && kind <= 4, "cv: Invalid umbrella mode specified.");
    umbrella_kernel<<<1, 32, 0, stream>>>(kind, cv0, kappa, width_flat, scale, d_cv, d_bias_in, d_bias_out, d_energy_out);
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}
