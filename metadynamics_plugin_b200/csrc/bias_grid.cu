// bias_grid.cu -- IntegratorMetaDynamics grid bias, fully device-resident, fp64.
//
// Reference behaviour restated (CPU path = parity target; IntegratorMetaDynamics.cc):
//   updateHistogram :1092-1119, updateSigmaGrid :1122-1155, sigmaDeterminant :1296-1313, well-tempered scale
//   :373-379, updateGrid :1002-1047 (GPU twin gpu_update_grid, IntegratorMetaDynamics.cu:6-93),
//   updateReweightedEstimator :1053-1090, delta merge :415-438, biasPotentialDerivative :738-776,
//   interpolateGrid :663-736, IndexGrid.cc (first CV fastest).
//
// The reference does all of this on the host (plus one deposit kernel whose result it copies back) and needs
// the CV value on the host every step.  Here one small kernel per step consumes the CV values from device
// memory and leaves dV/ds_i in device memory for the force kernels; the grid (<= 65 536 points in the configs,
// ten arrays) is touched only on deposit steps.  One CTA: the two grid-wide sums of the reweighting estimator
// need a barrier, and G is far too small to be bandwidth-relevant (every `stride` steps, ~5 MB at 256x256).
#include "common.cuh"

#include <vector>

namespace metad {

constexpr int kMaxCV = 4;
constexpr int kGridThreads = 1024;

struct GridParams {
    int d;
    unsigned n[kMaxCV], factor[kMaxCV];
    double cv_min[kMaxCV], cv_max[kMaxCV];
    double sigma_inv[kMaxCV * kMaxCV];
    double sigma_det;
    double W, T_shift, temp;
    unsigned G;
    int well_tempered;
};

struct GridArrays {
    double *grid, *grid_delta, *reweighted, *weight, *sigma_grid, *sigma_grid_delta;
    unsigned *hist, *hist_delta, *hist_gauss, *hist_gauss_delta;
    double* scalars;   // [0] curr_bias_potential [1] curr_reweight [2] num_gaussians [3] out-of-bounds count
};

__device__ inline double grid_delta_of(const GridParams& P, int i) {
    return (P.cv_max[i] - P.cv_min[i]) / (double)(P.n[i] - 1);
}

// interpolateGrid (:663-736).  Returns 0 and counts a warning when any CV is outside [min,max).
__device__ double grid_interpolate(const GridParams& P, const double* __restrict__ arr, const double* val, double* oob) {
    unsigned lo[kMaxCV], hi[kMaxCV];
    double rel[kMaxCV];
    for (int i = 0; i < P.d; ++i) {
        const double delta = grid_delta_of(P, i);
        if (val[i] < P.cv_min[i] || val[i] >= P.cv_max[i]) { *oob += 1.0; return 0.0; }
        int lower = (int)((val[i] - P.cv_min[i]) / delta);
        int upper = lower + 1;
        if (upper >= (int)P.n[i]) { lower--; upper--; }
        const double lb = P.cv_min[i] + delta * lower;
        const double ub = P.cv_min[i] + delta * upper;
        lo[i] = lower; hi[i] = upper;
        rel[i] = (val[i] - lb) / (ub - lb);
    }
    double res = 0.0;
    for (unsigned bits = 0; bits < (1u << P.d); ++bits) {
        double term = 1.0;
        unsigned idx = 0;
        for (int i = 0; i < P.d; ++i) {
            if (bits & (1u << i)) { idx += lo[i] * P.factor[i]; term *= (1.0 - rel[i]); }
            else { idx += hi[i] * P.factor[i]; term *= rel[i]; }
        }
        term *= arr[idx];
        res += term;
    }
    return res;
}

// biasPotentialDerivative (:738-776)
__device__ double grid_derivative(const GridParams& P, const double* grid, int cv, const double* val, double* oob) {
    const double delta = grid_delta_of(P, cv);
    double v1[kMaxCV], v2[kMaxCV];
    for (int i = 0; i < P.d; ++i) { v1[i] = val[i]; v2[i] = val[i]; }
    if (val[cv] - delta < P.cv_min[cv]) {
        v2[cv] += delta;
        const double y2 = grid_interpolate(P, grid, v2, oob), y1 = grid_interpolate(P, grid, val, oob);
        return (y2 - y1) / delta;
    } else if (val[cv] + delta > P.cv_max[cv]) {
        v2[cv] -= delta;
        const double y1 = grid_interpolate(P, grid, v2, oob), y2 = grid_interpolate(P, grid, val, oob);
        return (y2 - y1) / delta;
    }
    v1[cv] -= delta; v2[cv] += delta;
    const double y1 = grid_interpolate(P, grid, v1, oob), y2 = grid_interpolate(P, grid, v2, oob);
    return (y2 - y1) / (2.0 * delta);
}

// histogram bin (:1102-1116): Scalar -> unsigned conversion, off-grid if any coordinate >= n (or <= -1)
__device__ bool grid_bin(const GridParams& P, const double* val, unsigned* idx_out) {
    bool on = true;
    unsigned idx = 0;
    for (int i = 0; i < P.d; ++i) {
        const double q = (val[i] - P.cv_min[i]) / grid_delta_of(P, i);
        // the reference converts a Scalar to unsigned here; on its platform (x86-64) -1 < q < 0 lands in bin 0 and
        // q <= -1 off the grid (pinned against the reference binary, tests/test_reference_build.py)
        if (!(q > -1.0) || q >= 4294967296.0) { on = false; continue; }
        const unsigned c = q < 0.0 ? 0u : (unsigned)q;
        if (c >= P.n[i]) on = false;
        idx += c * P.factor[i];
    }
    *idx_out = idx;
    return on;
}

__global__ void __launch_bounds__(kGridThreads)
grid_step_kernel(GridParams P, GridArrays A, const double* __restrict__ cv_in, double* __restrict__ bias_out, int deposit, int phase) {
    // phase 0: the whole step.  Multiple walkers (IntegratorMetaDynamics.cc:392-410) put an all-reduce of the four delta
    // arrays between the deposit and the merge: phase 1 = histogram + sigma grid + Gaussian into the delta arrays,
    // phase 2 = reweighted estimator + merge of the (summed) deltas + hand-off of dV/ds, V(s) and the weight.
    __shared__ double cur[kMaxCV];
    __shared__ double sh_scal, sh_avg;
    __shared__ double red[32];
    __shared__ double sh_oob;

    // every thread reads the CV values itself: the hand-off lanes below do not wait for thread 0 on steps without a deposit
    double mine[kMaxCV];
    for (int i = 0; i < P.d; ++i) mine[i] = cv_in[i];
    if (threadIdx.x == 0) {
        sh_oob = 0.0;
        for (int i = 0; i < P.d; ++i) cur[i] = mine[i];
        unsigned idx;
        const bool on = phase != 2 && grid_bin(P, cur, &idx);
        if (on) atomicAdd(A.hist_delta + idx, 1u);            // no return value: the kernel does not wait for the round trip
        if (deposit && phase != 2) {
            if (on) { A.sigma_grid_delta[idx] += P.sigma_det; A.hist_gauss_delta[idx] += 1u; }
            double scal = 1.0;
            if (P.well_tempered) {
                const double V = grid_interpolate(P, A.grid, cur, &sh_oob);
                scal = exp(-V / P.T_shift);
            }
            sh_scal = scal;
        }
    }
    if (deposit) __syncthreads();

    if (deposit && phase != 2) {
        const double scal = sh_scal;
        for (unsigned g = threadIdx.x; g < P.G; g += blockDim.x) {
            // IndexGrid::getCoordinates: first CV fastest
            unsigned rest = g, c[kMaxCV];
            for (int i = P.d - 1; i >= 0; --i) { c[i] = rest / P.factor[i]; rest -= c[i] * P.factor[i]; }
            double dd[kMaxCV];
            for (int i = 0; i < P.d; ++i) dd[i] = (P.cv_min[i] + c[i] * grid_delta_of(P, i)) - cur[i];
            double gauss_exp = 0.0;
            for (int i = 0; i < P.d; ++i)
                for (int j = 0; j < P.d; ++j) {
                    const double sij = P.sigma_inv[i * P.d + j];
                    gauss_exp += dd[i] * dd[j] * 0.5 * (sij * sij);
                }
            A.grid_delta[g] = P.W * scal * exp(-gauss_exp);
        }
    }
    if (deposit && phase != 1) {
        double avg = 0.0, norm = 0.0;
        for (unsigned g = threadIdx.x; g < P.G; g += blockDim.x) {      // phase 0: every thread re-reads what it wrote above
            const double delta = A.grid_delta[g];
            const double rew = A.reweighted[g] + (double)A.hist_delta[g];
            A.reweighted[g] = rew;
            avg += rew * delta;
            norm += rew;
        }
        const double tavg = block_sum(avg, red);
        const double tnorm = block_sum(norm, red);
        if (threadIdx.x == 0) sh_avg = tavg / tnorm;
        __syncthreads();
        const double avg_dV = sh_avg;
        for (unsigned g = threadIdx.x; g < P.G; g += blockDim.x) {
            const double delta = A.grid_delta[g];
            const double fac = exp(-(delta - avg_dV) / P.temp);
            A.reweighted[g] *= fac;
            A.weight[g] /= fac;
            A.grid[g] += delta;
            A.sigma_grid[g] += A.sigma_grid_delta[g];
            A.hist[g] += A.hist_delta[g];
            A.hist_gauss[g] += A.hist_gauss_delta[g];
            A.grid_delta[g] = 0.0;
            A.sigma_grid_delta[g] = 0.0;
            A.hist_delta[g] = 0u;
            A.hist_gauss_delta[g] = 0u;
        }
        __syncthreads();
    }

    // The 2d + 2 interpolations of the hand-off (two samples per finite difference, V(s), weight(s)) are independent:
    // one lane each, so the dependent-load latency of this one-block kernel is paid once, not 2d + 2 times (it sits on
    // the critical path of every step).  Arithmetic per interpolation is unchanged (grid_derivative above is the
    // serial statement of the same rule).
    if (threadIdx.x < 32 && phase != 1) {
        const int lane = threadIdx.x, d = P.d;
        double y = 0.0, oob = 0.0;
        if (lane < 2 * d) {
            const int cv = lane >> 1, side = lane & 1;          // side 0: lower sample, 1: upper sample
            const double delta = grid_delta_of(P, cv);
            double v[kMaxCV];
            for (int i = 0; i < d; ++i) v[i] = mine[i];
            if (mine[cv] - delta < P.cv_min[cv]) { if (side) v[cv] += delta; }              // forward: (V(s+d) - V(s))/d
            else if (mine[cv] + delta > P.cv_max[cv]) { if (!side) v[cv] -= delta; }        // backward: (V(s) - V(s-d))/d
            else v[cv] += side ? delta : -delta;                                            // central
            y = grid_interpolate(P, A.grid, v, &oob);
        } else if (lane == 2 * d) {             // the per-thread copy `mine`, like the derivative lanes: on steps without a deposit
            y = grid_interpolate(P, A.grid, mine, &oob);     // no barrier orders thread 0's stores to the shared cur[] before these reads
        } else if (lane == 2 * d + 1) {
            y = grid_interpolate(P, A.weight, mine, &oob);
        }
        const double y_up = __shfl_down_sync(0xffffffffu, y, 1);
        double oob_sum = oob;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) oob_sum += __shfl_xor_sync(0xffffffffu, oob_sum, o);
        if (lane < 2 * d && !(lane & 1)) {
            const int cv = lane >> 1;
            const double delta = grid_delta_of(P, cv);
            const bool one_sided = (mine[cv] - delta < P.cv_min[cv]) || (mine[cv] + delta > P.cv_max[cv]);
            bias_out[cv] = (y_up - y) / (one_sided ? delta : 2.0 * delta);
        }
        if (lane == 2 * d) A.scalars[0] = y;
        if (lane == 2 * d + 1) A.scalars[1] = y;
        if (lane == 0) {
            if (deposit) A.scalars[2] += 1.0;
            A.scalars[3] += sh_oob + oob_sum;                   // lane 0 is thread 0: sh_oob is its own write
        }
    }
}

__global__ void grid_fill_kernel(double* a, unsigned n, double v) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a[i] = v;
}

}  // namespace metad

using namespace metad;

struct metad_grid {
    GridParams P;
    GridArrays A;
    unsigned stride = 1;
    bool add_bias = true;
    void* slab = nullptr;
};

namespace {
double det_small(const double* m_in, int d) {
    double m[kMaxCV * kMaxCV];
    for (int i = 0; i < d * d; ++i) m[i] = m_in[i];
    double det = 1.0;
    for (int c = 0; c < d; ++c) {
        int piv = c;
        for (int r = c + 1; r < d; ++r) if (fabs(m[r * d + c]) > fabs(m[piv * d + c])) piv = r;
        if (m[piv * d + c] == 0.0) return 0.0;
        if (piv != c) { for (int k = 0; k < d; ++k) std::swap(m[piv * d + k], m[c * d + k]); det = -det; }
        det *= m[c * d + c];
        for (int r = c + 1; r < d; ++r) {
            const double f = m[r * d + c] / m[c * d + c];
            for (int k = c; k < d; ++k) m[r * d + k] -= f * m[c * d + k];
        }
    }
    return det;
}
}  // namespace

extern "C" int metad_grid_create(metad_grid** out, int n_cv, const double* cv_min, const double* cv_max,
                                 const unsigned* num_points, const double* sigma, double W, double T_shift, double T,
                                 unsigned stride, int add_bias, int well_tempered) {
    METAD_REQUIRE(out && cv_min && cv_max && num_points && sigma, "metad_grid_create: null argument");
    METAD_REQUIRE(n_cv >= 1, "metad_grid_create: need at least one collective variable");
    if (n_cv > kMaxCV) { set_error("metad_grid_create: more than 4 collective variables on one grid is not supported"); return METAD_ERR_UNSUPPORTED; }
    METAD_REQUIRE(stride >= 1, "metad_grid_create: stride must be >= 1");
    auto* g = new metad_grid();
    GridParams& P = g->P;
    memset(&P, 0, sizeof P);
    P.d = n_cv;
    unsigned long long G = 1;
    for (int i = 0; i < n_cv; ++i) {
        // setGrid checks (IntegratorMetaDynamics.cc:798-812)
        if (!(cv_min[i] < cv_max[i])) { delete g; set_error("integrate.mode_metadynamics: Maximum grid value of collective variable has to be greater than minimum value."); return METAD_ERR_INVALID; }
        if (num_points[i] < 2) { delete g; set_error("integrate.mode_metadynamics: Number of grid points for collective variable has to be at least two."); return METAD_ERR_INVALID; }
        if (!(sigma[i] > 0.0)) { delete g; set_error("metad_grid_create: sigma must be positive"); return METAD_ERR_INVALID; }
        P.n[i] = num_points[i];
        P.factor[i] = (i == 0) ? 1u : P.n[i - 1] * P.factor[i - 1];
        P.cv_min[i] = cv_min[i];
        P.cv_max[i] = cv_max[i];
        P.sigma_inv[i * n_cv + i] = 1.0 / sigma[i];
        G *= num_points[i];
    }
    if (G > (1ull << 28)) { delete g; set_error("metad_grid_create: grid too large"); return METAD_ERR_INVALID; }
    P.G = (unsigned)G;
    P.sigma_det = det_small(P.sigma_inv, n_cv);
    P.W = W; P.T_shift = T_shift; P.temp = T;
    P.well_tempered = well_tempered != 0;
    g->stride = stride;
    g->add_bias = add_bias != 0;

    const size_t bytes = (size_t)P.G * (6 * sizeof(double) + 4 * sizeof(unsigned)) + 4 * sizeof(double);
    cudaError_t e = cudaMalloc(&g->slab, bytes);
    if (e != cudaSuccess) { delete g; return cuda_fail(e, "cudaMalloc(grid)", __FILE__, __LINE__); }
    e = cudaMemset(g->slab, 0, bytes);
    if (e != cudaSuccess) { cudaFree(g->slab); delete g; return cuda_fail(e, "cudaMemset(grid)", __FILE__, __LINE__); }
    double* dp = (double*)g->slab;
    GridArrays& A = g->A;
    A.scalars = dp; dp += 4;
    A.grid = dp; dp += P.G;
    A.grid_delta = dp; dp += P.G;
    A.reweighted = dp; dp += P.G;
    A.weight = dp; dp += P.G;
    A.sigma_grid = dp; dp += P.G;
    A.sigma_grid_delta = dp; dp += P.G;
    unsigned* up = (unsigned*)dp;
    A.hist = up; up += P.G;
    A.hist_delta = up; up += P.G;
    A.hist_gauss = up; up += P.G;
    A.hist_gauss_delta = up; up += P.G;
    // grid_weight starts at one (setupGrid :654-658); curr_reweight starts at one (ctor :56)
    grid_fill_kernel<<<64, 256>>>(A.weight, P.G, 1.0);
    grid_fill_kernel<<<1, 32>>>(A.scalars + 1, 1, 1.0);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(g->slab); delete g; return cuda_fail(e, "grid init", __FILE__, __LINE__); }
    *out = g;
    return METAD_OK;
}

extern "C" int metad_grid_destroy(metad_grid* g) {
    if (!g) return METAD_OK;
    cudaFree(g->slab);
    delete g;
    return METAD_OK;
}

extern "C" unsigned metad_grid_num_elements(const metad_grid* g) { return g ? g->P.G : 0; }

namespace {
int grid_step_phase(metad_grid* g, unsigned timestep, const double* d_cv_values, double* d_bias_out, int phase, cudaStream_t stream) {
    const int deposit = (g->add_bias && (timestep % g->stride == 0)) ? 1 : 0;
    grid_step_kernel<<<1, deposit ? kGridThreads : 32, 0, stream>>>(g->P, g->A, d_cv_values, d_bias_out, deposit, phase);
    METAD_LAUNCH_CHECK();
    return METAD_OK;
}
}  // namespace

extern "C" int metad_grid_step(metad_grid* g, unsigned timestep, const double* d_cv_values, double* d_bias_out,
                               metad_stream_t stream) {
    METAD_REQUIRE(g && d_cv_values && d_bias_out, "metad_grid_step: null argument");
    return grid_step_phase(g, timestep, d_cv_values, d_bias_out, 0, stream);
}

extern "C" int metad_grid_step_deposit(metad_grid* g, unsigned timestep, const double* d_cv_values, metad_stream_t stream) {
    METAD_REQUIRE(g && d_cv_values, "metad_grid_step_deposit: null argument");
    return grid_step_phase(g, timestep, d_cv_values, nullptr, 1, stream);
}

extern "C" int metad_grid_step_merge(metad_grid* g, unsigned timestep, const double* d_cv_values, double* d_bias_out,
                                     metad_stream_t stream) {
    METAD_REQUIRE(g && d_cv_values && d_bias_out, "metad_grid_step_merge: null argument");
    return grid_step_phase(g, timestep, d_cv_values, d_bias_out, 2, stream);
}

extern "C" int metad_grid_is_deposit_step(const metad_grid* g, unsigned timestep) {
    return (g && g->add_bias && (timestep % g->stride == 0)) ? 1 : 0;
}

// the four delta arrays as two contiguous device buffers (what a multiple-walker all-reduce sums): doubles
// [grid_delta | sigma_grid_delta], unsigned [hist_delta | hist_gauss_delta], G entries each
extern "C" int metad_grid_deltas_export(metad_grid* g, double* d_out_2G, unsigned* d_out_u_2G, metad_stream_t stream) {
    METAD_REQUIRE(g && d_out_2G && d_out_u_2G, "metad_grid_deltas_export: null argument");
    const size_t G = g->P.G;
    METAD_CUDA(cudaMemcpyAsync(d_out_2G, g->A.grid_delta, G * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    METAD_CUDA(cudaMemcpyAsync(d_out_2G + G, g->A.sigma_grid_delta, G * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    METAD_CUDA(cudaMemcpyAsync(d_out_u_2G, g->A.hist_delta, G * sizeof(unsigned), cudaMemcpyDeviceToDevice, stream));
    METAD_CUDA(cudaMemcpyAsync(d_out_u_2G + G, g->A.hist_gauss_delta, G * sizeof(unsigned), cudaMemcpyDeviceToDevice, stream));
    return METAD_OK;
}

extern "C" int metad_grid_deltas_import(metad_grid* g, const double* d_in_2G, const unsigned* d_in_u_2G, metad_stream_t stream) {
    METAD_REQUIRE(g && d_in_2G && d_in_u_2G, "metad_grid_deltas_import: null argument");
    const size_t G = g->P.G;
    METAD_CUDA(cudaMemcpyAsync(g->A.grid_delta, d_in_2G, G * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    METAD_CUDA(cudaMemcpyAsync(g->A.sigma_grid_delta, d_in_2G + G, G * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    METAD_CUDA(cudaMemcpyAsync(g->A.hist_delta, d_in_u_2G, G * sizeof(unsigned), cudaMemcpyDeviceToDevice, stream));
    METAD_CUDA(cudaMemcpyAsync(g->A.hist_gauss_delta, d_in_u_2G + G, G * sizeof(unsigned), cudaMemcpyDeviceToDevice, stream));
    return METAD_OK;
}

// adaptive Gaussians: install the inverse sigma matrix computed by computeSigma (IntegratorMetaDynamics.cc:1205-1294);
// the sigma grid then receives its determinant (sigmaDeterminant :1296-1313)
extern "C" int metad_grid_set_sigma_inv(metad_grid* g, const double* sigma_inv) {
    METAD_REQUIRE(g && sigma_inv, "metad_grid_set_sigma_inv: null argument");
    const int d = g->P.d;
    for (int i = 0; i < d * d; ++i) g->P.sigma_inv[i] = sigma_inv[i];
    g->P.sigma_det = det_small(g->P.sigma_inv, d);
    return METAD_OK;
}

extern "C" int metad_grid_set_flags(metad_grid* g, int add_bias, int well_tempered, unsigned stride) {
    METAD_REQUIRE(g, "metad_grid_set_flags: null grid");
    METAD_REQUIRE(stride >= 1, "metad_grid_set_flags: stride must be >= 1");
    g->add_bias = add_bias != 0;
    g->P.well_tempered = well_tempered != 0;
    g->stride = stride;
    return METAD_OK;
}

extern "C" int metad_grid_reset_histogram(metad_grid* g, metad_stream_t stream) {
    METAD_REQUIRE(g, "metad_grid_reset_histogram: null grid");
    METAD_CUDA(cudaMemsetAsync(g->A.hist, 0, sizeof(unsigned) * g->P.G, stream));
    METAD_CUDA(cudaMemsetAsync(g->A.hist_delta, 0, sizeof(unsigned) * g->P.G, stream));
    return METAD_OK;
}

namespace {
int grid_array(metad_grid* g, int which, void** ptr, size_t* bytes) {
    const size_t G = g->P.G;
    switch (which) {
        case 0: *ptr = g->A.grid; *bytes = G * sizeof(double); break;
        case 1: *ptr = g->A.reweighted; *bytes = G * sizeof(double); break;
        case 2: *ptr = g->A.weight; *bytes = G * sizeof(double); break;
        case 3: *ptr = g->A.sigma_grid; *bytes = G * sizeof(double); break;
        case 4: *ptr = g->A.hist; *bytes = G * sizeof(unsigned); break;
        case 5: *ptr = g->A.hist_gauss; *bytes = G * sizeof(unsigned); break;
        case 6: *ptr = g->A.hist_delta; *bytes = G * sizeof(unsigned); break;
        default: set_error("metad_grid: unknown array id"); return METAD_ERR_INVALID;
    }
    return METAD_OK;
}
}  // namespace

extern "C" int metad_grid_download(metad_grid* g, int which, void* h_out) {
    METAD_REQUIRE(g && h_out, "metad_grid_download: null argument");
    void* p; size_t b;
    int rc = grid_array(g, which, &p, &b);
    if (rc) return rc;
    METAD_CUDA(cudaDeviceSynchronize());
    METAD_CUDA(cudaMemcpy(h_out, p, b, cudaMemcpyDeviceToHost));
    return METAD_OK;
}

extern "C" int metad_grid_upload(metad_grid* g, int which, const void* h_in) {
    METAD_REQUIRE(g && h_in, "metad_grid_upload: null argument");
    void* p; size_t b;
    int rc = grid_array(g, which, &p, &b);
    if (rc) return rc;
    METAD_CUDA(cudaDeviceSynchronize());
    METAD_CUDA(cudaMemcpy(p, h_in, b, cudaMemcpyHostToDevice));
    return METAD_OK;
}

extern "C" int metad_grid_scalars(metad_grid* g, double* h_out4) {
    METAD_REQUIRE(g && h_out4, "metad_grid_scalars: null argument");
    METAD_CUDA(cudaDeviceSynchronize());
    METAD_CUDA(cudaMemcpy(h_out4, g->A.scalars, 4 * sizeof(double), cudaMemcpyDeviceToHost));
    return METAD_OK;
}

extern "C" int metad_grid_set_num_gaussians(metad_grid* g, unsigned n) {
    METAD_REQUIRE(g, "metad_grid_set_num_gaussians: null grid");
    const double v = (double)n;
    METAD_CUDA(cudaDeviceSynchronize());
    METAD_CUDA(cudaMemcpy(g->A.scalars + 2, &v, sizeof(double), cudaMemcpyHostToDevice));
    return METAD_OK;
}
