// ref_capi.cc -- C entry points into the REFERENCE's own CPU classes, compiled from the reference's sources where they lie
// (oracle/Makefile, target `_ref`; the HOOMD types they need come from the stand-in under oracle/ref_shim/).
// Used by tests/test_reference_build.py and tests/golden/make_ref_golden.py to check the oracle's restatement
// against the reference's own arithmetic.  TEST INFRASTRUCTURE ONLY -- never loaded by the product path.
//
// The access-specifier override below only affects THIS translation unit (the reference's .cc files are compiled
// unmodified); it lets the checker read the meshes and call the protected computeBiasForces directly.
#include <hoomd/ForceCompute.h>       // the stand-in and every standard header first, with their real access specifiers
#include <hoomd/extern/kiss_fftnd.h>
#include <hoomd/extern/pybind/include/pybind11/pybind11.h>
#include <hoomd/md/IntegratorTwoStep.h>
#include <chrono>
#include <cstdio>
#include <memory>
#include <string.h>
#define private public
#define protected public
#include "OrderParameterMesh.h"
#include "LamellarOrderParameter.h"
#include "AspectRatio.h"
#include "Density.h"
#include "IndexGrid.h"
#include "IntegratorMetaDynamics.h"
#undef private
#undef protected

namespace {
std::shared_ptr<SystemDefinition> make_system(const float* postype, unsigned N, const double* L, const double* tilt, unsigned ntypes) {
    BoxDim box((Scalar)L[0], (Scalar)L[1], (Scalar)L[2]);
    box.setTiltFactors((Scalar)tilt[0], (Scalar)tilt[1], (Scalar)tilt[2]);
    std::shared_ptr<ExecutionConfiguration> exec(new ExecutionConfiguration());
    std::shared_ptr<ParticleData> pdata(new ParticleData(N, box, ntypes, exec));
    ArrayHandle<Scalar4> h_pos(pdata->getPositions(), access_location::host, access_mode::overwrite);
    for (unsigned i = 0; i < N; ++i) {
        int type;
        memcpy(&type, postype + 4 * (size_t)i + 3, 4);
        h_pos.data[i] = make_scalar4((Scalar)postype[4 * (size_t)i], (Scalar)postype[4 * (size_t)i + 1], (Scalar)postype[4 * (size_t)i + 2],
                                     __int_as_scalar(type));
    }
    return std::shared_ptr<SystemDefinition>(new SystemDefinition(pdata));
}
void copy_force(ForceCompute& fc, unsigned N, double* out) {
    ArrayHandle<Scalar4> h(fc.getForceArray(), access_location::host, access_mode::read);
    for (unsigned i = 0; i < N; ++i) { out[4 * (size_t)i] = h.data[i].x; out[4 * (size_t)i + 1] = h.data[i].y; out[4 * (size_t)i + 2] = h.data[i].z; out[4 * (size_t)i + 3] = h.data[i].w; }
}
}  // namespace

namespace {
class PrescribedCV : public CollectiveVariable {
  public:
    PrescribedCV(std::shared_ptr<SystemDefinition> sysdef, const std::string& name) : CollectiveVariable(sysdef, name), m_value(0) {}
    Scalar getCurrentValue(unsigned int) { return m_value; }
    Scalar m_value;
};
// a CV with a prescribed value AND a prescribed gradient: computeBiasForces (called by computeDerivatives with bias = 1)
// leaves bias * gradient in the force array, which is what IntegratorMetaDynamics::computeSigma reads
class GradientCV : public CollectiveVariable {
  public:
    GradientCV(std::shared_ptr<SystemDefinition> sysdef, const std::string& name, const float* grad4, unsigned N, bool can)
        : CollectiveVariable(sysdef, name), m_value(0), m_grad(grad4, grad4 + 4 * (size_t)N), m_can(can) {}
    Scalar getCurrentValue(unsigned int) { return m_value; }
    virtual bool canComputeDerivatives() { return m_can; }
    virtual void computeBiasForces(unsigned int) {
        ArrayHandle<Scalar4> h(m_force, access_location::host, access_mode::overwrite);
        for (size_t n = 0; n < m_grad.size() / 4; ++n)
            h.data[n] = make_scalar4(m_bias * (Scalar)m_grad[4 * n], m_bias * (Scalar)m_grad[4 * n + 1], m_bias * (Scalar)m_grad[4 * n + 2], Scalar(0));
    }
    Scalar m_value;
    std::vector<float> m_grad;
    bool m_can;
};
template <class T, class U> void dump(const GPUArray<T>& a, U* out) {
    if (!out) return;
    ArrayHandle<T> h(a, access_location::host, access_mode::read);
    for (unsigned i = 0; i < a.getNumElements(); ++i) out[i] = (U)h.data[i];
}
}  // namespace

namespace {
// A CV object of the reference on a particle set that stays resident, as in a simulation: what bench.py times as the
// reference's own CPU implementation of the step (getCurrentValue(t) + setBiasFactor + computeBiasForces(t), new timestep
// every call so that the per-timestep cache of the CV never answers).
struct RefStepPlan {
    std::shared_ptr<SystemDefinition> sys;
    std::unique_ptr<CollectiveVariable> cv;
    unsigned t = 0;
};
}  // namespace

extern "C" {

int ref_scalar_bytes() { return (int)sizeof(Scalar); }

void* ref_step_create_mesh(unsigned nx, unsigned ny, unsigned nz, const double* mode, unsigned ntypes, const double* L, const double* tilt,
                           const float* postype, unsigned N) {
    try {
        auto* p = new RefStepPlan();
        p->sys = make_system(postype, N, L, tilt, ntypes);
        std::vector<Scalar> m(mode, mode + ntypes);
        p->cv.reset(new OrderParameterMesh(p->sys, nx, ny, nz, m));
        return p;
    } catch (const std::exception& e) { fprintf(stderr, "ref_step_create_mesh: %s\n", e.what()); return nullptr; }
}
void* ref_step_create_lamellar(const double* mode, unsigned ntypes, const int* lattice, unsigned n_wave, const double* L, const double* tilt,
                               const float* postype, unsigned N) {
    try {
        auto* p = new RefStepPlan();
        p->sys = make_system(postype, N, L, tilt, ntypes);
        std::vector<Scalar> m(mode, mode + ntypes);
        std::vector<int3> lv(n_wave);
        for (unsigned k = 0; k < n_wave; ++k) lv[k] = make_int3(lattice[3 * k], lattice[3 * k + 1], lattice[3 * k + 2]);
        p->cv.reset(new LamellarOrderParameter(p->sys, m, lv, ""));
        return p;
    } catch (const std::exception& e) { fprintf(stderr, "ref_step_create_lamellar: %s\n", e.what()); return nullptr; }
}
// one step; seconds[0] = getCurrentValue, seconds[1] = computeBiasForces (wall clock)
int ref_step(void* h, double bias, double* cv, double* seconds) {
    try {
        auto* p = (RefStepPlan*)h;
        ++p->t;
        const auto t0 = std::chrono::steady_clock::now();
        *cv = (double)p->cv->getCurrentValue(p->t);
        const auto t1 = std::chrono::steady_clock::now();
        p->cv->setBiasFactor((Scalar)bias);
        p->cv->computeBiasForces(p->t);
        const auto t2 = std::chrono::steady_clock::now();
        seconds[0] = std::chrono::duration<double>(t1 - t0).count();
        seconds[1] = std::chrono::duration<double>(t2 - t1).count();
        return 0;
    } catch (const std::exception& e) { fprintf(stderr, "ref_step: %s\n", e.what()); return -1; }
}
// force of particle i after the last step (spot checks)
int ref_step_force(void* h, unsigned i, double* out4) {
    auto* p = (RefStepPlan*)h;
    ArrayHandle<Scalar4> f(p->cv->getForceArray(), access_location::host, access_mode::read);
    out4[0] = f.data[i].x; out4[1] = f.data[i].y; out4[2] = f.data[i].z; out4[3] = f.data[i].w;
    return 0;
}
void ref_step_destroy(void* h) { delete (RefStepPlan*)h; }

// OrderParameterMesh: getCurrentValue (assignParticles + updateMeshes + computeCV), then computeBiasForces with `bias`.
// Outputs (double): cv, mode_sq, force[4N], rho[M] = Re(mesh), inv[M] = Re(inverse mesh), interp[M]
int ref_mesh(unsigned nx, unsigned ny, unsigned nz, const double* mode, unsigned ntypes, const double* L, const double* tilt,
             const float* postype, unsigned N, double bias, double* cv, double* mode_sq, double* force, double* rho, double* inv, double* interp) {
    try {
        auto sys = make_system(postype, N, L, tilt, ntypes);
        std::vector<Scalar> m(mode, mode + ntypes);
        OrderParameterMesh op(sys, nx, ny, nz, m);
        *cv = op.getCurrentValue(1);
        *mode_sq = op.m_mode_sq;
        op.setBiasFactor((Scalar)bias);
        op.computeBiasForces(1);
        copy_force(op, N, force);
        const size_t M = (size_t)nx * ny * nz;
        ArrayHandle<kiss_fft_cpx> h_mesh(op.m_mesh, access_location::host, access_mode::read);
        ArrayHandle<kiss_fft_cpx> h_inv(op.m_inv_fourier_mesh, access_location::host, access_mode::read);
        ArrayHandle<Scalar> h_interp(op.m_interpolation_f, access_location::host, access_mode::read);
        for (size_t i = 0; i < M; ++i) { rho[i] = h_mesh.data[i].r; inv[i] = h_inv.data[i].r; interp[i] = h_interp.data[i]; }
        return 0;
    } catch (const std::exception& e) { fprintf(stderr, "ref_mesh: %s\n", e.what()); return -1; }
}

// OrderParameterMesh::computeVirial with a kernel table (setTable + setUseTable): external virial for bias factor `bias`
int ref_mesh_virial(unsigned nx, unsigned ny, unsigned nz, const double* mode, unsigned ntypes, const double* L, const double* tilt,
                    const float* postype, unsigned N, const double* K, const double* dK, unsigned ntable, double kmin, double kmax,
                    double bias, double* out6) {
    try {
        auto sys = make_system(postype, N, L, tilt, ntypes);
        std::vector<Scalar> m(mode, mode + ntypes);
        OrderParameterMesh op(sys, nx, ny, nz, m);
        std::vector<Scalar> k(K, K + ntable), dk(dK, dK + ntable);
        op.setTable(k, dk, (Scalar)kmin, (Scalar)kmax);
        op.setUseTable(true);
        op.getCurrentValue(1);
        op.setBiasFactor((Scalar)bias);
        op.computeVirial();
        for (int i = 0; i < 6; ++i) out6[i] = op.getExternalVirial(i);
        return 0;
    } catch (const std::exception& e) { fprintf(stderr, "ref_mesh_virial: %s\n", e.what()); return -1; }
}

// OrderParameterMesh::computeQmax (the q*_max / sq_max log quantities): out4 = {q_max.x, q_max.y, q_max.z, sq_max}
int ref_mesh_qmax(unsigned nx, unsigned ny, unsigned nz, const double* mode, unsigned ntypes, const double* L, const double* tilt,
                  const float* postype, unsigned N, double* out4) {
    try {
        auto sys = make_system(postype, N, L, tilt, ntypes);
        std::vector<Scalar> m(mode, mode + ntypes);
        OrderParameterMesh op(sys, nx, ny, nz, m);
        op.computeQmax(1);
        out4[0] = op.m_q_max.x; out4[1] = op.m_q_max.y; out4[2] = op.m_q_max.z; out4[3] = op.m_sq_max;
        return 0;
    } catch (const std::exception& e) { fprintf(stderr, "ref_mesh_qmax: %s\n", e.what()); return -1; }
}

// LamellarOrderParameter: getCurrentValue, Fourier modes, computeBiasForces with `bias`
int ref_lamellar(const double* mode, unsigned ntypes, const int* lattice, unsigned n_wave, const double* L, const double* tilt,
                 const float* postype, unsigned N, double bias, double* cv, double* modes_out /* 2*n_wave */, double* force) {
    try {
        auto sys = make_system(postype, N, L, tilt, ntypes);
        std::vector<Scalar> m(mode, mode + ntypes);
        std::vector<int3> lv(n_wave);
        for (unsigned k = 0; k < n_wave; ++k) lv[k] = make_int3(lattice[3 * k], lattice[3 * k + 1], lattice[3 * k + 2]);
        LamellarOrderParameter op(sys, m, lv, "");
        *cv = op.getCurrentValue(1);
        {
            ArrayHandle<Scalar2> h(op.m_fourier_modes, access_location::host, access_mode::read);
            for (unsigned k = 0; k < n_wave; ++k) { modes_out[2 * k] = h.data[k].x; modes_out[2 * k + 1] = h.data[k].y; }
        }
        op.setBiasFactor((Scalar)bias);
        op.computeBiasForces(1);
        copy_force(op, N, force);
        return 0;
    } catch (const std::exception& e) { fprintf(stderr, "ref_lamellar: %s\n", e.what()); return -1; }
}

// CollectiveVariable::computeForces / getUmbrellaPotential through a Lamellar CV: returns the bias factor that reached
// computeBiasForces (read back from the force: F = bias * dCV-like term, so the ratio to a unit-bias run) and the umbrella energy
int ref_umbrella(int kind, double cv0, double kappa, double width_flat, double scale, const double* mode, unsigned ntypes,
                 const int* lattice, unsigned n_wave, const double* L, const float* postype, unsigned N, double bias_in,
                 double* cv, double* energy, double* force) {
    try {
        const double tilt[3] = {0, 0, 0};
        auto sys = make_system(postype, N, L, tilt, ntypes);
        std::vector<Scalar> m(mode, mode + ntypes);
        std::vector<int3> lv(n_wave);
        for (unsigned k = 0; k < n_wave; ++k) lv[k] = make_int3(lattice[3 * k], lattice[3 * k + 1], lattice[3 * k + 2]);
        LamellarOrderParameter op(sys, m, lv, "");
        op.setUmbrella((CollectiveVariable::umbrella_Enum)kind);
        op.setMinimum((Scalar)cv0); op.setKappa((Scalar)kappa); op.setWidthFlat((Scalar)width_flat); op.setScale((Scalar)scale);
        op.setBiasFactor((Scalar)bias_in);
        *cv = op.getCurrentValue(1);
        *energy = op.getUmbrellaPotential(1);
        op.compute(1);                       // ForceCompute::compute -> CollectiveVariable::computeForces
        copy_force(op, N, force);
        return 0;
    } catch (const std::exception& e) { fprintf(stderr, "ref_umbrella: %s\n", e.what()); return -1; }
}

int ref_aspect(unsigned dir1, unsigned dir2, const double* L, const double* tilt, double bias, double* cv, double* ext_virial6) {
    try {
        const float dummy[4] = {0, 0, 0, 0};
        auto sys = make_system(dummy, 1, L, tilt, 1);
        AspectRatio ar(sys, dir1, dir2);
        *cv = ar.getCurrentValue(1);
        ar.setBiasFactor((Scalar)bias);
        ar.computeBiasForces(1);
        for (int i = 0; i < 6; ++i) ext_virial6[i] = ar.getExternalVirial(i);
        return 0;
    } catch (const std::exception& e) { fprintf(stderr, "ref_aspect: %s\n", e.what()); return -1; }
}

// IntegratorMetaDynamics grid bias: prepRun (which performs the first updateBiasPotential) and then one
// updateBiasPotential per further step, driven with prescribed CV values through a trivial CollectiveVariable subclass.
// Outputs: the bias factors handed to the CVs after every step, and the grid state at the end.

int ref_grid_sequence(int ncv, const double* cv_min, const double* cv_max, const unsigned* num_points, const double* sigma, double W,
                      double T_shift, double T, unsigned stride, int add_bias, int well_tempered, const double* cv_values,
                      const unsigned* timesteps, int nsteps, double* bias_out, double* grid, double* reweighted, double* weight,
                      double* sigma_grid, unsigned* hist, unsigned* hist_gauss, unsigned* hist_delta, double* scalars3) {
    try {
        const double L[3] = {10, 10, 10}, tilt[3] = {0, 0, 0};
        const float dummy[4] = {0, 0, 0, 0};
        auto sys = make_system(dummy, 1, L, tilt, 1);
        IntegratorMetaDynamics imd(sys, Scalar(0.005), (Scalar)W, (Scalar)T_shift, (Scalar)T, stride, add_bias != 0, "", false,
                                   well_tempered ? IntegratorMetaDynamics::mode_well_tempered : IntegratorMetaDynamics::mode_standard);
        std::vector<std::shared_ptr<PrescribedCV> > cvs;
        for (int i = 0; i < ncv; ++i) {
            cvs.push_back(std::shared_ptr<PrescribedCV>(new PrescribedCV(sys, "cv" + std::to_string(i))));
            imd.registerCollectiveVariable(cvs.back(), (Scalar)sigma[i], (Scalar)cv_min[i], (Scalar)cv_max[i], num_points[i]);
        }
        imd.setGrid(true);
        for (int s = 0; s < nsteps; ++s) {
            for (int i = 0; i < ncv; ++i) cvs[i]->m_value = (Scalar)cv_values[(size_t)s * ncv + i];
            if (s == 0) imd.prepRun(timesteps[s]);
            else imd.updateBiasPotential(timesteps[s]);
            for (int i = 0; i < ncv; ++i) bias_out[(size_t)s * ncv + i] = cvs[i]->m_bias;
        }
        dump(imd.m_grid, grid); dump(imd.m_grid_reweighted, reweighted); dump(imd.m_grid_weight, weight); dump(imd.m_sigma_grid, sigma_grid);
        dump(imd.m_grid_hist, hist); dump(imd.m_grid_hist_gauss, hist_gauss); dump(imd.m_grid_hist_delta, hist_delta);
        scalars3[0] = imd.m_curr_bias_potential; scalars3[1] = imd.m_curr_reweight; scalars3[2] = imd.m_num_gaussians;
        return 0;
    } catch (const std::exception& e) { fprintf(stderr, "ref_grid_sequence: %s\n", e.what()); return -1; }
}

// Adaptive Gaussians (setAdaptive(true), IntegratorMetaDynamics.cc:333-341, computeSigma :1205-1294) with prescribed CV
// values and prescribed per-particle gradients: sigma_inv after every step, bias factors, final grid and sigma grid.
int ref_grid_adaptive(int ncv, const double* cv_min, const double* cv_max, const unsigned* num_points, const double* sigma, double W,
                      double T_shift, double T, unsigned stride, int well_tempered, double sigma_g, const float* grads /* [ncv][N][4] */,
                      const int* can, unsigned N, const double* cv_values, const unsigned* timesteps, int nsteps, double* bias_out,
                      double* sigma_inv_out /* [nsteps][ncv*ncv] */, double* grid, double* sigma_grid) {
    try {
        const double L[3] = {10, 10, 10}, tilt[3] = {0, 0, 0};
        std::vector<float> pt(4 * (size_t)N, 0.f);
        auto sys = make_system(pt.data(), N, L, tilt, 1);
        IntegratorMetaDynamics imd(sys, Scalar(0.005), (Scalar)W, (Scalar)T_shift, (Scalar)T, stride, true, "", false,
                                   well_tempered ? IntegratorMetaDynamics::mode_well_tempered : IntegratorMetaDynamics::mode_standard);
        std::vector<std::shared_ptr<GradientCV> > cvs;
        for (int i = 0; i < ncv; ++i) {
            cvs.push_back(std::shared_ptr<GradientCV>(new GradientCV(sys, "cv" + std::to_string(i), grads + 4 * (size_t)N * i, N, can[i] != 0)));
            imd.registerCollectiveVariable(cvs.back(), (Scalar)sigma[i], (Scalar)cv_min[i], (Scalar)cv_max[i], num_points[i]);
        }
        imd.setGrid(true);
        imd.setAdaptive(true);
        imd.setSigmaG((Scalar)sigma_g);
        for (int s = 0; s < nsteps; ++s) {
            for (int i = 0; i < ncv; ++i) cvs[i]->m_value = (Scalar)cv_values[(size_t)s * ncv + i];
            if (s == 0) imd.prepRun(timesteps[s]);
            else imd.updateBiasPotential(timesteps[s]);
            for (int i = 0; i < ncv; ++i) bias_out[(size_t)s * ncv + i] = cvs[i]->m_bias;
            ArrayHandle<Scalar> h(imd.m_sigma_inv, access_location::host, access_mode::read);
            for (int k = 0; k < ncv * ncv; ++k) sigma_inv_out[(size_t)s * ncv * ncv + k] = h.data[k];
        }
        dump(imd.m_grid, grid); dump(imd.m_sigma_grid, sigma_grid);
        return 0;
    } catch (const std::exception& e) { fprintf(stderr, "ref_grid_adaptive: %s\n", e.what()); return -1; }
}

// The reference's test/test_2d.py scenario through the reference's own classes, files and all: one particle, Density +
// AspectRatio CVs on a 20 x 30 grid, well-tempered, stride 1, grid dumped every step with the hills log switched on, the
// box rescaled between the two run(1) calls.  run(1) = prepRun(t) [first updateBiasPotential(t)] + update(t)
// [updateBiasPotential(t + 1)].  restart != 0: a fresh integrator restarts from <dir>/bias.dat_1, rescales, runs one step
// and dumps <dir>/bias_restart.dat_{0,1}.  Every file is written by IntegratorMetaDynamics itself.
int ref_test2d_files(const char* dir, int restart) {
    try {
        const double L0 = std::pow(10.0, 1.0 / 3.0), sc = std::pow(0.125, 1.0 / 3.0);
        const double L[3] = {L0, L0, L0}, tilt[3] = {0, 0, 0};
        const float one[4] = {0, 0, 0, 0};
        auto sys = make_system(one, 1, L, tilt, 1);
        const std::string d(dir);
        IntegratorMetaDynamics imd(sys, Scalar(0.005), Scalar(1.0), Scalar(1.0), Scalar(1.0), 1, true, restart ? "" : d + "/hills.dat", true,
                                   IntegratorMetaDynamics::mode_well_tempered);
        std::shared_ptr<ParticleGroup> all(new ParticleGroup(1));
        std::shared_ptr<Density> density(new Density(sys, all, ""));
        std::shared_ptr<AspectRatio> aspect(new AspectRatio(sys, 0, 1));
        imd.registerCollectiveVariable(density, Scalar(0.25), Scalar(0.0), Scalar(1.0), 20);
        imd.registerCollectiveVariable(aspect, Scalar(0.1), Scalar(0.0), Scalar(2.0), 30);
        imd.setGrid(true);
        auto rescale = [&]() { sys->getParticleData()->setGlobalBox(BoxDim((Scalar)(L0 * sc), (Scalar)(L0 * sc), (Scalar)(L0 * sc))); };
        if (!restart) {
            imd.dumpGrid(d + "/bias.dat", "", 1);
            imd.prepRun(0); imd.updateBiasPotential(1);
            rescale();
            imd.prepRun(1); imd.updateBiasPotential(2);
        } else {
            imd.restartFromGridFile(d + "/bias.dat_1");
            imd.dumpGrid(d + "/bias_restart.dat", "", 1);
            rescale();
            imd.prepRun(0); imd.updateBiasPotential(1);
        }
        return (int)imd.m_num_gaussians;
    } catch (const std::exception& e) { fprintf(stderr, "ref_test2d_files: %s\n", e.what()); return -1; }
}

unsigned ref_indexgrid_index(const unsigned* lengths, int d, const unsigned* coords) {
    IndexGrid g(std::vector<unsigned int>(lengths, lengths + d));
    return g.getIndex(std::vector<unsigned int>(coords, coords + d));
}
void ref_indexgrid_coords(const unsigned* lengths, int d, unsigned idx, unsigned* coords) {
    IndexGrid g(std::vector<unsigned int>(lengths, lengths + d));
    std::vector<unsigned int> c(d);
    g.getCoordinates(idx, c);
    for (int i = 0; i < d; ++i) coords[i] = c[i];
}
unsigned ref_indexgrid_num(const unsigned* lengths, int d) {
    IndexGrid g(std::vector<unsigned int>(lengths, lengths + d));
    return g.getNumElements();
}

}  // extern "C"
