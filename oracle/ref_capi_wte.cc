// ref_capi_wte.cc -- C entry point into the REFERENCE's own WellTemperedEnsemble (CPU branch), compiled from the reference's
// source where it lies.  Its header names the GPU autotuners outside ENABLE_CUDA guards, so this translation unit and
// WellTemperedEnsemble.cc are compiled with ENABLE_CUDA defined against inert stand-ins (oracle/ref_shim_cuda/); the CPU
// branch is the one that runs.  TEST INFRASTRUCTURE ONLY -- never loaded by the product path.
#include <hoomd/ForceCompute.h>
#include <hoomd/extern/pybind/include/pybind11/pybind11.h>
#include <cstdio>
#define private public
#define protected public
#include "WellTemperedEnsemble.h"
#undef private
#undef protected

// The GPU drivers the reference's GPU branch links against (defined in its .cu file, which is not compiled here): inert, never
// called because ExecutionConfiguration::exec_mode is CPU.
#include "WellTemperedEnsemble.cuh"
void gpu_scale_netforce(Scalar4*, Scalar4*, Scalar*, unsigned int, Scalar, const GPUPartition&, const unsigned int, const unsigned int) {}
void gpu_reduce_potential_energy(Scalar*, Scalar4*, const GPUPartition&, const unsigned int, const unsigned int, Scalar*, bool,
                                 const unsigned int) {}

// net_force / net_torque: N x 4, net_virial: 6 rows of `pitch` values (pitch = the stand-in's GPUArray pitch, returned);
// on return the arrays hold the values scaled by computeBiasForces and *pe the CV.
extern "C" int ref_wte(unsigned N, double* net_force, double* net_torque, double* net_virial6N, double external_energy,
                       double* external_virial6, double bias, double* pe) {
    try {
        BoxDim box(Scalar(10), Scalar(10), Scalar(10));
        std::shared_ptr<ExecutionConfiguration> exec(new ExecutionConfiguration());
        std::shared_ptr<ParticleData> pdata(new ParticleData(N, box, 1, exec));
        std::shared_ptr<SystemDefinition> sys(new SystemDefinition(pdata));
        const unsigned pitch = pdata->getNetVirial().getPitch();
        {
            ArrayHandle<Scalar4> f(pdata->getNetForce(), access_location::host, access_mode::overwrite);
            ArrayHandle<Scalar4> t(pdata->getNetTorqueArray(), access_location::host, access_mode::overwrite);
            ArrayHandle<Scalar> v(pdata->getNetVirial(), access_location::host, access_mode::overwrite);
            for (unsigned i = 0; i < N; ++i) {
                f.data[i] = make_scalar4((Scalar)net_force[4 * i], (Scalar)net_force[4 * i + 1], (Scalar)net_force[4 * i + 2], (Scalar)net_force[4 * i + 3]);
                t.data[i] = make_scalar4((Scalar)net_torque[4 * i], (Scalar)net_torque[4 * i + 1], (Scalar)net_torque[4 * i + 2], (Scalar)net_torque[4 * i + 3]);
                for (unsigned r = 0; r < 6; ++r) v.data[i + r * pitch] = (Scalar)net_virial6N[(size_t)r * N + i];
            }
        }
        pdata->setExternalEnergy((Scalar)external_energy);
        for (unsigned r = 0; r < 6; ++r) pdata->setExternalVirial(r, (Scalar)external_virial6[r]);
        WellTemperedEnsemble wte(sys, "wte");
        *pe = wte.getCurrentValue(1);
        wte.setBiasFactor((Scalar)bias);
        wte.computeBiasForces(1);
        {
            ArrayHandle<Scalar4> f(pdata->getNetForce(), access_location::host, access_mode::read);
            ArrayHandle<Scalar4> t(pdata->getNetTorqueArray(), access_location::host, access_mode::read);
            ArrayHandle<Scalar> v(pdata->getNetVirial(), access_location::host, access_mode::read);
            for (unsigned i = 0; i < N; ++i) {
                net_force[4 * i] = f.data[i].x; net_force[4 * i + 1] = f.data[i].y; net_force[4 * i + 2] = f.data[i].z; net_force[4 * i + 3] = f.data[i].w;
                net_torque[4 * i] = t.data[i].x; net_torque[4 * i + 1] = t.data[i].y; net_torque[4 * i + 2] = t.data[i].z; net_torque[4 * i + 3] = t.data[i].w;
                for (unsigned r = 0; r < 6; ++r) net_virial6N[(size_t)r * N + i] = v.data[i + r * pitch];
            }
        }
        for (unsigned r = 0; r < 6; ++r) external_virial6[r] = pdata->getExternalVirial(r);
        return 0;
    } catch (const std::exception& e) { fprintf(stderr, "ref_wte: %s\n", e.what()); return -1; }
}
