// hoomd_shim.h -- the few HOOMD-blue 2.x types the CV + bias-force path touches, as a
// stand-in so the host classes can be built and driven without HOOMD (which is neither installed nor vendored).
// With a real HOOMD the same host classes bind to hoomd::ParticleData / ForceCompute instead.
//
// Everything particle-sized lives in device memory (cudaMalloc); host copies happen only through the explicit
// upload / download calls, as with HOOMD's ArrayHandle(access_location::host).
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/metad_b200.h"

namespace shim {

typedef float Scalar;            // SINGLE_PRECISION build: matches the device arrays of the C ABI
struct Scalar3 { Scalar x, y, z; };
struct Scalar4 { Scalar x, y, z, w; };
struct int3_ { int x, y, z; };

inline void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}
// translate a C-ABI status into the reference's convention: msg->error() then throw std::runtime_error
inline void metad_check(int rc, const char* what) {
    if (rc != METAD_OK) throw std::runtime_error(std::string(what) + ": " + metad_last_error());
}

template <class T> class DeviceArray {
  public:
    DeviceArray() = default;
    explicit DeviceArray(size_t n) { resize(n); }
    ~DeviceArray() { if (m_ptr) cudaFree(m_ptr); }
    DeviceArray(const DeviceArray&) = delete;
    DeviceArray& operator=(const DeviceArray&) = delete;
    void resize(size_t n) {
        if (n == m_n) return;
        if (m_ptr) { cudaFree(m_ptr); m_ptr = nullptr; }
        m_n = n;
        if (n) { cuda_check(cudaMalloc(&m_ptr, n * sizeof(T)), "cudaMalloc"); cuda_check(cudaMemset(m_ptr, 0, n * sizeof(T)), "cudaMemset"); }
    }
    T* data() const { return m_ptr; }
    size_t size() const { return m_n; }
    void upload(const T* h, size_t n) { cuda_check(cudaMemcpy(m_ptr, h, n * sizeof(T), cudaMemcpyHostToDevice), "upload"); }
    void download(T* h, size_t n) const { cuda_check(cudaMemcpy(h, m_ptr, n * sizeof(T), cudaMemcpyDeviceToHost), "download"); }
    void zero() { if (m_n) cuda_check(cudaMemset(m_ptr, 0, m_n * sizeof(T)), "cudaMemset"); }

  private:
    T* m_ptr = nullptr;
    size_t m_n = 0;
};

// hoomd/BoxDim.h (orthorhombic + tilt factors); lo = -L/2
class BoxDim {
  public:
    BoxDim() : BoxDim(1, 1, 1) {}
    BoxDim(double Lx, double Ly, double Lz, double xy = 0, double xz = 0, double yz = 0) : m_xy(xy), m_xz(xz), m_yz(yz) {
        m_L[0] = Lx; m_L[1] = Ly; m_L[2] = Lz;
    }
    Scalar3 getL() const { return {(Scalar)m_L[0], (Scalar)m_L[1], (Scalar)m_L[2]}; }
    Scalar3 getLo() const { return {(Scalar)(-m_L[0] / 2), (Scalar)(-m_L[1] / 2), (Scalar)(-m_L[2] / 2)}; }
    double getTiltFactorXY() const { return m_xy; }
    double getTiltFactorXZ() const { return m_xz; }
    double getTiltFactorYZ() const { return m_yz; }
    double getVolume() const { return m_L[0] * m_L[1] * m_L[2]; }
    double L(int i) const { return m_L[i]; }
    metad_box pod() const { return metad_box{{m_L[0], m_L[1], m_L[2]}, {m_xy, m_xz, m_yz}}; }

  private:
    double m_L[3];
    double m_xy, m_xz, m_yz;
};

// hoomd Messenger: counts what the reference would print
struct Messenger {
    int notice_level = 2;
    unsigned n_warnings = 0, n_errors = 0;
    std::string last_warning, last_error;
    void warning(const std::string& s) { ++n_warnings; last_warning = s; if (notice_level >= 1) std::cerr << "*Warning*: " << s << std::endl; }
    void error(const std::string& s) { ++n_errors; last_error = s; std::cerr << "**ERROR**: " << s << std::endl; }
    void notice(int level, const std::string& s) { if (level <= notice_level) std::cout << s << std::endl; }
};

struct ExecutionConfiguration {
    std::shared_ptr<Messenger> msg = std::make_shared<Messenger>();
    cudaStream_t stream = nullptr;           // HOOMD launches on the default stream
    bool isCUDAEnabled() const { return true; }
    unsigned getRank() const { return 0; }
    unsigned getNRanks() const { return 1; }
};

// hoomd/ParticleData.h, the members the plugin reads
class ParticleData {
  public:
    ParticleData(unsigned N, const BoxDim& box, const std::vector<std::string>& type_names, std::shared_ptr<ExecutionConfiguration> exec)
        : m_N(N), m_N_global(N), m_box(box), m_global_box(box), m_types(type_names), m_exec(exec) {
        m_pos.resize(N); m_net_force.resize(N); m_net_torque.resize(N);
        m_pitch = (N + 15) / 16 * 16;
        m_net_virial.resize(6 * (size_t)m_pitch);
        for (auto& v : m_external_virial) v = 0;
    }
    unsigned getN() const { return m_N; }
    unsigned getNGlobal() const { return m_N_global; }
    void setNGlobal(unsigned n) { m_N_global = n; }       // particle-sharded runs: N local < N global
    unsigned getNTypes() const { return (unsigned)m_types.size(); }
    std::string getNameByType(unsigned i) const { return m_types.at(i); }
    const BoxDim& getBox() const { return m_box; }
    const BoxDim& getGlobalBox() const { return m_global_box; }
    void setGlobalBox(const BoxDim& b) { m_box = b; m_global_box = b; }
    DeviceArray<Scalar4>& getPositions() { return m_pos; }
    DeviceArray<Scalar4>& getNetForce() { return m_net_force; }
    DeviceArray<Scalar4>& getNetTorqueArray() { return m_net_torque; }
    DeviceArray<Scalar>& getNetVirial() { return m_net_virial; }
    unsigned getNetVirialPitch() const { return m_pitch; }
    Scalar getExternalEnergy() const { return m_external_energy; }
    void setExternalEnergy(Scalar e) { m_external_energy = e; }
    Scalar getExternalVirial(unsigned i) const { return m_external_virial[i]; }
    void setExternalVirial(unsigned i, Scalar v) { m_external_virial[i] = v; }
    // PDataFlags pressure_tensor / isotropic_virial (set by an NPT integrator or a pressure log in HOOMD): force computes
    // evaluate their virial only when one of them is set (e.g. OrderParameterMesh.cc:1062-1067)
    bool getPressureFlag() const { return m_pressure_flag; }
    void setPressureFlag(bool f) { m_pressure_flag = f; }
    std::shared_ptr<ExecutionConfiguration> getExecConf() const { return m_exec; }

  private:
    unsigned m_N, m_N_global, m_pitch;
    BoxDim m_box, m_global_box;
    std::vector<std::string> m_types;
    std::shared_ptr<ExecutionConfiguration> m_exec;
    DeviceArray<Scalar4> m_pos, m_net_force, m_net_torque;
    DeviceArray<Scalar> m_net_virial;
    Scalar m_external_energy = 0;
    bool m_pressure_flag = false;
    Scalar m_external_virial[6];
};

class SystemDefinition {
  public:
    SystemDefinition(unsigned N, const BoxDim& box, const std::vector<std::string>& type_names)
        : m_exec(std::make_shared<ExecutionConfiguration>()), m_pdata(std::make_shared<ParticleData>(N, box, type_names, m_exec)) {}
    std::shared_ptr<ParticleData> getParticleData() const { return m_pdata; }
    std::shared_ptr<ExecutionConfiguration> getExecConf() const { return m_exec; }
    unsigned getNDimensions() const { return 3; }

  private:
    std::shared_ptr<ExecutionConfiguration> m_exec;
    std::shared_ptr<ParticleData> m_pdata;
};

// hoomd/ForceCompute.h: per-particle force array + external virial; compute(timestep) runs computeForces once per step
class ForceCompute {
  public:
    explicit ForceCompute(std::shared_ptr<SystemDefinition> sysdef)
        : m_sysdef(sysdef), m_pdata(sysdef->getParticleData()), m_exec_conf(sysdef->getExecConf()) {
        m_force.resize(m_pdata->getN());
        for (auto& v : m_external_virial) v = 0;
    }
    virtual ~ForceCompute() {}
    void compute(unsigned timestep) {
        if (m_computed_once && m_last_computed == timestep) return;
        computeForces(timestep);
        m_last_computed = timestep;
        m_computed_once = true;
    }
    DeviceArray<Scalar4>& getForceArray() { return m_force; }
    // per-particle torque and virial (6 x pitch) of a force compute: allocated when first asked for (the CVs of this plugin
    // never fill them; CollectiveWrapper scales the ones of the force it wraps)
    DeviceArray<Scalar4>& getTorqueArray() { if (m_torque.size() != m_force.size()) m_torque.resize(m_force.size()); return m_torque; }
    DeviceArray<Scalar>& getVirialArray() { if (m_virial.size() != 6 * getVirialPitch()) m_virial.resize(6 * getVirialPitch()); return m_virial; }
    size_t getVirialPitch() const { return (m_force.size() + 15) / 16 * 16; }
    Scalar getExternalEnergy() const { return m_external_energy; }
    void setExternalEnergy(Scalar e) { m_external_energy = e; }
    // force computes whose external virial needs a device -> host round trip evaluate it when it is read (HOOMD reads it only
    // when the pressure is needed), not in every step
    Scalar getExternalVirial(unsigned i) { updateExternalVirial(); return m_external_virial[i]; }
    virtual void updateExternalVirial() {}
    virtual std::vector<std::string> getProvidedLogQuantities() { return {}; }
    virtual Scalar getLogValue(const std::string& quantity, unsigned) { throw std::runtime_error("Error querying log quantity " + quantity); }
    bool enabled = true;             // cv.potential_energy disables itself as a regular ForceCompute (cv.py:490)

  protected:
    virtual void computeForces(unsigned timestep) = 0;
    std::shared_ptr<SystemDefinition> m_sysdef;
    std::shared_ptr<ParticleData> m_pdata;
    std::shared_ptr<ExecutionConfiguration> m_exec_conf;
    DeviceArray<Scalar4> m_force, m_torque;
    DeviceArray<Scalar> m_virial;
    Scalar m_external_energy = 0;
    Scalar m_external_virial[6];
    unsigned m_last_computed = 0;
    bool m_computed_once = false;
};

// Stand-in for "any HOOMD force": a ForceCompute whose per-particle force / energy, torque and virial are prescribed
// arrays, re-installed at every new timestep (what a pair or bond force would recompute).  cv.wrap needs something to wrap.
class PrescribedForce : public ForceCompute {
  public:
    explicit PrescribedForce(std::shared_ptr<SystemDefinition> sysdef) : ForceCompute(sysdef) {}
    void setArrays(const std::vector<Scalar4>& force, const std::vector<Scalar4>& torque, const std::vector<Scalar>& virial6pitch, Scalar external_energy) {
        if (force.size() != m_force.size() || torque.size() != m_force.size() || virial6pitch.size() != 6 * getVirialPitch())
            throw std::runtime_error("PrescribedForce: array sizes do not match the particle number");
        m_h_force = force; m_h_torque = torque; m_h_virial = virial6pitch; m_external_energy = external_energy;
    }

  protected:
    void computeForces(unsigned) override {
        if (m_h_force.empty()) return;
        m_force.upload(m_h_force.data(), m_h_force.size());
        getTorqueArray().upload(m_h_torque.data(), m_h_torque.size());
        getVirialArray().upload(m_h_virial.data(), m_h_virial.size());
    }
    std::vector<Scalar4> m_h_force, m_h_torque;
    std::vector<Scalar> m_h_virial;
};

}  // namespace shim
